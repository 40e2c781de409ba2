#!/usr/bin/env python
"""Host profile of image_to_graph (dynamic quadtree, 229 x 361, 10 frames): wall per call, cProfile callees."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib, graph_csr

dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(12)
img = q.add_positional_encoding(torch.from_numpy(cube[:10]).to(dev))
img1 = img[:1].contiguous()
for label, im in (("10 frames", img), ("1 frame", img1)):
    build = lambda: q.image_to_graph(im, thresh=0.15, mask=mask, transform_func=bench.dist_from_05, use_edge_attrs=True)
    for _ in range(5):
        g = build()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        g = build()
    torch.cuda.synchronize()
    print(label, "wall us per build", (time.perf_counter() - t0) / 50 * 1e6, "N", g["data"].shape[1], "E", g["edge_index"].shape[1])
    t0 = time.perf_counter()
    for _ in range(50):
        g = build()
        N = int(g["data"].shape[1])
        csr = graph_csr.get_csr(g["edge_index"], g["edge_attrs"], N)
    torch.cuda.synchronize()
    print(label, "wall us per build + CSR", (time.perf_counter() - t0) / 50 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    g = build()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("tottime").print_stats(25)
st.print_callees("image_to_graph")
