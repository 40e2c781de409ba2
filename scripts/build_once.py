#!/usr/bin/env python
"""Three dynamic-quadtree graph builds + CSR at the ice grid (for an ncu launch list of the build kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import graph_csr

dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(12)
img = q.add_positional_encoding(torch.from_numpy(cube[:10]).to(dev))
for _ in range(3):
    g = q.image_to_graph(img, thresh=0.15, mask=mask, transform_func=bench.dist_from_05, use_edge_attrs=True)
    csr = graph_csr.get_csr(g["edge_index"], g["edge_attrs"], int(g["data"].shape[1]))
torch.cuda.synchronize()
print("N", g["data"].shape[1], "E", g["edge_index"].shape[1])
