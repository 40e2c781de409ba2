#!/usr/bin/env python
"""Phase timeline of the one-launch quadtree graph build (csrc/graph_build.cu): global-timer stamps of CTA 0 after every
grid-wide barrier, at the ice grid (229 x 361, thresh 0.15, 10 frames and 1 frame)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import graph_functions as gf

dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(12)
img = q.add_positional_encoding(torch.from_numpy(cube[:10]).to(dev))
names = ["level-0 planes, table fill, NaN count", "split pyramid (tiles)", "base-cell scan (1 CTA)", "labels + leaf rectangles",
         "pixel scan pass 1 + adjacency inserts", "pix_ptr + first-occurrence flags", "pixel lists + edge emission",
         "pooling + edge_index + CSR counts", "edge attributes + CSR scan pass 1", "CSR row pointers", "CSR fill",
         "CSR row sort", "in-CSR payload", "out-CSR payload"]
noise = img[:1].clone()
noise[..., 0] = torch.rand_like(noise[..., 0])          # what an untrained model forecasts: (almost) every pixel its own leaf
for label, im in (("10 frames", img), ("1 frame", img[:1].contiguous()), ("1 frame of noise (pixel-level mesh)", noise)):
    for _ in range(3):
        g = q.image_to_graph(im, thresh=0.15, mask=mask, transform_func=bench.dist_from_05, use_edge_attrs=True)
    torch.cuda.synchronize()
    arena = next(v[0] for k, v in gf._gb_arenas.items() if k[3] == im.shape[0])
    t = arena[64:64 + 8 * 16].view(torch.int64).cpu().tolist()
    print(f"{label}: N={g['data'].shape[1]} E={g['edge_index'].shape[1]}  kernel {((t[len(names)] - t[0]) / 1e3):.1f} us")
    for i, nm in enumerate(names):
        print(f"   {(t[i + 1] - t[i]) / 1e3:7.1f} us  {nm}")
