#!/usr/bin/env python
"""Graph-build half of the hot path at the ice grid (229 x 361): device time of image_to_graph (dynamic quadtree,
thresh 0.15, dist_from_05, max cell 64, mask), the pixel-wise mesh, the static heterogeneous mesh, pool / unpool --
CUDA events, warm -- next to the CPU oracle on the host cores (the reference's own numpy/numba code is what the
oracle restates; SURVEY.md section 6 quotes 0.52 s per dynamic build for the reference)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from oracle import graph_ref as G

dev = torch.device("cuda")
H, W = bench.H, bench.W
mask = bench.ocean_mask()
cube = bench.synthetic_cube(12)
x = cube[:10]


def gpu_time(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def cpu_time(fn, iters=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        out = fn()
    return (time.perf_counter() - t0) / iters * 1e3, out


xg = torch.from_numpy(x).to(dev)
xp = q.add_positional_encoding(xg)
xc = G.add_positional_encoding(torch.from_numpy(x))
kw = dict(thresh=0.15, max_grid_size=64, mask=mask, transform_func=bench.dist_from_05, use_edge_attrs=True)
rows = []
tg, g = gpu_time(lambda: q.image_to_graph(xp, **kw))
tc, c = cpu_time(lambda: G.image_to_graph(xc, **kw))
same = bool(torch.equal(g["edge_index"].cpu(), c["edge_index"]) and np.array_equal(g["labels"].cpu().numpy(), c["labels"]))
rows.append(("image_to_graph dynamic quadtree (N=%d, E=%d), bit-exact=%s" % (g["data"].shape[1], g["edge_index"].shape[1], same), tg, tc))
kwp = dict(thresh=-np.inf, mask=mask, use_edge_attrs=True)
tg, g2 = gpu_time(lambda: q.image_to_graph(xp, **kwp))
tc, c2 = cpu_time(lambda: G.image_to_graph(xc, **kwp))
rows.append(("image_to_graph pixel-wise (N=%d, E=%d), edge_index equal=%s" % (g2["data"].shape[1], g2["edge_index"].shape[1],
                                                                          bool(torch.equal(g2["edge_index"].cpu(), c2["edge_index"]))), tg, tc))
tg, g3 = gpu_time(lambda: q.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=dev), iters=5)
tc, c3 = cpu_time(lambda: G.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12), iters=1)
rows.append(("create_static_heterogeneous_graph max cell 4 (N=%d, E=%d), edge_index equal=%s" % (
    int(g3["n_pixels_per_node"].shape[0]), g3["edge_index"].shape[1], bool(torch.equal(g3["edge_index"].cpu(), c3["edge_index"]))), tg, tc))
tg, d = gpu_time(lambda: q.flatten(xp, g["mapping"], g["n_pixels_per_node"], mask))
tc, dc = cpu_time(lambda: G.pool(xc, c["mapping"], c["n_pixels_per_node"], mask))
rows.append(("flatten 10 frames x 7 channels onto the quadtree mesh", tg, tc))
tg, u = gpu_time(lambda: q.unflatten(d[0], g["mapping"], (H, W), mask))
tc, uc = cpu_time(lambda: G.unpool(dc[0], c["mapping"], (H, W), mask))
rows.append(("unflatten one frame", tg, tc))
print(f"{'step':100s} {'B200 ms':>9s} {'CPU oracle ms':>14s} {'x':>7s}")
for name, a, b in rows:
    print(f"{name:100s} {a:9.3f} {b:14.1f} {b / a:7.0f}")
