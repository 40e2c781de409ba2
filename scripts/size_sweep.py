#!/usr/bin/env python
"""Per-launch time of every hot C-ABI call against the mesh size: eager training samples (2 + 3 frames) on pixel-wise meshes
of growing grids, each call timed alone with CUDA events (warm L2, no flush).  Shows the size-independent part of each
kernel (prologue, weight image, allocation, tail) next to the part that scales with the nodes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib
from quadtree_mpnnlstm_b200.train import TrainStep

dev = torch.device("cuda")
_lib.set_pdl(False)
grids = [(12, 12), (40, 40), (64, 64), (128, 128), (229, 361)]
table = {}
sizes = []
for (h, w) in grids:
    mask = bench.ocean_mask(h, w)
    cube = bench.synthetic_cube(8, h=h, w=w)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    torch.manual_seed(21)
    model = q.Seq2Seq(**bench.model_kwargs(2, 3, 0.1), device=dev).to(dev).train()
    step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
    smp = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, 0, 2, 3)]
    step(*smp)
    step(*smp)
    N = int(model.graph.pyg.x.shape[0])
    sizes.append(N)
    records, orig = [], _lib.call

    def timed(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *args)
        e1.record()
        records.append((bench._classify(name, args)[1], e0, e1))

    _lib.call = timed
    try:
        for _ in range(3):
            step(*smp)
    finally:
        _lib.call = orig
    torch.cuda.synchronize()
    acc = {}
    for key, a, b in records:
        acc.setdefault(key, []).append(a.elapsed_time(b) * 1e3)
    for key, v in acc.items():
        v.sort()
        table.setdefault(key, {})[N] = v[len(v) // 2]
    del model, step
print("median us per call; columns = N nodes:", sizes)
for key in sorted(table, key=lambda k: -table[k].get(sizes[-1], 0)):
    print(f"{key[:72]:72s} " + " ".join(f"{table[key].get(n, float('nan')):8.1f}" for n in sizes))
