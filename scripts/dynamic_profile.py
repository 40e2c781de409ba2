#!/usr/bin/env python
"""Where one configs[2] sample (dynamic quadtree, eager) spends its host time: cProfile of a training step after warm-up, next
to the summed device time of its kernels (torch.profiler)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200.train import TrainStep

dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(bench.FRAMES + 8)
clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
if len(sys.argv) > 1 and sys.argv[1] == "mnist":        # configs[0]-like: 64 x 64 moving blob, ChebConv, hidden 16, 2 layers, 10 + 10
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import moving_blob
    rng = np.random.default_rng(1)
    mask = np.zeros((64, 64), bool)
    smp = []
    for d in range(6):
        x, y = moving_blob(rng, 10, 64, 64, size=28), moving_blob(rng, 10, 64, 64, size=28)
        smp.append([torch.from_numpy(a).to(dev) for a in (x, y, np.zeros((10, 64, 64, 1), np.float32))])
    torch.manual_seed(1)
    model = q.Seq2Seq(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=10, input_features=4, output_timesteps=10,
                      n_layers=2, device=dev).to(dev).train()
else:
    kw = bench.model_kwargs(dropout=0.1)
    kw["thresh"] = 0.15
    torch.manual_seed(21)
    model = q.Seq2Seq(**kw, device=dev).to(dev).train()
    smp = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, d)] for d in range(5)]
import quadtree_mpnnlstm_b200.seq2seq as _s2s
_sizes = []
_orig_i2g = _s2s.image_to_graph


def _rec_i2g(*a, **k):
    g = _orig_i2g(*a, **k)
    _sizes.append((int(g["data"].shape[1]), int(g["edge_index"].shape[1])))
    return g


_s2s.image_to_graph = _rec_i2g
step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
for s_ in smp[:3]:
    step(*s_)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
step(*smp[3])
torch.cuda.synchronize()
print("wall ms per sample", (time.perf_counter() - t0) * 1e3)
_ns = np.array([n for n, _ in _sizes[-(len(_sizes) // 4):]])
print("mesh sizes of the last sample (N per build): first", _ns[:5], "min", _ns.min(), "median", int(np.median(_ns)), "max", _ns.max())
pr = cProfile.Profile()
pr.enable()
step(*smp[3])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(*smp[2])
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev) if hasattr(ev[0], "device_time_total") else sum(e.cuda_time_total for e in ev)
print("summed device time ms", tot / 1e3)
print(ev.table(sort_by="device_time_total" if hasattr(ev[0], "device_time_total") else "cuda_time_total", row_limit=25, max_name_column_width=60))
print(ev.table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
