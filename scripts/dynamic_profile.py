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
kw = bench.model_kwargs(dropout=0.1)
kw["thresh"] = 0.15
torch.manual_seed(21)
model = q.Seq2Seq(**kw, device=dev).to(dev).train()
step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
smp = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, d)] for d in range(4)]
for s_ in smp[:2]:
    step(*s_)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
step(*smp[2])
torch.cuda.synchronize()
print("wall ms per sample", (time.perf_counter() - t0) * 1e3)
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
