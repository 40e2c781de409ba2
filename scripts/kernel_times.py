#!/usr/bin/env python
"""Per-entry-point GPU times of one training sample at the bench configuration (CUDA events around every C-ABI
call, eager mode, warm caches).  Usage: python scripts/kernel_times.py [t_in t_out]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib
from quadtree_mpnnlstm_b200.train import TrainStep

t_in, t_out = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 12)
dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(t_in + t_out + 8)
clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
torch.manual_seed(21)
model = q.Seq2Seq(**bench.model_kwargs(t_in, t_out), device=dev).to(dev).train()
step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
samples = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, d, t_in, t_out)] for d in range(3)]
for s in samples[:2]:
    step(*s)
torch.cuda.synchronize()

records = []
orig = _lib.call


def timed(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *args)
    e1.record()
    key = name
    if name.startswith("qmp_fused_fwd") or name.startswith("qmp_fused_bwd"):
        # N ptr ptr ea xa lda DA GA wa xb ldb DB GB shared wb mode
        key = f"{name}[DA={args[6]} GA={args[7]} DB={args[11]} GB={args[12]} mode={args[15]}]"
    elif name == "qmp_fused_wgrad":
        key = f"{name}[DA={args[3]} GA={args[4]} DB={args[7]} GB={args[8]} mode={args[10]}]"
    records.append((key, e0, e1))


_lib.call = timed
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
step(*samples[2])
t1.record()
torch.cuda.synchronize()
tot, cnt = collections.defaultdict(float), collections.Counter()
for k, a, b in records:
    tot[k] += a.elapsed_time(b) * 1e3
    cnt[k] += 1
T = sum(tot.values())
print(f"sample {t_in}+{t_out} frames: wall {t0.elapsed_time(t1):.1f} ms (eager), sum of qmp calls {T / 1e3:.1f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{v:10.0f} us {100 * v / T:5.1f}% n={cnt[k]:4d} avg {v / cnt[k]:8.1f} us  {k}")
