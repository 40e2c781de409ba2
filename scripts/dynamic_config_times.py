#!/usr/bin/env python
"""configs[2] (ice_exp with per-frame dynamic quadtree re-decomposition) and configs[0]-like (MNIST 64x64 ChebConv):
wall time of one fwd+bwd+Adam sample, eager (data-dependent mesh sizes: no CUDA graph), after 2 warm-up samples."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200.train import TrainStep

dev = torch.device("cuda")


def run(name, model, mask, samples, frames):
    step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
    for s in samples[:2]:
        step(*s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in samples[2:]:
        loss = step(*s)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / len(samples[2:])
    print(f"{name}: {dt * 1e3:.1f} ms / sample -> {frames / dt:.1f} graph-frames/s (loss {float(loss):.4f})", flush=True)


# configs[2]: ice grid, dynamic quadtree, TransformerConv
mask = bench.ocean_mask()
cube = bench.synthetic_cube(bench.FRAMES + 8)
clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
kw = bench.model_kwargs()
kw["thresh"] = 0.15
torch.manual_seed(21)
model = q.Seq2Seq(**kw, device=dev).to(dev).train()
samples = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, d)] for d in range(4)]
run("configs[2] ice dynamic quadtree (229x361, thresh 0.15, remesh every step, 10+90 frames)", model, mask, samples, bench.FRAMES)

# configs[0]: MNIST-like 64x64, ChebConv, hidden 16, 2 layers, 10+10 frames
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import moving_blob
rng = np.random.default_rng(1)
ms = np.zeros((64, 64), bool)
samples = []
for d in range(6):
    x, y = moving_blob(rng, 10, 64, 64, size=28), moving_blob(rng, 10, 64, 64, size=28)
    samples.append([torch.from_numpy(a).to(dev) for a in (x, y, np.zeros((10, 64, 64, 1), np.float32))])
torch.manual_seed(1)
m2 = q.Seq2Seq(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=10, input_features=4, output_timesteps=10, n_layers=2,
               device=dev).to(dev).train()
run("configs[0] MNIST-like 64x64 dynamic quadtree ChebConv (10+10 frames)", m2, ms, samples, 20)
