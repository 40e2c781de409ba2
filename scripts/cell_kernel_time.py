#!/usr/bin/env python
"""Decoder-cell forward at the bench mesh (N = 47 200, E = 187 808): the per-conv tcgen05 kernel (qmp_fused_fwd_tc) against
the gates-batched persistent kernel (qmp_fused_cell_fwd), CUDA events, L2 flushed between launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib, fused as FZ, graph_csr

dev = torch.device("cuda")
import numpy as np
mask = bench.ocean_mask()
x = torch.zeros(1, mask.shape[0], mask.shape[1], 3, device=dev)
gs = q.image_to_graph(x, thresh=-np.inf, mask=torch.as_tensor(mask), use_edge_attrs=True)
N = int(gs["data"].shape[1])
csr = graph_csr.get_csr(gs["edge_index"], gs["edge_attrs"], N)
E = csr.n_edges
print("N", N, "E", E)
gen = torch.Generator(device="cpu").manual_seed(0)
xa, xb, Cp = (torch.randn(N, w, generator=gen).to(dev) for w in (4, 32, 32))
wa = (torch.randn(4, FZ.conv_total(4), generator=gen) * 0.3).to(dev)
wb = (torch.randn(4, FZ.conv_total(32), generator=gen) * 0.2).to(dev)
prm = (torch.randn(13, 32, generator=gen) * 0.5).to(dev)
concat = torch.randn(N, generator=gen).to(dev)
z = lambda *s: torch.empty(s, device=dev)
o = dict(gates=z(N, 128), Craw=z(N, 32), O=z(N, 32), H=z(N, 32), C=z(N, 32), head=z(N, 36), logit=z(E, 8), mstat=z(N, 8), linv=z(N, 8))
ia, ib, ic = FZ.tc_image(wa, 4), FZ.tc_image(wb, 32), FZ.cell_image(wa, wb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def old():
    _lib.call("qmp_fused_fwd_tc", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, 4, 4, ia, xb, 32, 32, 4, 1, ib, 1, 0, 32, None, 256,
              Cp, prm, 1, 1, 1, 1e-5, o["gates"], o["Craw"], o["O"], o["H"], o["C"], o["head"], 36, concat, o["logit"], o["mstat"],
              o["linv"], 0.0, 1)


def new():
    _lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, xb, 32, ic, Cp, prm, 1, 1, 1, 1e-5, o["gates"],
              o["Craw"], o["O"], o["H"], o["C"], o["head"], 36, concat, o["logit"], o["mstat"], o["linv"], None, 0.0, 1)


modes = sys.argv[1:] or ["write"]
for mode in modes:
  print("L2 flush between launches:", mode)
  for name, fn in (("qmp_fused_fwd_tc", old), ("qmp_fused_cell_fwd", new)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        if mode == "write":
            flush.zero_()
        elif mode == "read":
            flush.view(torch.int32).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:24s} median {ts[len(ts) // 2]:7.1f} us   min {ts[0]:7.1f} us")
