#!/usr/bin/env python
"""What a programmatic dependent launch buys inside a replayed CUDA graph: chains of (a) the decoder-cell forward kernel at
the bench mesh, (b) the fc_out2 pair of small kernels (qmp_tconv1_fwd), captured with qmp_set_pdl(1) and qmp_set_pdl(0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib, fused as FZ, graph_csr

dev = torch.device("cuda")
mask = bench.ocean_mask()
x = torch.zeros(1, mask.shape[0], mask.shape[1], 3, device=dev)
gs = q.image_to_graph(x, thresh=-np.inf, mask=torch.as_tensor(mask), use_edge_attrs=True)
N = int(gs["data"].shape[1])
csr = graph_csr.get_csr(gs["edge_index"], gs["edge_attrs"], N)
E = csr.n_edges
gen = torch.Generator(device="cpu").manual_seed(0)
xa, xb, Cp = (torch.randn(N, w, generator=gen).to(dev) for w in (4, 32, 32))
wa = (torch.randn(4, FZ.conv_total(4), generator=gen) * 0.3).to(dev)
wb = (torch.randn(4, FZ.conv_total(32), generator=gen) * 0.2).to(dev)
prm = (torch.randn(13, 32, generator=gen) * 0.5).to(dev)
concat = torch.randn(N, generator=gen).to(dev)
z = lambda *s: torch.empty(s, device=dev)
o = dict(gates=z(N, 128), Craw=z(N, 32), O=z(N, 32), H=z(N, 32), C=z(N, 32), head=z(N, 36), logit=z(E, 8), mstat=z(N, 8), linv=z(N, 8))
ic = FZ.cell_image(wa, wb)
P1 = torch.randn(136, generator=gen).to(dev)
s4, y1 = z(N, 4), z(N)


def cell():
    _lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, xb, 32, ic, Cp, prm, 1, 1, 1, 1e-5, o["gates"],
              o["Craw"], o["O"], o["H"], o["C"], o["head"], 36, concat, o["logit"], o["mstat"], o["linv"], None, 0.0, 1)


def small():
    _lib.call("qmp_tconv1_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xb, 32, P1, s4, y1, 0.0, 1)


for label, fn, n_nodes in (("decoder-cell forward", cell, 200), ("qmp_tconv1_fwd (2 kernels)", small, 1000)):
    for pdl in (0, 1, 0, 1):
        _lib.set_pdl(bool(pdl))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(n_nodes):
                    fn()
            for _ in range(2):
                gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                gr.replay()
            e1.record(s)
            torch.cuda.synchronize()
            print(f"{label}: pdl={pdl}  {e0.elapsed_time(e1) / 5 / n_nodes * 1000:.2f} us per call", flush=True)
