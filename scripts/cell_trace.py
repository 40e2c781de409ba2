"""Timeline of CTA 0 of the decoder-cell kernel (build with QMP_CELL_TRACE=1): clock marks of threads 0 (issues the MMAs)
and 160.  Usage on the GPU box: QMP_CELL_TRACE=1 python quadtree_mpnnlstm_b200/csrc/build.py --force && python scripts/cell_trace.py"""
import ctypes
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
L = _lib.lib()
buf = (ctypes.c_float * 4096)()
for it in range(3):
    L.qmpx_cell_trace_dump(buf, 1)
    _lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, xb, 32, ic, Cp, prm, 1, 1, 1, 1e-5, o["gates"],
              o["Craw"], o["O"], o["H"], o["C"], o["head"], 36, concat, o["logit"], o["mstat"], o["linv"], None, 0.0, 1)
    torch.cuda.synchronize()
L.qmpx_cell_trace_dump(buf, 0)
d = np.frombuffer(buf, dtype=np.float32).reshape(2, 2048)
names = {1: "tile start (wait G1)", 6: "G1 complete", 8: "U dumped, first gathers issued (+sync)", 9: "edge phase done", 10: "sync",
         11: "A=z staged (+sync)", 12: "G2 issued (thread 0)", 5: "next tile: rows/indices prefetched, X convs done", 13: "G2 complete",
         14: "P dumped, next [h|x] staged (+sync)", 4: "G1(next) issued (thread 0)", 15: "epilogue done", 16: "sync"}
for th, label in ((0, "thread 0"), (1, "thread 160")):
    n = int(d[th, 0])
    tags = d[th, 1:1 + 2 * n:2].astype(int)
    clk = d[th, 2:2 + 2 * n:2].astype(np.int64)
    print(f"--- {label}: {n} marks")
    prev = t0 = clk[0]
    for tg, c in zip(tags, clk):
        print(f"{(c - t0) % (1 << 24):8d} (+{(c - prev) % (1 << 24):6d})  {names.get(tg, tg)}")
        prev = c

cta = (ctypes.c_ulonglong * 1024)()
L.qmpx_cell_cta_dump(cta)
c = np.frombuffer(cta, dtype=np.uint64).reshape(256, 4)[:148].astype(np.int64)
t0 = c[:, 0].min()
c = c - t0
print("--- per CTA, ns from the first CTA's entry (globaltimer): entry, prologue done (first G1 issued), last tile done, exit")
for k, name in enumerate(("entry", "prologue done", "tiles done", "exit")):
    print(f"{name:16s} min {c[:, k].min():7d}  median {int(np.median(c[:, k])):7d}  max {c[:, k].max():7d}")
print("CTA 0:", c[0].tolist(), " CTA 147:", c[147].tolist())
