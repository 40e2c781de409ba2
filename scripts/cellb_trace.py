"""Timeline of CTA 0 of the decoder-cell BACKWARD kernel (build with QMP_CELL_TRACE=1) + CUDA-event time of the launch."""
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
dP = torch.randn(N, 128, generator=gen).to(dev)
z = lambda *s: torch.empty(s, device=dev)
o = dict(gates=z(N, 128), Craw=z(N, 32), O=z(N, 32), H=z(N, 32), C=z(N, 32), head=z(N, 36), logit=z(E, 8), mstat=z(N, 8), linv=z(N, 8),
         usave=z(N, 128))
_lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, xb, 32, FZ.cell_image(wa, wb), Cp, prm, 1, 1, 1, 1e-5,
          o["gates"], o["Craw"], o["O"], o["H"], o["C"], o["head"], 36, None, o["logit"], o["mstat"], o["linv"], o["usave"], 0.0, 1)
b = dict(ZsA=z(N, 128), dUsA=z(N, 128), ZsB=z(N, 64), dUsB=z(N, 32), dxa=z(N, 4), dxb=z(N, 32))      # zB, duB, sd, sg of the panel layout
img = FZ.cell_bwd_image(wa, wb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    _lib.call("qmp_fused_cell_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xa, 4, xb, 32, img, o["usave"], dP, 128, None, None, None, None, 0, 0, 0, 0.0, None, None, None, None, 0, None, None, o["logit"],
              o["mstat"], o["linv"], b["ZsA"], b["dUsA"], b["ZsB"], b["dUsB"], b["dxa"], b["dxb"], 0.0, 1)


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"qmp_fused_cell_bwd (2 memsets + kernel)  median {ts[10]:7.1f} us   min {ts[0]:7.1f} us   N {N} E {E}")
L = _lib.lib()
if hasattr(L, "qmpx_cellb_trace_dump"):
    buf = (ctypes.c_float * 4096)()
    L.qmpx_cellb_trace_dump(buf, 1)
    run()
    torch.cuda.synchronize()
    L.qmpx_cellb_trace_dump(buf, 0)
    d = np.frombuffer(buf, dtype=np.float32).reshape(2, 2048)
    names = {1: "tile start", 2: "g rows staged", 3: "sync, G1 issued (thread 0)", 4: "first-pass reads issued, G1 complete", 5: "dz dumped",
             7: "sync", 8: "edge phase done", 9: "sync", 10: "X reads issued, [du|dw] staged", 6: "sync, G2 issued, X convs done",
             11: "G2 complete, sync", 12: "dx flushed", 13: "sync"}
    n = int(d[0, 0])
    tags = d[0, 1:1 + 2 * n:2].astype(int)
    clk = d[0, 2:2 + 2 * n:2].astype(np.int64)
    prev = t0 = clk[0]
    for tg, c in zip(tags, clk):
        print(f"{(c - t0) % (1 << 24):8d} (+{(c - prev) % (1 << 24):6d})  {names.get(tg, tg)}")
        prev = c
