"""Timeline of one CTA of the paired-warp forward kernel (build with QMP_PW_TRACE=1): clock64 marks of thread 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200 import _lib, fused
from quadtree_mpnnlstm_b200.graph_csr import get_csr
dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(4)
xg = q.add_positional_encoding(torch.from_numpy(cube[:1]).to(dev))
gs = q.image_to_graph(xg, thresh=-np.inf, mask=mask, use_edge_attrs=True)
N = gs["data"].shape[1]
csr = get_csr(gs["edge_index"], gs["edge_attrs"], N)
E, C = csr.n_edges, 32
torch.manual_seed(0)
model = q.Seq2Seq(**bench.model_kwargs(), device=dev).to(dev)
cell = model.decoder.rnns[0]
with torch.no_grad():
    wa = fused.tc_image(fused.pack_fused(cell._convs("x", 0), 4), 4)
    wb = fused.tc_image(fused.pack_fused(cell._convs("h", 0), C), C)
    prm = cell._gate_params(-1, model.decoder.norm_h, model.decoder.norm_c, model.decoder.norm_o).contiguous()
f32 = dict(dtype=torch.float32, device=dev)
X, Hs, Cs, cc = torch.randn(N, 4, **f32), torch.randn(N, C, **f32), torch.randn(N, C, **f32), torch.randn(N, **f32)
gates = torch.empty(N, 4 * C, **f32)
Craw, O, Hn, Cn = (torch.empty(N, C, **f32) for _ in range(4))
head = torch.empty(N, fused.HEADW, **f32)
logit, ms, li = torch.empty(E, 8, **f32), torch.empty(N, 8, **f32), torch.empty(N, 8, **f32)
dbg = torch.zeros(4096, **f32)
for it in range(3):
    dbg.zero_()
    _lib.call("qmp_fused_fwd_tc", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, X, 4, 4, 4, wa, Hs, C, C, 4, 1, wb,
              1, 0, C, dbg, 8 * C, Cs, prm, 1, 1, 1, 1e-5, gates, Craw, O, Hn, Cn, head, fused.HEADW, cc, logit, ms, li, 0.0, 0)
    torch.cuda.synchronize()
d = dbg.cpu().numpy()
n = int(d[0])
tags = d[1:1 + 2 * n:2].astype(int)
clk = d[2:2 + 2 * n:2].astype(np.int64)
t0 = clk[0]
names = {1: "Xconv start", 2: "Xconv end", 10: "H start", 11: "own row loaded", 12: "pending MMA2 waited", 13: "x staged+sync",
         14: "weights landed", 15: "MMA1 issued", 16: "gather0 issued/loaded", 17: "U ready", 18: "edges done", 19: "z staged+sync",
         20: "MMA2 issued", 30: "slot end", 31: "P ready", 32: "epilogue done"}
prev = t0
for tg, c in zip(tags, clk):
    dt = (c - prev) % (1 << 24)
    print(f"{(c - t0) % (1 << 24):8d} (+{dt:6d})  {names.get(tg, tg)}")
    prev = c
