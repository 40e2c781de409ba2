"""CUDA-event time of qmp_head_bwd (memset + kernel) at the bench mesh size, L2 flushed before every launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadtree_mpnnlstm_b200 import _lib, fused as FZ
from quadtree_mpnnlstm_b200.graph_csr import get_csr
import bench, quadtree_mpnnlstm_b200 as q
dev = torch.device("cuda")
mask = bench.ocean_mask()
gs = q.graph_functions.create_static_homogeneous_graph if False else None
img = torch.zeros(1, bench.H, bench.W, 1, device=dev)
g = q.image_to_graph(q.add_positional_encoding(img), thresh=-float("inf"), mask=mask, use_edge_attrs=True)
ei, ea = g["edge_index"], g["edge_attrs"]
N, E = int(g["data"].shape[1]), int(ei.shape[1])
csr = get_csr(ei, ea, N)
gen = torch.Generator().manual_seed(1)
xb = torch.randn(N, 36, generator=gen).to(dev)
wb = (torch.randn(1, FZ.conv_total(36), generator=gen) * 0.2).to(dev)
dP = torch.randn(N, 32, generator=gen).to(dev)
logit = torch.randn(E, 1, generator=gen).to(dev)
ptr = csr.in_ptr.long()
seg = torch.repeat_interleave(torch.arange(N, device=dev), ptr[1:] - ptr[:-1])
mstat = torch.full((N, 1), -1e30, device=dev).scatter_reduce(0, seg[:, None], logit, "amax")
ssum = torch.zeros(N, 1, device=dev).index_add_(0, seg, (logit - mstat[seg]).exp())
linv = 1.0 / ssum.clamp(min=1e-30)
Zs, dUs, dx = torch.empty(N, 40, device=dev), torch.empty(N, 40, device=dev), torch.empty(N, 36, device=dev)
image = FZ.head_bwd_image(wb)
flush = torch.empty(64 * 1024 * 1024, device=dev)
ts = []
for r in range(15):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("qmp_head_bwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, xb, 36, image, dP, 32, logit, mstat, linv, Zs, dUs, dx, 0.1, 7)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1000)
ts.sort()
print(f"qmp_head_bwd N {N} E {E}: median {ts[7]:.1f} us  min {ts[0]:.1f} us")
