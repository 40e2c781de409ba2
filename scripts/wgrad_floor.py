import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadtree_mpnnlstm_b200 import _lib
dev = "cuda"
tot = lambda dc: (dc + 2) * dc + dc + 4 + 32 * (dc + 4) + 32 * dc + 32
def run(N, GB, mode=0, C=32, iters=20):
    xb = torch.randn(N, 32, device=dev)
    lddp = 128 if mode == 1 else GB * C
    dP = torch.randn(N, lddp, device=dev)
    Zs = torch.randn(N, GB, 36, device=dev); dUs = torch.randn(N, GB, 36, device=dev)
    gwb = torch.zeros(GB, tot(32), device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for k in range(iters + 3):
        if k >= 3: ev[k-3][0].record()
        _lib.call("qmp_fused_wgrad", N, None, 0, 0, 0, xb, 32, 32, GB, 1, mode, C, dP, lddp, None, None, Zs, dUs, None, gwb)
        if k >= 3: ev[k-3][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters * 1e3
for N in (64, 1024, 9472, 18944, 47200, 94400):
    print(f"N={N:6d}  GB=1: {run(N,1):7.1f} us   GB=4 (mode 1): {run(N,4,1):7.1f} us", flush=True)
