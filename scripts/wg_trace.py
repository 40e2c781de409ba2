"""Clock marks of CTA 0 of the weight-gradient kernel for the decoder cell (build with QMP_CELL_TRACE=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadtree_mpnnlstm_b200 import _lib
dev = torch.device("cuda")
N = 47200
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
xa, xb, dP = r(N, 4), r(N, 32), r(N, 128)
ZsA, dUsA, ZsB, dUsB = r(N, 4, 8), r(N, 4, 8), r(N, 4, 36), r(N, 4, 36)
gwa = torch.zeros(4, 448, device=dev); gwb = torch.zeros(4, 3332, device=dev)
def run():
    _lib.call("qmp_fused_wgrad", N, xa, 4, 4, 4, xb, 32, 32, 4, 1, 1, 32, dP, 128, ZsA, dUsA, ZsB, dUsB, gwa, gwb)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("wgrad decoder cell: %.1f us" % (e0.elapsed_time(e1) * 1e3))
L = _lib.lib()
if hasattr(L, "qmpx_wg_trace_dump"):
    buf = (ctypes.c_longlong * 64)()
    L.qmpx_wg_trace_dump(buf)
    v = list(buf)
    t0 = v[0]
    print("prologue done (fetch 0 issued next):", v[1] - t0)
    k = 2
    while k + 2 < 60 and v[k] > 0 and v[k] >= v[k - 1]:
        print(f"tile: MMA(prev) waited +{v[k] - v[k-1]:6d} | staged +{v[k+1] - v[k]:6d} | sync + next fetch issued +{v[k+2] - v[k+1]:6d}   (t = {v[k+2] - t0})")
        k += 3
    print("loop end", v[60] - t0, "| last MMA done", v[61] - t0, "| flush done", v[62] - t0, "| exit", v[63] - t0)
