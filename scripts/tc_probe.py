import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import probe_lib as _lib
torch.manual_seed(0)
for (M, N, K) in [(128, 32, 8), (128, 32, 32), (128, 136, 32), (300, 128, 40), (1000, 256, 72), (128, 8, 8)]:
    for split in (0, 1):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.full((M, N), float("nan"), device="cuda")
        _lib.call("qmp_tc_gemm_probe", A, B, C, M, N, K, split)
        torch.cuda.synchronize()
        ref = A.double() @ B.double().T
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        print(f"M={M} N={N} K={K} split={split}: rel err {err:.3e}  nan={int(torch.isnan(C).sum())}", flush=True)
