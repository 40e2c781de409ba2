"""Sweep of the MN-major / 128-byte-swizzle operand conventions of tcgen05.mma kind::tf32 (csrc/probes/mn_probe.cu)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import probe_lib

torch.manual_seed(0)
K, NB = 16, 64
A = torch.randn(K, 128, device="cuda")
B = torch.randn(K, NB, device="cuda")
ref = (A.double().T @ B.double()).float()
dump = torch.zeros((4 + NB // 32) * 2048, dtype=torch.uint8, device="cuda")
MN = (1 << 15) | (1 << 16)
SW128, SW128_32B = 3, 4          # CUtensorMapSwizzle: 128B, 128B_ATOM_32B
for name, extra, lbo, sbo, lt, kstep, swz in [
        ("MN base32b lbo=panel sbo=512", MN, 2048, 512, 1, 1024, SW128_32B),
        ("MN sw128 lbo=panel sbo=1024, TMA 128B (returns zeros)", MN, 2048, 1024, 2, 1024, SW128),
        # measured: lbo = 512 / sbo = panel gives garbage (2.5e+01); sbo = 1024 and a 128B-swizzled TMA with the base32b
        # descriptor end in illegal-address faults
]:
    D = torch.full((128, NB), float("nan"), device="cuda")
    probe_lib.call("qmp_mn_probe", A, B, D, dump, K, NB, extra, lbo, sbo, lt, kstep, swz)
    torch.cuda.synchronize()
    sm = dump.view(torch.float32).view(-1, 16, 32)             # [panel][node][32 floats as stored]
    src = torch.cat([A.view(16, 4, 32).permute(1, 0, 2), B.view(16, NB // 32, 32).permute(1, 0, 2)])     # [panel][node][32]
    e16, e32 = torch.empty_like(sm), torch.empty_like(sm)
    for r in range(16):
        for c in range(8):      # 16-byte chunks XOR (row & 7)
            e16[:, r, 4 * (c ^ (r & 7)):4 * (c ^ (r & 7)) + 4] = src[:, r, 4 * c:4 * c + 4]
        for c in range(4):      # 32-byte chunks XOR (row & 3)
            e32[:, r, 8 * (c ^ (r & 3)):8 * (c ^ (r & 3)) + 8] = src[:, r, 8 * c:8 * c + 8]
    lay = "16B^row&7" if torch.equal(sm, e16) else "32B^row&3" if torch.equal(sm, e32) else "plain" if torch.equal(sm, src) else "other"
    err = float((D - ref).abs().max())
    print(f"{name:60s} smem layout {lay:10s} max|D-ref| {err:9.3e}  nonzero {int((D != 0).sum()):5d}  nan {int(torch.isnan(D).sum())}")
