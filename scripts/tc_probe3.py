import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import probe_lib as _lib
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for a_tmem in (0, 1):
    for rotate in (0, 1):
        for N in (16, 32, 48, 64, 128, 256):
            if rotate and N > 64:
                continue
            res = []
            for reps in (1, 8, 64, 256):
                _lib.call("qmp_tc_probe3", out, N, reps, a_tmem, rotate)
                _lib.call("qmp_tc_probe3", out, N, reps, a_tmem, rotate)
                torch.cuda.synchronize()
                res.append((reps, int(out.item())))
            per = (res[-1][1] - res[-2][1]) / (res[-1][0] - res[-2][0])
            print(f"A={'tmem' if a_tmem else 'smem'} rotate={rotate} N={N:3d}: " + " ".join(f"{r}:{c}" for r, c in res) + f"  -> {per:.1f} clk/MMA", flush=True)
