"""Which ATen ops launch kernels inside one training sample (eager), by Python call site: the glue left around the qmp kernels."""
import collections, os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import quadtree_mpnnlstm_b200 as q
from quadtree_mpnnlstm_b200.train import TrainStep

t_in, t_out = 2, 6
dev = torch.device("cuda")
mask = bench.ocean_mask()
cube = bench.synthetic_cube(t_in + t_out + 8)
clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
torch.manual_seed(21)
model = q.Seq2Seq(**bench.model_kwargs(t_in, t_out, 0.1), device=dev).to(dev).train()
step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
s = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in bench.sample(cube, clim, 0, t_in, t_out)]
for _ in range(2):
    step(*s)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(*s)
    torch.cuda.synchronize()
rows = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CPU and ev.name.startswith("aten::") and len(ev.kernels) > 0:
        st = [f for f in (ev.stack or []) if "quadtree_mpnnlstm_b200" in f or "bench.py" in f]
        rows[(ev.name, st[0] if st else "(autograd engine)")] += 1
print(f"{t_in}+{t_out} frames; kernel-launching ATen ops by call site:")
for (name, where), n in rows.most_common(40):
    print(f"{n:5d}  {name:28s} {where}")
