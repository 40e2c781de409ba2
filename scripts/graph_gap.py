"""Cost of one dependent kernel node inside a replayed CUDA graph (launch + drain gap), measured with a trivial kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadtree_mpnnlstm_b200 import _lib

dev = torch.device("cuda")
y = torch.randn(1024, device=dev); g = torch.randn(1024, device=dev); o = torch.empty(1024, device=dev)
big = [torch.randn(47200, 32, device=dev) for _ in range(3)]
for n_nodes, args, label in ((2000, (y, g, o, 1024), "tiny kernel (1 CTA)"), (2000, (big[0], big[1], big[2], 47200 * 32), "6 MB elementwise kernel")):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            _lib.call("qmp_relu_mask_to", *args)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(n_nodes):
                _lib.call("qmp_relu_mask_to", *args)
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            gr.replay()
        e1.record(s)
        torch.cuda.synchronize()
        print(f"{label}: {e0.elapsed_time(e1) / 5 / n_nodes * 1000:.2f} us per node")
