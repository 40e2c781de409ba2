"""Training-step execution for the static-mesh path: one optimizer step per launch date, as in the
reference trainer (model/mpnnlstm.py:219-257: forward, MSE on unmasked pixels, backward, clip_grad_norm_(10),
Adam), optionally captured once into a CUDA graph and replayed.

Why a graph: one sample is ~100 dependent timesteps of a few short kernels each (thousands of launches,
5-50 us each); issued from Python the step is launch-bound by an order of magnitude.  With a static mesh
every shape is fixed, so the whole step -- forward, loss, backward, gradient all-reduce, clip, Adam -- is
captured after a few eager warm-up steps and replayed with zero host work per launch.

Data parallelism: launch dates are sharded across ranks; gradients are averaged
with ONE NCCL all-reduce of a flat fp32 bucket per step.
"""
from __future__ import annotations

import os

import torch

from .graph_functions import flatten, unflatten


def shard_launch_dates(n_dates, rank, world, seed=0, pad=False):
    """Launch dates (sample indices) of one rank: a seed-fixed permutation dealt round-robin, the same count on every rank so
    that every optimizer step / the final gather has exactly one collective on all ranks.  Training (``pad=False``) drops the
    remainder; inference (``pad=True``) must forecast EVERY date, so short ranks repeat the last date and the caller trims
    the padding after the gather (infer.predict_sharded)."""
    import numpy as np
    perm = np.random.default_rng(seed).permutation(n_dates) if seed else np.arange(n_dates)
    if pad:
        per = -(-n_dates // world)
        mine = [int(d) for d in perm[rank::world]]
        return mine + [int(perm[-1])] * (per - len(mine)) if n_dates else []
    per = n_dates // world
    return [int(d) for d in perm[rank:per * world:world]]


_expandable_done = [False]


def _expandable_segments():
    """Eager steps on data-dependent meshes allocate blocks whose sizes change from forecast step to forecast step (N, E vary);
    the caching allocator's fixed 20 MB segments fragment under that pattern and every miss is a device-synchronising
    ``cudaMalloc`` (measured: 51 per 20-frame sample, 31 -> 85 ... 290 ms, when one more ~2 MB workspace per step joined the mix).
    Expandable segments (one growing virtual range per stream, physical pages mapped on demand) take the misses away:
    configs[2] 120 -> 112 ms per sample with ``PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True`` from process start.  Switching
    an allocator that already holds fixed segments is another matter -- inside ``bench.py`` (after the captured static-mesh step
    and the rollouts) the switch made the dynamic-mesh sample bimodal, 105 or 140 ... 200 ms -- so the first eager ``TrainStep`` only
    does it when asked to: ``QMP_EXPANDABLE_SEGMENTS=1``."""
    if _expandable_done[0] or os.environ.get("QMP_EXPANDABLE_SEGMENTS", "0") != "1":
        return
    _expandable_done[0] = True
    if "expandable_segments" in os.environ.get("PYTORCH_CUDA_ALLOC_CONF", ""):
        return
    try:
        setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or torch.cuda.memory._set_allocator_settings
        setter("expandable_segments:True")
    except Exception:                      # an allocator backend without the setting: keep going
        pass


class TrainStep:
    def __init__(self, model, mask, lr=1e-4, graph_structure=None, use_cuda_graph=True, process_group=None,
                 world_size=1, max_norm=10.0):
        self.model, self.mask, self.graph_structure = model, mask, graph_structure
        self.params = [p for p in model.parameters()]
        # capturable foreach Adam divides by per-parameter bias-correction TENSORS: two one-tensor kernels per parameter
        # (652 of the step's launches, 2 ms at the ice configuration); the fused implementation is a handful of launches
        # (eager steps -- dynamic meshes -- are bound by the number of launches, so they take the fused implementation too)
        fused = bool(self.params and all(p.is_cuda for p in self.params))
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=use_cuda_graph, fused=fused or None)
        if not use_cuda_graph and fused:
            _expandable_segments()
        self.use_cuda_graph, self.pg, self.world, self.max_norm = use_cuda_graph, process_group, world_size, max_norm
        self.graph = None
        self.static = None
        self.loss = None
        self.eager_steps = 0
        self.launches_per_replay = 0
        self._bucket = None
        self.stream = None
        self.topology = None
        # dropout under replay: kernel arguments (the per-call seeds) are frozen by the capture, so the seeded kernels also
        # mix in this device counter, bumped by one device op at the start of every (captured) step
        self.salt = None
        if use_cuda_graph and self.params and self.params[0].is_cuda:
            self.salt = torch.zeros(1, dtype=torch.int64, device=self.params[0].device)

    # -- one eager step -----------------------------------------------------------------------------
    def _step(self, x, y, concat):
        from .ops import dropout_salt
        if self.salt is not None:
            self.salt.add_(1)
        with dropout_salt(self.salt):
            return self._step_body(x, y, concat)

    def _step_body(self, x, y, concat):
        for p in self.params:
            p.grad = None
        out, maps = self.model(x, y, concat, teacher_forcing_ratio=0, mask=self.mask, graph_structure=self.graph_structure)
        if self.model.thresh == -float("inf") or self.graph_structure is not None:
            # static mesh: compare on the nodes (== y[:, ~mask] on a pixel mesh)
            mapping = self.model.graph.mapping
            y_nodes = flatten(y, mapping, self.model.graph.n_pixels_per_node, self.mask)
            loss = torch.nn.functional.mse_loss(torch.stack(out), y_nodes)
        else:
            # dynamic quadtree: the mesh differs per forecast step -> unpool every step and compare the unmasked pixels,
            # as the reference trainer does (model/mpnnlstm.py:243-246)
            shape = tuple(x.shape[1:3])
            done = getattr(self.model, "_unpooled", {})         # images the decoder already unpooled while remeshing (do_remesh)

            def image(t):
                hit = done.get(id(out[t]))
                if hit is not None and hit[0] is out[t] and hit[1] is maps[t] and tuple(hit[2].shape[:2]) == shape:
                    return hit[2]
                return unflatten(out[t], maps[t], shape, self.mask)
            y_hat = torch.stack([image(t) for t in range(len(out))])
            keep = self._keep_mask(x.device)
            loss = torch.nn.functional.mse_loss(y_hat[:, keep], y[:, keep])
        loss.backward()
        if self.world > 1:
            import torch.distributed as dist
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat, group=self.pg)
            flat.div_(self.world)
            for p, g in zip(self.params, torch._utils._unflatten_dense_tensors(flat, grads)):
                p.grad = g
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.max_norm)
        self.opt.step()
        return loss.detach()

    def _reserve_headroom(self, device, factor=1.25):
        """Eager steps on data-dependent meshes: the bytes a sample keeps alive for its backward pass vary from sample to sample,
        and every new high-water mark is a ``cudaMalloc`` -- a device-synchronising call of 2 ... 14 ms in the middle of a step
        (profiles/r02c_dynprof_ice_after.txt).  After the second sample, grow the caching allocator's pool once to ``factor`` x
        the peak seen so far; later samples then carve their blocks out of cached segments."""
        peak, reserved = torch.cuda.max_memory_allocated(device), torch.cuda.memory_reserved(device)
        free, _ = torch.cuda.mem_get_info(device)
        want = min(max(int(factor * peak) - reserved, 256 << 20), free // 2)
        if want > (8 << 20):
            try:
                del_me = torch.empty(want, dtype=torch.uint8, device=device)
                del del_me
            except RuntimeError:          # out of memory: keep going without the headroom
                pass

    def _keep_mask(self, device):
        if getattr(self, "_keep", None) is None or self._keep.device != device:
            import numpy as np
            m = self.mask if self.mask is not None else np.zeros(1, bool)
            self._keep = torch.from_numpy(~np.asarray(m, dtype=bool)).to(device)
        return self._keep

    def _drop_autograd_leftovers(self):
        """Packed-parameter caches and the model's mesh state keep the previous step's autograd graph alive;
        a capture must not see nodes created on another stream."""
        for m in self.model.modules():
            if hasattr(m, "_cache"):
                m._cache.clear()
        g = self.model.graph
        if g is not None:
            self.topology = (g.pyg.edge_index, g.pyg.edge_attr, int(g.pyg.x.shape[0]))
        self.model.graph = None

    # -- public -------------------------------------------------------------------------------------
    def stage(self, x, y, concat):
        """Start the host -> device copy of the NEXT sample (pinned host tensors) on a copy stream and return device tensors
        that ``__call__`` accepts: the copy of sample i + 1 overlaps the step of sample i instead of preceding its own."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            out = [t.to(cur.device, non_blocking=True) for t in (x, y, concat)]
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        for t in out:
            t.record_stream(cur)
        out[0]._qmp_ready = ev
        return out

    def __call__(self, x, y, concat, warmup_eager=3):
        """Run one optimizer step on device tensors x [T_in,H,W,c], y [T_out,H,W,1], concat [T_out,H,W,1].
        Returns the (device) loss tensor of this step."""
        ev = getattr(x, "_qmp_ready", None)
        if ev is not None:                               # staged by stage(): wait for its copy
            torch.cuda.current_stream().wait_event(ev)
        if not self.use_cuda_graph:
            loss = self._step(x, y, concat)
            self.eager_steps += 1
            if self.eager_steps == 2 and x.is_cuda and os.environ.get("QMP_NO_HEADROOM", "0") == "0":
                self._reserve_headroom(x.device)
            return loss
        if self.graph is None:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            if self.eager_steps < warmup_eager:        # let allocator / caches / lazy state settle first,
                self.eager_steps += 1                  # on the side stream the capture will use
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    loss = self._step(x, y, concat)
                torch.cuda.current_stream().wait_stream(self.stream)
                return loss
            self.static = [t.clone() for t in (x, y, concat)]
            self._drop_autograd_leftovers()
            torch.cuda.synchronize()
            from . import _lib
            before = _lib.kernel_launches()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.loss = self._step(*self.static)
            self._drop_autograd_leftovers()
            self.launches_per_replay = _lib.kernel_launches() - before   # qmp kernels inside one replay
            # the capture itself does not execute; fall through to the first replay
        for s, t in zip(self.static, (x, y, concat)):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss
