"""Rollout inference over many launch dates (reference: NextFramePredictorS2S.predict, model/mpnnlstm.py:402-440, driven by
ice_inf.py:60-130: per launch date forward -> unflatten -> stack to [n_dates, T_out, H, W, 1]).

* ``Rollout``: one launch date's no_grad forward + un-pooling.  On a static mesh (pixel-wise ``thresh = -inf`` or a preset
  ``graph_structure``, the ice_inf.py configuration) every shape is fixed, so the whole 100-frame rollout is captured ONCE
  into a CUDA graph and replayed per launch date: on the N = 4 066 heterogeneous mesh every kernel is a few microseconds and
  an eager Python loop is pure launch latency.  Dynamic-quadtree rollouts (data-dependent N, E) stay eager.
* ``predict_sharded``: launch dates are independent, so they are dealt round-robin to the ranks (one process per GPU) with NO
  collective on the data path; short ranks are padded so that every rank joins the single ``all_gather`` at the end, and the
  padding is trimmed afterwards -- every date is forecast.
* ``RolloutPool``: several ``Rollout`` lanes (own captured graph, static buffers and stream each) replayed CONCURRENTLY.  On
  the ice_inf.py mesh (N = 4 066) a rollout's persistent kernels are 32 CTAs on a 148-SM part and every launch is latency-
  bound; independent launch dates on different streams fill the other SMs."""
from __future__ import annotations

import torch

from .graph_functions import _Unpool, _as_mesh, unflatten
from .train import shard_launch_dates


def _static_mesh(model, graph_structure):
    return graph_structure is not None or model.thresh == -float("inf")


class Rollout:
    """``rollout(x [T_in,H,W,c], concat_layers [T_out,H,W,1]) -> [T_out, H, W, 1]`` (device tensor; with a captured graph it is
    the graph's static output buffer, valid until the next call)."""

    def __init__(self, model, mask, graph_structure=None, high_interest_region=None, remesh_every=1, use_cuda_graph=True,
                 warmup_eager=2):
        self.model, self.mask, self.gs, self.hir, self.remesh_every = model, mask, graph_structure, high_interest_region, remesh_every
        self.use_cuda_graph = bool(use_cuda_graph) and _static_mesh(model, graph_structure)
        self.warmup_eager, self.eager_calls = warmup_eager, 0
        self.graph = self.static = self.out = self.stream = None
        self.launches_per_replay = 0
        # predict() of the reference never calls eval(): with the model in train() mode the attention dropout stays on, and a
        # replayed graph must resample it like a fresh call would (ops.dropout_salt)
        self.salt = None
        p0 = next(model.parameters(), None)
        if self.use_cuda_graph and p0 is not None and p0.is_cuda:
            self.salt = torch.zeros(1, dtype=torch.int64, device=p0.device)

    @torch.no_grad()
    def _forward(self, x, cl):
        from .ops import dropout_salt
        if self.salt is not None:
            self.salt.add_(1)
        with dropout_salt(self.salt):
            return self._forward_body(x, cl)

    def _forward_body(self, x, cl):
        m = self.model
        y_hat, maps = m(x, None, cl, teacher_forcing_ratio=0, mask=self.mask, high_interest_region=self.hir,
                        graph_structure=self.gs, remesh_every=self.remesh_every)
        shape = tuple(x.shape[1:3])
        if _static_mesh(m, self.gs):        # one mesh for every step: un-pool the whole [T_out, N, 1] stack in ONE launch
            mesh = _as_mesh(maps[0], shape, self.mask, x.device)
            fill = float("nan") if (maps[0] is None and self.mask is not None) else 0.0
            return _Unpool.apply(torch.stack(y_hat).float(), mesh, fill)
        return torch.stack([unflatten(y_hat[t], maps[t], shape, self.mask) for t in range(len(y_hat))])

    def __call__(self, x, cl):
        if not self.use_cuda_graph:
            return self._forward(x, cl)
        if self.graph is None:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            if self.eager_calls < self.warmup_eager:      # allocator / mesh caches / lazy state settle on the capture stream
                self.eager_calls += 1
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    out = self._forward(x, cl)
                torch.cuda.current_stream().wait_stream(self.stream)
                return out
            from . import _lib
            self.static = [x.clone(), cl.clone()]
            torch.cuda.synchronize()
            before = _lib.kernel_launches()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = self._forward(*self.static)
            self.launches_per_replay = _lib.kernel_launches() - before
        self.static[0].copy_(x, non_blocking=True)
        self.static[1].copy_(cl, non_blocking=True)
        self.graph.replay()
        return self.out


def default_lanes(n_nodes, n_sm=148, most=8):
    """How many rollouts to keep in flight: the persistent kernels run one CTA per 128-node tile and SM, and a second set of
    lanes fills the launch gaps and tails of the first (measured on the N = 4 066 mesh, end to end: 2 lanes 203, 4 lanes 324,
    8 lanes 442, 12 lanes 453 launch dates/s; one lane 46)."""
    tiles = max(1, -(-int(n_nodes) // 128))
    return max(1, min(most, 2 * n_sm // tiles))


class RolloutPool:
    """``lanes`` captured rollouts of one model replayed on ``lanes`` streams: launch date i runs on lane i % lanes.  Every
    lane owns its graph, its static input / output buffers and its dropout salt (offset per lane, so that two lanes never
    draw the same attention-dropout masks); the lanes share the model's parameters, the mesh and its CSR, which a replay
    only reads."""

    def __init__(self, model, mask, graph_structure=None, high_interest_region=None, remesh_every=1, lanes=None, warmup_eager=2):
        if not _static_mesh(model, graph_structure):
            raise ValueError("RolloutPool needs a static mesh (thresh = -inf or a preset graph_structure): a dynamic-quadtree "
                             "rollout cannot be captured")
        if lanes is None:
            lanes = default_lanes(graph_structure["mapping"].n_nodes) if graph_structure is not None else 1
        self.lanes = [Rollout(model, mask, graph_structure, high_interest_region, remesh_every, True, warmup_eager)
                      for _ in range(max(1, int(lanes)))]
        for k, ro in enumerate(self.lanes):
            if ro.salt is not None:
                ro.salt.fill_(k << 32)
        self.streams = None

    @property
    def graph(self):
        return self.lanes[0].graph

    @property
    def launches_per_replay(self):
        return self.lanes[0].launches_per_replay

    def warm(self, x, cl):
        """Eager warm-ups and the capture of every lane (one after the other, on the calling stream)."""
        for ro in self.lanes:
            while ro.graph is None:
                ro(x, cl)
        torch.cuda.current_stream().synchronize()

    @torch.no_grad()
    def predict_many(self, xs, concat_layers, out=None):
        if self.streams is None:
            self.streams = [torch.cuda.Stream() for _ in self.lanes]
        main = torch.cuda.current_stream()
        for i, (x, cl) in enumerate(zip(xs, concat_layers)):
            k = i % len(self.lanes)
            ro, st = self.lanes[k], self.streams[k]
            if ro.graph is None:                    # still warming up / capturing: on the calling stream, nothing in flight
                for s2 in self.streams:
                    main.wait_stream(s2)
                y = ro(x, cl)
                if out is None:
                    out = torch.empty((len(xs),) + tuple(y.shape), dtype=y.dtype, device=y.device)
                out[i].copy_(y)
                continue
            if out is None:
                out = torch.empty((len(xs),) + tuple(ro.out.shape), dtype=ro.out.dtype, device=ro.out.device)
            st.wait_stream(main)                    # x, cl (and `out`) were produced on the calling stream
            with torch.cuda.stream(st):
                y = ro(x, cl)                       # static-buffer copies + the replay, in this lane's stream order
                out[i].copy_(y, non_blocking=True)
            x.record_stream(st)
            cl.record_stream(st)
        for st in self.streams:
            main.wait_stream(st)
        return out


@torch.no_grad()
def predict(model, xs, concat_layers, mask, graph_structure=None, high_interest_region=None, remesh_every=1,
            use_cuda_graph=False, rollout=None):
    """xs: list of [T_in, H, W, c] device tensors (one per launch date), concat_layers: matching list of
    [T_out, H, W, 1].  Returns [n, T_out, H, W, 1] on the device.  ``rollout``: a Rollout or a RolloutPool to reuse."""
    if not xs:
        return None
    if isinstance(rollout, RolloutPool):
        return rollout.predict_many(xs, concat_layers)
    ro = rollout or Rollout(model, mask, graph_structure, high_interest_region, remesh_every, use_cuda_graph)
    out = None
    for i, (x, cl) in enumerate(zip(xs, concat_layers)):
        y = ro(x, cl)
        if out is None:
            out = torch.empty((len(xs),) + tuple(y.shape), dtype=y.dtype, device=y.device)
        out[i].copy_(y)
    return out


@torch.no_grad()
def predict_sharded(model, load_sample, n_dates, mask, rank=0, world=1, process_group=None, **kw):
    """Shard ``n_dates`` launch dates over ``world`` ranks; ``load_sample(d)`` returns (x, concat_layers) on this
    rank's device.  Every rank returns the full [n_dates, T_out, H, W, 1] tensor in launch-date order after ONE
    all_gather (ranks that got one date fewer forecast a padding date, trimmed here)."""
    mine = shard_launch_dates(n_dates, rank, world, pad=True)
    pairs = [load_sample(d) for d in mine]
    local = predict(model, [p[0] for p in pairs], [p[1] for p in pairs], mask, **kw)
    if world == 1:
        return local
    import torch.distributed as dist
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=process_group)
    # rank r holds dates r, r + world, ...: interleave back to launch-date order, drop the padding
    full = torch.stack(parts, dim=1)                      # [per, world, ...]
    return full.reshape((-1,) + tuple(local.shape[1:]))[:n_dates]
