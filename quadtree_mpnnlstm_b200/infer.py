"""Rollout inference over many launch dates (reference: NextFramePredictorS2S.predict, model/mpnnlstm.py:402-440:
per launch date forward -> unflatten -> stack to [n_dates, T_out, H, W, 1]).  Launch dates are independent, so they
are sharded across ranks (one process per GPU) with NO collective on the data path; the forecasts are gathered once
at the end."""
from __future__ import annotations

import torch

from .graph_functions import unflatten
from .train import shard_launch_dates


@torch.no_grad()
def predict(model, xs, concat_layers, mask, graph_structure=None, high_interest_region=None, remesh_every=1):
    """xs: list of [T_in, H, W, c] device tensors (one per launch date), concat_layers: matching list of
    [T_out, H, W, 1].  Returns [n, T_out, H, W, 1] on the device."""
    out = []
    for x, cl in zip(xs, concat_layers):
        y_hat, maps = model(x, None, cl, teacher_forcing_ratio=0, mask=mask, high_interest_region=high_interest_region,
                            graph_structure=graph_structure, remesh_every=remesh_every)
        shape = tuple(x.shape[1:3])
        out.append(torch.stack([unflatten(y_hat[t], maps[t], shape, mask) for t in range(len(y_hat))]))
    return torch.stack(out) if out else None


@torch.no_grad()
def predict_sharded(model, load_sample, n_dates, mask, rank=0, world=1, process_group=None, **kw):
    """Shard ``n_dates`` launch dates over ``world`` ranks; ``load_sample(d)`` returns (x, concat_layers) on this
    rank's device.  Every rank returns the full [n_used, T_out, H, W, 1] tensor in launch-date order
    (n_used = n_dates rounded down to a multiple of world) after ONE all_gather."""
    mine = shard_launch_dates(n_dates, rank, world)
    pairs = [load_sample(d) for d in mine]
    local = predict(model, [p[0] for p in pairs], [p[1] for p in pairs], mask, **kw)
    if world == 1:
        return local
    import torch.distributed as dist
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=process_group)
    # rank r holds dates r, r + world, ...: interleave back to launch-date order
    full = torch.stack(parts, dim=1)                      # [per, world, ...]
    return full.reshape((-1,) + tuple(local.shape[1:]))
