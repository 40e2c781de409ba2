"""Mirror of the reference's ``model/utils.py`` (the functions the hot path and its callers import)."""
from __future__ import annotations

import datetime

import numpy as np
import torch

from . import _lib


def get_n_params(model):
    """Number of parameters in a PyTorch model (model/utils.py:19-27)."""
    return sum(int(np.prod(p.size())) for p in model.parameters())


class _AddPos(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, h, w, c = x.shape
        x = x.contiguous()
        out = torch.empty(n, h, w, c + 2, dtype=torch.float32, device=x.device)
        _lib.call("qmp_add_positional_encoding", x, n, h, w, c, out)
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx, g):
        return g[..., :ctx.c]


def add_positional_encoding(x):
    """(n_samples, w, h, c) -> (n_samples, w, h, c+2): appends ii = col / W and jj = row / H
    (model/utils.py:30-52).  Device tensors go through the CUDA kernel; numpy arrays (host-side data
    preparation in the reference's notebooks) are handled with numpy."""
    assert len(x.shape) == 4, f'array should be 4-dimensional (n_samples, w, h, c); got {x.shape}'
    if isinstance(x, torch.Tensor):
        if not _lib.on_device(x):
            raise _lib.QmpError("add_positional_encoding: CUDA tensor required (no CPU fallback)")
        return _AddPos.apply(x.float())
    n, rows, cols, _ = x.shape
    ii = np.broadcast_to(np.arange(cols)[None, :] / cols, (rows, cols))
    jj = np.broadcast_to(np.arange(rows)[:, None] / rows, (rows, cols))
    pos = np.broadcast_to(np.stack([ii, jj], -1)[None], (n, rows, cols, 2)).astype(x.dtype)
    return np.concatenate((x, pos), axis=-1)


def normalize(arr):
    """Per-variable min-max over all other axes (model/utils.py:70-73)."""
    min_ = np.min(arr, (0, 2, 3, 4))[:, None, None, None]
    max_ = np.max(arr, (0, 2, 3, 4))[:, None, None, None]
    return (arr - min_) / (max_ - min_)


def int_to_datetime(x):
    return datetime.datetime.fromtimestamp(x / 1e9)


def round_to_day(dt):
    return datetime.datetime(*dt.timetuple()[:3])
