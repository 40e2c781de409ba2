"""Shared synthetic inputs and comparison helpers for the parity tests."""
import numpy as np
import torch


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def dist_from_05(arr):
    return abs(abs(arr - 0.5) - 0.5)


def blob_frames(rng, T, H, W, c=1, density=0.9, scale=0.05):
    """Mostly-low field with sparse high pixels: gives a quadtree with cells of every size."""
    x = rng.random((T, H, W, c)).astype(np.float32)
    hi = rng.random((T, H, W)) > density
    x[..., 0] = np.where(hi, x[..., 0], x[..., 0] * scale)
    return x


def moving_blob(rng, T, H, W, size=12):
    """SURVEY 8(d) C1-style sample: N(0, 0.05) noise + a sparse blob translating 1 px / frame."""
    x = rng.normal(0, 0.05, (T, H, W, 1)).astype(np.float32)
    u = rng.random((size, size)).astype(np.float32)
    blob = (u > 0.6) * u
    for t in range(T):
        r, c = 5 + t, 7 + t
        x[t, r:r + size, c:c + size, 0] += blob[: max(0, min(size, H - r)), : max(0, min(size, W - c))]
    return x


def circ_close(a, b, tol=1e-5):
    d = (a - b).abs()
    return bool((torch.minimum(d, 1 - d) < tol).all())


def attrs_close(a, b, tol=1e-5):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.dim() == 1:
        return torch.allclose(a, b, atol=tol)
    return circ_close(a[:, 0], b[:, 0], tol) and torch.allclose(a[:, 1], b[:, 1], atol=tol)
