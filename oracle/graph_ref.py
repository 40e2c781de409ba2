"""Oracle (CPU, numpy + torch) for the graph half of the hot path.  TEST INFRASTRUCTURE.

Restates reference ``model/graph_functions.py`` and ``model/utils.py:30-52``.
Every function cites the reference lines it follows.  Written for clarity, not
speed: recursion instead of the reference's explicit stack, an explicit model of
CPython's small-set slot order instead of relying on ``set`` itself (the CUDA
kernel implements the same slot model, and ``tests/test_oracle_graph.py`` checks
the model against the real ``set``).

Pinned against the unmodified reference in ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CONDITIONS = ("max_larger_than", "max_smaller_than", "min_larger_than", "min_smaller_than")


# --------------------------------------------------------------------------- utils
def add_positional_encoding(x):
    """Append ii = col / W and jj = row / H channels (reference model/utils.py:30-52).

    The reference builds both planes in float64 numpy and casts to ``x.dtype``.
    """
    assert x.ndim == 4, f"array should be 4-dimensional (n_samples, w, h, c); got {tuple(x.shape)}"
    n, rows, cols, _ = x.shape
    ii = np.broadcast_to(np.arange(cols, dtype=np.float64)[None, :] / cols, (rows, cols))
    jj = np.broadcast_to(np.arange(rows, dtype=np.float64)[:, None] / rows, (rows, cols))
    pos = np.broadcast_to(np.stack([ii, jj], -1)[None], (n, rows, cols, 2))
    if isinstance(x, torch.Tensor):
        pos_t = torch.from_numpy(np.array(pos)).to(dtype=x.dtype, device=x.device)
        return torch.cat((x, pos_t), dim=-1)
    return np.concatenate((x, pos.astype(x.dtype)), axis=-1)


# --------------------------------------------------------------------------- quadtree
def _criterion(window, condition, thresh):
    # graph_functions.py:228-235 -- compare the float32 extreme (promoted) with a float64 thresh
    if condition == "max_larger_than":
        return float(window.max()) > thresh
    if condition == "max_smaller_than":
        return float(window.max()) < thresh
    if condition == "min_larger_than":
        return float(window.min()) > thresh
    return float(window.min()) < thresh


def quadtree_labels(img, thresh=0.05, max_size=8, mask=None, high_interest_region=None,
                    transform_func=None, condition="max_larger_than"):
    """Per-pixel node label, -1 = masked (graph_functions.py:145-259, padding=0)."""
    assert max_size & (max_size - 1) == 0 and max_size > 0
    assert condition in CONDITIONS
    img = np.asarray(img)
    n, m = img.shape
    n_pad = -(n // -max_size) * max_size
    m_pad = -(m // -max_size) * max_size
    padded = np.pad(img, ((0, n_pad - n), (0, m_pad - m)), mode="edge")      # :190
    crit = transform_func(padded) if transform_func is not None else padded   # :194
    return quadtree_labels_on_padded(crit, n, m, thresh, max_size, mask, high_interest_region, condition)


def quadtree_labels_on_padded(crit, n, m, thresh, max_size, mask=None, high_interest_region=None,
                              condition="max_larger_than"):
    """The traversal of graph_functions.py:196-259 on the padded, transformed criterion image.

    Order of labels: the reference pushes base cells in raster order on a LIFO stack and
    pushes the four children as (x,y) (x+s,y) (x,y+s) (x+s,y+s); the pop order is therefore
    the reverse.  A depth-first recursion that visits base cells in reverse raster order and
    children in order (x+s,y+s) (x,y+s) (x+s,y) (x,y) numbers the leaves identically.
    """
    n_pad, m_pad = crit.shape
    labels = np.full((n_pad, m_pad), -1, dtype=np.int64)
    # the reference clips BOTH axes with shape[1] (= m_pad) (:222-225); numpy slicing then
    # clips rows to the array's own extent
    row_cap = min(m_pad, n_pad)
    counter = [0]

    def visit(x, y, size):
        if x >= n or y >= m:                                                   # :208
            return
        if size == 1:                                                          # :214-220
            if mask is not None and mask[x, y]:
                return
            labels[x, y] = counter[0]
            counter[0] += 1
            return
        r1 = min(x + size + 1, row_cap)
        c1 = min(y + size + 1, m_pad)
        if r1 <= x:
            raise ValueError("window is empty: image taller than its padded width (reference reads "
                             "out of bounds here)")
        split = _criterion(crit[x:r1, y:c1], condition, thresh)
        if mask is not None and mask[x:r1, y:c1].any():                        # :239 (numpy clips to n, m)
            split = True
        if high_interest_region is not None and high_interest_region[x:r1, y:c1].any():   # :241-244
            split = True
        if split:                                                              # :249-254, reversed pop order
            h = size // 2
            visit(x + h, y + h, h)
            visit(x, y + h, h)
            visit(x + h, y, h)
            visit(x, y, h)
        else:                                                                  # :256-257
            labels[x:x + size, y:y + size] = counter[0]
            counter[0] += 1

    base = [(i * max_size, j * max_size) for i in range(n_pad // max_size) for j in range(m_pad // max_size)]
    for (bx, by) in reversed(base):                                            # LIFO over :199-202
        visit(bx, by, max_size)
    return labels[:n, :m]


def pixelwise_labels(mask):
    """Raster rank of unmasked pixels, -1 on masked (graph_functions.py:511)."""
    keep = ~np.asarray(mask, dtype=bool)
    lab = np.cumsum(keep.ravel()) - 1
    lab = np.where(keep.ravel(), lab, -1).astype(np.int64)
    return lab.reshape(keep.shape)


# --------------------------------------------------------------------------- mapping / pooling
class LabelMap:
    """Pixel -> node assignment.  Stands in for the reference's dense one-hot ``mapping[N, P]``
    (graph_functions.py:555-587, densified at :649); ``dense()`` materialises that matrix."""

    def __init__(self, labels, n_nodes=None):
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int64).reshape(-1)
        self.labels = lab
        self.n_nodes = int(lab.max().item()) + 1 if n_nodes is None else int(n_nodes)

    def counts(self):
        valid = self.labels >= 0
        return torch.bincount(self.labels[valid], minlength=self.n_nodes).to(torch.float32)

    def dense(self):
        m = torch.zeros(self.n_nodes, self.labels.numel(), dtype=torch.float32)
        idx = torch.nonzero(self.labels >= 0).squeeze(1)
        m[self.labels[idx], idx] = 1.0
        return m


def lane_tree_segment_sum(x, ptr):
    """Segment sums of ``x`` [B, K, C] (segment v = rows ptr[v]:ptr[v+1]) in a DEFINED floating-point order, the one the
    device kernel uses (csrc/pool.cu segment_sum_kernel): 32 partial sums per segment -- partial l adds the segment's rows
    l, l + 32, l + 64, ... one after the other starting from 0 -- then the butterfly partial[l] += partial[l ^ off] for
    off = 16, 8, 4, 2, 1; the result is partial[0].  (The reference sums through a dense GEMM whose order is unspecified.)"""
    ptr = torch.as_tensor(ptr).long()
    B, K, C = x.shape
    N = ptr.numel() - 1
    cnt = ptr[1:] - ptr[:-1]
    out = torch.zeros(B, N, C, dtype=x.dtype)
    if N == 0 or K == 0:
        return out
    one = cnt == 1
    if bool(one.any()):                                  # a single row: 0 + x, every other partial is 0
        out[:, one] = x[:, ptr[:-1][one]] + 0.0
    multi = torch.nonzero(cnt > 1).squeeze(1)
    if multi.numel() == 0:
        return out
    slot = torch.full((N,), -1, dtype=torch.long)
    slot[multi] = torch.arange(multi.numel())
    seg = torch.repeat_interleave(torch.arange(N), cnt)
    rank = torch.arange(K) - ptr[seg]
    keep = slot[seg] >= 0
    rows, seg_m, lane, rnd = torch.nonzero(keep).squeeze(1), slot[seg[keep]], rank[keep] % 32, rank[keep] // 32
    part = torch.zeros(B, multi.numel(), 32, C, dtype=x.dtype)
    order = torch.argsort(rnd, stable=True)
    rows, seg_m, lane, rnd = rows[order], seg_m[order], lane[order], rnd[order]
    bounds = torch.searchsorted(rnd, torch.arange(int(rnd.max()) + 2))
    for j in range(bounds.numel() - 1):                  # round j: every (segment, lane) pair appears at most once
        a, b = int(bounds[j]), int(bounds[j + 1])
        if a == b:
            continue
        part[:, seg_m[a:b], lane[a:b]] = part[:, seg_m[a:b], lane[a:b]] + x[:, rows[a:b]]
    lanes = torch.arange(32)
    for off in (16, 8, 4, 2, 1):
        part = part + part[:, :, lanes ^ off]
    out[:, multi] = part[:, :, 0]
    return out


def pool(img, mapping, n_pixels_per_node, mask=None):
    """Mean-pool pixels into nodes: [n,H,W,c] -> [n,N,c] (graph_functions.py:391-419).

    ``mapping is None`` is the pixel-wise shortcut ``img[:, ~mask, :]`` (:383-389).
    The reference sums through a dense GEMM (order unspecified); here the sum over a node's pixels (raster order) runs
    in the defined order of :func:`lane_tree_segment_sum`, then divides by the pixel count like the reference does.
    """
    assert img.ndim == 4, f"array should be 4-dimensional (n_samples, w, h, c); got {tuple(img.shape)}"
    n, h, w, c = img.shape
    if mapping is None:
        if mask is not None:
            return img[:, ~torch.as_tensor(np.asarray(mask), dtype=torch.bool)]
        return img.reshape(n, -1, c)
    flat = img.reshape(n, h * w, c)
    lab = mapping.labels
    idx = torch.nonzero(lab >= 0).squeeze(1)
    order = idx[torch.argsort(lab[idx], stable=True)]      # pixels of node 0 in raster order, then node 1, ...
    cnt = torch.bincount(lab[idx], minlength=mapping.n_nodes)
    ptr = torch.zeros(mapping.n_nodes + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(cnt, 0)
    out = lane_tree_segment_sum(flat[:, order], ptr)
    return out / n_pixels_per_node.to(img.dtype)[None, :, None]


def unpool(data, mapping, image_shape, mask=None):
    """Nodes back to pixels: [..., N, c] -> [..., H, W, c] (graph_functions.py:451-468).

    Quadtree: masked pixels read 0 (a zero column of the one-hot matrix); pixel-wise
    (``mapping is None``): masked pixels read NaN (:460-468).
    """
    h, w = image_shape
    if mapping is None:
        n_nodes, c = data.shape
        if mask is None:
            return data.reshape(h, w, c)
        keep = ~torch.as_tensor(np.asarray(mask), dtype=torch.bool)
        img = torch.full((h, w, c), float("nan"), dtype=data.dtype)
        img[keep] = data
        return img
    lab = mapping.labels
    lead = data.shape[:-2]
    c = data.shape[-1]
    safe = lab.clamp(min=0)
    img = data[..., safe, :] * (lab >= 0).to(data.dtype)[:, None]
    return img.reshape(*lead, h, w, c)


# --------------------------------------------------------------------------- adjacency
def cpython_small_set_order(values):
    """Iteration order of ``s = set(); [s.add(v) for v in values]; s.discard(-1)`` for at most
    four small ints, as CPython produces it (what graph_functions.py:308-343 iterates).

    Model (CPython >= 3.7 Objects/setobject.c, set_add_entry): 8-slot open-addressed table
    (PySet_MINSIZE), first slot ``hash & 7`` with ``hash(v) = v`` for v >= 0 and
    ``hash(-1) = -2``; linear probing is disabled for an 8-slot table
    (``i + LINEAR_PROBES(9) <= mask(7)`` is never true); on a collision
    ``perturb >>= 5; i = (5*i + 1 + perturb) & 7`` with ``perturb`` the hash as an unsigned
    64-bit word.  Four inserts never trigger a resize (fill*5 < mask*3).  Removing -1
    leaves a dummy; iteration walks slots 0..7 and skips it.
    """
    assert len(values) <= 4
    slots = [None] * 8
    for v in values:
        v = int(v)
        h = -2 if v == -1 else v
        perturb = h & 0xFFFFFFFFFFFFFFFF
        i = h & 7
        while True:
            if slots[i] is None:
                slots[i] = v
                break
            if slots[i] == v:
                break
            perturb >>= 5
            i = (5 * i + 1 + perturb) & 7
    return [v for v in slots if v is not None and v != -1]


def adjacency(labels):
    """Directed edge list in the reference's order (graph_functions.py:291-345).

    Raster scan; per pixel insert up, down, left, right labels into a set, drop -1, iterate
    in set order; emit (node, nb) the first time ``nb`` is seen for ``node``.  Self-loops are
    kept (the removal is commented out at :329-333).  Returns int64 [2, E].
    """
    labels = np.asarray(labels)
    rows, cols = labels.shape
    seen = {}
    src, dst = [], []
    for i in range(rows):
        for j in range(cols):
            node = int(labels[i, j])
            if node == -1:
                continue
            mine = seen.setdefault(node, set())
            cand = []
            if i != 0:
                cand.append(labels[i - 1, j])
            if i != rows - 1:
                cand.append(labels[i + 1, j])
            if j != 0:
                cand.append(labels[i, j - 1])
            if j != cols - 1:
                cand.append(labels[i, j + 1])
            for nb in cpython_small_set_order(cand):
                if nb not in mine:
                    mine.add(nb)
                    src.append(node)
                    dst.append(nb)
    return np.array([src, dst], dtype=np.int64).reshape(2, -1)


def adjacency_pixelwise(labels):
    """graph_functions.py:471-493: per pixel the candidates are [row+1, row-1, col+1, col-1];
    pairs touching -1 are dropped.  Returns int64 [2, E]."""
    labels = np.asarray(labels)
    rows, cols = labels.shape
    nb = np.full((rows, cols, 4), -1, dtype=np.int64)
    nb[:-1, :, 0] = labels[1:, :]
    nb[1:, :, 1] = labels[:-1, :]
    nb[:, :-1, 2] = labels[:, 1:]
    nb[:, 1:, 3] = labels[:, :-1]
    src = np.repeat(labels.reshape(-1), 4)
    dst = nb.reshape(-1)
    keep = (src != -1) & (dst != -1)
    return np.stack([src[keep], dst[keep]])


def edge_dist(e0, e1, xx, yy):
    """graph_functions.py:358-363 (float32 torch arithmetic)."""
    return torch.sqrt((yy[e0] - yy[e1]) ** 2 + (xx[e0] - xx[e1]) ** 2)


def edge_angle(e0, e1, xx, yy):
    """graph_functions.py:365-370: atan2(dx, dy) wrapped to [0, 1)."""
    two_pi = 2 * np.pi
    return torch.atan2(xx[e0] - xx[e1], yy[e0] - yy[e1]) % two_pi / two_pi


def edge_attributes(edge_index, xx, yy, use_edge_attrs):
    e0, e1 = edge_index[0], edge_index[1]
    if use_edge_attrs:
        return torch.stack((edge_angle(e0, e1, xx, yy), edge_dist(e0, e1, xx, yy))).T   # :347-351
    return edge_dist(e0, e1, xx, yy)                                                    # :353


# --------------------------------------------------------------------------- image -> graph
def image_to_graph_pixelwise(img, mask, use_edge_attrs=True, resolution=0.25):
    """graph_functions.py:506-539.  Requires a mask (the reference evaluates ``~mask``)."""
    mask = np.asarray(mask, dtype=bool)
    h, w = img.shape[1:3]
    labels = pixelwise_labels(mask)
    n_nodes = int((~mask).sum())
    data = pool(img, None, None, mask)
    xx = data[0, :, -2] * w * resolution
    yy = data[0, :, -1] * h * resolution
    size = torch.ones(data.shape[0], n_nodes, dtype=data.dtype) * (resolution ** 2)
    data = torch.cat([data, size.unsqueeze(-1)], -1)
    edge_index = torch.from_numpy(adjacency_pixelwise(labels))
    # the reference's get_adj_pixelwise is not given `resolution` and ignores it (:528)
    edge_attrs = edge_attributes(edge_index, xx, yy, True) if use_edge_attrs else None
    return dict(edge_index=edge_index, edge_attrs=edge_attrs, data=data,
                graph_nodes=torch.arange(n_nodes), mapping=None,
                n_pixels_per_node=torch.ones(n_nodes), labels=labels)


def image_to_graph(img, thresh=0.05, max_grid_size=64, mask=None, high_interest_region=None,
                   transform_func=None, condition="max_larger_than", use_edge_attrs=True, resolution=0.25):
    """graph_functions.py:590-681.  Extra key ``labels`` (the reference dropped it) is for tests."""
    assert img.ndim == 4, f"array should be 4-dimensional (n_samples, w, h, c); got {tuple(img.shape)}"
    if torch.isnan(img).any():
        raise ValueError(f"Found NaNs in image data {int(torch.isnan(img).sum())} / {img.numel()}")
    if thresh == -np.inf:
        return image_to_graph_pixelwise(img, mask, use_edge_attrs=use_edge_attrs, resolution=resolution)
    n, h, w, _ = img.shape
    frame = img[..., 0].max(dim=0).values.detach().numpy()                     # :632-636
    labels = quadtree_labels(frame, thresh=thresh, max_size=max_grid_size, mask=mask,
                             high_interest_region=high_interest_region,
                             transform_func=transform_func, condition=condition)
    mapping = LabelMap(labels)
    npix = mapping.counts()
    if (npix == 0).any():
        raise ValueError("label gap: a node without pixels")
    data = pool(img, mapping, npix)
    if torch.isnan(data).any():
        raise ValueError(f"Found NaNs in graph data {int(torch.isnan(data).sum())} / {data.numel()}")
    xx = data[0, :, -2] * w * resolution                                       # :657
    yy = data[0, :, -1] * h * resolution
    size = (npix / ((max_grid_size / 2) ** 2)).repeat(n, 1)                    # :665-666
    data = torch.cat([data, size.unsqueeze(-1)], -1)
    edge_index = torch.from_numpy(adjacency(labels))
    edge_attrs = edge_attributes(edge_index, xx, yy, use_edge_attrs)
    return dict(edge_index=edge_index, edge_attrs=edge_attrs, data=data,
                graph_nodes=np.arange(mapping.n_nodes), mapping=mapping,
                n_pixels_per_node=npix, labels=labels)


def create_static_heterogeneous_graph(image_shape, max_grid_size, mask, high_interest_region=None,
                                      use_edge_attrs=True, resolution=0.25):
    """graph_functions.py:683-699: zeros + positional encoding, thresh=+inf, so only mask / HIR
    overlap splits cells."""
    arr = add_positional_encoding(torch.zeros(1, *image_shape, 1))
    gs = image_to_graph(arr, thresh=np.inf, max_grid_size=max_grid_size, mask=mask,
                        high_interest_region=high_interest_region,
                        use_edge_attrs=use_edge_attrs, resolution=resolution)
    del gs["data"]
    return gs


def create_static_homogeneous_graph(image_shape, max_grid_size, mask, use_edge_attrs=True, resolution=0.25):
    """graph_functions.py:707-737: heterogeneous graph without a mask, then delete the nodes whose
    pixels are all masked, drop their edges, renumber the survivors 0..n-1."""
    gs = create_static_heterogeneous_graph(image_shape, max_grid_size, None, None, use_edge_attrs, resolution)
    lab = gs["mapping"].labels
    keep_px = ~torch.as_tensor(np.asarray(mask), dtype=torch.bool).reshape(-1)
    n_old = gs["mapping"].n_nodes
    unmasked = torch.zeros(n_old, dtype=torch.int64).index_add(0, lab, keep_px.to(torch.int64))
    alive = unmasked > 0                                                       # :701-702 get_nan_nodes
    renum = torch.cumsum(alive.to(torch.int64), 0) - 1
    renum = torch.where(alive, renum, torch.full_like(renum, -1))
    ei = gs["edge_index"]
    ekeep = alive[ei[0]] & alive[ei[1]]                                        # :716
    gs["edge_index"] = renum[ei[:, ekeep]]
    gs["edge_attrs"] = gs["edge_attrs"][ekeep]
    gs["n_pixels_per_node"] = gs["n_pixels_per_node"][alive]
    gs["graph_nodes"] = np.arange(int(alive.sum()))
    gs["mapping"] = LabelMap(renum[lab], n_nodes=int(alive.sum()))             # rows of `mapping` kept (:722)
    gs["labels"] = gs["mapping"].labels.reshape(image_shape).numpy()
    return gs
