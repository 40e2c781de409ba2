"""B200 mirror of the reference's ``model/graph_functions.py`` (same names, arguments, return
layouts and error behaviour), running on hand-written sm_100a kernels through the C ABI.

What changed underneath: the pixel -> node assignment is a :class:`Mesh` (label image
+ per-node pixel lists on the device) instead of a dense one-hot ``[N, P]`` matrix; the quadtree split,
label numbering, adjacency and edge attributes are computed on the GPU with one host read-back of the
two data-dependent sizes (N, E) per graph.  ``Mesh`` is accepted wherever the reference takes
``mapping`` (``flatten``, ``unflatten``, ``graph_structure['mapping']``); a dense matrix is still
accepted and converted once.  There is no CPU path: every function here needs a CUDA device.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from .utils import add_positional_encoding

CONDITIONS = [
    "max_larger_than",
    "max_smaller_than",
    "min_larger_than",
    "min_smaller_than",
]


def _device(device=None):
    if device is not None:
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.QmpError("quadtree_mpnnlstm_b200 runs on CUDA devices only (no CPU fallback)")
        return device
    if not torch.cuda.is_available():
        raise _lib.QmpError("quadtree_mpnnlstm_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


_mask_cache = {}


def _mask_u8(mask, device):
    """numpy / torch bool mask -> uint8 device tensor [H*W]; cached per array object."""
    if mask is None:
        return None
    if isinstance(mask, torch.Tensor):
        return mask.to(device=device, dtype=torch.uint8).reshape(-1).contiguous()
    key = (id(mask), str(device))
    hit = _mask_cache.get(key)
    if hit is not None and hit[0] is mask:
        return hit[1]
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))).to(device).reshape(-1)
    if len(_mask_cache) > 16:
        _mask_cache.clear()
    _mask_cache[key] = (mask, t)
    return t


class Mesh:
    """Pixel -> node assignment on the device; stands in for the reference's ``mapping`` matrix
    (graph_functions.py:555-587 / :649).  ``to_dense()`` materialises the ``[N, P]`` matrix."""

    def __init__(self, labels, n_nodes, npix, pix_ptr, pix_idx, image_shape, kind):
        self.labels, self.n_nodes, self.npix = labels, int(n_nodes), npix
        self.pix_ptr, self.pix_idx = pix_ptr, pix_idx
        self.image_shape, self.kind = tuple(image_shape), kind

    @property
    def shape(self):
        return (self.n_nodes, self.labels.numel())

    @property
    def device(self):
        return self.labels.device

    def to(self, *a, **k):
        return self

    def to_dense(self):
        lab = self.labels.long()
        m = torch.zeros(self.n_nodes, lab.numel(), dtype=torch.float32, device=lab.device)
        idx = torch.nonzero(lab >= 0).squeeze(1)
        m[lab[idx], idx] = 1.0
        return m

    @staticmethod
    def from_labels(labels, n_nodes, image_shape, kind="quadtree"):
        """Generic labels (any pixel sets).  Host-logic helper for the static-graph constructors and
        for callers that pass a dense matrix; the pixel lists come from one stable device sort."""
        lab = labels.reshape(-1).to(torch.int32).contiguous()
        valid = lab >= 0
        order = torch.argsort(torch.where(valid, lab, torch.full_like(lab, n_nodes)).long(), stable=True)
        counts = torch.bincount(lab[valid].long(), minlength=n_nodes)
        ptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=lab.device)
        ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        n_valid = int(valid.sum().item())
        return Mesh(lab, n_nodes, counts.to(torch.float32), ptr, order[:n_valid].to(torch.int32).contiguous(),
                    image_shape, kind)

    _dense_cache = {}

    @staticmethod
    def from_dense(mapping, image_shape):
        key = (mapping.data_ptr(), tuple(mapping.shape))
        hit = Mesh._dense_cache.get(key)
        if hit is not None and hit[0] is mapping:
            return hit[1]
        if mapping.is_sparse:
            mapping = mapping.to_dense()
        has = mapping.sum(0) > 0
        lab = torch.where(has, mapping.argmax(0), torch.full((mapping.shape[1],), -1, device=mapping.device))
        mesh = Mesh.from_labels(lab, mapping.shape[0], image_shape)
        if len(Mesh._dense_cache) > 8:
            Mesh._dense_cache.clear()
        Mesh._dense_cache[key] = (mapping, mesh)
        return mesh


_pixelwise_cache = {}


def pixelwise_mesh(mask, image_shape, device):
    """Mesh with one node per unmasked pixel (graph_functions.py:511-525)."""
    key = (id(mask), tuple(image_shape), str(device))
    hit = _pixelwise_cache.get(key)
    if hit is not None and hit[0] is mask:
        return hit[1]
    h, w = image_shape
    P = h * w
    m8 = _mask_u8(mask, device)
    i32 = dict(dtype=torch.int32, device=device)
    labels, pix_ptr, pix_idx = torch.empty(P, **i32), torch.empty(P + 1, **i32), torch.empty(P, **i32)
    npix = torch.empty(P, dtype=torch.float32, device=device)
    n_dev, keep, rank, bs = torch.zeros(1, **i32), torch.empty(P, **i32), torch.empty(P, **i32), torch.empty(P // 1024 + 2, **i32)
    _lib.call("qmp_mesh_pixelwise", m8, P, labels, pix_ptr, pix_idx, npix, n_dev, keep, rank, bs)
    n = int(n_dev.item())
    mesh = Mesh(labels, n, npix[:n], pix_ptr[:n + 1], pix_idx[:n], image_shape, "pixelwise")
    if len(_pixelwise_cache) > 8:
        _pixelwise_cache.clear()
    _pixelwise_cache[key] = (mask, mesh)
    return mesh


class Graph:
    """Mesh state holder (graph_functions.py:23-33).  ``pyg`` keeps the attribute names the reference's
    driver uses (``x``, ``edge_index``, ``edge_attr``, ``to``)."""

    class _Data:
        def __init__(self, edge_index=None, edge_attr=None, **kwargs):
            self.x, self.edge_index, self.edge_attr = None, edge_index, edge_attr
            for k, v in kwargs.items():
                setattr(self, k, v)

        def to(self, device, *a, **k):
            return self

    def __init__(self, edge_index, edge_attr, **kwargs):
        self.pyg = Graph._Data(edge_index=edge_index, edge_attr=edge_attr, **kwargs)
        self.mapping = None
        self.n_pixels_per_node = None
        self.hidden = None
        self.cell = None


# --------------------------------------------------------------------------- pool / unpool
class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, mesh):
        B, H, W, C = img.shape
        img = img.contiguous()
        out = torch.empty(B, mesh.n_nodes, C, dtype=torch.float32, device=img.device)
        _lib.call("qmp_segment_sum", img, B, H * W, C, mesh.pix_ptr, mesh.pix_idx, mesh.npix, mesh.n_nodes, None, 1, int(mesh.kind == "pixelwise"), out)
        ctx.mesh, ctx.shape = mesh, (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, C = ctx.shape
        mesh = ctx.mesh
        g = g.contiguous()
        dimg = torch.empty(B, H, W, C, dtype=torch.float32, device=g.device)
        _lib.call("qmp_gather_by_label", g, B, H * W, C, mesh.n_nodes, mesh.labels, mesh.npix, 1, 0.0, dimg)
        return dimg, None


class _PooledByBuild(torch.autograd.Function):
    """``data`` [B, N, c + 1] = (pooled ``img`` | node size) exactly as qmp_quadtree_graph wrote it, made differentiable in
    ``img``: the forward is the identity on the build's output (bit-identical to ``cat([_Pool(img), size])``), the backward is
    ``_Pool``'s (gather by label with the division); the size column carries no gradient."""

    @staticmethod
    def forward(ctx, img, data, mesh):
        ctx.mesh, ctx.shape = mesh, tuple(img.shape)
        return data.view_as(data)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C = ctx.shape
        mesh = ctx.mesh
        gc = g[..., :C].contiguous()
        dimg = torch.empty(B, H, W, C, dtype=torch.float32, device=g.device)
        _lib.call("qmp_gather_by_label", gc, B, H * W, C, mesh.n_nodes, mesh.labels, mesh.npix, 1, 0.0, dimg)
        return dimg, None, None


class _Unpool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, mesh, fill):
        B, N, C = data.shape
        data = data.contiguous()
        H, W = mesh.image_shape
        img = torch.empty(B, H, W, C, dtype=torch.float32, device=data.device)
        _lib.call("qmp_gather_by_label", data, B, H * W, C, N, mesh.labels, mesh.npix, 0, float(fill), img)
        ctx.mesh, ctx.shape = mesh, (B, N, C)
        return img

    @staticmethod
    def backward(ctx, g):
        B, N, C = ctx.shape
        mesh = ctx.mesh
        H, W = mesh.image_shape
        g = g.contiguous()
        dd = torch.empty(B, N, C, dtype=torch.float32, device=g.device)
        _lib.call("qmp_segment_sum", g, B, H * W, C, mesh.pix_ptr, mesh.pix_idx, mesh.npix, N, None, 0, int(mesh.kind == "pixelwise"), dd)
        return dd, None, None


class _Regrid(torch.autograd.Function):
    """``flatten(unflatten(a, mesh_s), mesh_d)`` for up to two node tensors [B, N_s, C] (hidden and cell state) as ONE launch each
    way (csrc/pool.cu regrid_kernel) -- the [B, H, W, C] images of model/seq2seq.py:440-476 are never materialised.  Values and
    gradients are bit-identical to the two-step path (same gather, same defined summation order)."""

    @staticmethod
    def forward(ctx, mesh_s, mesh_d, a, b):
        B, Ns, C = a.shape
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        H, W = mesh_s.image_shape
        Nd = mesh_d.n_nodes
        out_a = torch.empty(B, Nd, C, dtype=torch.float32, device=a.device)
        out_b = torch.empty(B, Nd, C, dtype=torch.float32, device=a.device) if b is not None else None
        _lib.call("qmp_regrid", a, b, B, H * W, C, Ns, mesh_s.labels, mesh_s.npix, 0, 0.0, mesh_d.pix_ptr, mesh_d.pix_idx,
                  mesh_d.npix, Nd, 1, out_a, out_b)
        ctx.meshes, ctx.shape, ctx.two = (mesh_s, mesh_d), (B, Ns, C), b is not None
        if b is None:
            return out_a
        return out_a, out_b

    @staticmethod
    def backward(ctx, ga, gb=None):
        mesh_s, mesh_d = ctx.meshes
        B, Ns, C = ctx.shape
        H, W = mesh_s.image_shape
        Nd = mesh_d.n_nodes
        if ga is None and gb is None:
            return None, None, None, None
        zero = lambda: torch.zeros(B, Nd, C, dtype=torch.float32, device=mesh_s.labels.device)
        ga = ga.contiguous() if ga is not None else zero()
        if ctx.two:
            gb = gb.contiguous() if gb is not None else zero()
        da = torch.empty(B, Ns, C, dtype=torch.float32, device=ga.device)
        db = torch.empty(B, Ns, C, dtype=torch.float32, device=ga.device) if ctx.two else None
        # pool backward (gather by D's labels with the division) + unpool backward (sum over S's pixel lists)
        _lib.call("qmp_regrid", ga, gb if ctx.two else None, B, H * W, C, Nd, mesh_d.labels, mesh_d.npix, 1, 0.0, mesh_s.pix_ptr,
                  mesh_s.pix_idx, mesh_s.npix, Ns, 0, da, db)
        return None, None, da, db


def regrid(mesh_s, mesh_d, a, b=None, image_shape=None, n_pixels_per_node=None):
    """Move node tensors ``a`` (and ``b``) [..., N_s, C] from mesh ``mesh_s`` onto mesh ``mesh_d``: the reference's
    unflatten -> flatten pair of do_remesh (model/seq2seq.py:440-476) in one pass.  Returns tensors [..., N_d, C]."""
    C = a.shape[-1]
    if not (isinstance(mesh_s, Mesh) and isinstance(mesh_d, Mesh)) or (b is not None and b.shape != a.shape):
        f = lambda t: flatten(unflatten(t, mesh_s, image_shape), mesh_d, n_pixels_per_node)      # dense mappings
        return f(a) if b is None else (f(a), f(b))
    lead = a.shape[:-2]
    a3 = a.float().reshape(-1, a.shape[-2], C)
    if b is None:
        return _Regrid.apply(mesh_s, mesh_d, a3, None).reshape(*lead, mesh_d.n_nodes, C)
    oa, ob = _Regrid.apply(mesh_s, mesh_d, a3, b.float().reshape(-1, b.shape[-2], C))
    return oa.reshape(*lead, mesh_d.n_nodes, C), ob.reshape(*lead, mesh_d.n_nodes, C)


def _as_mesh(mapping, image_shape, mask, device):
    if isinstance(mapping, Mesh):
        return mapping
    if mapping is None:
        if mask is None:
            mask = _no_mask(image_shape)
        return pixelwise_mesh(mask, image_shape, device)
    if isinstance(mapping, torch.Tensor):
        return Mesh.from_dense(mapping.to(device), image_shape)
    raise TypeError(f"unsupported mapping type {type(mapping)}")


_no_mask_cache = {}


def _no_mask(image_shape):
    m = _no_mask_cache.get(tuple(image_shape))
    if m is None:
        m = _no_mask_cache[tuple(image_shape)] = np.zeros(tuple(image_shape), dtype=bool)
    return m


def flatten(img, mapping, n_pixels_per_node, mask=None):
    """Mean-pool pixels into mesh nodes: (n_samples, w, h, c) -> (n_samples, N, c)
    (graph_functions.py:391-419; ``mapping is None`` = pixel-wise ``img[:, ~mask, :]``, :383-389)."""
    assert len(img.shape) == 4, f'array should be 4-dimensional (n_samples, w, h, c); got {img.shape}'
    mesh = _as_mesh(mapping, img.shape[1:3], mask, img.device)
    return _Pool.apply(img.float(), mesh)


def unflatten(data, mapping, image_shape, mask=None):
    """Nodes back to an image: (..., N, c) -> (..., w, h, c) (graph_functions.py:451-468).  Quadtree
    meshes write 0 on masked pixels, the pixel-wise mesh writes NaN when a mask is given (:460-468)."""
    if mapping is None:
        mesh = _as_mesh(None, image_shape, mask, data.device)
        fill = float("nan") if mask is not None else 0.0
        return _Unpool.apply(data.float().unsqueeze(0), mesh, fill).squeeze(0)
    mesh = _as_mesh(mapping, image_shape, mask, data.device)
    lead = data.shape[:-2]
    flat = data.float().reshape(-1, data.shape[-2], data.shape[-1])
    img = _Unpool.apply(flat, mesh, 0.0)
    return img.reshape(*lead, *image_shape, data.shape[-1])


# --------------------------------------------------------------------------- graph construction
def _apply_transform(transform_func, crit):
    """The reference applies ``transform_func`` to the padded numpy frame (graph_functions.py:194).
    Element-wise Python expressions work on the device tensor unchanged; anything that insists on
    numpy makes one host round trip of this single frame."""
    if transform_func is None:
        return crit
    try:
        out = transform_func(crit)
        if isinstance(out, torch.Tensor) and out.shape == crit.shape:
            return out.float().contiguous()
    except (TypeError, RuntimeError, AttributeError):
        pass
    out = transform_func(crit.cpu().numpy())
    return torch.as_tensor(np.ascontiguousarray(out), dtype=torch.float32).to(crit.device)


def _edge_buffers(e_cap, device):
    ei = torch.empty(2, e_cap, dtype=torch.int64, device=device)
    s32 = torch.empty(e_cap, dtype=torch.int32, device=device)
    d32 = torch.empty(e_cap, dtype=torch.int32, device=device)
    return ei, s32, d32


_pixelwise_topology = {}


def _pixelwise_edges(mesh, mask, h, w, c_pos_source, use_edge_attrs, resolution):
    """edge_index / edge_attrs of the pixel-wise mesh.  They depend only on the mask, the image shape and the
    resolution (node positions are the positional-encoding planes), so they are built once per mask and
    reused by every later sample -- no per-sample host read-back."""
    key = (id(mask), h, w, bool(use_edge_attrs), float(resolution), str(mesh.labels.device))
    hit = _pixelwise_topology.get(key)
    if hit is not None and hit[0] is mask:
        return hit[1], hit[2]
    dev = mesh.labels.device
    N, P = mesh.n_nodes, h * w
    i32 = dict(dtype=torch.int32, device=dev)
    e_cap = 4 * N
    ei, s32, d32 = _edge_buffers(e_cap, dev)
    n_edges, count, offset, bs = torch.zeros(1, **i32), torch.empty(P, **i32), torch.empty(P, **i32), torch.empty(P // 1024 + 2, **i32)
    _lib.call("qmp_adjacency_pixelwise", mesh.labels, h, w, ei[0], ei[1], s32, d32, n_edges, count, offset, bs)
    edge_attrs = None
    if use_edge_attrs:
        # node positions = pooled ii / jj planes of the positional encoding (graph_functions.py:519)
        pos = _Pool.apply(add_positional_encoding(torch.zeros(1, h, w, 1, device=dev))[..., 1:].contiguous(), mesh)
        edge_attrs = torch.empty(e_cap, 2, dtype=torch.float32, device=dev)
        _lib.call("qmp_edge_attrs", s32, d32, e_cap, n_edges, pos[0, :, 0:], pos[0, :, 1:], 2, w, h, float(resolution),
                  1, edge_attrs)
    E = int(n_edges.item())
    edge_index = ei[:, :E].contiguous()
    if edge_attrs is not None:
        edge_attrs = edge_attrs[:E].contiguous()
    if len(_pixelwise_topology) > 8:
        _pixelwise_topology.clear()
    _pixelwise_topology[key] = (mask, edge_index, edge_attrs)
    return edge_index, edge_attrs


def image_to_graph_pixelwise(img, mask=None, use_edge_attrs=True, resolution=0.25):
    """image_to_graph() if each pixel is treated as a node (graph_functions.py:506-539).  The last two
    channels of ``img`` must be the positional encoding, as in the reference (:519)."""
    if mask is None:
        raise TypeError("image_to_graph_pixelwise needs a mask (the reference evaluates ~mask)")
    n, h, w, c = img.shape
    dev = img.device
    mesh = pixelwise_mesh(mask, (h, w), dev)
    N = mesh.n_nodes
    data = _Pool.apply(img.float(), mesh)
    edge_index, edge_attrs = _pixelwise_edges(mesh, mask, h, w, c, use_edge_attrs, resolution)
    sizes = torch.full((n, N, 1), float(resolution) ** 2, dtype=torch.float32, device=dev)
    data = torch.cat([data, sizes], -1)
    return dict(edge_index=edge_index, edge_attrs=edge_attrs, data=data, graph_nodes=torch.arange(N),
                mapping=None, n_pixels_per_node=mesh.npix, labels=mesh.labels.view(h, w), mesh=mesh)


ONE_LAUNCH_BUILD = os.environ.get("QMP_ONE_LAUNCH_BUILD", "1") != "0"     # 0: the per-kernel entry points (cross-check path)
_gb_arenas = {}


def _gb_arena(h, w, S, T, C, dev):
    """Arena of qmp_quadtree_graph (scratch + the results at capacity) and its pinned count buffer, reused per shape and stream."""
    key = (h, w, S, T, C, dev.index, _lib.stream_ptr())
    hit = _gb_arenas.get(key)
    if hit is None:
        nbytes = int(_lib.lib().qmp_quadtree_graph_scratch_bytes(h, w, S, T, C))
        if len(_gb_arenas) > 8:
            _gb_arenas.clear()
        host = torch.zeros(4, dtype=torch.int32)
        hit = (torch.empty(nbytes, dtype=torch.uint8, device=dev), host.pin_memory() if torch.cuda.is_available() else host)
        _gb_arenas[key] = hit
    return hit


def _quadtree_graph_one_launch(img, crit, m8, h8, S, cond, thresh, use_edge_attrs, resolution, max_grid_size):
    """image_to_graph's quadtree branch on csrc/graph_build.cu: one cooperative launch builds labels, pixel lists, pooled node
    features, edges, edge attributes and both CSRs inside an arena; one read-back of (N, E, NaN count); one launch copies the
    compacted result into exact-size buffers.  The CSR is registered with graph_csr, so the conv modules never rebuild it."""
    from . import graph_csr
    n, h, w, c = img.shape
    dev = img.device
    P = h * w
    two = 1 if use_edge_attrs else 0
    arena, counts_host = _gb_arena(h, w, S, n, c, dev)
    _lib.call("qmp_quadtree_graph", img.detach(), n, h, w, c, crit, m8, h8, S, cond, float(thresh), float(resolution), two,
              counts_host.data_ptr(), arena)
    N, E, nans = counts_host.tolist()[:3]
    if nans:
        raise ValueError(f'Found NaNs in image data {nans} / {img.numel()}')
    r4 = lambda v: (v + 3) & ~3
    wd = 2 if two else 1
    isz = (P, N + 1, P, E, E, N + 1, E, E, N + 1, E, E)
    fsz = (N, n * N * (c + 1), E * wd, E * wd)
    ipack = torch.empty(sum(r4(v) for v in isz), dtype=torch.int32, device=dev)
    fpack = torch.empty(sum(r4(v) for v in fsz), dtype=torch.float32, device=dev)
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    _lib.call("qmp_quadtree_graph_export", arena, h, w, S, n, c, two, N, E, ipack, fpack, edge_index)
    iv, off = [], 0
    for v in isz:
        iv.append(ipack[off:off + v])
        off += r4(v)
    labels, pix_ptr, pix_idx, src32, dst32, in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin = iv
    fv, off = [], 0
    for v in fsz:
        fv.append(fpack[off:off + v])
        off += r4(v)
    npix, data, edge_attrs, edge_attr_in = fv
    data = data.view(n, N, c + 1)
    if two:
        edge_attrs, edge_attr_in = edge_attrs.view(E, 2), edge_attr_in.view(E, 2)
    mesh = Mesh(labels, N, npix, pix_ptr, pix_idx, (h, w), "quadtree")
    if img.requires_grad:            # differentiable pooling (seq2seq.py:440-476 regrids state): the build already pooled the
        data = _PooledByBuild.apply(img, data, mesh)      # frames (same summation order as _Pool); only the backward is added
    edge_index._qmp_trusted = True
    graph_csr.register(graph_csr.GraphCSR.from_parts(edge_index, edge_attrs, N, src32, dst32, in_ptr, in_src, in_eid, out_ptr, out_dst,
                                                     out_kin, edge_attr_in))
    return dict(edge_index=edge_index, edge_attrs=edge_attrs, data=data, graph_nodes=np.arange(N), mapping=mesh,
                n_pixels_per_node=npix, labels=labels.view(h, w))


def image_to_graph(img, thresh=0.05, max_grid_size=64, mask=None, high_interest_region=None, transform_func=None,
                   condition='max_larger_than', use_edge_attrs=True, resolution=0.25):
    """Convert an image (n_samples, height, width, channels) to its graph representation using quadtree
    decomposition (graph_functions.py:590-681).  Returns the reference's dict -- ``edge_index`` int64
    [2, E] in the reference's edge order, ``edge_attrs``, ``data`` [n, N, c+1] (last column = node size),
    ``graph_nodes``, ``mapping`` (a :class:`Mesh`), ``n_pixels_per_node`` -- plus ``labels`` [H, W]."""
    assert len(img.shape) == 4, f'array should be 4-dimensional (n_samples, w, h, c); got {img.shape}'
    assert max_grid_size & (max_grid_size - 1) == 0
    assert condition in CONDITIONS
    if not _lib.on_device(img):
        raise _lib.QmpError("image_to_graph: CUDA tensor required (no CPU fallback)")
    img = img.float().contiguous()
    if thresh == -np.inf:
        # (a CUDA-graph capture cannot read back; the eager warm-up steps before a capture do check)
        if not _lib.capturing():
            nans = int(torch.isnan(img.detach()).sum().item())
            if nans:
                raise ValueError(f'Found NaNs in image data {nans} / {img.numel()}')
        return image_to_graph_pixelwise(img, mask, use_edge_attrs=use_edge_attrs, resolution=resolution)

    n, h, w, c = img.shape
    dev = img.device
    P = h * w
    S = int(max_grid_size)
    n_pad, m_pad = -(h // -S) * S, -(w // -S) * S
    f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)

    crit = torch.empty(n_pad, m_pad, **f32)
    _lib.call("qmp_frame_max_pad", img.detach(), n, h, w, c, n_pad, m_pad, crit)
    crit = _apply_transform(transform_func, crit)

    m8, h8 = _mask_u8(mask, dev), _mask_u8(high_interest_region, dev)
    if S <= 64 and c >= 2 and ONE_LAUNCH_BUILD:
        return _quadtree_graph_one_launch(img, crit, m8, h8, S, CONDITIONS.index(condition), thresh, use_edge_attrs, resolution,
                                          max_grid_size)
    nan_flag = torch.isnan(img.detach()).sum()              # (the one-launch build counts NaNs itself)
    cells = _lib.lib().qmp_quadtree_pyramid_cells(h, w, S)
    labels, rect, npix = torch.empty(P, **i32), torch.empty(P, 4, **i32), torch.empty(P, **f32)
    counts = torch.zeros(2, **i32)                       # [n_nodes, n_edges]
    split, cnt = torch.empty(cells, dtype=torch.uint8, device=dev), torch.empty(cells, **i32)
    nb = (n_pad // S) * (m_pad // S)
    top = 2 * ((n_pad + 63) // 64) * ((m_pad + 63) // 64) + 2
    base_off, top_f, top_b = torch.empty(nb, **i32), torch.empty(top, **f32), torch.empty(top, dtype=torch.uint8, device=dev)
    _lib.call("qmp_quadtree_labels", crit, m8, h8, h, w, S, CONDITIONS.index(condition), float(thresh), labels, rect,
              npix, counts[0:], split, cnt, base_off, top_f, top_b)

    pix_ptr, pix_idx, tmp, bs = torch.empty(P + 1, **i32), torch.empty(P, **i32), torch.empty(P + 2, **i32), torch.empty(P // 1024 + 4, **i32)
    _lib.call("qmp_mesh_pixels_from_rects", labels, h, w, rect, npix, counts[0:], P, pix_ptr, pix_idx, tmp, bs)

    # pooled node features with node capacity P (true N still on the device)
    data_cap = torch.empty(n, P, c, **f32)
    _lib.call("qmp_segment_sum", img.detach(), n, P, c, pix_ptr, pix_idx, npix, P, counts[0:], 1, 0, data_cap)

    e_cap = 4 * P
    ei, s32, d32 = _edge_buffers(e_cap, dev)
    table_cap = 1 << max(10, int(np.ceil(np.log2(8 * P))))
    keys, vals = torch.empty(table_cap, dtype=torch.int64, device=dev), torch.empty(table_cap, **i32)
    emit, count, offset = torch.empty(P, dtype=torch.uint8, device=dev), torch.empty(P, **i32), torch.empty(P, **i32)
    _lib.call("qmp_adjacency_quadtree", labels, h, w, ei[0], ei[1], s32, d32, counts[1:], keys, vals, table_cap,
              emit, count, offset, bs)
    two = 1 if use_edge_attrs else 0
    edge_attrs = torch.empty((e_cap, 2) if two else (e_cap,), **f32)
    _lib.call("qmp_edge_attrs", s32, d32, e_cap, counts[1:], data_cap[0, :, c - 2:], data_cap[0, :, c - 1:], c, w, h,
              float(resolution), two, edge_attrs)

    # the one host read-back per graph: N, E and the NaN count
    N, E, nans = [int(v) for v in torch.cat([counts, nan_flag.reshape(1).to(torch.int32)]).tolist()]
    if nans:
        raise ValueError(f'Found NaNs in image data {nans} / {img.numel()}')
    mesh = Mesh(labels, N, npix[:N], pix_ptr[:N + 1], pix_idx, (h, w), "quadtree")
    if img.requires_grad:
        data = _Pool.apply(img, mesh)                    # differentiable pooling (seq2seq.py:440-476 regrids state)
    else:
        data = data_cap[:, :N] if n == 1 else data_cap[:, :N].contiguous()
    cell_sizes = (mesh.npix / ((max_grid_size / 2) ** 2)).reshape(1, N, 1).expand(n, N, 1)   # :665-666
    data = torch.cat([data, cell_sizes], -1)
    edge_index = ei[:, :E].contiguous()
    edge_index._qmp_trusted = True          # emitted by qmp_adjacency_quadtree: endpoints are in [0, N) by construction
    return dict(edge_index=edge_index, edge_attrs=edge_attrs[:E].contiguous(), data=data,
                graph_nodes=np.arange(N), mapping=mesh, n_pixels_per_node=mesh.npix, labels=labels.view(h, w))


def create_static_heterogeneous_graph(image_shape, max_grid_size, mask, high_interest_region=None, use_edge_attrs=True,
                                      resolution=0.25, device=None):
    """Static mesh that is denser near the mask / high-interest edges (graph_functions.py:683-699)."""
    arr = torch.zeros(size=(1, *image_shape, 1), device=_device(device))
    arr = add_positional_encoding(arr)
    graph_structure = image_to_graph(arr, thresh=np.inf, max_grid_size=max_grid_size, mask=mask,
                                     high_interest_region=high_interest_region, use_edge_attrs=use_edge_attrs,
                                     resolution=resolution)
    del graph_structure['data']
    return graph_structure


def create_static_homogeneous_graph(image_shape, max_grid_size, mask, use_edge_attrs=True, resolution=0.25, device=None):
    """Static mesh of one resolution (graph_functions.py:707-737): the heterogeneous mesh without a mask,
    minus the nodes whose pixels are all masked, renumbered 0..n-1.  Host-side list logic in the
    reference; here a handful of device tensor ops, run once."""
    gs = create_static_heterogeneous_graph(image_shape, max_grid_size, None, None, use_edge_attrs, resolution, device)
    mesh = gs['mapping']
    dev = mesh.labels.device
    lab = mesh.labels.long()
    keep_px = (_mask_u8(mask, dev) == 0).long()
    unmasked = torch.zeros(mesh.n_nodes, dtype=torch.int64, device=dev).index_add_(0, lab, keep_px)
    alive = unmasked > 0
    renum = torch.where(alive, torch.cumsum(alive.long(), 0) - 1, torch.full_like(unmasked, -1))
    ei = gs['edge_index']
    ekeep = alive[ei[0]] & alive[ei[1]]
    n_new = int(alive.sum().item())
    gs['edge_index'] = renum[ei[:, ekeep]].contiguous()
    gs['edge_attrs'] = gs['edge_attrs'][ekeep].contiguous()
    gs['graph_nodes'] = np.arange(n_new)
    new_mesh = Mesh.from_labels(renum[lab], n_new, image_shape)
    gs['mapping'] = new_mesh
    gs['n_pixels_per_node'] = new_mesh.npix
    gs['labels'] = new_mesh.labels.view(*image_shape)
    return gs


def plot_contours(ax, labels):
    """Plot cell contours for a label image (graph_functions.py:99-113); plotting helper, host side."""
    labels = labels.cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
    for i in range(labels.shape[0]):
        for j in range(labels.shape[1]):
            if j + 1 < labels.shape[1] and labels[i][j] != labels[i][j + 1]:
                ax.plot([j + 0.5, j + 0.5], [i - 0.5, i + 0.5], c='k', lw=0.5)
            if i + 1 < labels.shape[0] and labels[i][j] != labels[i + 1][j]:
                ax.plot([j - 0.5, j + 0.5], [i + 0.5, i + 0.5], c='k', lw=0.5)
