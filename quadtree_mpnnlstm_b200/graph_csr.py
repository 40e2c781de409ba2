"""Device CSR views of an ``edge_index`` (built once per graph by ``qmp_csr_from_edge_index``).

The reference passes ``edge_index[2, E]`` / ``edge_attr`` tensors to every PyG conv call
(model/model.py:96); the modules here keep that calling convention and look the CSR up in a small
cache keyed by the tensors' storage, so the per-call cost is a dictionary probe.
"""
from __future__ import annotations

import collections

import torch

from . import _lib


class GraphCSR:
    def __init__(self, edge_index, edge_attr, n_nodes, validate=True):
        assert edge_index.dim() == 2 and edge_index.shape[0] == 2, "edge_index must be [2, E]"
        if not _lib.on_device(edge_index):
            raise _lib.QmpError("edge_index must live on a CUDA device (no CPU fallback)")
        dev = edge_index.device
        ei = edge_index.to(torch.int64).contiguous()
        E, N = ei.shape[1], int(n_nodes)
        self.n_nodes, self.n_edges, self.device = N, E, dev
        i32 = dict(dtype=torch.int32, device=dev)
        self.src, self.dst = torch.empty(E, **i32), torch.empty(E, **i32)
        self.in_ptr, self.in_src, self.in_eid = torch.empty(N + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
        self.out_ptr, self.out_dst, self.out_kin = torch.empty(N + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
        bad = torch.zeros(1, **i32)
        tmp, bs = torch.empty(N + 2, **i32), torch.empty((N + 1) // 1024 + 4, **i32)
        eid_out, kin_of_edge = torch.empty(E, **i32), torch.empty(E, **i32)
        _lib.call("qmp_csr_from_edge_index", ei, E, N, self.src, self.dst, self.in_ptr, self.in_src, self.in_eid,
                  self.out_ptr, self.out_dst, self.out_kin, bad, tmp, bs, eid_out, kin_of_edge)
        if validate and int(bad.item()):
            raise IndexError(f"edge_index has {int(bad.item())} endpoints outside [0, {N})")
        self.edge_attr_in = None       # edge payload permuted into in-CSR order
        self.edge_dim = 0
        if edge_attr is not None:
            ea = edge_attr.detach().float().contiguous()
            width = 1 if ea.dim() == 1 else ea.shape[1]
            self.edge_dim = width
            self.edge_attr_in = torch.empty_like(ea)
            _lib.call("qmp_gather_rows", ea, self.in_eid, E, width, self.edge_attr_in)
        self._norm = {}
        self._keepalive = (edge_index, edge_attr)

    @classmethod
    def from_parts(cls, edge_index, edge_attr, n_nodes, src, dst, in_ptr, in_src, in_eid, out_ptr, out_dst, out_kin, edge_attr_in):
        """A CSR the graph build already produced on the device (qmp_quadtree_graph): no kernels, no validation read-back."""
        self = cls.__new__(cls)
        self.n_nodes, self.n_edges, self.device = int(n_nodes), int(edge_index.shape[1]), edge_index.device
        self.src, self.dst = src, dst
        self.in_ptr, self.in_src, self.in_eid = in_ptr, in_src, in_eid
        self.out_ptr, self.out_dst, self.out_kin = out_ptr, out_dst, out_kin
        self.edge_attr_in = edge_attr_in
        self.edge_dim = 0 if edge_attr is None else (1 if edge_attr.dim() == 1 else edge_attr.shape[1])
        self._norm = {}
        self._keepalive = (edge_index, edge_attr)
        return self

    def norm(self, mode):
        """Per-edge normalisation (in-CSR order) for 'gcn' (mode 0) or 'cheb' (mode 1); cached per graph."""
        val = self._norm.get(mode)
        if val is None:
            w = self.edge_attr_in
            if w is not None and w.dim() != 1:
                raise ValueError("GCNConv / ChebConv take a 1-D edge weight; got edge attributes of shape "
                                 f"{tuple(w.shape)}")
            val = torch.empty(self.n_edges, dtype=torch.float32, device=self.device)
            dis = torch.empty(self.n_nodes, dtype=torch.float32, device=self.device)
            _lib.call("qmp_edge_norm", 0 if mode == "gcn" else 1, self.n_nodes, self.in_ptr, self.in_src, self.out_ptr,
                      self.out_dst, self.out_kin, w, dis, val)
            self._norm[mode] = val
        return val


_cache = collections.OrderedDict()       # least recently used first
_MAX = 32


def _key(edge_index, edge_attr, n_nodes):
    return (edge_index.data_ptr(), tuple(edge_index.shape), int(n_nodes), edge_index._version,
            None if edge_attr is None else (edge_attr.data_ptr(), tuple(edge_attr.shape), edge_attr._version))


def register(csr):
    """Put a prebuilt GraphCSR where get_csr() finds it for its (edge_index, edge_attr) tensors."""
    ei, ea = csr._keepalive
    key = _key(ei, ea, csr.n_nodes)
    _cache[key] = csr
    _cache.move_to_end(key)
    while len(_cache) > _MAX:
        _cache.popitem(last=False)
    return csr


def get_csr(edge_index, edge_attr, n_nodes, validate=True):
    if isinstance(edge_index, GraphCSR):
        return edge_index
    key = _key(edge_index, edge_attr, n_nodes)
    hit = _cache.get(key)
    if hit is not None and hit._keepalive[0] is edge_index and hit._keepalive[1] is edge_attr:
        _cache.move_to_end(key)
        return hit
    # an edge_index this package's own graph build emitted needs no bounds check (a host read-back per mesh: with a mesh per
    # forecast step that is one pipeline drain per frame)
    csr = GraphCSR(edge_index, edge_attr, n_nodes, validate=validate and not getattr(edge_index, "_qmp_trusted", False))
    _cache[key] = csr                    # a stale entry under the same key (same storage, new tensor object) is replaced
    _cache.move_to_end(key)
    while len(_cache) > _MAX:
        _cache.popitem(last=False)
    return csr
