"""GPU parity: graph build, pooling and adjacency vs the CPU oracle (bit-exact labels / edge_index)."""
import numpy as np
import pytest
import torch

from helpers import attrs_close, blob_frames, dist_from_05, rel_err


CASES = [
    # H, W, max_size, thresh, cond, mask, hir, transform
    (64, 64, 64, 0.5, "max_larger_than", False, False, False),
    (37, 53, 16, 0.5, "max_smaller_than", True, False, False),
    (20, 70, 64, 0.5, "min_larger_than", False, False, True),
    (64, 64, 4, 0.5, "min_smaller_than", True, True, False),
    (50, 90, 32, 0.3, "max_larger_than", False, True, True),
    (33, 33, 64, 0.5, "max_smaller_than", True, True, False),
    (100, 130, 128, 0.5, "max_larger_than", True, False, False),   # max_size > 64: level-up kernels
    (7, 9, 1, 0.5, "max_larger_than", True, False, False),         # max_size 1: every pixel its own base cell
    (229, 361, 64, 0.15, "max_larger_than", True, True, True),     # ice grid
]


def _inputs(case, seed):
    H, W, S, thresh, cond, use_mask, use_hir, use_tf = case
    rng = np.random.default_rng(seed)
    x = blob_frames(rng, 3, H, W, c=2)
    mask = (rng.random((H, W)) > 0.9) if use_mask else None
    hir = (rng.random((H, W)) > 0.98) if use_hir else None
    return x, mask, hir, (dist_from_05 if use_tf else None)


@pytest.mark.parametrize("idx", range(len(CASES)))
@pytest.mark.parametrize("use_edge_attrs", [True, False])
def test_image_to_graph_matches_oracle(be, idx, use_edge_attrs):
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    case = CASES[idx]
    H, W, S, thresh, cond, *_ = case
    x, mask, hir, tf = _inputs(case, 100 + idx)
    xc = G.add_positional_encoding(torch.from_numpy(x))
    xg = q.add_positional_encoding(be.dev(torch.from_numpy(x)))
    assert torch.equal(xc, xg.cpu()), "positional encoding differs"
    ref = G.image_to_graph(xc, thresh=thresh, max_grid_size=S, mask=mask, high_interest_region=hir,
                           transform_func=tf, condition=cond, use_edge_attrs=use_edge_attrs)
    got = q.image_to_graph(xg, thresh=thresh, max_grid_size=S, mask=mask, high_interest_region=hir,
                           transform_func=tf, condition=cond, use_edge_attrs=use_edge_attrs)
    lab = got["labels"].cpu().long().numpy()
    assert np.array_equal(lab, ref["labels"]), f"labels differ at {np.argwhere(lab != ref['labels'])[:5]}"
    assert got["edge_index"].dtype == torch.int64
    assert torch.equal(got["edge_index"].cpu(), ref["edge_index"]), "edge_index differs (order or content)"
    assert torch.equal(got["n_pixels_per_node"].cpu(), ref["n_pixels_per_node"])
    assert torch.equal(got["data"].cpu(), ref["data"]), f"node data not bit-identical: {rel_err(got['data'], ref['data'])}"
    assert attrs_close(got["edge_attrs"], ref["edge_attrs"])
    assert got["mapping"].shape == (ref["mapping"].n_nodes, H * W)


@pytest.mark.parametrize("shape", [(40, 60), (229, 361), (5, 5)])
def test_pixelwise_graph_matches_oracle(be, shape):
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    H, W = shape
    rng = np.random.default_rng(7)
    x = rng.random((2, H, W, 3)).astype(np.float32)
    rr, cc = np.mgrid[0:H, 0:W]
    mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    ref = G.image_to_graph(G.add_positional_encoding(torch.from_numpy(x)), thresh=-np.inf, mask=mask)
    got = q.image_to_graph(q.add_positional_encoding(be.dev(torch.from_numpy(x))), thresh=-np.inf, mask=mask)
    assert torch.equal(got["edge_index"].cpu(), ref["edge_index"])
    assert torch.equal(got["data"].cpu(), ref["data"])
    assert torch.allclose(got["edge_attrs"].cpu(), ref["edge_attrs"], atol=1e-6)
    assert got["mapping"] is None
    img_r = G.unpool(ref["data"][0], None, (H, W), mask)
    img_g = q.unflatten(got["data"][0], None, (H, W), mask).cpu()
    assert torch.equal(torch.isnan(img_r), torch.isnan(img_g))
    assert torch.equal(torch.nan_to_num(img_r), torch.nan_to_num(img_g))


def test_static_graphs_match_oracle(be):
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    mask = np.zeros((50, 70), bool)
    mask[:10, :30] = True
    mask[30:, 50:] = True
    mask[20, 20] = True
    for fn in ("create_static_heterogeneous_graph", "create_static_homogeneous_graph"):
        a = getattr(G, fn)((50, 70), 4, mask, use_edge_attrs=True, resolution=1 / 12)
        b = getattr(q, fn)((50, 70), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=be.device)
        assert torch.equal(b["edge_index"].cpu(), a["edge_index"]), fn
        assert attrs_close(b["edge_attrs"], a["edge_attrs"]), fn
        assert torch.equal(b["n_pixels_per_node"].cpu(), a["n_pixels_per_node"]), fn
        assert torch.equal(b["mapping"].to_dense().cpu(), a["mapping"].dense()), fn
        assert "data" not in b


def test_pool_unpool_forward_backward(be):
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    rng = np.random.default_rng(3)
    H, W = 48, 40
    x = blob_frames(rng, 2, H, W, c=1)
    ref = G.image_to_graph(G.add_positional_encoding(torch.from_numpy(x)), thresh=0.5, max_grid_size=16)
    got = q.image_to_graph(q.add_positional_encoding(be.dev(torch.from_numpy(x))), thresh=0.5, max_grid_size=16)
    img = torch.from_numpy(rng.random((3, H, W, 5)).astype(np.float32))
    a = img.clone().requires_grad_(True)
    b = be.dev(img.clone()).requires_grad_(True)
    pa = G.pool(a, ref["mapping"], ref["n_pixels_per_node"])
    pb = q.flatten(b, got["mapping"], got["n_pixels_per_node"])
    assert torch.equal(pa, pb.cpu())
    w = torch.from_numpy(rng.random(tuple(pa.shape)).astype(np.float32))
    (pa * w).sum().backward()
    (pb * be.dev(w)).sum().backward()
    assert torch.allclose(a.grad, b.grad.cpu(), atol=1e-7)
    # unpool of [L, N, C] state tensors
    st = torch.from_numpy(rng.random((2, pa.shape[1], 4)).astype(np.float32))
    sa = st.clone().requires_grad_(True)
    sb = be.dev(st.clone()).requires_grad_(True)
    ua = G.unpool(sa, ref["mapping"], (H, W))
    ub = q.unflatten(sb, got["mapping"], (H, W))
    assert ua.shape == ub.shape and torch.equal(ua, ub.cpu())
    w2 = torch.from_numpy(rng.random(tuple(ua.shape)).astype(np.float32))
    (ua * w2).sum().backward()
    (ub * be.dev(w2)).sum().backward()
    assert torch.allclose(sa.grad, sb.grad.cpu(), rtol=1e-5, atol=1e-5)
    # a dense [N, P] matrix is still accepted where the reference takes `mapping`
    dense = got["mapping"].to_dense()
    pc = q.flatten(be.dev(img), dense, got["n_pixels_per_node"])
    assert torch.equal(pc.cpu(), pa.detach())


@pytest.mark.parametrize("shape,S,C", [((48, 40), 16, 8), ((229, 361), 64, 32), ((33, 70), 4, 4), ((64, 64), 64, 3)])
def test_regrid_is_bit_identical_to_unflatten_then_flatten(be, shape, S, C):
    """graph_functions.regrid (csrc/pool.cu regrid_kernel; model/seq2seq.py:440-476 do_remesh) against the two-step path it
    replaces: values AND gradients bit for bit, hidden and cell state in one launch, nodes of 1 ... S x S pixels."""
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.graph_functions import regrid
    if be.name == "cpu" and shape[0] > 100:
        pytest.skip("full grid on the GPU only")
    rng = np.random.default_rng(11)
    H, W = shape
    mask = rng.random((H, W)) > 0.9
    meshes = []
    for k in range(2):
        x = blob_frames(rng, 1, H, W, c=1)
        g = q.image_to_graph(q.add_positional_encoding(be.dev(torch.from_numpy(x))), thresh=0.3 + 0.3 * k, max_grid_size=S, mask=mask)
        meshes.append(g)
    ms, md = meshes[0]["mapping"], meshes[1]["mapping"]
    L = 2
    h0 = torch.from_numpy(rng.standard_normal((L, ms.n_nodes, C)).astype(np.float32))
    c0 = torch.from_numpy(rng.standard_normal((L, ms.n_nodes, C)).astype(np.float32))
    wa = be.dev(torch.from_numpy(rng.standard_normal((L, md.n_nodes, C)).astype(np.float32)))
    wb = be.dev(torch.from_numpy(rng.standard_normal((L, md.n_nodes, C)).astype(np.float32)))
    res = []
    for fused in (True, False):
        h = be.dev(h0.clone()).requires_grad_(True)
        c = be.dev(c0.clone()).requires_grad_(True)
        if fused:
            oh, oc = regrid(ms, md, h, c, (H, W), meshes[1]["n_pixels_per_node"])
        else:
            oh = q.flatten(q.unflatten(h, ms, (H, W)), md, meshes[1]["n_pixels_per_node"])
            oc = q.flatten(q.unflatten(c, ms, (H, W)), md, meshes[1]["n_pixels_per_node"])
        ((oh * wa).sum() + (oc * wb).sum()).backward()
        res.append((oh.detach().cpu(), oc.detach().cpu(), h.grad.cpu(), c.grad.cpu()))
    for a, b, name in zip(res[0], res[1], ("hidden", "cell", "d hidden", "d cell")):
        assert a.shape == b.shape and torch.equal(a, b), f"{name} differs: {rel_err(a, b)}"
    # one tensor only
    h = be.dev(h0.clone())
    assert torch.equal(regrid(ms, md, h, None, (H, W), meshes[1]["n_pixels_per_node"]).cpu(), res[1][0])


def test_nan_input_raises(be):
    import quadtree_mpnnlstm_b200 as q
    x = torch.zeros(1, 16, 16, 3, device=be.device)
    x[0, 3, 3, 0] = float("nan")
    with pytest.raises(ValueError):
        q.image_to_graph(x, thresh=0.5, max_grid_size=8)
    with pytest.raises(AssertionError):
        q.image_to_graph(torch.zeros(16, 16, 3, device=be.device))
    with pytest.raises(AssertionError):
        q.image_to_graph(torch.zeros(1, 16, 16, 3, device=be.device), max_grid_size=6)
    with pytest.raises(AssertionError):
        q.image_to_graph(torch.zeros(1, 16, 16, 3, device=be.device), condition="bogus")


def test_scan_kernel(be):
    from quadtree_mpnnlstm_b200 import _lib
    for n in (1, 5, 1024, 1025, 50000, 330676):
        v = torch.randint(0, 5, (n,), dtype=torch.int32, device=be.device)
        out = torch.empty_like(v)
        tot = torch.zeros(1, dtype=torch.int32, device=be.device)
        scratch = torch.empty(n // 1024 + 4, dtype=torch.int32, device=be.device)
        _lib.call("qmp_exclusive_scan_i32", v, out, n, tot, scratch)
        ref = torch.cumsum(v.long(), 0) - v.long()
        assert torch.equal(out.long(), ref), n
        assert int(tot) == int(v.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1, 3, 4, 7, 8])
def test_one_launch_build_matches_the_per_kernel_path(idx):
    """csrc/graph_build.cu (one cooperative launch: mesh + pooled features + edges + both CSRs) against the per-kernel entry
    points it replaces (quadtree.cu / pool.cu / edges.cu, then csr.cu from the emitted edge_index): every output bit for bit,
    pixel lists and CSR arrays included."""
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200 import graph_csr, graph_functions as gf
    case = CASES[idx]
    H, W, S, thresh, cond, *_ = case
    x, mask, hir, tf = _inputs(case, 300 + idx)
    xg = q.add_positional_encoding(torch.from_numpy(x).cuda())
    kw = dict(thresh=thresh, max_grid_size=S, mask=mask, high_interest_region=hir, transform_func=tf, condition=cond, use_edge_attrs=True)
    assert gf.ONE_LAUNCH_BUILD
    a = q.image_to_graph(xg, **kw)
    csr_a = graph_csr.get_csr(a["edge_index"], a["edge_attrs"], a["data"].shape[1])
    assert csr_a._keepalive[0] is a["edge_index"], "the build must have registered its CSR"
    gf.ONE_LAUNCH_BUILD = False
    try:
        b = q.image_to_graph(xg, **kw)
    finally:
        gf.ONE_LAUNCH_BUILD = True
    csr_b = graph_csr.GraphCSR(b["edge_index"], b["edge_attrs"], b["data"].shape[1])
    for k in ("edge_index", "edge_attrs", "data", "n_pixels_per_node", "labels"):
        assert torch.equal(a[k], b[k]), k
    ma, mb = a["mapping"], b["mapping"]
    assert ma.n_nodes == mb.n_nodes and torch.equal(ma.pix_ptr, mb.pix_ptr)
    nv = int(ma.pix_ptr[-1])
    assert torch.equal(ma.pix_idx[:nv], mb.pix_idx[:nv])
    for k in ("src", "dst", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_kin", "edge_attr_in"):
        assert torch.equal(getattr(csr_a, k), getattr(csr_b, k)), k
    # a second build on the same arena must not disturb the first result
    c = q.image_to_graph(xg.flip(0), **kw)
    assert torch.equal(a["edge_index"], b["edge_index"]) and torch.equal(csr_a.in_src, csr_b.in_src) and c["data"].shape[0] == xg.shape[0]
