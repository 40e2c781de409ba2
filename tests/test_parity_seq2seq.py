"""GPU parity: the full Seq2Seq driver vs the CPU oracle -- per-step forecasts within 1e-4 relative
(the north-star tolerance), gradients close, graph structure identical."""
import numpy as np
import pytest
import torch

from helpers import dist_from_05, moving_blob, rel_err

STEP_TOL = 1e-4


def _run_pair(be, kw, x, y, cl, mask, hir=None, graph_fn=None, remesh_every=1, grads=True):
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    from oracle.seq2seq_ref import Seq2Seq as OSeq
    torch.manual_seed(5)
    ref = OSeq(**kw)
    gpu = be.dev(q.Seq2Seq(**kw, device=be.device))
    gpu.load_state_dict(ref.state_dict())
    with torch.no_grad():          # exercise peepholes / biases / norms away from their init values
        for m in (ref, gpu):
            gen = torch.Generator().manual_seed(9)
            for k, p in m.named_parameters():
                if ".w_c_" in k or ".b_" in k or "norm" in k:
                    p.add_(0.1 * torch.randn(p.shape, generator=gen).to(p.device))
    ref.eval(); gpu.eval()
    gs_a = graph_fn(G, None) if graph_fn else None
    gs_b = graph_fn(q, be.device) if graph_fn else None
    oa, ma = ref(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask,
                 high_interest_region=hir, graph_structure=gs_a, remesh_every=remesh_every)
    ob, mb = gpu(be.dev(torch.from_numpy(x)), be.dev(torch.from_numpy(y)), be.dev(torch.from_numpy(cl)),
                 teacher_forcing_ratio=0, mask=mask, high_interest_region=hir, graph_structure=gs_b,
                 remesh_every=remesh_every)
    assert len(oa) == len(ob)
    H, W = x.shape[1:3]
    for t, (a, b) in enumerate(zip(oa, ob)):
        assert a.shape == b.shape, f"step {t}: node count {tuple(b.shape)} vs oracle {tuple(a.shape)}"
        assert rel_err(b, a) < STEP_TOL, f"step {t}: rel err {rel_err(b, a)}"
        ia = G.unpool(a, ma[t], (H, W), mask)
        ib = q.unflatten(b, mb[t], (H, W), mask).cpu()
        assert torch.equal(torch.isnan(ia), torch.isnan(ib))
        assert rel_err(torch.nan_to_num(ib), torch.nan_to_num(ia)) < STEP_TOL
    if grads:
        la = sum((o ** 2).mean() for o in oa)
        lb = sum((o ** 2).mean() for o in ob)
        la.backward(); lb.backward()
        worst = 0.0
        for (k, pa), (_, pb) in zip(ref.named_parameters(), gpu.named_parameters()):
            ga = pa.grad if pa.grad is not None else torch.zeros_like(pa)
            gb = pb.grad if pb.grad is not None else torch.zeros_like(pb)
            diff = float((ga - gb.cpu()).abs().max())
            err = diff / max(float(ga.abs().max()), 1e-4)
            worst = max(worst, err)
            assert err < 2e-3 or diff < 2e-5, f"grad {k}: rel {err} abs {diff}"
    return oa, ob


def _data(seed, T_in, T_out, H, W, c=1):
    rng = np.random.default_rng(seed)
    x = moving_blob(rng, T_in, H, W)
    if c > 1:
        x = np.concatenate([x, rng.random((T_in, H, W, c - 1)).astype(np.float32)], -1)
    y = moving_blob(rng, T_out, H, W)
    cl = rng.random((T_out, H, W, 1)).astype(np.float32)
    return x, y, cl


@pytest.mark.parametrize("path", ["tc", "ffma", "modular"])
def test_ice_like_pixelwise_transformer(be, path, monkeypatch):
    """configs[1] in miniature: pixel-wise static mesh, TransformerConv, 1 layer, 3 encoder conv layers; on the
    single-launch fused kernels (the default for this shape) and on the modular kernels."""
    import quadtree_mpnnlstm_b200.fused as FZ
    monkeypatch.setattr(FZ, "ENABLED", path != "modular")
    monkeypatch.setattr(FZ, "TC_FWD", path == "tc")
    monkeypatch.setattr(FZ, "TC_BWD", path == "tc")
    H, W = 24, 40
    x, y, cl = _data(1, 4, 6, H, W, c=5)
    rr, cc = np.mgrid[0:H, 0:W]
    mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    kw = dict(hidden_size=32, dropout=0.0, thresh=-np.inf, input_timesteps=4, input_features=8, output_timesteps=6,
              n_layers=1, n_conv_layers=3, convolution_type="TransformerConv", transform_func=dist_from_05)
    hits = FZ.ACC_HITS
    _run_pair(be, kw, x, y, cl, mask)
    if path != "modular":       # weight / gate-parameter gradients accumulate in place across the timesteps (fused._GradAccum)
        assert FZ.ACC_HITS > hits, "the shared gradient accumulators must be the path that runs"


def test_mnist_like_dynamic_quadtree_cheb(be):
    """configs[0] in miniature: dynamic quadtree rebuilt every decoder step, ChebConv, 2 layers."""
    H, W = 32, 32
    x, y, cl = _data(2, 4, 5, H, W)
    cl[:] = 0
    mask = np.zeros((H, W), bool)
    kw = dict(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=4, input_features=4, output_timesteps=5,
              n_layers=2, n_conv_layers=2)
    _run_pair(be, kw, x, y, cl, mask)


def test_dynamic_quadtree_transformer_with_mask(be):
    """configs[2] in miniature: TransformerConv on a quadtree mesh rebuilt every step, mask + HIR + transform."""
    H, W = 24, 40
    x, y, cl = _data(3, 3, 4, H, W, c=2)
    x[..., 0] = np.clip(x[..., 0] * 2, 0, 1)
    rng = np.random.default_rng(0)
    mask = rng.random((H, W)) > 0.85
    hir = rng.random((H, W)) > 0.97
    kw = dict(hidden_size=16, dropout=0.0, thresh=0.15, input_timesteps=3, input_features=5, output_timesteps=4,
              n_layers=1, n_conv_layers=2, convolution_type="TransformerConv", transform_func=dist_from_05)
    _run_pair(be, kw, x, y, cl, mask, hir=hir)


def test_static_heterogeneous_mesh_inference(be):
    """configs[4] in miniature: preset heterogeneous mesh (max cell 4), no_grad rollout."""
    H, W = 30, 44
    x, y, cl = _data(4, 3, 5, H, W, c=3)
    rng = np.random.default_rng(1)
    mask = rng.random((H, W)) > 0.8
    kw = dict(hidden_size=32, dropout=0.0, thresh=-np.inf, input_timesteps=3, input_features=6, output_timesteps=5,
              n_layers=1, n_conv_layers=3, convolution_type="TransformerConv")
    fn = lambda mod, dev: (mod.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12)
                           if dev is None else
                           mod.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=dev))
    with torch.no_grad():
        _run_pair(be, kw, x, y, cl, mask, graph_fn=fn, grads=False)


def test_gcn_quadtree_remesh_every_2(be):
    H, W = 32, 32
    x, y, cl = _data(5, 3, 4, H, W)
    mask = np.zeros((H, W), bool)
    kw = dict(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=3, input_features=4, output_timesteps=4,
              n_layers=2, n_conv_layers=2, convolution_type="GCNConv")
    _run_pair(be, kw, x, y, cl, mask, remesh_every=2)


def test_multi_head_transformer_seq2seq(be):
    """the driver with convolution_type='MHTransformerConv' (model/model.py:26-37, 52; edge attributes on,
    seq2seq.py:244) on a dynamic quadtree mesh."""
    H, W = 24, 32
    x, y, cl = _data(6, 3, 3, H, W, c=2)
    rng = np.random.default_rng(2)
    mask = rng.random((H, W)) > 0.85
    kw = dict(hidden_size=8, dropout=0.0, thresh=0.15, input_timesteps=3, input_features=5, output_timesteps=3,
              n_layers=1, n_conv_layers=2, convolution_type="MHTransformerConv", transform_func=dist_from_05)
    _run_pair(be, kw, x, y, cl, mask)


@pytest.mark.parametrize("kind", ["GATConv"])
def test_gat_seq2seq(be, kind):
    """the driver with convolution_type='GATConv' (model/model.py:43, 55) on a dynamic quadtree mesh with edge
    attributes: forecasts and gradients against the oracle driver.  ('GATv2Conv' cannot run through the reference driver: it is
    built with edge_dim=2, model/model.py:56, but seq2seq.py:243 hands it 1-D edge weights; it is covered at conv and cell level.)"""
    H, W = 24, 32
    x, y, cl = _data(7, 3, 3, H, W, c=2)
    mask = np.random.default_rng(3).random((H, W)) > 0.85
    kw = dict(hidden_size=8, dropout=0.0, thresh=0.15, input_timesteps=3, input_features=5, output_timesteps=3,
              n_layers=1, n_conv_layers=1, convolution_type=kind, transform_func=dist_from_05)
    _run_pair(be, kw, x, y, cl, mask)


def test_state_dict_keys_match_reference_layout(be):
    import quadtree_mpnnlstm_b200 as q
    m = q.Seq2Seq(hidden_size=8, dropout=0.1, thresh=-np.inf, input_features=8, n_layers=1, n_conv_layers=2,
                  convolution_type="TransformerConv")
    keys = set(m.state_dict().keys())
    for k in ("encoder.rnns.0.conv_x_i.convolutions.0.lin_key.weight", "encoder.rnns.0.conv_h_o.convolutions.1.lin_edge.weight",
              "encoder.rnns.0.w_c_i", "encoder.rnns.0.b_c", "encoder.norm_h.weight", "decoder.norm_o.bias",
              "decoder.fc_out1.lin_skip.bias", "decoder.fc_out2.lin_query.weight"):
        assert k in keys, k
    m2 = q.Seq2Seq(hidden_size=8, dropout=0.1, thresh=0.1, n_layers=2)
    keys2 = set(m2.state_dict().keys())
    assert "decoder.rnns.1.conv_h_f.convolutions.0.lins.2.weight" in keys2 and "decoder.fc_out1.bias" in keys2
