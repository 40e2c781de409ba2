"""Rollout inference lanes (infer.RolloutPool): launch dates forecast by several captured graphs replayed concurrently on
their own streams must equal the dates forecast one after the other without a graph (reference: NextFramePredictorS2S.predict,
model/mpnnlstm.py:402-440, a sequential loop over launch dates)."""
import numpy as np
import pytest
import torch

from helpers import dist_from_05, moving_blob

pytestmark = pytest.mark.gpu


def _setup(dev, static_mesh):
    import quadtree_mpnnlstm_b200 as q
    H, W, T_in, T_out = 40, 56, 3, 5
    rng = np.random.default_rng(5)
    rr, cc = np.mgrid[0:H, 0:W]
    mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    kw = dict(hidden_size=32, dropout=0.1, thresh=-np.inf, input_timesteps=T_in, input_features=8, output_timesteps=T_out,
              n_layers=1, n_conv_layers=2, convolution_type="TransformerConv", transform_func=dist_from_05)
    torch.manual_seed(3)
    model = q.Seq2Seq(**kw, device=dev).to(dev).eval()
    gs = None
    if static_mesh:
        gs = q.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=dev)
    dates = []
    for d in range(7):
        x = np.concatenate([moving_blob(rng, T_in, H, W, size=8), rng.random((T_in, H, W, 4)).astype(np.float32)], -1)
        cl = rng.random((T_out, H, W, 1)).astype(np.float32)
        dates.append((torch.from_numpy(x).to(dev), torch.from_numpy(cl).to(dev)))
    return model, mask, gs, dates


@pytest.mark.parametrize("static_mesh", [True, False])
def test_pool_lanes_match_sequential_eager_rollouts(static_mesh):
    from quadtree_mpnnlstm_b200.infer import Rollout, RolloutPool, predict
    dev = torch.device("cuda")
    model, mask, gs, dates = _setup(dev, static_mesh)
    xs, cls = [d[0] for d in dates], [d[1] for d in dates]
    want = predict(model, xs, cls, mask, graph_structure=gs, use_cuda_graph=False).clone()
    pool = RolloutPool(model, mask, graph_structure=gs, lanes=3)
    got_cold = predict(model, xs, cls, mask, rollout=pool).clone()          # includes every lane's warm-ups and capture
    assert pool.lanes[0].graph is not None            # dates 0, 3, 6: two eager warm-ups, then the capture
    pool.warm(xs[0], cls[0])
    assert all(ro.graph is not None for ro in pool.lanes), "every lane must replay a captured graph"
    got = predict(model, xs, cls, mask, rollout=pool)                      # pure replays, three dates in flight
    torch.cuda.synchronize()
    for name, g in (("cold", got_cold), ("replay", got)):
        assert g.shape == want.shape
        assert torch.equal(torch.isnan(g), torch.isnan(want)), name
        assert torch.allclose(torch.nan_to_num(g), torch.nan_to_num(want), atol=1e-6, rtol=1e-5), (name, (torch.nan_to_num(g) - torch.nan_to_num(want)).abs().max())
    # a second pass over the same dates reuses the static buffers: same forecasts
    again = predict(model, xs, cls, mask, rollout=pool)
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(again), torch.nan_to_num(got))


def test_default_lanes():
    from quadtree_mpnnlstm_b200.infer import default_lanes
    assert default_lanes(4066) == 8 and default_lanes(47200) == 1 and default_lanes(100) == 8 and default_lanes(9000) == 4
