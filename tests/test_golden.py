"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py, run in the build
container) versus (a) the CPU oracle -- this is what pins the oracle -- and (b) the product, on the CUDA
kernels (`-m gpu`) and on the test-only emulation.  Nothing here reads /root/reference.

Bars: labels / edge_index / node counts bit-exact; node data, edge attributes and forecasts within the float
tolerance written next to each check (north star: 1e-4 relative per forecast step).
"""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import attrs_close, dist_from_05, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA_TOL = 1e-5      # pooled node data: fp32 sums in a different order
STEP_TOL = 1e-4      # per-step forecasts (north star)

from golden.make_golden import GRAPH_CASES, SEQ_CASES  # noqa: E402  (case tables only; the generator is not run)


def _load(name):
    path = os.path.join(GOLD, name + ".npz")
    assert os.path.isfile(path), f"missing golden fixture {path}"
    return dict(np.load(path))


def test_fixture_inventory():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    want = {f"graph_{k}" for k in GRAPH_CASES} | {f"seq2seq_{k}" for k in SEQ_CASES} | {"graph_static_heterogeneous"}
    assert want <= have, want - have


def _graph_kwargs(name):
    seed, H, W, T, c, thresh, mgs, mask_p, hir_p, tf, cond, uea = GRAPH_CASES[name]
    return dict(thresh=thresh, max_grid_size=mgs, transform_func=dist_from_05 if tf else None, condition=cond,
                use_edge_attrs=uea)


def _check_graph(g, gold, to_cpu=lambda t: t):
    ei = to_cpu(g["edge_index"])
    assert tuple(ei.shape) == gold["edge_index"].shape, (tuple(ei.shape), gold["edge_index"].shape)
    assert np.array_equal(np.asarray(ei), gold["edge_index"]), "edge_index differs from the reference (order included)"
    if "labels" in gold:
        assert np.array_equal(np.asarray(to_cpu(g["labels"])).astype(np.int64), gold["labels"]), "labels differ"
    assert np.array_equal(np.asarray(to_cpu(g["n_pixels_per_node"])).astype(np.float32), gold["n_pixels_per_node"])
    assert attrs_close(torch.as_tensor(np.asarray(to_cpu(g["edge_attrs"]))), torch.from_numpy(gold["edge_attrs"]))
    if "data" in gold and g.get("data") is not None:
        d = torch.as_tensor(np.asarray(to_cpu(g["data"])))
        assert d.shape == gold["data"].shape
        assert torch.allclose(d, torch.from_numpy(gold["data"]), atol=DATA_TOL, rtol=DATA_TOL)


# ------------------------------------------------------------------------------- oracle vs reference vectors
@pytest.mark.parametrize("name", sorted(GRAPH_CASES))
def test_oracle_graph_matches_reference(name):
    from oracle import graph_ref as G
    gold = _load("graph_" + name)
    img = G.add_positional_encoding(torch.from_numpy(gold["x"]))
    g = G.image_to_graph(img, mask=gold.get("mask"), high_interest_region=gold.get("hir"), **_graph_kwargs(name))
    _check_graph(g, gold)
    if "unpooled" in gold:
        up = G.unpool(g["data"][0], g["mapping"], gold["x"].shape[1:3])
        assert torch.allclose(up, torch.from_numpy(gold["unpooled"]), atol=DATA_TOL)


def test_oracle_static_mesh_matches_reference():
    from oracle import graph_ref as G
    gold = _load("graph_static_heterogeneous")
    g = G.create_static_heterogeneous_graph(gold["mask"].shape, 4, gold["mask"], use_edge_attrs=True, resolution=1 / 12)
    _check_graph(g, gold)


def _seq_model(cls, name, gold, **extra):
    kw = dict(SEQ_CASES[name][6])
    if "Transformer" in kw.get("convolution_type", ""):
        kw["transform_func"] = dist_from_05
    m = cls(**kw, **extra)
    sd = {k[4:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("sd::")}
    missing = m.load_state_dict(sd, strict=True)          # same parameter names as the reference: part of the contract
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


@pytest.mark.parametrize("name", sorted(SEQ_CASES))
def test_oracle_seq2seq_matches_reference(name):
    from oracle import graph_ref as G
    from oracle.seq2seq_ref import Seq2Seq as OSeq
    gold = _load("seq2seq_" + name)
    model = _seq_model(OSeq, name, gold)
    with torch.no_grad():
        outs, maps = model(torch.from_numpy(gold["x"]), torch.from_numpy(gold["y"]), torch.from_numpy(gold["concat_layers"]),
                           teacher_forcing_ratio=0, mask=gold["mask"], remesh_every=SEQ_CASES[name][8])
    assert [o.shape[0] for o in outs] == gold["n_nodes"].tolist(), "mesh sizes differ from the reference's"
    H, W = gold["x"].shape[1:3]
    for t, (o, m) in enumerate(zip(outs, maps)):
        assert rel_err(o, torch.from_numpy(gold[f"out_{t}"])) < 1e-5, f"step {t}"
        fr = G.unpool(o, m, (H, W), gold["mask"])
        ref = torch.from_numpy(gold["frames"][t])
        assert torch.equal(torch.isnan(fr), torch.isnan(ref))
        assert rel_err(torch.nan_to_num(fr), torch.nan_to_num(ref)) < 1e-5


# ------------------------------------------------------------------------------- product vs reference vectors
@pytest.mark.parametrize("name", sorted(GRAPH_CASES))
def test_product_graph_matches_reference(be, name):
    import quadtree_mpnnlstm_b200 as q
    gold = _load("graph_" + name)
    img = q.add_positional_encoding(be.dev(torch.from_numpy(gold["x"])))
    g = q.image_to_graph(img, mask=gold.get("mask"), high_interest_region=gold.get("hir"), **_graph_kwargs(name))
    _check_graph(g, gold, to_cpu=lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t)
    if "unpooled" in gold:
        up = q.unflatten(g["data"][0], g["mapping"], gold["x"].shape[1:3]).cpu()
        assert torch.allclose(up, torch.from_numpy(gold["unpooled"]), atol=DATA_TOL)


def test_product_static_mesh_matches_reference(be):
    import quadtree_mpnnlstm_b200 as q
    gold = _load("graph_static_heterogeneous")
    g = q.create_static_heterogeneous_graph(gold["mask"].shape, 4, gold["mask"], use_edge_attrs=True, resolution=1 / 12,
                                            device=be.device)
    _check_graph(g, gold, to_cpu=lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t)


@pytest.mark.parametrize("name", sorted(SEQ_CASES))
def test_product_seq2seq_matches_reference(be, name):
    import quadtree_mpnnlstm_b200 as q
    gold = _load("seq2seq_" + name)
    model = be.dev(_seq_model(q.Seq2Seq, name, gold, device=be.device))
    with torch.no_grad():
        outs, maps = model(be.dev(torch.from_numpy(gold["x"])), be.dev(torch.from_numpy(gold["y"])),
                           be.dev(torch.from_numpy(gold["concat_layers"])), teacher_forcing_ratio=0, mask=gold["mask"],
                           remesh_every=SEQ_CASES[name][8])
    assert [o.shape[0] for o in outs] == gold["n_nodes"].tolist(), "mesh sizes differ from the reference's"
    H, W = gold["x"].shape[1:3]
    for t, (o, m) in enumerate(zip(outs, maps)):
        assert rel_err(o, torch.from_numpy(gold[f"out_{t}"])) < STEP_TOL, f"step {t}: {rel_err(o, torch.from_numpy(gold[f'out_{t}']))}"
        fr = q.unflatten(o, m, (H, W), gold["mask"]).cpu()
        ref = torch.from_numpy(gold["frames"][t])
        assert torch.equal(torch.isnan(fr), torch.isnan(ref))
        assert rel_err(torch.nan_to_num(fr), torch.nan_to_num(ref)) < STEP_TOL
