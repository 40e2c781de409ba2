#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.  TEST INFRASTRUCTURE.

    python tests/golden/make_golden.py            # build container only: needs /root/reference

The reference's hot-path files (model/graph_functions.py, model/utils.py, model/model.py, model/seq2seq.py)
are imported as they lie under /root/reference through ``oracle/ref_loader.py`` (import stubs for
matplotlib / torchviz, and a ``torch_geometric`` stand-in whose three convs are the restatements in
``oracle/convs_ref.py`` -- PyG 2.2.0 itself is not installable here, see oracle/__init__.py).  Outputs:

  graph_*.npz     inputs + labels / edge_index / edge_attrs / node data / pixel counts produced by the
                  reference's image_to_graph (graph_functions.py:590-681) and the pixel-wise /
                  static-mesh constructors (:506-539, :683-737)
  seq2seq_*.npz   inputs, a state dict, and the per-step forecasts + unpooled frames produced by the
                  reference's Seq2Seq.forward (seq2seq.py:402-418) on that state dict

The vectors are small (a few hundred kB in total) and committed; ``tests/test_golden.py`` checks the
oracle against them on CPU and the CUDA path against them on the GPU box (where /root/reference is absent).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import blob_frames, dist_from_05, moving_blob  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

GRAPH_CASES = {
    # name: (seed, H, W, T, c, thresh, max_grid_size, mask density, hir density, transform, condition, use_edge_attrs)
    "quadtree_plain": (1, 32, 32, 3, 1, 0.5, 8, None, None, None, "max_larger_than", True),
    "quadtree_ragged_mask_hir": (2, 37, 53, 2, 2, 0.3, 16, 0.15, 0.03, None, "max_larger_than", True),
    "quadtree_transform_weights": (3, 40, 52, 2, 1, 0.15, 64, 0.1, None, "dist_from_05", "max_larger_than", False),
    "quadtree_min_smaller": (4, 24, 40, 1, 1, 0.02, 8, None, None, None, "min_smaller_than", True),
    "quadtree_nothing_splits": (5, 16, 16, 1, 1, 2.0, 16, None, None, None, "max_larger_than", True),
    "quadtree_everything_splits": (6, 16, 24, 1, 1, -1.0, 8, None, None, None, "max_larger_than", False),
    "pixelwise_mask": (7, 23, 31, 2, 3, -np.inf, 64, 0.3, None, None, "max_larger_than", True),
}


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def graph_inputs(seed, H, W, T, c, mask_p, hir_p):
    rng = np.random.default_rng(seed)
    x = blob_frames(rng, T, H, W, c=c)
    mask = (rng.random((H, W)) < mask_p) if mask_p else None
    hir = (rng.random((H, W)) < hir_p) if hir_p else None
    return x, mask, hir


def make_graph_cases(ref):
    gf, ut = ref.graph_functions, ref.utils
    for name, (seed, H, W, T, c, thresh, mgs, mask_p, hir_p, tf, cond, uea) in GRAPH_CASES.items():
        x, mask, hir = graph_inputs(seed, H, W, T, c, mask_p, hir_p)
        img = ut.add_positional_encoding(torch.from_numpy(x))
        g = gf.image_to_graph(img, thresh=thresh, max_grid_size=mgs, mask=mask, high_interest_region=hir,
                              transform_func=dist_from_05 if tf else None, condition=cond, use_edge_attrs=uea)
        out = dict(x=x, edge_index=_np(g["edge_index"]).astype(np.int64), edge_attrs=_np(g["edge_attrs"]).astype(np.float32),
                   data=_np(g["data"]).astype(np.float32), n_pixels_per_node=_np(g["n_pixels_per_node"]).astype(np.float32),
                   graph_nodes=_np(g["graph_nodes"]).astype(np.int64))
        if mask is not None:
            out["mask"] = mask
        if hir is not None:
            out["hir"] = hir
        if g["mapping"] is not None:            # the dense [N, P] one-hot matrix -> per-pixel label image (-1 = masked)
            m = _np(g["mapping"].to_dense() if g["mapping"].is_sparse else g["mapping"])
            lab = np.where(m.sum(0) > 0, m.argmax(0), -1).reshape(H, W)
            out["labels"] = lab.astype(np.int64)
            # unpool of the node data must give back a piecewise-constant image (graph_functions.py:451-468)
            out["unpooled"] = _np(gf.unflatten(g["data"][0], g["mapping"], (H, W))).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"graph_{name}.npz"), **out)
        print(f"graph_{name}: N={out['data'].shape[1]} E={out['edge_index'].shape[1]}")

    # static meshes (ice_inf.py:60 / ice_exp.py -e 10)
    rng = np.random.default_rng(8)
    H, W = 30, 44
    mask = rng.random((H, W)) > 0.8
    g = gf.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12)
    m = _np(g["mapping"].to_dense() if g["mapping"].is_sparse else g["mapping"])
    np.savez_compressed(os.path.join(HERE, "graph_static_heterogeneous.npz"), mask=mask,
                        edge_index=_np(g["edge_index"]).astype(np.int64), edge_attrs=_np(g["edge_attrs"]).astype(np.float32),
                        n_pixels_per_node=_np(g["n_pixels_per_node"]).astype(np.float32),
                        labels=np.where(m.sum(0) > 0, m.argmax(0), -1).reshape(H, W).astype(np.int64))
    print("graph_static_heterogeneous: N=%d" % m.shape[0])


SEQ_CASES = {
    # name: (seed, H, W, T_in, T_out, c, model kwargs, mask kind, remesh_every)
    "ice_pixelwise_transformer": (11, 20, 28, 3, 4, 5, dict(hidden_size=32, dropout=0.0, thresh=-np.inf, input_timesteps=3,
                                                            input_features=8, output_timesteps=4, n_layers=1, n_conv_layers=3,
                                                            convolution_type="TransformerConv"), "ellipse", 1),
    "mnist_quadtree_cheb": (12, 32, 32, 3, 4, 1, dict(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=3,
                                                      input_features=4, output_timesteps=4, n_layers=2, n_conv_layers=2),
                            "none", 1),
    "ice_quadtree_transformer": (13, 24, 40, 3, 3, 2, dict(hidden_size=16, dropout=0.0, thresh=0.15, input_timesteps=3,
                                                           input_features=5, output_timesteps=3, n_layers=1, n_conv_layers=2,
                                                           convolution_type="TransformerConv"), "random", 1),
}


def seq_inputs(seed, H, W, T_in, T_out, c, mask_kind):
    rng = np.random.default_rng(seed)
    x = moving_blob(rng, T_in, H, W)
    if c > 1:
        x = np.concatenate([x, rng.random((T_in, H, W, c - 1)).astype(np.float32)], -1)
    y = moving_blob(rng, T_out, H, W)
    cl = rng.random((T_out, H, W, 1)).astype(np.float32)
    if mask_kind == "ellipse":
        rr, cc = np.mgrid[0:H, 0:W]
        mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    elif mask_kind == "random":
        mask = rng.random((H, W)) > 0.85
    else:
        mask = np.zeros((H, W), bool)
    return x, y, cl, mask


def make_seq_cases(ref):
    gf = ref.graph_functions
    for name, (seed, H, W, T_in, T_out, c, kw, mask_kind, remesh_every) in SEQ_CASES.items():
        x, y, cl, mask = seq_inputs(seed, H, W, T_in, T_out, c, mask_kind)
        kw = dict(kw)
        if "Transformer" in kw.get("convolution_type", ""):
            kw["transform_func"] = dist_from_05
        torch.manual_seed(seed)
        model = ref.seq2seq.Seq2Seq(**kw, device=torch.device("cpu"))
        with torch.no_grad():                      # peepholes / gate biases are zero-initialised: make them matter
            gen = torch.Generator().manual_seed(seed + 100)
            for k, p in model.named_parameters():
                if ".w_c_" in k or ".b_" in k or "norm" in k:
                    p.add_(0.1 * torch.randn(p.shape, generator=gen))
        model.eval()
        with torch.no_grad():
            outs, maps = model(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0,
                               mask=mask, remesh_every=remesh_every)
        frames = np.stack([_np(gf.unflatten(o, m, (H, W), mask)) for o, m in zip(outs, maps)]).astype(np.float32)
        blob = dict(x=x, y=y, concat_layers=cl, mask=mask, frames=frames,
                    n_nodes=np.array([o.shape[0] for o in outs], np.int64))
        for t, o in enumerate(outs):
            blob[f"out_{t}"] = _np(o).astype(np.float32)
        for k, v in model.state_dict().items():
            blob["sd::" + k] = _np(v)
        np.savez_compressed(os.path.join(HERE, f"seq2seq_{name}.npz"), **blob)
        print(f"seq2seq_{name}: nodes per step {blob['n_nodes'].tolist()}")


class RefWindows(torch.utils.data.Dataset):
    """What ice_dataset.py:20-68 serves: materialised (x, y, launch_date) windows + image_shape."""

    def __init__(self, cube, t_in, t_out, times, idx):
        self.x = [cube[a:a + t_in] for a in idx]
        self.y = [cube[a + t_in:a + t_in + t_out][..., :1] for a in idx]
        self.d = [times[a + t_in] for a in idx]
        self.image_shape = cube.shape[1:3]

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i], self.y[i], self.d[i]


TRAINER_CASES = {
    # name: (thresh, conv kwargs, T_in, T_out, truncated_backprop, epochs)
    "quadtree_cheb": (0.1, dict(hidden_size=8, dropout=0.0, n_layers=1, n_conv_layers=1), 3, 3, 0, 2),
    "pixelwise_transformer": (-np.inf, dict(hidden_size=32, dropout=0.0, n_layers=1, n_conv_layers=2, convolution_type="TransformerConv"),
                              3, 4, 0, 2),
    "quadtree_cheb_truncated": (0.1, dict(hidden_size=8, dropout=0.0, n_layers=1, n_conv_layers=1), 3, 4, 2, 2),
}


def make_trainer_cases(ref):
    """The UNMODIFIED reference trainer (model/mpnnlstm.py: NextFramePredictorS2S.train / predict) on a small cube: epoch
    losses, the forecasts of predict(), initial weights (the climatology is random((1, 366, H, W)) of seed ``clim_seed``).  eval() mode keeps the TransformerConv attention dropout
    off (the product cannot reproduce torch's dropout stream)."""
    import importlib
    import tempfile
    R = importlib.import_module("model.mpnnlstm")
    H, W, c, T = 16, 20, 2, 16
    for name, (thresh, kw, T_in, T_out, tb, epochs) in TRAINER_CASES.items():
        rng = np.random.default_rng(3)
        cube = moving_blob(rng, T, H, W)
        cube = np.concatenate([cube, rng.random((T, H, W, c - 1)).astype(np.float32) * 0.5], -1).astype(np.float32)
        times = (np.datetime64("2015-03-01").astype("datetime64[ns]").astype("int64") + np.arange(T, dtype=np.int64) * 86_400_000_000_000)
        mask = np.zeros((H, W), bool)
        mask[:3, :5] = True
        mask[10:, 15:] = True
        clim = np.random.default_rng(5).random((1, 366, H, W)).astype(np.float32)
        tr_idx, te_idx = [0, 2, 4], [6, 7]
        torch.manual_seed(4)
        tr = R.NextFramePredictorS2S(thresh, experiment_name="g", input_features=c, input_timesteps=T_in, output_timesteps=T_out,
                                     device=torch.device("cpu"), model_kwargs=kw,
                                     transform_func=dist_from_05 if "Transformer" in kw.get("convolution_type", "") else None)
        init = {k: _np(v).copy() for k, v in tr.model.state_dict().items()}
        mk = lambda idx: torch.utils.data.DataLoader(RefWindows(cube, T_in, T_out, times, idx), batch_size=1, shuffle=False)
        tr.model.eval()
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)                            # the reference writes TensorBoard runs/ into the working directory
            try:
                tr.train(mk(tr_idx), mk(te_idx), torch.from_numpy(clim), n_epochs=epochs, lr=0.01, lr_decay=0.5, mask=mask,
                         truncated_backprop=tb)
                pred = tr.predict(mk(te_idx), torch.from_numpy(clim), mask=mask)
            finally:
                os.chdir(cwd)
        blob = dict(cube=cube, times=times, mask=mask, clim_seed=np.array(5), tr_idx=np.array(tr_idx), te_idx=np.array(te_idx),
                    train_loss=np.array(tr.train_loss, np.float64), test_loss=np.array(tr.test_loss, np.float64),
                    predict=pred.astype(np.float32), last_lr=np.array(tr.scheduler.get_last_lr(), np.float64))
        for k, v in init.items():
            blob["init::" + k] = v
        np.savez_compressed(os.path.join(HERE, f"trainer_{name}.npz"), **blob)
        print(f"trainer_{name}: train {tr.train_loss} test {tr.test_loss}")


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("/root/reference is not available: golden vectors can only be generated in the build container")
    if "--trainer-only" not in sys.argv:
        make_graph_cases(ref)
        make_seq_cases(ref)
    make_trainer_cases(ref)


if __name__ == "__main__":
    main()
