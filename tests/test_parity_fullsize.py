"""GPU parity at BASELINE.json's full sizes: the CUDA path vs the CPU oracle on the very workload
bench.py times -- configs[1] (ice_exp.py:57-58, 153-162: 229 x 361 grid, pixel-wise mesh N = 47 200 / E = 187 808,
TransformerConv hidden 32, 10 input + 90 forecast steps) and configs[4] (ice_inf.py:60: static heterogeneous mesh, max
cell 4, N = 4 066 / E = 19 086, no_grad rollout).  Tolerances are the north-star ones: every forecast step within 1e-4
relative (max |a - b| / max |b| per step), gradients within 1e-3 relative per parameter tensor.

One measured fact shapes the configs[1] forward test.  With these (randomly initialised, perturbed) weights the 90-step
recurrence amplifies rounding differences by ~5 % per step: the ORACLE ITSELF, evaluated in float64 instead of float32
(oracle/precision.py -- same algorithm, exact arithmetic), moves by 1.2e-6 at forecast step 0, 1.4e-5 at step 40 and 2.0e-4
at step 89.  Two correct float32 implementations therefore cannot agree to 1e-4 on the late steps (the reference against
itself with another summation order would not), and the test states the bar in the only form that is decidable:
  * against the float32 oracle, 1e-4 at every step whose float32 noise floor (|oracle_f32 - oracle_f64|) is below 2.5e-5;
  * against the float64 oracle, at EVERY step: the CUDA path's error is at most max(1e-4, 1.5 x the oracle's own float32
    error at that step) -- the CUDA result is as close to the exact result as the reference's arithmetic is.

The oracle forward of one 100-frame sample takes ~15 s (float32) + ~90 s (float64) on the box's host cores."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEP_TOL = 1e-4
GRAD_TOL = 1e-3

pytestmark = pytest.mark.gpu


def _models(kw, dev, seed=21):
    import quadtree_mpnnlstm_b200 as q
    from oracle.seq2seq_ref import Seq2Seq as OSeq
    torch.manual_seed(seed)
    ref = OSeq(**kw)
    with torch.no_grad():          # peepholes / biases / norms away from their zero / one init values
        gen = torch.Generator().manual_seed(9)
        for k, p in ref.named_parameters():
            if ".w_c_" in k or ".b_" in k or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    gpu = q.Seq2Seq(**kw, device=dev).to(dev)
    gpu.load_state_dict(ref.state_dict())
    return ref.eval(), gpu.eval()


def _configs1_sample(t_in, t_out, day=1):
    import bench as B
    mask = B.ocean_mask()
    cube = B.synthetic_cube(t_in + t_out + 4)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    x, y, cl = (np.ascontiguousarray(a) for a in B.sample(cube, clim, day, t_in, t_out))
    return mask, x, y, cl


def test_configs1_full_size_every_forecast_step():
    """229 x 361, N = 47 200, 10 + 90 frames: every one of the 90 forecast steps against the oracle (see the module
    docstring for the form of the bar)."""
    import bench as B
    from oracle.precision import float64
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda")
    mask, x, y, cl = _configs1_sample(B.T_IN, B.T_OUT)
    assert x.shape == (10, 229, 361, 5) and y.shape == (90, 229, 361, 1)
    ref, gpu = _models(B.model_kwargs(), dev)
    with torch.no_grad():
        oa, _ = ref(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask)
        ob, _ = gpu(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(cl).to(dev),
                    teacher_forcing_ratio=0, mask=mask)
        ei32 = ref.graph.edge_index
        with float64():
            ref64 = ref.double()
            od, _ = ref64(torch.from_numpy(x).double(), torch.from_numpy(y).double(), torch.from_numpy(cl).double(),
                          teacher_forcing_ratio=0, mask=mask)
    assert len(oa) == len(ob) == len(od) == 90
    assert tuple(oa[0].shape) == tuple(ob[0].shape) == (47200, 1)
    assert gpu.graph.pyg.edge_index.shape[1] == 187808
    assert torch.equal(gpu.graph.pyg.edge_index.cpu(), ei32)
    e_gpu32 = [rel_err(b, a) for a, b in zip(oa, ob)]            # CUDA vs the oracle in the reference's arithmetic
    noise = [rel_err(a, d) for a, d in zip(oa, od)]              # the oracle's own float32 rounding noise
    e_gpu64 = [rel_err(b, d) for b, d in zip(ob, od)]            # CUDA vs the exact-arithmetic oracle
    for t in (0, 9, 19, 29, 39, 49, 59, 69, 79, 89):
        print("configs[1] step %2d: cuda-vs-f32-oracle %.2e | f32-oracle-vs-f64 (noise floor) %.2e | cuda-vs-f64 %.2e"
              % (t, e_gpu32[t], noise[t], e_gpu64[t]))
    strict = [t for t in range(90) if noise[t] < 2.5e-5]
    assert len(strict) >= 30, f"only {len(strict)} steps have a float32 noise floor under 2.5e-5"
    for t in strict:
        assert e_gpu32[t] < STEP_TOL, f"forecast step {t}: rel err {e_gpu32[t]} vs the float32 oracle (noise floor {noise[t]})"
    for t in range(90):
        bar = max(STEP_TOL, 1.5 * noise[t])
        assert e_gpu64[t] <= bar, f"forecast step {t}: {e_gpu64[t]} from the exact result, the oracle's float32 path {noise[t]}"
    print(f"configs[1]: strict 1e-4 bar held on the {len(strict)} steps with noise floor < 2.5e-5 (last: step {strict[-1]}); "
          f"worst cuda-vs-f64 {max(e_gpu64):.2e}, worst f32-oracle-vs-f64 {max(noise):.2e}")


def test_configs1_full_size_gradients():
    """The trainer's loss (MSE on the unmasked pixels == on the nodes of the pixel mesh) and the gradient of EVERY parameter
    tensor on the full mesh.  Frames: 10 + 90 when the host can hold the oracle's autograd tape of a full sample (~140 GB),
    else the same 1 : 9 encoder : decoder mix with fewer frames (bench.oracle_frames_for_memory)."""
    import bench as B
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda")
    t_in, t_out, mem = B.oracle_frames_for_memory()
    mask, x, y, cl = _configs1_sample(t_in, t_out)
    ref, gpu = _models(B.model_kwargs(t_in, t_out), dev)
    oa, _ = ref(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask)
    ob, _ = gpu(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(cl).to(dev),
                teacher_forcing_ratio=0, mask=mask)
    keep = torch.from_numpy(~mask)
    ya = torch.from_numpy(y)[:, keep]
    la = torch.nn.functional.mse_loss(torch.stack(oa), ya)
    lb = torch.nn.functional.mse_loss(torch.stack(ob), ya.to(dev))
    assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(la))
    la.backward()
    lb.backward()
    n_checked, worst_g = 0, (0.0, "")
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    for (k, pa), (_, pb) in zip(ref.named_parameters(), gpu.named_parameters()):
        if pa.grad is None:
            assert pb.grad is None or float(pb.grad.abs().max()) == 0.0, k
            continue
        ga = pa.grad
        gb = pb.grad.cpu() if pb.grad is not None else torch.zeros_like(ga)
        scale = float(ga.abs().max())
        if scale < 1e-7 * gmax:
            # analytically zero (lin_key.bias: the softmax does not see a per-target constant): the oracle holds rounding
            # residue there (~1e-11), the CUDA path an exact zero
            assert float(gb.abs().max()) <= 1e-6 * gmax, f"grad {k} should vanish: {float(gb.abs().max())}"
            continue
        err = float((ga - gb).abs().max()) / scale
        n_checked += 1
        if err > worst_g[0]:
            worst_g = (err, k)
        assert err < GRAD_TOL, f"grad {k}: rel {err} (scale {scale})"
    print(f"configs[1] gradients ({t_in}+{t_out} frames, host memory {mem:.0f} GB): {n_checked} parameter tensors within "
          f"{GRAD_TOL}, worst {worst_g[0]:.2e} ({worst_g[1]})")
    assert n_checked >= 50


def test_configs4_static_heterogeneous_mesh_full_size_rollout():
    import bench as B
    import quadtree_mpnnlstm_b200 as q
    from oracle import graph_ref as G
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda")
    mask = B.ocean_mask()
    cube = B.synthetic_cube(B.FRAMES + 4)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    x, y, cl = (np.ascontiguousarray(a) for a in B.sample(cube, clim, 2))
    ref, gpu = _models(B.model_kwargs(), dev, seed=22)
    gs_a = G.create_static_heterogeneous_graph((B.H, B.W), 4, mask, use_edge_attrs=True, resolution=1 / 12)
    gs_b = q.create_static_heterogeneous_graph((B.H, B.W), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=dev)
    assert torch.equal(gs_b["edge_index"].cpu(), gs_a["edge_index"])
    assert gs_a["edge_index"].shape[1] == 19086
    with torch.no_grad():
        oa, ma = ref(torch.from_numpy(x), None, torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask, graph_structure=gs_a)
        ob, mb = gpu(torch.from_numpy(x).to(dev), None, torch.from_numpy(cl).to(dev), teacher_forcing_ratio=0, mask=mask,
                     graph_structure=gs_b)
    assert len(oa) == len(ob) == 90 and tuple(ob[0].shape) == (4066, 1)
    errs = [rel_err(b, a) for a, b in zip(oa, ob)]
    worst = int(np.argmax(errs))
    print("configs[4] per-step rel err: step0 %.2e step44 %.2e step89 %.2e worst %.2e @%d" % (errs[0], errs[44], errs[89], errs[worst], worst))
    assert errs[worst] < STEP_TOL, f"forecast step {worst}: rel err {errs[worst]}"
    ia = G.unpool(oa[-1], ma[-1], (B.H, B.W), mask)
    ib = q.unflatten(ob[-1], mb[-1], (B.H, B.W), mask).cpu()
    assert torch.equal(torch.isnan(ia), torch.isnan(ib))
    assert rel_err(torch.nan_to_num(ib), torch.nan_to_num(ia)) < STEP_TOL
