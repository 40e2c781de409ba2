"""Dropout under CUDA-graph replay: the captured step's kernel arguments -- including the per-call dropout
seeds -- are frozen, so the seeded kernels also mix in a DEVICE counter (qmp_set_dropout_salt, ops.dropout_salt) that the
step bumps.  The reference resamples its masks on every call (PyG TransformerConv.message's F.dropout; nn.Dropout at
model/seq2seq.py:169)."""
import numpy as np
import pytest
import torch

from helpers import dist_from_05, moving_blob

pytestmark = pytest.mark.gpu


def _problem(dev, dropout=0.1):
    import quadtree_mpnnlstm_b200 as q
    H, W, T_in, T_out = 24, 32, 3, 4
    rng = np.random.default_rng(11)
    rr, cc = np.mgrid[0:H, 0:W]
    mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    x = np.concatenate([moving_blob(rng, T_in, H, W, size=8), rng.random((T_in, H, W, 4)).astype(np.float32)], -1)
    y = moving_blob(rng, T_out, H, W, size=8)
    cl = rng.random((T_out, H, W, 1)).astype(np.float32)
    kw = dict(hidden_size=32, dropout=dropout, thresh=-np.inf, input_timesteps=T_in, input_features=8, output_timesteps=T_out,
              n_layers=1, n_conv_layers=2, convolution_type="TransformerConv", transform_func=dist_from_05)
    torch.manual_seed(3)
    model = q.Seq2Seq(**kw, device=dev).to(dev)
    return model, mask, [torch.from_numpy(a).to(dev) for a in (x, y, cl)]


def test_replays_resample_the_dropout_masks():
    """train() mode, identical inputs, lr = 0 (the weights never move): consecutive replays must give DIFFERENT losses (new
    masks), an eval() model the SAME loss every replay."""
    from quadtree_mpnnlstm_b200.train import TrainStep
    dev = torch.device("cuda")
    for training in (True, False):
        model, mask, (x, y, cl) = _problem(dev)
        model.train(training)
        step = TrainStep(model, mask, lr=0.0, use_cuda_graph=True)
        losses = [float(step(x, y, cl)) for _ in range(8)]
        assert step.graph is not None, "the CUDA-graph step must be the path that runs"
        replayed = losses[4:]            # 3 eager warm-up steps, the capture step's first replay, then pure replays
        if training:
            assert len({round(v, 9) for v in replayed}) == len(replayed), f"replays reuse one dropout mask: {replayed}"
            assert max(replayed) - min(replayed) < 0.2 * abs(np.mean(replayed)), replayed      # same expectation
        else:
            assert max(replayed) - min(replayed) <= 1e-6 * abs(replayed[0]), replayed


def test_forward_and_backward_of_one_step_share_the_mask():
    """With the per-call seeds AND the salt pinned the step is a deterministic function of the weights: its analytic gradient
    must match a central finite difference along the gradient direction -- it would be off by O(p) if the backward kernels drew
    another mask than the forward ones -- for salt 0 (none), 1 and 2, whose losses must differ."""
    from quadtree_mpnnlstm_b200 import ops
    dev = torch.device("cuda")
    model, mask, (x, y, cl) = _problem(dev)
    model.train()
    params = [p for p in model.parameters()]
    direction = [torch.zeros_like(p) for p in params]
    salt = torch.zeros(1, dtype=torch.int64, device=dev)

    def loss_at(eps, salt_value, grad=False):
        ops._seed_counter[0] = 0x1234567
        salt.fill_(salt_value)
        with torch.no_grad():
            for p, d in zip(params, direction):
                p.add_(eps * d)
        try:
            with ops.dropout_salt(salt if salt_value else None), torch.set_grad_enabled(grad):
                out, _ = model(x, y, cl, teacher_forcing_ratio=0, mask=mask)
                loss = sum((o.double() ** 2).mean() for o in out)
                if grad:
                    for p in params:
                        p.grad = None
                    loss.backward()
        finally:
            with torch.no_grad():
                for p, d in zip(params, direction):
                    p.sub_(eps * d)
        return float(loss)

    seen = []
    for s in (0, 1, 2):
        base = loss_at(0.0, s, grad=True)
        # direction = the normalised gradient itself (largest signal over the fp32 rounding noise of the loss)
        gnorm = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params if p.grad is not None)))
        for p, d in zip(params, direction):
            d.copy_(p.grad / gnorm if p.grad is not None else torch.zeros_like(p))
        analytic = gnorm
        assert loss_at(0.0, s) == pytest.approx(base, rel=1e-6), "same seeds + same salt must give the same masks"
        eps = 1e-2 * min(1.0, 1.0 / gnorm)
        numeric = (loss_at(eps, s) - loss_at(-eps, s)) / (2 * eps)
        assert numeric == pytest.approx(analytic, rel=3e-2, abs=1e-4), (s, numeric, analytic)
        seen.append(base)
    assert len({round(v, 9) for v in seen}) == 3, f"the salt must change the masks: {seen}"
