"""Data-parallel path with world_size = 2 over gloo on CPU: launch dates are sharded across
ranks, every rank runs its own forward / backward, ONE all-reduce averages the flat gradient bucket, and the
result equals the single-process mean of the same two samples' gradients.  The kernels are the test-only CPU
emulation of the C ABI (tests/cpu_emulation.py); what is under test is the host logic of
quadtree_mpnnlstm_b200.train.TrainStep and the sharding helpers."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _setup():
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _problem():
    from helpers import dist_from_05, moving_blob
    H, W, T_in, T_out = 12, 16, 2, 3
    rr, cc = np.mgrid[0:H, 0:W]
    mask = ((rr - H / 2) ** 2 / (H / 2.2) ** 2 + (cc - W / 2) ** 2 / (W / 2.5) ** 2) > 1
    kw = dict(hidden_size=32, dropout=0.0, thresh=-np.inf, input_timesteps=T_in, input_features=6, output_timesteps=T_out,
              n_layers=1, n_conv_layers=2, convolution_type="TransformerConv", transform_func=dist_from_05)
    samples = []
    for d in range(2):
        rng = np.random.default_rng(40 + d)
        x = np.concatenate([moving_blob(rng, T_in, H, W, size=6), rng.random((T_in, H, W, 2)).astype(np.float32)], -1)
        y = moving_blob(rng, T_out, H, W, size=6)
        cl = rng.random((T_out, H, W, 1)).astype(np.float32)
        samples.append([torch.from_numpy(a) for a in (x, y, cl)])
    return kw, mask, samples


def _grads_after_step(model):
    return {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}


def _worker(rank, world, init_file, out_dir):
    _setup()
    from cpu_emulation import Emulated
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.train import TrainStep, shard_launch_dates
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    kw, mask, samples = _problem()
    mine = shard_launch_dates(len(samples), rank, world)
    assert mine == [rank], mine
    with Emulated():
        torch.manual_seed(3)
        model = q.Seq2Seq(**kw).eval()      # eval(): no attention dropout (the emulation covers p = 0); grads still flow
        step = TrainStep(model, mask, lr=1e-3, use_cuda_graph=False, world_size=world, max_norm=1e9)
        loss = step(*samples[mine[0]])
        torch.save({"grads": _grads_after_step(model), "loss": float(loss),
                    "params": {k: p.detach().clone() for k, p in model.named_parameters()}}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process_mean():
    _setup()
    from cpu_emulation import Emulated
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.train import TrainStep
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "rdzv")
        mp.spawn(_worker, args=(2, init_file, tmp), nprocs=2, join=True)
        r0, r1 = (torch.load(os.path.join(tmp, f"r{r}.pt")) for r in range(2))
    # both ranks hold the same averaged gradients and the same updated parameters
    for k in r0["grads"]:
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
        assert torch.equal(r0["params"][k], r1["params"][k]), k
    # single process: mean of the two samples' gradients
    kw, mask, samples = _problem()
    ref = None
    with Emulated():
        for s in samples:
            torch.manual_seed(3)
            model = q.Seq2Seq(**kw).eval()      # eval(): no attention dropout (the emulation covers p = 0); grads still flow
            step = TrainStep(model, mask, lr=1e-3, use_cuda_graph=False, world_size=1, max_norm=1e9)
            step(*s)
            g = _grads_after_step(model)
            ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    for k, v in ref.items():
        want = v / 2
        got = r0["grads"][k]
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-8, k


def test_shard_launch_dates_partitions_evenly():
    _setup()
    from quadtree_mpnnlstm_b200.train import shard_launch_dates
    for n, world in [(8, 2), (64, 8), (10, 4), (3, 4)]:
        shards = [shard_launch_dates(n, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1, "every rank gets the same number of launch dates (one all-reduce per step)"
        seen = [d for s in shards for d in s]
        assert len(seen) == len(set(seen)) and all(0 <= d < n for d in seen)
        assert len(seen) == (n // world) * world
        # inference: every date is forecast; short ranks are padded to the same count (trimmed after the gather)
        padded = [shard_launch_dates(n, r, world, pad=True) for r in range(world)]
        assert len({len(s) for s in padded}) == 1 and len(padded[0]) == -(-n // world)
        inter = [padded[r][i] for i in range(len(padded[0])) for r in range(world)]
        assert inter[:n] == list(range(n)), "round-robin interleave restores launch-date order"


def _infer_worker(rank, world, init_file, out_dir):
    _setup()
    from cpu_emulation import Emulated
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.infer import predict_sharded
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    kw, mask, samples = _problem()
    samples = (samples * 2)[:3]                            # 3 launch dates over 2 ranks: rank 1 pads, the gather trims
    with Emulated():
        torch.manual_seed(3)
        model = q.Seq2Seq(**kw).eval()
        full = predict_sharded(model, lambda d: (samples[d][0], samples[d][2]), len(samples), mask, rank=rank, world=world)
    torch.save(full, os.path.join(out_dir, f"inf{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_rollout_inference_sharded_over_two_ranks():
    """ice_inf-style rollout (configs[4]): launch dates sharded, one all_gather, same forecasts as one process."""
    _setup()
    from cpu_emulation import Emulated
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.infer import predict_sharded
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_infer_worker, args=(2, os.path.join(tmp, "rdzv"), tmp), nprocs=2, join=True)
        a, b = (torch.load(os.path.join(tmp, f"inf{r}.pt")) for r in range(2))
    assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    kw, mask, samples = _problem()
    samples = (samples * 2)[:3]
    with Emulated():
        torch.manual_seed(3)
        model = q.Seq2Seq(**kw).eval()
        ref = predict_sharded(model, lambda d: (samples[d][0], samples[d][2]), len(samples), mask)
    assert a.shape == ref.shape == (3, 3, 12, 16, 1), 'no launch date may be dropped'
    assert torch.equal(torch.isnan(a), torch.isnan(ref))
    assert torch.allclose(torch.nan_to_num(a), torch.nan_to_num(ref), atol=1e-6)
