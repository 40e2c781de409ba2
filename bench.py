#!/usr/bin/env python
"""bench.py -- graph-frames/s of the MPNNLSTM training step (fwd + bwd + clip + Adam) on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): the reference's `ice_exp.py` default --
229 x 361 grid, 5 variables + 3 mesh features, pixel-wise static mesh over a synthetic ocean mask
(47 200 nodes, 187 808 edges), Seq2Seq(hidden 32, 1 layer, 3 encoder conv layers, TransformerConv),
10 input + 90 forecast steps = 100 graph-frames per sample, one optimizer step per sample (batch 1, as
in the reference trainer, model/mpnnlstm.py:219-257).  One "step" = one sample.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the CPU oracle port on the host cores

Under torchrun every rank trains on its own launch dates and the gradients are all-reduced over NCCL
(data parallel, weak scaling).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, N_VARS, T_IN, T_OUT, HIDDEN = 229, 361, 5, 10, 90, 32
FRAMES = T_IN + T_OUT
METRIC = "graph-frames/sec MPNNLSTM fwd+bwd"


def dist_from_05(arr):
    return abs(abs(arr - 0.5) - 0.5)


def ocean_mask(h=H, w=W):
    rr, cc = np.mgrid[0:h, 0:w]
    return ((rr - h / 2) ** 2 / (h / 2.2) ** 2 + (cc - w / 2) ** 2 / (w / 2.5) ** 2) > 1


def synthetic_cube(n_days, seed=21, h=H, w=W):
    """[n_days, H, W, 5] in [0, 1]; channel 0 = sea-ice-like step edge drifting 0.5 px/day + noise."""
    rng = np.random.default_rng(seed)
    cube = rng.random((n_days, h, w, N_VARS), dtype=np.float32)
    rows = np.arange(h, dtype=np.float32)[None, :, None]
    edge = 0.55 * h + 0.5 * np.arange(n_days, dtype=np.float32)[:, None, None]
    sic = (rows < edge).astype(np.float32) + rng.normal(0, 0.01, (n_days, h, w)).astype(np.float32)
    cube[..., 0] = np.clip(sic, 0, 1)
    return cube


def model_kwargs(t_in=T_IN, t_out=T_OUT, dropout=0.0):
    return dict(hidden_size=HIDDEN, dropout=dropout, thresh=-np.inf, input_timesteps=t_in, input_features=N_VARS + 3,
                output_timesteps=t_out, n_layers=1, n_conv_layers=3, convolution_type="TransformerConv",
                rnn_type="LSTM", transform_func=dist_from_05)


def sample(cube, clim, day, t_in=T_IN, t_out=T_OUT):
    x = cube[day:day + t_in]
    y = cube[day + t_in:day + t_in + t_out, :, :, :1]
    cl = clim[(day + t_in + np.arange(t_out)) % clim.shape[0]]
    return x, y, cl


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].startswith("Active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(threads, seconds_target=20.0, t_in=2, t_out=4):
    """The oracle port (reference driver + cell restated, oracle/seq2seq_ref.py) on the host cores, on a bounded
    sample of the same workload: the full 229 x 361 mesh, t_in + t_out frames instead of 10 + 90."""
    from oracle.seq2seq_ref import Seq2Seq as OSeq
    torch.set_num_threads(threads)
    mask = ocean_mask()
    cube = synthetic_cube(t_in + t_out + 4)
    clim = cube[..., :1].mean(0, keepdims=True).repeat(8, 0)
    torch.manual_seed(21)
    model = OSeq(**model_kwargs(t_in, t_out)).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    keep = torch.from_numpy(~mask)
    done, t0 = 0, None
    for it in range(64):
        x, y, cl = sample(cube, clim, it % 3, t_in, t_out)
        opt.zero_grad()
        out, _ = model(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask)
        y_nodes = torch.from_numpy(y)[:, keep]
        loss = torch.nn.functional.mse_loss(torch.stack(out), y_nodes)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10)
        opt.step()
        if it == 0:
            t0 = time.perf_counter()       # first sample = warm-up
            continue
        done += 1
        if time.perf_counter() - t0 > seconds_target:
            break
    dt = time.perf_counter() - t0
    return done * (t_in + t_out) / dt, dt, done, f"{done} samples of {t_in}+{t_out} frames on the full 229x361 pixel-wise mesh (47200 nodes)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    rate, dt, done, what = cpu_oracle_rate(threads, seconds_target=min(60.0, 6.0 * steps))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "graph-frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ice_exp default (configs[1]): 229x361 pixel-wise mesh, TransformerConv, hidden 32",
                       "note": "reference CPU path = oracle port (the reference itself cannot travel to the GPU box; "
                               "torch-geometric is not installable): reference driver/cell restated + restated PyG convs"},
            "cpu_baseline": {"value": rate, "unit": "graph-frames/s", "cores": threads, "kind": "port", "sample": what},
            "e2e": {"value": rate, "unit": "graph-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def cell_roofline(model, csr, dev, iters=20):
    """Time the roofline kernel of the step -- the decoder-cell forward (qmp_fused_cell_fwd: 8 TransformerConvs'
    message passing over the CSR + their gate contractions on tcgen05 + the LSTM gate epilogue, ONE launch per
    forecast step) -- alone with CUDA events on the launching stream, flushing L2 between launches.
    Algorithmic bytes per launch: DESIGN.md section 5 (inputs X, H, C + CSR + edge attributes; outputs O, H', C',
    head input; activations saved for the backward pass: gates, raw cell state, edge logits, softmax statistics)."""
    from quadtree_mpnnlstm_b200 import _lib, fused
    N, E, C = csr.n_nodes, csr.n_edges, HIDDEN
    cell = model.decoder.rnns[0]
    F_in = cell.in_channels
    with torch.no_grad():
        pa, pb = fused.pack_fused(cell._convs("x", 0), fused.cap_of(F_in, True)), fused.pack_fused(cell._convs("h", 0), C)
        img = fused.cell_image(pa, pb)
        prm = cell._gate_params(-1, model.decoder.norm_h, model.decoder.norm_c, model.decoder.norm_o).contiguous()
    f32 = dict(dtype=torch.float32, device=dev)
    X, Hs, Cs, cc = torch.randn(N, F_in, **f32), torch.randn(N, C, **f32), torch.randn(N, C, **f32), torch.randn(N, **f32)
    gates = torch.empty(N, 4 * C, **f32)
    Craw, O, Hn, Cn = (torch.empty(N, C, **f32) for _ in range(4))
    head = torch.empty(N, fused.HEADW, **f32)
    logit, ms, li = torch.empty(E, 8, **f32), torch.empty(N, 8, **f32), torch.empty(N, 8, **f32)
    usave = torch.empty(N, 4 * C, **f32)          # logit projections of the H convs, saved for qmp_fused_cell_bwd
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for k in range(iters + 3):
        flush.zero_()
        if k >= 3:
            ev[k - 3][0].record()
        _lib.call("qmp_fused_cell_fwd", N, csr.in_ptr, csr.in_src, csr.edge_attr_in, X, F_in, Hs, C, img, Cs, prm, 1, 1, 1, 1e-5,
                  gates, Craw, O, Hn, Cn, head, fused.HEADW, cc, logit, ms, li, usave, 0.0, 0)
        if k >= 3:
            ev[k - 3][1].record()
    torch.cuda.synchronize()
    ms_avg = sum(a.elapsed_time(b) for a, b in ev) / iters
    reads = 4 * (N * (F_in + 2 * C) + N) + 4 * (N + 1 + E) + 8 * E          # X, H, C, concat; CSR; edge attributes
    writes = 4 * (3 * N * C + N * fused.HEADW)                                # O, H', C', head input
    saved = 4 * (4 * N * C + N * C + 8 * E + 16 * N + 4 * N * C)              # gates, raw C', logits, softmax max / 1/sum, u
    return ms_avg, reads + writes + saved


def run_gpu(args):
    import torch.distributed as dist
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200 import _lib
    from quadtree_mpnnlstm_b200.graph_csr import get_csr

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        # stdout carries exactly one JSON line: with NCCL_DEBUG=VERSION (set on the GPU boxes) NCCL prints its version banner
        # there (NCCL_DEBUG_FILE does not move it); WARN keeps real diagnostics and drops the banner
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()   # fail loudly if the CUDA library is missing

    mask = ocean_mask()
    n_dates = args.warmup + args.steps + 2
    cube = synthetic_cube(FRAMES + n_dates * world + 2)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    torch.manual_seed(21)
    model = q.Seq2Seq(**model_kwargs(dropout=args.dropout), device=dev).to(dev).train()
    from quadtree_mpnnlstm_b200.train import TrainStep
    step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=not args.no_graph, world_size=world)

    def host_sample(i):
        x, y, cl = sample(cube, clim, rank + world * i)
        return [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x, y, cl)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop -> value
    n_warm = max(args.warmup, 4 if not args.no_graph else args.warmup)   # 3 eager steps + the capture step
    dev_samples = [[t.to(dev) for t in host_sample(i)] for i in range(n_warm + args.steps)]
    for i in range(n_warm):
        step(*dev_samples[i])
    barrier()
    launches0 = _lib.kernel_launches()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(*dev_samples[n_warm + i])
        e1.record()
        barrier()
    launches = _lib.kernel_launches() - launches0
    if not args.no_graph:
        launches = step.launches_per_replay * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    del dev_samples
    # ---- host-buffer loop through the public API -> e2e (H2D of the step's inputs + D2H of the loss inside)
    hs = [host_sample(i) for i in range(args.steps)]
    h2d = sum(t.numel() * 4 for t in hs[0])

    def host_loop(samples):
        nxt = step.stage(*samples[0])           # pinned host -> device on a copy stream; sample i + 1 travels under step i
        for i in range(len(samples)):
            cur = nxt
            loss_i = step(*cur)                 # -> the graph's static inputs -> replay (asynchronous)
            if i + 1 < len(samples):
                nxt = step.stage(*samples[i + 1])   # the next sample's copy is enqueued while this step runs
            last = loss_i.item()                # the loss comes back to the host every step
        return last

    host_loop(hs[:2])                           # untimed: the copy stream's allocator pool and the staging path warm up
    barrier()
    e0.record()
    last = host_loop(hs)
    e1.record()
    barrier()
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2)

    if rank == 0:
        value = args.steps * FRAMES * world / (ms_total / 1e3)
        e2e = args.steps * FRAMES * world / (ms_e2e / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        if model.graph is not None:
            topo = (model.graph.pyg.edge_index, model.graph.pyg.edge_attr, int(model.graph.pyg.x.shape[0]))
        else:
            topo = step.topology
        csr = get_csr(*topo)
        k_ms, k_bytes = cell_roofline(model, csr, dev)
        traffic = None
        try:      # dram bytes of the same kernel from the committed ncu --set full capture (profiles/), per launch
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_kernel_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        achieved = k_bytes / (k_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "graph-frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "ice_exp default (configs[1]): 229x361 grid, pixel-wise static mesh N=%d E=%d, "
                                       "TransformerConv hidden 32, 10+90 frames, batch 1, fwd+bwd+clip+Adam" % (csr.n_nodes, csr.n_edges),
                           "dropout": args.dropout, "cuda_graph": not args.no_graph, "parallelism": f"dp{world}" if world > 1 else "single",
                           "l2": "inputs + saved activations per step (~10 GB) exceed the 126 MB L2; roofline kernel "
                                 "timed with a 256 MB flush write between launches",
                           "final_loss": last},
                "e2e": {"value": e2e, "unit": "graph-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches,
                "clocks": clocks.summary(),
                "roofline": {"bound": "hbm", "kernel": "fused_cell_fwd_kernel (decoder cell forward: 8 convs' message "
                                                         "passing + tcgen05 gate contractions + LSTM epilogue, one launch)",
                             "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": traffic, "ms_per_launch": k_ms, "algorithmic_bytes": k_bytes,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650"}}
        if world == 1 and not args.no_cpu:
            rate, dt, done, what = cpu_oracle_rate(os.cpu_count() or 1, seconds_target=15.0)
            line["cpu_baseline"] = {"value": rate, "unit": "graph-frames/s", "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": what}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL teardown with a captured graph still holding the communicator hung for the full time limit on a 2-GPU
        # box (the JSON line had already been printed): synchronise, meet once more, and leave without it.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
