#!/usr/bin/env python
"""bench.py -- graph-frames/s of the MPNNLSTM training step (fwd + bwd + clip + Adam) on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): the reference's `ice_exp.py` default --
229 x 361 grid, 5 variables + 3 mesh features, pixel-wise static mesh over a synthetic ocean mask
(47 200 nodes, 187 808 edges), Seq2Seq(hidden 32, 1 layer, 3 encoder conv layers, TransformerConv),
10 input + 90 forecast steps = 100 graph-frames per sample, one optimizer step per sample (batch 1, as
in the reference trainer, model/mpnnlstm.py:219-257).  One "step" = one sample.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the CPU oracle port on the host cores
    python bench.py --mode infer --gpus N                     # configs[4]: rollout inference, launch dates sharded

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
WORKLOAD = ("ice_exp default (configs[1]): 229x361 grid, pixel-wise static mesh N=%d E=%d, TransformerConv hidden 32, "
            "10+90 frames, batch 1, fwd+bwd+clip+Adam")


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
def oracle_frames_for_memory():
    """(t_in, t_out) of the CPU oracle's timed samples.  The oracle's autograd tape of one full 10 + 90 sample is ~140 GB of
    host memory (E x 32 edge tensors of 24 convs per encoder frame, 10 per decoder frame); on a host that cannot hold it the
    sample keeps the encoder : decoder frame mix of the workload (1 : 9, so the per-frame cost is the same) with fewer frames."""
    mem = host_memory_gb()
    if mem >= 220:
        return T_IN, T_OUT, mem
    if mem >= 48:
        return 2, 18, mem
    return 1, 9, mem


def cpu_oracle_rate(threads, max_samples=1, seconds_cap=150.0, dropout=0.1, t_in=T_IN, t_out=T_OUT):
    """The oracle port (reference driver + cell restated, oracle/seq2seq_ref.py) on the host cores, on the SAME workload as the
    GPU arm: full 10 + 90-frame samples on the full 229 x 361 mesh, train() mode (attention dropout 0.1 active, decoder dropout
    as given), fwd + bwd + clip + Adam.  Bounded sample: a 1 + 1-frame mini-sample warms up (allocator, thread pool), then
    ``max_samples`` full samples are timed (stops early once ``seconds_cap`` is exceeded)."""
    from oracle.seq2seq_ref import Seq2Seq as OSeq
    torch.set_num_threads(threads)
    mask = ocean_mask()
    cube = synthetic_cube(t_in + t_out + max_samples + 4)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    keep = torch.from_numpy(~mask)

    def one(model, opt, day, t_in, t_out):
        x, y, cl = sample(cube, clim, day, t_in, t_out)
        opt.zero_grad()
        out, _ = model(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(cl), teacher_forcing_ratio=0, mask=mask)
        loss = torch.nn.functional.mse_loss(torch.stack(out), torch.from_numpy(y)[:, keep])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10)
        opt.step()
        return float(loss)

    torch.manual_seed(21)
    warm = OSeq(**model_kwargs(1, 1, dropout)).train()
    one(warm, torch.optim.Adam(warm.parameters(), lr=1e-4), 0, 1, 1)
    del warm
    torch.manual_seed(21)
    model = OSeq(**model_kwargs(t_in, t_out, dropout)).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    done, t0 = 0, time.perf_counter()
    for it in range(max_samples):
        one(model, opt, it, t_in, t_out)
        done += 1
        if time.perf_counter() - t0 > seconds_cap:
            break
    dt = time.perf_counter() - t0
    what = (f"{done} sample(s) of {t_in}+{t_out} frames on the full 229x361 pixel-wise mesh (47200 nodes, 187808 edges), "
            f"train mode, fwd+bwd+clip+Adam, after a 1+1-frame warm-up"
            + ("" if (t_in, t_out) == (T_IN, T_OUT) else f"; same 1:9 encoder:decoder frame mix as 10+90, fewer frames because the "
               f"oracle's autograd tape of a full sample (~140 GB) does not fit this host ({host_memory_gb():.0f} GB)"))
    return done * (t_in + t_out) / dt, dt, done, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    t_in, t_out, _ = oracle_frames_for_memory()
    rate, dt, done, what = cpu_oracle_rate(threads, max_samples=steps, seconds_cap=150.0, dropout=args.dropout, t_in=t_in, t_out=t_out)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "graph-frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (47200, 187808), "dropout": args.dropout, "cuda_graph": False,
                       "parallelism": "cpu",
                       "note": "reference CPU path = oracle port (the reference itself cannot travel to the GPU box; "
                               "torch-geometric is not installable): reference driver/cell restated + restated PyG convs; "
                               f"{done} of the {steps} requested steps timed (each step = one {t_in}+{t_out}-frame sample, bounded at 150 s)"},
            "cpu_baseline": {"value": rate, "unit": "graph-frames/s", "cores": threads, "kind": "port", "sample": what},
            "e2e": {"value": rate, "unit": "graph-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def host_memory_gb():
    """Memory this process may use (cgroup limit when there is one, else what the host reports as available)."""
    lim = None
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v.isdigit() and int(v) < (1 << 60):
                lim = int(v)
                break
        except Exception:
            pass
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = None
    cands = [v for v in (lim, avail) if v]
    return min(cands) / 2 ** 30 if cands else 0.0


def algorithmic_bytes(N, E, C=HIDDEN, d_e=2):
    """SURVEY.md section 8(d), verbatim: compulsory fp32 traffic per graph-frame of each part of the step."""
    def b_f(F, S):
        return 4 * (N * (F + 2 * C) + 3 * N * C) + 4 * (E + N + 1) + 4 * E * d_e + (S - 1) * 2 * 4 * 8 * N * C

    def b_b(F, S):
        return 4 * (3 * N * C + N * (F + 2 * C) + 6 * N * C) + 4 * (E + N + 1) + 4 * E * d_e + (S - 1) * 2 * 4 * 8 * N * C

    saved = 4 * 6 * N * C
    head = 4 * (N * (C + 1) + 2 * N * C + N)
    return {"decoder_cell_fwd": b_f(4, 1) + saved,      # B_f + the gate activations a training forward must write
            "decoder_cell_bwd": b_b(4, 1),              # gate backward + message-passing backward + weight gradients
            "decoder_head_fwd": head,
            "decoder_head_bwd": 2 * head,               # not in 8(d): inputs re-read + one gradient per forward tensor
            "encoder_fwd": b_f(8, 3) + saved,
            "encoder_bwd": b_b(8, 3)}


def _classify(name, args):
    """(group, short kernel key) of one C-ABI call of the training step."""
    if name == "qmp_fused_cell_fwd":
        return "decoder_cell_fwd", name
    if name in ("qmp_fused_cell_bwd", "qmp_fused_cell_bwd_full", "qmp_cell_wgrad"):
        return "decoder_cell_bwd", name
    if name == "qmp_lstm_gates_bwd":        # args: N C gates Craw Cp prm norm_h norm_c norm_o ...
        return ("decoder_cell_bwd" if args[8] else "encoder_bwd"), name
    if name.startswith("qmp_fused_fwd") or name.startswith("qmp_fused_bwd"):
        DA, GA, DB, GB, mode = args[6], args[7], args[11], args[12], args[15]
        key = f"{name}[DA={DA} GA={GA} DB={DB} GB={GB} mode={mode}]"
        fwd = "_fwd" in name
        if DB == 36:
            return ("decoder_head_fwd" if fwd else "decoder_head_bwd"), key
        if GA == 4 and DA == 4:
            return ("decoder_cell_fwd" if fwd else "decoder_cell_bwd"), key
        return ("encoder_fwd" if fwd else "encoder_bwd"), key
    if name in ("qmp_fused_wgrad", "qmp_fused_wgrad_tma"):
        DA, GA, DB, GB, mode = args[3], args[4], args[7], args[8], args[10]
        key = f"{name}[DA={DA} GA={GA} DB={DB} GB={GB} mode={mode}]"
        if DB == 36:
            return "decoder_head_bwd", key
        return ("decoder_cell_bwd" if (GA == 4 and DA == 4) else "encoder_bwd"), key
    if name in ("qmp_tconv1_fwd", "qmp_head_finish_fwd", "qmp_head_tail_fwd"):
        return "decoder_head_fwd", name
    if name in ("qmp_tconv1_bwd", "qmp_head_finish_bwd", "qmp_relu_mask_to", "qmp_relu_mask", "qmp_panel_wgrad", "qmp_head_tail_bwd", "qmp_head_bwd"):
        return "decoder_head_bwd", name
    if name.startswith("qmp_pack_") or name.startswith("qmp_fused_pack") or name in ("qmp_add_positional_encoding", "qmp_segment_sum",
                                                                                   "qmp_gather_by_label"):
        return "per_step", name          # once per optimizer step (weight packs / images, input pooling), whatever the frame count
    return "other", name


def per_kernel_times(dev, mask, cube, clim, dropout, shapes=((2, 4), (3, 5)), reps=2):
    """Every C-ABI call of one eager training sample (full mesh) timed ALONE with CUDA events on the launching stream, L2
    flushed by a 256 MB write before each call.  Two sample shapes (t_in, t_out) are run so that the number of calls of
    every kernel in a T_IN + T_OUT-frame step follows from a linear fit (a kernel that runs once per sample -- e.g. the
    first forecast step's -- is not charged to every frame).  Returns {key: (group, avg us, calls per frame at 10 + 90)}."""
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200 import _lib
    from quadtree_mpnnlstm_b200.train import TrainStep
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    tot, cnt = {}, {}
    N = E = 0
    pdl_was = _lib.set_pdl(False)        # a kernel timed alone starts after the flush write has finished
    for si, (t_in, t_out) in enumerate(shapes):
        torch.manual_seed(21)
        model = q.Seq2Seq(**model_kwargs(t_in, t_out, dropout), device=dev).to(dev).train()
        step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
        smp = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sample(cube, clim, 0, t_in, t_out)]
        step(*smp)
        records, orig = [], _lib.call

        def timed(name, *args):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig(name, *args)
            e1.record()
            records.append((_classify(name, args), e0, e1))

        _lib.call = timed
        try:
            for _ in range(reps):
                step(*smp)
        finally:
            _lib.call = orig
        torch.cuda.synchronize()
        for gk, a, b in records:
            tot[gk] = tot.get(gk, 0.0) + a.elapsed_time(b) * 1e3
            cnt.setdefault(gk, [0] * len(shapes))[si] += 1
        N, E = int(model.graph.pyg.x.shape[0]), int(model.graph.pyg.edge_index.shape[1])
        del model, step
    _lib.set_pdl(pdl_was)
    out = {}
    (i0, o0), (i1, o1) = shapes
    for (group, key), t in tot.items():
        c0, c1 = (c / reps for c in cnt[(group, key)])
        if group == "per_step":
            frames, calls = 1, max(c0, c1)
        else:       # calls = slope * frames + once-per-sample part, from the two shapes
            f0, f1, frames = (i0, i1, T_IN) if group.startswith("encoder") else (o0, o1, T_OUT)
            slope = (c1 - c0) / (f1 - f0)
            calls = slope * frames + (c0 - slope * f0)
        out[key] = (group, t / sum(cnt[(group, key)]), calls / frames)
    return out, N, E


def roofline_report(times, N, E, hbm, peaks_known):
    """Groups of section 8(d) with their live times; the roofline is reported on the group with the largest share of a
    10 + 90-frame step (today: the decoder-cell backward)."""
    bytes_of = algorithmic_bytes(N, E)
    frames_of = lambda g: 1 if g == "per_step" else (T_IN if g.startswith("encoder") else T_OUT)
    groups = {}
    for key, (group, us, per_frame) in times.items():
        g = groups.setdefault(group, {"us_per_frame": 0.0, "kernels": []})
        g["us_per_frame"] += us * per_frame
        g["kernels"].append({"kernel": key, "us": round(us, 2), "launches_per_frame": round(per_frame, 3)})
    step_us = sum(g["us_per_frame"] * frames_of(name) for name, g in groups.items())
    per_group = []
    for name, g in sorted(groups.items(), key=lambda kv: -kv[1]["us_per_frame"] * frames_of(kv[0])):
        row = {"group": name, "us_per_frame": round(g["us_per_frame"], 2), "frames_per_step": frames_of(name),
               "share_of_step": round(g["us_per_frame"] * frames_of(name) / step_us, 4),
               "kernels": sorted(g["kernels"], key=lambda k: -k["us"] * k["launches_per_frame"])}
        if name in bytes_of:
            gbs = bytes_of[name] / (g["us_per_frame"] * 1e-6) / 1e9
            row.update({"algorithmic_bytes": bytes_of[name], "achieved_gbs": round(gbs, 1), "frac": round(gbs / hbm, 4)})
        per_group.append(row)
    top = next(r for r in per_group if "frac" in r)
    traffic = None
    try:      # dram bytes of the same kernels from the committed ncu --set full capture (profiles/), per frame
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_kernel_traffic.json")))
        traffic = t.get(top["group"], {}).get("dram_bytes_per_frame")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": top["group"] + ": " + " + ".join(k["kernel"] for k in top["kernels"]),
            "achieved": top["achieved_gbs"], "peak": hbm, "unit": "GB/s", "frac": top["frac"], "traffic": traffic,
            "ms_per_launch": top["us_per_frame"] / 1e3, "algorithmic_bytes": top["algorithmic_bytes"],
            "bytes_formula": "SURVEY.md 8(d): fwd B_f + 4*6NC, bwd B_b (F=4, C=32, S=1, d_e=2 for the decoder cell)",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks_known else "fallback 6650",
            "whole_step_frac": None, "per_kernel": per_group}
    return roof


def _time_samples(step, samples):
    """(median, mean) ms per sample; every sample bracketed by its own pair of CUDA events."""
    torch.cuda.synchronize()
    evs = []
    for s_ in samples:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(*s_)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2], sum(t) / len(t)


def extra_dynamic_quadtree(dev, mask, cube, clim, dropout, hbm):
    """configs[2]: the ice grid with the quadtree rebuilt every forecast step (thresh 0.15, dist_from_05, 91 mesh builds per
    sample, data-dependent N / E: eager), and the graph build alone against the HBM roofline (section 8(d) bytes)."""
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.train import TrainStep
    kw = model_kwargs(dropout=dropout)
    kw["thresh"] = 0.15
    torch.manual_seed(21)
    model = q.Seq2Seq(**kw, device=dev).to(dev).train()
    step = TrainStep(model, mask, lr=1e-4, use_cuda_graph=False)
    n_days = max(1, cube.shape[0] - FRAMES)
    smp = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sample(cube, clim, d % n_days)] for d in range(8)]
    for s_ in smp[:3]:          # warm-up samples: lazy kernel attributes, allocator pool for the data-dependent mesh sizes
        step(*s_)
    ms, ms_mean = _time_samples(step, smp[3:])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the graph build alone
    img = q.add_positional_encoding(smp[0][0])
    build = lambda: q.image_to_graph(img, thresh=0.15, mask=mask, transform_func=dist_from_05, use_edge_attrs=True)
    g = build()
    N, E = int(g["data"].shape[1]), int(g["edge_index"].shape[1])
    torch.cuda.synchronize()
    iters = 20
    e0.record()
    for _ in range(iters):
        build()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    P, c, T = H * W, int(img.shape[-1]), int(img.shape[0])
    # 8(d): read image 4P per frame of the criterion channel (+ mask P), write labels 4P, edges 16E, attrs 4E*d_e; pool 4P*c in, 4N*c out
    nbytes = 4 * P * T + P + 4 * P + 16 * E + 4 * E * 2 + T * (4 * P * c + 4 * N * c)
    return {"workload": "configs[2]: 229x361, quadtree thresh 0.15 + dist_from_05 + mask, remesh every forecast step, 10+90 frames, "
                        "fwd+bwd+clip+Adam, eager (data-dependent mesh)",
            "graph_frames_per_s": FRAMES / (ms / 1e3), "ms_per_sample": ms, "ms_per_sample_mean": ms_mean,
            "timing": "median over 5 samples, each bracketed by CUDA events (a host-bound eager sample: the mean carries the box's hiccups)",
            "graph_build": {"us": us, "N": N, "E": E, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (us * 1e-6) / 1e9,
                            "frac_of_hbm_peak": nbytes / (us * 1e-6) / 1e9 / hbm,
                            "what": "image_to_graph of the 10 input frames (quadtree + pixel lists + pooling + adjacency + edge attributes), wall per call incl. its one host read-back"}}


def extra_cheb_dynamic(dev):
    """configs[0]-like (moving_mnist_example: 64 x 64 frames, Seq2Seq defaults = ChebConv K 3, hidden 16, 2 layers, quadtree thresh
    0.1 rebuilt every forecast step, 10 + 10 frames): fwd + bwd + clip + Adam per sample, eager (data-dependent mesh)."""
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200.train import TrainStep
    rng = np.random.default_rng(1)

    def blob(T=10, size=28):          # N(0, 0.05) noise + a sparse blob translating 1 px / frame (SURVEY 8(d) C1-style sample)
        x = rng.normal(0, 0.05, (T, 64, 64, 1)).astype(np.float32)
        u = rng.random((size, size)).astype(np.float32)
        b = (u > 0.6) * u
        for t in range(T):
            x[t, 5 + t:5 + t + size, 7 + t:7 + t + size, 0] += b
        return x

    smp = [[torch.from_numpy(a).to(dev) for a in (blob(), blob(), np.zeros((10, 64, 64, 1), np.float32))] for _ in range(10)]
    torch.manual_seed(1)
    model = q.Seq2Seq(hidden_size=16, dropout=0.0, thresh=0.1, input_timesteps=10, input_features=4, output_timesteps=10, n_layers=2,
                      device=dev).to(dev).train()
    step = TrainStep(model, np.zeros((64, 64), bool), lr=1e-4, use_cuda_graph=False)
    for s_ in smp[:3]:
        step(*s_)
    ms, ms_mean = _time_samples(step, smp[3:])
    return {"ms_per_sample_mean": ms_mean, "timing": "median over 7 samples, each bracketed by CUDA events",
            "workload": "configs[0]-like: 64x64 moving blob, Seq2Seq defaults (ChebConv K=3), hidden 16, 2 layers, dynamic quadtree "
                        "thresh 0.1, 10+10 frames, fwd+bwd+clip+Adam, eager", "graph_frames_per_s": 20 / (ms / 1e3), "ms_per_sample": ms}


def inference_setup(dev, mask, static_mesh=True):
    import quadtree_mpnnlstm_b200 as q
    torch.manual_seed(21)
    model = q.Seq2Seq(**model_kwargs(), device=dev).to(dev).eval()
    gs = (q.create_static_heterogeneous_graph((H, W), 4, mask, use_edge_attrs=True, resolution=1 / 12, device=dev)
          if static_mesh else None)
    return model, gs


def extra_inference(dev, mask, cube, clim, n_dates=16):
    """configs[4] (ice_inf.py:60): rollout inference on the static heterogeneous mesh (max cell 4), 10 + 90 frames per launch
    date, no_grad, the whole rollout one CUDA-graph replay per launch date; independent launch dates replayed concurrently on
    infer.RolloutPool's lanes (a rollout's persistent kernels are 32 CTAs on this mesh), next to one lane alone."""
    from quadtree_mpnnlstm_b200.infer import RolloutPool
    model, gs = inference_setup(dev, mask)
    xs = [[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sample(cube, clim, d % 4)] for d in range(4)]
    res = {}
    for label, lanes in (("one_lane", 1), ("pool", None)):
        pool = RolloutPool(model, mask, graph_structure=gs, lanes=lanes)
        pool.warm(xs[0][0], xs[0][2])
        xl, cl = [xs[d % 4][0] for d in range(n_dates)], [xs[d % 4][2] for d in range(n_dates)]
        pool.predict_many(xl, cl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pool.predict_many(xl, cl)
        e1.record()
        torch.cuda.synchronize()
        res[label] = (e0.elapsed_time(e1) / n_dates, len(pool.lanes), pool.launches_per_replay)
    ms, lanes, launches = res["pool"]
    return {"workload": "configs[4]: static heterogeneous mesh max cell 4 (N=%d, E=%d), 10+90 frames per launch date, no_grad, "
                        "one CUDA-graph replay per date, %d dates in flight on their own streams"
                        % (int(gs["mapping"].n_nodes), int(gs["edge_index"].shape[1]), lanes),
            "launch_dates_per_s": 1e3 / ms, "graph_frames_per_s": FRAMES * 1e3 / ms, "ms_per_launch_date": ms, "lanes": lanes,
            "one_lane_launch_dates_per_s": 1e3 / res["one_lane"][0], "qmp_launches_per_date": launches}


_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner to file
    descriptor 1 when NCCL_DEBUG asks for it -- and the environment's NCCL_DEBUG is left alone), so descriptor 1 is pointed
    at stderr for the run and the line goes to a private duplicate of the original stdout."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _init_dist(dev):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world


def _finish(world):
    """NCCL teardown with a captured graph still holding the communicator hung for the full time limit on a 2-GPU box (the JSON
    line had already been printed): synchronise, meet once more, and leave without it."""
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_gpu(args):
    import torch.distributed as dist
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    world = _init_dist(dev)
    _lib.lib()   # fail loudly if the CUDA library is missing
    _lib.set_pdl(bool(args.pdl))

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
        peaks = _peaks()
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        del step, model
        torch.cuda.empty_cache()
        times, N, E = per_kernel_times(dev, mask, cube, clim, args.dropout)
        roof = roofline_report(times, N, E, hbm, bool(peaks))
        # whole step against the 8(d) speed of light: every group's compulsory bytes x its frames, over the measured step
        ab = algorithmic_bytes(N, E)
        step_bytes = T_OUT * (ab["decoder_cell_fwd"] + ab["decoder_cell_bwd"] + ab["decoder_head_fwd"] + ab["decoder_head_bwd"]) \
            + T_IN * (ab["encoder_fwd"] + ab["encoder_bwd"])
        roof["whole_step_frac"] = round(step_bytes / (ms_total / args.steps * 1e-3) / 1e9 / hbm, 4)
        roof["whole_step_algorithmic_bytes"] = step_bytes
        line = {"metric": METRIC, "value": value, "unit": "graph-frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD % (N, E),
                           "dropout": args.dropout, "attention_dropout": 0.1, "cuda_graph": not args.no_graph, "pdl": bool(args.pdl),
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "l2": "inputs + saved activations per step (~10 GB) exceed the 126 MB L2; per-kernel times taken "
                                 "with a 256 MB flush write before every launch",
                           "final_loss": last},
                "e2e": {"value": e2e, "unit": "graph-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches,
                "clocks": clocks.summary(),
                "roofline": roof}
        if world == 1 and not args.no_extras:
            extra = {}
            dev_samples = hs = None          # the timed loop's inputs: release them (and the allocator's cache) before the extras
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            # (the eager dynamic-mesh steps switch the allocator to expandable segments: they run after the captured rollouts)
            for name, fn in (("inference", lambda: extra_inference(dev, mask, cube, clim)),
                             ("dynamic_quadtree", lambda: extra_dynamic_quadtree(dev, mask, cube, clim, args.dropout, hbm)),
                             ("cheb_dynamic", lambda: extra_cheb_dynamic(dev))):
                try:
                    extra[name] = fn()
                except Exception as exc:        # an extra must never take the headline line down with it
                    extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.empty_cache()
            line["extra"] = extra
        if world == 1 and not args.no_cpu:
            t_in, t_out, _ = oracle_frames_for_memory()
            rate, dt, done, what = cpu_oracle_rate(os.cpu_count() or 1, max_samples=1, dropout=args.dropout, t_in=t_in, t_out=t_out)
            line["cpu_baseline"] = {"value": rate, "unit": "graph-frames/s", "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": what, "same_config": (t_in, t_out) == (T_IN, T_OUT),
                                    "same_frame_mix": True}
        _emit(line)
    _finish(world)


def run_infer(args):
    """configs[4] at N GPUs: launch dates sharded round-robin over the ranks (infer.predict_sharded: no data-path
    collective, one all_gather of the forecasts at the end), every rank replaying its captured rollout per date."""
    import torch.distributed as dist
    from quadtree_mpnnlstm_b200 import _lib
    from quadtree_mpnnlstm_b200.infer import Rollout, RolloutPool, predict_sharded
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    world = _init_dist(dev)
    _lib.lib()
    mask = ocean_mask()
    per_rank = max(args.steps, 1)
    n_dates = per_rank * world
    cube = synthetic_cube(FRAMES + 8)
    clim = np.ascontiguousarray(cube[..., :1].mean(0, keepdims=True).repeat(366, 0))
    model, gs = inference_setup(dev, mask, static_mesh=not args.pixel_mesh)
    if args.no_graph:
        ro = Rollout(model, mask, graph_structure=gs, use_cuda_graph=False)
    else:       # independent launch dates in flight on their own streams (one lane on the pixel-wise mesh: it fills the GPU)
        ro = RolloutPool(model, mask, graph_structure=gs, lanes=args.lanes or None)
    pinned = [[torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in sample(cube, clim, d)] for d in range(4)]

    def load(d):                # host -> device inside the timed region (e2e: inputs start in pinned host memory)
        x, _, cl = pinned[d % 4]
        return x.to(dev, non_blocking=True), cl.to(dev, non_blocking=True)

    if isinstance(ro, RolloutPool):
        ro.warm(*load(0))
        ro.predict_many(*zip(*[load(d) for d in range(2 * len(ro.lanes))]))
    else:
        for _ in range(max(args.warmup, 4)):        # eager warm-ups
            ro(*load(0))
    # the result buffer of a serving loop: pinned once, reused for every batch of launch dates
    host_out = torch.empty((per_rank, T_OUT, H, W, 1), dtype=torch.float32).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.kernel_launches()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1 and isinstance(ro, RolloutPool):
            # every lane copies its forecast into the pinned result buffer in its own stream order, under the other lanes' replays
            pairs = [load(d) for d in range(n_dates)]
            host = ro.predict_many([p[0] for p in pairs], [p[1] for p in pairs], out=host_out)
        else:
            full = predict_sharded(model, load, n_dates, mask, rank=rank, world=world, rollout=ro)
            host = host_out.copy_(full[rank::world] if world > 1 else full, non_blocking=True)   # the forecasts come back to the host
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(ms) / 1e3
        launches = ro.launches_per_replay * per_rank if ro.graph is not None else _lib.kernel_launches() - launches0
        mesh = "static heterogeneous mesh max cell 4" if gs is not None else "pixel-wise mesh"
        N = int(gs["mapping"].n_nodes) if gs is not None else 47200
        line = {"metric": "launch-dates/sec rollout inference (10+90 frames per date)", "value": n_dates / sec, "unit": "launch-dates/s",
                "graph_frames_per_s": n_dates * FRAMES / sec, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sec / per_rank, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "mode": "infer",
                "config": {"workload": f"ice_inf (configs[4]): 229x361, {mesh} N={N}, TransformerConv hidden 32, 10+90 frames per "
                                       f"launch date, no_grad, {per_rank} launch dates per GPU, forecasts copied into a pinned host buffer",
                           "cuda_graph": ro.graph is not None, "parallelism": f"dates/{world}",
                           "lanes": len(ro.lanes) if isinstance(ro, RolloutPool) else 1},
                "e2e": {"value": n_dates / sec, "unit": "launch-dates/s",
                        "h2d_bytes_per_step": sum(t.numel() * 4 for t in (pinned[0][0], pinned[0][2])),
                        "d2h_bytes_per_step": int(host[0].numel() * 4)},
                "gpu_launches": launches, "clocks": clocks.summary()}
        _emit(line)
    _finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--dropout", type=float, default=0.1, help="decoder dropout (ice_exp.py:155 trains with 0.1); the "
                    "TransformerConv attention dropout is 0.1 in train mode regardless, as in the reference")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.dynamic_quadtree / extra.inference")
    ap.add_argument("--pdl", type=int, default=0, help="programmatic dependent launch of the hot kernels (qmp_set_pdl); 0 = plain stream order")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying a CUDA graph")
    ap.add_argument("--lanes", type=int, default=0, help="--mode infer: launch dates in flight per GPU (0 = by mesh size)")
    ap.add_argument("--pixel-mesh", action="store_true", help="--mode infer on the pixel-wise mesh (N = 47 200) instead of configs[4]'s")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "infer":
        run_infer(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
