"""Build libqmp_b200.so (nvcc, sm_100a only) and regenerate include/qmp_b200.h from the sources.

    python quadtree_mpnnlstm_b200/csrc/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is written in-tree next to the Python package so that it
travels to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libqmp_b200.so")
PROBE_OUT = os.path.join(HERE, "probes", "libqmp_probe.so")      # test hooks only: NOT part of the product library / ABI
HEADER = os.path.join(REPO, "include", "qmp_b200.h")

NVCC_FLAGS = os.environ.get("QMP_EXTRA_FLAGS", "").split() + (["-DQMP_TIMING_TF32X1"] if os.environ.get("QMP_TIMING_TF32X1") else []) + (["-DQMP_PW_TRACE"] if os.environ.get("QMP_PW_TRACE") else []) + (["-DQMP_CELL_TRACE"] if os.environ.get("QMP_CELL_TRACE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]

REPLACES = {  # entry point -> reference interface it stands in for
    "qmp_frame_max_pad": "model/graph_functions.py:632 (max over frames of channel 0) and :190 (edge pad)",
    "qmp_quadtree_labels": "model/graph_functions.py:145-259 quadtree_decompose",
    "qmp_quadtree_pyramid_cells": "(scratch sizing for qmp_quadtree_labels)",
    "qmp_quadtree_graph": "model/graph_functions.py:590-681 image_to_graph (quadtree_decompose :145-259, get_mapping :555-587, flatten :391-419, get_adj :261-345, dist_angle :358-370) + PyG's per-edge scatter (the CSR) -- one cooperative launch",
    "qmp_quadtree_graph_export": "(copies the compacted result of qmp_quadtree_graph out of its arena into exact-size buffers)",
    "qmp_quadtree_graph_scratch_bytes": "(arena sizing for qmp_quadtree_graph)",
    "qmp_mesh_pixels_from_rects": "model/graph_functions.py:555-587 get_mapping (+ :649 to_dense)",
    "qmp_mesh_pixelwise": "model/graph_functions.py:511-525 pixel-wise labels / graph_nodes / n_pixels_per_node",
    "qmp_segment_sum": "model/graph_functions.py:391-419 flatten (and the backward of unflatten)",
    "qmp_gather_by_label": "model/graph_functions.py:451-468 unflatten (and the backward of flatten)",
    "qmp_cheb_cell_fwd": "model/model.py:430-447 GConvLSTM gate pre-activations with ChebConv / GCNConv stacks (GraphConv :60-97; PyG ChebConv / GCNConv forward) -- the launch sequence of one cell step issued from C++",
    "qmp_cheb_cell_bwd": "autograd of the same (weight / bias gradients accumulated in place, dX, dH)",
    "qmp_cheb_stack_fwd": "model/seq2seq.py:182-187 decoder head fc_out2(relu(fc_out1(.))) and model/model.py:60-97 GraphConv with ChebConv / GCNConv layers -- the launch sequence of the chain issued from C++",
    "qmp_cheb_stack_bwd": "autograd of the same",
    "qmp_regrid": "model/seq2seq.py:434-491 do_remesh: unflatten (graph_functions.py:451-468) + flatten (:391-419) of a recurrent state as one pass, and their backward",
    "qmp_adjacency_quadtree": "model/graph_functions.py:261-345 get_adj",
    "qmp_adjacency_pixelwise": "model/graph_functions.py:471-493 get_adj_pixelwise",
    "qmp_edge_attrs": "model/graph_functions.py:358-370 dist / dist_angle (+ :347-353, :657)",
    "qmp_add_positional_encoding": "model/utils.py:30-52 add_positional_encoding",
    "qmp_csr_from_edge_index": "PyG MessagePassing.propagate's per-edge scatter (model/model.py:96 call site)",
    "qmp_gather_rows": "(edge payload reordering for the CSR)",
    "qmp_gemm": "PyG Linear inside GCNConv/ChebConv/TransformerConv (model/model.py:39-57, :96)",
    "qmp_gemm_tn_acc": "autograd weight gradients of those Linears",
    "qmp_attn_fwd": "PyG TransformerConv.message/softmax/aggregate (model/model.py:51, :96)",
    "qmp_attn_bwd_target": "autograd of TransformerConv message passing (target side)",
    "qmp_attn_bwd_source": "autograd of TransformerConv message passing (source side)",
    "qmp_edge_norm": "PyG gcn_norm / ChebConv.__norm__ (model/model.py:50, :53)",
    "qmp_spmm": "PyG GCNConv/ChebConv propagate (model/model.py:96)",
    "qmp_lstm_gates_fwd": "model/model.py:394-463 GConvLSTM gates; model/seq2seq.py:59-66, 138-165 norms + head input",
    "qmp_lstm_gates_bwd": "autograd of the above",
    "qmp_gru_gates1_fwd": "model/model.py:240-250 GConvGRU update / reset gates and H * R",
    "qmp_gru_gates1_bwd": "autograd of the above",
    "qmp_gru_gates2_fwd": "model/model.py:251-258 GConvGRU candidate state and H'",
    "qmp_gru_gates2_bwd": "autograd of the above",
    "qmp_head_finish_fwd": "model/seq2seq.py:167-178 (dropout, tanh, residual, sigmoid) and :427-428 (next input)",
    "qmp_head_finish_bwd": "autograd of the above",
    "qmp_relu_mask": "model/seq2seq.py:184 F.relu backward",
    "qmp_relu_mask_to": "model/seq2seq.py:184 F.relu backward (out of place)",
    "qmp_fused_fwd": "model/model.py:394-463 GConvLSTM.forward around PyG TransformerConv + model/seq2seq.py:59-66, 138-165 (one launch)",
    "qmp_fused_bwd_target": "autograd of the above, target side",
    "qmp_fused_bwd_source": "autograd of the above, source side",
    "qmp_fused_fwd_tc": "model/model.py:394-463 GConvLSTM.forward around PyG TransformerConv + model/seq2seq.py:59-66, 138-165 (one launch, dense contractions on tcgen05)",
    "qmp_fused_bwd_target_tc": "autograd of qmp_fused_fwd_tc, target side (dense contractions on tcgen05)",
    "qmp_fused_bwd_source_tc": "autograd of qmp_fused_fwd_tc, source side (dense contractions on tcgen05)",
    "qmp_fused_bwd_onepass_tc": "autograd of qmp_fused_fwd_tc, target and source side of every edge in one launch (source rows by vector reductions)",
    "qmp_fused_pack_tc": "(weight images for the tcgen05 fused kernels: hi / lo TF32 split, canonical K-major layout)",
    "qmp_fused_tc_image_bytes": "(size of one conv's weight image)",
    "qmp_fused_wgrad": "autograd weight gradients of the above (tcgen05 3xTF32 reduction over the mesh nodes)",
    "qmp_fused_cell_fwd": "model/model.py:394-463 GConvLSTM.forward of the decoder cell (4 X convs + 4 H convs, gates, norms, head input) -- one persistent launch, gates batched on tcgen05, edge phase 8 lanes per node",
    "qmp_fused_pack_cell": "(weight image of qmp_fused_cell_fwd: the eight convs of the decoder cell side by side)",
    "qmp_fused_cell_image_bytes": "(size of that image)",
    "qmp_fused_cell_bwd": "autograd of qmp_fused_cell_fwd w.r.t. X and H (target and source side of every edge in one persistent launch) + rows for qmp_cell_wgrad",
    "qmp_cell_wgrad": "autograd weight gradients of the decoder cell's eight convs (one streaming launch: TMA panels as MN-major tcgen05 operands, 3xTF32)",
    "qmp_fused_wgrad_tma": "autograd weight gradients of a fused layer group (same contract as qmp_fused_wgrad): streaming TMA panels as MN-major tcgen05 operands, one product per conv and 8 nodes",
    "qmp_panel_wgrad": "autograd weight gradients of one wide-input TransformerConv (the decoder head's fc_out1, model/seq2seq.py:117-121): streaming TMA panels -> one tcgen05 product per 8 nodes",
    "qmp_fused_pack_cell_bwd": "(weight image of qmp_fused_cell_bwd)",
    "qmp_fused_cell_bwd_image_bytes": "(size of that image)",
    "qmp_tconv1_fwd": "PyG TransformerConv(hidden, 1) = the decoder's fc_out2 (model/seq2seq.py:117-121, 182-187): scalar query / key / value records",
    "qmp_head_tail_fwd": "model/seq2seq.py:167-187, 427-428: fc_out2 (TransformerConv hidden -> 1) + dropout, tanh, residual, sigmoid, next input -- two launches",
    "qmp_head_tail_bwd": "autograd of the above (incl. the relu mask of fc_out1's output, model/seq2seq.py:184)",
    "qmp_head_bwd": "autograd of the decoder head conv fc_out1 (model/seq2seq.py:117-121, 182-187) w.r.t. its input + rows for the weight gradients: persistent launch, octet edge phase, source side by vector reductions",
    "qmp_pack_head_bwd": "(weight image of qmp_head_bwd)",
    "qmp_head_bwd_image_bytes": "(size of that image)",
    "qmp_pack_tconv_fwd": "(weight pack of a TransformerConv group from the PyG-named parameters, model/model.py:51: one launch instead of ~40 tensor ops)",
    "qmp_pack_tconv_bwd": "autograd of the above (contiguous per-parameter gradient slices)",
    "qmp_gat_fwd": "PyG GATConv / GATv2Conv edge phase, one head (model/model.py:43-44, 55-56): additive-attention segment softmax + aggregate",
    "qmp_gat_bwd": "autograd of the above",
    "qmp_tconv1_bwd": "autograd of the above (input gradient + parameter gradients)",
    "qmp_set_fused_paired": "(switch: two threads per node (paired warps) or one in the tcgen05 fused kernels)",
    "qmp_set_pdl": "(switch: programmatic dependent launch of the hot kernels; stream order is what eager PyTorch gives the reference)",
    "qmp_set_tensor_cores": "(switch: tcgen05 3xTF32 contractions on/off; parity tests run both)",
    "qmp_exclusive_scan_i32": "(utility: numpy cumsum at model/graph_functions.py:511)",
    "qmp_last_error": "(error text; the reference raises Python exceptions)",
    "qmp_version": "(ABI version)",
}


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def write_header():
    decls = []
    for path in sources():
        src = open(path).read()
        for m in re.finditer(r"((?:^//[^\n]*\n)*)QMP_API\s+([\w\s\*]+?)\s*\b(qmp_\w+)\s*\(([^)]*)\)\s*\{", src, re.M):
            comment, ret, name, args = m.group(1), m.group(2).strip(), m.group(3), " ".join(m.group(4).split())
            decls.append((os.path.basename(path), comment, ret, name, args or "void"))
    lines = [
        "/* qmp_b200.h -- C ABI of libqmp_b200.so (GENERATED by quadtree_mpnnlstm_b200/csrc/build.py; do not edit).",
        " *",
        " * Drop-in boundary for the hot path of zach-gousseau/Quadtree-MPNNLSTM (quadtree graph build ->",
        " * graph-conv LSTM cell -> seq2seq driver).  The reference has no FFI: its boundary is the Python",
        " * module surface (model/seq2seq.py, model/model.py, model/graph_functions.py, model/utils.py).  The",
        " * Python mirror in quadtree_mpnnlstm_b200/ keeps that surface and calls the entry points below through",
        " * ctypes; INTEGRATION.md shows the binding.",
        " *",
        " * Conventions: every pointer is a DEVICE pointer borrowed from the caller (no ownership transfer,",
        " * outputs and scratch pre-allocated by the caller); sizes are plain ints; `stream` is a cudaStream_t;",
        " * calls are asynchronous on that stream, re-entrant per stream, and keep no global state.  Return",
        " * value 0 = success, otherwise a cudaError_t or -1 (argument error); qmp_last_error() gives the text",
        " * (thread-local).  Data-dependent sizes (node / edge counts) are written to device ints and read back",
        " * once by the host.  float = IEEE binary32 everywhere; edge_index is int64 [2, E] (row 0 = source,",
        " * row 1 = target), as in the reference.",
        " */",
        "#ifndef QMP_B200_H",
        "#define QMP_B200_H",
        "#include <stdint.h>",
        "#ifdef __cplusplus",
        'extern "C" {',
        "#endif",
        "",
    ]
    for fname, comment, ret, name, args in decls:
        lines.append(f"/* [{fname}] replaces: {REPLACES.get(name, '?')}")
        for c in comment.strip().splitlines():
            lines.append(" * " + c.lstrip("/ ").rstrip())
        lines.append(" */")
        lines.append(f"{ret} {name}({args});")
        lines.append("")
    lines += ["#ifdef __cplusplus", "}", "#endif", "#endif /* QMP_B200_H */", ""]
    os.makedirs(os.path.dirname(HEADER), exist_ok=True)
    text = "\n".join(lines)
    if not os.path.isfile(HEADER) or open(HEADER).read() != text:
        open(HEADER, "w").write(text)


OBJ_DIR = os.path.join(HERE, "_obj")


def _deps_mtime():
    return max(os.path.getmtime(p) for p in glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(HERE, "*.inl")))


def build(force=False, verbose=False):
    """Compile every .cu to an object (in parallel, only what changed) and link libqmp_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    write_header()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _deps_mtime()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = ["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, res in ex.map(compile_one, jobs):
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
                if verbose:
                    sys.stderr.write(res.stderr)
    if jobs or not os.path.isfile(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        res = subprocess.run(["nvcc"] + NVCC_FLAGS + objs + ["-o", OUT], capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking libqmp_b200.so")
    build_probes(force)
    return OUT


def build_probes(force=False):
    """csrc/probes/*.cu -> csrc/probes/libqmp_probe.so: hardware-convention probes and timing hooks used by tests/ and
    scripts/ only (tests/probe_lib.py); nothing in the product imports or links it."""
    srcs = sorted(glob.glob(os.path.join(HERE, "probes", "*.cu"))) + [os.path.join(HERE, "core.cu")]
    deps = srcs + glob.glob(os.path.join(HERE, "*.cuh"))
    if not force and os.path.isfile(PROBE_OUT) and os.path.getmtime(PROBE_OUT) >= max(os.path.getmtime(p) for p in deps):
        return PROBE_OUT
    res = subprocess.run(["nvcc"] + NVCC_FLAGS + srcs + ["-o", PROBE_OUT], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libqmp_probe.so")
    return PROBE_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
