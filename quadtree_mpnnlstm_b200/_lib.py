"""ctypes binding of ``libqmp_b200.so`` -- the C-ABI boundary declared in ``include/qmp_b200.h``.

There is no CPU fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.  PyTorch is used only to own device memory and streams; every entry
point takes raw device pointers, sizes and a ``cudaStream_t``.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMP_LIB_PATH") or os.path.join(_HERE, "libqmp_b200.so")   # override: kernel experiments (scripts/)

_P, _I, _L, _F, _D, _U = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_uint64
_CODES = {"p": _P, "i": _I, "l": _L, "f": _F, "d": _D, "u": _U}

# name -> argument codes (p pointer, i int, l int64, f float, d double, u uint64); all return int
SIGNATURES = {
    "qmp_exclusive_scan_i32": "ppippp",
    "qmp_frame_max_pad": "piiiiiipp",
    "qmp_quadtree_labels": "pppiiiidppppppppp p".replace(" ", ""),
    "qmp_mesh_pixels_from_rects": "piippp ipppp p".replace(" ", ""),
    "qmp_mesh_pixelwise": "pipppppppp p".replace(" ", ""),
    "qmp_segment_sum": "piiipppipiipp",
    "qmp_quadtree_graph": "piiiipppiidfippp",
    "qmp_quadtree_graph_export": "piiiiiiiipppp",
    "qmp_gather_by_label": "piiiippifpp",
    "qmp_regrid": "ppiiiippifpppiippp",
    "qmp_cheb_cell_fwd": "iiiiii" + "p" * 16,
    "qmp_cheb_cell_bwd": "iiiiii" + "p" * 15 + "ii" + "ppp",
    "qmp_cheb_stack_fwd": "i" * 11 + "p" * 13,
    "qmp_cheb_stack_bwd": "i" * 11 + "p" * 14 + "i" + "pp",
    "qmp_adjacency_quadtree": "piippppppplppppp",
    "qmp_adjacency_pixelwise": "piippppppppp",
    "qmp_edge_attrs": "ppipppiiifipp",
    "qmp_add_positional_encoding": "piiiipp",
    "qmp_csr_from_edge_index": "plippppppppppppp p".replace(" ", ""),
    "qmp_gather_rows": "pplipp",
    "qmp_gemm": "ppppiiiiiilllliiiip",
    "qmp_gemm_tn_acc": "pppiiiiiillliip",
    "qmp_attn_fwd": "iiipppp iippppp fup".replace(" ", ""),
    "qmp_attn_bwd_target": "iiipppp iippppp pfup".replace(" ", ""),
    "qmp_attn_bwd_source": "iiipppppppppp iiiifup".replace(" ", ""),
    "qmp_edge_norm": "iipppppp ppp".replace(" ", ""),
    "qmp_spmm": "iippppp iffpipip".replace(" ", ""),
    "qmp_lstm_gates_fwd": "iipippiiifppppppipp",
    "qmp_lstm_gates_bwd": "iippppiiifppppipippp",
    "qmp_gru_gates1_fwd": "lppppppppp",
    "qmp_gru_gates1_bwd": "lpppppppppp",
    "qmp_gru_gates2_fwd": "lppppppp",
    "qmp_gru_gates2_bwd": "lpppppppp",
    "qmp_head_finish_fwd": "ppiiifuppp",
    "qmp_head_finish_bwd": "pppppiiifuppp",
    "qmp_relu_mask": "pplp",
    "qmp_relu_mask_to": "ppplp",
    "qmp_fused_fwd": "ippppiiippiiiipiiipippiiifppppppippppfup",
    "qmp_fused_bwd_target": "ippppiiippiiiipiipippppppppppfup",
    "qmp_fused_bwd_source": "ippppiiippiiiipiipippppppfup",
    "qmp_fused_wgrad": "ipiiipiiiiiipippppppp",
    "qmp_fused_fwd_tc": "ippppiiippiiiipiiipippiiifppppppippppfup",
    "qmp_fused_pack_tc": "piiipp",
    "qmp_fused_bwd_target_tc": "ippppiiippiiiipiipippppppppppfup",
    "qmp_fused_bwd_source_tc": "ippppiiippiiiipiipippppppfup",
    "qmp_fused_bwd_onepass_tc": "ippppiiippiiiipiipippppppppppfup",
    "qmp_fused_pack_cell": "pppp",
    "qmp_fused_pack_cell_bwd": "pppp",
    "qmp_fused_cell_bwd": "ipppp" "i" "p" "i" "ppp" "i" "pppp" "iiif" "pppp" "i" "pp" "ppp" "pppp" "pp" "fup",
    "qmp_cell_wgrad": "ipipippppppp",
    "qmp_panel_wgrad": "ipiiipipppp",
    "qmp_fused_wgrad_tma": "ipiiipiiiiiipippppppp",
    "qmp_head_tail_fwd": "ipppp" "i" "pp" "ii" "fufu" "pppp" "p",
    "qmp_head_tail_bwd": "ipppp" "i" "ppppp" "ii" "fufu" "ppp" "p" "i" "i" "pp" "p",
    "qmp_pack_head_bwd": "ppp",
    "qmp_head_bwd": "ipppp" "i" "p" "p" "i" "ppp" "pp" "p" "fup",
    "qmp_pack_tconv_fwd": "piiiipp",
    "qmp_pack_tconv_bwd": "piiiippp",
    "qmp_gat_fwd": "iiippp" "pi" "ppp" "pipp" "f" "pi" "p" "p",
    "qmp_gat_bwd": "iiippp" "pi" "ppp" "pipp" "f" "p" "pi" "p" "pppp" "ppp" "p",
    "qmp_tconv1_fwd": "ipppp" "i" "ppp" "fup",
    "qmp_tconv1_bwd": "ipppp" "i" "ppppp" "i" "p" "fup",
    "qmp_fused_cell_fwd": "ippppipippp" "iiif" "pppppp" "i" "ppppp" "fup",
}


# kernels launched by one call of each entry point (for bench.py's gpu_launches accounting)
KERNELS_PER_CALL = {
    "qmp_exclusive_scan_i32": 3, "qmp_frame_max_pad": 1, "qmp_quadtree_labels": 3, "qmp_mesh_pixels_from_rects": 5,
    "qmp_mesh_pixelwise": 5, "qmp_segment_sum": 1, "qmp_quadtree_graph": 1, "qmp_quadtree_graph_export": 1, "qmp_gather_by_label": 1, "qmp_regrid": 1, "qmp_cheb_cell_fwd": 14, "qmp_cheb_cell_bwd": 21, "qmp_cheb_stack_fwd": 8, "qmp_cheb_stack_bwd": 12, "qmp_adjacency_quadtree": 8,
    "qmp_adjacency_pixelwise": 5, "qmp_edge_attrs": 1, "qmp_add_positional_encoding": 1,
    "qmp_csr_from_edge_index": 16, "qmp_gather_rows": 1, "qmp_gemm": 1, "qmp_gemm_tn_acc": 1, "qmp_attn_fwd": 1,
    "qmp_attn_bwd_target": 1, "qmp_attn_bwd_source": 1, "qmp_edge_norm": 2, "qmp_spmm": 1, "qmp_lstm_gates_fwd": 1,
    "qmp_lstm_gates_bwd": 1, "qmp_head_finish_fwd": 1, "qmp_head_finish_bwd": 1, "qmp_relu_mask": 1, "qmp_relu_mask_to": 1, "qmp_fused_fwd": 1, "qmp_fused_bwd_target": 1, "qmp_fused_bwd_source": 1,
    "qmp_fused_wgrad": 1, "qmp_fused_fwd_tc": 1, "qmp_fused_pack_tc": 1, "qmp_fused_bwd_target_tc": 1, "qmp_fused_bwd_source_tc": 1, "qmp_fused_bwd_onepass_tc": 1,
    "qmp_fused_pack_cell": 1, "qmp_fused_cell_fwd": 1, "qmp_gat_fwd": 1, "qmp_gat_bwd": 1, "qmp_tconv1_fwd": 2, "qmp_tconv1_bwd": 2, "qmp_head_tail_fwd": 2, "qmp_head_tail_bwd": 2, "qmp_fused_pack_cell_bwd": 1, "qmp_fused_cell_bwd": 1, "qmp_cell_wgrad": 1, "qmp_panel_wgrad": 1, "qmp_fused_wgrad_tma": 1,
}
CALL_COUNTS = {}


def kernel_launches():
    """Kernels launched through this binding so far (sum over entry points of calls x kernels per call)."""
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in CALL_COUNTS.items())


class QmpError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise QmpError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        L = ctypes.CDLL(LIB_PATH)
        L.qmp_last_error.restype = ctypes.c_char_p
        L.qmp_last_error.argtypes = []
        L.qmp_version.restype = _I
        L.qmp_quadtree_pyramid_cells.restype = _L
        L.qmp_quadtree_pyramid_cells.argtypes = [_I, _I, _I]
        L.qmp_quadtree_graph_scratch_bytes.restype = _L
        L.qmp_quadtree_graph_scratch_bytes.argtypes = [_I, _I, _I, _I, _I]
        L.qmp_set_tensor_cores.restype = _I
        L.qmp_set_tensor_cores.argtypes = [_I]
        L.qmp_set_fused_paired.restype = _I
        L.qmp_set_fused_paired.argtypes = [_I]
        L.qmp_fused_tc_image_bytes.restype = _L
        L.qmp_fused_tc_image_bytes.argtypes = [_I, _I]
        L.qmp_fused_cell_image_bytes.restype = _L
        L.qmp_fused_cell_image_bytes.argtypes = []
        L.qmp_fused_cell_bwd_image_bytes.restype = _L
        L.qmp_fused_cell_bwd_image_bytes.argtypes = []
        L.qmp_head_bwd_image_bytes.restype = _L
        L.qmp_head_bwd_image_bytes.argtypes = []
        L.qmp_set_dropout_salt.restype = _I
        L.qmp_set_dropout_salt.argtypes = [_P]
        L.qmp_set_pdl.restype = _I
        L.qmp_set_pdl.argtypes = [_I]
        for name, sig in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = _I
            fn.argtypes = [_CODES[c] for c in sig]
        _lib = L
    return _lib


def on_device(t):
    """True when ``t`` lives on a CUDA device (the only place this package computes)."""
    return bool(t.is_cuda)


def capturing():
    """True while the current CUDA stream is being captured into a graph (host read-backs are illegal then)."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise QmpError("qmp_b200 kernels take CUDA tensors only (no CPU fallback)")
        return t.data_ptr()
    return t


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw-handle accessors cost a fraction of a
    microsecond; torch.cuda.current_stream() builds a Stream object per call, which an eager dynamic-mesh sample pays
    ~3 000 times)."""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


_Tensor = torch.Tensor


def call(name, *args):
    """Invoke ``qmp_<name>`` on the current CUDA stream (appended as the last argument)."""
    L = _lib if _lib is not None else lib()
    fn = getattr(L, name)
    conv = []
    for a in args:
        if isinstance(a, _Tensor):
            if not a.is_cuda:
                raise QmpError("qmp_b200 kernels take CUDA tensors only (no CPU fallback)")
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    CALL_COUNTS[name] = CALL_COUNTS.get(name, 0) + 1
    rc = fn(*conv, stream_ptr())
    if rc != 0:
        raise QmpError(f"{name} failed (code {rc}): {L.qmp_last_error().decode(errors='replace')}")


def set_dropout_salt(t):
    """Point the seeded kernels at a device uint64 salt (``t``: int64 CUDA tensor with one element) or clear it (None);
    see qmp_set_dropout_salt (csrc/core.cu)."""
    L = lib()
    fn = getattr(L, "qmp_set_dropout_salt", None)
    if fn is not None:
        fn(_ptr(t))


def set_pdl(on):
    """Programmatic dependent launch of the hot kernels on / off (default off); returns the previous setting (csrc/core.cu)."""
    return bool(lib().qmp_set_pdl(1 if on else 0))


def exported_symbols():
    return ["qmp_set_dropout_salt", "qmp_set_pdl", "qmp_last_error", "qmp_version", "qmp_quadtree_pyramid_cells", "qmp_quadtree_graph_scratch_bytes", "qmp_set_tensor_cores", "qmp_fused_tc_image_bytes", "qmp_set_fused_paired", "qmp_fused_cell_image_bytes", "qmp_fused_cell_bwd_image_bytes", "qmp_head_bwd_image_bytes"] + list(SIGNATURES)
