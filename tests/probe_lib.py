"""ctypes access to csrc/probes/libqmp_probe.so -- hardware-convention probes and timing hooks (test infrastructure; the
product library libqmp_b200.so does not contain them and nothing under quadtree_mpnnlstm_b200/ loads this)."""
import ctypes
import os

import torch

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quadtree_mpnnlstm_b200", "csrc", "probes",
                    "libqmp_probe.so")
_CODES = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float}
SIGNATURES = {
    "qmp_tc_gemm_probe": "pppiiiip",
    "qmp_tc_probe2": "pppiiip",
    "qmp_tc_probe3": "piiiip",
    "qmp_mn_probe": "ppppiiiiiiiip",
}
_lib = None


def call(name, *args):
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
        _lib.qmp_last_error.restype = ctypes.c_char_p
    fn = getattr(_lib, name)
    fn.restype = ctypes.c_int
    fn.argtypes = [_CODES[c] for c in SIGNATURES[name]]
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    rc = fn(*conv, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {_lib.qmp_last_error().decode(errors='replace')}")
