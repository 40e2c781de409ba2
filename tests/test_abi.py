"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/qmp_b200.h declares;
the ctypes table agrees with the sources.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from quadtree_mpnnlstm_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "qmp_b200.h")).read()
    declared = re.findall(r"\b(qmp_\w+)\s*\(", header)
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_lib.exported_symbols())
    lib.qmp_version.restype = ctypes.c_int
    assert lib.qmp_version() == 100


def test_ctypes_signatures_match_sources():
    sys.path.insert(0, os.path.join(ROOT, "quadtree_mpnnlstm_b200", "csrc"))
    import gen_abi
    from quadtree_mpnnlstm_b200 import _lib
    protos = {name: "".join(gen_abi.code_of(t) for t, _ in args) for _, name, ret, args in gen_abi.prototypes() if ret == "int" and name not in ("qmp_version", "qmp_set_tensor_cores", "qmp_set_fused_paired", "qmp_set_dropout_salt", "qmp_set_pdl")}
    assert protos == _lib.SIGNATURES


def test_sass_is_sm100a_only():
    import subprocess
    from quadtree_mpnnlstm_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    import quadtree_mpnnlstm_b200 as q
    from quadtree_mpnnlstm_b200._lib import QmpError
    with pytest.raises(QmpError):
        q.add_positional_encoding(torch.zeros(1, 4, 4, 1))
    with pytest.raises(QmpError):
        q.image_to_graph(torch.zeros(1, 8, 8, 3), thresh=0.5, max_grid_size=8)
    m = q.Seq2Seq(hidden_size=8, dropout=0.0, thresh=0.1)
    with pytest.raises(QmpError):
        m(torch.zeros(3, 8, 8, 1))


def test_product_does_not_import_oracle():
    import subprocess
    code = "import sys; import quadtree_mpnnlstm_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
