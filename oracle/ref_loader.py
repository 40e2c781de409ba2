"""Import the UNMODIFIED reference from /root/reference (build container only).  TEST INFRASTRUCTURE.

The reference is a flat script repo whose hot-path files import ``torch_geometric``,
``matplotlib`` and ``torchviz``; none is installed here.  ``load_reference()`` puts the stubs of
``oracle/ref_shim`` and ``/root/reference`` on ``sys.path`` and imports the reference's own
``model`` package.  Nothing is copied or edited.  Returns ``None`` when the reference is absent
(the GPU box), so callers must skip.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QMP_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "seq2seq.py"))


def load_reference():
    """Returns a namespace with .graph_functions .seq2seq .model .utils (reference modules)."""
    if not reference_available():
        return None
    for p in (_REPO, _SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    cached = sys.modules.get("model")
    if cached is not None and not (getattr(cached, "__file__", None) or "").startswith(REFERENCE_ROOT) \
            and REFERENCE_ROOT not in "".join(getattr(cached, "__path__", [])):
        raise RuntimeError("a different top-level 'model' package is already imported")
    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("model.utils")
    ns.graph_functions = importlib.import_module("model.graph_functions")
    ns.model = importlib.import_module("model.model")
    ns.seq2seq = importlib.import_module("model.seq2seq")
    return ns
