"""Run the oracle in float64.  TEST INFRASTRUCTURE.

The oracle restates the reference in the reference's own arithmetic (float32).  Over a 90-step recurrent rollout float32
rounding differences between two correct implementations are amplified step by step, so a parity bar on late forecast steps
needs the size of that amplification: ``float64()`` evaluates the SAME restated algorithm in double precision (the graph half
keeps producing float32 tensors like the reference does; they are widened here), which gives the exact-arithmetic result of
the reference's algorithm.  ``|oracle_f32 - oracle_f64|`` is then the reference's own float32 rounding noise at each step,
and ``|cuda - oracle_f64|`` the CUDA path's.  Used by tests/test_parity_fullsize.py only.
"""
from __future__ import annotations

import contextlib

import torch

from . import graph_ref as G


@contextlib.contextmanager
def float64():
    """Inside: ``graph_ref`` hands out float64 node data / edge attributes / positional encodings and torch's default dtype
    is float64 (zero-initialised states).  Move the model with ``model.double()`` and pass float64 inputs."""
    saved = (G.pool, G.image_to_graph, G.add_positional_encoding, torch.get_default_dtype())
    dt = torch.float64

    def pool(*a, **k):
        return saved[0](*a, **k).to(dt)

    def image_to_graph(*a, **k):
        g = saved[1](*a, **k)
        for key in ("data", "edge_attrs", "n_pixels_per_node"):
            if torch.is_tensor(g.get(key)) and g[key].is_floating_point():
                g[key] = g[key].to(dt)
        return g

    def add_positional_encoding(x, *a, **k):
        # the positional planes are INPUTS of the model: keep the float32 values the reference feeds it, widened
        return saved[2](x.float(), *a, **k).to(dt)

    G.pool, G.image_to_graph, G.add_positional_encoding = pool, image_to_graph, add_positional_encoding
    torch.set_default_dtype(dt)
    try:
        yield
    finally:
        G.pool, G.image_to_graph, G.add_positional_encoding = saved[:3]
        torch.set_default_dtype(saved[3])
