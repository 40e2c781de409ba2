"""CPU oracle for the Quadtree-MPNNLSTM hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / CPU PyTorch, the algorithm of the
reference's hot path (quadtree graph build -> graph-conv LSTM cell -> seq2seq
driver).  It exists so that the CUDA product in ``quadtree_mpnnlstm_b200`` can
be checked against it; it is never shipped, never measured as the product, and
the product never imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it.

Pinning status
--------------
* graph half (``graph_ref``): PINNED.  Checked bit-exactly (labels, edge_index)
  and to float tolerance (node data, edge attrs) against the reference's own
  ``model/graph_functions.py`` imported unmodified from ``/root/reference`` in the
  build container (``tests/test_oracle_pinned.py``), and against the three
  docstring examples the reference holds (graph_functions.py:266-282, 560-574,
  427-431).  Golden vectors produced by the reference itself are committed under
  ``tests/golden/`` with the generating script.
* driver + cell (``seq2seq_ref``, ``cell_ref``): PINNED against the reference's
  own ``model/seq2seq.py`` / ``model/model.py`` imported unmodified on top of the
  conv restatements below (same state dict, same inputs).
* conv arithmetic (``convs_ref``): PARITY UNPINNED.  The arithmetic lives in the
  third-party dependency ``torch-geometric==2.2.0`` (reference requirements.txt:13;
  also torch-scatter==2.1.0, torch-sparse==0.6.15), which is absent from
  ``/root/reference`` and cannot be installed here (no network).  ``convs_ref``
  restates PyG 2.2.0's published algorithm for GCNConv / ChebConv /
  TransformerConv and is validated structurally against dense-matrix formulas
  (``tests/test_oracle_convs.py``); no reference test pins a number at this
  boundary.
"""
