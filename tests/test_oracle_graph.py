"""CPU: the graph oracle against the reference's own docstring examples and against CPython itself."""
import itertools

import numpy as np
import torch

from oracle import graph_ref as G


def test_set_order_model_matches_cpython_exhaustively_and_randomly():
    """The slot model the CUDA kernel implements == what `set` really does (graph_functions.py:308-343)."""
    pool = [-1, 0, 1, 5, 6, 7, 8, 13, 14, 15, 16, 22, 30, 38, 46, 54, 62, 70, 126, 134, 1022, 4093, 4094, 65534, 65535]
    for r in (1, 2, 3, 4):
        for combo in itertools.product(pool, repeat=r):
            s = set()
            for v in combo:
                s.add(np.int64(v))
            s.discard(-1)
            assert [int(v) for v in s] == G.cpython_small_set_order(combo), combo
    rng = np.random.default_rng(0)
    for _ in range(100000):
        combo = [int(v) for v in rng.integers(-1, 100000, size=int(rng.integers(1, 5)))]
        s = set()
        for v in combo:
            s.add(np.int64(v))
        s.discard(-1)
        assert [int(v) for v in s] == G.cpython_small_set_order(combo), combo


def test_get_adj_docstring_example():
    """graph_functions.py:266-282: labels [[0,0,1],[2,3,3],[2,3,3]] connect (0,1) (0,2) (0,3) (1,3) (2,3).
    The code emits both directions plus one self-loop per multi-pixel node (the docstring lists pairs once)."""
    labels = np.array([[0, 0, 1], [2, 3, 3], [2, 3, 3]])
    ei = G.adjacency(labels)
    pairs = {(int(a), int(b)) for a, b in ei.T}
    undirected = {(0, 1), (0, 2), (0, 3), (1, 3), (2, 3)}
    assert {tuple(sorted(p)) for p in pairs if p[0] != p[1]} == undirected
    assert all((b, a) in pairs for a, b in pairs)
    assert {p for p in pairs if p[0] == p[1]} == {(0, 0), (2, 2), (3, 3)}      # multi-pixel nodes only
    assert ei.shape[1] == 2 * len(undirected) + 3


def test_get_mapping_docstring_example():
    """graph_functions.py:560-574: the 4 x 9 one-hot matrix and counts [2, 1, 2, 4]."""
    m = G.LabelMap(np.array([[0, 0, 1], [2, 3, 3], [2, 3, 3]]))
    expect = torch.tensor([[1, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0, 0],
                           [0, 0, 0, 1, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 1, 0, 1, 1]], dtype=torch.float32)
    assert torch.equal(m.dense(), expect)
    assert m.counts().tolist() == [2, 1, 2, 4]


def test_grouped_mean_docstring_example():
    """graph_functions.py:427-431: arr [1..5], labels [0,1,1,2,2] -> [1, 2.5, 4.5] (the pooling rule)."""
    img = torch.tensor([1.0, 2, 3, 4, 5]).reshape(1, 1, 5, 1)
    m = G.LabelMap(np.array([[0, 1, 1, 2, 2]]))
    assert G.pool(img, m, m.counts()).reshape(-1).tolist() == [1.0, 2.5, 4.5]


def test_quadtree_basics():
    img = np.zeros((8, 8), np.float32)
    assert (G.quadtree_labels(img, thresh=0.5, max_size=8) == 0).all()              # nothing splits
    img[0, 0] = 1.0
    lab = G.quadtree_labels(img, thresh=0.5, max_size=8)
    assert lab.max() + 1 == 10 and lab[7, 7] == 0 and lab[0, 0] == 9                 # reverse-DFS numbering
    mask = np.zeros((8, 8), bool)
    mask[7, 7] = True
    lab = G.quadtree_labels(np.zeros((8, 8), np.float32), thresh=0.5, max_size=8, mask=mask)
    assert lab[7, 7] == -1 and (lab >= 0).sum() == 63
    # the (size+1) window reaches one pixel into the neighbouring cell
    img = np.zeros((8, 8), np.float32)
    img[4, 0] = 1.0
    lab = G.quadtree_labels(img, thresh=0.5, max_size=4)
    assert len(np.unique(lab[:4, :4])) > 1, "top-left base cell must split because its window sees row 4"


def test_pixelwise_and_unpool_fill():
    mask = np.array([[False, True], [False, False]])
    assert G.pixelwise_labels(mask).tolist() == [[0, -1], [1, 2]]
    ei = G.adjacency_pixelwise(G.pixelwise_labels(mask))
    assert ei.T.tolist() == [[0, 1], [1, 0], [1, 2], [2, 1]]
    img = G.unpool(torch.tensor([[1.0], [2.0], [3.0]]), None, (2, 2), mask)
    assert torch.isnan(img[0, 1, 0]) and img[1, 1, 0] == 3.0
