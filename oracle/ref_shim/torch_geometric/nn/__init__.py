import torch.nn as nn
from oracle.convs_ref import GCNConv, ChebConv, TransformerConv, GATConv, GATv2Conv  # noqa: F401  (TransformerConv: any number of heads, so the
# reference's own MHTransformerConv subclass, model/model.py:26-37, runs on it unmodified)
from .conv import MessagePassing  # noqa: F401


class _Unsupported(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__}: no config of the hot path selects it (SURVEY.md section 2)")


class GraphConv(_Unsupported):
    pass
