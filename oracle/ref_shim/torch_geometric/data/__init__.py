import torch


class Data:
    """Attribute bag with .to(device), enough for graph_functions.Graph (graph_functions.py:28)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, **kwargs):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device, *args, **kwargs):
        if device is None:
            return self
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self
