import torch.nn as nn


class MessagePassing(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
