from quadtree_mpnnlstm_b200.model import (CONVOLUTION_KWARGS, CONVOLUTIONS, GConvGRU, GConvLSTM, GraphConv, MPNNLSTM,  # noqa: F401
                                          MPNNLSTMI)
