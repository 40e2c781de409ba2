from quadtree_mpnnlstm_b200.model import (CONVOLUTION_KWARGS, CONVOLUTIONS, GConvLSTM, GraphConv, MPNNLSTM,  # noqa: F401
                                          MPNNLSTMI)
