"""quadtree_mpnnlstm_b200 -- B200-native hot path of zach-gousseau/Quadtree-MPNNLSTM.

Quadtree graph build -> graph-conv LSTM cell -> seq2seq driver, behind the reference's own PyTorch
module surface, on hand-written sm_100a CUDA reached through a C ABI (``include/qmp_b200.h``,
``libqmp_b200.so``).  See DESIGN.md / INTEGRATION.md.  CUDA only: nothing here falls back to the CPU.
"""
from .graph_functions import (Graph, Mesh, create_static_heterogeneous_graph, create_static_homogeneous_graph,  # noqa: F401
                              flatten, image_to_graph, image_to_graph_pixelwise, plot_contours, unflatten)
from .model import CONVOLUTION_KWARGS, CONVOLUTIONS, GConvGRU, GConvLSTM, GraphConv, MPNNLSTM, MPNNLSTMI  # noqa: F401
from .seq2seq import Decoder, Encoder, Seq2Seq  # noqa: F401
from .mpnnlstm import DeviceWindowDataset, NextFramePredictor, NextFramePredictorS2S  # noqa: F401
from .utils import add_positional_encoding, get_n_params, int_to_datetime, normalize  # noqa: F401

__version__ = "0.1.0"
