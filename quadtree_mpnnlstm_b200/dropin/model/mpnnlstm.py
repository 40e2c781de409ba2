from quadtree_mpnnlstm_b200.mpnnlstm import DeviceWindowDataset, NextFramePredictor, NextFramePredictorS2S  # noqa: F401
