from quadtree_mpnnlstm_b200.seq2seq import Decoder, Encoder, Seq2Seq  # noqa: F401
