from quadtree_mpnnlstm_b200.utils import (add_positional_encoding, get_n_params, int_to_datetime, normalize,  # noqa: F401
                                          round_to_day)
