"""Empty stub: the hot path never plots."""
