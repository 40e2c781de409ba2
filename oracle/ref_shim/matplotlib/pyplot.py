def __getattr__(name):
    raise NotImplementedError("matplotlib stub: plotting is outside the hot path")
