from oracle.convs_ref import Linear  # noqa: F401
