from oracle.convs_ref import glorot, zeros  # noqa: F401
