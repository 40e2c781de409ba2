"""Stand-in for torch-geometric 2.2.0 (see ../README.md).  Test infrastructure only."""
__version__ = "2.2.0+oracle-shim"
