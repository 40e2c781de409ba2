"""Launch the decoder-cell kernel a few times at the bench mesh (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], "none"]
import cell_kernel_time  # noqa
