"""dune_grid_b200 — B200-native bulk implementation of dune-grid's YaspGrid leaf-sweep hot path.
The CUDA/C-ABI library (libdgb.so) is mandatory; importing this package without it fails loudly."""
from . import capi  # noqa: F401  (raises ImportError when libdgb.so is missing)
from .grid import *  # noqa: F401,F403
