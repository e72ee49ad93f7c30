"""alphazero-multi-game_b200 — B200-native batched AlphaZero self-play engine (hot path only).

The product is ``libaz_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/az_b200.h``);
this package is the thin Python host side: a ctypes binding (``engine.Engine``), the network definition used
to create / export weights (``net``), and the reference-shaped host classes (``host``).  The directory name
carries a hyphen, so import it through ``az_b200_loader.load()`` at the repo root (or add it to sys.path and
import the submodules directly).
"""
from .engine import Engine, EngineConfig, build_library, library_path, LibraryMissing  # noqa: F401
