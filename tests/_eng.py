"""Shared helpers for the GPU tests: engine construction through the C ABI (ctypes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader  # noqa: E402

az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E  # noqa: E402
from alphazero_multi_game_b200 import net as N  # noqa: E402


def hash_engine(n_slots, board=15, sims=800, **kw):
    cfg = dict(game=E.GOMOKU, board_size=board, n_slots=n_slots, num_simulations=sims, evaluator=E.EVAL_HASH,
               deterministic=1, auto_restart=0, max_nodes_per_tree=2 * (sims + 1) * board * board + 1, n_streams=2)
    cfg.update(kw)
    return E.Engine(**cfg)
