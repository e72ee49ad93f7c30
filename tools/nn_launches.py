"""Profiling helper: a few network forwards at the BASELINE batch (4096 boards) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=4096, evaluator=E.EVAL_RESNET, net_blocks=10, num_simulations=8,
               max_nodes_per_tree=2048, deterministic=1)
eng.load_weights(N.export_weights(N.make_random_model(seed=0)))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
print("forward ms", min(eng.nn_bench(4096, reps) for _ in range(3 if reps > 3 else 1)), "PDL off" if os.environ.get("AZ_CONV_NO_PDL") else "PDL on")
