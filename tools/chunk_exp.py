"""Profiling experiment (not a bench): the network forward over 4096 Gomoku boards as ONE pass vs as sub-batches small enough for the
activations to stay in L2 between layers (engine capacity = chunk size; nn_bench walks the boards chunk by chunk)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
blob = N.export_weights(N.make_random_model(seed=0))
total = 4096
for chunk in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,1036,518,512").split(",")]:
    eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=chunk, evaluator=E.EVAL_RESNET, net_blocks=10, num_simulations=8,
                   max_nodes_per_tree=2048, deterministic=1)
    eng.load_weights(blob)
    ms = min(eng.nn_bench(total, 10) for _ in range(3))
    long_ms = eng.nn_bench(total, 200)          # ~1 s of back-to-back forwards: the power-capped steady state
    print(f"chunk {chunk:5d}: forward over {total} boards {ms:.3f} ms (best of 3 x 10), {long_ms:.3f} ms sustained (200 reps)", flush=True)
    eng.close()
