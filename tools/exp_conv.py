"""Profiling experiment (not a bench): time the 128->128 conv kernel alone under the AZ_CONV_DBG variants
(conv_trunk.cu): 1 = alternate accumulators, 2 = epilogue without global traffic, 4 = no weight re-streaming,
8 = no activation re-load.  Tells which resource bounds the kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=slots, evaluator=E.EVAL_RESNET, net_blocks=1, num_simulations=8,
               max_nodes_per_tree=2048, deterministic=1)
eng.load_weights(N.export_weights(N.make_random_model(seed=0, blocks=1)))
flop = 225 * 9 * 128 * 128 * 2 * slots
for dbg in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 2, 8, 10, 0]):
    os.environ["AZ_CONV_DBG"] = str(dbg)
    ms = min(eng.conv_bench(slots, 20) for _ in range(3))
    print(f"dbg={dbg:2d}  {ms*1e3:8.1f} us  {flop/ms/1e9:8.1f} TFLOP/s", flush=True)
