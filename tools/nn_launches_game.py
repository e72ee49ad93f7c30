"""Profiling helper: a few network forwards for a given game config (ncu launch list): python tools/nn_launches_game.py chess|go9|go19"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
g = sys.argv[1] if len(sys.argv) > 1 else "chess"
game, board, planes, actions, slots, blocks, ch = {"chess": (E.CHESS, 8, 18, 20480, 1024, 10, 128), "go9": (E.GO, 9, 8, 82, 2048, 10, 128),
                                                   "go19": (E.GO, 19, 8, 362, 1024, 20, 256)}[g]
eng = E.Engine(game=game, board_size=board, n_slots=slots, evaluator=E.EVAL_RESNET, net_blocks=blocks, net_channels=ch, num_simulations=8,
               max_nodes_per_tree=4096, deterministic=1)
eng.load_weights(N.export_weights(N.make_random_model(seed=0, in_planes=planes, board=board, actions=actions, blocks=blocks, channels=ch)))
print("forward ms", eng.nn_bench(slots, 3))
