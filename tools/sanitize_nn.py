"""Small-size exerciser of the network kernels (not a test; was written as a compute-sanitizer target, which this pool does not allow): — stem, fused trunk (strided and board-aligned groups, incl. a
last group that overhangs the stream), fused heads kernel, FC GEMMs, softmax, and the examples kernels.
   compute-sanitizer --tool memcheck python tools/sanitize_nn.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
for game, board, planes, actions, n in ((E.GOMOKU, 15, 11, 225, 9), (E.GO, 9, 8, 82, 37), (E.CHESS, 8, 18, 20480, 5), (E.GO, 19, 8, 362, 3)):
    m = N.make_random_model(seed=1, blocks=2, in_planes=planes, board=board, actions=actions)
    eng = E.Engine(game=game, board_size=board, n_slots=n, evaluator=E.EVAL_RESNET, net_blocks=2, num_simulations=4, max_nodes_per_tree=4096, deterministic=0, auto_restart=1, seed=2)
    eng.load_weights(N.export_weights(m))
    x = (np.random.default_rng(0).random((n, planes, board, board)) < 0.2).astype(np.float32)
    pol, val, _ = eng.nn_forward(x, want_logits=True)
    eng.play(2)
    smp = eng.drain_samples()
    print("game", game, board, "ok", float(pol.sum()), float(val.mean()), len(smp), flush=True)
    eng.close()
