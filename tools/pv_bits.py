"""Bit comparison of k_policy_value_reg (registers) with k_policy_value (AZ_PV_OLD=1): prints a digest of the policy / value / logits bytes of one
network forward per board geometry; run it once with and once without the variable and compare the lines.   python tools/pv_bits.py"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
for game, board, planes, actions, nb in ((E.GOMOKU, 15, 11, 225, 300), (E.GO, 9, 8, 82, 700), (E.GO, 13, 8, 170, 100), (E.GO, 19, 8, 362, 64)):
    eng = E.Engine(game=game, board_size=board, n_slots=max(nb, 64), evaluator=E.EVAL_RESNET, net_blocks=2, net_channels=128, num_simulations=4, deterministic=0, auto_restart=1)
    eng.load_weights(N.export_weights(N.make_random_model(seed=3, in_planes=planes, board=board, actions=actions, blocks=2, channels=128)))
    x = (np.random.default_rng(board).random((nb, planes, board, board)) < 0.3).astype(np.float32)
    pol, val, lg = eng.nn_forward(x, want_logits=True)
    print(game, board, hashlib.sha256(pol.tobytes() + val.tobytes() + lg.tobytes()).hexdigest()[:24], float(pol.sum()), flush=True)
    eng.close()
