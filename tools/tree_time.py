"""Profiling helper: cost of the tree kernels alone (hash evaluator instead of the network) per wave at BASELINE tree sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E
slots, sims = 4096, 800
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=slots, evaluator=E.EVAL_HASH, num_simulations=sims, deterministic=0,
               auto_restart=1, n_streams=streams)
for mv in range(8):
    eng.event_record(0); eng.play(1); eng.event_record(1)
    ms = eng.event_elapsed(0, 1)
    print(f"move {mv}: play(1) {ms:8.1f} ms = {ms / (sims + 1) * 1e3:7.1f} us / wave (select + hash eval + expand/backup, + move commit)", flush=True)
print(eng.stats())
