"""compute-sanitizer target (not a test): small searches on every game with the hash evaluator + the state API, so memcheck /
racecheck see every tree / rules kernel.   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E
for game, board, sims in ((E.GOMOKU, 9, 24), (E.GOMOKU, 15, 12), (E.GO, 9, 24), (E.GO, 19, 6), (E.CHESS, 8, 12)):
    mc = {E.GOMOKU: board * board, E.GO: board * board + 1, E.CHESS: 256}[game]
    eng = E.Engine(game=game, board_size=board, n_slots=7, num_simulations=sims, evaluator=E.EVAL_HASH, deterministic=0, auto_restart=1,
                   max_nodes_per_tree=2 * (sims + 2) * mc + 1, n_streams=2, seed=3)
    eng.play(6)
    eng.search(); st = eng.root_stats(3)
    eng.advance([int(st["actions"][-1])] * 7)
    eng.add_dirichlet_noise(0.3, 0.25)
    eng.drain_samples()
    r = eng.rules_replay([[int(st["actions"][0])] if game != E.CHESS else [], []])
    print("game", game, board, "ok", eng.stats()["simulations"], r["n_legal"].tolist(), flush=True)
    eng.close()
# the evaluation cache behind the hash evaluator, small enough to evict all the time (chess: the table model takes its place)
for game, board, sims in ((E.GOMOKU, 9, 24), (E.GO, 9, 24)):
    mc = {E.GOMOKU: board * board, E.GO: board * board + 1}[game]
    eng = E.Engine(game=game, board_size=board, n_slots=7, num_simulations=sims, evaluator=E.EVAL_HASH, deterministic=0, auto_restart=1,
                   max_nodes_per_tree=2 * (sims + 2) * mc + 1, n_streams=2, seed=3, eval_cache_entries=128)
    eng.play(8)
    st = eng.stats()
    print("cache", game, board, "ok", st["simulations"], st["eval_cached"], st["eval_shared"], flush=True)
    eng.close()
