"""Soak run (not a test): long self-play with game turnover on every game, throughput mode (noise, temperature, auto-restart) — the move-commit
path (region re-cut, re-root copy, sample ring, restarts) far past the few moves the bench plays.  Checks the counters' invariants.
    python tools/soak.py [resnet]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
resnet = len(sys.argv) > 1 and sys.argv[1] == "resnet"
cases = [(E.GOMOKU, 9, 512, 48, 120, 11, 81), (E.GO, 9, 512, 48, 200, 8, 82), (E.CHESS, 8, 256, 32, 560, 18, 20480), (E.GOMOKU, 15, 1024, 64, 60, 11, 225)]
for game, board, T, sims, moves, planes, actions in cases:
    kw = dict(game=game, board_size=board, n_slots=T, num_simulations=sims, deterministic=0, auto_restart=1, seed=5, sample_ring_capacity=T * 600)
    if resnet:
        eng = E.Engine(evaluator=E.EVAL_RESNET, net_blocks=2, **kw)
        eng.load_weights(N.export_weights(N.make_random_model(seed=1, blocks=2, in_planes=planes, board=board, actions=actions)))
    else:
        eng = E.Engine(evaluator=E.EVAL_HASH, **kw)
    t0 = time.time(); n_samples = 0; games_seen = set()
    for m in range(0, moves, 10):
        eng.play(10)
        smp = eng.drain_samples(cap=T * 600)
        n_samples += len(smp)
        if len(smp):
            assert np.all(np.abs(smp["z"]) <= 1) and np.all(smp["result"] >= 1) and np.all(smp["ply"] >= 0)
            games_seen.update(zip(smp["slot"].tolist(), smp["game_id"].tolist()))
    st = eng.stats()
    assert st["pool_overflows"] == 0 and st["samples_dropped"] == 0, st
    assert st["moves"] == (moves // 10) * 10 * T or st["games"] > 0, st
    assert len(games_seen) == st["games"] or st["games"] - len(games_seen) <= T, (len(games_seen), st["games"])
    print(f"game {game} board {board}: {st['moves']} moves, {st['games']} games finished, {n_samples} samples, {st['simulations']} simulations, "
          f"eval_shared {st['eval_shared']}, {time.time() - t0:.1f} s", flush=True)
    eng.close()
print("soak ok")
