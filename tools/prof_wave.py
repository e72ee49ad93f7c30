"""Profiling helper: `moves` self-play moves of a BASELINE config with the ResNet evaluator (random-init weights), for ncu captures of the
wave kernels at production tree sizes:  python tools/prof_wave.py chess|go9|gomoku15 [moves]   (AZ_NO_WAVE_GRAPH=1: plain launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
CFG = {"gomoku15": dict(game=E.GOMOKU, board=15, slots=4096, sims=800, planes=11, actions=225),
       "go9": dict(game=E.GO, board=9, slots=2048, sims=400, planes=8, actions=82),
       "chess": dict(game=E.CHESS, board=8, slots=1024, sims=800, planes=18, actions=20480)}
c = CFG[sys.argv[1] if len(sys.argv) > 1 else "chess"]
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = E.Engine(game=c["game"], board_size=c["board"], n_slots=c["slots"], num_simulations=c["sims"], evaluator=E.EVAL_RESNET, net_blocks=10, net_channels=128,
               deterministic=0, auto_restart=1)
eng.load_weights(N.export_weights(N.make_random_model(seed=0, in_planes=c["planes"], board=c["board"], actions=c["actions"], blocks=10, channels=128)))
for mv in range(moves):
    eng.event_record(0); eng.play(1); eng.event_record(1)
    print(f"move {mv}: {eng.event_elapsed(0, 1):8.1f} ms", flush=True)
print(eng.stats())
