"""Profiling helper: one 128->128 slice launch of the Go 19x19 / 256-channel trunk (k_conv3x3_pair_wide) at several board counts —
per-item time in steady state vs per-launch overhead.   python tools/go19_conv.py [channels]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import az_b200_loader; az_b200_loader.load()
from alphazero_multi_game_b200 import engine as E, net as N
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = E.Engine(game=E.GO, board_size=19, n_slots=4096, evaluator=E.EVAL_RESNET, net_blocks=2, net_channels=ch, num_simulations=4,
               max_nodes_per_tree=2048, deterministic=1)
eng.load_weights(N.export_weights(N.make_random_model(seed=0, in_planes=8, board=19, actions=362, blocks=2, channels=ch)))
for nb in (128, 256, 512, 1024, 2048, 4096):
    ms = min(eng.conv_bench(nb, 20) for _ in range(3))
    items = (nb * 400 + 255) // 256
    per_pair = -(-items // 74)
    print(f"boards {nb:5d}: {ms * 1e3:8.1f} us per launch, {items} items, {per_pair} per pair -> {ms * 1e3 / per_pair:6.2f} us per item-slot, "
          f"{nb * 361 * 9 * 128 * 128 * 2 / ms / 1e9:7.1f} TFLOP/s algorithmic", flush=True)
