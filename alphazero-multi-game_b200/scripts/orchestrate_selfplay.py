#!/usr/bin/env python3
"""orchestrate_selfplay.py — the reference's multi-process self-play orchestrator (python/scripts/orchestrate_selfplay.py:92-160 arguments).

The reference starts --processes OS processes (self_play.py or the self_play binary), pins them to core ranges and polls psutil / nvidia-smi
while they fill per-process output directories.  On the B200 engine concurrency lives on the device: ONE process drives all games of a GPU
as slots of one engine, and several GPUs are driven by one SelfPlayManager (engine + host thread per device, NCCL sample gather).  This
command therefore keeps the reference's options and output layout (proc_<i>/ directories, a final summary JSON with the same headline
keys) and maps --processes to GPUs: process i = GPU i (at most the visible device count; extra processes fold onto the GPUs as more game
slots).  --use-cpp-binary runs the C++ `self_play` command instead of the Python driver, as in the reference."""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="AlphaZero Self-Play Orchestrator (B200 engine)")
    p.add_argument("--model", type=str, default="")
    p.add_argument("--game", type=str, default="gomoku", choices=["gomoku", "chess", "go"])
    p.add_argument("--size", type=int, default=0)
    p.add_argument("--num-games", type=int, default=100)
    p.add_argument("--simulations", type=int, default=800)
    p.add_argument("--threads", type=int, default=0)
    p.add_argument("--processes", type=int, default=1, help="reference: OS processes; here: GPUs (capped at the visible device count)")
    p.add_argument("--output-dir", type=str, default="data/games")
    p.add_argument("--batch-size", type=int, default=16)
    p.add_argument("--batch-timeout", type=int, default=10)
    p.add_argument("--temperature", type=float, default=1.0)
    p.add_argument("--temp-drop", type=int, default=30)
    p.add_argument("--final-temp", type=float, default=0.0)
    p.add_argument("--dirichlet-alpha", type=float, default=0.03)
    p.add_argument("--dirichlet-epsilon", type=float, default=0.25)
    p.add_argument("--variant", action="store_true")
    p.add_argument("--no-gpu", action="store_true")
    p.add_argument("--fp16", action="store_true")
    p.add_argument("--monitor-interval", type=int, default=5)
    p.add_argument("--create-random-model", action="store_true")
    p.add_argument("--use-cpp-binary", action="store_true")
    p.add_argument("--profile", action="store_true")
    p.add_argument("--c-puct", type=float, default=1.5)
    p.add_argument("--fpu-reduction", type=float, default=0.1)
    p.add_argument("--virtual-loss", type=int, default=3)
    p.add_argument("--no-tt", action="store_true")
    p.add_argument("--progressive-widening", action="store_true")
    for flag in ("--optimize-batch", "--optimize-threads", "--pin-threads"):
        p.add_argument(flag, action="store_true", help="(recorded only)")
    p.add_argument("--cache-size", type=int, default=2097152, help="(recorded only)")
    p.add_argument("--compact-size", type=int, default=0, help="(recorded only)")
    return p.parse_args(argv)


def visible_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return max(1, sum(1 for ln in out.splitlines() if ln.startswith("GPU ")))
    except Exception:
        return 1


def main(argv=None):
    a = parse_args(argv)
    gpus = min(max(1, a.processes), visible_gpus())
    out_dir = os.path.join(a.output_dir, "proc_0")            # one driver process: the reference's per-process layout with a single entry
    os.makedirs(out_dir, exist_ok=True)
    common = ["--game", a.game, "--num-games", str(a.num_games), "--simulations", str(a.simulations), "--output-dir", out_dir,
              "--temperature", str(a.temperature), "--temp-drop", str(a.temp_drop), "--final-temp", str(a.final_temp),
              "--dirichlet-alpha", str(a.dirichlet_alpha), "--dirichlet-epsilon", str(a.dirichlet_epsilon), "--c-puct", str(a.c_puct),
              "--virtual-loss", str(a.virtual_loss), "--gpus", str(gpus)]
    if a.size:
        common += ["--size", str(a.size)]
    if a.model:
        common += ["--model", a.model]
    if a.use_cpp_binary:
        cmd = [os.path.join(PKG, "self_play")] + common
    else:
        cmd = [sys.executable, os.path.join(HERE, "self_play.py")] + common + (["--create-random-model"] if a.create_random_model else [])
    print(f"Starting 1 driver process over {gpus} GPU(s): {' '.join(cmd)}")
    t0 = time.time()
    proc = subprocess.Popen(cmd)
    while proc.poll() is None:
        time.sleep(max(1, a.monitor_interval))
        n = len([f for f in os.listdir(out_dir) if f.endswith(".json") and not f.startswith("metadata")])
        print(f"[monitor] {time.time() - t0:7.1f} s  games written: {n}/{a.num_games}")
    dur = time.time() - t0
    games = [f for f in os.listdir(out_dir) if f.endswith(".json") and not f.startswith("metadata")]
    moves = 0
    for f in games:
        try:
            moves += len(json.load(open(os.path.join(out_dir, f)))["moves"])
        except Exception:
            pass
    summary = {"total_games": len(games), "total_moves": moves, "total_time_seconds": dur, "games_per_second": len(games) / dur if dur else 0.0,
               "moves_per_second": moves / dur if dur else 0.0, "processes": a.processes, "gpus": gpus, "return_code": proc.returncode}
    json.dump(summary, open(os.path.join(a.output_dir, f"orchestration_summary_{time.strftime('%Y%m%d_%H%M%S')}.json"), "w"), indent=2)
    print(json.dumps(summary))
    return proc.returncode


if __name__ == "__main__":
    sys.exit(main())
