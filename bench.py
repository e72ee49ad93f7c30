#!/usr/bin/env python
"""bench.py — MCTS simulations/s of batched self-play, Gomoku 15x15 @ 800 sims, 10-block/128-ch ResNet (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own serial CPU search, same metric

A step = one self-play move for every game slot on this rank: root expansion, `sims` waves (one simulation per
tree per wave: select → encode → ResNet forward → expand/backup), move choice, sample record, re-root, game
turnover.  value = simulations completed by all ranks / device time of the K timed steps (max over ranks).
Prints ONE JSON line on rank 0.  Timing: CUDA events on the engine's own stream, barrier + synchronize on both
sides; the per-step working set (node pools ~ tens of GB, activations 3 x 268 MB) is far larger than L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"
BOARD, ACTIONS, PLANES, BLOCKS, CHANNELS = 15, 225, 11, 10, 128
# algorithmic FLOPs of one 128->128 3x3 conv over one board (real cells only)
CONV_FLOP_PER_BOARD = 225 * 9 * 128 * 128 * 2
NET_FLOP_PER_EVAL = 1.336e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16_tflops=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------- reference CPU arm
def reference_searcher(threads):
    """The reference's serial ParallelMCTS (patched build oracle/_ref; else the oracle port) on Gomoku 15x15 with the
    fp32 PyTorch network evaluated on the host CPU through the evaluator callback.  Returns (run(n_sims) -> seconds, kind)."""
    import numpy as np
    import torch
    import _orc
    import az_b200_loader
    az_b200_loader.load()
    from alphazero_multi_game_b200 import net as N
    torch.set_num_threads(threads)
    model = N.make_random_model(seed=0, in_planes=PLANES, board=BOARD, actions=ACTIONS, blocks=BLOCKS, channels=CHANNELS)
    K = _orc.reference() if _orc.have_ref() else None
    kind = "reference" if K is not None else "port"
    if K is None:
        K = _orc.oracle()

    def cb(planes, c, h, w, a, pol_out, val_out, user):
        x = torch.from_numpy(np.ctypeslib.as_array(planes, shape=(1, c, h, w)).copy())
        with torch.no_grad():
            p, v = model(x)
            p = torch.softmax(p[0, :a], 0)      # TorchNeuralNetwork::predictBatch softmax (torch_neural_network.cpp:298-316)
        np.ctypeslib.as_array(pol_out, shape=(a,))[:] = p.numpy()
        val_out[0] = float(v)

    cbf = _orc.EVAL_CB(cb)
    state = K.new_state(_orc.GOMOKU, BOARD)
    holder = dict(m=K.mcts_new(state, 100, 1.5, 3, 1, cbf, None), done=0, cb=cbf, state=state)

    def run(n_sims, chunk=100):
        t0 = time.perf_counter()
        left = n_sims
        while left > 0:
            c = min(chunk, left)
            K.mcts_set_sims(holder["m"], c)
            K.mcts_search(holder["m"])
            left -= c; holder["done"] += c
            if holder["done"] >= 800:             # 800 sims per move, then play it (playSingleGame loop)
                a = K.mcts_select_action(holder["m"], 1, 1.0)
                K.mcts_update_with_move(holder["m"], a); K.state_make_move(state, a); holder["done"] = 0
                if K.state_is_terminal(state):
                    holder["state"] = K.new_state(_orc.GOMOKU, BOARD)
                    holder["m"] = K.mcts_new(holder["state"], 100, 1.5, 3, 1, cbf, None)
        return time.perf_counter() - t0

    return run, kind


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    run, kind = reference_searcher(threads)
    sims_per_step = args.ref_sims_per_step
    for _ in range(args.warmup):
        run(sims_per_step)
    t = 0.0
    for _ in range(args.steps):
        t += run(sims_per_step)
    v = sims_per_step * args.steps / t
    sample = f"{args.steps} steps x {sims_per_step} simulations of one Gomoku 15x15 game, serial search, fp32 net on {threads} CPU threads"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, world):
    name = {"go9": "Go 9x9 (capture/ko/superko)", "go19": "Go 19x19", "chess": "Chess (20480-action head)"}.get(args.game, "Gomoku 15x15")
    return {"workload": f"{name} batched self-play, {args.slots} concurrent games per GPU, {args.sims} sims/move, "
                        f"{BLOCKS}-block {CHANNELS}-ch random-init ResNet (BASELINE.json configs[{dict(go9=2, go19=3, chess=4).get(args.game, 1)}])",
            "slots_per_gpu": args.slots, "sims_per_move": args.sims, "parallelism": f"games sharded over {world} GPU(s)",
            "step": "one self-play move on every slot (root expansion + sims waves + move commit)", "stream_groups": args.streams,
            "l2": "working set (node pools, 3 x 268 MB activations) >> 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------- engine arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--game", default="gomoku15", choices=["gomoku15", "go9", "go19", "chess"],
                    help="gomoku15 = BASELINE.json configs[1] (the headline metric); extra lines: go9 = configs[2] (Go 9x9, 2048 games, "
                         "400 sims), go19 = configs[3] per GPU (Go 19x19, 20-block 256-ch, 1024 games), chess = configs[4] (1024 games, 800 sims, 20480-action policy head)")
    ap.add_argument("--slots", type=int, default=None)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--streams", type=int, default=1, help="stream groups the slots are split into (tree kernels of one overlap the network pass of the other)")
    ap.add_argument("--ref-sims-per-step", type=int, default=200)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global BOARD, ACTIONS, PLANES, BLOCKS, CHANNELS, CONV_FLOP_PER_BOARD, NET_FLOP_PER_EVAL
    if args.game == "go9":
        BOARD, ACTIONS, PLANES = 9, 82, 8
        CONV_FLOP_PER_BOARD = 81 * 9 * 128 * 128 * 2
        NET_FLOP_PER_EVAL = 81 * 9 * 8 * 128 * 2 + 20 * CONV_FLOP_PER_BOARD + 2 * (64 * 128 * 32 * 2) + 2048 * 82 * 2 + 2048 * 256 * 2 + 512
    if args.game == "go19":
        BOARD, ACTIONS, PLANES, BLOCKS, CHANNELS = 19, 362, 8, 20, 256
        CONV_FLOP_PER_BOARD = 361 * 9 * 128 * 128 * 2                      # one 128 -> 128 slice launch; a 256 -> 256 layer is 4 of them
        NET_FLOP_PER_EVAL = 361 * 9 * 8 * 256 * 2 + 40 * 4 * CONV_FLOP_PER_BOARD + 2 * (64 * 256 * 32 * 2) + 2048 * 362 * 2 + 2048 * 256 * 2 + 512
    if args.game == "chess":
        BOARD, ACTIONS, PLANES = 8, 20480, 18
        CONV_FLOP_PER_BOARD = 64 * 9 * 128 * 128 * 2
        NET_FLOP_PER_EVAL = 64 * 9 * 18 * 128 * 2 + 20 * CONV_FLOP_PER_BOARD + 2 * (64 * 128 * 32 * 2) + 2048 * 20480 * 2 + 2048 * 256 * 2 + 512
    args.slots = args.slots or {"go9": 2048, "chess": 1024, "go19": 1024}.get(args.game, 4096)
    args.sims = args.sims or {"go9": 400, "go19": 400}.get(args.game, 800)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import az_b200_loader
    az_b200_loader.load()
    from alphazero_multi_game_b200 import engine as E, net as N
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    eng = E.Engine(game={"go9": E.GO, "go19": E.GO, "chess": E.CHESS}.get(args.game, E.GOMOKU), board_size=BOARD, n_slots=args.slots, num_simulations=args.sims, evaluator=E.EVAL_RESNET,
                   net_blocks=BLOCKS, net_channels=CHANNELS, deterministic=0, auto_restart=1, device=local, seed=1234 + rank,
                   n_streams=args.streams)
    model = N.make_random_model(seed=0, in_planes=PLANES, board=BOARD, actions=ACTIONS, blocks=BLOCKS, channels=CHANNELS)
    blob = N.export_weights(model)
    eng.load_weights(blob)
    pinned = torch.empty(max(32 * args.slots, 4096) * eng.sample_layout().record_bytes, dtype=torch.uint8).pin_memory()   # = the engine's default sample ring
    samples_np = pinned.numpy().view(eng.sample_dtype())

    def barrier():
        eng.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def allreduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(args.warmup):
        eng.play(1)
    eng.drain_samples(out=samples_np)

    # ---- timed region 1: device-resident throughput --------------------------------------------------------
    clocks = ClockSampler(local); clocks.start()
    barrier()
    s0 = eng.stats()
    live0 = eng.conv_sampled()
    eng.event_record(0)
    for _ in range(args.steps):
        eng.play(1)
    eng.event_record(1)
    ms = eng.event_elapsed(0, 1)
    barrier()
    s1 = eng.stats()
    live1 = eng.conv_sampled()
    clk = clocks.stop()
    ms_max = allreduce(ms, dist.ReduceOp.MAX) if dist else ms
    sims = allreduce(float(s1["simulations"] - s0["simulations"]), dist.ReduceOp.SUM) if dist else float(s1["simulations"] - s0["simulations"])
    moves = allreduce(float(s1["moves"] - s0["moves"]), dist.ReduceOp.SUM) if dist else float(s1["moves"] - s0["moves"])
    evals = float(s1["evaluations"] - s0["evaluations"])
    launches = int(s1["kernel_launches"] - s0["kernel_launches"])
    value = sims / (ms_max / 1e3)

    # ---- timed region 2: end to end through the C ABI with host buffers ------------------------------------
    # every step: (H2D) a fresh fp32 weight blob from host memory is folded / converted and uploaded — the trainer →
    # self-play hand-off of a real AlphaZero loop, here once per move, i.e. far more often than in production;
    # one self-play move; (D2H) finished-game samples into pinned host memory + the chosen actions + the counters;
    # for N > 1 the finished-game samples are also all-gathered over NCCL (the path's only exchange step).
    h2d = len(blob)
    blob_pinned = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()      # the step's input lives in pinned host memory
    if dist is not None:
        from alphazero_multi_game_b200 import gather as GA
        cap = 32 * args.slots
        rec_bytes = eng.sample_layout().record_bytes
        dev_samples = torch.zeros(cap * rec_bytes, dtype=torch.uint8, device="cuda")
        GA.all_gather_samples(dist, dev_samples, 1, rec_bytes)      # untimed warm-up of the collective (NCCL sets up its channels on first use)
    barrier()
    e0 = eng.stats()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        eng.load_weights(blob_pinned.data_ptr(), h2d)
        eng.play(1)
        if dist is None:
            smp = eng.drain_samples(out=samples_np)
            d2h += smp.nbytes
        else:
            n = eng.drain_samples_device(dev_samples.data_ptr(), cap)
            all_samples, per_rank = GA.all_gather_samples(dist, dev_samples, n, rec_bytes)     # NCCL; the path's only exchange step
            d2h += 8 * world
        acts = eng.last_actions(); d2h += acts.nbytes
        eng.stats(); d2h += 88
    barrier()
    t_e2e = time.perf_counter() - t0
    e1 = eng.stats()
    t_e2e = allreduce(t_e2e, dist.ReduceOp.MAX) if dist else t_e2e
    sims_e2e = allreduce(float(e1["simulations"] - e0["simulations"]), dist.ReduceOp.SUM) if dist else float(e1["simulations"] - e0["simulations"])
    e2e = {"value": sims_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // max(args.steps, 1),
           "what": "load_weights(pinned host blob) + play(1) + drain_samples(pinned host) + last_actions + stats per step"
                   + ("; + NCCL all-gather of finished-game samples" if dist else "")}

    # ---- roofline of the dominant kernel (3x3 conv 128->128 on tcgen05), timed alone with CUDA events ------
    pk = peaks()
    boards_per_launch = (args.slots + args.streams - 1) // args.streams      # the production launch shape: one stream group
    conv_ms_alone = eng.conv_bench(boards_per_launch, 20)
    conv_flop = CONV_FLOP_PER_BOARD * boards_per_launch
    # live: CUDA events around the conv launches of every 64th wave INSIDE the timed region above (engine stream); the launch
    # shape in a wave is the number of non-terminal leaves, ~ all slots
    live_n = live1[1] - live0[1]
    conv_ms = (live1[0] - live0[0]) / live_n if live_n else conv_ms_alone
    achieved = conv_flop / (conv_ms / 1e3) / 1e12
    nn_ms = eng.nn_bench(args.slots, 5)
    roofline = {"bound": "tensor", "kernel": (f"k_conv3x3_pair_wide (one 128->128 slice launch over {boards_per_launch} boards; weight-stationary CTA pair with K-split activation stages for the 21-row halo)" if args.game == "go19" else
                           f"k_trunk_pair (the trunk's 20 128->128 3x3 conv layers over the {boards_per_launch} boards of one stream group as ONE persistent launch of the weight-stationary CTA-pair kernel, cta_group::2; figures are per layer = launch / 20)" if (args.game == "gomoku15" and not os.environ.get("AZ_TRUNK_LAYERED")) else
                           f"k_conv3x3_pair (one 128->128 3x3 conv layer over the {boards_per_launch} boards of one stream group; weight-stationary CTA pair, cta_group::2)"), "achieved": achieved,
                # kernel timed inside a long step -> the SUSTAINED measured cuBLAS bf16 rate is the denominator; timed alone -> the burst one
                "peak": pk["bf16_sustained"] if live_n else pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / (pk["bf16_sustained"] if live_n else pk["bf16_tflops"]),
                "peak_source": pk["source"] + (" sustained bf16 (kernel timed inside the step)" if live_n else " burst bf16 (kernel timed alone)"),
                "frac_note": "the sustained peak is a power-capped cuBLAS bf16 GEMM (MEASURED_PEAKS.json), not a hardware limit: frac can exceed 1; nominal dense bf16 is 2250 TFLOP/s and the tensor pipe also multiplies the 12 % padding rows of the position stream, which `achieved` does not count",
                "timed_alone": {"achieved": conv_flop / (conv_ms_alone / 1e3) / 1e12, "peak": pk["bf16_tflops"],
                                "frac": conv_flop / (conv_ms_alone / 1e3) / 1e12 / pk["bf16_tflops"], "peak_source": pk["source"] + " burst bf16"},
                # DRAM bytes per launch of this kernel from the ncu --set full capture in profiles/r1_summary.md (4096 Gomoku boards,
                # layer without residual): dram__bytes_read.sum 269.9 MB + dram__bytes_write.sum 224.3 MB; algorithmic 268 + 268 MB
                # fused trunk (k_trunk_pair, profiles/r1f_trunk_pair_ncu_raw.csv): dram__bytes_read.sum 814 MB + dram__bytes_write.sum 3062 MB per launch
                # of 20 layers = 193.8 MB per layer — below the algorithmic 536-805 MB because a group's activations stay in L2
                "traffic": (193.8e6 if not os.environ.get("AZ_TRUNK_LAYERED") else 494.2e6) if (args.game == "gomoku15" and boards_per_launch == 4096) else None,
                "launch_ms": conv_ms, "launch_ms_source": f"live: {live_n} layers bracketed by CUDA events inside the timed steps" if live_n else "timed alone (no sampled wave in the timed region)",
                "launch_ms_timed_alone": conv_ms_alone, "flop_per_launch": conv_flop,
                "whole_net_ms": nn_ms, "whole_net_tflops": NET_FLOP_PER_EVAL * args.slots / (nn_ms / 1e3) / 1e12,
                "step_share_note": f"{20 * args.streams} of these launches per wave (20 per stream group); see profiles/ for the ncu launch list"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own serial search on the host cores --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.game == "gomoku15":
        try:
            threads = os.cpu_count() or 1
            run, kind = reference_searcher(threads)
            run(20)
            n, t = 0, 0.0
            while t < args.cpu_baseline_seconds:
                t += run(50); n += 50
            cpu = {"value": n / t, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{n} simulations of one Gomoku 15x15 game (800 sims/move), serial ParallelMCTS, fp32 10x128 net on {threads} host threads, {t:.1f} s"}
        except Exception as ex:      # the checker is optional equipment on the GPU box
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                          "moves_per_sec": moves / (ms_max / 1e3), "nn_evals_per_sec_rank0": evals / (ms / 1e3),
                          "tensor_roofline_frac_in_step": (evals / (ms / 1e3)) * NET_FLOP_PER_EVAL / 1e12 / pk["bf16_sustained"],
                          "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
                          "games_finished": int(e1["games"]), "samples_dropped": int(e1["samples_dropped"]), "pool_overflows": int(e1["pool_overflows"])}))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
