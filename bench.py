#!/usr/bin/env python
"""bench.py — MCTS simulations/s of batched self-play, Gomoku 15x15 @ 800 sims, 10-block/128-ch ResNet (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own serial CPU search, same metric

A step = one self-play move for every game slot on this rank: root expansion, `sims` waves (one simulation per
tree per wave: select → encode → ResNet forward → expand/backup), move choice, sample record, re-root + region re-cut,
game turnover.  value = simulations completed by all ranks / device time of the K timed steps (max over ranks).
Prints ONE JSON line on rank 0: the headline is BASELINE.json configs[1] (Gomoku 15x15, 4096 games, 800 sims); the same line
carries `other_configs` = configs[2] (Go 9x9), configs[4] (chess) and configs[3] per GPU (Go 19x19, 20 x 256), each with its own
value / e2e / roofline fraction / clocks.  Timing: CUDA events on the engine's own stream, barrier + synchronize on both
sides; the per-step working set (node pools ~ tens of GB, activations 3 x 268 MB) is far larger than L2.
Exit status 3 when any expansion failed for lack of node-pool room (pool_overflows > 0): such a run is not a measurement.
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

# BASELINE.json configs: game, board, actions, planes, blocks, channels, slots per GPU, simulations per move
CONFIGS = {
    "gomoku15": dict(idx=1, name="Gomoku 15x15", game="GOMOKU", board=15, actions=225, planes=11, blocks=10, channels=128, slots=4096, sims=800),
    "go9": dict(idx=2, name="Go 9x9 (capture/ko/superko)", game="GO", board=9, actions=82, planes=8, blocks=10, channels=128, slots=2048, sims=400),
    "go19": dict(idx=3, name="Go 19x19", game="GO", board=19, actions=362, planes=8, blocks=20, channels=256, slots=1024, sims=400),
    "chess": dict(idx=4, name="Chess (20480-action head)", game="CHESS", board=8, actions=20480, planes=18, blocks=10, channels=128, slots=1024, sims=800),
}


def flops(c):
    """(algorithmic FLOPs of one 128->128 3x3 conv launch over one board — real cells only; FLOPs of one whole network evaluation)"""
    cells = c["board"] * c["board"]
    conv = cells * 9 * 128 * 128 * 2
    ns = c["channels"] // 128
    # policy FC: all A logits, except on the wide chess head, where the waves compute the legal moves' logits only (~35 per position)
    policy_fc = 2048 * (c["actions"] if c["actions"] <= 1024 else 35) * 2
    net = (cells * 9 * c["planes"] * c["channels"] * 2 + 2 * c["blocks"] * ns * ns * conv + 2 * (64 * c["channels"] * 32 * 2)
           + policy_fc + 2048 * 256 * 2 + 512)
    return conv, net


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
def reference_searcher(threads, c):
    """The reference's serial ParallelMCTS (patched build oracle/_ref; else the oracle port) on Gomoku 15x15 with the
    fp32 PyTorch network evaluated on the host CPU through the evaluator callback.  Returns (run(n_sims) -> seconds, kind)."""
    import numpy as np
    import torch
    import _orc
    import az_b200_loader
    az_b200_loader.load()
    from alphazero_multi_game_b200 import net as N
    torch.set_num_threads(threads)
    model = N.make_random_model(seed=0, in_planes=c["planes"], board=c["board"], actions=c["actions"], blocks=c["blocks"], channels=c["channels"])
    K = _orc.reference() if _orc.have_ref() else None
    kind = "reference" if K is not None else "port"
    if K is None:
        K = _orc.oracle()

    def cb(planes, ch, h, w, a, pol_out, val_out, user):
        x = torch.from_numpy(np.ctypeslib.as_array(planes, shape=(1, ch, h, w)).copy())
        with torch.no_grad():
            p, v = model(x)
            p = torch.softmax(p[0, :a], 0)      # TorchNeuralNetwork::predictBatch softmax (torch_neural_network.cpp:298-316)
        np.ctypeslib.as_array(pol_out, shape=(a,))[:] = p.numpy()
        val_out[0] = float(v)

    cbf = _orc.EVAL_CB(cb)
    state = K.new_state(_orc.GOMOKU, c["board"])
    holder = dict(m=K.mcts_new(state, 100, 1.5, 3, 1, cbf, None), done=0, cb=cbf, state=state)

    def run(n_sims, chunk=100):
        t0 = time.perf_counter()
        left = n_sims
        while left > 0:
            k = min(chunk, left)
            K.mcts_set_sims(holder["m"], k)
            K.mcts_search(holder["m"])
            left -= k; holder["done"] += k
            if holder["done"] >= 800:             # 800 sims per move, then play it (playSingleGame loop)
                a = K.mcts_select_action(holder["m"], 1, 1.0)
                K.mcts_update_with_move(holder["m"], a); K.state_make_move(holder["state"], a); holder["done"] = 0
                if K.state_is_terminal(holder["state"]):
                    holder["state"] = K.new_state(_orc.GOMOKU, c["board"])
                    holder["m"] = K.mcts_new(holder["state"], 100, 1.5, 3, 1, cbf, None)
        return time.perf_counter() - t0

    return run, kind


def run_reference_arm(args, rank):
    if rank != 0:
        return
    c = CONFIGS["gomoku15"]
    threads = os.cpu_count() or 1
    run, kind = reference_searcher(threads, c)
    sims_per_step = args.ref_sims_per_step
    for _ in range(args.warmup):
        run(sims_per_step)
    t = 0.0
    for _ in range(args.steps):
        t += run(sims_per_step)
    v = sims_per_step * args.steps / t
    sample = f"{args.steps} steps x {sims_per_step} simulations of one Gomoku 15x15 game, serial search, fp32 net on {threads} CPU threads"
    cfg = {"workload": f"Gomoku 15x15 @800 sims/move, ONE game: the reference's serial ParallelMCTS (numThreads 1) with the fp32 10-block 128-ch net on {threads} host "
                       f"CPU threads (BASELINE.json configs[0]) — the reference's CPU implementation of the path the engine arm runs for 4096 concurrent games (configs[1])",
           "slots_per_gpu": 1, "sims_per_move": 800, "parallelism": "host CPU only",
           "step": f"a bounded sample: {sims_per_step} simulations of the running game (a move is played every 800)"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(c, args, world, slots, sims):
    return {"workload": f"{c['name']} batched self-play, {slots} concurrent games per GPU, {sims} sims/move, "
                        f"{c['blocks']}-block {c['channels']}-ch random-init ResNet (BASELINE.json configs[{c['idx']}])",
            "slots_per_gpu": slots, "sims_per_move": sims, "parallelism": f"games sharded over {world} GPU(s)",
            "step": "one self-play move on every slot (root expansion + sims waves + move commit)", "stream_groups": args.streams,
            "preroll_moves": args.preroll,
            "net_precision": args.precision,
            "l2": "working set (node pools, activations) >> 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------- engine arm
def measure(key, args, steps, warmup, e2e_steps, rank, world, local, dist, slots=None, sims=None, want_cpu=False):
    """One BASELINE config on this rank's GPU: device-timed value, e2e through the C ABI with host buffers, roofline of the conv kernel."""
    import numpy as np
    import torch
    from alphazero_multi_game_b200 import engine as E, net as N
    c = CONFIGS[key]
    slots = slots or c["slots"]; sims = sims or c["sims"]
    conv_flop_board, net_flop = flops(c)
    eng = E.Engine(game=getattr(E, c["game"]), board_size=c["board"], n_slots=slots, num_simulations=sims, evaluator=E.EVAL_RESNET,
                   net_blocks=c["blocks"], net_channels=c["channels"], deterministic=0, auto_restart=1, device=local, seed=1234 + rank,
                   n_streams=args.streams, net_precision=E.NET_BF16 if args.precision == "bf16" else E.NET_FP16, eval_cache_entries=args.cache)
    model = N.make_random_model(seed=0, in_planes=c["planes"], board=c["board"], actions=c["actions"], blocks=c["blocks"], channels=c["channels"])
    blob = N.export_weights(model)
    eng.load_weights(blob)
    rec_bytes = eng.sample_layout().record_bytes
    ring = max(32 * slots, 4096)                                     # = the engine's default sample ring
    pinned = torch.empty(ring * rec_bytes * (world if dist is not None else 1), dtype=torch.uint8).pin_memory()
    samples_np = pinned[:ring * rec_bytes].numpy().view(eng.sample_dtype())

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

    # Pre-roll (untimed, not part of the W warm-up steps): every slot starts from the initial position, so in the first moves thousands of games
    # walk through the same openings and the evaluation cache / in-wave sharing serve a share of the leaves that a long self-play run (games at
    # every phase) never sees (chess: 44 % at moves 3-8, 13 % at moves 6-11, 9 % at moves 40-60).  A few cheap moves (noise + temperature on, 48
    # simulations each) take the games apart first; the timed steps then run the full configuration on mid-game positions.
    if args.preroll > 0:
        eng.set_num_simulations(min(48, sims))
        eng.play(args.preroll)
        eng.set_num_simulations(sims)
    for _ in range(warmup):
        eng.play(1)
    eng.drain_samples(out=samples_np)

    # ---- timed region 1: device-resident throughput --------------------------------------------------------
    clocks = ClockSampler(local); clocks.start()
    barrier()
    s0 = eng.stats()
    live0 = eng.conv_sampled()
    tm0 = eng.timing()
    ring_dev = torch.empty(ring * rec_bytes, dtype=torch.uint8, device="cuda") if steps > 8 else None
    eng.event_record(0)
    for i in range(steps):
        eng.play(1)
        # long runs: games finish by the hundred; their samples are moved out of the engine's ring into a device buffer (still HBM-resident) so
        # that the ring does not overflow (a full ring drops records and counts them in samples_dropped)
        if ring_dev is not None and i % 8 == 7:
            eng.drain_samples_device(ring_dev.data_ptr(), ring)
    eng.event_record(1)
    ms = eng.event_elapsed(0, 1)
    barrier()
    s1 = eng.stats()
    live1 = eng.conv_sampled()
    tm1 = eng.timing()
    clk = clocks.stop()
    # per-step budget from the engine's own sampled CUDA events (every 64th wave kernel by kernel, every move commit): (sims + 1) waves + one commit
    nw = max(tm1["waves_sampled"] - tm0["waves_sampled"], 1); nm = max(tm1["moves_sampled"] - tm0["moves_sampled"], 1)
    budget = {k[:-3]: (tm1[k] - tm0[k]) / nw * (sims + 1) for k in ("select_ms", "dedup_encode_ms", "stem_ms", "trunk_ms", "head_conv_ms", "conv1x1_gemm_ms",
                                                                     "policy_fc_ms", "value_fc_ms", "policy_value_ms", "expand_backup_ms")}
    budget["move_commit"] = (tm1["commit_ms"] - tm0["commit_ms"]) / nm
    budget_sum = sum(budget.values())
    ms_max = allreduce(ms, dist.ReduceOp.MAX) if dist else ms
    n_sims = allreduce(float(s1["simulations"] - s0["simulations"]), dist.ReduceOp.SUM) if dist else float(s1["simulations"] - s0["simulations"])
    moves = allreduce(float(s1["moves"] - s0["moves"]), dist.ReduceOp.SUM) if dist else float(s1["moves"] - s0["moves"])
    leaf_evals = float(s1["evaluations"] - s0["evaluations"])
    shared = float(s1["eval_shared"] - s0["eval_shared"])             # leaves served by another tree's evaluation of the same input in the same wave
    cached = float(s1["eval_cached"] - s0["eval_cached"])             # ... by the evaluation cache (an earlier wave's evaluation of the same input)
    evals = leaf_evals - shared - cached                               # network evaluations actually run (what the FLOP accounting uses)
    launches = int(s1["kernel_launches"] - s0["kernel_launches"])
    value = n_sims / (ms_max / 1e3)

    # ---- timed region 2: end to end through the C ABI with host buffers ------------------------------------
    # every step: (H2D) a fresh fp32 weight blob from pinned host memory is uploaded, folded and converted — the trainer →
    # self-play hand-off of a real AlphaZero loop, here once per move, i.e. far more often than in production;
    # one self-play move; (D2H) finished-game samples into pinned host memory + the chosen actions + the counters.
    # N > 1: the finished-game samples of all ranks are all-gathered over NCCL (the path's only exchange step) and the
    # gathered records are then copied to pinned host memory on every rank.
    h2d = len(blob)
    blob_pinned = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
    if dist is not None:
        from alphazero_multi_game_b200 import gather as GA
        dev_samples = torch.zeros(ring * rec_bytes, dtype=torch.uint8, device="cuda")
        GA.all_gather_samples(dist, dev_samples, 1, rec_bytes)      # untimed warm-up of the collective (NCCL sets up its channels on first use)
    barrier()
    e0 = eng.stats()
    t0 = time.perf_counter()
    d2h = 0
    gathered_records = 0
    for _ in range(e2e_steps):
        eng.load_weights(blob_pinned.data_ptr(), h2d)
        eng.play(1)
        if dist is None:
            smp = eng.drain_samples(out=samples_np)
            d2h += smp.nbytes
        else:
            n = eng.drain_samples_device(dev_samples.data_ptr(), ring)
            all_samples, per_rank = GA.all_gather_samples(dist, dev_samples, n, rec_bytes)     # NCCL all-gather
            nb = all_samples.numel()
            if nb:
                pinned[:nb].copy_(all_samples.reshape(-1), non_blocking=True); torch.cuda.current_stream().synchronize()
            d2h += nb + 8 * world; gathered_records += nb // rec_bytes
        acts = eng.last_actions(); d2h += acts.nbytes
        eng.stats(); d2h += 88
    barrier()
    t_e2e = time.perf_counter() - t0
    e1 = eng.stats()
    t_e2e = allreduce(t_e2e, dist.ReduceOp.MAX) if dist else t_e2e
    sims_e2e = allreduce(float(e1["simulations"] - e0["simulations"]), dist.ReduceOp.SUM) if dist else float(e1["simulations"] - e0["simulations"])
    e2e = {"value": sims_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // max(e2e_steps, 1), "steps": e2e_steps,
           "what": "load_weights(pinned host blob) + play(1) + drain_samples(pinned host) + last_actions + stats per step"
                   + (f"; N > 1: NCCL all-gather of the finished-game samples, gathered records copied to pinned host memory ({gathered_records} records in the timed region)" if dist else "")}

    # ---- roofline of the dominant kernel (3x3 conv 128->128 on tcgen05) ------------------------------------
    pk = peaks()
    boards_per_launch = (slots + args.streams - 1) // args.streams      # the production launch shape: one stream group
    conv_ms_alone = eng.conv_bench(boards_per_launch, 20)
    conv_flop = conv_flop_board * boards_per_launch
    # live: CUDA events around the conv launches of every 64th wave INSIDE the timed region above (engine stream); the launch
    # shape in a wave is the number of non-terminal leaves, ~ all slots
    live_n = live1[1] - live0[1]
    conv_ms = (live1[0] - live0[0]) / live_n if live_n else conv_ms_alone
    achieved = conv_flop / (conv_ms / 1e3) / 1e12
    nn_ms = eng.nn_bench(slots, 5)
    fused = not os.environ.get("AZ_TRUNK_LAYERED") and key != "go19"
    traffic, traffic_src = None, None
    if key == "gomoku15" and boards_per_launch == 4096:
        # DRAM bytes per launch from an ncu --set full capture of this kernel at this shape (not measured in this run)
        tf = os.path.join(ROOT, "profiles", "r2_trunk_traffic.json")
        if os.path.exists(tf):
            d = json.load(open(tf)); traffic, traffic_src = d.get("per_layer_bytes"), "profiles/r2_trunk_traffic.json: " + d.get("source", "")
        else:
            traffic, traffic_src = 193.8e6, "profiles/r1f_trunk_pair_ncu_raw.csv (round-1 capture: 814 MB read + 3062 MB written per 20-layer launch)"
    roofline = {"bound": "tensor",
                "kernel": (f"k_conv3x3_pair_wide (one 128->128 slice launch over {boards_per_launch} boards; weight-stationary CTA pair with K-split activation stages for the 21-row halo)" if key == "go19" else
                           f"k_trunk_pair (the trunk's {2 * c['blocks']} 128->128 3x3 conv layers over the {boards_per_launch} boards of one stream group as ONE persistent launch of the weight-stationary CTA-pair kernel, cta_group::2; figures are per layer = launch / {2 * c['blocks']})" if fused else
                           f"k_conv3x3_pair (one 128->128 3x3 conv layer over the {boards_per_launch} boards of one stream group; weight-stationary CTA pair, cta_group::2)"),
                "achieved": achieved,
                # kernel timed inside a long step -> the SUSTAINED measured cuBLAS bf16 rate is the denominator; timed alone -> the burst one
                "peak": pk["bf16_sustained"] if live_n else pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / (pk["bf16_sustained"] if live_n else pk["bf16_tflops"]),
                "peak_source": pk["source"] + (" sustained bf16 (kernel timed inside the step)" if live_n else " burst bf16 (kernel timed alone)") + "; fp16 and bf16 share the kind::f16 tensor rate",
                "frac_note": "the sustained peak is a power-capped cuBLAS bf16 GEMM (MEASURED_PEAKS.json), not a hardware limit: frac can exceed 1; nominal dense 16-bit is 2250 TFLOP/s and the tensor pipe also multiplies the padding rows of the position stream, which `achieved` does not count",
                "timed_alone": {"achieved": conv_flop / (conv_ms_alone / 1e3) / 1e12, "peak": pk["bf16_tflops"],
                                "frac": conv_flop / (conv_ms_alone / 1e3) / 1e12 / pk["bf16_tflops"], "peak_source": pk["source"] + " burst bf16",
                                "what": "one k_conv3x3_pair layer launch alone (az_engine_conv_bench)"},
                "traffic": traffic, "traffic_source": traffic_src,
                "launch_ms": conv_ms, "launch_ms_source": f"live: {live_n} layers bracketed by CUDA events inside the timed steps" if live_n else "timed alone (no sampled wave in the timed region)",
                "launch_ms_timed_alone": conv_ms_alone, "flop_per_launch": conv_flop,
                "whole_net_ms": nn_ms, "whole_net_tflops": net_flop * slots / (nn_ms / 1e3) / 1e12}

    # ---- the synchronised opening phase beside it (rank 0; skipped for configs whose step takes seconds) -------
    # every slot restarted at move 0 with an empty evaluation cache, 3 untimed + 3 timed moves: what a fresh batch of games sees while all
    # of them still walk through the same openings (the cache / in-wave sharing serve a large share of the leaves).  Reported, not the value.
    opening = None
    if rank == 0 and args.preroll > 0 and (ms / steps < 1500.0 or key == args.game):
        eng.load_weights(blob); eng.reset_games()
        eng.play(3); eng.drain_samples(out=samples_np)
        eng.sync(); o0 = eng.stats(); eng.event_record(2)
        eng.play(3)
        eng.event_record(3); oms = eng.event_elapsed(2, 3); o1 = eng.stats()
        ol = float(o1["evaluations"] - o0["evaluations"])
        opening = {"value_rank0": float(o1["simulations"] - o0["simulations"]) / (oms / 1e3), "unit": UNIT, "moves": "3-5 of games that all start together at move 0",
                   "eval_cached_frac": float(o1["eval_cached"] - o0["eval_cached"]) / max(ol, 1.0), "eval_shared_frac": float(o1["eval_shared"] - o0["eval_shared"]) / max(ol, 1.0)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own serial search on the host cores --
    cpu = None
    if want_cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            run, kind = reference_searcher(threads, c)
            run(20)
            n, t = 0, 0.0
            while t < args.cpu_baseline_seconds:
                t += run(50); n += 50
            cpu = {"value": n / t, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{n} simulations of one Gomoku 15x15 game (800 sims/move), serial ParallelMCTS, fp32 10x128 net on {threads} host threads, {t:.1f} s"}
        except Exception as ex:      # the checker is optional equipment on the GPU box
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    out = {"value": value, "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
           "config": workload_config(c, args, world, slots, sims),
           "moves_per_sec": moves / (ms_max / 1e3), "nn_evals_per_sec_rank0": evals / (ms / 1e3),
           "eval_shared_frac_rank0": shared / max(leaf_evals, 1.0), "eval_cached_frac_rank0": cached / max(leaf_evals, 1.0),
           "tensor_roofline_frac_in_step": (evals / (ms / 1e3)) * net_flop / 1e12 / pk["bf16_sustained"],
           "step_budget_ms": {**{k: round(v, 3) for k, v in budget.items()}, "sum": round(budget_sum, 2), "sum_over_ms_per_step": round(budget_sum / (ms / steps), 4),
                              "source": f"az_engine_get_timing: {int(nw)} sampled waves and {int(nm)} move commits of rank 0 inside the timed steps, scaled to {sims} + 1 waves and one commit; the sampled waves are launched kernel by kernel with events in between, the others are CUDA-graph replays, so on short waves (chess, Go 9x9) the sum overstates the step by the launch gaps the replay removes"},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "opening_phase": opening, "gpu_launches": launches, "clocks": clk,
           "games_finished": int(e1["games"]), "samples_dropped": int(e1["samples_dropped"]), "pool_overflows": int(e1["pool_overflows"])}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--game", default="gomoku15", choices=list(CONFIGS),
                    help="headline config of the line; default gomoku15 = BASELINE.json configs[1]")
    ap.add_argument("--slots", type=int, default=None)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--streams", type=int, default=1, help="stream groups the slots are split into (tree kernels of one overlap the network pass of the other)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"], help="16-bit storage of activations / conv weights (az_config.net_precision)")
    ap.add_argument("--preroll", type=int, default=8, help="untimed 48-simulation moves played before the warm-up so that the games are at different positions (0 = start every game at move 0)")
    ap.add_argument("--cache", type=int, default=0, help="evaluation cache entries (az_config.eval_cache_entries): 0 = default (4 M), -1 = off")
    ap.add_argument("--others", default="go9,chess,go19", help="comma-separated BASELINE configs measured into `other_configs` ('' = none)")
    ap.add_argument("--other-steps", type=int, default=5)
    ap.add_argument("--ref-sims-per-step", type=int, default=200)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import az_b200_loader
    az_b200_loader.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    head = measure(args.game, args, args.steps, args.warmup, args.steps, rank, world, local, dist, slots=args.slots, sims=args.sims,
                   want_cpu=args.game == "gomoku15")
    others = {}
    if args.game == "gomoku15" and args.slots is None and args.sims is None:
        for k in [x for x in args.others.split(",") if x]:
            o = measure(k, args, max(args.other_steps, 5), 3 if k != "go19" else 2, max(args.other_steps, 5) if k != "go19" else 3, rank, world, local, dist)
            others[k] = {kk: o[kk] for kk in ("value", "unit", "steps", "warmup", "ms_per_step", "config", "moves_per_sec", "eval_shared_frac_rank0", "eval_cached_frac_rank0", "tensor_roofline_frac_in_step", "step_budget_ms", "e2e", "opening_phase",
                                              "gpu_launches", "clocks", "games_finished", "samples_dropped", "pool_overflows")}
            others[k]["roofline"] = {kk: o["roofline"][kk] for kk in ("bound", "kernel", "achieved", "peak", "unit", "frac", "launch_ms", "launch_ms_source", "whole_net_ms", "whole_net_tflops")}
    overflow = head["pool_overflows"] + sum(o["pool_overflows"] for o in others.values())
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "dtype_note": "16-bit tensor-core operands (tcgen05 kind::f16), fp32 accumulation; fp16 is the reference's own half-precision mode "
                                                       "(TorchNeuralNetworkConfig::useFp16) and meets the KL <= 1e-3 tolerance on the BASELINE network; --precision bf16 runs the bf16 storage at the same rate",
                "data": "synthetic"}
        line.update({k: head[k] for k in ("config", "moves_per_sec", "nn_evals_per_sec_rank0", "eval_shared_frac_rank0", "eval_cached_frac_rank0", "tensor_roofline_frac_in_step", "step_budget_ms", "roofline", "cpu_baseline", "e2e", "opening_phase",
                                          "gpu_launches", "clocks", "games_finished", "samples_dropped", "pool_overflows")})
        line["other_configs"] = others
        if overflow:
            line["error"] = f"{overflow} expansion(s) failed for lack of node-pool room: not a valid measurement"
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if overflow:
        sys.exit(3)


if __name__ == "__main__":
    main()
