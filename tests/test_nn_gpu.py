"""GPU parity of the bf16 tcgen05 trunk + heads against the fp32 PyTorch network (the TorchScript-equivalent
oracle, SURVEY.md §8c "third-party arithmetic": parity unpinned by the reference's own tests, so the oracle is
torch fp32 on CPU).  Tolerance (BASELINE.json north_star): value |Δ| <= 1e-2, policy KL(fp32 || bf16) <= 1e-3."""
import numpy as np
import pytest

import _orc

pytestmark = pytest.mark.gpu


def _positions(n, board=15, seed=0, game=_orc.GOMOKU):
    O = _orc.oracle()
    rng = np.random.default_rng(seed)
    xs = []
    for g in range(n):
        s = O.new_state(game, board)
        for _ in range(int(rng.integers(0, 70))):
            O.state_make_move(s, int(rng.choice(O.legal(s))))
            if O.state_is_terminal(s):
                break
        xs.append(O.tensor(s))
    return np.stack(xs)


def _check(model, n_pos, slots, tag, saturated=False, game=_orc.GOMOKU, board=15, precision=None):
    import torch
    import torch.nn.functional as F
    from _eng import E, N
    eng = E.Engine(game=game, board_size=board, n_slots=slots, evaluator=E.EVAL_RESNET, net_blocks=model.blocks_n,
                   net_channels=model.channels, num_simulations=8, max_nodes_per_tree=4096, deterministic=1,
                   net_precision=E.NET_FP16 if precision is None else precision)
    eng.load_weights(N.export_weights(model))
    x = _positions(n_pos, board=board, game=game)
    pol, val, logits = eng.nn_forward(x, want_logits=True)
    with torch.no_grad():
        torch.set_num_threads(8)
        p32, v32 = model(torch.tensor(x))
    lp32 = F.log_softmax(p32, 1)
    lp = torch.log(torch.tensor(pol).clamp_min(1e-30))
    kl = (lp32.exp() * (lp32 - lp)).sum(1).numpy()
    verr = np.abs(val - v32.numpy().reshape(-1))
    dl = np.abs(logits - p32.numpy())
    line = (f"[{tag}] max KL {kl.max():.3e}  q99 KL {np.quantile(kl, 0.99):.3e}  max |dv| {verr.max():.3e}  "
            f"max |dlogit| {dl.max():.3e}  logit std {p32.std():.2f}")
    print(line)
    import os
    log_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(log_dir):
        open(os.path.join(log_dir, "nn_parity.txt"), "a").write(line + "\n")
    assert np.all(np.isfinite(pol)) and np.allclose(pol.sum(1), 1.0, atol=1e-4)
    assert verr.max() <= 1e-2, f"{tag}: value error {verr.max()}"
    if not saturated:
        assert kl.max() <= 1e-3, f"{tag}: policy KL {kl.max()}"
    else:
        # bf16 storage only.  Reference-init weights give logits with std ~38 (near one-hot policies, value saturated at -1); the
        # 8-bit significand of bf16 WEIGHTS alone moves the logits by 0.16 rms / 0.7 max (tools/kl_rounding_exp.py,
        # profiles/r2_kl_rounding_experiment.md), which exceeds KL 1e-3 wherever the two top logits nearly tie.  Stated tolerance for
        # the bf16 mode on this model: 99 % of positions within 1e-3, worst case within 1e-2, logit error within 3 % of the logit spread.
        assert np.quantile(kl, 0.99) <= 1e-3 and kl.max() <= 1e-2, f"{tag}: KL q99 {np.quantile(kl, 0.99)} max {kl.max()}"
        assert dl.max() <= 0.03 * float(p32.std()), f"{tag}: logit error {dl.max()}"
    eng.close()


@pytest.mark.parametrize("blocks", [0, 1, 10])
def test_trunk_matches_fp32_calibrated_heads(blocks):
    """Non-saturated heads + randomised BatchNorm statistics so folding, residual adds and the heads are all
    exercised with outputs in their sensitive range; blocks=0 isolates the stem conv + heads, 1 adds one
    residual block (both 128-channel conv variants), 10 is the BASELINE network depth."""
    import torch
    from _eng import N
    m = N.make_random_model(seed=1, randomize_bn=True, blocks=blocks)
    with torch.no_grad():
        scale = {0: 1.0, 1: 0.5, 10: 0.03}[blocks]
        gen = torch.Generator().manual_seed(11 + blocks)   # fixed: the heads' calibration must not depend on the global RNG
        m.p_fc.weight *= scale; m.v_fc1.weight *= scale * 0.7; m.v_fc2.weight *= 0.2
        m.p_fc.bias.copy_(torch.rand(m.p_fc.bias.shape, generator=gen) - 0.5)
        m.v_fc1.bias.copy_(0.2 * torch.rand(m.v_fc1.bias.shape, generator=gen) - 0.1)
        m.v_fc2.bias.copy_(0.4 * torch.rand(m.v_fc2.bias.shape, generator=gen) - 0.2)
    _check(m, 37, 64, f"calibrated-{blocks}")


def test_trunk_matches_fp32_reference_init_model():
    """The BASELINE network — `random_model_gomoku_15x15` equivalent: reference init, seed 0, 10 blocks x 128 channels, heads as initialised
    (logit std 38) — in the engine's default fp16 storage (the reference's own half-precision mode, TorchNeuralNetworkConfig::useFp16):
    policy KL <= 1e-3 and |dv| <= 1e-2 on every one of 600 positions, no relaxation."""
    from _eng import N
    _check(N.make_random_model(seed=0), 600, 512, "reference-init-fp16")


def test_trunk_bf16_mode_reference_init_model():
    """The same network with net_precision = AZ_NET_BF16: the relaxed, stated tolerance (see _check)."""
    from _eng import E, N
    _check(N.make_random_model(seed=0), 300, 512, "reference-init-bf16", saturated=True, precision=E.NET_BF16)


@pytest.mark.parametrize("blocks", [1, 10])
def test_bf16_mode_calibrated_heads(blocks):
    """bf16 storage on calibrated heads: the north-star tolerance holds (as in round 1)."""
    import torch
    from _eng import E, N
    m = N.make_random_model(seed=1, randomize_bn=True, blocks=blocks)
    with torch.no_grad():
        scale = {1: 0.5, 10: 0.03}[blocks]
        m.p_fc.weight *= scale; m.v_fc1.weight *= scale * 0.7; m.v_fc2.weight *= 0.2
    _check(m, 37, 64, f"calibrated-bf16-{blocks}", precision=E.NET_BF16)


@pytest.mark.parametrize("game,board,planes", [(_orc.GO, 9, 8), (_orc.GO, 13, 8), (_orc.GOMOKU, 9, 11), (_orc.GO, 19, 8), (_orc.CHESS, 8, 18)])
def test_trunk_matches_fp32_other_boards(game, board, planes):
    """Go 9x9 / 13x13 (8 planes, A = N*N + 1, boards that do not fill whole 128-row MMA tiles), Gomoku 9x9, Go 19x19 (row
    pitch 20: the single-CTA conv kernel, 6 policy tiles) and chess (18 planes -> 32-channel stem, 20480-wide policy head):
    same kernels, same tolerance."""
    import torch
    from _eng import N
    actions = 20480 if game == _orc.CHESS else board * board + (1 if game == _orc.GO else 0)
    m = N.make_random_model(seed=2, randomize_bn=True, blocks=3, in_planes=planes, board=board, actions=actions)
    gen = torch.Generator().manual_seed(7)                 # fixed: the heads' calibration must not depend on the global RNG
    with torch.no_grad():
        # chess: 8x8 is not pooled, so the value FC sees 4x the feature energy of a pooled 15x15 board
        m.p_fc.weight *= 0.2; m.v_fc1.weight *= (0.05 if game == _orc.CHESS else 0.1); m.v_fc2.weight *= 0.2
        m.p_fc.bias.copy_(torch.rand(m.p_fc.bias.shape, generator=gen) - 0.5)
        m.v_fc2.bias.copy_(0.4 * torch.rand(m.v_fc2.bias.shape, generator=gen) - 0.2)
    _check(m, 41, 64, f"game{game}-{board}x{board}", game=game, board=board)


@pytest.mark.parametrize("game,board,planes", [(_orc.GOMOKU, 15, 11), (_orc.GO, 19, 8)])
def test_trunk_256_channels_matches_fp32(game, board, planes):
    """256-channel trunk (BASELINE.json configs[3]: Go 19x19, 256 channels): every 256 -> 256 layer runs as 2 x 2 launches of
    the 128 -> 128 kernels over channel slices, partial sums accumulated through the residual path — weight-stationary CTA
    pairs at 15x15, the single-CTA kernel at 19x19."""
    import torch
    from _eng import N
    actions = board * board + (1 if game == _orc.GO else 0)
    m = N.make_random_model(seed=5, randomize_bn=True, blocks=2, channels=256, in_planes=planes, board=board, actions=actions)
    gen = torch.Generator().manual_seed(13)
    with torch.no_grad():
        m.p_fc.weight *= 0.2; m.v_fc1.weight *= 0.1; m.v_fc2.weight *= 0.2
        m.p_fc.bias.copy_(torch.rand(m.p_fc.bias.shape, generator=gen) - 0.5)
    _check(m, 23, 32, f"256ch-game{game}-{board}x{board}", game=game, board=board)


def test_two_kernel_heads_path_still_matches_fp32():
    """The fused heads kernel (head_conv.cu) is the default on 128-channel trunks; the k_pool + 1x1-GEMM path it replaced stays in use for
    256-channel trunks and Go 19x19.  AZ_NO_HEAD_FUSION=1 forces it on the Gomoku network too (the switch is read once per process, so this
    runs the calibrated 0 / 1 / 10-block checks in a child process)."""
    import os, subprocess, sys
    env = dict(os.environ, AZ_NO_HEAD_FUSION="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_nn_gpu.py") + "::test_trunk_matches_fp32_calibrated_heads", "-x", "-q",
                        "-m", "gpu"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _forward_dump(n, blocks, path, game=0, board=15, planes=11, actions=225):
    """nn_forward over n random positions (default Gomoku 15x15) → npz (policy, value, logits); run in-process or as a child with other switches."""
    from _eng import E, N
    m = N.make_random_model(seed=3, randomize_bn=True, blocks=blocks, in_planes=planes, board=board, actions=actions)
    eng = E.Engine(game=game, board_size=board, n_slots=n, evaluator=E.EVAL_RESNET, net_blocks=blocks, num_simulations=8,
                   max_nodes_per_tree=2048, deterministic=1)
    eng.load_weights(N.export_weights(m))
    rng = np.random.default_rng(5)
    x = (rng.random((n, planes, board, board)) < 0.15).astype(np.float32)
    pol, val, logits = eng.nn_forward(x, want_logits=True)
    np.savez(path, pol=pol, val=val, logits=logits)
    eng.close()


@pytest.mark.parametrize("n,blocks,game,board,planes,actions", [(100, 2, 0, 15, 11, 225), (600, 2, 0, 15, 11, 225), (898, 2, 0, 15, 11, 225), (1100, 10, 0, 15, 11, 225),
                                                               (1500, 2, 2, 9, 8, 82), (700, 2, 2, 13, 8, 170), (2000, 2, 1, 8, 18, 20480), (333, 2, 0, 9, 11, 81)])
def test_fused_trunk_bit_identical_to_layered(n, blocks, game, board, planes, actions, tmp_path):
    """k_trunk_pair (all layers in one persistent launch, per-pair groups of 7 boards) against the per-layer launches (AZ_TRUNK_LAYERED=1 in a
    child process: the switch is read once): same MMAs and epilogue arithmetic → bit-identical outputs.  The board counts put 2 (n = 100),
    7 + 1 / 2 (n = 600), 7 + 5 / 6 (n = 898) and 7 + 7 + 0 / 1 (n = 1100) items on a CTA pair, i.e. full groups (two publications per layer), shorter
    batched groups and the short-tail path (one publication per layer).  Go 9x9 / 13x13, chess and Gomoku 9x9: boards that are not one
    256-row work item each (board-aligned groups on a pair-local item grid, rows past a group's end computed but not stored)."""
    import os, subprocess, sys
    a, b = str(tmp_path / "fused.npz"), str(tmp_path / "layered.npz")
    _forward_dump(n, blocks, a, game, board, planes, actions)
    here = os.path.dirname(os.path.abspath(__file__))
    code = f"import sys; sys.path.insert(0, {here!r}); import test_nn_gpu as t; t._forward_dump({n}, {blocks}, {b!r}, {game}, {board}, {planes}, {actions})"
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, AZ_TRUNK_LAYERED="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fa, fb = np.load(a), np.load(b)
    for k in ("pol", "val", "logits"):
        assert np.array_equal(fa[k].view(np.uint32), fb[k].view(np.uint32)), k


def test_fused_trunk_repeatable_at_full_batch():
    """4096 boards x 20 layers in one persistent launch, ten times on the same inputs: every repetition bit-identical (the kernel synchronises
    its two CTAs per layer through global memory + barriers; a missed ordering would show up as run-to-run differences)."""
    from _eng import E, N
    n = 4096
    m = N.make_random_model(seed=4, randomize_bn=True, blocks=10)
    eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=n, evaluator=E.EVAL_RESNET, net_blocks=10, num_simulations=8,
                   max_nodes_per_tree=2048, deterministic=1)
    eng.load_weights(N.export_weights(m))
    rng = np.random.default_rng(6)
    x = (rng.random((n, 11, 15, 15)) < 0.2).astype(np.float32)
    p0, v0, l0 = eng.nn_forward(x, want_logits=True)
    assert np.all(np.isfinite(l0))
    for _ in range(9):
        p, v, l = eng.nn_forward(x, want_logits=True)
        assert np.array_equal(l.view(np.uint32), l0.view(np.uint32)) and np.array_equal(v.view(np.uint32), v0.view(np.uint32))
    # a permutation of the boards permutes the outputs (batch-position independence across the pairs' board groups)
    perm = rng.permutation(n)
    p, v, l = eng.nn_forward(x[perm], want_logits=True)
    assert np.array_equal(l.view(np.uint32), l0[perm].view(np.uint32)) and np.array_equal(v.view(np.uint32), v0[perm].view(np.uint32))
    eng.close()
