"""Which rounding step of the 16-bit trunk costs the policy KL on the BASELINE network (seed 0, 10 x 128, reference init)?

CPU emulation (fp32 accumulation, selective rounding) of the engine's forward on the 300 positions of
tests/test_nn_gpu.py::test_trunk_matches_fp32_reference_init_model.  Each variant rounds a different subset of
{conv weights, block outputs (the residual stream), mid-block activations, input planes} to bf16 / fp16 and reports
KL(fp32 || variant) and the logit error.  Results: profiles/r2_kl_rounding_experiment.md.

    python tools/kl_rounding_exp.py            (about two minutes on 8 cores)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _orc                                       # noqa: E402
import az_b200_loader                             # noqa: E402

az_b200_loader.load()
from alphazero_multi_game_b200 import net as N  # noqa: E402


def positions(n, board=15, seed=0):
    O = _orc.oracle()
    rng = np.random.default_rng(seed)
    xs = []
    for _ in range(n):
        s = O.new_state(_orc.GOMOKU, board)
        for _ in range(int(rng.integers(0, 70))):
            O.state_make_move(s, int(rng.choice(O.legal(s))))
            if O.state_is_terminal(s):
                break
        xs.append(O.tensor(s))
    return torch.tensor(np.stack(xs))


def rnd(x, kind):
    if kind == "fp32":
        return x
    if kind == "bf16":
        return x.bfloat16().float()
    if kind == "fp16":
        return x.half().float()
    if kind == "bf16x2":                          # hi + lo pair of bf16 values (16 mantissa bits)
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    raise ValueError(kind)


def fold(conv, bn):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv.weight * s[:, None, None, None], bn.bias - bn.running_mean * s


@torch.no_grad()
def forward(m, x, w_kind, stream_kind, mid_kind, in_kind, conv_in_kind=None):
    """stream_kind: storage of block outputs; conv_in_kind: what the next conv reads of it (default: the same);
    mid_kind: storage of the first conv's output inside a block; in_kind: input planes."""
    conv_in_kind = conv_in_kind or stream_kind
    x = rnd(x, in_kind)
    w, b = fold(m.stem, m.stem_bn)
    h = rnd(F.relu(F.conv2d(x, rnd(w, w_kind), b, padding=1)), stream_kind)
    for blk in m.blocks:
        w1, b1 = fold(blk.conv1, blk.bn1)
        w2, b2 = fold(blk.conv2, blk.bn2)
        y = rnd(F.relu(F.conv2d(rnd(h, conv_in_kind), rnd(w1, w_kind), b1, padding=1)), mid_kind)
        h = rnd(F.relu(F.conv2d(y, rnd(w2, w_kind), b2, padding=1) + h), stream_kind)
    # heads as the engine computes them: exact w.r.t. the trunk output it is given (hi/lo weights, hi/lo features)
    h = rnd(h, conv_in_kind)
    if h.shape[-1] != m.pool:
        h = F.adaptive_avg_pool2d(h, (m.pool, m.pool))
    p = m.p_fc(F.relu(m.p_bn(m.p_conv(h))).flatten(1))
    v = torch.tanh(m.v_fc2(F.relu(m.v_fc1(F.relu(m.v_bn(m.v_conv(h))).flatten(1)))))
    return p, v


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    m = N.make_random_model(seed=0)
    x = positions(int(os.environ.get("N_POS", "300")))
    with torch.no_grad():
        p32, v32 = m(x)
    lp32 = F.log_softmax(p32, 1)
    variants = [
        # name, weights, stream, mid, input, conv_in
        ("engine (bf16 everywhere)", "bf16", "bf16", "bf16", "bf16", None),
        ("fp32 weights, bf16 activations", "fp32", "bf16", "bf16", "bf16", None),
        ("bf16 weights, fp32 activations", "bf16", "fp32", "fp32", "fp32", None),
        ("bf16, exact input planes", "bf16", "bf16", "bf16", "fp32", None),
        ("bf16, hi/lo residual stream (convs read hi)", "bf16", "bf16x2", "bf16", "bf16", "bf16"),
        ("bf16, fp32 mid-block activation only", "bf16", "bf16", "fp32", "bf16", None),
        ("bf16 weights, fp16 activations", "bf16", "fp16", "fp16", "fp16", None),
        ("fp16 weights + activations", "fp16", "fp16", "fp16", "fp16", None),
        ("bf16x2 weights, bf16 activations", "bf16x2", "bf16", "bf16", "bf16", None),
        ("fp16 weights, bf16 activations", "fp16", "bf16", "bf16", "bf16", None),
        ("fp16 weights, bf16 activations, exact inputs", "fp16", "bf16", "bf16", "fp32", None),
        ("fp16 weights, bf16 mid, hi/lo stream (read hi)", "fp16", "bf16x2", "bf16", "bf16", "bf16"),
        ("fp16 weights, bf16 mid, fp16 stream", "fp16", "fp16", "bf16", "bf16", None),
        ("fp16 weights, fp16 mid, bf16 stream", "fp16", "bf16", "fp16", "bf16", None),
    ]
    print(f"{'variant':48s} {'max KL':>10s} {'q99 KL':>10s} {'max|dlogit|':>12s} {'rms dlogit':>11s} {'max|dv|':>9s}")
    for name, wk, sk, mk, ik, ck in variants:
        p, v = forward(m, x, wk, sk, mk, ik, ck)
        kl = (lp32.exp() * (lp32 - F.log_softmax(p, 1))).sum(1).numpy()
        dl = (p - p32)
        print(f"{name:48s} {kl.max():10.3e} {np.quantile(kl, 0.99):10.3e} {dl.abs().max():12.3e} {dl.pow(2).mean().sqrt():11.3e} {(v - v32).abs().max():9.2e}")
    print(f"logit std {p32.std():.2f}")


if __name__ == "__main__":
    main()
