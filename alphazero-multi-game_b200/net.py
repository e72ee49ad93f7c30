"""Policy/value ResNet definition (fp32, PyTorch) + weight export for the engine.

Architecture = the reference's head shape (python/alphazero/models/ddw_randwire.py:175-184,203-235) behind a
plain residual trunk (ResidualBlock of ddw_randwire.py:27-44 without the SE branch) — SURVEY.md §8a N1/N1b:
stem conv3x3(no bias)+BN+ReLU → `blocks` x [conv3x3+BN+ReLU, conv3x3+BN, +skip, ReLU] → adaptive_avg_pool2d to
min(8,H) → policy: conv1x1(→32, no bias)+BN+ReLU → FC(32*8*8 → A) raw logits; value: conv1x1(→32)+BN+ReLU →
FC(→256)+ReLU → FC(256→1) → tanh.  Random init per _initialize_weights (ddw_randwire.py:189-201).

This module is the fp32 parity reference for the bf16 tcgen05 trunk (tests/test_nn_gpu.py) and the
source of the AZW1 weight blob; it is never on the engine's compute path.
"""
import struct

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class ResidualBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False); self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False); self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class PolicyValueNet(nn.Module):
    def __init__(self, in_planes=11, board=15, actions=225, blocks=10, channels=128):
        super().__init__()
        self.in_planes, self.board, self.actions, self.blocks_n, self.channels = in_planes, board, actions, blocks, channels
        self.stem = nn.Conv2d(in_planes, channels, 3, padding=1, bias=False); self.stem_bn = nn.BatchNorm2d(channels)
        self.blocks = nn.ModuleList([ResidualBlock(channels) for _ in range(blocks)])
        self.pool = min(8, board)
        feat = 32 * self.pool * self.pool
        self.p_conv = nn.Conv2d(channels, 32, 1, bias=False); self.p_bn = nn.BatchNorm2d(32); self.p_fc = nn.Linear(feat, actions)
        self.v_conv = nn.Conv2d(channels, 32, 1, bias=False); self.v_bn = nn.BatchNorm2d(32)
        self.v_fc1 = nn.Linear(feat, 256); self.v_fc2 = nn.Linear(256, 1)
        self._init()

    def _init(self):  # ddw_randwire.py:189-201
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1); nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        x = F.relu(self.stem_bn(self.stem(x)))
        for b in self.blocks:
            x = b(x)
        if x.shape[-1] != self.pool or x.shape[-2] != self.pool:
            x = F.adaptive_avg_pool2d(x, (self.pool, self.pool))
        p = F.relu(self.p_bn(self.p_conv(x))).flatten(1)
        p = self.p_fc(p)
        v = F.relu(self.v_bn(self.v_conv(x))).flatten(1)
        v = torch.tanh(self.v_fc2(F.relu(self.v_fc1(v))))
        return p, v


def make_random_model(seed=0, randomize_bn=False, **kw):
    """`random_model_gomoku_15x15` equivalent: torch.manual_seed(seed), eval mode, default BN statistics.
    randomize_bn=True perturbs BN affine/running stats (used by tests so BN folding is actually exercised)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = PolicyValueNet(**kw).eval()
    if randomize_bn:
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, nn.BatchNorm2d):
                    mod.weight.uniform_(0.6, 1.2); mod.bias.uniform_(-0.2, 0.2)
                    mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.6, 1.4)
    torch.random.set_rng_state(g)
    return m


def export_weights(model: PolicyValueNet) -> bytes:
    """AZW1 blob consumed by az_engine_load_weights (csrc/engine.cu: Net::load): header + fp32 tensors."""
    out = [b"AZW1", struct.pack("<7i", 1, model.blocks_n, model.channels, model.in_planes, model.board, model.board, model.actions)]

    def t(x):
        out.append(np.ascontiguousarray(x.detach().cpu().numpy(), np.float32).tobytes())

    def bn(b):
        t(b.weight); t(b.bias); t(b.running_mean); t(b.running_var)

    t(model.stem.weight); bn(model.stem_bn)
    for blk in model.blocks:
        t(blk.conv1.weight); bn(blk.bn1); t(blk.conv2.weight); bn(blk.bn2)
    t(model.p_conv.weight.reshape(32, -1)); bn(model.p_bn); t(model.p_fc.weight); t(model.p_fc.bias)
    t(model.v_conv.weight.reshape(32, -1)); bn(model.v_bn); t(model.v_fc1.weight); t(model.v_fc1.bias)
    t(model.v_fc2.weight.reshape(-1)); t(model.v_fc2.bias)
    return b"".join(out)


def trace_torchscript(model: PolicyValueNet):
    """TorchScript export traced at the real board size (tracing at 8x8 as the reference's
    python/scripts/simple_export.py:149 does would bake out the pooling branch)."""
    ex = torch.zeros(1, model.in_planes, model.board, model.board)
    return torch.jit.trace(model, ex)
