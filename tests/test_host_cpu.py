"""CPU suite, part 2: host-side logic and the C-ABI surface (no GPU compute)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    import az_b200_loader
    az_b200_loader.load()
    from alphazero_multi_game_b200 import engine
    path = engine.library_path()
    if not os.path.exists(path):
        engine.build_library()
    hdr = open(os.path.join(ROOT, "include", "az_b200.h")).read()
    declared = sorted(set(re.findall(r"AZ_API\s+[\w\s\*]+?\b(az_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/az_b200.h but not exported"
    assert sorted(engine.EXPORTS) == declared


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from _eng import E
    with pytest.raises(E.EngineError, match="no CUDA device"):
        E.Engine(n_slots=4)


def test_product_rules_header_matches_oracle_on_host():
    """The product's bit-board rules (csrc/gomoku.cuh, host+device code) compiled with g++ and checked against
    the oracle: result / winner (incl. black-exactly-5), HashEvaluator key, 11 feature planes."""
    so = "/tmp/az_rules_host_shim.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "alphazero-multi-game_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "rules_host_shim.cpp"), "-o", so])
    L = C.CDLL(so)
    O = _orc.oracle()
    rng = np.random.default_rng(3)

    def check(n, moves, s):
        mv = np.array(moves, np.int32)
        res, pl, win = C.c_int(), C.c_int(), C.c_int()
        key = C.c_ulonglong()
        planes = np.zeros((11, n, n), np.float32)
        assert L.host_gomoku_replay(n, mv.ctypes.data, len(moves), C.byref(res), C.byref(pl), C.byref(key),
                                    planes.ctypes.data, C.byref(win)) == 0
        assert res.value == O.state_result(s) and pl.value == O.state_current_player(s)
        assert key.value == O.state_key(s)
        assert np.array_equal(planes, O.tensor(s))

    for n, games in ((15, 60), (9, 60)):
        for _ in range(games):
            s = O.new_state(_orc.GOMOKU, n); moves = []
            while True:
                check(n, moves, s)
                if O.state_is_terminal(s):
                    break
                l = O.legal(s)
                if moves and rng.random() < 0.5:      # bias towards lines so wins and overlines occur
                    ref = moves[-2] if len(moves) > 1 else moves[-1]
                    a = int(min(l, key=lambda c: abs(c - ref - 1) + rng.random() * 3))
                else:
                    a = int(rng.choice(l))
                O.state_make_move(s, a); moves.append(a)
    # explicit overline cases (QUIRK G3)
    for first_white in (False, True):
        s = O.new_state(_orc.GOMOKU, 15); moves = []
        seq = [0, 30, 1, 32, 2, 34, 4, 36, 5, 38, 3]
        if first_white:
            seq = [224] + [x + (100 if i % 2 == 1 else 0) for i, x in enumerate(seq)]
            seq = [224, 0, 130, 1, 132, 2, 134, 4, 136, 5, 138, 3]
        for a in seq:
            O.state_make_move(s, a); moves.append(a)
        check(15, moves, s)


def test_weight_blob_layout_roundtrip():
    import struct
    from _eng import N
    m = N.make_random_model(seed=3, blocks=2)
    blob = N.export_weights(m)
    assert blob[:4] == b"AZW1"
    ver, blocks, ch, inp, h, w, a = struct.unpack("<7i", blob[4:32])
    assert (ver, blocks, ch, inp, h, w, a) == (1, 2, 128, 11, 15, 15, 225)
    n_float = (len(blob) - 32) // 4
    expect = 128 * 11 * 9 + 4 * 128 + 2 * 2 * (128 * 128 * 9 + 4 * 128) + 2 * (32 * 128 + 128) + 225 * 2048 + 225 + 256 * 2048 + 256 + 256 + 1
    assert n_float == expect
