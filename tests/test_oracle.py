"""CPU suite, part 1: pin the oracle (CPU restatement) to the reference.

* against tests/golden/*.json — generated from the patched reference itself (oracle/_ref) by
  tests/golden/gen_golden.py; bit-exact (floats compared as IEEE bit patterns);
* against oracle/_ref/libaz_ref.so directly when it is present (authoring container / GPU box);
* against the known-answer cases of the reference's own gtest files.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import _orc
from _orc import GOMOKU, GO, DRAW, ONGOING, WIN_P1, WIN_P2

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32).tolist()


def _run_search_case(K, case):
    s = K.new_state(case["game"], case["board"])
    m = K.mcts_new(s, case["sims"], case["cpuct"], case["virtual_loss"], 0, None, None)
    for mv, g in enumerate(case["moves"]):
        K.mcts_search(m)
        st = K.root_stats(m)
        assert st["actions"].tolist() == g["actions"], f"child order, move {mv}"
        assert st["N"].tolist() == g["N"], f"visit counts, move {mv}"
        assert bits(st["W"]) == g["W"], f"valueSum bits, move {mv}"
        assert bits(st["P"]) == g["P"], f"prior bits, move {mv}"
        assert st["rootN"] == g["rootN"] and bits([st["rootW"]])[0] == g["rootW"]
        assert bits([K.mcts_root_value(m)])[0] == g["root_value"]
        a = K.mcts_select_action(m, 1, 1.0)
        assert a == g["action"]
        K.mcts_update_with_move(m, a)
    assert K.mcts_eval_calls(m) == case["evals"]
    K.mcts_free(m)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_search_matches_golden(idx):
    case = json.load(open(os.path.join(GOLD, "search_hash_eval.json")))[idx]
    _run_search_case(_orc.oracle(), case)


def _ply_digest(K, s):
    legal = K.legal(s)
    h = hashlib.sha256()
    h.update(legal.astype(np.int32).tobytes())
    h.update(bytes([K.state_is_terminal(s), K.state_result(s), K.state_current_player(s)]))
    h.update(K.tensor(s).astype(np.float32).tobytes())
    return h.hexdigest()[:16]


def test_oracle_state_playouts_match_golden():
    K = _orc.oracle()
    for case in json.load(open(os.path.join(GOLD, "state_playouts.json"))):
        s = K.new_state(case["game"], case["board"])
        for ply, dg in enumerate(case["ply_digest"]):
            assert _ply_digest(K, s) == dg, (case["game"], case["board"], case["seed"], ply)
            if ply < len(case["moves"]):
                assert K.state_make_move(s, case["moves"][ply]) == 0
        assert K.state_is_terminal(s) == case["final_terminal"]
        assert K.state_result(s) == case["final_result"]


def test_first_fill_order_matches_golden():
    gold = json.load(open(os.path.join(GOLD, "gomoku_first_fill_order.json")))
    for n, order in gold.items():
        n = int(n)
        assert _orc.first_fill_order(np.arange(n * n)).tolist() == order
        K = _orc.oracle()
        s = K.new_state(GOMOKU, n)
        assert K.legal(s).tolist() == order
        K.state_make_move(s, 0)
        assert K.legal(s).tolist() == list(range(n * n - 1, 0, -1))   # every later fill: descending


@pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_reference_live():
    """Lock-step random playouts + searches, oracle vs the patched reference (fresh seeds)."""
    O, R = _orc.oracle(), _orc.reference()
    rng = np.random.default_rng(7)
    for game, n, ngames in [(GOMOKU, 15, 6), (GO, 9, 10)]:
        for _ in range(ngames):
            so, sr = O.new_state(game, n), R.new_state(game, n)
            for ply in range(300):
                lo, lr = O.legal(so), R.legal(sr)
                assert np.array_equal(lo, lr)
                assert O.state_is_terminal(so) == R.state_is_terminal(sr)
                assert O.state_result(so) == R.state_result(sr)
                assert np.array_equal(O.tensor(so), R.tensor(sr))
                assert O.state_key(so) == R.state_key(sr)
                if O.state_is_terminal(so):
                    break
                cand = lo[1:] if (game == GO and len(lo) > 1 and rng.random() > 0.03) else lo
                a = int(rng.choice(cand))
                assert O.state_make_move(so, a) == 0 and R.state_make_move(sr, a) == 0
    # enumerate=True: both states had their legal moves listed before the search (later fills are
    # descending); enumerate=False: a never-enumerated 6-ply state ⇒ the root's first fill is the
    # piecewise libstdc++ order over 219 cells (QUIRK G2) — the oracle must reproduce both.
    for game, n, sims, moves, enumerate_ in [(GOMOKU, 15, 300, 5, True), (GOMOKU, 15, 300, 3, False),
                                             (GO, 9, 200, 5, True)]:
        so, sr = O.new_state(game, n), R.new_state(game, n)
        # start from a random 6-ply opening so the trees differ from the golden ones
        occupied = set()
        for _ in range(6):
            if enumerate_:
                l = O.legal(so); R.legal(sr)
                a = int(rng.choice(l[1:] if game == GO else l))
            else:
                a = int(rng.choice([x for x in range(n * n) if x not in occupied]))
            occupied.add(a)
            O.state_make_move(so, a); R.state_make_move(sr, a)
        mo, mr = O.mcts_new(so, sims, 1.5, 3, 0, None, None), R.mcts_new(sr, sims, 1.5, 3, 0, None, None)
        for mv in range(moves):
            O.mcts_search(mo); R.mcts_search(mr)
            a, b = O.root_stats(mo), R.root_stats(mr)
            for k in ("actions", "N"):
                assert np.array_equal(a[k], b[k]), (k, mv)
            assert bits(a["W"]) == bits(b["W"]) and bits(a["P"]) == bits(b["P"])
            assert a["rootN"] == b["rootN"]
            x, y = O.mcts_select_action(mo, 1, 1.0), R.mcts_select_action(mr, 1, 1.0)
            assert x == y
            O.mcts_update_with_move(mo, x); R.mcts_update_with_move(mr, y)


@pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_reference_live_other_boards():
    """The board sizes the GPU tests compare with the restatement only (Go 13x13 and 19x19 — BASELINE configs[3] —, Gomoku 9x9): lock-step
    random playouts (legal order, terminal flag, result, planes, key on every position) and serial searches from mid-game positions,
    oracle vs the patched reference, bit for bit."""
    O, R = _orc.oracle(), _orc.reference()
    rng = np.random.default_rng(23)
    for game, n, ngames, plies in [(GO, 13, 2, 260), (GO, 19, 1, 420), (GOMOKU, 9, 6, 81)]:
        for _ in range(ngames):
            so, sr = O.new_state(game, n), R.new_state(game, n)
            for ply in range(plies):
                lo, lr = O.legal(so), R.legal(sr)
                assert np.array_equal(lo, lr), (game, n, ply)
                assert O.state_is_terminal(so) == R.state_is_terminal(sr)
                assert O.state_result(so) == R.state_result(sr)
                if ply % 7 == 0:
                    assert np.array_equal(O.tensor(so), R.tensor(sr)), (game, n, ply)
                assert O.state_key(so) == R.state_key(sr)
                if O.state_is_terminal(so):
                    break
                cand = lo[1:] if (game == GO and len(lo) > 1 and rng.random() > 0.02) else lo
                a = int(rng.choice(cand))
                assert O.state_make_move(so, a) == 0 and R.state_make_move(sr, a) == 0
    for game, n, sims, opening in [(GO, 13, 400, 60), (GO, 19, 300, 150), (GOMOKU, 9, 200, 8)]:
        so, sr = O.new_state(game, n), R.new_state(game, n)
        for _ in range(opening):
            l = O.legal(so); R.legal(sr)
            a = int(rng.choice(l[1:] if game == GO else l))
            assert O.state_make_move(so, a) == 0 and R.state_make_move(sr, a) == 0
        assert not O.state_is_terminal(so)
        mo, mr = O.mcts_new(so, sims, 1.5, 3, 0, None, None), R.mcts_new(sr, sims, 1.5, 3, 0, None, None)
        for mv in range(2):
            O.mcts_search(mo); R.mcts_search(mr)
            a, b = O.root_stats(mo), R.root_stats(mr)
            for k in ("actions", "N"):
                assert np.array_equal(a[k], b[k]), (game, n, k, mv)
            assert bits(a["W"]) == bits(b["W"]) and bits(a["P"]) == bits(b["P"]), (game, n, mv)
            assert a["rootN"] == b["rootN"]
            x, y = O.mcts_select_action(mo, 1, 1.0), R.mcts_select_action(mr, 1, 1.0)
            assert x == y
            O.mcts_update_with_move(mo, x); R.mcts_update_with_move(mr, y)


# ------------------------------------------------------------------ reference gtest known answers
def _checkers():
    ks = [_orc.oracle()]
    if _orc.have_ref():
        ks.append(_orc.reference())
    return ks


def test_gomoku_known_answers():
    # tests/games/gomoku/gomoku_state_test.cpp:73-94 (five in a row), :96-124 (legal set / bounds)
    for K in _checkers():
        s = K.new_state(GOMOKU, 15)
        assert len(K.legal(s)) == 225
        assert K.state_make_move(s, -1) != 0 and K.state_make_move(s, 225) != 0
        c = 7 * 15 + 7
        assert K.state_make_move(s, c) == 0
        l = K.legal(s)
        assert len(l) == 224 and c not in l
        assert K.state_make_move(s, c) != 0            # occupied
        s = K.new_state(GOMOKU, 15)
        for i in range(5):                              # black A1..E1 (row 0), white on row 1
            assert K.state_is_terminal(s) == 0
            K.state_make_move(s, i)
            if i < 4:
                K.state_make_move(s, 15 + i)
        assert K.state_is_terminal(s) == 1 and K.state_result(s) == WIN_P1
        # QUIRK G3: a black overline (6) is not a win, a white one is
        for colour_first, expect in ((1, ONGOING), (2, WIN_P2)):
            s = K.new_state(GOMOKU, 15)
            if colour_first == 2:
                K.state_make_move(s, 224)               # burn a black move
            order = [0, 1, 2, 4, 5, 3]                  # fill the gap last → run of 6
            other = [30, 32, 34, 36, 38, 40]
            for i in range(6):
                K.state_make_move(s, order[i])
                if i < 5:
                    K.state_make_move(s, other[i] + (100 if colour_first == 2 else 0))
            assert K.state_result(s) == expect


def test_go_known_answers():
    # tests/games/go/go_state_test.cpp:68-83 (two passes), :109-134 (suicide),
    # tests/integration/go_integration_test.cpp:187-226 (capture), :248-295 (ko bookkeeping)
    for K in _checkers():
        s = K.new_state(GO, 9)
        l = K.legal(s)
        assert len(l) == 82 and l[0] == -1 and l[1] == 0
        assert K.go_ko(s) == -1
        K.state_make_move(s, -1)
        assert K.go_ko(s) == -1 and K.state_is_terminal(s) == 0
        K.state_make_move(s, -1)
        assert K.state_is_terminal(s) == 1
        # suicide: black surrounds E5 (4,4); white may not play there
        s = K.new_state(GO, 9)
        P = lambda x, y: y * 9 + x
        for b, w in zip([P(4, 3), P(3, 4), P(5, 4), P(4, 5)], [P(0, 0), P(8, 8), P(0, 8), None]):
            K.state_make_move(s, b)
            if w is not None:
                K.state_make_move(s, w)
        assert P(4, 4) not in K.legal(s)
        # capture of a single stone
        s = K.new_state(GO, 9)
        seq = [P(1, 0), P(0, 0), P(0, 1)]               # B, W(corner), B captures
        for a in seq:
            assert K.state_make_move(s, a) == 0
        assert K.go_stone(s, P(0, 0)) == 0 and K.go_captured(s, 1) == 1
        assert K.go_ko(s) == P(0, 0)                    # Go4: one group of one stone ⇒ ko point


def test_augment_example_restatement_is_the_dihedral_group():
    """Dataset::augmentExample restatement (dataset.cpp:245-436): 7 distinct images of an asymmetric plane, each a bijection of the
    cells, policy entries moved with their cells, the entry past N*N (Go's pass) untouched, rot90 applied 4 times = identity."""
    rng = np.random.default_rng(0)
    n, c, a = 5, 3, 26
    pl = rng.random((c, n, n)).astype(np.float32); po = rng.random(a).astype(np.float32)
    opl, opo = _orc.augment_example(pl, po)
    imgs = [pl] + [opl[k] for k in range(7)]
    assert len({im.tobytes() for im in imgs}) == 8
    for k in range(7):
        assert np.array_equal(np.sort(opl[k].ravel()), np.sort(pl.ravel())) and opo[k][25] == po[25]
        src = {float(pl[0].ravel()[i]): float(po[i]) for i in range(25)}          # the policy entry follows its cell
        assert all(src[float(opl[k][0].ravel()[i])] == float(opo[k][i]) for i in range(25))
    assert np.array_equal(opl[0][0], np.rot90(pl[0], -1)) and np.array_equal(opl[1][0], np.rot90(pl[0], 2)) and np.array_equal(opl[3][0], pl[0][:, ::-1])
    r = pl
    for _ in range(4):
        r = _orc.augment_example(r, po)[0][0]
    assert np.array_equal(r, pl)
