"""CPU suite: pins the oracle's CHESS restatement and Dataset::augmentExample / extractExamples restatement to the REFERENCE ITSELF.

oracle/_ref now contains the reference's own chess (src/games/chess/*.cpp, with the moveExposesKing -> cloneWithMove -> makeMove
recursion cut by shim 6 of oracle/build_ref.sh — the move is applied without the legality re-check) and its own
src/selfplay/{dataset,game_record}.cpp.  Two layers, as for Gomoku / Go:

* live lock-step against oracle/_ref/libaz_ref.so when it is present (authoring container and GPU box);
* tests/golden/chess_*.json / dataset_examples.json — generated from the reference by tests/golden/gen_golden.py — everywhere.

Rows of SURVEY.md §8: C1-C7 (action code, movegen order, terminal rules, 18 planes, the literal pawn-attack quirk, castling, every
insufficient-material case), (f)1 (Dataset / augmentation / GameRecord JSON)."""
import hashlib
import json
import os

import numpy as np
import pytest

import _orc
from _orc import CHESS, GOMOKU, GO, DRAW, ONGOING, WIN_P1, WIN_P2

GOLD = os.path.join(os.path.dirname(__file__), "golden")
need_ref = pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref not built")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32).tolist()


def code(frm, to, promo=0):
    return (promo << 12) | (frm << 6) | to


# QUIRK C5 positions: (FEN, side whose king is examined is the side to move)
C5_FENS = [
    "8/8/8/4k3/3P4/8/8/4K3 b - - 0 1",      # white pawn d4 really attacks e5: NOT seen by the literal test
    "8/8/3P4/4k3/8/8/8/4K3 b - - 0 1",      # white pawn d6 "behind" the king: seen
    "4k3/8/8/8/3p4/4K3/8/8 w - - 0 1",      # black pawn d4 really attacks e3: not seen
    "4k3/8/8/8/8/4K3/3p4/8 w - - 0 1",      # black pawn d2 behind: seen
    "4k3/8/8/8/8/8/4P3/3K4 w - - 0 1",      # no check at all
    "rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR w KQkq e6 0 2",
]
# C7: every insufficient-material case of chess_rules.cpp:231-389 (+ near misses that are NOT draws)
C7_FENS = [
    ("8/8/4k3/8/8/3K4/8/8 w - - 0 1", DRAW),            # K-K
    ("8/8/4k3/8/8/3KN3/8/8 w - - 0 1", DRAW),           # K+N-K
    ("8/8/4kn2/8/8/3K4/8/8 w - - 0 1", DRAW),           # K-K+N
    ("8/8/4k3/8/8/3KB3/8/8 w - - 0 1", DRAW),           # K+B-K
    ("8/8/4kb2/8/8/3K4/8/8 b - - 0 1", DRAW),           # K-K+B
    ("8/8/4kb2/8/8/3KB3/8/8 w - - 0 1", DRAW),          # K+B-K+B, bishops on the same square colour (f6, e3)
    ("8/8/4kb2/8/8/3BK3/8/8 w - - 0 1", ONGOING),       # K+B-K+B, opposite colours (f6, d3): not a draw
    ("8/8/4k3/8/8/2NKN3/8/8 w - - 0 1", DRAW),          # K+NN-K
    ("8/8/3nkn2/8/8/3K4/8/8 w - - 0 1", DRAW),          # K-K+NN
    ("8/8/4kn2/8/8/3KN3/8/8 w - - 0 1", DRAW),          # K+N-K+N
    ("8/8/4kb2/8/8/3KN3/8/8 w - - 0 1", DRAW),          # K+N-K+B
    ("8/8/4kn2/8/8/3KB3/8/8 w - - 0 1", DRAW),          # K+B-K+N
    ("8/8/4k3/8/8/3KR3/8/8 w - - 0 1", ONGOING),        # rook: sufficient
    ("8/8/4k3/8/8/3KP3/8/8 w - - 0 1", ONGOING),        # pawn: sufficient
    ("8/8/4k3/8/8/2BKN3/8/8 w - - 0 1", ONGOING),       # K+B+N-K: sufficient
    ("8/8/4k3/8/8/3KR3/8/8 w - - 99 1", ONGOING),       # fifty-move boundary
    ("8/8/4k3/8/8/3KR3/8/8 w - - 100 1", DRAW),
]


def _digest(K, s):
    legal = K.legal(s)
    h = hashlib.sha256()
    h.update(legal.astype(np.int32).tobytes())
    h.update(bytes([K.state_is_terminal(s), K.state_result(s), K.state_current_player(s), K.chess_in_check(s)]))
    h.update(K.tensor(s).astype(np.float32).tobytes())
    return h.hexdigest()[:16]


def _same_position(O, R, so, sr, ctx):
    lo, lr = O.legal(so), R.legal(sr)
    assert np.array_equal(lo, lr), ("legal moves / order", ctx)
    assert O.state_is_terminal(so) == R.state_is_terminal(sr), ("terminal", ctx)
    assert O.state_result(so) == R.state_result(sr), ("result", ctx)
    assert O.state_current_player(so) == R.state_current_player(sr), ctx
    assert np.array_equal(O.tensor(so), R.tensor(sr)), ("18 planes", ctx)
    assert O.chess_in_check(so) == R.chess_in_check(sr), ("isInCheck", ctx)
    assert O.state_key(so) == R.state_key(sr), ctx
    return lo


@need_ref
def test_chess_restatement_matches_reference_live():
    """Random games (a third of the moves prefer captures, so the games reach sparse endgames, promotions and all draw rules):
    legal moves in order, terminal flag, result, player, the 18 planes, isInCheck and the hash-evaluator key on EVERY position."""
    O, R = _orc.oracle(), _orc.reference()
    rng = np.random.default_rng(11)
    n_pos, results = 0, set()
    for g in range(24):
        so, sr = O.new_state(CHESS, 8), R.new_state(CHESS, 8)
        for ply in range(450):
            lo = _same_position(O, R, so, sr, (g, ply))
            n_pos += 1
            if O.state_is_terminal(so):
                results.add(O.state_result(so))
                break
            caps = [a for a in lo if O.chess_piece(so, int(a) & 63) != 0]
            a = int(rng.choice(caps)) if (caps and rng.random() < 0.35) else int(rng.choice(lo))
            assert O.state_make_move(so, a) == 0 and R.state_make_move(sr, a) == 0
        O.state_free(so); R.state_free(sr)
    assert n_pos > 3000 and DRAW in results


@need_ref
def test_chess_fen_cases_match_reference_live():
    """C5 (literal pawn-attack direction) and C7 (all seven insufficient-material patterns, fifty-move boundary) on FEN positions,
    plus the reference tests' own FENs (castling / e.p. / promotion / stalemate / Fool's mate) and an illegal-move rejection."""
    O, R = _orc.oracle(), _orc.reference()
    ref_fens = ["r1bqkbnr/pppp1ppp/2n5/4p3/4P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 2 3", "rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3",
                "rnbqkbnr/pppppPpp/8/8/8/8/PPPPPP1P/RNBQKBNR w KQkq - 0 1", "rnbqkbnr/pppp1ppp/8/4p3/2B1P3/5Q2/PPPP1PPP/RNB1K1NR b KQkq - 3 3",
                "8/8/8/8/8/6k1/5q2/7K w - - 0 1", "rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 1 3", "8/P7/8/8/8/8/8/k6K w - - 0 1",
                "r3k2r/pppppppp/8/8/8/8/PPPPPPPP/R3K2R w KQkq - 0 1", "r3k2r/8/8/8/8/8/8/R3K2R b KQkq - 0 1", "r3k2r/8/8/4q3/8/8/8/R3K2R w KQkq - 0 1"]
    for fen in C5_FENS + [f for f, _ in C7_FENS] + ref_fens:
        so, sr = O.chess_from_fen(fen), R.chess_from_fen(fen)
        lo = _same_position(O, R, so, sr, fen)
        for a in lo[:6]:                                   # one ply further from every FEN
            co, cr = O.state_clone(so), R.state_clone(sr)
            assert O.state_make_move(co, int(a)) == 0 and R.state_make_move(cr, int(a)) == 0
            _same_position(O, R, co, cr, (fen, int(a)))
        assert O.state_make_move(so, code(0, 63)) == R.state_make_move(sr, code(0, 63)) == -1     # "Illegal move attempted"
    # the quirk itself, stated: the really-attacking pawn is not seen, the one "behind" is
    assert not R.chess_in_check(R.chess_from_fen(C5_FENS[0])) and R.chess_in_check(R.chess_from_fen(C5_FENS[1]))
    for fen, want in C7_FENS:
        assert R.state_result(R.chess_from_fen(fen)) == want, fen


@need_ref
def test_chess_perft_and_search_match_reference_live():
    O, R = _orc.oracle(), _orc.reference()
    O.chess_set_fide(0)
    so, sr = O.new_state(CHESS, 8), R.new_state(CHESS, 8)
    assert [R.chess_perft(sr, d) for d in (1, 2, 3)] == [O.chess_perft(so, d) for d in (1, 2, 3)] == [20, 400, 8902]
    kiwi = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
    assert R.chess_perft(R.chess_from_fen(kiwi), 2) == O.chess_perft(O.chess_from_fen(kiwi), 2)
    # serial ParallelMCTS over chess with the hash evaluator: the reference's search on the reference's chess vs the oracle's
    rng = np.random.default_rng(5)
    for sims, moves, plies in [(120, 3, 0), (100, 3, 12)]:
        so, sr = O.new_state(CHESS, 8), R.new_state(CHESS, 8)
        for _ in range(plies):
            a = int(rng.choice(O.legal(so)))
            O.state_make_move(so, a); R.state_make_move(sr, a)
        mo, mr = O.mcts_new(so, sims, 1.5, 3, 0, None, None), R.mcts_new(sr, sims, 1.5, 3, 0, None, None)
        for mv in range(moves):
            O.mcts_search(mo); R.mcts_search(mr)
            a, b = O.root_stats(mo), R.root_stats(mr)
            assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), mv
            assert bits(a["W"]) == bits(b["W"]) and bits(a["P"]) == bits(b["P"]) and a["rootN"] == b["rootN"]
            assert bits([O.mcts_root_value(mo)]) == bits([R.mcts_root_value(mr)])
            x, y = O.mcts_select_action(mo, 1, 1.0), R.mcts_select_action(mr, 1, 1.0)
            assert x == y
            O.mcts_update_with_move(mo, x); R.mcts_update_with_move(mr, y)
        O.mcts_free(mo); R.mcts_free(mr)


def test_chess_restatement_matches_reference_golden():
    """The same pins from fixtures generated from the reference's chess (runs where oracle/_ref is absent)."""
    O = _orc.oracle()
    O.chess_set_fide(0)
    gold = json.load(open(os.path.join(GOLD, "chess_reference.json")))
    for case in gold["playouts"]:
        s = O.new_state(CHESS, 8)
        for ply, dg in enumerate(case["ply_digest"]):
            assert _digest(O, s) == dg, (case["seed"], ply)
            if ply < len(case["moves"]):
                assert O.state_make_move(s, case["moves"][ply]) == 0
        assert O.state_result(s) == case["final_result"]
    for fen, dg in gold["fens"].items():
        assert _digest(O, O.chess_from_fen(fen)) == dg, fen
    for case in gold["search"]:
        s = O.new_state(CHESS, 8)
        for a in case["opening"]:
            assert O.state_make_move(s, a) == 0
        m = O.mcts_new(s, case["sims"], 1.5, 3, 0, None, None)
        for mv, g in enumerate(case["moves"]):
            O.mcts_search(m)
            st = O.root_stats(m)
            assert st["actions"].tolist() == g["actions"] and st["N"].tolist() == g["N"], mv
            assert bits(st["W"]) == g["W"] and bits(st["P"]) == g["P"] and st["rootN"] == g["rootN"]
            a = O.mcts_select_action(m, 1, 1.0)
            assert a == g["action"]
            O.mcts_update_with_move(m, a)
        O.mcts_free(m)


# ------------------------------------------------------------------------------------------------ Dataset / GameRecord
def _random_example(rng, c, n, p):
    return rng.random((c, n, n)).astype(np.float32), rng.random(p).astype(np.float32)


def _ref_augment(R, pl, po, game):
    c, n, _ = pl.shape
    opl = np.zeros((7, c, n, n), np.float32); opo = np.zeros((7, len(po)), np.float32)
    R.augment_example(np.ascontiguousarray(pl).ctypes.data, c, n, np.ascontiguousarray(po).ctypes.data, len(po), game, opl.ctypes.data, opo.ctypes.data)
    return opl, opo


@need_ref
def test_augment_example_matches_reference_live():
    """orc_augment_example against the reference's own Dataset::augmentExample (dataset.cpp:245-436): the 7 images in the reference's
    order, planes and policy bit for bit — square policies, Go's N*N + 1 (pass entry stays) and policies SHORTER than N*N (child-ordered
    MoveData.policy: an entry moves only when its old and new index both fit)."""
    R = _orc.reference()
    rng = np.random.default_rng(2)
    for c, n, p, game in [(11, 15, 225, GOMOKU), (8, 9, 82, GO), (8, 19, 362, GO), (11, 9, 81, GOMOKU), (11, 15, 97, GOMOKU), (8, 9, 30, GO), (3, 5, 26, GO), (11, 15, 1, GOMOKU)]:
        pl, po = _random_example(rng, c, n, p)
        a_pl, a_po = _orc.augment_example(pl, po)
        r_pl, r_po = _ref_augment(R, pl, po, game)
        assert np.array_equal(a_pl.view(np.uint32), r_pl.view(np.uint32)), (c, n, p)
        assert np.array_equal(a_po.view(np.uint32), r_po.view(np.uint32)), (c, n, p)


def _oracle_examples(O, game, board, moves, policies, result, augment):
    """Dataset::extractExamples restated with the oracle's state replay + orc_augment_example, in the reference's pre-shuffle order."""
    s = O.new_state(game, board)
    planes, pols, vals = [], [], []
    for i in range(len(moves)):
        if i > 0:
            assert O.state_make_move(s, int(moves[i - 1])) == 0
        t = O.tensor(s)
        gv = 1.0 if result == WIN_P1 else (-1.0 if result == WIN_P2 else 0.0)
        if O.state_current_player(s) == 2:
            gv = -gv
        planes.append(t); pols.append(policies[i]); vals.append(gv)
        if augment and game != CHESS:
            apl, apo = _orc.augment_example(t, policies[i])
            for k in range(7):
                planes.append(apl[k]); pols.append(apo[k]); vals.append(gv)
    return np.stack(planes), np.stack(pols), np.asarray(vals, np.float32)


def _canon(planes, pols, vals):
    """order-free form of an example list (extractExamples shuffles with a random_device seed)"""
    return sorted(hashlib.sha256(planes[i].tobytes() + pols[i].tobytes() + np.float32(vals[i]).tobytes()).hexdigest() for i in range(len(vals)))


def _random_game(O, rng, game, board, plies):
    s = O.new_state(game, board)
    mv = []
    for _ in range(plies):
        if O.state_is_terminal(s):
            break
        lg = O.legal(s)
        cand = lg[1:] if (game == GO and len(lg) > 1 and rng.random() > 0.05) else lg
        a = int(rng.choice(cand))
        assert O.state_make_move(s, a) == 0
        mv.append(a)
    return mv, O.state_result(s)


@need_ref
@pytest.mark.parametrize("game,board,plen", [(GOMOKU, 9, 81), (GO, 9, 82), (GOMOKU, 15, 40), (CHESS, 8, 37)])
def test_extract_examples_matches_reference_live(game, board, plen):
    """Dataset::addGameRecord + extractExamples (dataset.cpp:64-114) run by the reference on a GameRecord vs the oracle's restatement
    (state replay + augmentExample): same multiset of (planes, policy, value) examples, bit for bit; chess is not augmented; the value
    is the game result seen from the player to move; an illegal recorded move throws."""
    O, R = _orc.oracle(), _orc.reference()
    rng = np.random.default_rng(17 + board)
    for result_override in (None, WIN_P2, DRAW):
        mv, res = _random_game(O, rng, game, board, 14)
        res = res if result_override is None else result_override
        if res == ONGOING:
            res = WIN_P1
        pols = rng.random((len(mv), plen)).astype(np.float32)
        for aug in (1, 0):
            k = 8 if (aug and game != CHESS) else 1
            c = O.state_tensor(O.new_state(game, board), None)
            n_ex = R.dataset_extract(game, board, np.asarray(mv, np.int32).ctypes.data, len(mv), pols.ctypes.data, plen, res, aug, None, None, None)
            assert n_ex == len(mv) * k
            r_pl = np.zeros((n_ex, c, board, board), np.float32); r_po = np.zeros((n_ex, plen), np.float32); r_va = np.zeros(n_ex, np.float32)
            assert R.dataset_extract(game, board, np.asarray(mv, np.int32).ctypes.data, len(mv), pols.ctypes.data, plen, res, aug, r_pl.ctypes.data, r_po.ctypes.data, r_va.ctypes.data) == n_ex
            o_pl, o_po, o_va = _oracle_examples(O, game, board, mv, pols, res, aug)
            assert _canon(o_pl, o_po, o_va) == _canon(r_pl, r_po, r_va)
    bad = [0, 0] if game != CHESS else [code(0, 63)] * 2
    assert R.dataset_extract(game, board, np.asarray(bad + [1], np.int32).ctypes.data, 3, np.zeros((3, 4), np.float32).ctypes.data, 4, WIN_P1, 0, None, None, None) == -1


def test_augment_and_examples_match_reference_golden():
    """The same from fixtures generated from the reference's Dataset (digests of every image of every example)."""
    O = _orc.oracle()
    gold = json.load(open(os.path.join(GOLD, "dataset_reference.json")))
    for case in gold["augment"]:
        rng = np.random.default_rng(case["seed"])
        pl, po = _random_example(rng, case["c"], case["n"], case["p"])
        a_pl, a_po = _orc.augment_example(pl, po)
        assert [hashlib.sha256(a_pl[k].tobytes() + a_po[k].tobytes()).hexdigest()[:16] for k in range(7)] == case["images"]
    for case in gold["extract"]:
        pols = np.random.default_rng(case["seed"]).random((len(case["moves"]), case["plen"])).astype(np.float32)
        o = _oracle_examples(O, case["game"], case["board"], case["moves"], pols, case["result"], case["augment"])
        assert hashlib.sha256("".join(_canon(*o)).encode()).hexdigest()[:16] == case["digest"], case["game"]


@need_ref
def test_game_record_json_matches_reference_live():
    """GameRecord::toJson (game_record.cpp:64-90) of the host mirror against the reference's own, key for key and value for value
    (the timestamp is wall-clock in both; compared for format only)."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(_orc.ROOT, "alphazero-multi-game_b200"))
    try:
        import _alphazero_cpp as az
    except ImportError:
        pytest.skip("host mirror not built")
    R = _orc.reference()
    rng = np.random.default_rng(4)
    for game, board, variant, gt in [(GOMOKU, 15, 0, az.GameType.GOMOKU), (GO, 9, 0, az.GameType.GO), (CHESS, 8, 1, az.GameType.CHESS)]:
        n = 5
        moves = rng.integers(0, 60, n).astype(np.int32)
        plens = rng.integers(1, 9, n).astype(np.int32)
        pols = rng.random(int(plens.sum())).astype(np.float32)
        vals = (rng.random(n).astype(np.float32) * 2 - 1)
        think = rng.integers(0, 5000, n).astype(np.int64)
        buf = C.create_string_buffer(1 << 16)
        ln = R.game_record_json(game, board, variant, moves.ctypes.data, n, pols.ctypes.data, plens.ctypes.data, vals.ctypes.data, think.ctypes.data, WIN_P2, buf, 1 << 16)
        ref = json.loads(buf.value[:ln].decode())
        rec = az.GameRecord(gt, board, bool(variant))
        off = 0
        for i in range(n):
            rec.addMove(int(moves[i]), pols[off:off + plens[i]].tolist(), float(vals[i]), int(think[i])); off += int(plens[i])
        rec.setResult(az.GameResult.WIN_PLAYER2)
        mine = json.loads(rec.toJson())
        assert list(mine.keys()) == list(ref.keys())
        for k in ref:
            if k == "timestamp":
                assert type(mine[k]) is type(ref[k])
            elif k == "moves":
                assert len(mine[k]) == len(ref[k])
                for a, b in zip(mine[k], ref[k]):
                    assert list(a.keys()) == list(b.keys()) and a == b
            else:
                assert mine[k] == ref[k], k
        # and the reference's fromJson semantics: the mirror reads the reference's JSON back to the same record
        back = az.GameRecord.fromJson(buf.value[:ln].decode())
        assert [m.action for m in back.getMoves()] == moves.tolist() and back.getResult() == az.GameResult.WIN_PLAYER2


@need_ref
def test_chess_transposition_table_quirk_c8_live():
    """QUIRK C8: the reference's TranspositionTable is keyed by ChessState::getHash(), which after the first move covers the piece
    placement only (chess_state.cpp:230-241, 976-1095: makeMove never marks the hash dirty) — a leaf whose placement was seen before in
    the game gets that earlier position's cached policy / value even when side to move, castling rights or e.p. square differ.  Start
    position, 200 simulations x 3 moves: the third search has such a hit (599 network evaluations instead of 600).  The oracle with its
    TT model equals the reference bit for bit; the TT-free search (what the device engine runs: every leaf on its own input) differs
    exactly there.  Gomoku / Go keys cover the whole evaluator input of the hash evaluator, so their searches are TT-transparent."""
    O, R = _orc.oracle(), _orc.reference()

    def run(K, tt):
        s = K.new_state(CHESS, 8); m = K.mcts_new(s, 200, 1.5, 3, 0, None, None)
        if tt is not None:
            K.mcts_set_tt(m, tt)
        out = []
        for _ in range(3):
            K.mcts_search(m); st = K.root_stats(m); a = K.mcts_select_action(m, 1, 1.0)
            out.append((st["N"].tolist(), bits(st["W"]), bits(st["P"]), a)); K.mcts_update_with_move(m, a)
        ev = K.mcts_eval_calls(m); K.mcts_free(m)
        return out, ev

    ref, ref_ev = run(R, None)
    on, on_ev = run(O, 1)
    off, off_ev = run(O, 0)
    assert ref == on and ref_ev == on_ev == 599
    assert off_ev == 600 and off[:2] == on[:2] and off[2] != on[2]
    # Gomoku / Go: same results with and without the table
    for game, n, sims in [(GOMOKU, 9, 150), (GO, 9, 150)]:
        res = []
        for tt in (1, 0):
            s = O.new_state(game, n); m = O.mcts_new(s, sims, 1.5, 3, 0, None, None); O.mcts_set_tt(m, tt)
            o = []
            for _ in range(4):
                O.mcts_search(m); st = O.root_stats(m); a = O.mcts_select_action(m, 1, 1.0)
                o.append((st["N"].tolist(), bits(st["W"]), a)); O.mcts_update_with_move(m, a)
            res.append(o); O.mcts_free(m)
        assert res[0] == res[1]
