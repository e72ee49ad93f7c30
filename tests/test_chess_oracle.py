"""Pins the chess restatement of the oracle (oracle/az_oracle.cpp: struct Chess).  The reference's own chess cannot run
(unbounded makeMove -> isLegalMove -> moveExposesKing -> cloneWithMove -> makeMove recursion), so the pins
are the known answers the reference's tests state — tests/games/chess/chess_state_test.cpp:27-65 (32 pieces, white to
move, all castling rights, 20 legal moves, player alternation), :92-180 (the castling / en-passant / promotion / Scholar's
mate / stalemate FENs), tests/integration/chess_integration_test.cpp:96-121 (Fool's mate: white in check, no legal moves,
terminal, WIN_PLAYER2) — plus perft from the start position as external sanity."""
import numpy as np

import _orc
from _orc import CHESS

A1, E1, G1, C1, E8 = 56, 60, 62, 58, 4


def code(frm, to, promo=0):
    return (promo << 12) | (frm << 6) | to


def test_initial_state_known_answers():
    O = _orc.oracle()
    s = O.new_state(CHESS, 8)
    assert O.state_board_size(s) == 8 and O.state_current_player(s) == 1 and O.state_action_space(s) == 20480
    assert sum(1 for q in range(64) if O.chess_piece(s, q) != 0) == 32
    assert O.chess_piece(s, E1) == 6 + 8 * 1 and O.chess_piece(s, E8) == 6 + 8 * 2          # kings, different colours
    lg = O.legal(s)
    assert len(lg) == 20                                                                    # chess_state_test.cpp:64
    # movegen order (chess_rules.cpp:57-98): squares ascending, so a2-a3, a2-a4, b2-b3, ... then the knights
    assert lg[:4].tolist() == [code(48, 40), code(48, 32), code(49, 41), code(49, 33)]
    assert lg[16:].tolist() == [code(57, 40), code(57, 42), code(62, 45), code(62, 47)]
    assert O.state_make_move(s, int(lg[0])) == 0 and O.state_current_player(s) == 2
    assert not np.array_equal(O.legal(s), lg)
    assert O.state_make_move(s, code(48, 40)) != 0                                         # illegal: throws in the reference
    t = O.tensor(s)
    assert t.shape == (18, 8, 8) and t[12].max() == 0.0 and t[13].min() == 1.0 and t[:12].sum() == 32


def test_perft_from_start_position():
    O = _orc.oracle()
    for fide in (0, 1):
        O.chess_set_fide(fide)
        s = O.new_state(CHESS, 8)
        assert [O.chess_perft(s, d) for d in (1, 2, 3)] == [20, 400, 8902]
    O.chess_set_fide(1)
    assert O.chess_perft(O.new_state(CHESS, 8), 4) == 197281        # standard perft(4); needs real pawn attacks only deeper
    O.chess_set_fide(0)
    assert O.chess_perft(O.new_state(CHESS, 8), 4) == 197281        # ... and the literal pawn test (QUIRK C5) agrees up to here


def test_reference_fen_positions():
    O = _orc.oracle()
    # castling available (chess_state_test.cpp:92-114): O-O is generated, and generated last
    s = O.chess_from_fen("r1bqkbnr/pppp1ppp/2n5/4p3/4P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 2 3")
    lg = O.legal(s)
    assert lg[-1] == code(E1, G1) and code(E1, C1) not in lg.tolist()
    assert O.state_make_move(s, code(E1, G1)) == 0
    assert O.chess_piece(s, G1) == 6 + 8 and O.chess_piece(s, 61) == 4 + 8 and O.chess_piece(s, 63) == 0 and O.state_current_player(s) == 2
    assert O.tensor(s)[13].max() == 0.5                                                     # white's rights are gone
    # en passant (:116-133): e5xf6
    s = O.chess_from_fen("rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3")
    assert code(28, 21) in O.legal(s).tolist() and O.tensor(s)[14, 2, 5] == 1.0
    assert O.state_make_move(s, code(28, 21)) == 0 and O.chess_piece(s, 29) == 0 and O.chess_piece(s, 21) == 1 + 8
    # promotion (:135-150): f7xg8 and f7xe8 with Q, R, B, N (f8 is occupied)
    s = O.chess_from_fen("rnbqkbnr/pppppPpp/8/8/8/8/PPPPPP1P/RNBQKBNR w KQkq - 0 1")
    lg = O.legal(s).tolist()
    assert lg[:8] == [code(13, 4, p) for p in (1, 2, 3, 4)] + [code(13, 6, p) for p in (1, 2, 3, 4)]
    assert O.state_make_move(s, code(13, 6, 4)) == 0 and O.chess_piece(s, 6) == 2 + 8
    # Scholar's mate set-up, black to move (:152-166): not terminal yet
    s = O.chess_from_fen("rnbqkbnr/pppp1ppp/8/4p3/2B1P3/5Q2/PPPP1PPP/RNB1K1NR b KQkq - 3 3")
    assert O.state_current_player(s) == 2 and len(O.legal(s)) > 0 and not O.state_is_terminal(s)
    # stalemate (:168-182)
    s = O.chess_from_fen("8/8/8/8/8/6k1/5q2/7K w - - 0 1")
    assert len(O.legal(s)) == 0 and O.state_is_terminal(s) and O.state_result(s) == _orc.DRAW and not O.chess_in_check(s)
    # Fool's mate (chess_integration_test.cpp:96-121)
    s = O.chess_from_fen("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 1 3")
    assert O.chess_in_check(s) and len(O.legal(s)) == 0 and O.state_is_terminal(s) and O.state_result(s) == _orc.WIN_P2


def test_fools_mate_by_moves_and_draw_rules():
    O = _orc.oracle()
    s = O.new_state(CHESS, 8)
    for a in (code(53, 45), code(12, 28), code(54, 38), code(3, 39)):      # f3 e5 g4 Qh4#
        assert O.state_make_move(s, a) == 0
    assert O.state_is_terminal(s) and O.state_result(s) == _orc.WIN_P2
    # threefold by piece placement (QUIRK: the repetition key never sees side to move / rights, chess_state.cpp:233-245)
    s = O.new_state(CHESS, 8)
    cyc = (code(62, 45), code(6, 21), code(45, 62), code(21, 6))           # Nf3 Nf6 Ng1 Ng8
    for a in cyc:
        assert not O.state_is_terminal(s)
        assert O.state_make_move(s, a) == 0
    assert not O.state_is_terminal(s) and abs(O.tensor(s)[17, 0, 0] - 2 / 3) < 1e-7
    for a in cyc:
        assert O.state_make_move(s, a) == 0
    assert O.state_is_terminal(s) and O.state_result(s) == _orc.DRAW
    # insufficient material and the fifty-move rule
    assert O.state_result(O.chess_from_fen("8/8/4k3/8/8/3K4/8/8 w - - 0 1")) == _orc.DRAW
    assert O.state_result(O.chess_from_fen("8/8/4k3/8/8/3KN3/8/8 w - - 0 1")) == _orc.DRAW
    assert O.state_result(O.chess_from_fen("8/8/4k3/8/8/3KR3/8/8 w - - 99 1")) == _orc.ONGOING
    assert O.state_result(O.chess_from_fen("8/8/4k3/8/8/3KR3/8/8 w - - 100 1")) == _orc.DRAW


def test_literal_pawn_attack_quirk_is_reproduced():
    """QUIRK C5 (chess_rules.cpp:134-148): a black king on e5 next to a white pawn on d4 — really in check — is not seen
    as attacked by the literal test, which looks one rank the other way (d6 / f6)."""
    O = _orc.oracle()
    fen = "8/8/8/4k3/3P4/8/8/4K3 b - - 0 1"
    O.chess_set_fide(0)
    assert not O.chess_in_check(O.chess_from_fen(fen))
    O.chess_set_fide(1)
    assert O.chess_in_check(O.chess_from_fen(fen))
    O.chess_set_fide(0)
    assert O.chess_in_check(O.chess_from_fen("8/8/3P4/4k3/8/8/8/4K3 b - - 0 1"))       # literal mode: the pawn "behind" attacks
