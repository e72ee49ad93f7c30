"""CPU suite: SURVEY 8f.4 — the variant rules behind createGameState(type, boardSize, variantRules=True) (Renju / Omok, Chess960).

Probing the reference itself (oracle/_ref, its own sources compiled by oracle/build_ref.sh) shows that none of them survives its first
use at the reference's HEAD, so there is no behaviour to be bit-exact with; the drop-in raises an error where the reference dies:

* Renju: `GomokuRules::renju_double_four_or_more` (src/games/gomoku/gomoku_rules.cpp:198-220) replaces the rules' board accessor `is_bit_set`
  by a lambda that itself calls `this->is_bit_set` — unbounded recursion, stack overflow (SIGSEGV) on the first forbidden-move test, i.e. on
  Black's first getLegalMoves() / isLegalMove() (the reference's own tests/games/gomoku/gomoku_state_test.cpp:126-139 would crash there).
* Omok: the same pattern in `omok_check_double_three_strict` (:360-397).
* Chess960: ChessState(chess960=true) uses position number 518; `Chess960::getPermutation` (src/games/chess/chess960.cpp:466-480) maps
  knight configuration 5 to the same index twice, one square stays empty: assert failure (debug) / "Invalid piece index in Chess960
  generation" (release).  192 of the 960 position numbers are affected; the other 768 are reachable only through the C++ constructor.

Each probe runs in a child process (the crash is the result)."""
import os
import signal
import subprocess
import sys

import pytest

import _orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alphazero-multi-game_b200")
needs_ref = pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref/libaz_ref.so not built")

PROBE = r"""
import ctypes, sys
L = ctypes.CDLL(sys.argv[1])
L.ref_state_new_variant.restype = ctypes.c_void_p
L.ref_state_new_variant.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
L.ref_state_legal_moves.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
L.ref_state_is_legal.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.ref_state_make_move.argtypes = [ctypes.c_void_p, ctypes.c_int]
what = sys.argv[2]
if what == "chess960":
    s = L.ref_state_new_variant(1, 0, 1)
    print("constructed", flush=True)
elif what == "chess960_n":
    print("constructible", L.ref_chess960_constructible(int(sys.argv[3])), flush=True)
else:
    variant = 1 if what.startswith("renju") else 2 if what.startswith("omok") else 0
    s = L.ref_state_new_variant(0, 15, variant)
    print("constructed", flush=True)
    if what.endswith("_white"):               # White is never tested for forbidden moves: the crash needs Black to move
        buf = (ctypes.c_int * 256)()
        # standard rules state is used to get past Black's move: not possible with the variant state itself
        sys.exit(0)
    if what.endswith("_islegal"):
        print("legal", L.ref_state_is_legal(s, 112), flush=True)
    else:
        buf = (ctypes.c_int * 256)()
        print("n", L.ref_state_legal_moves(s, buf, 256), flush=True)
"""


def probe(*args):
    r = subprocess.run([sys.executable, "-c", PROBE, _orc.ref_path(), *map(str, args)], capture_output=True, text=True, timeout=120)
    return r.returncode, r.stdout


@needs_ref
@pytest.mark.parametrize("what", ["renju", "renju_islegal", "omok", "omok_islegal"])
def test_reference_renju_and_omok_overflow_the_stack_on_blacks_first_move(what):
    rc, out = probe(what)
    assert "constructed" in out                       # the constructor is fine
    assert rc == -signal.SIGSEGV, (rc, out)           # ... the first forbidden-move test is not
    assert "n " not in out and "legal " not in out


@needs_ref
def test_reference_standard_gomoku_probe_is_alive():
    rc, out = probe("plain")
    assert rc == 0 and "n 225" in out


@needs_ref
def test_reference_chess960_default_position_cannot_be_constructed():
    rc, out = probe("chess960")
    assert rc == -signal.SIGABRT and "constructed" not in out, (rc, out)      # assert in Chess960::getPermutation (the oracle build keeps asserts)
    rc, out = probe("chess960_n", 518)
    assert rc == -signal.SIGABRT
    rc, out = probe("chess960_n", 96)                 # a position number whose knights land on two squares
    assert rc == 0 and "constructible 1" in out


def test_host_mirror_raises_where_the_reference_dies():
    sys.path.insert(0, PKG)
    import _alphazero_cpp as az
    with pytest.raises(RuntimeError, match="Renju"):
        az.createGameState(az.GameType.GOMOKU, 15, True)
    with pytest.raises(RuntimeError, match="Chess960"):
        az.createGameState(az.GameType.CHESS, 0, True)
    with pytest.raises(RuntimeError):
        az.GomokuState(15, True)
    with pytest.raises(RuntimeError):
        az.GomokuState(15, False, True)
    g = az.createGameState(az.GameType.GO, 9, True)   # Go ignores the flag (game_factory.cpp:106-111)
    assert g.getBoardSize() == 9 and len(g.getLegalMoves()) == 82
