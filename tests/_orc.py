"""ctypes front-end shared by the tests for the two CPU checkers:

* ``oracle/libaz_oracle.so``      — the CPU restatement (prefix ``orc_``), built by ``make -C oracle``
* ``oracle/_ref/libaz_ref.so``    — the patched reference itself (prefix ``ref_``), built by
  ``oracle/build_ref.sh`` in the authoring container (it travels to the GPU box prebuilt).

Both export the same C entry points, so one wrapper serves both.  TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOMOKU, CHESS, GO = 0, 1, 2
ONGOING, DRAW, WIN_P1, WIN_P2 = 0, 1, 2, 3

EVAL_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


class Checker:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        f = self._f
        f("state_new", C.c_void_p, [C.c_int, C.c_int])
        f("state_free", None, [C.c_void_p])
        f("state_clone", C.c_void_p, [C.c_void_p])
        f("state_make_move", C.c_int, [C.c_void_p, C.c_int])
        f("state_legal_moves", C.c_int, [C.c_void_p, C.c_void_p, C.c_int])
        f("state_is_terminal", C.c_int, [C.c_void_p])
        f("state_result", C.c_int, [C.c_void_p])
        f("state_current_player", C.c_int, [C.c_void_p])
        f("state_action_space", C.c_int, [C.c_void_p])
        f("state_board_size", C.c_int, [C.c_void_p])
        f("state_tensor", C.c_int, [C.c_void_p, C.c_void_p])
        f("state_key", C.c_uint64, [C.c_void_p])
        f("hash_eval", None, [C.c_void_p, C.c_void_p, C.c_void_p])
        # chess: the restatement, and the reference itself with its legality recursion cut (oracle/build_ref.sh shim 6)
        f("chess_set_fen", C.c_int, [C.c_void_p, C.c_char_p])
        f("chess_set_fide", None, [C.c_int])
        f("chess_perft", C.c_long, [C.c_void_p, C.c_int])
        f("chess_piece", C.c_int, [C.c_void_p, C.c_int])
        f("chess_in_check", C.c_int, [C.c_void_p])
        if prefix == "ref_":     # the reference's Dataset / GameRecord (src/selfplay/dataset.cpp, game_record.cpp)
            f("augment_example", None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p])
            f("dataset_extract", C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
            f("game_record_json", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int])
        f("go_stone", C.c_int, [C.c_void_p, C.c_int])
        f("go_ko", C.c_int, [C.c_void_p])
        f("go_captured", C.c_int, [C.c_void_p, C.c_int])
        f("mcts_new", C.c_void_p, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p])
        f("mcts_free", None, [C.c_void_p])
        f("mcts_search", None, [C.c_void_p])
        f("mcts_set_sims", None, [C.c_void_p, C.c_int])
        f("mcts_eval_calls", C.c_long, [C.c_void_p])
        f("mcts_root_stats", C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p])
        f("mcts_select_action", C.c_int, [C.c_void_p, C.c_int, C.c_float])
        f("mcts_action_probs", C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_int])
        f("mcts_root_value", C.c_float, [C.c_void_p])
        f("mcts_update_with_move", None, [C.c_void_p, C.c_int])
        if prefix == "orc_":
            f("mcts_set_tt", None, [C.c_void_p, C.c_int])
            f("mcts_tt_hits", C.c_long, [C.c_void_p])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, name, fn)

    # ---- convenience -------------------------------------------------------------------------
    def new_state(self, game, n):
        h = self.state_new(game, n)
        assert h, "state_new failed"
        return h

    def legal(self, s):
        buf = np.zeros(512, np.int32)
        n = self.state_legal_moves(s, buf.ctypes.data, 512)
        assert n <= 512
        return buf[:n].copy()

    def chess_from_fen(self, fen):
        s = self.new_state(CHESS, 8)
        assert self.chess_set_fen(s, fen.encode()) == 0
        return s

    def tensor(self, s):
        c = self.state_tensor(s, None)
        n = self.state_board_size(s)
        out = np.zeros((c, n, n), np.float32)
        self.state_tensor(s, out.ctypes.data)
        return out

    def hash_policy_value(self, s):
        a = self.state_action_space(s)
        pol = np.zeros(a, np.float32)
        v = C.c_float()
        self.hash_eval(s, pol.ctypes.data, C.byref(v))
        return pol, np.float32(v.value)

    def root_stats(self, m):
        a = np.zeros(512, np.int32); n_ = np.zeros(512, np.int32)
        w = np.zeros(512, np.float32); p = np.zeros(512, np.float32)
        rn = C.c_int(); rw = C.c_float()
        n = self.mcts_root_stats(m, a.ctypes.data, n_.ctypes.data, w.ctypes.data, p.ctypes.data, 512,
                                 C.byref(rn), C.byref(rw))
        return dict(actions=a[:n].copy(), N=n_[:n].copy(), W=w[:n].copy(), P=p[:n].copy(),
                    rootN=rn.value, rootW=np.float32(rw.value))

    def probs(self, m, temperature):
        out = np.zeros(512, np.float32)
        n = self.mcts_action_probs(m, temperature, out.ctypes.data, 512)
        return out[:n].copy()


_cache = {}


def oracle():
    """The CPU restatement; (re)built on demand with gcc — it is a few hundred lines."""
    if "orc" not in _cache:
        so = os.path.join(ROOT, "oracle", "libaz_oracle.so")
        src = os.path.join(ROOT, "oracle", "az_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libaz_oracle.so"],
                                  stdout=subprocess.DEVNULL)
        _cache["orc"] = Checker(so, "orc_")
        ag = _cache["orc"].lib.orc_augment_example
        ag.restype = None
        ag.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        fn = _cache["orc"].lib.orc_first_fill_order
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _cache["orc"]


def ref_path():
    return os.path.join(ROOT, "oracle", "_ref", "libaz_ref.so")


def have_ref():
    return os.path.exists(ref_path())


def reference():
    """The patched reference itself (None when oracle/_ref was never built)."""
    if "ref" not in _cache:
        _cache["ref"] = Checker(ref_path(), "ref_") if have_ref() else None
    return _cache["ref"]


def augment_example(planes, policy):
    """Dataset::augmentExample restatement: (planes [C,N,N], policy [A]) -> (planes [7,C,N,N], policy [7,A])."""
    pl = np.ascontiguousarray(planes, np.float32); po = np.ascontiguousarray(policy, np.float32)
    c, n, _ = pl.shape
    opl = np.zeros((7, c, n, n), np.float32); opo = np.zeros((7, len(po)), np.float32)
    oracle().lib.orc_augment_example(pl.ctypes.data, c, n, po.ctypes.data, len(po), opl.ctypes.data, opo.ctypes.data)
    return opl, opo


def first_fill_order(empties):
    e = np.asarray(empties, np.int32)
    out = np.zeros(len(e), np.int32)
    oracle().lib.orc_first_fill_order(e.ctypes.data, len(e), out.ctypes.data)
    return out
