"""ctypes binding of the C ABI declared in include/az_b200.h (libaz_b200.so).

No compute happens in Python and there is no fallback: if the CUDA library is missing, or no B200-class
device is present, construction raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

GOMOKU, CHESS, GO = 0, 1, 2
ONGOING, DRAW, WIN_PLAYER1, WIN_PLAYER2 = 0, 1, 2, 3
EVAL_HASH, EVAL_RESNET, EVAL_HASH_PEAKED, EVAL_EXTERNAL = 0, 1, 2, 3
NET_FP16, NET_BF16 = 0, 1


class LibraryMissing(RuntimeError):
    pass


class EngineConfig(C.Structure):
    """az_config (include/az_b200.h) — MCTSConfig + SelfPlayManager exploration params + sizing."""
    _fields_ = [("game", C.c_int32), ("board_size", C.c_int32), ("n_slots", C.c_int32),
                ("num_simulations", C.c_int32), ("c_puct", C.c_float), ("virtual_loss", C.c_int32),
                ("evaluator", C.c_int32), ("net_blocks", C.c_int32), ("net_channels", C.c_int32),
                ("max_nodes_per_tree", C.c_int32), ("deterministic", C.c_int32),
                ("dirichlet_alpha", C.c_float), ("dirichlet_epsilon", C.c_float),
                ("init_temperature", C.c_float), ("final_temperature", C.c_float),
                ("temperature_drop_move", C.c_int32), ("auto_restart", C.c_int32),
                ("sample_ring_capacity", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
                ("n_streams", C.c_int32), ("net_precision", C.c_int32), ("tt_entries", C.c_int32), ("eval_dedup", C.c_int32), ("eval_cache_entries", C.c_int32), ("dense_policy", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("simulations", "evaluations", "terminal_leaves", "nodes_created",
                                          "nodes_expanded", "pool_overflows", "moves", "games",
                                          "samples_dropped", "kernel_launches", "waves", "eval_shared", "eval_cached")]


class Timing(C.Structure):
    _fields_ = [("waves_sampled", C.c_uint64), ("moves_sampled", C.c_uint64)] + [(n, C.c_double) for n in (
        "select_ms", "dedup_encode_ms", "evaluator_ms", "expand_backup_ms", "commit_ms", "stem_ms", "trunk_ms", "head_conv_ms", "conv1x1_gemm_ms",
        "policy_fc_ms", "value_fc_ms", "policy_value_ms")]


class SampleLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("record_bytes", "off_game_id", "off_slot", "off_ply", "off_action",
                                         "off_player", "off_z", "off_result", "off_root_value",
                                         "off_root_visits", "off_state", "state_bytes", "off_visits", "n_visits")]


EXPORTS = ["az_config_default", "az_last_error", "az_engine_create", "az_engine_destroy",
           "az_engine_load_weights", "az_engine_reset_games", "az_engine_set_root", "az_engine_search",
           "az_engine_root_stats", "az_engine_advance", "az_engine_play", "az_engine_add_dirichlet_noise", "az_engine_last_actions",
           "az_engine_slot_state", "az_engine_sample_layout", "az_engine_drain_samples",
           "az_engine_drain_samples_device", "az_engine_make_examples", "az_engine_examples_from_games", "az_engine_get_stats", "az_engine_sync", "az_engine_nn_forward",
           "az_engine_nn_bench", "az_engine_conv_bench", "az_engine_conv_sampled", "az_engine_event_record", "az_engine_event_elapsed",
           "az_rules_replay", "az_engine_set_search_params", "az_engine_node_stats", "az_device_count", "az_device_alloc", "az_device_free",
           "az_device_memcpy", "az_device_sync", "az_engine_get_timing", "az_engine_set_external_evaluator", "az_engine_set_num_simulations"]


def library_path():
    return os.path.join(HERE, "libaz_b200.so")


def build_library(verbose=False):
    """Compile every CUDA source for sm_100a into libaz_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["bash", os.path.join(HERE, "csrc", "build.sh")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libaz_b200.so failed")
    return library_path()


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not built — run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.az_last_error.restype = C.c_char_p
    vp, i32p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    sig = {
        "az_config_default": [C.POINTER(EngineConfig)],
        "az_engine_create": [C.POINTER(EngineConfig), C.POINTER(vp)],
        "az_engine_destroy": [vp],
        "az_engine_load_weights": [vp, vp, C.c_size_t],
        "az_engine_reset_games": [vp],
        "az_engine_set_root": [vp, C.c_int, i32p, C.c_int, i32p, C.c_int],
        "az_engine_search": [vp, C.c_int],
        "az_engine_root_stats": [vp, C.c_int, i32p, i32p, f32p, f32p, i32p, i32p, f32p],
        "az_engine_advance": [vp, i32p, C.c_int],
        "az_engine_play": [vp, C.c_int],
        "az_engine_add_dirichlet_noise": [vp, C.c_float, C.c_float],
        "az_engine_last_actions": [vp, i32p, C.c_int],
        "az_engine_slot_state": [vp, C.c_int, i32p, i32p, i32p],
        "az_engine_sample_layout": [vp, C.POINTER(SampleLayout)],
        "az_engine_drain_samples": [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)],
        "az_engine_drain_samples_device": [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)],
        "az_engine_make_examples": [vp, vp, C.c_size_t, C.c_int, f32p, f32p, f32p],
        "az_engine_examples_from_games": [vp, i32p, i32p, vp, C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p, f32p, f32p],
        "az_engine_get_stats": [vp, C.POINTER(Stats)],
        "az_engine_sync": [vp],
        "az_engine_nn_forward": [vp, f32p, C.c_int, f32p, f32p, f32p],
        "az_engine_nn_bench": [vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
        "az_engine_conv_bench": [vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
        "az_engine_conv_sampled": [vp, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)],
        "az_engine_event_record": [vp, C.c_int],
        "az_engine_event_elapsed": [vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
        "az_rules_replay": [vp, i32p, i32p, C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, f32p],
        "az_engine_set_search_params": [vp, C.c_float, C.c_int],
        "az_engine_set_num_simulations": [vp, C.c_int],
        "az_engine_node_stats": [vp, C.c_int, i32p, C.c_int, i32p, i32p, f32p, f32p, i32p, i32p, f32p, f32p, i32p],
        "az_device_count": [i32p],
        "az_engine_get_timing": [vp, C.POINTER(Timing)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        if name != "az_config_default":
            fn.restype = C.c_int
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def default_config(**kw):
    lib = load_library()
    cfg = EngineConfig()
    lib.az_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class Engine:
    """One engine per GPU / process.  Thin: every method is one C-ABI call."""

    def __init__(self, **kw):
        self.lib = load_library()
        self.cfg = default_config(**kw)
        h = C.c_void_p()
        self._check(self.lib.az_engine_create(C.byref(self.cfg), C.byref(h)))
        self.h = h
        self.n_slots = self.cfg.n_slots
        self.board = self.cfg.board_size if self.cfg.game != CHESS else 8
        self.cells = self.board * self.board
        # policy length / getActionSpaceSize: N*N for Gomoku, N*N + 1 for Go (go_state.cpp:345-347); a node has at most
        # `max_children` children (Go: pass + every cell)
        self.actions = self.cells + (1 if self.cfg.game == GO else 0)
        self.max_children = self.actions
        self.planes = 8 if self.cfg.game == GO else 11
        if self.cfg.game == CHESS:      # action = promo << 12 | from << 6 | to (chess_state.h:117); at most 218 legal moves
            self.actions, self.max_children, self.planes = 20480, 256, 18

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.az_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.az_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weights(self, blob, nbytes=None):
        """blob: bytes, or the address (int) of `nbytes` bytes of host memory (e.g. a pinned buffer)."""
        if isinstance(blob, int):
            self._check(self.lib.az_engine_load_weights(self.h, C.c_void_p(blob), nbytes))
        else:
            self._check(self.lib.az_engine_load_weights(self.h, blob, len(blob)))

    def reset_games(self):
        self._check(self.lib.az_engine_reset_games(self.h))

    def set_root(self, slot, moves, first_fill_order=None):
        mv = np.ascontiguousarray(moves, np.int32)
        if first_fill_order is None:
            self._check(self.lib.az_engine_set_root(self.h, slot, mv.ctypes.data, len(mv), None, 0))
        else:
            od = np.ascontiguousarray(first_fill_order, np.int32)
            self._check(self.lib.az_engine_set_root(self.h, slot, mv.ctypes.data, len(mv), od.ctypes.data, len(od)))

    def search(self, sims=0):
        self._check(self.lib.az_engine_search(self.h, sims))

    def root_stats(self, slot):
        cap = self.actions + 1
        a = np.zeros(cap, np.int32); n = np.zeros(cap, np.int32)
        w = np.zeros(cap, np.float32); p = np.zeros(cap, np.float32)
        cnt = C.c_int32(cap); rn = C.c_int32(); rw = C.c_float()
        self._check(self.lib.az_engine_root_stats(self.h, slot, a.ctypes.data, n.ctypes.data, w.ctypes.data,
                                                  p.ctypes.data, C.addressof(cnt), C.addressof(rn), C.addressof(rw)))
        k = cnt.value
        return dict(actions=a[:k].copy(), N=n[:k].copy(), W=w[:k].copy(), P=p[:k].copy(), rootN=rn.value,
                    rootW=np.float32(rw.value))

    def advance(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        self._check(self.lib.az_engine_advance(self.h, a.ctypes.data, len(a)))

    def play(self, n_moves=1):
        self._check(self.lib.az_engine_play(self.h, n_moves))

    def set_num_simulations(self, sims):
        self._check(self.lib.az_engine_set_num_simulations(self.h, int(sims)))

    def add_dirichlet_noise(self, alpha=0.03, epsilon=0.25):
        self._check(self.lib.az_engine_add_dirichlet_noise(self.h, alpha, epsilon))

    def last_actions(self):
        a = np.zeros(self.n_slots, np.int32)
        self._check(self.lib.az_engine_last_actions(self.h, a.ctypes.data, len(a)))
        return a

    def slot_state(self, slot):
        r, ply, pl = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.az_engine_slot_state(self.h, slot, C.addressof(r), C.addressof(ply), C.addressof(pl)))
        return r.value, ply.value, pl.value

    def sample_layout(self):
        lay = SampleLayout()
        self._check(self.lib.az_engine_sample_layout(self.h, C.byref(lay)))
        return lay

    def sample_dtype(self):
        L = self.sample_layout()
        return np.dtype({"names": ["game_id", "slot", "ply", "action", "player", "z", "result", "root_value",
                                   "root_visits", "state", "visits"],
                         "formats": ["<u4", "<i4", "<i2", "<i2", "i1", "i1", "i1", "<f4", "<i4",
                                     (np.uint8, L.state_bytes), ("<u2", L.n_visits)],
                         "offsets": [L.off_game_id, L.off_slot, L.off_ply, L.off_action, L.off_player, L.off_z,
                                     L.off_result, L.off_root_value, L.off_root_visits, L.off_state, L.off_visits],
                         "itemsize": L.record_bytes})

    def timing(self):
        """az_engine_get_timing: sums (ms) over the sampled waves / moves, as a dict."""
        t = Timing()
        self._check(self.lib.az_engine_get_timing(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timing._fields_}

    def drain_samples(self, cap=None, out=None):
        """Copy finished-game samples to HOST memory (pinned `out` if given) and empty the device ring."""
        dt = self.sample_dtype()
        if out is None:
            cap = cap or max(32 * self.n_slots, 4096, int(self.cfg.sample_ring_capacity))      # the whole ring (engine default: max(32 * slots, 4096))
            out = np.zeros(cap, dt)
        n = C.c_size_t()
        self._check(self.lib.az_engine_drain_samples(self.h, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def drain_samples_device(self, dev_ptr, cap_records):
        n = C.c_size_t()
        self._check(self.lib.az_engine_drain_samples_device(self.h, C.c_void_p(dev_ptr), cap_records, C.byref(n)))
        return n.value

    def make_examples(self, samples, augment=True):
        """Dataset.extractExamples(includeAugmentations) on the device: sample records -> (planes [n*k, C, N, N], policy [n*k, A],
        value [n*k]), k = 8 with augmentation (1 for chess)."""
        dt = self.sample_dtype()
        smp = np.ascontiguousarray(samples)
        if smp.dtype != dt:          # e.g. np.concatenate re-packs a padded structured dtype: restore the engine's record layout
            fixed = np.zeros(len(smp), dt)
            for name in dt.names:
                fixed[name] = smp[name]
            smp = fixed
        n = len(smp)
        k = 8 if (augment and self.cfg.game != CHESS) else 1
        planes = np.zeros((n * k, self.planes, self.board, self.board), np.float32)
        policy = np.zeros((n * k, self.actions), np.float32)
        value = np.zeros(n * k, np.float32)
        if n:
            self._check(self.lib.az_engine_make_examples(self.h, smp.ctypes.data, n, 1 if augment else 0, planes.ctypes.data,
                                                         policy.ctypes.data, value.ctypes.data))
        return planes, policy, value

    def examples_from_games(self, games, results, policies, augment=True):
        """Dataset.addGameRecord + extractExamples on the device for game records given as move lists: games = list of action lists,
        results = GameResult code per game (0 ONGOING, 1 DRAW, 2 WIN_PLAYER1, 3 WIN_PLAYER2), policies = [sum(len(g)), P] (one vector
        per recorded move, game order).  Returns (planes [n*k, C, N, N], policy [n*k, P], value [n*k])."""
        n_games = len(games)
        max_moves = max(1, max((len(g) for g in games), default=1))
        mv = np.zeros((n_games, max_moves), np.int32)
        nm = np.zeros(n_games, np.int32)
        for i, g in enumerate(games):
            mv[i, :len(g)] = g
            nm[i] = len(g)
        res = np.ascontiguousarray(results, np.int8)
        pol = np.ascontiguousarray(policies, np.float32)
        n = int(nm.sum())
        assert pol.ndim == 2 and pol.shape[0] == n and len(res) == n_games
        P = pol.shape[1]
        k = 8 if (augment and self.cfg.game != CHESS) else 1
        planes = np.zeros((n * k, self.planes, self.board, self.board), np.float32)
        policy = np.zeros((n * k, P), np.float32)
        value = np.zeros(n * k, np.float32)
        self._check(self.lib.az_engine_examples_from_games(self.h, mv.ctypes.data, nm.ctypes.data, res.ctypes.data, n_games, max_moves, pol.ctypes.data, P,
                                                           1 if augment else 0, planes.ctypes.data, policy.ctypes.data, value.ctypes.data))
        return planes, policy, value

    def stats(self):
        s = Stats()
        self._check(self.lib.az_engine_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def sync(self):
        self._check(self.lib.az_engine_sync(self.h))

    def nn_forward(self, planes, want_logits=False):
        x = np.ascontiguousarray(planes, np.float32)
        n = x.shape[0]
        pol = np.zeros((n, self.actions), np.float32); val = np.zeros(n, np.float32)
        lg = np.zeros((n, self.actions), np.float32) if want_logits else None
        self._check(self.lib.az_engine_nn_forward(self.h, x.ctypes.data, n, pol.ctypes.data, val.ctypes.data,
                                                  lg.ctypes.data if want_logits else None))
        return (pol, val, lg) if want_logits else (pol, val)

    def nn_bench(self, n_boards, reps):
        ms = C.c_float()
        self._check(self.lib.az_engine_nn_bench(self.h, n_boards, reps, C.byref(ms)))
        return ms.value

    def conv_bench(self, n_boards, reps):
        ms = C.c_float()
        self._check(self.lib.az_engine_conv_bench(self.h, n_boards, reps, C.byref(ms)))
        return ms.value

    def conv_sampled(self):
        """(accumulated ms, launches) of the 128->128 conv launches timed live inside the waves (every 64th network pass)."""
        ms = C.c_double(); n = C.c_ulonglong()
        self._check(self.lib.az_engine_conv_sampled(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def event_record(self, idx):
        self._check(self.lib.az_engine_event_record(self.h, idx))

    def event_elapsed(self, i, j):
        ms = C.c_float()
        self._check(self.lib.az_engine_event_elapsed(self.h, i, j, C.byref(ms)))
        return ms.value

    def rules_replay(self, games, want_planes=True, planes_c=None):
        """games: list of move lists.  Returns per game: legal (reference order for non-first fills), terminal,
        result, player, planes."""
        ng = len(games)
        mx = max(1, max(len(g) for g in games))
        mv = np.zeros((ng, mx), np.int32); nm = np.zeros(ng, np.int32)
        for i, g in enumerate(games):
            mv[i, :len(g)] = g; nm[i] = len(g)
        A = self.max_children
        planes_c = planes_c or self.planes
        legal = np.zeros((ng, A), np.int32); nl = np.zeros(ng, np.int32)
        term = np.zeros(ng, np.int32); res = np.zeros(ng, np.int32); pl = np.zeros(ng, np.int32)
        planes = np.zeros((ng, planes_c, self.board, self.board), np.float32) if want_planes else None
        self._check(self.lib.az_rules_replay(self.h, mv.ctypes.data, nm.ctypes.data, ng, mx, legal.ctypes.data,
                                             nl.ctypes.data, term.ctypes.data, res.ctypes.data, pl.ctypes.data,
                                             planes.ctypes.data if want_planes else None))
        return dict(legal=[legal[i, :max(nl[i], 0)].copy() for i in range(ng)], n_legal=nl, terminal=term,
                    result=res, player=pl, planes=planes)
