"""Child process of tests/test_pybind_cpu.py::test_state_string_api_matches_reference_live: drives the reference's own IGameState
implementations (oracle/_ref/libaz_ref.so) through seeded random games and prints, as JSON, what the rest of the IGameState surface returns on
the way (actionToString / stringToAction / toString / getTensorRepresentation / getMoveHistory / undoMove / validate / equals).
Runs in a process of its own and imports nothing but the standard library: the reference build carries a static libstdc++, and its iostream
code crashes when another C++ runtime (numpy's, torch's) was loaded into the process first."""
import ctypes as C
import hashlib
import json
import random
import sys

lib = C.CDLL(sys.argv[1])
lib.ref_state_new.restype = C.c_void_p
lib.ref_state_clone.restype = C.c_void_p
BUF = C.create_string_buffer(1 << 14)
ARR = (C.c_int * 32768)()
FL = (C.c_float * (32 * 19 * 19))()


def s_of(fn, *a):
    n = fn(*a, BUF, len(BUF))
    return None if n < 0 else BUF.value.decode()


def s2a(r, text):
    act = C.c_int(0)
    rc = lib.ref_state_string_to_action(r, text.encode(), C.byref(act))
    return act.value if rc == 0 else (None if rc == 1 else "THROW")


out = []
rng = random.Random(11)
BAD = ["", "Z99", "I5", "a1", "h8", "H8", "e2e4", "e7e8q", "e2e5", "pass", "PASS", "A0", "A16", "e2", "xx", "a2a4q", "H 8", "p1", "T19", "j10", "J10"]
for game, board, plies in ((0, 15, 60), (0, 9, 40), (2, 9, 70), (2, 19, 60), (1, 8, 90)):
    for g in range(2):
        r = C.c_void_p(lib.ref_state_new(game, board))
        rec = {"game": game, "board": board, "steps": [], "bad": {t: s2a(r, t) for t in BAD},
               "bad_a2s": {str(a): s_of(lib.ref_state_action_to_string, r, a) for a in (-1, -2, board * board, board * board + 5, 99999)}}
        for ply in range(plies):
            n = lib.ref_state_legal_moves(r, ARR, len(ARR))
            lm = list(ARR[:n])
            if lib.ref_state_is_terminal(r) or not lm:
                break
            pick = rng.sample(lm, min(4, len(lm)))
            strs = [s_of(lib.ref_state_action_to_string, r, a) for a in pick]
            c = lib.ref_state_basic_tensor(r, None)
            lib.ref_state_basic_tensor(r, FL)
            step = {"pick": pick, "a2s": strs, "s2a": [s2a(r, t) for t in strs], "to_string": s_of(lib.ref_state_to_string, r),
                    "basic_c": c, "basic_sha": hashlib.sha256(bytes(memoryview(FL))[:4 * c * board * board]).hexdigest()[:16],
                    "validate": lib.ref_state_validate(r)}
            a = rng.choice(lm)
            assert lib.ref_state_make_move(r, a) == 0
            step["move"] = a
            h = lib.ref_state_history(r, ARR, len(ARR))
            step["history_len"] = h
            step["history_tail"] = list(ARR[:h])[-3:]
            if ply % 7 == 3 and game == 1:
                # chess: the reference's undoMove does not restore the position (legal moves differ afterwards and the undone move is refused):
                # recorded on a clone, the game itself goes on
                cl = C.c_void_p(lib.ref_state_clone(r))
                step["chess_undo"] = lib.ref_state_undo(cl)
                n2 = lib.ref_state_legal_moves(cl, ARR, len(ARR))
                step["chess_legal_after_undo"] = n2
                step["chess_remake"] = lib.ref_state_make_move(cl, a)
                lib.ref_state_free(cl)
            elif ply % 7 == 3:
                cl = C.c_void_p(lib.ref_state_clone(r))
                step["equals_clone"] = lib.ref_state_equals(r, cl)
                step["undo"] = lib.ref_state_undo(r)
                step["equals_after_undo"] = lib.ref_state_equals(r, cl)
                n2 = lib.ref_state_legal_moves(r, ARR, len(ARR))
                step["legal_after_undo"] = list(ARR[:n2])
                step["player_after_undo"] = lib.ref_state_current_player(r)
                step["remake"] = lib.ref_state_make_move(r, a)          # playing the undone move again
                lib.ref_state_free(cl)
                if step["remake"] != 0:
                    rec["steps"].append(step)
                    break
            rec["steps"].append(step)
        if game == 2 and not lib.ref_state_is_terminal(r):      # two passes: the printout of a finished game (scores, winner)
            assert lib.ref_state_make_move(r, -1) == 0 and lib.ref_state_make_move(r, -1) == 0 and lib.ref_state_is_terminal(r)
            rec["final_to_string"] = s_of(lib.ref_state_to_string, r)
            c = lib.ref_state_basic_tensor(r, None)
        r0 = C.c_void_p(lib.ref_state_new(game, board))
        rec["undo_on_fresh"] = lib.ref_state_undo(r0)
        out.append(rec)
json.dump(out, sys.stdout)
