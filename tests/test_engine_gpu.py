"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the
oracle (CPU restatement, pinned to the reference) and the golden fixtures generated from the reference itself.

Bar (BASELINE.json north_star): legal moves, terminal flags and per-root visit counts bit-exact; Q within 1e-6
(here W and P are compared as IEEE bit patterns, which is stronger)."""
import json
import os

import numpy as np
import pytest

import _orc
from _orc import GOMOKU

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def test_rules_kernels_match_golden_and_oracle():
    """Device rules (apply / win / draw / legal order / 11-plane encoder) on the golden playouts."""
    from _eng import hash_engine
    O = _orc.oracle()
    for board in (15, 9):
        eng = hash_engine(4, board=board, sims=8)
        games, expect = [], []
        for case in json.load(open(os.path.join(GOLD, "state_playouts.json"))):
            if case["game"] != GOMOKU or case["board"] != board:
                continue
            s = O.new_state(GOMOKU, board)
            for ply in range(len(case["moves"]) + 1):
                games.append(case["moves"][:ply])
                expect.append((np.sort(O.legal(s))[::-1].copy(), O.state_is_terminal(s), O.state_result(s),
                               O.state_current_player(s), O.tensor(s)))
                if ply < len(case["moves"]):
                    O.state_make_move(s, case["moves"][ply])
        assert len(games) > 100
        r = eng.rules_replay(games)
        for i, (legal, term, res, pl, planes) in enumerate(expect):
            assert np.array_equal(r["legal"][i], legal), (board, i)
            assert r["terminal"][i] == term and r["result"][i] == res and r["player"][i] == pl, (board, i)
            assert np.array_equal(r["planes"][i], planes), (board, i)
        # illegal replays are reported, not applied (reference make_move throws, gomoku_state.cpp:681-689)
        bad = eng.rules_replay([[0, 0], [board * board]])
        assert bad["n_legal"][0] == -1 and bad["n_legal"][1] == -1
        eng.close()


@pytest.mark.parametrize("idx,cache", [(0, 0), (2, 0), (2, 1 << 12)])
def test_search_matches_reference_golden(idx, cache):
    """Serial-search parity on the golden cases generated from the patched reference (Gomoku 15x15 @800 sims,
    9x9 @200 sims): child order, visit counts, valueSum and prior bits, root leak, chosen move.  cache > 0: with the evaluation cache
    (M16) on behind the hash evaluator, deliberately small (4096 entries per engine: evictions happen) — still bit-equal."""
    from _eng import hash_engine
    case = json.load(open(os.path.join(GOLD, "search_hash_eval.json")))[idx]
    assert case["game"] == GOMOKU
    eng = hash_engine(3, board=case["board"], sims=case["sims"], eval_cache_entries=cache)
    for mv, g in enumerate(case["moves"]):
        eng.search()
        for slot in range(3):                      # identical games in every slot
            st = eng.root_stats(slot)
            assert st["actions"].tolist() == g["actions"], (mv, slot)
            assert st["N"].tolist() == g["N"], (mv, slot)
            assert bits(st["W"]).tolist() == g["W"], (mv, slot)
            assert bits(st["P"]).tolist() == g["P"], (mv, slot)
            assert st["rootN"] == g["rootN"] and int(bits([st["rootW"]])[0]) == g["rootW"]
        eng.advance([g["action"]] * 3)
        if eng.slot_state(0)[0] != 0:
            break
    st = eng.stats()
    assert st["pool_overflows"] == 0
    eng.close()


def test_search_matches_oracle_random_openings():
    """Different position in every slot, searched together wave by wave, each compared with the oracle's serial
    search: covers enumerated roots (descending order) and never-enumerated roots (first-fill order, QUIRK G2)."""
    from _eng import hash_engine
    O = _orc.oracle()
    rng = np.random.default_rng(11)
    T, sims, board = 12, 300, 15
    eng = hash_engine(T, board=board, sims=sims)
    states, searches = [], []
    for t in range(T):
        s = O.new_state(GOMOKU, board)
        nply = int(rng.integers(0, 40))
        moves = rng.choice(board * board, nply, replace=False).tolist()
        enumerated = (t % 2 == 0)
        for a in moves:
            if enumerated:
                O.legal(s)
            assert O.state_make_move(s, int(a)) == 0
        if O.state_is_terminal(O.state_clone(s)):       # (clone: do not disturb s's cache state)
            moves = []; s = O.new_state(GOMOKU, board); enumerated = False
        if enumerated:
            O.legal(s)
            eng.set_root(t, moves)
        else:
            empties = [a for a in range(board * board) if a not in set(moves)]
            eng.set_root(t, moves, first_fill_order=_orc.first_fill_order(empties))
        states.append(s)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 0, None, None))
    for mv in range(4):
        eng.search()
        acts = []
        for t in range(T):
            O.mcts_search(searches[t])
            a, b = eng.root_stats(t), O.root_stats(searches[t])
            assert np.array_equal(a["actions"], b["actions"]), (mv, t)
            assert np.array_equal(a["N"], b["N"]), (mv, t)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (mv, t)
            assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0]
            q_e = a["W"][a["N"] > 0] / a["N"][a["N"] > 0]
            q_o = b["W"][b["N"] > 0] / b["N"][b["N"] > 0]
            assert np.max(np.abs(q_e - q_o), initial=0.0) <= 1e-6          # the stated Q tolerance
            act = O.mcts_select_action(searches[t], 1, 1.0)
            acts.append(act)
            O.mcts_update_with_move(searches[t], act)
        eng.advance(acts)
    assert eng.stats()["pool_overflows"] == 0
    eng.close()


def test_selfplay_loop_matches_oracle_full_game():
    """az_engine_play (search → choose → record → re-root → turnover) in deterministic mode against the oracle's
    playSingleGame restatement, a whole 9x9 game: same moves, same recorded visit counts, same result, and the
    drained sample records carry the right z."""
    from _eng import hash_engine
    O = _orc.oracle()
    board, sims = 9, 120
    eng = hash_engine(2, board=board, sims=sims)
    s = O.new_state(GOMOKU, board)
    m = O.mcts_new(s, sims, 1.5, 3, 0, None, None)
    moves, visit_rows, rvals = [], [], []
    while not O.state_is_terminal(s):
        O.mcts_search(m)
        st = O.root_stats(m)
        v = np.zeros(board * board, np.int64); v[st["actions"]] = st["N"]
        a = O.mcts_select_action(m, 1, 1.0)
        visit_rows.append(v); moves.append(a); rvals.append(O.mcts_root_value(m))
        O.state_make_move(s, a); O.mcts_update_with_move(m, a)
    result = O.state_result(s)
    for i, a in enumerate(moves):
        eng.play(1)
        assert eng.last_actions().tolist() == [a, a], i
    assert eng.slot_state(0)[0] == result
    smp = eng.drain_samples()
    assert len(smp) == 2 * len(moves)
    s0 = np.sort(smp[smp["slot"] == 0], order="ply")
    assert s0["action"].tolist() == moves
    for i in range(len(moves)):
        assert np.array_equal(s0["visits"][i][:board * board].astype(np.int64), visit_rows[i]), i
        assert bits([s0["root_value"][i]])[0] == bits([rvals[i]])[0]
        pl = 1 + (i % 2)
        z = 0 if result == 1 else (1 if (result == 2) == (pl == 1) else -1)
        assert s0["z"][i] == z and s0["result"][i] == result and s0["player"][i] == pl
    st = eng.stats()
    assert st["games"] == 2 and st["moves"] == 2 * len(moves) and st["pool_overflows"] == 0
    eng.close()


def test_selfplay_auto_restart_and_noise_smoke():
    """Throughput-mode loop (Dirichlet noise + temperature sampling + auto-restart): invariants only —
    visit counts sum to sims (+ reused subtree), games finish, slots restart, samples are well-formed."""
    from _eng import hash_engine
    board, sims, T = 9, 64, 64
    eng = hash_engine(T, board=board, sims=sims, deterministic=0, auto_restart=1, seed=7)
    total = 0
    for _ in range(30):
        eng.play(3)
        smp = eng.drain_samples()
        total += len(smp)
        if len(smp):
            assert np.all(smp["visits"][:, :board * board].sum(1) >= sims - 1)
            assert np.all(np.abs(smp["z"]) <= 1) and np.all(smp["result"] >= 1)
            played = smp["visits"][np.arange(len(smp)), smp["action"]]
            assert np.all(played >= 1)
    st = eng.stats()
    assert st["games"] >= 1 and total >= 1 and st["pool_overflows"] == 0 and st["samples_dropped"] == 0
    assert st["moves"] == 90 * T
    eng.close()


def test_full_size_4096_slots_identical_and_golden():
    """BASELINE.json configs[1] at full size (4096 slots, 800 simulations) with the hash evaluator in deterministic mode:
    size-independent property — every slot plays the same game, so all 4096 trees must be bit-identical to each other and
    slot 0 must equal the golden search generated from the reference (Gomoku 15x15 @800)."""
    from _eng import hash_engine
    case = json.load(open(os.path.join(GOLD, "search_hash_eval.json")))[0]
    assert case["game"] == GOMOKU and case["board"] == 15 and case["sims"] == 800
    T = 4096
    eng = hash_engine(T, board=15, sims=800, n_streams=1)
    for mv, g in enumerate(case["moves"][:3]):
        eng.search()
        for slot in (0, 1, 17, 2047, 2048, 4095):
            st = eng.root_stats(slot)
            assert st["actions"].tolist() == g["actions"] and st["N"].tolist() == g["N"], (mv, slot)
            assert bits(st["W"]).tolist() == g["W"] and bits(st["P"]).tolist() == g["P"], (mv, slot)
            assert st["rootN"] == g["rootN"] and int(bits([st["rootW"]])[0]) == g["rootW"]
        eng.play(0)
        eng.advance([g["action"]] * T)
    st = eng.stats()
    assert st["simulations"] == 3 * 800 * T and st["pool_overflows"] == 0 and st["moves"] == 3 * T
    eng.close()


def test_full_size_resnet_batch_position_independence():
    """4096 slots with the bf16 ResNet evaluator in deterministic mode: every slot holds the same position, so the network
    must give every board of the batch the same bits wherever it sits in the 4096-board launch (tile position, CTA pair,
    accumulator stage), and the searches built on it must be identical across slots — at the BASELINE batch size."""
    import torch
    from _eng import E, N
    T, sims = 4096, 24
    model = N.make_random_model(seed=3, blocks=10)
    with torch.no_grad():
        model.p_fc.weight *= 0.03; model.v_fc1.weight *= 0.02
    blob = N.export_weights(model)
    first = None
    # eval_dedup = -1: all 4096 boards really go through the network (with the sharing on, 4095 of them would ride on slot 0's evaluation);
    # then the default engine (in-wave sharing + evaluation cache): one evaluation per wave, the same bits in every slot
    for dedup in (-1, 0):
        eng = E.Engine(game=E.GOMOKU, board_size=15, n_slots=T, evaluator=E.EVAL_RESNET, net_blocks=10, num_simulations=sims,
                       deterministic=1, auto_restart=0, max_nodes_per_tree=2 * (sims + 2) * 225 + 1, eval_dedup=dedup)
        eng.load_weights(blob)
        seen = []
        for mv in range(2):
            eng.search()
            ref = eng.root_stats(0)
            assert int(ref["N"].sum()) == sims * (mv + 1) or int(ref["N"].sum()) >= sims
            for slot in (1, 255, 256, 1000, 2049, 4094, 4095):
                st = eng.root_stats(slot)
                assert np.array_equal(st["actions"], ref["actions"]) and np.array_equal(st["N"], ref["N"]), (dedup, mv, slot)
                assert np.array_equal(bits(st["W"]), bits(ref["W"])) and np.array_equal(bits(st["P"]), bits(ref["P"])), (dedup, mv, slot)
            seen.append(ref)
            eng.play(0)
            a = int(ref["actions"][int(np.argmax(ref["N"]))])
            eng.advance([a] * T)
        st = eng.stats()
        assert st["pool_overflows"] == 0
        assert (st["eval_shared"] == 0) if dedup < 0 else (st["eval_shared"] + st["eval_cached"] >= st["evaluations"] * 4090 // 4096)
        eng.close()
        if first is None:
            first = seen
        else:
            for a, b in zip(first, seen):
                assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"])
                assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"]))


def test_c_abi_error_behaviour():
    """Status codes + az_last_error where the reference throws (igamestate.h:36-52, std::runtime_error): unsupported game /
    board, evaluator mismatch, bad weight blobs, illegal roots, out-of-range slots; the engine stays usable afterwards."""
    import struct
    from _eng import E, N, hash_engine
    with pytest.raises(E.EngineError, match="unsupported game"):
        E.Engine(game=E.GOMOKU, board_size=7, n_slots=2, evaluator=E.EVAL_HASH)
    with pytest.raises(E.EngineError, match="n_slots"):
        E.Engine(game=E.GOMOKU, board_size=9, n_slots=0, evaluator=E.EVAL_HASH)
    eng = hash_engine(2, board=9, sims=8)
    with pytest.raises(E.EngineError, match="hash evaluator"):
        eng.load_weights(b"AZW1" + b"\0" * 64)
    with pytest.raises(E.EngineError, match="illegal move"):
        eng.set_root(0, [40, 40])
    with pytest.raises(E.EngineError, match="illegal move"):
        eng.set_root(0, [81])
    with pytest.raises(E.EngineError, match="slot out of range"):
        eng.set_root(2, [])
    with pytest.raises(E.EngineError, match="slot out of range"):
        eng.root_stats(-1)
    with pytest.raises(E.EngineError, match="one action per slot"):
        eng.advance([0])
    eng.set_root(0, [40]); eng.set_root(1, [40])
    eng.search()
    st = eng.root_stats(0)
    assert int(st["N"].sum()) == 8 and 40 not in st["actions"].tolist()
    eng.advance([-2, int(st["actions"][0])])                        # -2 = leave the slot alone
    assert eng.slot_state(0)[1] == 1 and eng.slot_state(1)[1] == 2
    eng.close()
    net = E.Engine(game=E.GOMOKU, board_size=9, n_slots=4, evaluator=E.EVAL_RESNET, net_blocks=1, num_simulations=4, max_nodes_per_tree=2048)
    with pytest.raises(E.EngineError, match="no network weights"):
        net.search()
    with pytest.raises(E.EngineError, match="magic"):
        net.load_weights(b"XXXX" + b"\0" * 64)
    good = N.export_weights(N.make_random_model(seed=0, blocks=1, board=9, actions=81))
    with pytest.raises(E.EngineError, match="truncated"):
        net.load_weights(good[:len(good) // 2])
    wrong = N.export_weights(N.make_random_model(seed=0, blocks=1, board=15, actions=225))
    with pytest.raises(E.EngineError, match="does not match"):
        net.load_weights(wrong)
    net.load_weights(good)
    net.play(1)
    assert net.stats()["moves"] == 4
    net.close()


def test_dirichlet_noise_schedule_matches_playSingleGame():
    """SelfPlayManager::playSingleGame (self_play_manager.cpp:184, 209-211): addDirichletNoise once before the first search and again after
    every EVEN move number — so the searches at plies 0, 1, 3, 5, ... see noisy root priors and those at plies 2, 4, 6, ... the pristine
    ones (normalised network policy, also when the root is a reused subtree).  The noise values themselves are unpinned (libstdc++
    gamma_distribution vs Philox); the schedule is deterministic and checked here: pristine priors are compared bit for bit with the
    oracle's expansion arithmetic."""
    from _eng import hash_engine
    O = _orc.oracle()
    board, sims, T = 9, 60, 3
    eng = hash_engine(T, board=board, sims=sims, deterministic=0, auto_restart=0, init_temperature=0.0, final_temperature=0.0, seed=3, n_streams=1)
    states = [O.new_state(GOMOKU, board) for _ in range(T)]
    for ply in range(8):
        eng.search()
        acts = []
        for t in range(T):
            st = eng.root_stats(t)
            pol, _ = O.hash_policy_value(states[t])
            raw = pol[st["actions"]]
            tot = np.float32(0.0)
            for x in raw:
                tot = np.float32(tot + x)
            pristine = np.array_equal((raw / tot).astype(np.float32).view(np.uint32), st["P"].view(np.uint32))
            noisy_ply = ply == 0 or (ply - 1) % 2 == 0
            assert pristine == (not noisy_ply), (ply, t)
            assert abs(float(st["P"].sum()) - 1.0) < 1e-4
            a = int(st["actions"][int(np.argmax(st["N"]))])
            assert O.state_make_move(states[t], a) == 0
            acts.append(a)
        eng.advance(acts)
    eng.close()


def test_advisor_regressions_groups_order_drain():
    """(1) stream-group counts that do not divide the slots (9 / 4 left an empty group: invalid launch; 5 / 4 a negative one); (2) a caller's
    first-fill order is validated; (3) a drain buffer smaller than the ring keeps the rest for the next call."""
    from _eng import E, hash_engine
    for n_slots, n_streams in [(9, 4), (5, 4), (7, 3)]:
        eng = hash_engine(n_slots, board=9, sims=12, n_streams=n_streams)
        eng.search()
        assert all(int(eng.root_stats(t)["N"].sum()) == 12 for t in range(n_slots))
        eng.close()
    eng = hash_engine(2, board=9, sims=8)
    with pytest.raises(E.EngineError, match="first_fill_order"):
        eng.set_root(0, [40], first_fill_order=list(range(81)))             # 40 is occupied
    with pytest.raises(E.EngineError, match="first_fill_order"):
        eng.set_root(0, [40], first_fill_order=[0] * 80)                    # repeated
    with pytest.raises(E.EngineError, match="first_fill_order"):
        eng.set_root(0, [40], first_fill_order=list(range(40)))             # incomplete
    eng.set_root(0, [40], first_fill_order=[a for a in range(81) if a != 40])
    eng.close()
    eng = hash_engine(16, board=9, sims=8, deterministic=0, auto_restart=1, seed=11)
    got = 0
    for _ in range(60):
        eng.play(2)
    st = eng.stats()
    first = eng.drain_samples(cap=5)
    rest = eng.drain_samples()
    assert len(first) == 5 and st["samples_dropped"] == 0
    assert len(first) + len(rest) >= st["games"] and len(rest) > 0         # nothing was thrown away by the short drain
    assert not set(zip(first["game_id"].tolist(), first["slot"].tolist(), first["ply"].tolist())) & set(zip(rest["game_id"].tolist(), rest["slot"].tolist(), rest["ply"].tolist()))
    eng.close()


@pytest.mark.parametrize("game,board", [(GOMOKU, 15), (_orc.GO, 9), (_orc.CHESS, 8)])
def test_in_wave_eval_dedup_is_result_transparent(game, board):
    """In-wave evaluation dedup (the TranspositionTable's role, M16, inside one wave): trees whose leaves present the same network input in
    the same wave share ONE ResNet evaluation.  The network output of a board does not depend on its batch position, so the searches are
    bit-identical with the dedup on and off — same child order, visit counts, valueSum and prior bits on every slot over several moves
    (slots 0-5 play the same game, so most of their leaves are shared; the others start from different openings)."""
    from _eng import E, N
    O = _orc.oracle()
    planes, actions = {GOMOKU: (11, board * board), _orc.GO: (8, board * board + 1), _orc.CHESS: (18, 20480)}[game]
    m = N.make_random_model(seed=4, randomize_bn=True, blocks=2, in_planes=planes, board=board, actions=actions)
    blob = N.export_weights(m)
    rng = np.random.default_rng(3)
    openings = [[] for _ in range(6)]
    for _ in range(6):
        s = O.new_state(game, board); mv = []
        for _ in range(int(rng.integers(1, 12))):
            lg = O.legal(s)
            a = int(rng.choice(lg[lg >= 0] if game == _orc.GO else lg))
            O.state_make_move(s, a); mv.append(a)
        openings.append(mv)
    out = []
    for dedup in (0, -1):
        eng = E.Engine(game=game, board_size=board, n_slots=len(openings), evaluator=E.EVAL_RESNET, net_blocks=2, num_simulations=48, deterministic=1,
                       auto_restart=0, eval_dedup=dedup, eval_cache_entries=-1)
        eng.load_weights(blob)
        for t, mv in enumerate(openings):
            eng.set_root(t, mv)
        res = []
        for move in range(3):
            eng.search()
            res.append([eng.root_stats(t) for t in range(len(openings))])
            eng.advance([int(r["actions"][int(np.argmax(r["N"]))]) if len(r["N"]) else -2 for r in res[-1]])      # (-2: the game in this slot is over — Go: two passes)
        st = eng.stats()
        out.append((res, st))
        eng.close()
    (on, st_on), (off, st_off) = out
    assert st_on["eval_shared"] > 0 and st_off["eval_shared"] == 0 and st_on["evaluations"] == st_off["evaluations"]
    assert st_on["eval_shared"] >= st_on["evaluations"] // 3                    # six identical games: at least 5/12 of the leaves are shared
    for move in range(3):
        for t in range(len(openings)):
            a, b = on[move][t], off[move][t]
            assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (move, t)
            assert np.array_equal(a["W"].view(np.uint32), b["W"].view(np.uint32)) and np.array_equal(a["P"].view(np.uint32), b["P"].view(np.uint32)), (move, t)


@pytest.mark.parametrize("game,board", [(GOMOKU, 15), (_orc.GO, 9), (_orc.CHESS, 8)])
def test_eval_cache_across_waves_is_result_transparent(game, board):
    """Evaluation cache across waves (M16 / SURVEY 8f.3: the reference's TranspositionTable, 64-bit key -> (policy, value)) behind the ResNet
    evaluator.  Key = the whole network input, policies kept in fp32: a hit returns exactly what the network would compute, so searches are
    bit-identical with the cache on and off.  Each engine searches the same roots TWICE (fresh trees the second time): with the cache on
    the second pass is served almost entirely from the cache; all four runs must agree bit for bit on every slot and move."""
    from _eng import E, N
    O = _orc.oracle()
    planes, actions = {GOMOKU: (11, board * board), _orc.GO: (8, board * board + 1), _orc.CHESS: (18, 20480)}[game]
    m = N.make_random_model(seed=5, randomize_bn=True, blocks=2, in_planes=planes, board=board, actions=actions)
    blob = N.export_weights(m)
    rng = np.random.default_rng(11)
    openings = []
    for _ in range(8):
        s = O.new_state(game, board); mv = []
        for _ in range(int(rng.integers(0, 10))):
            lg = O.legal(s)
            a = int(rng.choice(lg[lg >= 0] if game == _orc.GO else lg))
            O.state_make_move(s, a); mv.append(a)
        openings.append(mv)
    runs = {}
    for cache in (1 << 16, -1):
        eng = E.Engine(game=game, board_size=board, n_slots=len(openings), evaluator=E.EVAL_RESNET, net_blocks=2, num_simulations=64, deterministic=1,
                       auto_restart=0, eval_cache_entries=cache)
        eng.load_weights(blob)
        for rep in range(2):
            for t, mv in enumerate(openings):
                eng.set_root(t, mv)
            res = []
            for move in range(2):
                eng.search()
                res.append([eng.root_stats(t) for t in range(len(openings))])
                eng.advance([int(r["actions"][int(np.argmax(r["N"]))]) if len(r["N"]) else -2 for r in res[-1]])
            runs[(cache, rep)] = (res, eng.stats())
        eng.close()
    ref = runs[(-1, 0)][0]
    for key, (res, _) in runs.items():
        for move in range(2):
            for t in range(len(openings)):
                a, b = res[move][t], ref[move][t]
                assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (key, move, t)
                assert np.array_equal(a["W"].view(np.uint32), b["W"].view(np.uint32)) and np.array_equal(a["P"].view(np.uint32), b["P"].view(np.uint32)), (key, move, t)
    on0, on1, off1 = runs[(1 << 16, 0)][1], runs[(1 << 16, 1)][1], runs[(-1, 1)][1]
    assert off1["eval_cached"] == 0 and on1["evaluations"] == off1["evaluations"]
    second_pass = on1["evaluations"] - on0["evaluations"]
    assert on1["eval_cached"] - on0["eval_cached"] >= 0.9 * second_pass, (on0, on1)      # the second pass re-evaluates (almost) nothing


def test_set_num_simulations_between_moves():
    """az_engine_set_num_simulations (ParallelMCTS::setNumSimulations, parallel_mcts.cpp:1183-1185): az_engine_play uses the new count from the
    next move on; growing it back re-cuts the node regions before the search; a count the node pool cannot hold is refused."""
    from _eng import hash_engine
    eng = hash_engine(8, board=9, sims=300, deterministic=0, auto_restart=1, max_nodes_per_tree=0)
    eng.set_num_simulations(40); eng.play(1)
    assert eng.stats()["simulations"] == 8 * 40
    eng.set_num_simulations(12); eng.play(2)
    assert eng.stats()["simulations"] == 8 * (40 + 2 * 12)
    eng.set_num_simulations(300); eng.play(1)
    st = eng.stats()
    assert st["simulations"] == 8 * (40 + 24 + 300) and st["pool_overflows"] == 0
    with pytest.raises(RuntimeError):
        eng.set_num_simulations(0)
    with pytest.raises(RuntimeError, match="node pool"):
        eng.set_num_simulations(100000)           # beyond what the pool was sized for at creation
    eng.close()


def test_eval_cache_tiny_capacity_with_stream_groups_is_result_transparent():
    """The evaluation cache under pressure: 128 entries per stream group (32 buckets of 4 ways: almost every store evicts), two stream
    groups (one table each), ResNet evaluator, Go 9x9 — bit-identical to the same engine with the cache off, and the cache does serve leaves:
    odd slots start one pass into the game, i.e. on the position every even slot's tree evaluates in its first simulation (pass is child 0
    and unvisited children are taken in order), one wave after the odd slots' root expansion stored it.  What gets stored depends on warp
    timing; what a hit returns must not."""
    from _eng import E, N
    m = N.make_random_model(seed=9, randomize_bn=True, blocks=2, in_planes=8, board=9, actions=82)
    blob = N.export_weights(m)
    runs = []
    for cache in (256, -1):
        eng = E.Engine(game=E.GO, board_size=9, n_slots=12, evaluator=E.EVAL_RESNET, net_blocks=2, num_simulations=160, deterministic=1,
                       auto_restart=0, n_streams=2, eval_cache_entries=cache)
        eng.load_weights(blob)
        for t in range(12):
            eng.set_root(t, [-1] if t % 2 else [])
        res = []
        for move in range(4):
            eng.search()
            r = [eng.root_stats(t) for t in range(12)]
            res.append(r)
            eng.advance([int(x["actions"][int(np.argmax(x["N"]))]) if len(x["N"]) else -2 for x in r])
        runs.append((res, eng.stats()))
        eng.close()
    (on, st_on), (off, st_off) = runs
    assert st_on["eval_cached"] > 0 and st_off["eval_cached"] == 0 and st_on["evaluations"] == st_off["evaluations"] and st_on["pool_overflows"] == 0
    for move in range(4):
        for t in range(12):
            a, b = on[move][t], off[move][t]
            assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (move, t)
            assert np.array_equal(a["W"].view(np.uint32), b["W"].view(np.uint32)) and np.array_equal(a["P"].view(np.uint32), b["P"].view(np.uint32)), (move, t)
