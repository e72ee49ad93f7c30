"""GPU parity tests for Go (BASELINE.json configs[2]): device rules (capture, ko, positional
superko, suicide, area score, legal-move order with pass first, 8-plane encoder) and the batched search, called through the
C ABI, against the oracle (CPU restatement pinned to the patched reference) and the golden fixtures generated from the
reference itself.  Bar: legal moves, terminal flags, results, planes and per-root visit counts bit-exact."""
import json
import os

import numpy as np
import pytest

import _orc
from _orc import GO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def go_engine(n_slots, board=9, sims=400, **kw):
    from _eng import E
    cfg = dict(game=E.GO, board_size=board, n_slots=n_slots, num_simulations=sims, evaluator=E.EVAL_HASH, deterministic=1,
               auto_restart=0, max_nodes_per_tree=2 * (sims + 1) * (board * board + 1) + 1, n_streams=2)
    cfg.update(kw)
    return E.Engine(**cfg)


@pytest.mark.parametrize("board", [9, 19])
def test_go_rules_kernels_match_golden_and_oracle(board):
    """Every prefix of the golden playouts (generated from the patched reference: captures, ko, superko, passes):
    getLegalMoves in order, isTerminal, getGameResult, getCurrentPlayer, the 8 feature planes."""
    O = _orc.oracle()
    eng = go_engine(2, board=board, sims=4)
    games, expect = [], []
    for case in json.load(open(os.path.join(GOLD, "state_playouts.json"))):
        if case["game"] != GO or case["board"] != board:
            continue
        s = O.new_state(GO, board)
        step = 1 if board == 9 else 3
        for ply in range(len(case["moves"]) + 1):
            if ply % step == 0 or ply == len(case["moves"]):
                games.append(case["moves"][:ply])
                expect.append((O.legal(s), O.state_is_terminal(s), O.state_result(s), O.state_current_player(s), O.tensor(s)))
            if ply < len(case["moves"]):
                assert O.state_make_move(s, case["moves"][ply]) == 0
    assert len(games) > 80
    r = eng.rules_replay(games)
    for i, (legal, term, res, pl, planes) in enumerate(expect):
        assert r["n_legal"][i] >= 0, (board, i)
        assert np.array_equal(r["legal"][i], legal), (board, i, len(games[i]))
        assert r["terminal"][i] == term and r["result"][i] == res and r["player"][i] == pl, (board, i)
        assert np.array_equal(r["planes"][i], planes), (board, i)
    eng.close()


def test_go_rules_edge_cases():
    """Occupied cell, suicide, simple ko and out-of-range moves are rejected like the reference's makeMove (it throws);
    two passes end the game and the empty board scores for White (komi 7.5)."""
    O = _orc.oracle()
    eng = go_engine(2, board=9, sims=4)
    # a ko: black 1,9+? build the classic shape around (x=1,y=1)
    #   . B W .        black: (1,0) (0,1) (1,2)   white: (2,0) (3,1) (2,2), then white plays (1,1)?? no:
    # black stones at 1, 9, 19; white stones at 2, 12, 20; black plays 11 (x=2,y=1); white captures at 10 (x=1,y=1)
    ko = [1, 2, 9, 12, 19, 20, 11, 10]
    cases = [[0, 0], [81], [-2], ko + [11], ko, [-1, -1], [-1, -1, 5]]
    r = eng.rules_replay(cases)
    exp_bad = []
    for c in cases:
        s = O.new_state(GO, 9)
        bad = False
        for a in c:
            if O.state_make_move(s, a) != 0:
                bad = True
                break
        exp_bad.append(bad)
        if not bad:
            i = len(exp_bad) - 1
            assert np.array_equal(r["legal"][i], O.legal(s)), c
            assert r["terminal"][i] == O.state_is_terminal(s) and r["result"][i] == O.state_result(s), c
    assert [int(n) == -1 for n in r["n_legal"]] == exp_bad
    assert exp_bad[3] is True and exp_bad[4] is False          # immediate ko recapture is illegal, the ko itself is fine
    assert r["terminal"][5] == 1 and r["result"][5] == _orc.WIN_P2
    # suicide: white fills its own last liberty in the corner surrounded by black
    su = [1, 40, 9, 0]
    s = O.new_state(GO, 9)
    ok = [O.state_make_move(s, a) for a in su]
    rr = eng.rules_replay([su])
    assert (rr["n_legal"][0] == -1) == (ok[-1] != 0)
    eng.close()


@pytest.mark.parametrize("cache", [0, 1 << 14])
def test_go_search_matches_reference_golden(cache):
    """Serial-search parity on the golden case generated from the patched reference (Go 9x9 @400 sims):
    child order (pass first), visit counts, valueSum and prior bits, root leak, chosen move.  cache > 0: the same with the evaluation
    cache (M16, the reference's TranspositionTable) switched on behind the hash evaluator — transpositions inside the trees are served
    from the cache and the search still equals the reference's bit for bit (the reference ran with ITS table on)."""
    cases = [c for c in json.load(open(os.path.join(GOLD, "search_hash_eval.json"))) if c["game"] == GO]
    assert cases
    case = cases[0]
    eng = go_engine(3, board=case["board"], sims=case["sims"], eval_cache_entries=cache)
    for rep in range(2 if cache else 1):           # second pass (fresh trees, same roots): every evaluation comes out of the cache
        for slot in range(3):
            eng.set_root(slot, [])
        for mv, g in enumerate(case["moves"]):
            eng.search()
            for slot in range(3):
                st = eng.root_stats(slot)
                assert st["actions"].tolist() == g["actions"], (rep, mv, slot)
                assert st["N"].tolist() == g["N"], (rep, mv, slot)
                assert bits(st["W"]).tolist() == g["W"], (rep, mv, slot)
                assert bits(st["P"]).tolist() == g["P"], (rep, mv, slot)
                assert st["rootN"] == g["rootN"] and int(bits([st["rootW"]])[0]) == g["rootW"]
            eng.advance([g["action"]] * 3)
    st = eng.stats()
    assert st["pool_overflows"] == 0
    assert (st["eval_cached"] > 0) == (cache > 0), st
    eng.close()


def test_go_search_matches_oracle_random_positions():
    """A different mid-game position (random legal playouts with captures) in every slot, searched together wave by wave,
    each compared with the oracle's serial search for several consecutive moves (subtree reuse included)."""
    O = _orc.oracle()
    rng = np.random.default_rng(5)
    T, sims, board = 10, 200, 9
    eng = go_engine(T, board=board, sims=sims)
    searches = []
    for t in range(T):
        s = O.new_state(GO, board)
        moves = []
        for _ in range(int(rng.integers(0, 70))):
            lg = O.legal(s)
            cand = lg[lg >= 0] if (len(lg) > 1 and rng.random() < 0.97) else lg
            a = int(rng.choice(cand))
            assert O.state_make_move(s, a) == 0
            moves.append(a)
            if O.state_is_terminal(s):
                break
        if O.state_is_terminal(s):
            s = O.new_state(GO, board); moves = []
        eng.set_root(t, moves)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 0, None, None))
    for mv in range(3):
        eng.search()
        acts = []
        for t in range(T):
            O.mcts_search(searches[t])
            a, b = eng.root_stats(t), O.root_stats(searches[t])
            assert np.array_equal(a["actions"], b["actions"]), (mv, t)
            assert np.array_equal(a["N"], b["N"]), (mv, t)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (mv, t)
            assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0]
            act = O.mcts_select_action(searches[t], 1, 1.0)
            acts.append(act)
            O.mcts_update_with_move(searches[t], act)
        eng.advance(acts)
    assert eng.stats()["pool_overflows"] == 0
    eng.close()


@pytest.mark.parametrize("board,sims", [(13, 400), (19, 500)])
def test_go_search_matches_oracle_13_and_19(board, sims):
    """The larger boards of BASELINE.json configs[3] (19x19: 6-word bitboards, 362 children per node) and 13x13: mid-game positions with
    captures from random legal playouts, searched together and compared with the oracle's serial search bit for bit over two moves."""
    O = _orc.oracle()
    rng = np.random.default_rng(board)
    T = 4
    eng = go_engine(T, board=board, sims=sims)
    searches = []
    for t in range(T):
        s = O.new_state(GO, board)
        moves = []
        for _ in range([0, 60, 150, 260][t] * board * board // 361):
            lg = O.legal(s)
            cand = lg[lg >= 0] if len(lg) > 1 else lg
            a = int(rng.choice(cand))
            assert O.state_make_move(s, a) == 0
            moves.append(a)
        assert not O.state_is_terminal(s)
        eng.set_root(t, moves)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 0, None, None))
    for mv in range(2):
        eng.search()
        acts = []
        for t in range(T):
            O.mcts_search(searches[t])
            a, b = eng.root_stats(t), O.root_stats(searches[t])
            assert np.array_equal(a["actions"], b["actions"]), (board, mv, t)
            assert np.array_equal(a["N"], b["N"]), (board, mv, t)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (board, mv, t)
            assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0]
            act = O.mcts_select_action(searches[t], 1, 1.0)
            acts.append(act)
            O.mcts_update_with_move(searches[t], act)
        eng.advance(acts)
    assert eng.stats()["pool_overflows"] == 0
    eng.close()


def test_go_selfplay_loop_matches_oracle():
    """az_engine_play in deterministic mode against the oracle's playSingleGame restatement for the first 30 moves of a
    9x9 game: same moves, same recorded visit counts (pass = last entry of the visit vector)."""
    O = _orc.oracle()
    board, sims = 9, 100
    eng = go_engine(2, board=board, sims=sims)
    s = O.new_state(GO, board)
    m = O.mcts_new(s, sims, 1.5, 3, 0, None, None)
    for i in range(30):
        if O.state_is_terminal(s):
            break
        O.mcts_search(m)
        a = O.mcts_select_action(m, 1, 1.0)
        O.state_make_move(s, a); O.mcts_update_with_move(m, a)
        eng.play(1)
        assert eng.last_actions().tolist() == [a, a], i
        r, ply, pl = eng.slot_state(0)
        assert ply == i + 1 and pl == O.state_current_player(s) and r == O.state_result(s)
    eng.close()


def test_go_selfplay_auto_restart_and_noise_smoke():
    """Throughput-mode loop on Go 9x9 (Dirichlet noise, temperature sampling, auto-restart, move cap): invariants only."""
    board, sims, T = 9, 32, 64
    eng = go_engine(T, board=board, sims=sims, deterministic=0, auto_restart=1, seed=3)
    total = 0
    for _ in range(60):
        eng.play(3)
        smp = eng.drain_samples()
        total += len(smp)
        if len(smp):
            assert np.all(smp["visits"].sum(1) >= sims - 1)
            assert np.all(np.abs(smp["z"]) <= 1) and np.all(smp["result"] >= 1)
            idx = np.where(smp["action"] < 0, board * board, smp["action"])
            assert np.all(smp["visits"][np.arange(len(smp)), idx] >= 1)
            assert np.all(smp["ply"] < 2 * board * board)
    st = eng.stats()
    assert st["games"] >= 1 and total >= 1 and st["pool_overflows"] == 0 and st["samples_dropped"] == 0
    assert st["moves"] == 180 * T
    eng.close()


def test_go_selfplay_with_resnet_evaluator():
    """Go 9x9 self-play with the tcgen05 ResNet evaluator in the loop (random-init 2-block net): the wave pipeline
    (select -> 8-plane encode -> trunk + heads -> expand/backup) runs, every simulation is accounted for, root visit
    counts add up, and the evaluator's policy for the root position matches the fp32 network."""
    import torch
    from _eng import E, N
    O = _orc.oracle()
    board, sims, T = 9, 24, 48
    model = N.make_random_model(seed=4, blocks=2, in_planes=8, board=board, actions=board * board + 1)
    with torch.no_grad():
        model.p_fc.weight *= 0.2; model.v_fc1.weight *= 0.2
    eng = E.Engine(game=E.GO, board_size=board, n_slots=T, evaluator=E.EVAL_RESNET, net_blocks=2, num_simulations=sims,
                   deterministic=1, auto_restart=1)
    eng.load_weights(N.export_weights(model))
    eng.search()
    st = eng.root_stats(5)
    assert st["actions"].tolist() == [-1] + list(range(board * board))
    assert int(st["N"].sum()) == sims
    x = O.tensor(O.new_state(GO, board))[None]
    with torch.no_grad():
        p32 = torch.softmax(model(torch.tensor(x))[0], 1).numpy()[0]
    prior = st["P"][1:]                                      # children 1.. are the cells in ascending order
    ref = p32[:board * board] / p32[:board * board].sum()   # expandNodeWithPolicy renormalises over the legal moves
    assert st["P"][0] == 0.0 and np.abs(prior - ref).max() <= 2e-3
    eng.play(3)
    s = eng.stats()
    assert s["moves"] == 3 * T and s["simulations"] == 4 * sims * T and s["pool_overflows"] == 0
    eng.close()


def test_go_full_size_2048_slots_identical_and_golden():
    """BASELINE.json configs[2] at full size (Go 9x9, 2048 slots, 400 simulations), hash evaluator, deterministic mode:
    every slot must hold the same tree, bit-identical to the golden search generated from the reference."""
    cases = [c for c in json.load(open(os.path.join(GOLD, "search_hash_eval.json"))) if c["game"] == GO]
    case = cases[0]
    assert case["board"] == 9 and case["sims"] == 400
    T = 2048
    eng = go_engine(T, board=9, sims=400, n_streams=1)
    for mv, g in enumerate(case["moves"][:3]):
        eng.search()
        for slot in (0, 3, 1023, 1024, 2047):
            st = eng.root_stats(slot)
            assert st["actions"].tolist() == g["actions"] and st["N"].tolist() == g["N"], (mv, slot)
            assert bits(st["W"]).tolist() == g["W"] and bits(st["P"]).tolist() == g["P"], (mv, slot)
        eng.advance([g["action"]] * T)
    st = eng.stats()
    assert st["simulations"] == 3 * 400 * T and st["pool_overflows"] == 0
    eng.close()


def test_go19_full_size_1024_slots_identical_and_oracle():
    """BASELINE.json configs[3] at its per-GPU size (Go 19x19, 1024 slots, 400 simulations), hash evaluator, deterministic mode: the slots
    alternate between the empty board and a 150-move middle game with captures; every slot of a root must hold the same tree, equal to the
    oracle's serial search bit for bit, over two moves (subtree reuse, region re-cut at 362 children per node)."""
    O = _orc.oracle()
    board, sims, T = 19, 400, 1024
    rng = np.random.default_rng(19)
    s1 = O.new_state(GO, board)
    mid = []
    for _ in range(150):
        lg = O.legal(s1)
        a = int(rng.choice(lg[lg >= 0]))
        assert O.state_make_move(s1, a) == 0
        mid.append(a)
    assert not O.state_is_terminal(s1)
    roots, states = [[], mid], [O.new_state(GO, board), s1]
    R = len(roots)
    searches = [O.mcts_new(st, sims, 1.5, 3, 0, None, None) for st in states]
    eng = go_engine(T, board=board, sims=sims, n_streams=1)
    for t in range(T):
        eng.set_root(t, roots[t % R])
    for move in range(2):
        eng.search()
        acts = []
        for r in range(R):
            O.mcts_search(searches[r])
            b = O.root_stats(searches[r])
            a = eng.root_stats(r)
            assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (move, r)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (move, r)
            assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0], (move, r)
            acts.append(O.mcts_select_action(searches[r], 1, 1.0))
            O.mcts_update_with_move(searches[r], acts[-1])
        first = [eng.root_stats(r) for r in range(R)]
        for t in range(R, T):
            a = eng.root_stats(t)
            assert np.array_equal(a["N"], first[t % R]["N"]) and np.array_equal(bits(a["W"]), bits(first[t % R]["W"])), (move, t)
        eng.advance([acts[t % R] for t in range(T)])
    st = eng.stats()
    assert st["simulations"] == 2 * sims * T and st["pool_overflows"] == 0, st
    eng.close()


def test_go_peaked_policy_deep_trees_no_overflow():
    """Node-pool behaviour under a PEAKED policy (what a trained network produces; AZ_EVAL_HASH_PEAKED = the hash evaluator with one
    action's raw prior x 4096): after the first sweep over the children most simulations go below one child, that child is played, and
    its subtree — 450-520 visits on top of which the next 400 land — is kept move after move, so a tree holds more than two searches'
    worth of expansions: more than the fixed 2 x (sims + 1) x A nodes per tree of round 1, which overflowed here.  Go 9x9 @400, 20
    consecutive moves on 4 roots; regions are re-cut at every move commit from the kept-subtree sizes (tree.cuh).  Asserts 0 failed
    expansions and bit-equality with the oracle (== the reference, tests/test_oracle.py) on every move."""
    O = _orc.oracle()
    board, sims = 9, 400
    from _eng import E
    openings = [[], [40], [30, 50], [20, 60, 41]]
    eng = go_engine(len(openings), board=board, sims=sims, evaluator=E.EVAL_HASH_PEAKED, max_nodes_per_tree=int(2.6 * (sims + 1) * 82))
    searches = []
    for t, mv in enumerate(openings):
        s = O.new_state(GO, board)
        for a in mv:
            assert O.state_make_move(s, a) == 0
        eng.set_root(t, mv)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 2, None, None))
    shares, deepest = [], 0
    for move in range(20):
        eng.search()
        acts = []
        for t in range(len(openings)):
            O.mcts_search(searches[t])
            a, b = eng.root_stats(t), O.root_stats(searches[t])
            assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (move, t)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (move, t)
            assert a["rootN"] == b["rootN"]
            shares.append(a["N"].max() / max(1, a["N"].sum())); deepest = max(deepest, int(a["N"].sum()))
            act = O.mcts_select_action(searches[t], 1, 1.0)
            acts.append(act)
            O.mcts_update_with_move(searches[t], act)
        eng.advance(acts)
    assert eng.stats()["pool_overflows"] == 0
    # the trees really are deeper than one search: the root's children hold more than 2 x sims visits, the played child about half of them
    assert deepest > 2 * sims and np.mean(shares) > 0.45
    eng.close()


def test_pool_exhaustion_is_an_error_not_silent():
    """A pool that cannot hold the search: az_engine_search fails (deterministic / parity mode) instead of leaving leaves unexpanded."""
    from _eng import E
    eng = go_engine(2, board=9, sims=200, evaluator=E.EVAL_HASH_PEAKED, max_nodes_per_tree=60 * 82)
    with pytest.raises(E.EngineError, match="node pool exhausted"):
        for _ in range(6):
            eng.search()
            eng.advance([int(eng.root_stats(t)["actions"][int(np.argmax(eng.root_stats(t)["N"]))]) for t in range(2)])
    eng.close()
