"""GPU parity of az_engine_make_examples (device-side Dataset::extractExamples + augmentExample, src/selfplay/dataset.cpp:64-114,
245-436): sample records of self-played games -> planes / policy / value with the 8 dihedral images, against the
oracle (state replay for the planes, the augmentExample restatement for the images).  Bit-exact."""
import numpy as np
import pytest

import _orc

pytestmark = pytest.mark.gpu


def _play_and_collect(game, board, sims, slots, steps, **kw):
    from _eng import E
    mc = {E.GOMOKU: board * board, E.GO: board * board + 1, E.CHESS: 256}[game]
    eng = E.Engine(game=game, board_size=board, n_slots=slots, num_simulations=sims, evaluator=E.EVAL_HASH, deterministic=0, auto_restart=1,
                   max_nodes_per_tree=2 * (sims + 2) * mc + 1, seed=11, **kw)
    smp = []
    for _ in range(steps):
        eng.play(4)
        s = eng.drain_samples()
        if len(s):
            smp.append(s.copy())
        if sum(len(x) for x in smp) > 300:
            break
    return eng, (np.concatenate(smp) if smp else None)


def _games(smp):
    """Group records into games (same slot + game id), ply order."""
    out = {}
    for r in smp:
        out.setdefault((int(r["slot"]), int(r["game_id"])), []).append(r)
    return [sorted(v, key=lambda r: int(r["ply"])) for v in out.values()]


@pytest.mark.parametrize("game,board", [(_orc.GOMOKU, 9), (_orc.GO, 9)])
def test_examples_match_oracle_replay_and_augmentation(game, board):
    O = _orc.oracle()
    eng, smp = _play_and_collect(game, board, sims=20, slots=24, steps=60)
    assert smp is not None and len(smp) >= 40
    planes, policy, value = eng.make_examples(smp, augment=True)
    assert planes.shape[0] == 8 * len(smp)
    idx = {(int(r["slot"]), int(r["game_id"]), int(r["ply"])): i for i, r in enumerate(smp)}
    checked = 0
    for g in _games(smp)[:6]:
        s = O.new_state(game, board)
        assert int(g[0]["ply"]) == 0
        for r in g:
            i = idx[(int(r["slot"]), int(r["game_id"]), int(r["ply"]))]
            t = O.tensor(s)
            v = r["visits"].astype(np.float32)
            a = board * board + (1 if game == _orc.GO else 0)
            pol = (v[:a] / np.float32(v[:a].sum())).astype(np.float32)
            res = int(r["result"]); pl = O.state_current_player(s)
            z = 0.0 if res == 1 else (1.0 if (res == 2) == (pl == 1) else -1.0)      # dataset.cpp:84-96
            apl, apo = _orc.augment_example(t, pol)
            assert np.array_equal(planes[8 * i], t) and np.array_equal(policy[8 * i], pol) and value[8 * i] == z
            for k in range(7):
                assert np.array_equal(planes[8 * i + 1 + k], apl[k]), (i, k)
                assert np.array_equal(policy[8 * i + 1 + k], apo[k]), (i, k)
                assert value[8 * i + 1 + k] == z
            assert O.state_make_move(s, int(r["action"])) == 0
            checked += 1
    assert checked >= 30
    p1, q1, v1 = eng.make_examples(smp[:5], augment=False)
    assert np.array_equal(p1, planes[0:40:8]) and np.array_equal(q1, policy[0:40:8]) and np.array_equal(v1, value[0:40:8])
    eng.close()


def test_chess_examples_match_oracle_replay():
    """Chess: no augmentation (dataset.cpp:250-253); 18 planes incl. the repetition plane carried in the snapshot; dense 20480-entry
    policy scattered from the (action, count) pairs."""
    O = _orc.oracle()
    eng, smp = _play_and_collect(_orc.CHESS, 8, sims=12, slots=48, steps=140, sample_ring_capacity=48 * 600)
    if smp is None:
        pytest.skip("no chess game finished within the step budget")
    planes, policy, value = eng.make_examples(smp, augment=True)
    assert planes.shape == (len(smp), 18, 8, 8) and policy.shape == (len(smp), 20480)
    idx = {(int(r["slot"]), int(r["game_id"]), int(r["ply"])): i for i, r in enumerate(smp)}
    checked = 0
    for g in _games(smp)[:4]:
        s = O.new_state(_orc.CHESS, 8)
        for r in g:
            i = idx[(int(r["slot"]), int(r["game_id"]), int(r["ply"]))]
            assert np.array_equal(planes[i], O.tensor(s)), (i, int(r["ply"]))
            pairs = r["visits"].reshape(-1, 2).astype(np.int64)
            legal = O.legal(s)
            assert pairs[:len(legal), 0].tolist() == legal.tolist()
            tot = np.float32(pairs[:, 1].sum())
            dense = np.zeros(20480, np.float32); dense[legal] = pairs[:len(legal), 1].astype(np.float32) / tot
            assert np.array_equal(policy[i], dense)
            assert O.state_make_move(s, int(r["action"])) == 0
            checked += 1
    assert checked >= 10
    eng.close()


@pytest.mark.parametrize("game,board", [(_orc.GOMOKU, 9), (_orc.GO, 9), (_orc.CHESS, 8)])
def test_examples_from_game_records_match_sample_path_and_oracle(game, board):
    """az_engine_examples_from_games (Dataset::addGameRecord + extractExamples for records that arrive as move lists) against
    (1) az_engine_make_examples on the sample records of the same games — bit-identical planes / policy / value — and (2) the oracle's
    augmentExample restatement for a policy SHORTER than N*N (child-ordered policies: an entry moves only when old and new index fit)."""
    from _eng import E
    kw = dict(sample_ring_capacity=48 * 600) if game == _orc.CHESS else {}
    eng, smp = _play_and_collect(game, board, sims=12 if game == _orc.CHESS else 20, slots=48 if game == _orc.CHESS else 24,
                                 steps=140 if game == _orc.CHESS else 60, **kw)
    if smp is None:
        pytest.skip("no game finished within the step budget")
    games = _games(smp)[:8]
    recs = np.concatenate([np.array(g, dtype=smp.dtype) for g in games])
    planes, policy, value = eng.make_examples(recs, augment=True)
    moves = [[int(r["action"]) for r in g] for g in games]
    results = [int(g[0]["result"]) for g in games]
    k = 1 if game == _orc.CHESS else 8
    dense = policy[::k].copy()                                       # the action-indexed targets of the original images
    pl2, po2, va2 = eng.examples_from_games(moves, results, dense, augment=True)
    assert pl2.shape == planes.shape and np.array_equal(pl2, planes)
    assert np.array_equal(po2, policy)
    assert np.array_equal(va2, value)
    if game != _orc.CHESS:
        short = np.ascontiguousarray(dense[:, :board * 4 + 3])       # P < N*N
        pl3, po3, _ = eng.examples_from_games(moves, results, short, augment=True)
        assert np.array_equal(pl3, planes)
        for i in range(0, len(short), 7):
            apl, apo = _orc.augment_example(planes[8 * i], short[i])
            assert np.array_equal(po3[8 * i], short[i])
            for j in range(7):
                assert np.array_equal(po3[8 * i + 1 + j], apo[j]), (i, j)
                assert np.array_equal(pl3[8 * i + 1 + j], apl[j]), (i, j)
    # an illegal recorded move is an error, as the reference's makeMove throws (dataset.cpp:76)
    bad = [list(moves[0])]
    if len(bad[0]) >= 2:
        bad[0][1] = bad[0][0] if game != _orc.CHESS else 0           # occupied point / a1a1
        with pytest.raises(E.EngineError):
            eng.examples_from_games(bad, results[:1], dense[:len(bad[0])], augment=False)
    eng.close()
