"""GPU: the reference-shaped Python API (`_alphazero_cpp`) end to end against the oracle — the test reads like the
reference's own usage (python/scripts/self_play.py:339-372, python/tests/test_binding.py)."""
import os
import sys

import numpy as np
import pytest

import _orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-multi-game_b200"))


def test_parallel_mcts_api_matches_oracle():
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 15)
    state = az.createGameState(az.GameType.GOMOKU, 15, False)
    o_state = O.new_state(_orc.GOMOKU, 15)
    mcts = az.ParallelMCTS(state, nn, None, 1, 300, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    om = O.mcts_new(o_state, 300, 1.5, 3, 0, None, None)
    for mv in range(4):
        mcts.search(); O.mcts_search(om)
        actions, visits, wsum, priors, root_n, root_w = mcts.getRootChildren()
        b = O.root_stats(om)
        assert actions == b["actions"].tolist() and visits == b["N"].tolist()
        assert np.array_equal(np.array(wsum, np.float32).view(np.uint32), b["W"].view(np.uint32))
        assert np.array_equal(np.array(mcts.getActionProbabilities(1.0), np.float32), O.probs(om, 1.0))
        assert np.float32(mcts.getRootValue()) == np.float32(O.mcts_root_value(om))
        a = mcts.selectAction(True, 1.0)
        assert a == O.mcts_select_action(om, 1, 1.0)
        state.makeMove(a); mcts.updateWithMove(a); O.mcts_update_with_move(om, a)
    assert "Total visits" in mcts.getSearchInfo()


def test_selfplay_manager_generate_games_matches_oracle():
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 9)
    mgr = az.SelfPlayManager(nn, 3, 80, 1)
    mgr.setDeterministic(True)
    mgr.setConcurrentGames(3)
    seen = []
    mgr.setProgressCallback(lambda g, m, tg, tm: seen.append((g, m, tg, tm)))
    games = mgr.generateGames(az.GameType.GOMOKU, 9, False)
    assert len(games) == 3 and mgr.getCompletedGamesCount() == 3 and len(seen) == 3
    s = O.new_state(_orc.GOMOKU, 9)
    m = O.mcts_new(s, 80, 1.5, 3, 0, None, None)
    moves = []
    while not O.state_is_terminal(s):
        O.mcts_search(m)
        a = O.mcts_select_action(m, 1, 1.0)
        moves.append(a); O.state_make_move(s, a); O.mcts_update_with_move(m, a)
    for g in games:            # deterministic mode: every slot plays the oracle's game
        assert [mv.action for mv in g.getMoves()] == moves
        assert int(g.getResult()) == O.state_result(s)
        assert abs(sum(g.getMoves()[0].policy) - 1.0) < 1e-5


def test_parallel_mcts_api_on_go_matches_oracle():
    """Same reference-shaped calls on Go 9x9 (createGameState(GO) -> host GoState, device Go rules in the search)."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.GO, 9)
    state = az.createGameState(az.GameType.GO, 9, False)
    o_state = O.new_state(_orc.GO, 9)
    for a in (40, 41, 31, 49):
        state.makeMove(a); O.state_make_move(o_state, a)
    mcts = az.ParallelMCTS(state, nn, None, 1, 200, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    om = O.mcts_new(o_state, 200, 1.5, 3, 0, None, None)
    for mv in range(3):
        mcts.search(); O.mcts_search(om)
        actions, visits, wsum, priors, root_n, root_w = mcts.getRootChildren()
        b = O.root_stats(om)
        assert actions == b["actions"].tolist() and actions[0] == -1 and visits == b["N"].tolist()
        assert np.array_equal(np.array(wsum, np.float32).view(np.uint32), b["W"].view(np.uint32))
        a = mcts.selectAction(True, 1.0)
        assert a == O.mcts_select_action(om, 1, 1.0)
        state.makeMove(a); mcts.updateWithMove(a); O.mcts_update_with_move(om, a)
    pol, val = nn.predict(state)
    po, vo = O.hash_policy_value(o_state) if False else (None, None)
    assert len(pol) == 82


def test_selfplay_manager_generate_go_games():
    import _alphazero_cpp as az
    nn = az.createNeuralNetwork("hash", az.GameType.GO, 9)
    mgr = az.SelfPlayManager(nn, 4, 24, 1)
    mgr.setConcurrentGames(4)
    games = mgr.generateGames(az.GameType.GO, 9, False)
    assert len(games) == 4
    for g in games:
        mv = g.getMoves()
        assert 2 <= len(mv) <= 162 and int(g.getResult()) in (1, 2, 3)
        assert len(mv[0].policy) == 82 and abs(sum(mv[0].policy) - 1.0) < 1e-5
        assert all(-1 <= m.action < 81 for m in mv)


def test_parallel_mcts_and_selfplay_on_chess():
    """The reference-shaped API on chess: createGameState(CHESS) -> host ChessState, ParallelMCTS against the oracle's serial
    search (hash evaluator), SelfPlayManager.generateGames returning child-ordered policies."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.CHESS, 8)
    state = az.createGameState(az.GameType.CHESS, 0, False)
    o_state = O.new_state(_orc.CHESS, 8)
    for mv in ("e2e4", "e7e5", "g1f3"):
        a = state.stringToAction(mv); state.makeMove(a); assert O.state_make_move(o_state, a) == 0
    mcts = az.ParallelMCTS(state, nn, None, 1, 100, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    om = O.mcts_new(o_state, 100, 1.5, 3, 0, None, None)
    for mv in range(2):
        mcts.search(); O.mcts_search(om)
        actions, visits, wsum, priors, root_n, root_w = mcts.getRootChildren()
        b = O.root_stats(om)
        assert actions == b["actions"].tolist() and visits == b["N"].tolist()
        assert np.array_equal(np.array(wsum, np.float32).view(np.uint32), b["W"].view(np.uint32))
        a = mcts.selectAction(True, 1.0)
        assert a == O.mcts_select_action(om, 1, 1.0)
        state.makeMove(a); mcts.updateWithMove(a); O.mcts_update_with_move(om, a)
    mgr = az.SelfPlayManager(nn, 2, 16, 1)
    mgr.setConcurrentGames(8)
    games = mgr.generateGames(az.GameType.CHESS, 0, False)
    assert len(games) == 2
    for g in games:
        mv = g.getMoves()
        assert 2 <= len(mv) <= 512 and int(g.getResult()) in (1, 2, 3)
        assert 1 <= len(mv[0].policy) <= 218 and abs(sum(mv[0].policy) - 1.0) < 1e-5
