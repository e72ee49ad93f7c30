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


@pytest.mark.parametrize("game,board", [("GOMOKU", 9), ("GO", 9)])
def test_dataset_extract_examples_matches_oracle(game, board):
    """Dataset.addGameRecord + extractExamples (dataset.cpp:60-114, 245-436) on records produced by SelfPlayManager: planes = the oracle's
    tensor of the replayed state, the 7 images = the augmentExample restatement, value = result seen from the player to move."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    gt = getattr(az.GameType, game)
    og = {"GOMOKU": _orc.GOMOKU, "GO": _orc.GO}[game]
    nn = az.createNeuralNetwork("hash", gt, board)
    mgr = az.SelfPlayManager(nn, 3, 24, 1)
    mgr.setConcurrentGames(6)
    games = mgr.generateGames(gt, board, False)
    assert len(games) == 3
    d = az.Dataset()
    d.setShuffleOnExtract(False)
    for g in games:
        d.addGameRecord(g)
    d.extractExamples(True)
    n_moves = sum(len(g.getMoves()) for g in games)
    assert d.size() == 8 * n_moves
    ex = d.examples()
    i = 0
    for g in games:
        s = O.new_state(og, board)
        res = int(g.getResult())
        for mv in g.getMoves():
            t = O.tensor(s)
            pol = np.array(mv.policy, np.float32)
            z = 0.0 if res in (0, 1) else (1.0 if (res == 2) == (O.state_current_player(s) == 1) else -1.0)
            apl, apo = _orc.augment_example(t, pol)
            assert np.array_equal(np.array(ex[i].state, np.float32), t) and np.array_equal(np.array(ex[i].policy, np.float32), pol) and ex[i].value == z
            for k in range(7):
                assert np.array_equal(np.array(ex[i + 1 + k].state, np.float32), apl[k]), (i, k)
                assert np.array_equal(np.array(ex[i + 1 + k].policy, np.float32), apo[k]), (i, k)
                assert ex[i + 1 + k].value == z
            i += 8
            assert O.state_make_move(s, mv.action) == 0
    d.extractExamples(False)
    assert d.size() == n_moves
    d.setShuffleOnExtract(True); d.extractExamples(True)
    st, po, va = d.getBatch(5)
    assert len(st) == 5 and len(st[0]) == (11 if game == "GOMOKU" else 8) and len(po[0]) == len(games[0].getMoves()[0].policy)


def test_dataset_chess_no_augmentation_and_mixed_policy_lengths():
    import _alphazero_cpp as az
    O = _orc.oracle()
    r = az.GameRecord(az.GameType.CHESS, 8, False)
    s = O.new_state(_orc.CHESS, 8)
    tensors = []
    for ply in range(6):
        legal = O.legal(s)
        tensors.append(O.tensor(s))
        a = int(legal[(7 * ply + 3) % len(legal)])
        r.addMove(a, [1.0 / (ply + 1)] * (ply + 1), 0.0, 1)           # child-ordered vectors of different lengths
        assert O.state_make_move(s, a) == 0
    r.setResult(az.GameResult.WIN_PLAYER2)
    d = az.Dataset(); d.setShuffleOnExtract(False); d.addGameRecord(r); d.extractExamples(True)
    assert d.size() == 6                                             # dataset.cpp:250-253
    for ply, e in enumerate(d.examples()):
        assert np.array_equal(np.array(e.state, np.float32), tensors[ply])
        assert len(e.policy) == ply + 1 and e.value == (-1.0 if ply % 2 == 0 else 1.0)


def test_self_play_command_reference_flags_and_outputs(tmp_path):
    """The `self_play` command (src/selfplay/selfplay_main.cpp:87-117 flags, :352-388 metadata): GameRecord JSON per game + metadata JSON."""
    import json, subprocess
    exe = os.path.join(ROOT, "alphazero-multi-game_b200", "self_play")
    out = tmp_path / "games"
    r = subprocess.run([exe, "--model", "hash", "--game", "gomoku", "--size", "9", "--num-games", "3", "--simulations", "30", "--slots", "3", "--output-dir", str(out),
                        "--temperature", "1.0", "--temp-drop", "10", "--c-puct", "1.5", "--virtual-loss", "3", "--threads", "2", "--batch-size", "16"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Self-play completed!" in r.stdout and "Average moves per second" in r.stdout
    files = sorted(os.listdir(out))
    games = [f for f in files if not f.startswith("metadata_")]
    meta = [f for f in files if f.startswith("metadata_")]
    assert len(games) == 3 and len(meta) == 1
    g = json.loads((out / games[0]).read_text())
    assert g["game_type"] == 0 and g["board_size"] == 9 and g["result"] in (1, 2, 3) and len(g["moves"]) >= 9
    m = json.loads((out / meta[0]).read_text())
    assert list(m) == ["game", "board_size", "num_games_requested", "num_games_completed", "simulations", "threads", "temperature", "temp_drop", "final_temp",
                       "dirichlet_alpha", "dirichlet_epsilon", "variant", "model_path", "total_moves", "avg_moves_per_game", "total_time_seconds",
                       "avg_moves_per_second", "use_gpu", "batch_size", "batch_timeout", "fp16_used", "c_puct", "fpu_reduction", "virtual_loss",
                       "use_transposition_table", "progressive_widening"]
    assert m["num_games_completed"] == 3 and m["simulations"] == 30 and m["threads"] == 2 and m["use_gpu"] is True
    bad = subprocess.run([exe, "--model", "hash", "--variant"], capture_output=True, text=True)
    assert bad.returncode == 1 and "variant" in bad.stderr


def test_parallel_mcts_inspection_and_setters():
    """The ParallelMCTS members the reference binds beyond the search loop (python_bindings.cpp:245-253, 303-314): MCTSNode snapshots
    (getNode), printSearchPath, setConfig / setCPuct (take effect on the next search: the search then equals an oracle run with the new
    constant), setSelectionStrategy (PUCT only), setNeuralNetwork, setTranspositionTable / batch knobs (recorded, nothing to act on)."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 9)
    state = az.createGameState(az.GameType.GOMOKU, 9, False)
    mcts = az.ParallelMCTS(state, nn, az.TranspositionTable(1024, 4), 1, 150, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    mcts.setCPuct(2.25)                                   # before the first search
    mcts.search()
    om = O.mcts_new(O.new_state(_orc.GOMOKU, 9), 150, 2.25, 3, 0, None, None)
    O.mcts_search(om)
    b = O.root_stats(om)
    root = mcts.getNode()
    assert root.actions == b["actions"].tolist() and root.childVisits == b["N"].tolist() and root.visitCount == b["rootN"] and root.isExpanded
    assert np.array_equal(np.array(root.childValueSums, np.float32).view(np.uint32), b["W"].view(np.uint32))
    best = root.getBestAction()
    assert best == b["actions"][int(np.argmax(b["N"]))]
    child = mcts.getNode([best])
    assert child.visitCount == int(b["N"].max()) and abs(child.prior - float(b["P"][int(np.argmax(b["N"]))])) == 0.0
    assert sum(child.childVisits) == child.visitCount - 1 and "Node(V=" in child.toString()
    d = root.getVisitCountDistribution(1.0)
    assert abs(sum(d) - 1.0) < 1e-5 and np.allclose(d, mcts.getActionProbabilities(1.0))
    assert child.getUcbScore(1.5, 1, 0.0, root.visitCount) > child.getValue()
    mcts.printSearchPath(best); mcts.printSearchPath(-7)
    cfg = az.MCTSConfig(); cfg.numSimulations = 60; cfg.cPuct = 1.5; cfg.virtualLoss = 3
    mcts.setConfig(cfg); mcts.enableBatchedMCTS(True); mcts.setBatchSize(32); mcts.setBatchTimeout(5); mcts.setTranspositionTable(None); mcts.setNeuralNetwork(nn)
    with pytest.raises(RuntimeError, match="PUCT"):
        mcts.setSelectionStrategy(az.MCTSNodeSelection.RAVE)
    mcts.setSelectionStrategy(az.MCTSNodeSelection.PUCT)
    mcts.search()                                         # 60 more simulations with cPuct 1.5 on the same tree
    assert sum(mcts.getNode().childVisits) == 150 + 60


def test_selfplay_manager_over_two_gpus_nccl():
    """SelfPlayManager.setDevices([0, 1]): games sharded over two GPUs (one engine + host thread each), finished-game samples all-gathered
    with ncclAllGather, counters summed with ncclAllReduce — all from the C++ host layer, through the reference's own API.  Needs 2 GPUs
    (`gpurun --gpus 2`); deterministic mode, so the two devices' games are the same game (every slot plays the oracle's game) and each
    record can be checked against the single-GPU run."""
    import _alphazero_cpp as az
    import ctypes as C
    from _eng import E
    n = C.c_int32(0)
    E.load_library().az_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 9)
    one = az.SelfPlayManager(nn, 4, 48, 1); one.setConcurrentGames(4); one.setDeterministic(True)
    ref_games = one.generateGames(az.GameType.GOMOKU, 9, False)
    two = az.SelfPlayManager(nn, 8, 48, 1); two.setConcurrentGames(4); two.setDeterministic(True); two.setDevices([0, 1])
    games = two.generateGames(az.GameType.GOMOKU, 9, False)
    assert len(games) == 8 and two.getLastGatheredSampleBytes() > 0
    want = [m.action for m in ref_games[0].getMoves()]
    for g in games:
        assert [m.action for m in g.getMoves()] == want and g.getResult() == ref_games[0].getResult()
    st1, st2 = one.getLastRunStats(), two.getLastRunStats()
    assert st2[3] >= 8 and st2[0] == 2 * st1[0] and st2[2] == 2 * st1[2]          # games; simulations and moves: twice the one-GPU run (all-reduced)
    # throughput mode with the ResNet on both devices: runs, finishes, returns well-formed records
    from _eng import N
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "w.azw")
        open(p, "wb").write(N.export_weights(N.make_random_model(seed=0, blocks=1, board=9, actions=81)))
        net = az.createNeuralNetwork(p, az.GameType.GOMOKU, 9)
        mgr = az.SelfPlayManager(net, 6, 16, 1); mgr.setConcurrentGames(8); mgr.setDevices([0, 1])
        gs = mgr.generateGames(az.GameType.GOMOKU, 9, False)
        assert len(gs) == 6 and all(len(g.getMoves()) >= 5 and g.getResult() != az.GameResult.ONGOING for g in gs)


def test_python_self_play_driver_and_orchestrator(tmp_path):
    """scripts/self_play.py (the reference's python/scripts/self_play.py flow: createNeuralNetwork → SelfPlayManager → setMctsConfig(dict) → progress
    callback → GameRecord files + metadata JSON with the reference's keys) and scripts/orchestrate_selfplay.py (per-process output layout, summary)."""
    import json, subprocess, sys
    scripts = os.path.join(ROOT, "alphazero-multi-game_b200", "scripts")
    out = tmp_path / "py"
    r = subprocess.run([sys.executable, os.path.join(scripts, "self_play.py"), "--model", "hash", "--game", "gomoku", "--size", "9", "--num-games", "3", "--simulations", "24",
                        "--slots", "3", "--output-dir", str(out), "--threads", "2", "--batch-size", "16", "--seed", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Self-play finished." in r.stdout and "Progress:" in r.stdout
    files = sorted(os.listdir(out))
    meta = [f for f in files if f.startswith("metadata_")]
    assert len([f for f in files if not f.startswith("metadata_")]) == 3 and len(meta) == 1
    m = json.loads((out / meta[0]).read_text())
    ref_keys = ["timestamp", "game", "board_size", "num_games_requested", "num_games_completed", "simulations", "threads", "temperature", "temp_drop", "final_temp",
                "dirichlet_alpha", "dirichlet_epsilon", "variant", "model_path_arg", "total_moves", "avg_moves_per_game", "total_time_seconds", "avg_time_per_game_seconds",
                "avg_moves_per_second", "use_gpu", "fp16_used", "batch_size_used", "batch_timeout_used", "seed", "nn_loaded", "nn_avg_inference_ms", "nn_device_info",
                "nn_batch_size", "nn_batch_timeout", "nn_fp16_enabled", "fpu_reduction", "c_puct", "virtual_loss", "use_transposition_table", "progressive_widening"]
    assert list(m)[:len(ref_keys)] == ref_keys and m["num_games_completed"] == 3
    out2 = tmp_path / "orch"
    r = subprocess.run([sys.executable, os.path.join(scripts, "orchestrate_selfplay.py"), "--model", "hash", "--game", "gomoku", "--size", "9", "--num-games", "2",
                        "--simulations", "16", "--processes", "1", "--output-dir", str(out2), "--monitor-interval", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert os.path.isdir(out2 / "proc_0") and any(f.startswith("orchestration_summary_") for f in os.listdir(out2))
    s = json.loads(r.stdout.strip().splitlines()[-1])
    assert s["total_games"] == 2 and s["return_code"] == 0


@pytest.mark.parametrize("game,board,moves", [(_orc.GOMOKU, 9, 4), (_orc.GO, 9, 3)])
def test_parallel_mcts_with_an_arbitrary_host_evaluator(game, board, moves):
    """ParallelMCTS accepts ANY nn::NeuralNetwork (neural_network.h:19-131), not only the device evaluators: a network that is not a
    B200NeuralNetwork runs through AZ_EVAL_EXTERNAL — every wave's leaves go to the host as move sequences, the states are rebuilt from the
    root clone and evaluated with predictBatch, policies / values go back.  "hash-host" is the hash evaluator computed by a plain host
    NeuralNetwork: the search must equal the oracle (== the device hash evaluator) bit for bit, incl. subtree reuse across moves."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    gt = az.GameType.GOMOKU if game == _orc.GOMOKU else az.GameType.GO
    nn = az.createNeuralNetwork("hash-host", gt, board)
    assert "host hash evaluator" in nn.getDeviceInfo() and not nn.isGpuAvailable()
    state = az.createGameState(gt, board, False)
    o_state = O.new_state(game, board)
    for a in ([40, 41] if game == _orc.GOMOKU else [30]):
        state.makeMove(a); assert O.state_make_move(o_state, a) == 0
    mcts = az.ParallelMCTS(state, nn, None, 1, 90, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    om = O.mcts_new(o_state, 90, 1.5, 3, 0, None, None)
    for mv in range(moves):
        mcts.search(); O.mcts_search(om)
        actions, visits, wsum, priors, root_n, root_w = mcts.getRootChildren()
        b = O.root_stats(om)
        assert actions == b["actions"].tolist() and visits == b["N"].tolist() and root_n == b["rootN"], mv
        assert np.array_equal(np.array(wsum, np.float32).view(np.uint32), b["W"].view(np.uint32)), mv
        assert np.array_equal(np.array(priors, np.float32).view(np.uint32), b["P"].view(np.uint32)), mv
        a = mcts.selectAction(True, 1.0)
        assert a == O.mcts_select_action(om, 1, 1.0)
        mcts.updateWithMove(a); O.mcts_update_with_move(om, a)
    assert "evaluations" in nn.getDeviceInfo()


def test_transposition_table_object_is_the_device_evaluation_cache():
    """mcts::TranspositionTable (python_bindings.cpp:255-266) on the B200 engine: the table lives on the device (az_config.eval_cache_entries),
    the Python object carries the requested size into the engine and collects lookups / hits.  Searching a Go position twice from fresh
    ParallelMCTS objects is NOT served across engines (one device table per engine, as the reference keeps one table per ParallelMCTS,
    self_play_manager.cpp:159,175), but inside one engine a second search of the same root after set-up hits; results equal the oracle's."""
    import _alphazero_cpp as az
    O = _orc.oracle()
    nn = az.createNeuralNetwork("hash", az.GameType.GO, 9)
    tt = az.TranspositionTable(4096, 4)
    state = az.createGameState(az.GameType.GO, 9, False)
    mcts = az.ParallelMCTS(state, nn, tt, 1, 200, 1.5, 0.0, 3)
    mcts.setDeterministicMode(True)
    om = O.mcts_new(O.new_state(_orc.GO, 9), 200, 1.5, 3, 0, None, None)
    for mv in range(3):
        mcts.search(); O.mcts_search(om)
        actions, visits, wsum, priors, root_n, root_w = mcts.getRootChildren()
        b = O.root_stats(om)
        assert actions == b["actions"].tolist() and visits == b["N"].tolist()
        assert np.array_equal(np.array(wsum, np.float32).view(np.uint32), b["W"].view(np.uint32))
        a = mcts.selectAction(True, 1.0)
        assert a == O.mcts_select_action(om, 1, 1.0)
        state.makeMove(a); mcts.updateWithMove(a); O.mcts_update_with_move(om, a)
    assert tt.getSize() == 4096 and 3 * 150 <= tt.getLookups() <= 3 * 201      # one lookup per evaluated leaf (terminal leaves need none)
    assert 0 <= tt.getHits() <= tt.getLookups() and 0.0 <= tt.getHitRate() <= 1.0
    assert tt.getEntryCount() == min(tt.getLookups() - tt.getHits(), 4096)
    tt.clear()
    assert tt.getLookups() == 0
