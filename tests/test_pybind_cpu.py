"""CPU suite, part 3: the reference-shaped pybind module `_alphazero_cpp` (host side above the C ABI): same names,
argument meaning and error behaviour as src/pybind/python_bindings.cpp for the self-play path."""
import json
import os
import sys

import numpy as np
import pytest

import _orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alphazero-multi-game_b200")


def _mod():
    sys.path.insert(0, PKG)
    try:
        import _alphazero_cpp as az
    except ImportError:
        import subprocess
        subprocess.check_call(["bash", os.path.join(PKG, "host", "build.sh")])
        import _alphazero_cpp as az
    return az


def test_module_surface_matches_reference_names():
    az = _mod()
    for name in ["GameType", "GameResult", "MCTSNodeSelection", "MCTSSearchMode", "IGameState", "GomokuState", "createGameState",
                 "NeuralNetwork", "createNeuralNetwork", "MCTSConfig", "MCTSStats", "TranspositionTable", "ParallelMCTS", "MoveData",
                 "GameRecord", "TrainingExample", "Dataset", "SelfPlayManager"]:
        assert hasattr(az, name), name
    for meth in ["search", "selectAction", "getActionProbabilities", "getRootValue", "updateWithMove", "addDirichletNoise", "setNumSimulations",
                 "setDeterministicMode", "getSearchInfo", "getMemoryUsage"]:
        assert hasattr(az.ParallelMCTS, meth), meth
    for meth in ["generateGames", "setExplorationParams", "setProgressCallback", "setBatchConfig", "setSaveGames", "setAbort", "isRunning",
                 "setMctsConfig", "getCompletedGamesCount", "getTotalMovesCount"]:
        assert hasattr(az.SelfPlayManager, meth), meth
    for meth in ["addGameRecord", "extractExamples", "size", "getBatch", "shuffle", "saveToFile", "loadFromFile", "getRandomSubset"]:
        assert hasattr(az.Dataset, meth), meth
    c = az.MCTSConfig()
    assert (c.numSimulations, c.virtualLoss) == (800, 3) and abs(c.cPuct - 1.5) < 1e-7 and c.fpuReduction == 0.0
    assert int(az.GameType.GO) == 2 and int(az.GameResult.WIN_PLAYER2) == 3


def test_gomoku_state_matches_oracle_including_legal_order():
    az = _mod()
    O = _orc.oracle()
    rng = np.random.default_rng(5)
    for n in (15, 9):
        for g in range(6):
            s = az.GomokuState(n); o = O.new_state(_orc.GOMOKU, n)
            enumerate_always = g % 2 == 0
            for ply in range(n * n + 1):
                if enumerate_always or ply % 7 == 3:
                    assert s.getLegalMoves() == O.legal(o).tolist(), (n, g, ply)    # incl. first-fill order (QUIRK G2)
                assert s.isTerminal() == bool(O.state_is_terminal(o))
                assert int(s.getGameResult()) == O.state_result(o)
                assert s.getCurrentPlayer() == O.state_current_player(o)
                if ply % 5 == 0:
                    assert np.array_equal(np.array(s.getEnhancedTensorRepresentation(), np.float32), O.tensor(o))
                if s.isTerminal():
                    break
                a = int(rng.choice(O.legal(o)))
                if not enumerate_always:
                    pass
                s.makeMove(a); O.state_make_move(o, a)
    s = az.GomokuState(15)
    with pytest.raises(RuntimeError):
        s.makeMove(225)
    s.makeMove(112)
    with pytest.raises(RuntimeError):
        s.makeMove(112)
    assert s.is_occupied(112) and not s.is_occupied(0) and s.get_board()[7][7] == 1
    assert s.actionToString(112) == "H8" and s.stringToAction("H8") == 112
    assert az.createGameState(az.GameType.GOMOKU).getBoardSize() == 15


def test_chess_state_matches_oracle():
    """createGameState(CHESS): the host-side ChessState (csrc/chess.cuh compiled for the host) against the oracle's chess
    restatement on random games: legal moves in order, terminal, result, player, 18 planes; illegal moves raise."""
    az = _mod()
    O = _orc.oracle()
    rng = np.random.default_rng(21)
    for g in range(4):
        s = az.createGameState(az.GameType.CHESS); o = O.new_state(_orc.CHESS, 8)
        assert s.getBoardSize() == 8 and s.getActionSpaceSize() == 20480
        for ply in range(140):
            lg = O.legal(o)
            assert s.getLegalMoves() == lg.tolist(), (g, ply)
            assert s.isTerminal() == bool(O.state_is_terminal(o)) and int(s.getGameResult()) == O.state_result(o)
            assert s.getCurrentPlayer() == O.state_current_player(o)
            if ply % 9 == 0:
                assert np.array_equal(np.array(s.getEnhancedTensorRepresentation(), np.float32), O.tensor(o))
            if s.isTerminal() or len(lg) == 0:
                break
            a = int(rng.choice(lg))
            s.makeMove(a); assert O.state_make_move(o, a) == 0
    s = az.createGameState(az.GameType.CHESS)
    assert len(s.getLegalMoves()) == 20 and s.actionToString(s.getLegalMoves()[1]) == "a2a4" and s.stringToAction("e2e4") == (52 << 6) | 36
    with pytest.raises(RuntimeError):
        s.makeMove((48 << 6) | 56)                 # a pawn moving backwards
    for mv in ("f2f3", "e7e5", "g2g4", "d8h4"):
        s.makeMove(s.stringToAction(mv))
    assert s.isTerminal() and int(s.getGameResult()) == 3 and s.getMoveHistory()[-1] == s.stringToAction("d8h4")      # Fool's mate: WIN_PLAYER2


def test_go_state_matches_oracle():
    """createGameState(GO): the host-side GoState (same rules header as the kernels, compiled for the host) against the
    oracle on random playouts with captures / ko / superko / passes: legal moves in order, terminal, result, planes."""
    az = _mod()
    O = _orc.oracle()
    rng = np.random.default_rng(9)
    assert az.createGameState(az.GameType.GO).getBoardSize() == 19
    for n, plies in ((9, 220), (13, 120)):
        for g in range(3):
            s = az.createGameState(az.GameType.GO, n, False); o = O.new_state(_orc.GO, n)
            assert s.getActionSpaceSize() == n * n + 1
            for ply in range(plies):
                lg = O.legal(o)
                if ply % 3 == 0 or ply > plies - 20:
                    assert s.getLegalMoves() == lg.tolist(), (n, g, ply)
                assert s.isTerminal() == bool(O.state_is_terminal(o))
                assert int(s.getGameResult()) == O.state_result(o)
                assert s.getCurrentPlayer() == O.state_current_player(o)
                if ply % 11 == 0:
                    assert np.array_equal(np.array(s.getEnhancedTensorRepresentation(), np.float32), O.tensor(o))
                if s.isTerminal():
                    break
                cand = lg[lg >= 0] if (len(lg) > 1 and rng.random() < 0.96) else lg
                a = int(rng.choice(cand))
                s.makeMove(a); assert O.state_make_move(o, a) == 0
            assert s.getMoveHistory()[-1] == a
    s = az.createGameState(az.GameType.GO, 9, False)
    s.makeMove(40)
    with pytest.raises(RuntimeError):
        s.makeMove(40)
    with pytest.raises(RuntimeError):
        s.makeMove(81)
    assert s.actionToString(-1) == "pass" and s.stringToAction("pass") == -1 and s.stringToAction(s.actionToString(40)) == 40
    assert not s.isTerminal()
    c = az.createGameState(az.GameType.GO, 9, False); c.makeMove(40); c.makeMove(-1); c.makeMove(-1)
    assert c.isTerminal() and c.undoMove() and not c.isTerminal() and c.getMoveHistory() == [40, -1]
    c.makeMove(-1)
    assert c.isTerminal() and int(c.getGameResult()) == 2      # one black stone owns the whole board: 81 > 7.5


def test_game_record_json_roundtrip_reference_format():
    az = _mod()
    r = az.GameRecord(az.GameType.GOMOKU, 15, False)
    r.addMove(112, [0.25, 0.75], 0.5, 12)
    r.addMove(113, [1.0], -0.25, 7)
    r.setResult(az.GameResult.WIN_PLAYER1)
    j = json.loads(r.toJson())
    assert set(j) == {"game_type", "board_size", "use_variant_rules", "result", "timestamp", "moves"}      # game_record.cpp:64-90
    assert j["game_type"] == 0 and j["board_size"] == 15 and j["result"] == 2 and len(j["moves"]) == 2
    assert set(j["moves"][0]) == {"action", "policy", "value", "thinking_time_ms"}
    r2 = az.GameRecord.fromJson(r.toJson())
    assert [m.action for m in r2.getMoves()] == [112, 113] and r2.getResult() == az.GameResult.WIN_PLAYER1
    with pytest.raises(RuntimeError):
        az.GameRecord.fromJson("{not json")


def test_training_example_and_dataset_file_format(tmp_path):
    """TrainingExample JSON (dataset.cpp:14-54) and the Dataset file format {"examples": [{state, policy, value}]} (:144-216)."""
    az = _mod()
    e = az.TrainingExample()
    e.state = [[[0.0, 1.0], [0.5, 0.25]], [[1.0, 1.0], [0.0, 0.0]]]
    e.policy = [0.125, 0.875]
    e.value = -1.0
    j = json.loads(e.toJson())
    assert set(j) == {"state", "policy", "value"} and j["state"][0][1] == [0.5, 0.25] and j["value"] == -1.0
    e2 = az.TrainingExample.fromJson(e.toJson())
    assert e2.state == e.state and e2.policy == e.policy and e2.value == e.value
    fn = tmp_path / "ds.json"
    fn.write_text(json.dumps({"examples": [j, j, j]}))
    d = az.Dataset()
    assert d.size() == 0 and d.loadFromFile(str(fn)) and d.size() == 3
    st, po, va = d.getBatch(2)
    assert len(st) == 2 and po[0] == e.policy and va == [-1.0, -1.0]
    assert len(d.getRandomSubset(10)) == 3
    out = tmp_path / "out.json"
    assert d.saveToFile(str(out)) and json.loads(out.read_text()) == {"examples": [j, j, j]}
    assert not d.loadFromFile(str(tmp_path / "missing.json"))


def test_search_classes_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    az = _mod()
    nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 15)
    assert "Hash" in nn.getModelInfo()
    pol, v = nn.predict(az.GomokuState(15))        # the hash evaluator is host arithmetic (test evaluator), no engine needed
    O = _orc.oracle()
    po, vo = O.hash_policy_value(O.new_state(_orc.GOMOKU, 15))
    assert np.array_equal(np.array(pol, np.float32), po) and np.float32(v) == vo
    with pytest.raises(RuntimeError, match="no CUDA device"):
        az.ParallelMCTS(az.GomokuState(15), nn)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        az.SelfPlayManager(nn, 2, 10, 1).generateGames(az.GameType.GOMOKU, 15, False)
    d = az.Dataset()
    r = az.GameRecord(az.GameType.GOMOKU, 9, False); r.addMove(40, [1.0], 0.0, 1); r.setResult(az.GameResult.DRAW)
    d.addGameRecord(r)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        d.extractExamples(True)


def test_self_play_command_help_and_loud_failure_without_gpu():
    import subprocess, torch
    _mod()
    exe = os.path.join(PKG, "self_play")
    if not os.path.exists(exe):
        subprocess.check_call(["bash", os.path.join(PKG, "host", "build.sh")])
    h = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert h.returncode == 0
    for flag in ["--model", "--game", "--size", "--num-games", "--simulations", "--output-dir", "--temperature", "--temp-drop", "--final-temp",
                 "--dirichlet-alpha", "--dirichlet-epsilon", "--c-puct", "--virtual-loss", "--threads", "--batch-size"]:      # selfplay_main.cpp:87-117
        assert flag in h.stdout, flag
    assert subprocess.run([exe], capture_output=True, text=True).returncode == 1           # --model is required
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--model", "hash", "--num-games", "1", "--output-dir", "/tmp/az_sp_none"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_python_caller_clis_accept_the_reference_arguments():
    """scripts/self_play.py and scripts/orchestrate_selfplay.py take every option of the reference's drivers
    (python/scripts/self_play.py:76-136, python/scripts/orchestrate_selfplay.py:92-160), with the reference's defaults."""
    sys.path.insert(0, os.path.join(PKG, "scripts"))
    import self_play as sp
    import orchestrate_selfplay as orch
    a = sp.parse_args(["--model", "m.azw", "--game", "go", "--size", "9", "--num-games", "7", "--simulations", "50", "--threads", "3", "--output-dir", "o",
                       "--temperature", "0.8", "--temp-drop", "12", "--final-temp", "0.1", "--dirichlet-alpha", "0.3", "--dirichlet-epsilon", "0.2", "--seed", "5",
                       "--batch-size", "32", "--batch-timeout", "7", "--no-batched-search", "--fp16", "--create-random-model", "--fpu-reduction", "0.2",
                       "--c-puct", "2.0", "--virtual-loss", "2", "--use-transposition-table", "--progressive-widening", "--profile"])
    assert (a.game, a.size, a.num_games, a.simulations, a.temp_drop, a.c_puct, a.virtual_loss) == ("go", 9, 7, 50, 12, 2.0, 2)
    d = sp.parse_args([])
    assert (d.game, d.num_games, d.simulations, d.output_dir, d.temperature, d.temp_drop, d.final_temp, d.dirichlet_alpha, d.dirichlet_epsilon, d.fpu_reduction,
            d.c_puct, d.virtual_loss) == ("gomoku", 100, 800, "data/games", 1.0, 30, 0.0, 0.03, 0.25, 0.1, 1.5, 3)
    o = orch.parse_args(["--processes", "4", "--model", "x", "--monitor-interval", "2", "--use-cpp-binary", "--no-tt", "--optimize-batch", "--cache-size", "1024",
                         "--optimize-threads", "--compact-size", "3", "--pin-threads", "--batch-size", "8", "--batch-timeout", "5"])
    assert o.processes == 4 and o.use_cpp_binary and o.batch_size == 8
    od = orch.parse_args([])
    assert (od.processes, od.batch_size, od.batch_timeout, od.monitor_interval, od.cache_size) == (1, 16, 10, 5, 2097152)


@pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref not built")
def test_state_string_api_matches_reference_live():
    """The rest of the IGameState surface the pybind module binds (python_bindings.cpp:57-73) — actionToString, stringToAction (incl. the
    strings the reference accepts by accident: lower case, 'H 8', a Go prefix of 'e2e4'), toString (board printouts character for
    character, Go's finished-game block with scores, chess castling / e.p. / clocks / FEN), getTensorRepresentation (3 / 3 / 12 planes),
    getMoveHistory, undoMove — host mirror vs the reference's own GomokuState / GoState / ChessState on seeded random games.
    The reference runs in a child process (tests/_ref_state_api_child.py: its static libstdc++ must not meet numpy's in one process).
    Not reproduced, on purpose: the reference's chess undoMove does not restore the position (the undone move is refused afterwards,
    asserted below on the reference's own output), and ChessState::actionToString prints squares for out-of-range actions."""
    import hashlib
    import subprocess
    az = _mod()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_state_api_child.py"), _orc.ref_path()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    recs = json.loads(out.stdout)
    assert len(recs) == 10

    def mk(game, board):
        if game == 0:
            return az.GomokuState(board)
        return az.createGameState(az.GameType.GO if game == 2 else az.GameType.CHESS, board if game == 2 else 0, False)

    n_strings = n_undo = 0
    for rec in recs:
        g, b = rec["game"], rec["board"]
        s = mk(g, b)
        for text, want in rec["bad"].items():
            assert s.stringToAction(text) == want, (g, b, text)
        for a, want in rec["bad_a2s"].items():
            if g != 1:
                assert s.actionToString(int(a)) == want, (g, b, a)
        assert bool(s.undoMove()) == bool(rec["undo_on_fresh"])
        for i, st in enumerate(rec["steps"]):
            s.getLegalMoves()                              # the child enumerates at every ply (Gomoku's legal order depends on it, QUIRK G2)
            for a, x, y in zip(st["pick"], st["a2s"], st["s2a"]):
                assert s.actionToString(a) == x and s.stringToAction(x) == y == a, (g, b, i, a)
                n_strings += 1
            assert s.toString() == st["to_string"], (g, b, i)
            t = np.array(s.getTensorRepresentation(), np.float32)
            assert t.shape == (st["basic_c"], b, b) and hashlib.sha256(t.tobytes()).hexdigest()[:16] == st["basic_sha"], (g, b, i)
            s.makeMove(st["move"])
            h = s.getMoveHistory()
            assert len(h) == st["history_len"] and h[-3:] == st["history_tail"], (g, b, i)
            if "undo" in st:
                assert st["equals_clone"] == 1 and st["equals_after_undo"] == 0 and st["remake"] == 0
                assert bool(s.undoMove()) == bool(st["undo"])
                assert s.getLegalMoves() == st["legal_after_undo"] and s.getCurrentPlayer() == st["player_after_undo"], (g, b, i)
                s.makeMove(st["move"])
                n_undo += 1
            if "chess_undo" in st:
                assert st["chess_undo"] == 1 and st["chess_remake"] == -1          # the reference defect, pinned as such
        if "final_to_string" in rec:
            s.makeMove(-1); s.makeMove(-1)
            assert s.isTerminal() and s.toString() == rec["final_to_string"], (g, b)
    assert n_strings > 1500 and n_undo > 30


@pytest.mark.skipif(not _orc.have_ref(), reason="oracle/_ref not built")
def test_game_record_and_dataset_files_roundtrip_through_the_reference(tmp_path):
    """Files written by the host mirror (GameRecord::saveToFile, game_record.cpp:119-131; Dataset::saveToFile, dataset.cpp:151-188) are read
    by the reference's own loaders and written back by its own writers BYTE-identically (but for the time stamp, which the reference re-stamps on load), and the mirror reads the reference's files back to the
    same records — existing data directories stay readable in both directions.  (Reference in a child process, standard library only.)"""
    import random
    import re
    import subprocess
    az = _mod()
    rng = random.Random(1)

    def ref_roundtrip(fn, src, dst):
        code = "import ctypes as C, sys; lib = C.CDLL(sys.argv[1]); print(getattr(lib, sys.argv[2])(sys.argv[3].encode(), sys.argv[4].encode()))"
        out = subprocess.run([sys.executable, "-c", code, _orc.ref_path(), fn, str(src), str(dst)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-1000:]
        return int(out.stdout.strip())

    for gt, board, acts in ((az.GameType.GOMOKU, 15, 225), (az.GameType.GO, 9, 82), (az.GameType.CHESS, 8, 20480)):
        rec = az.GameRecord(gt, board, False)
        n = rng.randrange(3, 12)
        for _ in range(n):
            rec.addMove(rng.randrange(-1 if gt == az.GameType.GO else 0, acts - 1), [rng.random() for _ in range(rng.randrange(1, 9))], rng.random() * 2 - 1, rng.randrange(5000))
        rec.setResult(az.GameResult.WIN_PLAYER2)
        a, b = tmp_path / f"rec_{board}_a.json", tmp_path / f"rec_{board}_b.json"
        assert rec.saveToFile(str(a))
        assert ref_roundtrip("ref_game_record_file_roundtrip", a, b) == n
        # the reference's fromJson does not read the timestamp back (game_record.cpp:92-117): its record carries the time of the load
        stamp = re.compile(rb'"timestamp": "[^"]*"')
        assert stamp.sub(b'"timestamp": ""', a.read_bytes()) == stamp.sub(b'"timestamp": ""', b.read_bytes()) and len(stamp.findall(a.read_bytes())) == 1
        back = az.GameRecord.loadFromFile(str(b))
        assert [m.action for m in back.getMoves()] == [m.action for m in rec.getMoves()] and back.getResult() == rec.getResult()
        assert [m.policy for m in back.getMoves()] == [m.policy for m in rec.getMoves()]
    ex = [{"state": [[[float(rng.random() < 0.5) for _ in range(3)] for _ in range(3)] for _ in range(4)], "policy": [rng.random() for _ in range(10)], "value": v}
          for v in (-1.0, 0.0, 1.0, 0.5)]
    c, e, f = tmp_path / "c.json", tmp_path / "e.json", tmp_path / "f.json"
    c.write_text(json.dumps({"examples": ex}))
    ds = az.Dataset()
    assert ds.loadFromFile(str(c)) and ds.size() == 4 and ds.saveToFile(str(e))
    assert ref_roundtrip("ref_dataset_file_roundtrip", e, f) == 4
    assert e.read_bytes() == f.read_bytes()
    ds2 = az.Dataset()
    assert ds2.loadFromFile(str(f)) and ds2.size() == 4 and sorted(ds2.getBatch(4)[2]) == sorted(ds.getBatch(4)[2]) == [-1.0, 0.0, 0.5, 1.0]      # getBatch draws at random
    assert ref_roundtrip("ref_dataset_file_roundtrip", tmp_path / "missing.json", f) == -1


@pytest.mark.skipif(not os.path.isdir("/root/reference/python/alphazero"), reason="reference tree not present")
def test_reference_python_package_imports_against_this_module():
    """python/alphazero/__init__.py:12-28 of the reference does `from _alphazero_cpp import (GameType, ..., SelfPlayManager)` and then imports its
    own Python sub-packages: with this repo's module first on the path the reference's package imports unchanged (child process: it must
    not pick up a module imported earlier in this one)."""
    import subprocess
    code = ("import sys; sys.path.insert(0, sys.argv[1]); sys.path.insert(0, '/root/reference/python'); import alphazero, _alphazero_cpp; "
            "assert alphazero.SelfPlayManager is _alphazero_cpp.SelfPlayManager and alphazero.ParallelMCTS is _alphazero_cpp.ParallelMCTS; "
            "assert _alphazero_cpp.__file__.startswith(sys.argv[1]); print(alphazero.__version__)")
    _mod()
    out = subprocess.run([sys.executable, "-c", code, PKG], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "1.0.0", out.stderr[-1500:]
