import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-multi-game_b200"))
import _alphazero_cpp as az
nn = az.createNeuralNetwork("hash", az.GameType.GOMOKU, 15)
state = az.createGameState(az.GameType.GOMOKU, 15, False)
mcts = az.ParallelMCTS(state, nn, None, 1, 100, 1.5, 0.0, 3)
mcts.setDeterministicMode(True)
mcts.search(); print("searched", flush=True)
a = mcts.selectAction(True, 1.0); print("action", a, flush=True)
state.makeMove(a); print("state moved", flush=True)
mcts.updateWithMove(a); print("updated", flush=True)
