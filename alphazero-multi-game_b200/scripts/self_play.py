#!/usr/bin/env python3
"""self_play.py — the reference's Python self-play driver (python/scripts/self_play.py:76-136 arguments, :280-520 flow and metadata JSON)
on the B200 engine: same options, same console flow, one GameRecord JSON per game in --output-dir and metadata_<time>.json with the
reference's keys, through the same pybind calls (createGameState / createNeuralNetwork / SelfPlayManager.setBatchConfig /
setExplorationParams / setSaveGames / setMctsConfig(dict) / setProgressCallback / generateGames).

Differences (all visible in --help): --model takes an AZW1 weight blob (alphazero-multi-game_b200/net.py:export_weights) or the word
`hash`; --create-random-model builds the plain 10-block 128-channel ResNet of BASELINE.json (the reference builds a DDW-RandWire net, out
of scope) and exports it as AZW1; --no-gpu and --variant are errors (no CPU path, no variant rules); --threads / --batch-size /
--batch-timeout / --no-batched-search are accepted and recorded but have nothing to act on (every wave is one batch on the device);
--fp16 is the engine's default storage type anyway; new: --gpus N (games sharded over N GPUs, NCCL sample gather), --slots, --deterministic.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.dirname(PKG))


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="AlphaZero Self-Play (B200 engine)")
    p.add_argument("--model", type=str, default="", help="AZW1 weight blob, or `hash` for the deterministic test evaluator")
    p.add_argument("--game", type=str, default="gomoku", choices=["gomoku", "chess", "go"], help="Game type")
    p.add_argument("--size", type=int, default=0, help="Board size (0 for default)")
    p.add_argument("--num-games", type=int, default=100, help="Number of games to generate")
    p.add_argument("--simulations", type=int, default=800, help="Number of MCTS simulations per move")
    p.add_argument("--threads", type=int, default=0, help="(recorded only)")
    p.add_argument("--output-dir", type=str, default="data/games", help="Output directory for game records")
    p.add_argument("--temperature", type=float, default=1.0, help="Initial temperature")
    p.add_argument("--temp-drop", type=int, default=30, help="Move number to drop temperature")
    p.add_argument("--final-temp", type=float, default=0.0, help="Final temperature")
    p.add_argument("--dirichlet-alpha", type=float, default=0.03, help="Dirichlet noise alpha")
    p.add_argument("--dirichlet-epsilon", type=float, default=0.25, help="Dirichlet noise weight")
    p.add_argument("--variant", action="store_true", help="(error: variant rules are not built)")
    p.add_argument("--seed", type=int, default=None, help="Random seed")
    p.add_argument("--batch-size", type=int, default=0, help="(recorded only)")
    p.add_argument("--batch-timeout", type=int, default=0, help="(recorded only)")
    p.add_argument("--no-gpu", action="store_true", help="(error: the engine has no CPU path)")
    p.add_argument("--no-batched-search", action="store_true", help="(recorded only)")
    p.add_argument("--fp16", action="store_true", help="fp16 network storage (the engine's default)")
    p.add_argument("--create-random-model", action="store_true", help="Create and export a random-init 10 x 128 ResNet if no model is given")
    p.add_argument("--fpu-reduction", type=float, default=0.1, help="(recorded only: the FPU branch is unreachable in the serial search)")
    p.add_argument("--c-puct", type=float, default=1.5, help="PUCT exploration constant")
    p.add_argument("--virtual-loss", type=int, default=3, help="Virtual loss amount")
    p.add_argument("--use-transposition-table", action="store_true", default=True, help="(in-wave evaluation sharing is always on)")
    p.add_argument("--progressive-widening", action="store_true", help="(recorded only)")
    p.add_argument("--profile", action="store_true", help="cProfile the driver")
    p.add_argument("--gpus", type=int, default=1, help="Shard the games over GPUs 0..N-1 (NCCL sample all-gather)")
    p.add_argument("--slots", type=int, default=0, help="Concurrent games per GPU (default min(num-games, 4096))")
    p.add_argument("--deterministic", action="store_true", help="Noise off, first max-visit move (parity runs)")
    return p.parse_args(argv)


def create_neural_network(az, args, game_type, board_size):
    state = az.createGameState(game_type, board_size, False)
    input_channels = len(state.getEnhancedTensorRepresentation())
    action_size = state.getActionSpaceSize()
    path = args.model
    if args.create_random_model and not path:
        import az_b200_loader
        az_b200_loader.load()
        from alphazero_multi_game_b200 import net as N
        print("Creating a random-init 10-block 128-channel policy/value ResNet...")
        model = N.make_random_model(seed=args.seed or 0, in_planes=input_channels, board=board_size, actions=action_size, blocks=10, channels=128)
        os.makedirs("models", exist_ok=True)
        path = os.path.join("models", f"random_model_{args.game}_{board_size}x{board_size}.azw")
        open(path, "wb").write(N.export_weights(model))
        print(f"Random model exported to {path}")
    if not path:
        raise SystemExit("Error: --model PATH | hash (or --create-random-model) is required: the engine has no random-policy CPU evaluator")
    print(f"Attempting to load model with C++ API: {path}")
    nn = az.createNeuralNetwork(path, game_type, board_size, True)
    print(f"Neural network loaded into C++ API: {nn.getDeviceInfo()}")
    print(f"C++ NN Batch size: {nn.getBatchSize()}")
    return nn


def run_self_play(args):
    import _alphazero_cpp as az
    if args.no_gpu:
        raise SystemExit("Error: --no-gpu: the B200 engine has no CPU path")
    if args.variant:
        raise SystemExit("Error: --variant: variant rules (Renju, Chess960) are not built")
    game_type = {"gomoku": az.GameType.GOMOKU, "chess": az.GameType.CHESS, "go": az.GameType.GO}[args.game]
    board_size = args.size if args.size > 0 else {"gomoku": 15, "chess": 8, "go": 19}[args.game]
    os.makedirs(args.output_dir, exist_ok=True)
    print("Initializing Neural Network...")
    nn = create_neural_network(az, args, game_type, board_size)
    print("Initializing Self-Play Manager...")
    sp = az.SelfPlayManager(nn, args.num_games, args.simulations, max(1, args.threads))
    sp.setBatchConfig(args.batch_size or 64, args.batch_timeout or 10)
    sp.setExplorationParams(args.dirichlet_alpha, args.dirichlet_epsilon, args.temperature, args.temp_drop, args.final_temp)
    sp.setSaveGames(True, args.output_dir)
    sp.setMctsConfig({"useBatchedMCTS": not args.no_batched_search, "batchSize": args.batch_size or 64, "batchTimeoutMs": args.batch_timeout or 10,
                      "searchMode": "BATCHED" if not args.no_batched_search else "PARALLEL", "fpuReduction": args.fpu_reduction, "cPuct": args.c_puct,
                      "virtualLoss": args.virtual_loss, "useFmapCache": args.use_transposition_table, "useTemporalDifference": False,
                      "useProgressiveWidening": args.progressive_widening})
    print("Advanced MCTS configuration applied.")
    if args.slots:
        sp.setConcurrentGames(args.slots)
    if args.deterministic:
        sp.setDeterministic(True)
    if args.gpus > 1:
        sp.setDevices(list(range(args.gpus)))
    print("Starting self-play generation...")
    print("-" * 40)
    print(f"Game:               {args.game.upper()}\nBoard size:         {board_size}x{board_size}\nVariant rules:      {args.variant}\n"
          f"Number of games:    {args.num_games}\nSimulations/move:   {args.simulations}\nGPUs:               {args.gpus}\n"
          f"Output directory:   {args.output_dir}\nModel path:         {args.model or 'random-init (exported)'}\nNN Device Info:     {nn.getDeviceInfo()}")
    print("-" * 40)
    t0 = time.time()
    state = dict(last=t0, games=0, moves=0)

    def progress(game_id, move_num, total_games, total_moves):
        now = time.time(); dt = now - state["last"]
        if dt > 0 and game_id > state["games"]:
            print(f"Progress: {game_id}/{total_games} games | {total_moves} moves | {(game_id - state['games']) / dt:.2f} games/sec | "
                  f"{(total_moves - state['moves']) / dt:.1f} moves/sec")
        else:
            print(f"Progress: {game_id}/{total_games} games | {total_moves} moves")
        state.update(last=now, games=game_id, moves=total_moves)

    sp.setProgressCallback(progress)
    games = sp.generateGames(game_type, board_size, False)
    dur = time.time() - t0
    n_games, n_moves = len(games), sp.getTotalMovesCount()
    print("--- Self-Play Results ---")
    print(f"Completed {n_games} games in {dur:.2f} seconds")
    if n_games == 0:
        print("No complete games were generated.")
        return 0
    print(f"Total moves:        {n_moves}\nAvg moves/game:     {n_moves / n_games:.1f}\nAvg time/game:      {dur / n_games:.2f} seconds\nAvg moves/second:   {n_moves / dur:.1f}")
    meta = {   # same keys, same order as python/scripts/self_play.py:470-508
        "timestamp": time.strftime("%Y-%m-%d %H:%M:%S"), "game": args.game, "board_size": board_size, "num_games_requested": args.num_games,
        "num_games_completed": n_games, "simulations": args.simulations, "threads": args.threads, "temperature": args.temperature,
        "temp_drop": args.temp_drop, "final_temp": args.final_temp, "dirichlet_alpha": args.dirichlet_alpha, "dirichlet_epsilon": args.dirichlet_epsilon,
        "variant": args.variant, "model_path_arg": args.model, "total_moves": n_moves, "avg_moves_per_game": n_moves / n_games,
        "total_time_seconds": dur, "avg_time_per_game_seconds": dur / n_games, "avg_moves_per_second": n_moves / dur, "use_gpu": True,
        "fp16_used": True, "batch_size_used": nn.getBatchSize(), "batch_timeout_used": args.batch_timeout, "seed": args.seed, "nn_loaded": True,
        "nn_avg_inference_ms": nn.getInferenceTimeMs(), "nn_device_info": nn.getDeviceInfo(), "nn_batch_size": nn.getBatchSize(),
        "nn_batch_timeout": args.batch_timeout, "nn_fp16_enabled": True, "fpu_reduction": args.fpu_reduction, "c_puct": args.c_puct,
        "virtual_loss": args.virtual_loss, "use_transposition_table": args.use_transposition_table, "progressive_widening": args.progressive_widening,
        "gpus": args.gpus, "simulations_total": (sp.getLastRunStats() or [0])[0]}
    path = os.path.join(args.output_dir, f"metadata_{time.strftime('%Y%m%d_%H%M%S')}.json")
    json.dump(meta, open(path, "w"), indent=2)
    print(f"Metadata saved to {path}\nSelf-play finished.")
    return n_games


if __name__ == "__main__":
    a = parse_args()
    print("Checking C++ module availability...")
    import _alphazero_cpp  # noqa: F401
    print("C++ module loaded successfully.")
    if a.profile:
        import cProfile
        cProfile.run("run_self_play(a)", f"selfplay_profile_{time.strftime('%Y%m%d_%H%M%S')}.prof")
    else:
        run_self_play(a)
