// selfplay_main.cpp — the reference's `self_play` command (src/selfplay/selfplay_main.cpp:87-117 flags, :352-388 metadata JSON)
// on the B200 engine: same options, same output layout (one GameRecord JSON per game in --output-dir + metadata_<ticks>.json), so
// pipelines that shell out to `self_play` (scripts/run_alphazero_pipeline.sh) run unchanged.  Differences, all stated in --help:
// --model takes an AZW1 weight blob (net.py:export_weights) or the word `hash` (the deterministic test evaluator); --threads,
// --batch-size, --batch-timeout, --no-batched-search, --fp16, --use-tt and --progressive-widening are accepted and recorded but have
// no effect (the engine batches every wave on the device in bf16); --no-gpu and --variant are errors (no CPU path, no variant rules);
// --slots = concurrent games on the GPU.
#include "alphazero_host.hpp"

#include <chrono>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <unordered_map>

using namespace alphazero;

namespace {
struct Args {
    std::unordered_map<std::string, std::string> kv;
    Args(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (a.rfind("--", 0) != 0) continue;
            a = a.substr(2);
            if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) kv[a] = argv[++i]; else kv[a] = "true";
        }
    }
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string str(const std::string& k, const std::string& d) const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    int num(const std::string& k, int d) const { try { return has(k) ? std::stoi(kv.at(k)) : d; } catch (...) { return d; } }
    float flt(const std::string& k, float d) const { try { return has(k) ? std::stof(kv.at(k)) : d; } catch (...) { return d; } }
    bool flag(const std::string& k, bool d) const {
        if (!has(k)) return d;
        const std::string& v = kv.at(k);
        if (v == "true" || v == "yes" || v == "1") return true;
        if (v == "false" || v == "no" || v == "0") return false;
        return d;
    }
};

void usage() {
    std::cout << "AlphaZero Multi-Game Self-Play Generator (B200 engine)\n"
                 "Usage: self_play [options]\n\n"
                 "Options:\n"
                 "  --model PATH          AZW1 weight blob (net.py:export_weights), or `hash` for the deterministic test evaluator\n"
                 "  --game TYPE           Game type: gomoku, chess, go (default: gomoku)\n"
                 "  --size SIZE           Board size (default: depends on game)\n"
                 "  --num-games NUM       Number of games to generate (default: 100)\n"
                 "  --simulations SIMS    Number of MCTS simulations per move (default: 800)\n"
                 "  --slots N             Concurrent games per GPU (default: min(num-games, 4096))\n"
                 "  --gpus N              Shard the games over GPUs 0..N-1 of this node (sample all-gather + counter all-reduce over NCCL)\n"
                 "  --output-dir DIR      Output directory (default: data/games)\n"
                 "  --temperature TEMP    Initial temperature (default: 1.0)\n"
                 "  --temp-drop MOVE      Move to drop temperature (default: 30)\n"
                 "  --final-temp TEMP     Final temperature (default: 0.0)\n"
                 "  --dirichlet-alpha A   Dirichlet noise alpha (default: 0.03)\n"
                 "  --dirichlet-epsilon E Dirichlet noise weight (default: 0.25)\n"
                 "  --c-puct VALUE        Exploration constant (default: 1.5)\n"
                 "  --virtual-loss VALUE  Virtual loss amount (default: 3)\n"
                 "  --deterministic       Noise off, first max-visit move (parity runs)\n"
                 "  accepted for compatibility, no effect: --threads --batch-size --batch-timeout --no-batched-search --fp16\n"
                 "                                         --fpu-reduction --progressive-widening\n"
                 "  --no-tt               No transposition table = no device evaluation cache (results are identical either way)\n"
                 "  not supported (error): --no-gpu, --variant\n"
                 "  --help                Display this help message\n";
}
}  // namespace

int main(int argc, char** argv) {
    Args a(argc, argv);
    if (a.has("help")) { usage(); return 0; }
    try {
        const std::string modelPath = a.str("model", ""), gameStr = a.str("game", "gomoku");
        if (modelPath.empty()) { std::cerr << "Error: Model path is required (--model PATH | hash)\n"; usage(); return 1; }
        if (a.flag("no-gpu", false)) { std::cerr << "Error: --no-gpu: the B200 engine has no CPU path\n"; return 1; }
        if (a.flag("variant", false)) { std::cerr << "Error: --variant: variant rules (Renju, Chess960) are not built\n"; return 1; }
        const core::GameType gt = gameStr == "chess" ? core::GameType::CHESS : (gameStr == "go" ? core::GameType::GO : core::GameType::GOMOKU);
        int boardSize = a.num("size", 0);
        if (boardSize <= 0) boardSize = gt == core::GameType::CHESS ? 8 : (gt == core::GameType::GO ? 19 : 15);
        const int numGames = a.num("num-games", 100), sims = a.num("simulations", 800), threads = a.num("threads", 0);
        const std::string outputDir = a.str("output-dir", "data/games");
        const float temperature = a.flt("temperature", 1.0f), finalTemp = a.flt("final-temp", 0.0f);
        const int tempDrop = a.num("temp-drop", 30);
        const float alpha = a.flt("dirichlet-alpha", 0.03f), eps = a.flt("dirichlet-epsilon", 0.25f);
        const int batchSize = a.num("batch-size", 8), batchTimeout = a.num("batch-timeout", 10);
        const float cPuct = a.flt("c-puct", 1.5f), fpu = a.flt("fpu-reduction", 0.1f);
        const int virtualLoss = a.num("virtual-loss", 3);
        const bool useTT = !a.has("no-tt") && a.flag("use-tt", true), pw = a.flag("progressive-widening", false), fp16 = a.flag("fp16", false);

        std::filesystem::create_directories(outputDir);
        std::cout << "Loading model from " << modelPath << std::endl;
        auto nn = nn::NeuralNetwork::create(modelPath, gt, boardSize, true);
        std::cout << nn->getDeviceInfo() << std::endl;

        selfplay::SelfPlayManager sp(nn.get(), numGames, sims, threads > 0 ? threads : 1);
        sp.setExplorationParams(alpha, eps, temperature, tempDrop, finalTemp);
        sp.setBatchConfig(batchSize, batchTimeout);
        sp.setSaveGames(true, outputDir);
        if (a.has("slots")) sp.setConcurrentGames(a.num("slots", 0));
        const int gpus = a.num("gpus", 1);
        if (gpus > 1) { std::vector<int> dv; for (int d = 0; d < gpus; ++d) dv.push_back(d); sp.setDevices(dv); }
        if (a.flag("deterministic", false)) sp.setDeterministic(true);
        mcts::MCTSConfig mc; mc.numSimulations = sims; mc.cPuct = cPuct; mc.fpuReduction = fpu; mc.virtualLoss = virtualLoss;
        mc.useDirichletNoise = true; mc.dirichletAlpha = alpha; mc.dirichletEpsilon = eps; mc.useProgressiveWidening = pw;
        if (!useTT) mc.transpositionTableSize = 0;      // --no-tt (what the reference's main tests, selfplay_main.cpp:188): no evaluation cache
        sp.setMctsConfig(mc);
        sp.setProgressCallback([](int gameId, int moves, int totalGames, int totalMoves) {
            std::cout << "Game " << (gameId + 1) << "/" << totalGames << " finished after " << moves << " moves (total moves played: " << totalMoves << ")" << std::endl;
        });

        std::cout << "Generating " << numGames << " " << gameStr << " games, " << sims << " simulations per move" << std::endl;
        const auto t0 = std::chrono::high_resolution_clock::now();
        auto records = sp.generateGames(gt, boardSize, false);
        const double secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        const long long duration = (long long)secs;

        const int totalMoves = sp.getTotalMovesCount();
        const float movesPerGame = records.empty() ? 0.0f : (float)totalMoves / records.size();
        const float movesPerSecond = secs > 0 ? (float)(totalMoves / secs) : 0.0f;
        std::cout << "\nSelf-play completed!\nGenerated " << records.size() << " games in " << duration << " seconds\nTotal moves: " << totalMoves
                  << "\nAverage moves per game: " << std::fixed << std::setprecision(1) << movesPerGame
                  << "\nAverage moves per second: " << movesPerSecond << std::endl;

        const std::string metaPath = outputDir + "/metadata_" + std::to_string(std::chrono::system_clock::now().time_since_epoch().count()) + ".json";
        std::ofstream m(metaPath);
        if (m.is_open()) {          // same keys, same order as selfplay_main.cpp:357-384
            auto b = [](bool v) { return v ? "true" : "false"; };
            m << "{\n  \"game\": \"" << gameStr << "\",\n  \"board_size\": " << boardSize << ",\n  \"num_games_requested\": " << numGames
              << ",\n  \"num_games_completed\": " << records.size() << ",\n  \"simulations\": " << sims << ",\n  \"threads\": " << threads
              << ",\n  \"temperature\": " << temperature << ",\n  \"temp_drop\": " << tempDrop << ",\n  \"final_temp\": " << finalTemp
              << ",\n  \"dirichlet_alpha\": " << alpha << ",\n  \"dirichlet_epsilon\": " << eps << ",\n  \"variant\": false,\n  \"model_path\": \"" << modelPath
              << "\",\n  \"total_moves\": " << totalMoves << ",\n  \"avg_moves_per_game\": " << movesPerGame << ",\n  \"total_time_seconds\": " << duration
              << ",\n  \"avg_moves_per_second\": " << movesPerSecond << ",\n  \"use_gpu\": true,\n  \"batch_size\": " << batchSize
              << ",\n  \"batch_timeout\": " << batchTimeout << ",\n  \"fp16_used\": " << b(fp16) << ",\n  \"c_puct\": " << cPuct << ",\n  \"fpu_reduction\": " << fpu
              << ",\n  \"virtual_loss\": " << virtualLoss << ",\n  \"use_transposition_table\": " << b(useTT) << ",\n  \"progressive_widening\": " << b(pw) << "\n}\n";
            std::cout << "Metadata saved to " << metaPath << std::endl;
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
