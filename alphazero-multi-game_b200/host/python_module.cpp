// python_module.cpp — pybind11 module `_alphazero_cpp`: the same Python-visible names as the reference's module
// (src/pybind/python_bindings.cpp:26-458) for the self-play path, bound to the B200 host classes (alphazero_host.hpp).
// Out-of-scope reference bindings (DDWRandWireResNet, Dataset, TrainingExample, MCTSNode) are not provided.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/functional.h>

#include "alphazero_host.hpp"

namespace py = pybind11;
using namespace alphazero;

PYBIND11_MODULE(_alphazero_cpp, m) {
    m.doc() = "AlphaZero Multi-Game AI Engine bindings — B200-native self-play engine (libaz_b200)";

    py::enum_<core::GameType>(m, "GameType").value("GOMOKU", core::GameType::GOMOKU).value("CHESS", core::GameType::CHESS).value("GO", core::GameType::GO).export_values();
    py::enum_<core::GameResult>(m, "GameResult").value("ONGOING", core::GameResult::ONGOING).value("DRAW", core::GameResult::DRAW)
        .value("WIN_PLAYER1", core::GameResult::WIN_PLAYER1).value("WIN_PLAYER2", core::GameResult::WIN_PLAYER2).export_values();
    py::enum_<mcts::MCTSNodeSelection>(m, "MCTSNodeSelection").value("UCB", mcts::MCTSNodeSelection::UCB).value("PUCT", mcts::MCTSNodeSelection::PUCT)
        .value("PROGRESSIVE_BIAS", mcts::MCTSNodeSelection::PROGRESSIVE_BIAS).value("RAVE", mcts::MCTSNodeSelection::RAVE).export_values();
    py::enum_<mcts::MCTSSearchMode>(m, "MCTSSearchMode").value("SERIAL", mcts::MCTSSearchMode::SERIAL).value("PARALLEL", mcts::MCTSSearchMode::PARALLEL)
        .value("BATCHED", mcts::MCTSSearchMode::BATCHED).export_values();

    py::class_<core::IGameState>(m, "IGameState")
        .def("getLegalMoves", &core::IGameState::getLegalMoves).def("isLegalMove", &core::IGameState::isLegalMove)
        .def("makeMove", &core::IGameState::makeMove).def("undoMove", &core::IGameState::undoMove)
        .def("isTerminal", &core::IGameState::isTerminal).def("getGameResult", &core::IGameState::getGameResult)
        .def("getCurrentPlayer", &core::IGameState::getCurrentPlayer).def("getBoardSize", &core::IGameState::getBoardSize)
        .def("getActionSpaceSize", &core::IGameState::getActionSpaceSize)
        .def("getTensorRepresentation", &core::IGameState::getTensorRepresentation)
        .def("getEnhancedTensorRepresentation", &core::IGameState::getEnhancedTensorRepresentation)
        .def("actionToString", &core::IGameState::actionToString).def("stringToAction", &core::IGameState::stringToAction)
        .def("toString", &core::IGameState::toString).def("getMoveHistory", &core::IGameState::getMoveHistory)
        .def("getGameType", &core::IGameState::getGameType);

    py::class_<gomoku::GomokuState, core::IGameState>(m, "GomokuState")
        .def(py::init<int, bool, bool, int, bool>(), py::arg("board_size") = 15, py::arg("use_renju") = false, py::arg("use_omok") = false,
             py::arg("seed") = 0, py::arg("use_pro_long_opening") = false)
        .def("is_occupied", &gomoku::GomokuState::is_occupied).def("get_board", &gomoku::GomokuState::get_board);

    m.def("createGameState", &core::createGameState, py::arg("type"), py::arg("boardSize") = 0, py::arg("variantRules") = false);

    py::class_<nn::NeuralNetwork>(m, "NeuralNetwork")
        .def("predict", [](nn::NeuralNetwork& self, const core::IGameState& s) { py::gil_scoped_release r; return self.predict(s); })
        .def("predictBatch", [](nn::NeuralNetwork& self, const std::vector<const core::IGameState*>& states) {
            std::vector<std::reference_wrapper<const core::IGameState>> refs;
            for (auto* s : states) refs.push_back(std::cref(*s));
            std::vector<std::vector<float>> pol; std::vector<float> val;
            { py::gil_scoped_release r; self.predictBatch(refs, pol, val); }
            return std::make_pair(pol, val);
        })
        .def("isGpuAvailable", &nn::NeuralNetwork::isGpuAvailable).def("getDeviceInfo", &nn::NeuralNetwork::getDeviceInfo)
        .def("getInferenceTimeMs", &nn::NeuralNetwork::getInferenceTimeMs).def("getBatchSize", &nn::NeuralNetwork::getBatchSize)
        .def("getModelInfo", &nn::NeuralNetwork::getModelInfo).def("getModelSizeBytes", &nn::NeuralNetwork::getModelSizeBytes)
        .def("benchmark", [](nn::NeuralNetwork& self, int it, int bs) { py::gil_scoped_release r; self.benchmark(it, bs); }, py::arg("numIterations") = 100, py::arg("batchSize") = 16)
        .def("enableDebugMode", &nn::NeuralNetwork::enableDebugMode)
        .def("is_gil_safe", [](nn::NeuralNetwork&) { return true; });
    m.def("createNeuralNetwork", &nn::NeuralNetwork::create, py::arg("modelPath"), py::arg("gameType"), py::arg("boardSize") = 0, py::arg("useGpu") = true);

    py::class_<mcts::MCTSConfig>(m, "MCTSConfig").def(py::init<>())
        .def_readwrite("numThreads", &mcts::MCTSConfig::numThreads).def_readwrite("numSimulations", &mcts::MCTSConfig::numSimulations)
        .def_readwrite("cPuct", &mcts::MCTSConfig::cPuct).def_readwrite("fpuReduction", &mcts::MCTSConfig::fpuReduction)
        .def_readwrite("virtualLoss", &mcts::MCTSConfig::virtualLoss).def_readwrite("maxSearchDepth", &mcts::MCTSConfig::maxSearchDepth)
        .def_readwrite("useDirichletNoise", &mcts::MCTSConfig::useDirichletNoise).def_readwrite("dirichletAlpha", &mcts::MCTSConfig::dirichletAlpha)
        .def_readwrite("dirichletEpsilon", &mcts::MCTSConfig::dirichletEpsilon).def_readwrite("useBatchInference", &mcts::MCTSConfig::useBatchInference)
        .def_readwrite("useTemporalDifference", &mcts::MCTSConfig::useTemporalDifference).def_readwrite("tdLambda", &mcts::MCTSConfig::tdLambda)
        .def_readwrite("useProgressiveWidening", &mcts::MCTSConfig::useProgressiveWidening).def_readwrite("minVisitsForWidening", &mcts::MCTSConfig::minVisitsForWidening)
        .def_readwrite("progressiveWideningBase", &mcts::MCTSConfig::progressiveWideningBase)
        .def_readwrite("progressiveWideningExponent", &mcts::MCTSConfig::progressiveWideningExponent)
        .def_readwrite("selectionStrategy", &mcts::MCTSConfig::selectionStrategy).def_readwrite("batchSize", &mcts::MCTSConfig::batchSize)
        .def_readwrite("useBatchedMCTS", &mcts::MCTSConfig::useBatchedMCTS).def_readwrite("batchTimeoutMs", &mcts::MCTSConfig::batchTimeoutMs)
        .def_readwrite("searchMode", &mcts::MCTSConfig::searchMode);
    py::class_<mcts::MCTSStats>(m, "MCTSStats").def(py::init<>())
        .def_readonly("nodesCreated", &mcts::MCTSStats::nodesCreated).def_readonly("nodesExpanded", &mcts::MCTSStats::nodesExpanded)
        .def_readonly("nodesTotalVisits", &mcts::MCTSStats::nodesTotalVisits).def_readonly("simulationCount", &mcts::MCTSStats::simulationCount)
        .def_readonly("evaluationCalls", &mcts::MCTSStats::evaluationCalls).def_readonly("cacheHits", &mcts::MCTSStats::cacheHits)
        .def_readonly("cacheMisses", &mcts::MCTSStats::cacheMisses).def_readonly("batchedEvaluations", &mcts::MCTSStats::batchedEvaluations)
        .def_readonly("totalBatches", &mcts::MCTSStats::totalBatches);
    py::class_<mcts::MCTSNode>(m, "MCTSNode")
        .def("getUcbScore", &mcts::MCTSNode::getUcbScore, py::arg("cPuct"), py::arg("currentPlayer"), py::arg("fpuReduction") = 0.0f, py::arg("parentVisits") = 0)
        .def("getTerminalValue", &mcts::MCTSNode::getTerminalValue).def("getValue", &mcts::MCTSNode::getValue).def("getBestAction", &mcts::MCTSNode::getBestAction)
        .def("getVisitCountDistribution", &mcts::MCTSNode::getVisitCountDistribution, py::arg("temperature") = 1.0f)
        .def("toString", &mcts::MCTSNode::toString, py::arg("maxDepth") = 1)
        .def_readonly("visitCount", &mcts::MCTSNode::visitCount).def_readonly("valueSum", &mcts::MCTSNode::valueSum).def_readonly("prior", &mcts::MCTSNode::prior)
        .def_readonly("isTerminal", &mcts::MCTSNode::isTerminal).def_readonly("isExpanded", &mcts::MCTSNode::isExpanded).def_readonly("actions", &mcts::MCTSNode::actions)
        .def_readonly("childVisits", &mcts::MCTSNode::childVisits).def_readonly("childValueSums", &mcts::MCTSNode::childValueSums).def_readonly("childPriors", &mcts::MCTSNode::childPriors);
    py::class_<mcts::TranspositionTable>(m, "TranspositionTable")
        .def(py::init<size_t, size_t>(), py::arg("size") = 1048576, py::arg("numShards") = 1024)
        .def("getSize", &mcts::TranspositionTable::getSize).def("getHitRate", &mcts::TranspositionTable::getHitRate)
        .def("getLookups", &mcts::TranspositionTable::getLookups).def("getHits", &mcts::TranspositionTable::getHits)
        .def("getEntryCount", &mcts::TranspositionTable::getEntryCount).def("getMemoryUsageBytes", &mcts::TranspositionTable::getMemoryUsageBytes)
        .def("clear", &mcts::TranspositionTable::clear).def("resize", &mcts::TranspositionTable::resize);

    py::class_<mcts::ParallelMCTS>(m, "ParallelMCTS")
        .def(py::init<const core::IGameState&, nn::NeuralNetwork*, mcts::TranspositionTable*, int, int, float, float, int>(),
             py::arg("rootState"), py::arg("nn") = nullptr, py::arg("tt") = nullptr, py::arg("numThreads") = 1, py::arg("numSimulations") = 800,
             py::arg("cPuct") = 1.5f, py::arg("fpuReduction") = 0.0f, py::arg("virtualLoss") = 3, py::keep_alive<1, 3>())
        .def(py::init<const core::IGameState&, const mcts::MCTSConfig&, nn::NeuralNetwork*, mcts::TranspositionTable*>(),
             py::arg("rootState"), py::arg("config"), py::arg("nn") = nullptr, py::arg("tt") = nullptr, py::keep_alive<1, 4>())
        .def("search", [](mcts::ParallelMCTS& self) { py::gil_scoped_release r; self.search(); })
        .def("selectAction", &mcts::ParallelMCTS::selectAction, py::arg("isTraining") = false, py::arg("temperature") = 1.0f)
        .def("getActionProbabilities", &mcts::ParallelMCTS::getActionProbabilities, py::arg("temperature") = 1.0f)
        .def("getRootValue", &mcts::ParallelMCTS::getRootValue).def("updateWithMove", &mcts::ParallelMCTS::updateWithMove)
        .def("addDirichletNoise", &mcts::ParallelMCTS::addDirichletNoise, py::arg("alpha") = 0.03f, py::arg("epsilon") = 0.25f)
        .def("setNumThreads", &mcts::ParallelMCTS::setNumThreads).def("setNumSimulations", &mcts::ParallelMCTS::setNumSimulations)
        .def("setCPuct", &mcts::ParallelMCTS::setCPuct).def("setFpuReduction", &mcts::ParallelMCTS::setFpuReduction)
        .def("setVirtualLoss", &mcts::ParallelMCTS::setVirtualLoss).def("setDeterministicMode", &mcts::ParallelMCTS::setDeterministicMode)
        .def("setNeuralNetwork", &mcts::ParallelMCTS::setNeuralNetwork, py::keep_alive<1, 2>()).def("setTranspositionTable", &mcts::ParallelMCTS::setTranspositionTable, py::keep_alive<1, 2>())
        .def("setSelectionStrategy", &mcts::ParallelMCTS::setSelectionStrategy).def("setConfig", &mcts::ParallelMCTS::setConfig)
        .def("enableBatchedMCTS", &mcts::ParallelMCTS::enableBatchedMCTS).def("setBatchSize", &mcts::ParallelMCTS::setBatchSize)
        .def("setBatchTimeout", &mcts::ParallelMCTS::setBatchTimeout).def("printSearchPath", &mcts::ParallelMCTS::printSearchPath)
        .def("getNode", &mcts::ParallelMCTS::getNode, py::arg("path") = std::vector<int>{})
        .def("setDebugMode", &mcts::ParallelMCTS::setDebugMode).def("printSearchStats", &mcts::ParallelMCTS::printSearchStats)
        .def("getSearchInfo", &mcts::ParallelMCTS::getSearchInfo).def("getMemoryUsage", &mcts::ParallelMCTS::getMemoryUsage)
        .def("getRootChildren", [](const mcts::ParallelMCTS& self) { auto r = self.rootStats(); return py::make_tuple(r.actions, r.visits, r.valueSums, r.priors, r.rootVisits, r.rootValueSum); });

    py::class_<selfplay::MoveData>(m, "MoveData").def(py::init<>())
        .def_readwrite("action", &selfplay::MoveData::action).def_readwrite("policy", &selfplay::MoveData::policy)
        .def_readwrite("value", &selfplay::MoveData::value).def_readwrite("thinking_time_ms", &selfplay::MoveData::thinking_time_ms);
    py::class_<selfplay::GameRecord>(m, "GameRecord")
        .def(py::init<core::GameType, int, bool>(), py::arg("gameType"), py::arg("boardSize"), py::arg("useVariantRules") = false)
        .def("addMove", &selfplay::GameRecord::addMove).def("setResult", &selfplay::GameRecord::setResult)
        .def("getMetadata", &selfplay::GameRecord::getMetadata).def("getMoves", &selfplay::GameRecord::getMoves)
        .def("getResult", &selfplay::GameRecord::getResult).def("toJson", &selfplay::GameRecord::toJson)
        .def("saveToFile", &selfplay::GameRecord::saveToFile).def_static("fromJson", &selfplay::GameRecord::fromJson)
        .def_static("loadFromFile", &selfplay::GameRecord::loadFromFile);

    py::class_<selfplay::TrainingExample>(m, "TrainingExample").def(py::init<>())
        .def_readwrite("state", &selfplay::TrainingExample::state).def_readwrite("policy", &selfplay::TrainingExample::policy)
        .def_readwrite("value", &selfplay::TrainingExample::value).def("toJson", &selfplay::TrainingExample::toJson)
        .def_static("fromJson", &selfplay::TrainingExample::fromJson);
    py::class_<selfplay::Dataset>(m, "Dataset").def(py::init<>())
        .def("addGameRecord", &selfplay::Dataset::addGameRecord, py::arg("record"), py::arg("useEnhancedFeatures") = true)
        .def("extractExamples", [](selfplay::Dataset& self, bool aug) { py::gil_scoped_release r; self.extractExamples(aug); }, py::arg("includeAugmentations") = true)
        .def("size", &selfplay::Dataset::size).def("getBatch", &selfplay::Dataset::getBatch).def("shuffle", &selfplay::Dataset::shuffle)
        .def("saveToFile", &selfplay::Dataset::saveToFile).def("loadFromFile", &selfplay::Dataset::loadFromFile)
        .def("getRandomSubset", &selfplay::Dataset::getRandomSubset)
        .def("setShuffleOnExtract", &selfplay::Dataset::setShuffleOnExtract).def("examples", &selfplay::Dataset::examples);

    py::class_<selfplay::SelfPlayManager>(m, "SelfPlayManager")
        .def(py::init<nn::NeuralNetwork*, int, int, int>(), py::arg("neuralNetwork"), py::arg("numGames") = 100, py::arg("numSimulations") = 800,
             py::arg("numThreads") = 4, py::keep_alive<1, 2>())
        .def("generateGames", [](selfplay::SelfPlayManager& self, core::GameType t, int bs, bool v) { py::gil_scoped_release r; return self.generateGames(t, bs, v); },
             py::arg("gameType"), py::arg("boardSize") = 0, py::arg("useVariantRules") = false)
        .def("setExplorationParams", &selfplay::SelfPlayManager::setExplorationParams, py::arg("dirichletAlpha") = 0.03f, py::arg("dirichletEpsilon") = 0.25f,
             py::arg("initialTemperature") = 1.0f, py::arg("temperatureDropMove") = 30, py::arg("finalTemperature") = 0.0f)
        .def("setProgressCallback", [](selfplay::SelfPlayManager& self, std::function<void(int, int, int, int)> cb) {
            self.setProgressCallback([cb](int a, int b, int c, int d) { py::gil_scoped_acquire g; cb(a, b, c, d); });
        })
        .def("setBatchConfig", &selfplay::SelfPlayManager::setBatchConfig).def("setSaveGames", &selfplay::SelfPlayManager::setSaveGames,
             py::arg("saveGames"), py::arg("outputDir") = "games")
        .def("setAbort", &selfplay::SelfPlayManager::setAbort).def("isRunning", &selfplay::SelfPlayManager::isRunning)
        .def("setMctsConfig", [](selfplay::SelfPlayManager& self, py::dict d) {
            mcts::MCTSConfig c;
            auto geti = [&](const char* k, int& v) { if (d.contains(k) && py::isinstance<py::int_>(d[k])) v = d[k].cast<int>(); };
            auto getf = [&](const char* k, float& v) { if (d.contains(k) && py::isinstance<py::float_>(d[k])) v = d[k].cast<float>(); };
            auto getb = [&](const char* k, bool& v) { if (d.contains(k) && py::isinstance<py::bool_>(d[k])) v = d[k].cast<bool>(); };
            geti("numThreads", c.numThreads); geti("numSimulations", c.numSimulations); getf("cPuct", c.cPuct); getf("fpuReduction", c.fpuReduction);
            geti("virtualLoss", c.virtualLoss); getb("useDirichletNoise", c.useDirichletNoise); getf("dirichletAlpha", c.dirichletAlpha);
            getf("dirichletEpsilon", c.dirichletEpsilon); getb("useBatchInference", c.useBatchInference); getb("useBatchedMCTS", c.useBatchedMCTS);
            geti("batchSize", c.batchSize); geti("batchTimeoutMs", c.batchTimeoutMs);
            getb("useTemporalDifference", c.useTemporalDifference); getb("useProgressiveWidening", c.useProgressiveWidening); getb("useFmapCache", c.useFmapCache);
            if (d.contains("searchMode") && py::isinstance<py::str>(d["searchMode"])) {      // python_bindings.cpp:436-444 (every mode runs the wave engine)
                const std::string m = d["searchMode"].cast<std::string>();
                if (m == "SERIAL") c.searchMode = mcts::MCTSSearchMode::SERIAL; else if (m == "PARALLEL") c.searchMode = mcts::MCTSSearchMode::PARALLEL;
                else if (m == "BATCHED") c.searchMode = mcts::MCTSSearchMode::BATCHED;
            }
            self.setMctsConfig(c);
        })
        .def("getCompletedGamesCount", &selfplay::SelfPlayManager::getCompletedGamesCount).def("getTotalMovesCount", &selfplay::SelfPlayManager::getTotalMovesCount)
        .def("setConcurrentGames", &selfplay::SelfPlayManager::setConcurrentGames).def("setDeterministic", &selfplay::SelfPlayManager::setDeterministic)
        .def("setDevices", &selfplay::SelfPlayManager::setDevices).def("getDevices", &selfplay::SelfPlayManager::getDevices)
        .def("getLastRunStats", &selfplay::SelfPlayManager::getLastRunStats).def("getLastGatheredSampleBytes", &selfplay::SelfPlayManager::getLastGatheredSampleBytes);
}
