// alphazero_host.hpp — C++ host side above the C ABI: the reference's class names, signatures and error behaviour
// for the self-play path, implemented on top of libaz_b200.so (include/az_b200.h).  No search or network arithmetic
// happens here: these classes hold moves / configuration and forward to the engine.
//
// Mirrors (reference file:line):
//   core::IGameState                 include/alphazero/core/igamestate.h:60-223
//   gomoku::GomokuState              include/alphazero/games/gomoku/gomoku_state.h:21-171 (standard rules)
//   nn::NeuralNetwork                include/alphazero/nn/neural_network.h:19-131
//   mcts::MCTSConfig / ParallelMCTS  include/alphazero/mcts/parallel_mcts.h:41-74, 131-282
//   selfplay::MoveData / GameRecord  include/alphazero/selfplay/game_record.h:18-126
//   selfplay::SelfPlayManager        include/alphazero/selfplay/self_play_manager.h:28-196
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <random>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../../include/az_b200.h"

namespace alphazero {

namespace core {

enum class GameType { GOMOKU, CHESS, GO };
enum class GameResult { ONGOING, DRAW, WIN_PLAYER1, WIN_PLAYER2 };

class GameStateException : public std::runtime_error {
public:
    explicit GameStateException(const std::string& m) : std::runtime_error(m) {}
};
class IllegalMoveException : public GameStateException {
public:
    IllegalMoveException(const std::string& m, int action) : GameStateException(m), action_(action) {}
    int getAction() const { return action_; }
private:
    int action_;
};

class IGameState {
public:
    explicit IGameState(GameType t) : gameType_(t) {}
    virtual ~IGameState() = default;
    virtual std::vector<int> getLegalMoves() const = 0;
    virtual bool isLegalMove(int action) const = 0;
    virtual void makeMove(int action) = 0;
    virtual bool undoMove() = 0;
    virtual bool isTerminal() const = 0;
    virtual GameResult getGameResult() const = 0;
    virtual int getCurrentPlayer() const = 0;
    virtual int getBoardSize() const = 0;
    virtual int getActionSpaceSize() const = 0;
    virtual std::vector<std::vector<std::vector<float>>> getTensorRepresentation() const = 0;
    virtual std::vector<std::vector<std::vector<float>>> getEnhancedTensorRepresentation() const = 0;
    virtual uint64_t getHash() const = 0;
    virtual std::unique_ptr<IGameState> clone() const = 0;
    virtual std::string actionToString(int action) const = 0;
    virtual std::optional<int> stringToAction(const std::string& s) const = 0;
    virtual std::string toString() const = 0;
    virtual bool equals(const IGameState& other) const = 0;
    virtual std::vector<int> getMoveHistory() const = 0;
    virtual bool validate() const = 0;
    GameType getGameType() const { return gameType_; }
protected:
    GameType gameType_;
};

std::unique_ptr<IGameState> createGameState(GameType type, int boardSize = 0, bool variantRules = false);

// include/alphazero/core/game_factory.h:16-75.  The reference's factory goes through its GameRegistry, in which no game is ever registered
// (SURVEY.md Appendix B), so at HEAD every create* call throws and isGameSupported is false for every game; this one constructs the states.
class GameFactory {
public:
    static std::unique_ptr<IGameState> createGomokuState(int boardSize = 15, bool useRenju = false, bool useOmok = false, int seed = 0, bool useProLongOpening = false);
    static std::unique_ptr<IGameState> createChessState(bool chess960 = false, const std::string& fen = "");
    static std::unique_ptr<IGameState> createGoState(int boardSize = 19, float komi = 7.5f, bool chineseRules = true);
    static bool isGameSupported(GameType type);
    static int getDefaultBoardSize(GameType type);
};

}  // namespace core

namespace gomoku {

constexpr int BLACK = 1, WHITE = 2;

// Host-side state container: bitboards + history, standard rules.  Rules arithmetic is the SAME header the kernels
// compile (csrc/gomoku.cuh, host+device) for 15x15 and 9x9; other sizes use the generic loop below it.
class GomokuState : public core::IGameState {
public:
    GomokuState(int board_size = 15, bool use_renju = false, bool use_omok = false, int seed = 0, bool use_pro_long_opening = false);
    std::vector<int> getLegalMoves() const override;
    bool isLegalMove(int action) const override;
    void makeMove(int action) override;
    bool undoMove() override;
    bool isTerminal() const override;
    core::GameResult getGameResult() const override;
    int getCurrentPlayer() const override { return current_player; }
    int getBoardSize() const override { return board_size; }
    int getActionSpaceSize() const override { return board_size * board_size; }
    std::vector<std::vector<std::vector<float>>> getTensorRepresentation() const override;
    std::vector<std::vector<std::vector<float>>> getEnhancedTensorRepresentation() const override;
    uint64_t getHash() const override;
    std::unique_ptr<core::IGameState> clone() const override { return std::make_unique<GomokuState>(*this); }
    std::string actionToString(int action) const override;
    std::optional<int> stringToAction(const std::string& s) const override;
    std::string toString() const override;
    bool equals(const core::IGameState& other) const override;
    std::vector<int> getMoveHistory() const override { return move_history; }
    bool validate() const override;
    bool is_occupied(int action) const;
    std::vector<std::vector<int>> get_board() const;
    // the rest of the reference's public helpers (gomoku_state.h:88-117), used by its unit tests
    bool is_bit_set(int player_index, int action) const noexcept;      // player_index 0 = BLACK, 1 = WHITE
    int count_total_stones() const noexcept;
    bool board_equal(const GomokuState& other) const;
    bool isUsingRenjuRules() const { return false; }                   // Renju / Omok states cannot be constructed
    bool isUsingOmokRules() const { return false; }
    // true until the legal moves of this lineage were enumerated once (QUIRK G2: first-fill order)
    bool neverEnumerated() const { return never_filled_; }

    int board_size;
    int current_player;
private:
    int winner() const;
    std::vector<int8_t> cells_;                 // 0 empty, 1 black, 2 white; index a = x*N + y
    std::vector<int> move_history;
    // legal-move cache with the reference's container, so its iteration order is the reference's
    mutable std::unordered_set<int> cached_valid_moves_;
    mutable bool valid_moves_dirty_ = true;
    mutable bool never_filled_ = true;
};

}  // namespace gomoku

namespace go {

// Host-side Go state (reference go::GoState, include/alphazero/games/go/go_state.h:30-219): Chinese rules, positional
// superko, komi 7.5, pass = action -1.  Rules arithmetic is the SAME header the kernels compile (csrc/go.cuh, host+device)
// for 9x9 / 13x13 / 19x19.
class GoState : public core::IGameState {
public:
    explicit GoState(int board_size = 19, float komi = 7.5f, bool chinese_rules = true, bool enforce_superko = true);
    GoState(const GoState& o);
    ~GoState() override;
    std::vector<int> getLegalMoves() const override;
    bool isLegalMove(int action) const override;
    void makeMove(int action) override;
    bool undoMove() override;
    bool isTerminal() const override;
    core::GameResult getGameResult() const override;
    int getCurrentPlayer() const override;
    int getBoardSize() const override { return board_size_; }
    int getActionSpaceSize() const override { return board_size_ * board_size_ + 1; }      // go_state.cpp:345-347
    std::vector<std::vector<std::vector<float>>> getTensorRepresentation() const override;       // 3 planes: black, white, side to move (go_state.cpp:349-378)
    std::vector<std::vector<std::vector<float>>> getEnhancedTensorRepresentation() const override;
    uint64_t getHash() const override;
    std::unique_ptr<core::IGameState> clone() const override { return std::make_unique<GoState>(*this); }
    std::string actionToString(int action) const override;
    std::optional<int> stringToAction(const std::string& s) const override;
    std::string toString() const override;
    bool equals(const core::IGameState& other) const override;
    std::vector<int> getMoveHistory() const override { return move_history_; }
    bool validate() const override { return true; }
    int getStone(int pos) const;                 // 0 empty, 1 black, 2 white; pos = y*N + x
    int getKoPoint() const;
    // go_state.h:112-150
    int getCapturedStones(int player) const { return (player == 1 || player == 2) ? captured_[player] : 0; }
    float getKomi() const { return komi_; }
    bool isChineseRules() const { return true; }
    bool isEnforcingSuperko() const { return true; }
    std::pair<int, int> actionToCoord(int action) const { return (action < 0 || action >= board_size_ * board_size_) ? std::pair<int, int>{-1, -1} : std::pair<int, int>{action % board_size_, action / board_size_}; }
    int coordToAction(int x, int y) const { return (x < 0 || y < 0 || x >= board_size_ || y >= board_size_) ? -1 : y * board_size_ + x; }
    uint64_t hashEvaluatorKey() const;           // SURVEY Appendix C key (B200NeuralNetwork("hash").predict)
    struct Impl;
private:
    int board_size_;
    std::unique_ptr<Impl> impl_;
    std::vector<int> move_history_;
    int captured_[3] = {0, 0, 0};                // stones captured BY player 1 / 2 (captured_stones_, go_state.cpp:245)
    float komi_ = 7.5f;                          // any komi for the state's own scoring; the engine is built for 7.5
};

}  // namespace go

namespace chess {

// include/alphazero/games/chess/chess_state.h:21-82
enum class PieceType { NONE = 0, PAWN = 1, KNIGHT = 2, BISHOP = 3, ROOK = 4, QUEEN = 5, KING = 6 };
enum class PieceColor { NONE = 0, WHITE = 1, BLACK = 2 };
struct CastlingRights {
    bool white_kingside = true, white_queenside = true, black_kingside = true, black_queenside = true;
    bool operator==(const CastlingRights& o) const { return white_kingside == o.white_kingside && white_queenside == o.white_queenside && black_kingside == o.black_kingside && black_queenside == o.black_queenside; }
};
struct Piece {
    PieceType type = PieceType::NONE;
    PieceColor color = PieceColor::NONE;
    bool operator==(const Piece& o) const { return type == o.type && color == o.color; }
    bool operator!=(const Piece& o) const { return !(*this == o); }
    bool is_empty() const { return type == PieceType::NONE; }
};

// Host-side chess state (reference chess::ChessState, include/alphazero/games/chess/chess_state.h:84-406; standard chess
// from the initial position).  Rules arithmetic is the SAME header the kernels compile (csrc/chess.cuh, host+device),
// including the reference's literal attack test (QUIRK C5) and placement-only repetition key.
class ChessState : public core::IGameState {
public:
    explicit ChessState(bool chess960 = false, const std::string& fen = "", int position_number = -1);
    ChessState(const ChessState& o);
    ~ChessState() override;
    std::vector<int> getLegalMoves() const override;
    bool isLegalMove(int action) const override;
    void makeMove(int action) override;
    bool undoMove() override;
    bool isTerminal() const override;
    core::GameResult getGameResult() const override;
    int getCurrentPlayer() const override;
    int getBoardSize() const override { return 8; }
    int getActionSpaceSize() const override { return 64 * 64 * 5; }                    // chess_state.h:117
    std::vector<std::vector<std::vector<float>>> getTensorRepresentation() const override;        // 12 piece planes
    std::vector<std::vector<std::vector<float>>> getEnhancedTensorRepresentation() const override;   // 18 planes
    uint64_t getHash() const override;
    std::unique_ptr<core::IGameState> clone() const override { return std::make_unique<ChessState>(*this); }
    std::string actionToString(int action) const override;                             // "e2e4", "e7e8q"
    std::optional<int> stringToAction(const std::string& s) const override;
    std::string toString() const override;
    std::string toFEN() const;                                                         // chess_state.cpp:270-378
    bool equals(const core::IGameState& other) const override;
    std::vector<int> getMoveHistory() const override { return move_history_; }
    bool validate() const override { return true; }
    int getPieceCode(int square) const;          // type | colour << 3 (0 = empty)
    // chess_state.h:150-215: square = rank * 8 + file with rank 0 = the eighth rank (the reference's getSquare)
    Piece getPiece(int square) const;
    CastlingRights getCastlingRights() const;
    int getEnPassantSquare() const;
    int getHalfmoveClock() const;
    int getFullmoveNumber() const;
    bool isFromFEN() const { return !fen_start_.empty(); }
    bool setFromFEN(const std::string& fen);     // chess_state.cpp:382-495 (standard castling letters; Chess960 rook-file letters are not read)
    bool isInCheck() const;
    uint64_t hashEvaluatorKey() const;
    struct Impl;
private:
    std::unique_ptr<Impl> impl_;
    std::vector<int> move_history_;
    std::string fen_start_;                      // FEN the game was set up from ("" = the initial position): undoMove replays from it
};

}  // namespace chess

namespace nn {

class NeuralNetwork {
public:
    virtual ~NeuralNetwork() = default;
    virtual std::pair<std::vector<float>, float> predict(const core::IGameState& state) = 0;
    virtual void predictBatch(const std::vector<std::reference_wrapper<const core::IGameState>>& states,
                              std::vector<std::vector<float>>& policies, std::vector<float>& values) = 0;
    virtual bool isGpuAvailable() const = 0;
    virtual std::string getDeviceInfo() const = 0;
    virtual float getInferenceTimeMs() const = 0;
    virtual int getBatchSize() const = 0;
    virtual std::string getModelInfo() const = 0;
    virtual size_t getModelSizeBytes() const = 0;
    virtual void benchmark(int numIterations = 100, int batchSize = 16) = 0;
    virtual void enableDebugMode(bool enable) = 0;
    virtual void printModelSummary() const = 0;
    // modelPath: an AZW1 weight blob (net.py:export_weights); "" or "hash" → the stateless HashEvaluator (parity runs)
    static std::unique_ptr<NeuralNetwork> create(const std::string& modelPath, core::GameType gameType, int boardSize = 0, bool useGpu = true);
};

// The evaluator the engine runs on the device: either the bf16 ResNet (weights from an AZW1 blob) or the hash evaluator.
class B200NeuralNetwork : public NeuralNetwork {
public:
    B200NeuralNetwork(const std::string& modelPath, core::GameType gameType, int boardSize);
    ~B200NeuralNetwork() override;
    std::pair<std::vector<float>, float> predict(const core::IGameState& state) override;
    void predictBatch(const std::vector<std::reference_wrapper<const core::IGameState>>& states,
                      std::vector<std::vector<float>>& policies, std::vector<float>& values) override;
    bool isGpuAvailable() const override { return true; }
    std::string getDeviceInfo() const override;
    float getInferenceTimeMs() const override { return lastMs_; }
    int getBatchSize() const override { return batch_; }
    std::string getModelInfo() const override;
    size_t getModelSizeBytes() const override { return blob_.size(); }
    void benchmark(int numIterations = 100, int batchSize = 16) override;
    void enableDebugMode(bool) override {}
    void printModelSummary() const override;
    bool isHash() const { return hash_; }
    const std::vector<uint8_t>& blob() const { return blob_; }
    core::GameType gameType() const { return gameType_; }
    int boardSize() const { return boardSize_; }
    int blocks() const { return blocks_; }
    int channels() const { return channels_; }
private:
    void ensureEngine();
    std::vector<uint8_t> blob_;
    bool hash_ = false;
    core::GameType gameType_;
    int boardSize_, blocks_ = 0, channels_ = 128, batch_ = 64;
    az_engine* eng_ = nullptr;
    float lastMs_ = 0.0f;
};

}  // namespace nn

namespace mcts {

enum class MCTSNodeSelection { UCB, PUCT, PROGRESSIVE_BIAS, RAVE };
enum class MCTSSearchMode { SERIAL, PARALLEL, BATCHED };

struct MCTSConfig {   // include/alphazero/mcts/parallel_mcts.h:41-74 (same names, same defaults)
    int numThreads = 1;
    int numSimulations = 800;
    float cPuct = 1.5f;
    float fpuReduction = 0.0f;
    int virtualLoss = 3;
    int maxSearchDepth = 1000;
    bool useDirichletNoise = false;
    float dirichletAlpha = 0.03f;
    float dirichletEpsilon = 0.25f;
    bool useBatchInference = false;
    bool useTemporalDifference = false;
    float tdLambda = 0.8f;
    bool useProgressiveWidening = false;
    int minVisitsForWidening = 10;
    float progressiveWideningBase = 2.0f;
    float progressiveWideningExponent = 0.5f;
    MCTSNodeSelection selectionStrategy = MCTSNodeSelection::PUCT;
    int maxRetries = 3;
    int transpositionTableSize = 1048576;
    uint64_t cacheEntryMaxAge = 60000;
    bool useFmapCache = false;
    int batchSize = 16;
    bool useBatchedMCTS = false;
    int batchTimeoutMs = 5;
    MCTSSearchMode searchMode = MCTSSearchMode::PARALLEL;
    bool pinThreads = false;
    bool deterministic = false;
    int cacheSize = 2097152;
};

struct MCTSStats {
    size_t nodesCreated = 0, nodesExpanded = 0, nodesTotalVisits = 0, simulationCount = 0, evaluationCalls = 0,
           cacheHits = 0, cacheMisses = 0, batchedEvaluations = 0, totalBatches = 0;
};

// The reference's TT is an evaluation cache that is result-transparent with a deterministic evaluator;
// the engine has no use for it.  Kept so existing call sites construct and pass one.
// mcts::TranspositionTable (include/alphazero/mcts/transposition_table.h, src/mcts/transposition_table.cpp:44-84, 128-176): the evaluation cache.
// On the B200 engine the table itself lives on the device (az_config.eval_cache_entries, csrc/tree.cuh EvalCache: 64-bit network-input key ->
// fp32 policy + value, 4-way buckets, filled and probed inside the waves); this object carries the size the caller asks for into the engines
// built with it and collects their lookup / hit counters.  Searches are bit-identical with it present or absent.
class TranspositionTable {
public:
    explicit TranspositionTable(size_t size = 1048576, size_t numShards = 1024) : size_(size) { (void)numShards; }
    size_t getSize() const { return size_; }
    float getHitRate() const { return lookups_ ? (float)hits_ / (float)lookups_ : 0.0f; }
    size_t getLookups() const { return lookups_; }
    size_t getHits() const { return hits_; }
    size_t getEntryCount() const { return std::min<size_t>(lookups_ - hits_, size_); }      // one store per miss
    size_t getMemoryUsageBytes() const { return size_ * 16; }                                // key / stamp / value; the policies are per game type
    void clear() { lookups_ = hits_ = 0; }                                                   // (device tables are cleared when an engine loads weights)
    void resize(size_t s) { size_ = s; }
    void addStats(size_t lookups, size_t hits) { lookups_ += lookups; hits_ += hits; }       // called by ParallelMCTS::search
private:
    size_t size_, lookups_ = 0, hits_ = 0;
};

// mcts::MCTSNode (include/alphazero/mcts/mcts_node.h:54-75, 80-119, 200-230) as a read-only SNAPSHOT of one node of the device tree: the
// fields and the inspection methods the reference binds to Python (python_bindings.cpp:245-253); children are held as plain statistics.
class MCTSNode {
public:
    int visitCount = 0, virtualLoss = 0; float valueSum = 0.0f, prior = 0.0f;
    bool isTerminal = false, isExpanded = false; core::GameResult gameResult = core::GameResult::ONGOING;
    std::vector<int> actions, childVisits; std::vector<float> childValueSums, childPriors;
    float getValue() const { return visitCount == 0 ? 0.0f : valueSum / visitCount; }                       // mcts_node.h:80-86
    float getTerminalValue(int currentPlayer) const;                                                       // mcts_node.cpp:387-399
    float getUcbScore(float cPuct, int currentPlayer, float fpuReduction = 0.0f, int parentVisits = 0) const;   // mcts_node.cpp:121-166
    int getBestAction() const;                                                                             // first child with the most visits
    std::vector<float> getVisitCountDistribution(float temperature = 1.0f) const;                          // mcts_node.cpp:289-322
    std::string toString(int maxDepth = 1) const;
};

class ParallelMCTS {
public:
    ParallelMCTS(const core::IGameState& rootState, nn::NeuralNetwork* nn = nullptr, TranspositionTable* tt = nullptr,
                 int numThreads = 1, int numSimulations = 800, float cPuct = 1.5f, float fpuReduction = 0.0f, int virtualLoss = 3);
    ParallelMCTS(const core::IGameState& rootState, const MCTSConfig& config, nn::NeuralNetwork* nn = nullptr, TranspositionTable* tt = nullptr);
    ~ParallelMCTS();
    void search();
    int selectAction(bool isTraining = false, float temperature = 1.0f);
    std::vector<float> getActionProbabilities(float temperature = 1.0f) const;
    float getRootValue() const;
    void updateWithMove(int action);
    void addDirichletNoise(float alpha = 0.03f, float epsilon = 0.25f);
    void setNumThreads(int n) { config_.numThreads = n; }
    void setNumSimulations(int n) { config_.numSimulations = n; }
    void setCPuct(float c);
    void setFpuReduction(float f) { config_.fpuReduction = f; }
    void setVirtualLoss(int v);
    void setDeterministicMode(bool enable) { config_.useBatchInference = enable; }
    void setDebugMode(bool) {}
    // parallel_mcts.cpp:1173-1261.  The evaluator runs inside the device waves: a new network is loaded into the engine (tree kept, as in
    // the reference); the transposition table, batch size / timeout and batched-search switches have nothing to act on (every wave is one
    // batch) and are recorded only; PUCT is the one selection strategy built (the reference's default and the only one self-play uses).
    void setNeuralNetwork(nn::NeuralNetwork* nn);
    void setTranspositionTable(TranspositionTable* tt) { tt_ = tt; }
    void setSelectionStrategy(MCTSNodeSelection s);
    void setConfig(const MCTSConfig& config);
    void enableBatchedMCTS(bool enable) { config_.useBatchedMCTS = enable; }
    void setBatchSize(int n) { config_.batchSize = n; }
    void setBatchTimeout(int ms) { config_.batchTimeoutMs = ms; }
    void printSearchPath(int action) const;                     // parallel_mcts.cpp:1390-1450
    MCTSNode getNode(const std::vector<int>& path = {}) const;  // snapshot of the node reached from the root by `path` ({} = the root)
    std::string getSearchInfo() const;
    void printSearchStats() const;
    size_t getMemoryUsage() const;
    MCTSStats getStats() const;
    // root children in the reference's child order
    struct RootStats { std::vector<int> actions, visits; std::vector<float> valueSums, priors; int rootVisits = 0; float rootValueSum = 0; };
    RootStats rootStats() const;
private:
    void build(const core::IGameState& rootState);
    MCTSConfig config_;
    nn::NeuralNetwork* nn_;
    TranspositionTable* tt_ = nullptr;
    size_t ttLookups_ = 0, ttHits_ = 0;      // engine counters already credited to tt_
    az_engine* eng_ = nullptr;
    bool external_ = false;      // nn_ is not a B200NeuralNetwork: the leaves are evaluated on the host through nn_->predictBatch (AZ_EVAL_EXTERNAL)
    std::string evalError_;
    static int evalTrampoline(int n, const int32_t* slot, const int32_t* paths, const int32_t* lens, int maxLen, int actions, float* policy, float* value, void* user);
    std::unique_ptr<core::IGameState> rootState_;
    bool searched_ = false;
};

}  // namespace mcts

namespace selfplay {

struct MoveData {
    int action = -1;
    std::vector<float> policy;
    float value = 0.0f;
    int64_t thinking_time_ms = 0;
    std::string toJson() const;
    static MoveData fromJson(const std::string& jsonStr);
};

class GameRecord {
public:
    GameRecord(core::GameType gameType, int boardSize, bool useVariantRules = false);
    void addMove(int action, const std::vector<float>& policy, float value, int64_t thinkingTimeMs);
    void setResult(core::GameResult r) { result_ = r; }
    std::tuple<core::GameType, int, bool> getMetadata() const { return {gameType_, boardSize_, useVariantRules_}; }
    const std::vector<MoveData>& getMoves() const { return moves_; }
    core::GameResult getResult() const { return result_; }
    std::string toJson() const;                                  // src/selfplay/game_record.cpp:64-90 format
    static GameRecord fromJson(const std::string& jsonStr);
    bool saveToFile(const std::string& filename) const;
    static GameRecord loadFromFile(const std::string& filename);
private:
    core::GameType gameType_; int boardSize_; bool useVariantRules_;
    core::GameResult result_;
    std::vector<MoveData> moves_;
    std::chrono::system_clock::time_point timestamp_;
};

// include/alphazero/selfplay/dataset.h:21-128
struct TrainingExample {
    std::vector<std::vector<std::vector<float>>> state;   // [plane][row][col]
    std::vector<float> policy;
    float value = 0.0f;
    std::string toJson() const;
    static TrainingExample fromJson(const std::string& json);
};

// Dataset (src/selfplay/dataset.cpp): same methods and file format.  extractExamples replays the records, evaluates the feature
// planes and builds the dihedral images on the GPU (az_engine_examples_from_games); everything else is host bookkeeping.
class Dataset {
public:
    Dataset();
    void addGameRecord(const GameRecord& record, bool useEnhancedFeatures = true);
    void extractExamples(bool includeAugmentations = true);
    size_t size() const { return examples_.size(); }
    std::tuple<std::vector<std::vector<std::vector<std::vector<float>>>>, std::vector<std::vector<float>>, std::vector<float>> getBatch(size_t batchSize) const;
    void shuffle();
    bool saveToFile(const std::string& filename) const;
    bool loadFromFile(const std::string& filename);
    std::vector<TrainingExample> getRandomSubset(size_t count) const;
    // engine knobs without a reference counterpart
    void setShuffleOnExtract(bool s) { shuffleOnExtract_ = s; }  // false: examples stay in record order (parity tests)
    const std::vector<TrainingExample>& examples() const { return examples_; }
private:
    std::vector<GameRecord> gameRecords_;
    std::vector<TrainingExample> examples_;
    mutable std::mt19937 rng_;
    bool shuffleOnExtract_ = true;
};

class SelfPlayManager {
public:
    SelfPlayManager(nn::NeuralNetwork* neuralNetwork, int numGames = 100, int numSimulations = 800, int numThreads = 4);
    ~SelfPlayManager();
    std::vector<GameRecord> generateGames(core::GameType gameType, int boardSize = 0, bool useVariantRules = false);
    void setExplorationParams(float dirichletAlpha = 0.03f, float dirichletEpsilon = 0.25f, float initialTemperature = 1.0f,
                              int temperatureDropMove = 30, float finalTemperature = 0.0f);
    void setProgressCallback(std::function<void(int, int, int, int)> cb) { progressCallback_ = std::move(cb); }
    void setBatchConfig(int batchSize, int batchTimeoutMs) { batchSize_ = batchSize; batchTimeoutMs_ = batchTimeoutMs; }
    void setSaveGames(bool saveGames, const std::string& outputDir = "games") { saveGames_ = saveGames; outputDir_ = outputDir; }
    void setAbort(bool abort) { abort_ = abort; }
    bool isRunning() const { return running_; }
    void setMctsConfig(const mcts::MCTSConfig& c) { mctsConfig_ = c; }
    int getCompletedGamesCount() const { return completedGames_; }
    int getTotalMovesCount() const { return totalMoves_; }
    // engine knobs without a reference counterpart
    void setConcurrentGames(int n) { concurrentGames_ = n; }     // game slots per GPU (default min(numGames, 4096))
    void setDeterministic(bool d) { deterministic_ = d; }        // noise off + first-max-visit move (parity runs)
    // GPUs the games are sharded over (default: device 0 only).  One engine + one host thread per device, no exchange during the waves;
    // finished-game samples are drained on each device, all-gathered with ncclAllGather and the counters summed with ncclAllReduce
    //  — libnccl.so.2 is loaded at run time when more than one device is named.
    void setDevices(const std::vector<int>& devices) { devices_ = devices; }
    const std::vector<int>& getDevices() const { return devices_; }
    // counters summed over the devices by the last generateGames: simulations, evaluations, moves, games (ncclAllReduce when > 1 device)
    std::vector<unsigned long long> getLastRunStats() const { return lastStats_; }
    size_t getLastGatheredSampleBytes() const { return lastGatheredBytes_; }
private:
    nn::NeuralNetwork* nn_;
    int numGames_, numSimulations_, numThreads_;
    float dirichletAlpha_ = 0.03f, dirichletEpsilon_ = 0.25f, initialTemperature_ = 1.0f, finalTemperature_ = 0.0f;
    int temperatureDropMove_ = 30;
    bool saveGames_ = false; std::string outputDir_ = "games";
    std::function<void(int, int, int, int)> progressCallback_;
    std::atomic<bool> abort_{false}, running_{false};
    std::atomic<int> completedGames_{0}, totalMoves_{0};
    int batchSize_ = 64, batchTimeoutMs_ = 10, concurrentGames_ = 0;
    bool deterministic_ = false;
    mcts::MCTSConfig mctsConfig_;
    std::vector<int> devices_;
    std::vector<unsigned long long> lastStats_;
    size_t lastGatheredBytes_ = 0;
    std::vector<GameRecord> generateGamesMultiGpu(core::GameType gameType, int bs, const az_config& base);
    void appendRecords(const uint8_t* data, size_t n, const az_sample_layout& L, core::GameType gameType, int bs, int A, int64_t ms, std::vector<GameRecord>& done);
};

}  // namespace selfplay
}  // namespace alphazero
