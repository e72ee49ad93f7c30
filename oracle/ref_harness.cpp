// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// A thin C-ABI wrapper around the *unmodified-but-shimmed* reference classes
// (GomokuState / GoState / ParallelMCTS) so Python tests can drive the real
// reference search and state API.  Compiled by oracle/build_ref.sh against a
// temporary patched copy of /root/reference (patch shim: SURVEY.md §8c) into
// oracle/_ref/libaz_ref.so.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
//
// Evaluators offered to the reference search (both are nn::NeuralNetwork subclasses,
// include/alphazero/nn/neural_network.h:19-131):
//   * HashEvaluator  — stateless integer-mix evaluator of SURVEY.md Appendix C
//                      (exactly-rounded fp32 ops only → bit-reproducible on the GPU)
//   * CallbackEvaluator — hands the reference's own enhanced tensor
//                      (getEnhancedTensorRepresentation) to a C callback that returns
//                      (policy[A], value); used to time the reference search with the
//                      fp32 TorchScript-equivalent network on the host CPU.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <string>
#include <future>
#include <algorithm>

#include "alphazero/core/igamestate.h"
#include "alphazero/games/gomoku/gomoku_state.h"
#include "alphazero/games/go/go_state.h"
#include "alphazero/games/chess/chess_state.h"
#include "alphazero/selfplay/game_record.h"
#include "alphazero/core/registry.h"
#define private public      /* Dataset::augmentExample is private: the harness pins its image ORDER, which extractExamples' final shuffle hides */
#include "alphazero/selfplay/dataset.h"
#undef private
#include "alphazero/mcts/parallel_mcts.h"
#include "alphazero/mcts/mcts_node.h"
#include "alphazero/mcts/transposition_table.h"
#include "alphazero/nn/neural_network.h"

using alphazero::core::IGameState;
using alphazero::core::GameType;

namespace {

inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

// Canonical key of a state, independent of the (time-seeded) Zobrist tables.
uint64_t state_key(const IGameState& s) {
    uint64_t h = 1469598103934665603ULL;
    if (s.getGameType() == GameType::GOMOKU) {
        const auto& g = dynamic_cast<const alphazero::gomoku::GomokuState&>(s);
        for (int p = 0; p < 2; ++p)
            for (uint64_t w : g.player_bitboards[p]) h = mix64(h ^ w);
        h = mix64(h ^ (uint64_t)g.current_player);
    } else if (s.getGameType() == GameType::GO) {
        const auto& g = dynamic_cast<const alphazero::go::GoState&>(s);
        int n = g.getBoardSize();
        for (int pos = 0; pos < n * n; ++pos) h = mix64(h ^ (uint64_t)g.getStone(pos));
        h = mix64(h ^ (uint64_t)g.getCurrentPlayer());
        h = mix64(h ^ (uint64_t)(int64_t)(g.getKoPoint() + 1));
    } else if (s.getGameType() == GameType::CHESS) {
        // this repo's chess key (oracle/az_oracle.cpp Chess::key): pieces, side, castling rights, e.p. square
        const auto& c = dynamic_cast<const alphazero::chess::ChessState&>(s);
        for (int sq = 0; sq < 64; ++sq) { const auto pc = c.getPiece(sq); h = mix64(h ^ (uint64_t)((int)pc.type + 8 * (int)pc.color)); }
        h = mix64(h ^ (uint64_t)c.getCurrentPlayer());
        const auto cr = c.getCastlingRights();
        h = mix64(h ^ (uint64_t)((cr.white_kingside ? 1 : 0) | (cr.white_queenside ? 2 : 0) | (cr.black_kingside ? 4 : 0) | (cr.black_queenside ? 8 : 0)));
        h = mix64(h ^ (uint64_t)(int64_t)(c.getEnPassantSquare() + 1));
    }
    return h;
}

void hash_eval(uint64_t h, int A, float* policy, float* value, bool peaked = false) {
    float sum = 0.0f;
    const int peak = peaked ? (int)(mix64(h ^ 0x5EEDULL) % (uint64_t)A) : -1;     // evaluator 2: one action's raw prior x 4096
    for (int i = 0; i < A; ++i) {
        uint64_t r = mix64(h + (uint64_t)i * 0x9E3779B97F4A7C15ULL) >> 40;
        float raw = (float)(r + 1) / (float)(1 << 24);
        if (i == peak) raw = raw * 4096.0f;
        policy[i] = raw;
        sum += raw;
    }
    for (int i = 0; i < A; ++i) policy[i] = policy[i] / sum;
    float v = ((float)(mix64(h ^ 0xABCDEFULL) >> 40) / (float)(1 << 24)) * 2.0f - 1.0f;
    *value = v * 0.5f;
}

class EvaluatorBase : public alphazero::nn::NeuralNetwork {
public:
    long calls = 0;
    void predictBatch(const std::vector<std::reference_wrapper<const IGameState>>& states,
                      std::vector<std::vector<float>>& policies, std::vector<float>& values) override {
        policies.clear(); values.clear();
        for (auto& s : states) { auto pv = predict(s.get()); policies.push_back(pv.first); values.push_back(pv.second); }
    }
    std::future<std::pair<std::vector<float>, float>> predictAsync(const IGameState& state) override {
        std::promise<std::pair<std::vector<float>, float>> p; p.set_value(predict(state)); return p.get_future();
    }
    bool isGpuAvailable() const override { return false; }
    std::string getDeviceInfo() const override { return "cpu"; }
    float getInferenceTimeMs() const override { return 0.0f; }
    int getBatchSize() const override { return 1; }
    std::string getModelInfo() const override { return "oracle-evaluator"; }
    size_t getModelSizeBytes() const override { return 0; }
    void benchmark(int, int) override {}
    void enableDebugMode(bool) override {}
    void printModelSummary() const override {}
};

class HashEvaluator : public EvaluatorBase {
public:
    bool peaked = false;
    std::pair<std::vector<float>, float> predict(const IGameState& s) override {
        ++calls;
        int A = s.getActionSpaceSize();
        std::vector<float> pol(A); float v;
        hash_eval(state_key(s), A, pol.data(), &v, peaked);
        return {pol, v};
    }
};

typedef void (*eval_cb_t)(const float* planes, int C, int H, int W, int A, float* policy_out, float* value_out, void* user);

class CallbackEvaluator : public EvaluatorBase {
public:
    eval_cb_t cb; void* user;
    CallbackEvaluator(eval_cb_t c, void* u) : cb(c), user(u) {}
    std::pair<std::vector<float>, float> predict(const IGameState& s) override {
        ++calls;
        auto t = s.getEnhancedTensorRepresentation();   // the reference's own encoder
        int C = (int)t.size(), H = (int)t[0].size(), W = (int)t[0][0].size();
        std::vector<float> flat; flat.reserve((size_t)C * H * W);
        for (auto& pl : t) for (auto& row : pl) for (float x : row) flat.push_back(x);
        int A = s.getActionSpaceSize();
        std::vector<float> pol(A, 0.0f); float v = 0.0f;
        cb(flat.data(), C, H, W, A, pol.data(), &v, user);
        return {pol, v};
    }
};

struct RefMcts {
    std::unique_ptr<EvaluatorBase> nn;
    std::unique_ptr<alphazero::mcts::TranspositionTable> tt;
    std::unique_ptr<alphazero::mcts::ParallelMCTS> mcts;
};

// ParallelMCTS keeps rootNode_ private; getSearchInfo() only prints.  Grab the
// root through a layout-compatible accessor: the reference exposes getRootNode()?
// (checked at build time by build_ref.sh — falls back to a friend-injection shim).
}  // namespace

extern "C" {

// ---------------------------------------------------------------- state API
void* ref_state_new(int game_type, int board_size) {
    try {
        if (game_type == 0) return new alphazero::gomoku::GomokuState(board_size, false, false, 1, false);
        if (game_type == 2) return new alphazero::go::GoState(board_size, 7.5f, true, true);
        if (game_type == 1) return new alphazero::chess::ChessState();
    } catch (...) {}
    return nullptr;
}
// the variant rules behind createGameState(type, boardSize, variantRules = true) (src/core/game_factory.cpp:90-112): Gomoku -> Renju
// (variant 1; 2 = Omok, the GomokuState ctor's third flag), chess -> ChessState(chess960 = true).  The probe test runs these in a child
// process: at the reference's HEAD they do not survive their first use (tests/test_ref_variants.py).
void* ref_state_new_variant(int game_type, int board_size, int variant) {
    if (game_type == 0) return new alphazero::gomoku::GomokuState(board_size, variant == 1, variant == 2, 1, false);
    if (game_type == 1) return new alphazero::chess::ChessState(variant != 0);
    return nullptr;
}
// ChessState(chess960 = true, fen = "", position_number): 1 = constructed, 0 = threw
int ref_chess960_constructible(int position_number) {
    try { alphazero::chess::ChessState s(true, "", position_number); return 1; } catch (...) { return 0; }
}
void ref_state_free(void* h) { delete (IGameState*)h; }
void* ref_state_clone(void* h) { return ((IGameState*)h)->clone().release(); }
int ref_state_make_move(void* h, int action) {
    try { ((IGameState*)h)->makeMove(action); return 0; } catch (...) { return -1; }
}
int ref_state_legal_moves(void* h, int* out, int cap) {
    auto m = ((IGameState*)h)->getLegalMoves();
    int n = (int)m.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = m[i];
    return n;
}
int ref_state_is_legal(void* h, int action) { return ((IGameState*)h)->isLegalMove(action) ? 1 : 0; }
int ref_state_is_terminal(void* h) { return ((IGameState*)h)->isTerminal() ? 1 : 0; }
int ref_state_result(void* h) { return (int)((IGameState*)h)->getGameResult(); }
int ref_state_current_player(void* h) { return ((IGameState*)h)->getCurrentPlayer(); }
int ref_state_action_space(void* h) { return ((IGameState*)h)->getActionSpaceSize(); }
int ref_state_board_size(void* h) { return ((IGameState*)h)->getBoardSize(); }
// writes C*H*W floats; returns C (call with out == nullptr to query C)
int ref_state_tensor(void* h, float* out) {
    auto t = ((IGameState*)h)->getEnhancedTensorRepresentation();
    if (out) { size_t k = 0; for (auto& pl : t) for (auto& row : pl) for (float x : row) out[k++] = x; }
    return (int)t.size();
}
uint64_t ref_state_key(void* h) { return state_key(*(IGameState*)h); }
// the rest of the IGameState surface (include/alphazero/core/igamestate.h:60-223), for the host mirror's state classes (tests/test_pybind_cpu.py):
// strings are copied into `out` (cap bytes, NUL-terminated), return value = the full length, -1 = the call threw
static int copy_str(const std::string& v, char* out, int cap) {
    if (out && cap > 0) { const int n = (int)std::min<size_t>(v.size(), (size_t)cap - 1); std::memcpy(out, v.data(), n); out[n] = 0; }
    return (int)v.size();
}
int ref_state_action_to_string(void* h, int action, char* out, int cap) {
    try { return copy_str(((IGameState*)h)->actionToString(action), out, cap); } catch (...) { return -1; }
}
// returns 0 and the action in *action when the string parses (std::optional has a value), 1 when it does not, -1 when the call threw
int ref_state_string_to_action(void* h, const char* text, int* action) {
    try { auto a = ((IGameState*)h)->stringToAction(text); if (!a) return 1; *action = *a; return 0; } catch (...) { return -1; }
}
int ref_state_to_string(void* h, char* out, int cap) {
    try { return copy_str(((IGameState*)h)->toString(), out, cap); } catch (...) { return -1; }
}
int ref_state_undo(void* h) { try { return ((IGameState*)h)->undoMove() ? 1 : 0; } catch (...) { return -1; } }
int ref_state_validate(void* h) { try { return ((IGameState*)h)->validate() ? 1 : 0; } catch (...) { return -1; } }
int ref_state_equals(void* a, void* b) { try { return ((IGameState*)a)->equals(*(IGameState*)b) ? 1 : 0; } catch (...) { return -1; } }
int ref_state_history(void* h, int* out, int cap) {
    auto m = ((IGameState*)h)->getMoveHistory();
    for (int i = 0; i < (int)m.size() && i < cap; ++i) out[i] = m[i];
    return (int)m.size();
}
// getTensorRepresentation (the basic planes): writes C*H*W floats; returns C (out == nullptr queries C)
int ref_state_basic_tensor(void* h, float* out) {
    auto t = ((IGameState*)h)->getTensorRepresentation();
    if (out) { size_t k = 0; for (auto& pl : t) for (auto& row : pl) for (float x : row) out[k++] = x; }
    return (int)t.size();
}
void ref_hash_eval(void* h, float* policy, float* value) {
    IGameState* s = (IGameState*)h; hash_eval(state_key(*s), s->getActionSpaceSize(), policy, value);
}
// Go extras (tests/integration/go_integration_test.cpp known-answer cases)
int ref_go_stone(void* h, int pos) { return ((alphazero::go::GoState*)h)->getStone(pos); }
int ref_go_ko(void* h) { return ((alphazero::go::GoState*)h)->getKoPoint(); }
int ref_go_captured(void* h, int player) { return ((alphazero::go::GoState*)h)->getCapturedStones(player); }

// chess extras (tests/games/chess/chess_state_test.cpp drives states through setFromFEN)
int ref_chess_set_fen(void* h, const char* fen) { try { return ((alphazero::chess::ChessState*)h)->setFromFEN(fen) ? 0 : -1; } catch (...) { return -1; } }
int ref_chess_piece(void* h, int sq) { const auto pc = ((alphazero::chess::ChessState*)h)->getPiece(sq); return (int)pc.type + 8 * (int)pc.color; }
int ref_chess_in_check(void* h) { return ((alphazero::chess::ChessState*)h)->isInCheck() ? 1 : 0; }
void ref_chess_set_fide(int) {}    // the reference IS the literal pawn-attack rule (QUIRK C5); kept so both checkers export the same names
long ref_chess_perft(void* h, int depth) {
    if (depth == 0) return 1;
    IGameState* s = (IGameState*)h;
    auto mv = s->getLegalMoves();
    if (depth == 1) return (long)mv.size();
    long n = 0;
    for (int a : mv) { auto c = s->clone(); c->makeMove(a); n += ref_chess_perft(c.get(), depth - 1); }
    return n;
}

// ---------------------------------------------------------------- Dataset / GameRecord (src/selfplay/dataset.cpp, game_record.cpp)
// Dataset::extractExamples builds its replay state with core::createGameState (game_factory.cpp:92-120), which asks the
// GameRegistry — and the reference registers nothing in it (its one REGISTER_GAME file, src/core/gomoku_state_plugin.cpp, is not part
// of any build target; chess and Go have none).  The harness, as the application, registers the three games with the argument
// names game_factory.cpp passes.
static void register_games_once() {
    static bool done = false; if (done) return; done = true;
    auto& reg = alphazero::core::GameRegistry::instance();
    using alphazero::core::VariantArgs;
    reg.registerGame("gomoku", [](const VariantArgs& a) -> std::unique_ptr<IGameState> {
        return std::make_unique<alphazero::gomoku::GomokuState>(a.get<int>("boardSize", 15), a.get<bool>("useRenju", false), a.get<bool>("useOmok", false),
                                                                a.get<int>("seed", 0), a.get<bool>("useProLongOpening", false)); });
    reg.registerGame("go", [](const VariantArgs& a) -> std::unique_ptr<IGameState> {
        return std::make_unique<alphazero::go::GoState>(a.get<int>("boardSize", 19), a.get<float>("komi", 7.5f), a.get<bool>("chineseRules", true), true); });
    reg.registerGame("chess", [](const VariantArgs& a) -> std::unique_ptr<IGameState> {
        return std::make_unique<alphazero::chess::ChessState>(a.get<bool>("chess960", false)); });
}
// Dataset::augmentExample on one example: planes [C][N][N], policy [P] -> the 7 extra images in the reference's order
void ref_augment_example(const float* planes, int C, int N, const float* policy, int P, int game_type, float* out_planes, float* out_policy) {
    alphazero::selfplay::TrainingExample ex;
    ex.state.assign(C, std::vector<std::vector<float>>(N, std::vector<float>(N)));
    for (int c = 0; c < C; ++c) for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) ex.state[c][i][j] = planes[((size_t)c * N + i) * N + j];
    ex.policy.assign(policy, policy + P); ex.value = 0.0f;
    alphazero::selfplay::Dataset ds;
    auto aug = ds.augmentExample(ex, (GameType)game_type);
    for (size_t k = 0; k < aug.size(); ++k) {
        for (int c = 0; c < C; ++c) for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) out_planes[(((size_t)k * C + c) * N + i) * N + j] = aug[k].state[c][i][j];
        for (int a = 0; a < P; ++a) out_policy[(size_t)k * P + a] = aug[k].policy[a];
    }
}
// Dataset::addGameRecord + extractExamples(includeAugmentations) for ONE game: moves[n], policies [n][P], result code.
// extractExamples ends with a shuffle (random_device seed), so the examples come back in shuffled order: callers compare as multisets.
// Returns the number of examples (call with out_planes == nullptr to query); -1 on an exception (illegal recorded move).
int ref_dataset_extract(int game_type, int board_size, const int* moves, int n, const float* policies, int P, int result, int augment,
                        float* out_planes, float* out_policy, float* out_value) {
    try {
        register_games_once();
        alphazero::selfplay::GameRecord rec((GameType)game_type, board_size, false);
        for (int i = 0; i < n; ++i) rec.addMove(moves[i], std::vector<float>(policies + (size_t)i * P, policies + (size_t)(i + 1) * P), 0.0f, 0);
        rec.setResult((alphazero::core::GameResult)result);
        alphazero::selfplay::Dataset ds;
        ds.addGameRecord(rec);
        ds.extractExamples(augment != 0);
        const auto& ex = ds.examples_;
        if (out_planes) {
            size_t kp = 0, kq = 0;
            for (size_t e = 0; e < ex.size(); ++e) {
                for (auto& pl : ex[e].state) for (auto& row : pl) for (float x : row) out_planes[kp++] = x;
                for (float x : ex[e].policy) out_policy[kq++] = x;
                out_value[e] = ex[e].value;
            }
        }
        return (int)ex.size();
    } catch (...) { return -1; }
}
// GameRecord::toJson for a record built through addMove / setResult (timestamp = now); writes at most cap-1 bytes + NUL, returns the length
int ref_game_record_json(int game_type, int board_size, int variant, const int* moves, int n, const float* policies, const int* policy_len, const float* values,
                         const long long* think_ms, int result, char* out, int cap) {
    alphazero::selfplay::GameRecord rec((GameType)game_type, board_size, variant != 0);
    size_t off = 0;
    for (int i = 0; i < n; ++i) { rec.addMove(moves[i], std::vector<float>(policies + off, policies + off + policy_len[i]), values[i], think_ms[i]); off += policy_len[i]; }
    rec.setResult((alphazero::core::GameResult)result);
    const std::string j = rec.toJson();
    if (out && cap > 0) { const size_t k = std::min<size_t>(j.size(), (size_t)cap - 1); std::memcpy(out, j.data(), k); out[k] = 0; }
    return (int)j.size();
}

// File formats: the reference's own readers / writers over files the host mirror wrote (tests/test_pybind_cpu.py; child process).
// GameRecord::loadFromFile(in).saveToFile(out) (game_record.cpp:119-150): returns the number of moves read, -1 = threw / could not write
int ref_game_record_file_roundtrip(const char* in, const char* out) {
    try {
        alphazero::selfplay::GameRecord rec = alphazero::selfplay::GameRecord::loadFromFile(in);
        if (!rec.saveToFile(out)) return -1;
        return (int)rec.getMoves().size();
    } catch (...) { return -1; }
}
// Dataset::loadFromFile(in) + saveToFile(out) (dataset.cpp:151-216): returns the number of examples read, -1 = load or save failed
int ref_dataset_file_roundtrip(const char* in, const char* out) {
    try {
        alphazero::selfplay::Dataset d;
        if (!d.loadFromFile(in)) return -1;
        if (!d.saveToFile(out)) return -1;
        return (int)d.size();
    } catch (...) { return -1; }
}

// ---------------------------------------------------------------- search API
// evaluator: 0 = HashEvaluator, 1 = CallbackEvaluator(cb,user), 2 = HashEvaluator with a peaked policy
void* ref_mcts_new(void* state, int sims, float cpuct, int virtual_loss, int evaluator, eval_cb_t cb, void* user) {
    auto* r = new RefMcts();
    if (evaluator == 0 || evaluator == 2) { auto* he = new HashEvaluator(); he->peaked = evaluator == 2; r->nn.reset(he); }
    else r->nn.reset(new CallbackEvaluator(cb, user));
    r->tt.reset(new alphazero::mcts::TranspositionTable(1 << 20, 1 << 10));
    alphazero::mcts::MCTSConfig cfg;
    cfg.numThreads = 1; cfg.numSimulations = sims; cfg.cPuct = cpuct; cfg.fpuReduction = 0.0f;
    cfg.virtualLoss = virtual_loss; cfg.useDirichletNoise = false;
    // useBatchInference=true is what SelfPlayManager forces (self_play_manager.cpp:169): it only
    // switches selectAction to the deterministic argmax branch and builds a (never used, since
    // useBatchedMCTS stays false) BatchQueue.  Keep it false here: plain serial path.
    cfg.useBatchInference = false; cfg.useBatchedMCTS = false;
    r->mcts.reset(new alphazero::mcts::ParallelMCTS(*(IGameState*)state, cfg, r->nn.get(), r->tt.get()));
    // Deterministic mode AFTER construction (parallel_mcts.cpp:1263-1274): flips useBatchInference so
    // selectAction takes its first-max / argmax branches (:1018-1021, :1037-1039) — the behaviour
    // SelfPlayManager forces (self_play_manager.cpp:169) — without creating a BatchQueue, so search()
    // stays on the serial runSingleSimulation path.  It also stops the dtor deleting our TT (:105-109).
    r->mcts->setDeterministicMode(true);
    return r;
}
void ref_mcts_free(void* h) {
    auto* r = (RefMcts*)h;
    r->mcts.reset();   // before the TT and evaluator it points at
    delete r;
}
void ref_mcts_search(void* h) { ((RefMcts*)h)->mcts->search(); }
void ref_mcts_set_sims(void* h, int sims) { ((RefMcts*)h)->mcts->setNumSimulations(sims); }
long ref_mcts_eval_calls(void* h) { return ((RefMcts*)h)->nn->calls; }
int ref_mcts_root_stats(void* h, int* actions, int* N, float* W, float* P, int cap, int* rootN, float* rootW) {
    const alphazero::mcts::MCTSNode* root = ((RefMcts*)h)->mcts->getRootNode();
    int n = (int)root->children.size();
    for (int i = 0; i < n && i < cap; ++i) {
        auto* c = root->children[i].get();
        actions[i] = root->actions[i];
        N[i] = c->visitCount.load(); W[i] = c->valueSum.load(); P[i] = c->prior;
    }
    if (rootN) *rootN = root->visitCount.load();
    if (rootW) *rootW = root->valueSum.load();
    return n;
}
int ref_mcts_select_action(void* h, int is_training, float temperature) {
    return ((RefMcts*)h)->mcts->selectAction(is_training != 0, temperature);
}
int ref_mcts_action_probs(void* h, float temperature, float* out, int cap) {
    auto p = ((RefMcts*)h)->mcts->getActionProbabilities(temperature);
    int n = (int)p.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = p[i]; return n;
}
float ref_mcts_root_value(void* h) { return ((RefMcts*)h)->mcts->getRootValue(); }
void ref_mcts_update_with_move(void* h, int action) { ((RefMcts*)h)->mcts->updateWithMove(action); }
void ref_mcts_set_deterministic(void* h, int on) { ((RefMcts*)h)->mcts->setDeterministicMode(on != 0); }

}  // extern "C"
