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

#include "alphazero/core/igamestate.h"
#include "alphazero/games/gomoku/gomoku_state.h"
#include "alphazero/games/go/go_state.h"
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
    }
    return h;
}

void hash_eval(uint64_t h, int A, float* policy, float* value) {
    float sum = 0.0f;
    for (int i = 0; i < A; ++i) {
        uint64_t r = mix64(h + (uint64_t)i * 0x9E3779B97F4A7C15ULL) >> 40;
        float raw = (float)(r + 1) / (float)(1 << 24);
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
    std::pair<std::vector<float>, float> predict(const IGameState& s) override {
        ++calls;
        int A = s.getActionSpaceSize();
        std::vector<float> pol(A); float v;
        hash_eval(state_key(s), A, pol.data(), &v);
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
    } catch (...) {}
    return nullptr;
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
void ref_hash_eval(void* h, float* policy, float* value) {
    IGameState* s = (IGameState*)h; hash_eval(state_key(*s), s->getActionSpaceSize(), policy, value);
}
// Go extras (tests/integration/go_integration_test.cpp known-answer cases)
int ref_go_stone(void* h, int pos) { return ((alphazero::go::GoState*)h)->getStone(pos); }
int ref_go_ko(void* h) { return ((alphazero::go::GoState*)h)->getKoPoint(); }
int ref_go_captured(void* h, int player) { return ((alphazero::go::GoState*)h)->getCapturedStones(player); }

// ---------------------------------------------------------------- search API
// evaluator: 0 = HashEvaluator, 1 = CallbackEvaluator(cb,user)
void* ref_mcts_new(void* state, int sims, float cpuct, int virtual_loss, int evaluator, eval_cb_t cb, void* user) {
    auto* r = new RefMcts();
    if (evaluator == 0) r->nn.reset(new HashEvaluator()); else r->nn.reset(new CallbackEvaluator(cb, user));
    r->tt.reset(new alphazero::mcts::TranspositionTable(1 << 16, 16));
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
