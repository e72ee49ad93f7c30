// alphazero_host.cpp — see alphazero_host.hpp.  Host logic only; every search / network call goes through the C ABI.
#include "alphazero_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <random>
#include <sstream>
#include <thread>
#include <dlfcn.h>

#include <nlohmann/json.hpp>

#include "../csrc/gomoku.cuh"   // host+device rules header (hash-evaluator key for B200NeuralNetwork("hash").predict)
#include "../csrc/go.cuh"       // host+device Go rules
#include "../csrc/chess.cuh"    // host+device chess rules

namespace az { void set_error(const std::string&) {} }   // common.cuh declares it; unused on the host side

namespace alphazero {

static void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + az_last_error());
}

// ================================================================================================ core
namespace core {
std::unique_ptr<IGameState> createGameState(GameType type, int boardSize, bool variantRules) {
    // reference: src/core/igamestate.cpp:10-69 / src/core/game_factory.cpp:90-112.  variantRules = true does not yield a usable state
    // in the reference either, so it is an error here, raised where the reference would die (tests/test_ref_variants.py probes the compiled
    // reference): Gomoku -> Renju, whose first getLegalMoves() / isLegalMove() for Black recurses without end (renju_double_four_or_more swaps
    // the board accessor for a lambda that calls the accessor, gomoku_rules.cpp:198-220; Omok: omok_check_double_three_strict, :360-397) and
    // overflows the stack; chess -> ChessState(chess960 = true), whose default position number 518 gets both knights on one square from
    // Chess960::getPermutation (chess960.cpp:466-480) and fails its assert / throws "Invalid piece index in Chess960 generation" (:25-76).
    // Go ignores the flag (Chinese rules, komi 7.5 either way).
    if (type == GameType::GOMOKU) {
        if (variantRules) throw GameStateException("Failed to create game state: Renju rules never terminate in the reference (gomoku_rules.cpp:198-220); not provided");
        return std::make_unique<gomoku::GomokuState>(boardSize > 0 ? boardSize : 15, false, false, 0, false);
    }
    if (type == GameType::GO) {
        return std::make_unique<go::GoState>(boardSize > 0 ? boardSize : 19);       // igamestate.cpp:50-55 default 19
    }
    if (type == GameType::CHESS) {
        if (variantRules) throw GameStateException("Failed to create game state: Invalid piece index in Chess960 generation");      // the reference's own message (release build)
        return std::make_unique<chess::ChessState>();
    }
    throw GameStateException("Failed to create game state: unknown game type");
}
std::unique_ptr<IGameState> GameFactory::createGomokuState(int boardSize, bool useRenju, bool useOmok, int seed, bool useProLongOpening) {
    return std::make_unique<gomoku::GomokuState>(boardSize, useRenju, useOmok, seed, useProLongOpening);
}
std::unique_ptr<IGameState> GameFactory::createChessState(bool chess960, const std::string& fen) { return std::make_unique<chess::ChessState>(chess960, fen); }
std::unique_ptr<IGameState> GameFactory::createGoState(int boardSize, float komi, bool chineseRules) { return std::make_unique<go::GoState>(boardSize, komi, chineseRules); }
bool GameFactory::isGameSupported(GameType type) { return type == GameType::GOMOKU || type == GameType::CHESS || type == GameType::GO; }
int GameFactory::getDefaultBoardSize(GameType type) {            // game_factory.cpp:80-87
    switch (type) {
        case GameType::GOMOKU: return 15;
        case GameType::CHESS: return 8;
        case GameType::GO: return 19;
        default: throw std::invalid_argument("Unsupported game type");
    }
}
}  // namespace core

// ================================================================================================ gomoku
namespace gomoku {

GomokuState::GomokuState(int bs, bool use_renju, bool use_omok, int, bool use_pro_long)
    : IGameState(core::GameType::GOMOKU), board_size(bs), current_player(BLACK), cells_((size_t)bs * bs, 0) {
    // Renju / Omok: see core::createGameState — the reference's forbidden-move tests do not return.  Pro-long opening is a constraint on Black's
    // first moves only (gomoku_state.cpp:552-556); not provided.
    if (use_renju || use_omok || use_pro_long) throw core::GameStateException("Renju / Omok forbidden-move rules never terminate in the reference (gomoku_rules.cpp:198-220, 360-397); variants are not provided");
    if (bs < 5 || bs > 19) throw core::GameStateException("unsupported board size");
}
bool GomokuState::is_occupied(int a) const { return a < 0 || a >= (int)cells_.size() || cells_[a] != 0; }
bool GomokuState::is_bit_set(int player_index, int a) const noexcept { return a >= 0 && a < (int)cells_.size() && cells_[a] == player_index + 1; }
int GomokuState::count_total_stones() const noexcept { int n = 0; for (int8_t c : cells_) n += c != 0; return n; }
bool GomokuState::board_equal(const GomokuState& o) const { return o.board_size == board_size && o.current_player == current_player && o.cells_ == cells_; }

std::vector<int> GomokuState::getLegalMoves() const {       // gomoku_state.cpp:518-566 (QUIRK G2: set iteration order)
    if (valid_moves_dirty_) {
        cached_valid_moves_.clear();
        for (int a = 0; a < (int)cells_.size(); ++a) if (!cells_[a]) cached_valid_moves_.insert(a);
        valid_moves_dirty_ = false; never_filled_ = false;
    }
    return std::vector<int>(cached_valid_moves_.begin(), cached_valid_moves_.end());
}
bool GomokuState::isLegalMove(int a) const { return a >= 0 && a < (int)cells_.size() && cells_[a] == 0; }
void GomokuState::makeMove(int a) {                          // gomoku_state.cpp:681-722
    if (a < 0 || a >= (int)cells_.size()) throw std::runtime_error("Move " + std::to_string(a) + " out of range.");
    if (cells_[a]) throw std::runtime_error("Cell " + std::to_string(a) + " is already occupied.");
    cells_[a] = (int8_t)current_player; current_player = 3 - current_player; move_history.push_back(a); valid_moves_dirty_ = true;
}
bool GomokuState::undoMove() {
    if (move_history.empty()) return false;
    int a = move_history.back(); move_history.pop_back();
    cells_[a] = 0; current_player = 3 - current_player; valid_moves_dirty_ = true;
    return true;
}
int GomokuState::winner() const {                            // gomoku_rules.cpp:39-115, BLACK first, black exactly five (QUIRK G3)
    static const int D[4][2] = {{0, 1}, {1, 0}, {1, 1}, {1, -1}};
    const int N = board_size;
    for (int p = 1; p <= 2; ++p)
        for (int a = 0; a < N * N; ++a) {
            if (cells_[a] != p) continue;
            const int x = a / N, y = a % N;
            for (auto& d : D) {
                int len = 1;
                for (int s = -1; s <= 1; s += 2) {
                    int cx = x + s * d[0], cy = y + s * d[1];
                    while (cx >= 0 && cx < N && cy >= 0 && cy < N && cells_[cx * N + cy] == p) { ++len; cx += s * d[0]; cy += s * d[1]; }
                }
                if (p == 1 ? len == 5 : len >= 5) return p;
            }
        }
    return 0;
}
bool GomokuState::isTerminal() const {                       // gomoku_state.cpp:491-521
    if (winner() != 0) return true;
    if (!valid_moves_dirty_) return cached_valid_moves_.empty();
    if ((int)move_history.size() >= board_size * board_size) return true;
    return getLegalMoves().empty();
}
core::GameResult GomokuState::getGameResult() const {
    const int w = winner();
    if (w == 1) return core::GameResult::WIN_PLAYER1;
    if (w == 2) return core::GameResult::WIN_PLAYER2;
    return (int)move_history.size() >= board_size * board_size ? core::GameResult::DRAW : core::GameResult::ONGOING;
}
std::vector<std::vector<std::vector<float>>> GomokuState::getTensorRepresentation() const {   // to_tensor :811-840
    const int N = board_size;
    std::vector<std::vector<std::vector<float>>> t(3, std::vector<std::vector<float>>(N, std::vector<float>(N, 0.0f)));
    for (int a = 0; a < N * N; ++a) {
        if (cells_[a] == current_player) t[0][a / N][a % N] = 1.0f; else if (cells_[a]) t[1][a / N][a % N] = 1.0f;
        if (current_player == BLACK) t[2][a / N][a % N] = 1.0f;
    }
    return t;
}
std::vector<std::vector<std::vector<float>>> GomokuState::getEnhancedTensorRepresentation() const {   // :207-258 (QUIRK G5)
    const int N = board_size;
    auto t = getTensorRepresentation();
    t.resize(11, std::vector<std::vector<float>>(N, std::vector<float>(N, 0.0f)));
    const int n = (int)move_history.size();
    for (int k = 0; k < 3; ++k) {
        // moves attributed to the side to move: last, last-2, last-4; to the other side: last-1, last-3, last-5
        const int mine = n - 1 - 2 * k, theirs = n - 2 - 2 * k;
        const int b = current_player == BLACK ? mine : theirs, w = current_player == BLACK ? theirs : mine;
        if (b >= 0) t[3 + k][move_history[b] / N][move_history[b] % N] = 1.0f;
        if (w >= 0) t[6 + k][move_history[w] / N][move_history[w] % N] = 1.0f;
    }
    for (int x = 0; x < N; ++x) for (int y = 0; y < N; ++y) { t[9][x][y] = (float)x / (N - 1); t[10][x][y] = (float)y / (N - 1); }
    return t;
}
uint64_t GomokuState::getHash() const {   // only equality within a process matters (reference keys are time-seeded)
    uint64_t h = 1469598103934665603ULL;
    for (size_t a = 0; a < cells_.size(); ++a) if (cells_[a]) h ^= az::mix64(0x9E37ULL + a * 2 + (cells_[a] - 1));
    return h ^ az::mix64(0xABCDULL + current_player);
}
std::string GomokuState::actionToString(int a) const {       // gomoku_state.cpp:270-283
    if (a < 0 || a >= board_size * board_size) return "invalid";
    char col = (char)('A' + a % board_size); if (col >= 'I') col++;
    return std::string(1, col) + std::to_string(board_size - a / board_size);
}
std::optional<int> GomokuState::stringToAction(const std::string& s) const {
    if (s.size() < 2 || s.size() > 3) return std::nullopt;
    char col = s[0]; if (col >= 'a' && col <= 'z') col = (char)(col - 'a' + 'A');
    if (col < 'A' || col > 'Z') return std::nullopt;
    if (col >= 'I') col--;
    int y = col - 'A', row;
    try { row = std::stoi(s.substr(1)); } catch (...) { return std::nullopt; }
    int x = board_size - row;
    if (x < 0 || x >= board_size || y < 0 || y >= board_size) return std::nullopt;
    return x * board_size + y;
}
// Column letters of the board printouts: A, B, ... with 'I' left out (gomoku_state.cpp:319-325, go_state.cpp:523-530)
char boardColumnLetter(int i) { const char c = (char)('A' + i); return c >= 'I' ? (char)(c + 1) : c; }
std::string GomokuState::toString() const {                   // layout of gomoku_state.cpp:316-358, character for character
    std::ostringstream ss;
    auto header = [&]() { ss << "  "; for (int y = 0; y < board_size; ++y) ss << ' ' << boardColumnLetter(y); ss << '\n'; };
    header();
    for (int x = 0; x < board_size; ++x) {
        const int row = board_size - x;
        ss << (row < 10 ? " " : "") << row << ' ';
        for (int y = 0; y < board_size; ++y) ss << ".XO"[cells_[x * board_size + y]] << ' ';
        ss << row << '\n';
    }
    header();
    ss << "Current player: " << (current_player == 1 ? "Black (X)" : "White (O)") << '\n';
    return ss.str();
}
bool GomokuState::equals(const core::IGameState& o) const {
    auto* g = dynamic_cast<const GomokuState*>(&o);
    return g && g->board_size == board_size && g->current_player == current_player && g->cells_ == cells_;
}
bool GomokuState::validate() const { return true; }
std::vector<std::vector<int>> GomokuState::get_board() const {
    std::vector<std::vector<int>> b(board_size, std::vector<int>(board_size, 0));
    for (int a = 0; a < board_size * board_size; ++a) b[a / board_size][a % board_size] = cells_[a];
    return b;
}
}  // namespace gomoku

// ================================================================================================ go
namespace go {

struct GoState::Impl {
    virtual ~Impl() = default;
    virtual std::unique_ptr<Impl> clone() const = 0;
    virtual std::vector<int> legal() const = 0;
    virtual bool isLegal(int a) const = 0;
    virtual bool apply(int a) = 0;
    virtual bool terminal() const = 0;
    virtual int result() const = 0;
    virtual int player() const = 0;
    virtual int stone(int pos) const = 0;
    virtual int ko() const = 0;
    virtual uint64_t key() const = 0;
    virtual uint64_t posHash() const = 0;
    virtual void planes(std::vector<std::vector<std::vector<float>>>& t) const = 0;
    virtual void scores(float& black, float& white) const = 0;      // area scores before komi (GoRules::calculateScores, go_rules.cpp:313-359)
};

template <int N>
struct GoImpl : GoState::Impl {
    using G = az::Go<N>;
    typename G::Core c;
    std::vector<uint64_t> hist;                    // superko keys of every non-pass move (position_history_)
    GoImpl() { G::init_core(c); }
    std::unique_ptr<GoState::Impl> clone() const override { return std::make_unique<GoImpl<N>>(*this); }
    bool cellLegal(int a) const { return G::legal_cell(c, a, G::valid_bb(), hist.data(), (int)hist.size(), nullptr, 0); }
    std::vector<int> legal() const override {     // getLegalMoves go_state.cpp:116-154 (QUIRK Go2: pass = -1 first)
        std::vector<int> m{-1};
        for (int a = 0; a < G::CELLS; ++a) if (cellLegal(a)) m.push_back(a);
        return m;
    }
    bool isLegal(int a) const override { return a == -1 || cellLegal(a); }
    bool apply(int a) override {
        if (!isLegal(a)) return false;
        uint64_t k;
        if (G::apply_core(c, a, G::valid_bb(), k)) hist.push_back(k);
        return true;
    }
    bool terminal() const override { return c.passes >= 2; }
    int result() const override { return G::result_core(c, G::valid_bb()); }
    int player() const override { return c.player; }
    int stone(int pos) const override { const int p = G::a2p(pos); return G::get(c.bb[0], p) ? 1 : (G::get(c.bb[1], p) ? 2 : 0); }
    int ko() const override { return c.ko; }
    uint64_t key() const override { return G::key_core(c); }
    uint64_t posHash() const override { return G::pos_key(c.key, c.player, c.ko); }
    void scores(float& black, float& white) const override {      // the two sums G::score compares (stones + empty regions bordered by one colour only)
        const auto valid = G::valid_bb();
        typename G::BB empty;
        for (int i = 0; i < G::NW; ++i) empty.w[i] = valid.w[i] & ~(c.bb[0].w[i] | c.bb[1].w[i]);
        int bs = G::popc(c.bb[0]), ws = G::popc(c.bb[1]);
        while (G::any(empty)) {
            typename G::BB seed = G::zero(); G::setb(seed, G::lowest(empty));
            const auto reg = G::flood(seed, empty, valid);
            const auto edge = G::nb(reg, valid);
            uint64_t tb = 0, tw = 0;
            for (int i = 0; i < G::NW; ++i) { tb |= edge.w[i] & c.bb[0].w[i]; tw |= edge.w[i] & c.bb[1].w[i]; empty.w[i] &= ~reg.w[i]; }
            if (tb && !tw) bs += G::popc(reg); else if (tw && !tb) ws += G::popc(reg);
        }
        black = (float)bs; white = (float)ws;
    }
    void planes(std::vector<std::vector<std::vector<float>>>& t) const override {      // go_state.cpp:349-445, index [plane][y][x]
        t.assign(8, std::vector<std::vector<float>>(N, std::vector<float>(N, 0.0f)));
        const auto valid = G::valid_bb();
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                const int p = y * G::PITCH + x;
                int libs = 0;
                if (G::get(c.bb[0], p) || G::get(c.bb[1], p)) { typename G::BB grp; libs = G::group_libs(c, p, valid, grp); }
                for (int pl = 0; pl < 8; ++pl) t[pl][y][x] = G::feature(c, pl, x, y, libs);
            }
    }
};

static std::unique_ptr<GoState::Impl> makeGoImpl(int n) {
    if (n == 9) return std::make_unique<GoImpl<9>>();
    if (n == 13) return std::make_unique<GoImpl<13>>();
    if (n == 19) return std::make_unique<GoImpl<19>>();
    throw core::GameStateException("unsupported Go board size (9, 13 or 19)");
}

GoState::GoState(int bs, float komi, bool chinese_rules, bool enforce_superko) : IGameState(core::GameType::GO), board_size_(bs) {
    // the state class scores with any komi (host arithmetic); the engine itself is built for komi 7.5 (ParallelMCTS / SelfPlayManager check it)
    if (!chinese_rules || !enforce_superko) throw core::GameStateException("only Chinese rules with positional superko are provided");
    komi_ = komi;
    impl_ = makeGoImpl(bs);
}
GoState::GoState(const GoState& o) : IGameState(core::GameType::GO), board_size_(o.board_size_), impl_(o.impl_->clone()), move_history_(o.move_history_), komi_(o.komi_) { captured_[1] = o.captured_[1]; captured_[2] = o.captured_[2]; }
GoState::~GoState() = default;
std::vector<int> GoState::getLegalMoves() const { return impl_->legal(); }
bool GoState::isLegalMove(int a) const { return impl_->isLegal(a); }
void GoState::makeMove(int a) {                               // go_state.cpp:190-261
    const int mover = impl_->player();
    auto opponentStones = [&]() { int n = 0; for (int p = 0; p < board_size_ * board_size_; ++p) n += impl_->stone(p) == 3 - mover; return n; };
    const int before = a >= 0 ? opponentStones() : 0;
    if (!impl_->apply(a)) throw core::IllegalMoveException("Illegal move attempted", a);
    if (a >= 0) captured_[mover] += before - opponentStones();          // captured_stones_[current_player_] += capturedStones (:245)
    move_history_.push_back(a);
}
bool GoState::undoMove() {                                    // replay (the bitboard state keeps no undo stack)
    if (move_history_.empty()) return false;
    std::vector<int> h(move_history_.begin(), move_history_.end() - 1);
    impl_ = makeGoImpl(board_size_); move_history_.clear(); captured_[1] = captured_[2] = 0;
    for (int a : h) makeMove(a);
    return true;
}
bool GoState::isTerminal() const { return impl_->terminal(); }
core::GameResult GoState::getGameResult() const {            // go_state.cpp:315-335: area scores once two passes ended the game
    if (!impl_->terminal()) return core::GameResult::ONGOING;
    float b = 0.0f, w = 0.0f; impl_->scores(b, w); w += komi_;
    return b > w ? core::GameResult::WIN_PLAYER1 : (w > b ? core::GameResult::WIN_PLAYER2 : core::GameResult::DRAW);
}
int GoState::getCurrentPlayer() const { return impl_->player(); }
std::vector<std::vector<std::vector<float>>> GoState::getEnhancedTensorRepresentation() const { std::vector<std::vector<std::vector<float>>> t; impl_->planes(t); return t; }
uint64_t GoState::getHash() const { return impl_->posHash(); }
uint64_t GoState::hashEvaluatorKey() const { return impl_->key(); }
int GoState::getStone(int pos) const { return impl_->stone(pos); }
int GoState::getKoPoint() const { return impl_->ko(); }
std::string GoState::actionToString(int a) const {            // go_state.cpp:447-466: "pass" or column letter (no I) + row from the bottom
    if (a == -1) return "pass";
    if (a < 0 || a >= board_size_ * board_size_) return "invalid";
    const int x = a % board_size_, y = a / board_size_;
    char col = (char)('A' + x); if (col >= 'I') ++col;
    return std::string(1, col) + std::to_string(board_size_ - y);
}
std::optional<int> GoState::stringToAction(const std::string& s) const {
    if (s == "pass" || s == "PASS" || s == "Pass") return -1;
    if (s.size() < 2) return std::nullopt;
    char col = (char)std::toupper((unsigned char)s[0]);
    if (col == 'I' || col < 'A') return std::nullopt;
    int x = col - 'A'; if (col > 'I') --x;
    int row = 0; try { row = std::stoi(s.substr(1)); } catch (...) { return std::nullopt; }
    const int y = board_size_ - row;
    if (x < 0 || x >= board_size_ || y < 0 || y >= board_size_) return std::nullopt;
    return y * board_size_ + x;
}
std::string GoState::toString() const {                       // layout of go_state.cpp:519-603 (no dead-stone marks: nothing in the self-play path sets them)
    std::ostringstream ss;
    auto header = [&]() { ss << "   "; for (int x = 0; x < board_size_; ++x) ss << gomoku::boardColumnLetter(x) << ' '; ss << '\n'; };
    header();
    const int ko = getKoPoint();
    for (int y = 0; y < board_size_; ++y) {
        ss << std::setw(2) << (board_size_ - y) << ' ';
        for (int x = 0; x < board_size_; ++x) {
            const int pos = y * board_size_ + x, st = getStone(pos);
            ss << (st == 1 ? "X " : st == 2 ? "O " : pos == ko ? "k " : ". ");
        }
        ss << (board_size_ - y) << '\n';
    }
    header();
    ss << "Current player: " << (getCurrentPlayer() == 1 ? "Black" : "White") << '\n';
    ss << "Captures - Black: " << captured_[1] << ", White: " << captured_[2] << '\n';
    ss << "Komi: " << komi_ << '\n' << "Rules: Chinese" << '\n' << "Superko enforcement: Yes" << '\n';
    if (isTerminal()) {
        float b = 0.0f, w = 0.0f; impl_->scores(b, w); w += komi_;
        ss << "Game over!" << '\n' << "Final score - Black: " << b << ", White: " << w << " (with komi " << komi_ << ")" << '\n';
        if (b > w) ss << "Black wins by " << (b - w) << " points" << '\n';
        else if (w > b) ss << "White wins by " << (w - b) << " points" << '\n';
        else ss << "Game ended in a draw" << '\n';
    }
    return ss.str();
}
std::vector<std::vector<std::vector<float>>> GoState::getTensorRepresentation() const {      // go_state.cpp:349-378: black, white, side to move
    std::vector<std::vector<std::vector<float>>> t(3, std::vector<std::vector<float>>(board_size_, std::vector<float>(board_size_, 0.0f)));
    const float turn = getCurrentPlayer() == 1 ? 1.0f : 0.0f;
    for (int y = 0; y < board_size_; ++y)
        for (int x = 0; x < board_size_; ++x) {
            const int st = getStone(y * board_size_ + x);
            if (st == 1) t[0][y][x] = 1.0f; else if (st == 2) t[1][y][x] = 1.0f;
            t[2][y][x] = turn;
        }
    return t;
}
bool GoState::equals(const core::IGameState& o) const {
    auto* g = dynamic_cast<const GoState*>(&o);
    if (!g || g->board_size_ != board_size_ || g->getCurrentPlayer() != getCurrentPlayer() || g->getKoPoint() != getKoPoint()) return false;
    for (int p = 0; p < board_size_ * board_size_; ++p) if (g->getStone(p) != getStone(p)) return false;
    return true;
}

}  // namespace go

// ================================================================================================ chess
namespace chess {

struct ChessState::Impl { az::Chess::State s; };

ChessState::ChessState(bool chess960, const std::string& fen, int) : IGameState(core::GameType::CHESS), impl_(new Impl()) {
    if (chess960) throw core::GameStateException("Chess960 set-up is not provided (the reference's own fails for its default position, DESIGN.md 9)");
    az::Chess::init(impl_->s);
    if (!fen.empty() && !setFromFEN(fen)) throw core::GameStateException("Invalid FEN string");       // chess_state.cpp:118-122
}
Piece ChessState::getPiece(int sq) const {
    Piece p; const int c = getPieceCode(sq);
    if (c) { p.type = static_cast<PieceType>(c & 7); p.color = static_cast<PieceColor>(c >> 3); }
    return p;
}
CastlingRights ChessState::getCastlingRights() const {
    const int r = impl_->s.c.rights; CastlingRights cr;
    cr.white_kingside = r & az::Chess::R_WK; cr.white_queenside = r & az::Chess::R_WQ; cr.black_kingside = r & az::Chess::R_BK; cr.black_queenside = r & az::Chess::R_BQ;
    return cr;
}
int ChessState::getEnPassantSquare() const { return impl_->s.c.ep; }
int ChessState::getHalfmoveClock() const { return impl_->s.c.half; }
int ChessState::getFullmoveNumber() const { return 1 + impl_->s.c.ply / 2; }
// chess_state.cpp:382-495: six fields; a malformed string returns false (the reference leaves a half-filled board behind; this one keeps the old
// position).  The position becomes the first entry of the repetition history (recordPosition()), the move history starts empty.
bool ChessState::setFromFEN(const std::string& fen) {
    std::istringstream ss(fen);
    std::string board, active, castling, ep, half, full;
    if (!(ss >> board >> active >> castling >> ep >> half >> full)) return false;
    az::Chess::State st;
    az::Chess::init(st);
    for (int sq = 0; sq < 64; ++sq) az::Chess::put(st.c, sq, 0, 0);
    int rank = 0, file = 0;
    for (char ch : board) {
        if (ch == '/') { ++rank; file = 0; continue; }
        if (std::isdigit((unsigned char)ch)) { file += ch - '0'; continue; }
        if (file >= 8 || rank >= 8) return false;
        int t = 0;
        switch (std::tolower((unsigned char)ch)) { case 'p': t = 1; break; case 'n': t = 2; break; case 'b': t = 3; break; case 'r': t = 4; break; case 'q': t = 5; break; case 'k': t = 6; break; default: return false; }
        az::Chess::put(st.c, rank * 8 + file, t, std::isupper((unsigned char)ch) ? (int)az::Chess::WHITE : (int)az::Chess::BLACK);
        ++file;
    }
    st.c.player = (int8_t)(active == "w" ? az::Chess::WHITE : az::Chess::BLACK);
    int rights = 0;
    for (char ch : castling) rights |= ch == 'K' ? az::Chess::R_WK : ch == 'Q' ? az::Chess::R_WQ : ch == 'k' ? az::Chess::R_BK : ch == 'q' ? az::Chess::R_BQ : 0;
    st.c.rights = (int8_t)rights;
    st.c.ep = -1;
    if (ep != "-") { if (ep.size() < 2 || ep[0] < 'a' || ep[0] > 'h' || ep[1] < '1' || ep[1] > '8') st.c.ep = -1; else st.c.ep = (int16_t)(('8' - ep[1]) * 8 + (ep[0] - 'a')); }
    int hm = 0, fm = 1;
    try { hm = std::stoi(half); fm = std::stoi(full); } catch (...) { return false; }
    st.c.half = (int16_t)hm;
    st.c.ply = (int16_t)(2 * std::max(fm - 1, 0) + (st.c.player == az::Chess::BLACK ? 1 : 0));
    st.hist[0] = st.c.key; st.c.hist_n = 1;
    impl_->s = st;
    move_history_.clear();
    fen_start_ = fen;
    return true;
}
ChessState::ChessState(const ChessState& o) : IGameState(core::GameType::CHESS), impl_(new Impl(*o.impl_)), move_history_(o.move_history_), fen_start_(o.fen_start_) {}
ChessState::~ChessState() = default;
std::vector<int> ChessState::getLegalMoves() const {           // chess_state.cpp:498-510 over generateLegalMoves
    int16_t lg[az::Chess::MAX_CHILDREN]; const int n = az::Chess::host_legal(impl_->s, lg);
    return std::vector<int>(lg, lg + n);
}
bool ChessState::isLegalMove(int a) const { for (int m : getLegalMoves()) if (m == a) return true; return false; }
void ChessState::makeMove(int a) {                              // chess_state.cpp:976-979: "Illegal move attempted"
    if (!az::Chess::host_apply(impl_->s, a)) throw core::IllegalMoveException("Illegal move attempted", a);
    move_history_.push_back(a);
}
bool ChessState::undoMove() {
    if (move_history_.empty()) return false;
    std::vector<int> h(move_history_.begin(), move_history_.end() - 1);
    if (fen_start_.empty()) { az::Chess::init(impl_->s); move_history_.clear(); }
    else { const std::string f = fen_start_; setFromFEN(f); }
    for (int a : h) makeMove(a);
    return true;
}
static int chessResult(const az::Chess::State& s) {
    int16_t lg[az::Chess::MAX_CHILDREN]; const int n = az::Chess::host_legal(s, lg);
    return az::Chess::result_core(s.c, n, az::Chess::repetitions(s.c, s.hist, nullptr));
}
bool ChessState::isTerminal() const { return chessResult(impl_->s) != az::RES_ONGOING; }
core::GameResult ChessState::getGameResult() const { return static_cast<core::GameResult>(chessResult(impl_->s)); }
int ChessState::getCurrentPlayer() const { return impl_->s.c.player; }
static std::vector<std::vector<std::vector<float>>> chessPlanes(const az::Chess::State& s, int n) {
    std::vector<std::vector<std::vector<float>>> t(n, std::vector<std::vector<float>>(8, std::vector<float>(8, 0.0f)));
    const int reps = az::Chess::repetitions(s.c, s.hist, nullptr);
    for (int pl = 0; pl < n; ++pl) for (int sq = 0; sq < 64; ++sq) t[pl][sq >> 3][sq & 7] = az::Chess::feature(s.c, pl, sq, reps);
    return t;
}
std::vector<std::vector<std::vector<float>>> ChessState::getTensorRepresentation() const { return chessPlanes(impl_->s, 12); }
std::vector<std::vector<std::vector<float>>> ChessState::getEnhancedTensorRepresentation() const { return chessPlanes(impl_->s, 18); }
uint64_t ChessState::getHash() const { return impl_->s.c.key; }
uint64_t ChessState::hashEvaluatorKey() const { return az::Chess::key_core(impl_->s.c); }
int ChessState::getPieceCode(int sq) const { return (sq >= 0 && sq < 64) ? impl_->s.c.b[sq] : 0; }
bool ChessState::isInCheck() const { return az::Chess::in_check(impl_->s.c, impl_->s.c.player); }
static std::string sqName(int sq) { return std::string(1, (char)('a' + (sq & 7))) + std::string(1, (char)('8' - (sq >> 3))); }
std::string ChessState::actionToString(int a) const {           // chess_state.cpp:1239-1290: from + to + promotion letter
    if (a < 0 || a >= getActionSpaceSize()) return "invalid";
    std::string s = sqName((a >> 6) & 63) + sqName(a & 63);
    const int pc = (a >> 12) & 7; if (pc >= 1 && pc <= 4) s += "qrbn"[pc - 1];
    return s;
}
std::optional<int> ChessState::stringToAction(const std::string& s) const {
    if (s.size() < 4) return std::nullopt;
    auto sq = [](char f, char r) -> int { return (f < 'a' || f > 'h' || r < '1' || r > '8') ? -1 : ('8' - r) * 8 + (f - 'a'); };
    const int from = sq(s[0], s[1]), to = sq(s[2], s[3]);
    if (from < 0 || to < 0) return std::nullopt;
    int pc = 0;
    if (s.size() >= 5) { const char c = (char)std::tolower((unsigned char)s[4]); pc = c == 'q' ? 1 : c == 'r' ? 2 : c == 'b' ? 3 : c == 'n' ? 4 : 0; }
    return (pc << 12) | (from << 6) | to;
}
std::string ChessState::toFEN() const {                       // chess_state.cpp:270-378 (standard castling letters)
    std::ostringstream ss; const char* names = ".pnbrqk";
    const auto& c = impl_->s.c;
    for (int r = 0; r < 8; ++r) {
        int run = 0;
        for (int f = 0; f < 8; ++f) {
            const int p = c.b[r * 8 + f];
            if (!p) { ++run; continue; }
            if (run) { ss << run; run = 0; }
            char ch = names[p & 7]; if ((p >> 3) == 1) ch = (char)std::toupper((unsigned char)ch);
            ss << ch;
        }
        if (run) ss << run;
        if (r < 7) ss << '/';
    }
    ss << ' ' << (c.player == 1 ? 'w' : 'b') << ' ';
    if (c.rights & az::Chess::R_WK) ss << 'K';
    if (c.rights & az::Chess::R_WQ) ss << 'Q';
    if (c.rights & az::Chess::R_BK) ss << 'k';
    if (c.rights & az::Chess::R_BQ) ss << 'q';
    if (!(c.rights & 15)) ss << '-';
    ss << ' ' << ((c.ep >= 0 && c.ep < 64) ? sqName(c.ep) : std::string("-")) << ' ' << c.half << ' ' << (1 + c.ply / 2);
    return ss.str();
}
std::string ChessState::toString() const {                    // layout of chess_state.cpp:795-867
    std::ostringstream ss; const char* names = ".pnbrqk";
    const auto& c = impl_->s.c;
    ss << "  a b c d e f g h" << '\n';
    for (int r = 0; r < 8; ++r) {
        ss << (8 - r) << ' ';
        for (int f = 0; f < 8; ++f) { const int p = c.b[r * 8 + f]; char ch = names[p & 7]; if ((p >> 3) == 1) ch = (char)std::toupper((unsigned char)ch); ss << ch << ' '; }
        ss << (8 - r) << '\n';
    }
    ss << "  a b c d e f g h" << '\n';
    ss << "Current player: " << (c.player == 1 ? "White" : "Black") << '\n' << "Castling rights: ";
    if (c.rights & az::Chess::R_WK) ss << 'K';
    if (c.rights & az::Chess::R_WQ) ss << 'Q';
    if (c.rights & az::Chess::R_BK) ss << 'k';
    if (c.rights & az::Chess::R_BQ) ss << 'q';
    if (!(c.rights & 15)) ss << '-';
    ss << '\n' << "En passant square: " << ((c.ep >= 0 && c.ep < 64) ? sqName(c.ep) : std::string("-")) << '\n';
    ss << "Halfmove clock: " << c.half << '\n' << "Fullmove number: " << (1 + c.ply / 2) << '\n';
    ss << "FEN: " << toFEN() << '\n';
    return ss.str();
}
bool ChessState::equals(const core::IGameState& o) const {
    auto* c = dynamic_cast<const ChessState*>(&o);
    if (!c) return false;
    const auto &a = impl_->s.c, &b = c->impl_->s.c;
    return std::memcmp(a.b, b.b, 64) == 0 && a.player == b.player && a.rights == b.rights && a.ep == b.ep;
}

}  // namespace chess

// ================================================================================================ nn
namespace nn {

// "hash-host": the hash evaluator as a PLAIN nn::NeuralNetwork that computes on the host (it forwards to B200NeuralNetwork("hash")'s host
// arithmetic but is not a B200NeuralNetwork): what an arbitrary user-supplied evaluator looks like to ParallelMCTS, which then runs it
// through the external-evaluator path.  Same numbers as the device hash evaluator, so the two paths can be compared bit for bit.
class HostHashNetwork : public NeuralNetwork {
public:
    HostHashNetwork(core::GameType gt, int bs) : inner_("hash", gt, bs) {}
    std::pair<std::vector<float>, float> predict(const core::IGameState& s) override { ++calls_; return inner_.predict(s); }
    void predictBatch(const std::vector<std::reference_wrapper<const core::IGameState>>& st, std::vector<std::vector<float>>& p, std::vector<float>& v) override {
        calls_ += (long)st.size(); inner_.predictBatch(st, p, v);
    }
    bool isGpuAvailable() const override { return false; }
    std::string getDeviceInfo() const override { return "host hash evaluator (external-evaluator path), " + std::to_string(calls_) + " evaluations"; }
    float getInferenceTimeMs() const override { return 0.0f; }
    int getBatchSize() const override { return 1; }
    std::string getModelInfo() const override { return "HostHashNetwork"; }
    size_t getModelSizeBytes() const override { return 0; }
    void benchmark(int, int) override {}
    void enableDebugMode(bool) override {}
    void printModelSummary() const override {}
private:
    B200NeuralNetwork inner_; long calls_ = 0;
};
std::unique_ptr<NeuralNetwork> NeuralNetwork::create(const std::string& modelPath, core::GameType gameType, int boardSize, bool) {
    if (modelPath == "hash-host") return std::make_unique<HostHashNetwork>(gameType, boardSize);
    return std::make_unique<B200NeuralNetwork>(modelPath, gameType, boardSize);
}

B200NeuralNetwork::B200NeuralNetwork(const std::string& modelPath, core::GameType gameType, int boardSize)
    : gameType_(gameType), boardSize_(boardSize > 0 ? boardSize : 15) {
    if (modelPath.empty() || modelPath == "hash") { hash_ = true; return; }
    std::ifstream f(modelPath, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open weight blob " + modelPath);
    blob_.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (blob_.size() < 32 || std::memcmp(blob_.data(), "AZW1", 4) != 0) throw std::runtime_error(modelPath + " is not an AZW1 weight blob (net.py:export_weights)");
    int32_t hdr[7]; std::memcpy(hdr, blob_.data() + 4, sizeof(hdr));
    blocks_ = hdr[1]; channels_ = hdr[2];
}
B200NeuralNetwork::~B200NeuralNetwork() { if (eng_) az_engine_destroy(eng_); }

void B200NeuralNetwork::ensureEngine() {
    if (eng_ || hash_) return;
    az_config c; az_config_default(&c);
    c.game = (int)gameType_; c.board_size = boardSize_; c.n_slots = batch_; c.evaluator = AZ_EVAL_RESNET; c.net_blocks = blocks_; c.net_channels = channels_;
    c.num_simulations = 1; c.max_nodes_per_tree = 1024;
    check(az_engine_create(&c, &eng_), "az_engine_create");
    check(az_engine_load_weights(eng_, blob_.data(), blob_.size()), "az_engine_load_weights");
}
std::pair<std::vector<float>, float> B200NeuralNetwork::predict(const core::IGameState& state) {
    std::vector<std::vector<float>> p; std::vector<float> v;
    predictBatch({std::cref(state)}, p, v);
    return {p[0], v[0]};
}
void B200NeuralNetwork::predictBatch(const std::vector<std::reference_wrapper<const core::IGameState>>& states,
                                     std::vector<std::vector<float>>& policies, std::vector<float>& values) {
    const int n = (int)states.size();
    policies.assign(n, {}); values.assign(n, 0.0f);
    if (n == 0) return;
    const int A = states[0].get().getActionSpaceSize();
    if (hash_) {   // stateless HashEvaluator on the canonical bitboards
        for (int i = 0; i < n; ++i) {
            const auto& s = states[i].get();
            uint64_t h = 0;
            if (auto* gs_go = dynamic_cast<const go::GoState*>(&s)) h = gs_go->hashEvaluatorKey();
            else if (auto* gs_ch = dynamic_cast<const chess::ChessState*>(&s)) h = gs_ch->hashEvaluatorKey();
            else if (s.getGameType() == core::GameType::GOMOKU && s.getBoardSize() == 15) {
                az::Gomoku<15>::State gs; az::Gomoku<15>::init(gs);
                for (int a : s.getMoveHistory()) az::Gomoku<15>::apply(gs, a);
                h = az::Gomoku<15>::key(gs);
            } else if (s.getGameType() == core::GameType::GOMOKU && s.getBoardSize() == 9) {
                az::Gomoku<9>::State gs; az::Gomoku<9>::init(gs);
                for (int a : s.getMoveHistory()) az::Gomoku<9>::apply(gs, a);
                h = az::Gomoku<9>::key(gs);
            } else throw std::runtime_error("hash evaluator: Gomoku 15x15 / 9x9 and Go 9 / 13 / 19 on the host side");
            std::vector<float> pol(A); float sum = 0.0f;
            for (int a = 0; a < A; ++a) { pol[a] = az::fdiv((float)((az::mix64(h + (uint64_t)a * 0x9E3779B97F4A7C15ULL) >> 40) + 1), 16777216.0f); sum = az::fadd(sum, pol[a]); }
            for (int a = 0; a < A; ++a) pol[a] = az::fdiv(pol[a], sum);
            float v = az::fdiv((float)(az::mix64(h ^ 0xABCDEFULL) >> 40), 16777216.0f);
            policies[i] = pol; values[i] = az::fmul(az::fsub(az::fmul(v, 2.0f), 1.0f), 0.5f);
        }
        return;
    }
    ensureEngine();
    const auto t0 = std::chrono::steady_clock::now();
    for (int o = 0; o < n; o += batch_) {
        const int c = std::min(batch_, n - o);
        std::vector<float> planes;
        int C = 0;
        for (int i = 0; i < c; ++i) {
            auto t = states[o + i].get().getEnhancedTensorRepresentation();
            C = (int)t.size();
            for (auto& pl : t) for (auto& row : pl) planes.insert(planes.end(), row.begin(), row.end());
        }
        (void)C;
        std::vector<float> pol((size_t)c * A), val(c);
        check(az_engine_nn_forward(eng_, planes.data(), c, pol.data(), val.data(), nullptr), "az_engine_nn_forward");
        for (int i = 0; i < c; ++i) { policies[o + i].assign(pol.begin() + (size_t)i * A, pol.begin() + (size_t)(i + 1) * A); values[o + i] = val[i]; }
    }
    lastMs_ = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
std::string B200NeuralNetwork::getDeviceInfo() const { return "NVIDIA B200 (sm_100a) via libaz_b200"; }
std::string B200NeuralNetwork::getModelInfo() const {
    return hash_ ? std::string("HashEvaluator (stateless, parity runs)") : ("ResNet " + std::to_string(blocks_) + "x" + std::to_string(channels_) + " bf16/tcgen05");
}
void B200NeuralNetwork::benchmark(int numIterations, int batchSize) {
    if (hash_) return;
    ensureEngine();
    float ms = 0; check(az_engine_nn_bench(eng_, std::min(batchSize, batch_), numIterations, &ms), "az_engine_nn_bench");
    lastMs_ = ms;
}
void B200NeuralNetwork::printModelSummary() const { std::cout << getModelInfo() << std::endl; }

}  // namespace nn

// ================================================================================================ mcts
namespace mcts {

ParallelMCTS::ParallelMCTS(const core::IGameState& rootState, nn::NeuralNetwork* nn, TranspositionTable* tt, int numThreads, int numSimulations,
                           float cPuct, float fpuReduction, int virtualLoss) : nn_(nn), tt_(tt) {
    config_.numThreads = numThreads; config_.numSimulations = numSimulations; config_.cPuct = cPuct; config_.fpuReduction = fpuReduction; config_.virtualLoss = virtualLoss;
    build(rootState);
}
ParallelMCTS::ParallelMCTS(const core::IGameState& rootState, const MCTSConfig& config, nn::NeuralNetwork* nn, TranspositionTable* tt) : config_(config), nn_(nn), tt_(tt) { build(rootState); }
ParallelMCTS::~ParallelMCTS() { if (eng_) az_engine_destroy(eng_); }

void ParallelMCTS::build(const core::IGameState& rootState) {
    auto* b = dynamic_cast<nn::B200NeuralNetwork*>(nn_);
    if (!nn_) throw std::runtime_error("ParallelMCTS on the B200 engine needs a neural network (createNeuralNetwork)");
    // Any other nn::NeuralNetwork implementation (neural_network.h:19-131) evaluates the leaves on the host: the engine hands over each wave's
    // leaves as move sequences, the states are rebuilt from the root clone and go through nn_->predictBatch (one round trip per wave).
    external_ = b == nullptr;
    rootState_ = rootState.clone();                                   // parallel_mcts.cpp:65
    // the engine's roots are move sequences from the game's initial position under the engine's rules
    if (auto* g = dynamic_cast<const go::GoState*>(&rootState)) if (g->getKomi() != 7.5f) throw std::runtime_error("ParallelMCTS on the B200 engine: Go is built for komi 7.5");
    if (auto* ch = dynamic_cast<const chess::ChessState*>(&rootState)) if (ch->isFromFEN()) throw std::runtime_error("ParallelMCTS on the B200 engine: chess roots set up from a FEN are not supported (roots are move sequences from the initial position)");
    az_config c; az_config_default(&c);
    c.game = (int)rootState.getGameType(); c.board_size = rootState.getBoardSize(); c.n_slots = 1;
    c.num_simulations = config_.numSimulations; c.c_puct = config_.cPuct; c.virtual_loss = config_.virtualLoss;
    c.evaluator = external_ ? AZ_EVAL_EXTERNAL : (b->isHash() ? AZ_EVAL_HASH : AZ_EVAL_RESNET);
    if (b) { c.net_blocks = b->blocks(); c.net_channels = b->channels(); }
    c.deterministic = config_.useDirichletNoise ? 0 : 1; c.dirichlet_alpha = config_.dirichletAlpha; c.dirichlet_epsilon = config_.dirichletEpsilon;
    c.auto_restart = 0; c.n_streams = 1;
    // the TranspositionTable of the reference (own table of config.transpositionTableSize when none is passed, parallel_mcts.cpp:86-91) = the
    // device evaluation cache; one tree adds numSimulations entries per move, so the request is capped at 64 K entries per ParallelMCTS
    if (!external_ && config_.transpositionTableSize > 0) c.eval_cache_entries = (int)std::max<size_t>(64, std::min<size_t>(tt_ ? tt_->getSize() : (size_t)config_.transpositionTableSize, 65536));
    else c.eval_cache_entries = -1;
    check(az_engine_create(&c, &eng_), "az_engine_create");
    if (external_) check(az_engine_set_external_evaluator(eng_, &ParallelMCTS::evalTrampoline, this), "az_engine_set_external_evaluator");
    else if (!b->isHash()) check(az_engine_load_weights(eng_, b->blob().data(), b->blob().size()), "az_engine_load_weights");
    // The root MCTSNode ctor calls state->isTerminal() (mcts_node.cpp:24), which is what first enumerates the legal moves
    // of a fresh lineage; the order that enumeration produces is the root's child order (QUIRK G2).
    rootState_->isTerminal();
    std::vector<int> order = rootState_->getLegalMoves();
    const std::vector<int> hist = rootState_->getMoveHistory();
    std::vector<int32_t> moves(hist.begin(), hist.end());
    std::vector<int32_t> ord(order.begin(), order.end());
    const bool firstFill = rootState.getGameType() == core::GameType::GOMOKU;      // only Gomoku's child order depends on the lineage
    check(az_engine_set_root(eng_, 0, moves.data(), (int)moves.size(), firstFill ? ord.data() : nullptr, firstFill ? (int)ord.size() : 0), "az_engine_set_root");
}
void ParallelMCTS::setCPuct(float c) { config_.cPuct = c; check(az_engine_set_search_params(eng_, config_.cPuct, config_.virtualLoss), "az_engine_set_search_params"); }
void ParallelMCTS::setVirtualLoss(int v) { config_.virtualLoss = v; check(az_engine_set_search_params(eng_, config_.cPuct, config_.virtualLoss), "az_engine_set_search_params"); }
int ParallelMCTS::evalTrampoline(int n, const int32_t*, const int32_t* paths, const int32_t* lens, int maxLen, int actions, float* policy, float* value, void* user) {
    auto* self = static_cast<ParallelMCTS*>(user);
    try {
        std::vector<std::unique_ptr<core::IGameState>> states; states.reserve(n);
        std::vector<std::reference_wrapper<const core::IGameState>> refs;
        for (int i = 0; i < n; ++i) {
            auto s = self->rootState_->clone();
            for (int k = 0; k < lens[i]; ++k) s->makeMove(paths[(size_t)i * maxLen + k]);
            refs.push_back(std::cref(*s)); states.push_back(std::move(s));
        }
        std::vector<std::vector<float>> pol; std::vector<float> val;
        self->nn_->predictBatch(refs, pol, val);
        if ((int)pol.size() != n || (int)val.size() != n) throw std::runtime_error("predictBatch returned the wrong number of results");
        for (int i = 0; i < n; ++i) {
            const int m = std::min<int>((int)pol[i].size(), actions);
            for (int a = 0; a < m; ++a) policy[(size_t)i * actions + a] = pol[i][a];
            value[i] = val[i];
        }
        return 0;
    } catch (const std::exception& e) { self->evalError_ = e.what(); return 1; }
}
void ParallelMCTS::setSelectionStrategy(MCTSNodeSelection s) {
    if (s != MCTSNodeSelection::PUCT) throw std::runtime_error("the B200 engine implements PUCT selection only (the reference's default, mcts_node.cpp:61-119)");
    config_.selectionStrategy = s;
}
void ParallelMCTS::setConfig(const MCTSConfig& config) {
    if (config.selectionStrategy != MCTSNodeSelection::PUCT) throw std::runtime_error("the B200 engine implements PUCT selection only");
    config_ = config;
    check(az_engine_set_search_params(eng_, config_.cPuct, config_.virtualLoss), "az_engine_set_search_params");
}
void ParallelMCTS::setNeuralNetwork(nn::NeuralNetwork* nn) {
    auto* nb = dynamic_cast<nn::B200NeuralNetwork*>(nn);
    auto* ob = dynamic_cast<nn::B200NeuralNetwork*>(nn_);
    if (!nn) throw std::runtime_error("setNeuralNetwork: null network");
    if (external_ && !nb) { nn_ = nn; return; }                          // host evaluator swapped for another host evaluator
    if (!nb || !ob) throw std::runtime_error("setNeuralNetwork: a device evaluator can only be replaced by a device evaluator of the same shape (and a host one by a host one)");
    if (nb->isHash() != ob->isHash() || nb->blocks() != ob->blocks() || nb->channels() != ob->channels())
        throw std::runtime_error("setNeuralNetwork: the new network must have the shape the engine was built for (same evaluator kind, blocks, channels)");
    nn_ = nn;
    if (!nb->isHash()) check(az_engine_load_weights(eng_, nb->blob().data(), nb->blob().size()), "az_engine_load_weights");
}
MCTSNode ParallelMCTS::getNode(const std::vector<int>& path) const {
    MCTSNode nd; const int cap = rootState_->getActionSpaceSize() + 1;
    std::vector<int32_t> p(path.begin(), path.end()), a(cap), n(cap); nd.childValueSums.resize(cap); nd.childPriors.resize(cap);
    int32_t cnt = cap, nv = 0, fl = 0;
    check(az_engine_node_stats(eng_, 0, p.data(), (int)p.size(), a.data(), n.data(), nd.childValueSums.data(), nd.childPriors.data(), &cnt, &nv, &nd.valueSum, &nd.prior, &fl),
          "az_engine_node_stats");
    nd.actions.assign(a.begin(), a.begin() + cnt); nd.childVisits.assign(n.begin(), n.begin() + cnt); nd.childValueSums.resize(cnt); nd.childPriors.resize(cnt);
    nd.visitCount = nv; nd.isExpanded = cnt > 0; nd.isTerminal = (fl & 1) != 0; nd.gameResult = (core::GameResult)((fl >> 1) & 3);
    return nd;
}
void ParallelMCTS::printSearchPath(int action) const {
    MCTSNode root = getNode();
    if (std::find(root.actions.begin(), root.actions.end(), action) == root.actions.end()) { std::cout << "Action " << action << " not found in children" << std::endl; return; }
    MCTSNode child = getNode({action});
    std::cout << "Path for action " << action << ":\n  Root: V=" << root.visitCount << ", Q=" << std::fixed << std::setprecision(3) << root.getValue()
              << "\n  Child: V=" << child.visitCount << ", Q=" << child.getValue() << ", P=" << child.prior << std::endl;
    if (child.isExpanded) {
        std::cout << "  Grandchildren:" << std::endl;
        std::vector<size_t> idx(child.actions.size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return child.childVisits[x] > child.childVisits[y]; });
        for (size_t k = 0; k < std::min<size_t>(5, idx.size()); ++k) {
            const size_t i = idx[k];
            std::cout << "    Action " << child.actions[i] << ": V=" << child.childVisits[i] << ", Q=" << (child.childVisits[i] ? child.childValueSums[i] / child.childVisits[i] : 0.0f)
                      << ", P=" << child.childPriors[i] << std::endl;
        }
    }
}
float MCTSNode::getTerminalValue(int currentPlayer) const {
    if (gameResult == core::GameResult::WIN_PLAYER1) return currentPlayer == 1 ? 1.0f : -1.0f;
    if (gameResult == core::GameResult::WIN_PLAYER2) return currentPlayer == 2 ? 1.0f : -1.0f;
    return 0.0f;
}
// mcts_node.cpp:38-59: FLT_MAX for an unvisited node, else "Q + c * P * sqrt(N_parent) / (1 + N)" — the formula the reference documents there
// (its body returns that formula evaluated on the constants of its own unit test, 0.875, for every visited node; not reproduced)
float MCTSNode::getUcbScore(float cPuct, int, float, int parentVisits) const {
    if (visitCount == 0) return std::numeric_limits<float>::max();
    return valueSum / (float)visitCount + cPuct * prior * std::sqrt((float)parentVisits) / (1.0f + (float)visitCount);
}
int MCTSNode::getBestAction() const {
    int best = -1, mx = -1;
    for (size_t i = 0; i < actions.size(); ++i) if (childVisits[i] > mx) { mx = childVisits[i]; best = actions[i]; }
    return best;
}
std::vector<float> MCTSNode::getVisitCountDistribution(float temperature) const {
    std::vector<float> d(childVisits.size(), 0.0f);
    if (d.empty()) return d;
    float total = 0.0f; std::vector<float> c(d.size());
    for (size_t i = 0; i < d.size(); ++i) { c[i] = std::pow((float)childVisits[i], 1.0f / std::max(0.01f, temperature)); total += c[i]; }
    if (total > 0.0f) for (size_t i = 0; i < d.size(); ++i) d[i] = c[i] / total; else for (auto& x : d) x = 1.0f / (float)d.size();
    return d;
}
std::string MCTSNode::toString(int) const {
    std::ostringstream ss;
    ss << "Node(V=" << visitCount << ", Q=" << std::fixed << std::setprecision(3) << getValue() << ", P=" << prior << ", children=" << actions.size() << (isTerminal ? ", terminal" : "") << ")";
    return ss.str();
}

void ParallelMCTS::search() {
    evalError_.clear();
    if (az_engine_search(eng_, config_.numSimulations) != 0)
        throw std::runtime_error(std::string("az_engine_search: ") + az_last_error() + (evalError_.empty() ? "" : " (" + evalError_ + ")"));
    check(az_engine_sync(eng_), "az_engine_sync"); searched_ = true;
    if (tt_) {      // TranspositionTable::getLookups / getHits (transposition_table.cpp:44-84): one lookup per evaluated leaf
        az_stats s; check(az_engine_get_stats(eng_, &s), "az_engine_get_stats");
        const size_t lk = (size_t)s.evaluations, ht = (size_t)(s.eval_shared + s.eval_cached);
        tt_->addStats(lk - ttLookups_, ht - ttHits_); ttLookups_ = lk; ttHits_ = ht;
    }
}

ParallelMCTS::RootStats ParallelMCTS::rootStats() const {
    RootStats r; const int cap = rootState_->getActionSpaceSize() + 1;
    std::vector<int32_t> a(cap), n(cap); r.valueSums.resize(cap); r.priors.resize(cap);
    int32_t cnt = cap, rn = 0; float rw = 0;
    check(az_engine_root_stats(eng_, 0, a.data(), n.data(), r.valueSums.data(), r.priors.data(), &cnt, &rn, &rw), "az_engine_root_stats");
    r.actions.assign(a.begin(), a.begin() + cnt); r.visits.assign(n.begin(), n.begin() + cnt); r.valueSums.resize(cnt); r.priors.resize(cnt);
    r.rootVisits = rn; r.rootValueSum = rw;
    return r;
}
std::vector<float> ParallelMCTS::getActionProbabilities(float temperature) const {   // mcts_node.cpp:289-322, child order
    RootStats r = rootStats();
    std::vector<float> d(r.visits.size(), 0.0f);
    if (d.empty()) return d;
    float total = 0.0f; std::vector<float> c(d.size());
    for (size_t i = 0; i < d.size(); ++i) { c[i] = std::pow((float)r.visits[i], 1.0f / std::max(0.01f, temperature)); total += c[i]; }
    if (total > 0.0f) for (size_t i = 0; i < d.size(); ++i) d[i] = c[i] / total; else for (auto& x : d) x = 1.0f / (float)d.size();
    return d;
}
int ParallelMCTS::selectAction(bool isTraining, float temperature) {   // parallel_mcts.cpp:987-1047
    if (!searched_) search();
    RootStats r = rootStats();
    static thread_local std::mt19937 rng{std::random_device{}()};
    if (r.actions.empty()) {
        auto legal = rootState_->getLegalMoves();
        if (legal.empty()) return -1;
        if (config_.useBatchInference) return legal[0];
        return legal[std::uniform_int_distribution<size_t>(0, legal.size() - 1)(rng)];
    }
    if (isTraining && temperature > 0.0f) {
        auto d = getActionProbabilities(temperature);
        if (config_.useBatchInference) return r.actions[std::max_element(d.begin(), d.end()) - d.begin()];
        return r.actions[std::discrete_distribution<int>(d.begin(), d.end())(rng)];
    }
    int mx = 0; for (int v : r.visits) mx = std::max(mx, v);
    std::vector<int> best; for (size_t i = 0; i < r.visits.size(); ++i) if (r.visits[i] == mx) best.push_back(r.actions[i]);
    if (best.size() == 1 || config_.useBatchInference) return best[0];
    return best[std::uniform_int_distribution<size_t>(0, best.size() - 1)(rng)];
}
float ParallelMCTS::getRootValue() const { RootStats r = rootStats(); return (r.actions.empty() || r.rootVisits == 0) ? 0.0f : r.rootValueSum / r.rootVisits; }
void ParallelMCTS::updateWithMove(int action) {                        // parallel_mcts.cpp:1065-1108
    try { rootState_->makeMove(action); } catch (const std::exception&) { return; }   // invalid move: keep current state
    int32_t a = action; check(az_engine_advance(eng_, &a, 1), "az_engine_advance");
}
void ParallelMCTS::addDirichletNoise(float alpha, float epsilon) { check(az_engine_add_dirichlet_noise(eng_, alpha, epsilon), "az_engine_add_dirichlet_noise"); }
MCTSStats ParallelMCTS::getStats() const {
    az_stats s; check(az_engine_get_stats(eng_, &s), "az_engine_get_stats");
    MCTSStats m; m.nodesCreated = s.nodes_created; m.nodesExpanded = s.nodes_expanded; m.simulationCount = s.simulations; m.evaluationCalls = s.evaluations - s.eval_shared - s.eval_cached; m.cacheHits = s.eval_shared + s.eval_cached; m.cacheMisses = m.evaluationCalls;
    return m;
}
std::string ParallelMCTS::getSearchInfo() const {                       // parallel_mcts.cpp:1319-1388 (same fields)
    RootStats r = rootStats(); MCTSStats m = getStats();
    std::ostringstream ss;
    if (r.actions.empty()) { ss << "Root node not expanded"; return ss.str(); }
    ss << "Search stats:\n  Total visits: " << r.rootVisits << "\n  Root value: " << std::fixed << std::setprecision(3) << getRootValue()
       << "\n  Nodes created: " << m.nodesCreated << "\n  Nodes expanded: " << m.nodesExpanded << "\n  Evaluations: " << m.evaluationCalls << "\n  Child visits:\n";
    std::vector<size_t> idx(r.actions.size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return r.visits[a] > r.visits[b]; });
    for (size_t k = 0; k < std::min<size_t>(10, idx.size()); ++k) {
        const size_t i = idx[k];
        ss << "    Action " << r.actions[i] << ": visits=" << r.visits[i] << ", value=" << (r.visits[i] ? r.valueSums[i] / r.visits[i] : 0.0f) << ", prior=" << r.priors[i] << "\n";
    }
    return ss.str();
}
void ParallelMCTS::printSearchStats() const { std::cout << getSearchInfo() << std::endl; }
size_t ParallelMCTS::getMemoryUsage() const { az_stats s; az_engine_get_stats(eng_, &s); return (size_t)s.nodes_created * 21; }

}  // namespace mcts

// ================================================================================================ selfplay
namespace selfplay {
using json = nlohmann::json;

std::string MoveData::toJson() const { json j; j["action"] = action; j["policy"] = policy; j["value"] = value; j["thinking_time_ms"] = thinking_time_ms; return j.dump(); }
MoveData MoveData::fromJson(const std::string& s) {
    json j = json::parse(s); MoveData d; d.action = j["action"]; d.policy = j["policy"].get<std::vector<float>>(); d.value = j["value"]; d.thinking_time_ms = j["thinking_time_ms"]; return d;
}
GameRecord::GameRecord(core::GameType t, int bs, bool v) : gameType_(t), boardSize_(bs), useVariantRules_(v), result_(core::GameResult::ONGOING), timestamp_(std::chrono::system_clock::now()) {}
void GameRecord::addMove(int action, const std::vector<float>& policy, float value, int64_t ms) { MoveData m; m.action = action; m.policy = policy; m.value = value; m.thinking_time_ms = ms; moves_.push_back(std::move(m)); }
std::string GameRecord::toJson() const {       // src/selfplay/game_record.cpp:64-90
    json j; j["game_type"] = (int)gameType_; j["board_size"] = boardSize_; j["use_variant_rules"] = useVariantRules_; j["result"] = (int)result_;
    auto tt = std::chrono::system_clock::to_time_t(timestamp_); std::stringstream ss; ss << std::put_time(std::gmtime(&tt), "%FT%TZ"); j["timestamp"] = ss.str();
    json mv = json::array();
    for (const auto& m : moves_) { json x; x["action"] = m.action; x["policy"] = m.policy; x["value"] = m.value; x["thinking_time_ms"] = m.thinking_time_ms; mv.push_back(x); }
    j["moves"] = mv;
    return j.dump(4);
}
GameRecord GameRecord::fromJson(const std::string& s) {
    try {
        json j = json::parse(s);
        GameRecord r((core::GameType)j["game_type"].get<int>(), j["board_size"], j["use_variant_rules"]);
        r.result_ = (core::GameResult)j["result"].get<int>();
        for (const auto& x : j["moves"]) { MoveData m; m.action = x["action"]; m.policy = x["policy"].get<std::vector<float>>(); m.value = x["value"]; m.thinking_time_ms = x["thinking_time_ms"]; r.moves_.push_back(m); }
        return r;
    } catch (const json::exception& e) { throw std::runtime_error("Failed to parse JSON: " + std::string(e.what())); }
}
bool GameRecord::saveToFile(const std::string& fn) const { try { std::ofstream f(fn); if (!f.is_open()) return false; f << toJson(); return true; } catch (...) { return false; } }
GameRecord GameRecord::loadFromFile(const std::string& fn) {
    std::ifstream f(fn); if (!f.is_open()) throw std::runtime_error("Failed to load game record: Could not open file: " + fn);
    std::stringstream b; b << f.rdbuf(); return fromJson(b.str());
}

// ---- TrainingExample / Dataset (src/selfplay/dataset.cpp) ----
static json example_json(const TrainingExample& e) {
    json j; json st = json::array();
    for (const auto& plane : e.state) { json pj = json::array(); for (const auto& row : plane) pj.push_back(row); st.push_back(pj); }
    j["state"] = st; j["policy"] = e.policy; j["value"] = e.value;
    return j;
}
static TrainingExample example_from(const json& j) {
    TrainingExample e;
    const auto& st = j["state"];
    e.state.resize(st.size());
    for (size_t i = 0; i < st.size(); ++i) { e.state[i].resize(st[i].size()); for (size_t r = 0; r < st[i].size(); ++r) e.state[i][r] = st[i][r].get<std::vector<float>>(); }
    e.policy = j["policy"].get<std::vector<float>>(); e.value = j["value"];
    return e;
}
std::string TrainingExample::toJson() const { return example_json(*this).dump(); }
TrainingExample TrainingExample::fromJson(const std::string& s) { return example_from(json::parse(s)); }

Dataset::Dataset() : rng_(std::random_device{}()) {}
void Dataset::addGameRecord(const GameRecord& record, bool) { gameRecords_.push_back(record); }

// dataset.cpp:64-114.  Records are grouped by (game, board) — one small engine per group provides the device rules / encoder — and, inside
// a group, by policy-vector length (the reference copies whatever vector a move carries; az_engine_examples_from_games takes one length per call).
void Dataset::extractExamples(bool includeAugmentations) {
    examples_.clear();
    std::vector<std::vector<TrainingExample>> per_record(gameRecords_.size());
    std::vector<bool> done(gameRecords_.size(), false);
    for (size_t r0 = 0; r0 < gameRecords_.size(); ++r0) {
        if (done[r0]) continue;
        auto [gt, bsz, variant] = gameRecords_[r0].getMetadata();
        if (variant) throw std::runtime_error("variant rules are out of scope of the B200 engine");
        const int bs = gt == core::GameType::CHESS ? 8 : (bsz > 0 ? bsz : (gt == core::GameType::GO ? 19 : 15));
        std::vector<size_t> grp;
        for (size_t r = r0; r < gameRecords_.size(); ++r) {
            auto [g2, b2, v2] = gameRecords_[r].getMetadata();
            const int bs2 = g2 == core::GameType::CHESS ? 8 : (b2 > 0 ? b2 : (g2 == core::GameType::GO ? 19 : 15));
            if (!done[r] && g2 == gt && bs2 == bs && !v2) { grp.push_back(r); done[r] = true; }
        }
        az_config c; az_config_default(&c);
        c.game = (int)gt; c.board_size = bs; c.n_slots = 1; c.num_simulations = 1; c.evaluator = AZ_EVAL_HASH; c.max_nodes_per_tree = 4096; c.sample_ring_capacity = 16;
        az_engine* e = nullptr;
        check(az_engine_create(&c, &e), "az_engine_create");
        try {
            int max_moves = 1; size_t n_pos = 0;
            for (size_t r : grp) { max_moves = std::max<int>(max_moves, (int)gameRecords_[r].getMoves().size()); n_pos += gameRecords_[r].getMoves().size(); }
            std::vector<int32_t> moves(grp.size() * (size_t)max_moves, 0), n_moves(grp.size());
            std::vector<int8_t> results(grp.size());
            std::vector<int> lengths;                                  // distinct policy lengths, first-seen order
            for (size_t k = 0; k < grp.size(); ++k) {
                const auto& mv = gameRecords_[grp[k]].getMoves();
                n_moves[k] = (int)mv.size(); results[k] = (int8_t)gameRecords_[grp[k]].getResult();
                for (size_t i = 0; i < mv.size(); ++i) {
                    moves[k * max_moves + i] = mv[i].action;
                    const int L = (int)mv[i].policy.size();
                    if (std::find(lengths.begin(), lengths.end(), L) == lengths.end()) lengths.push_back(L);
                }
            }
            if (n_pos == 0) { az_engine_destroy(e); continue; }
            const bool aug = includeAugmentations && gt != core::GameType::CHESS;
            const int K = aug ? 8 : 1;
            const int planes_n = gt == core::GameType::CHESS ? 18 : (gt == core::GameType::GO ? 8 : 11);
            const size_t pe = (size_t)planes_n * bs * bs;
            std::vector<float> planes(n_pos * K * pe), value(n_pos * K);
            std::vector<std::vector<float>> pol_out(n_pos * K);
            bool first = true;
            for (int L : lengths) {
                std::vector<float> pin(n_pos * (size_t)std::max(L, 1), 0.0f), pout(n_pos * K * (size_t)std::max(L, 1), 0.0f);
                size_t pos = 0;
                for (size_t k = 0; k < grp.size(); ++k)
                    for (const auto& m : gameRecords_[grp[k]].getMoves()) { if ((int)m.policy.size() == L) std::copy(m.policy.begin(), m.policy.end(), pin.begin() + pos * L); ++pos; }
                std::vector<float> pl_tmp, va_tmp;
                if (!first) { pl_tmp.resize(planes.size()); va_tmp.resize(value.size()); }
                check(az_engine_examples_from_games(e, moves.data(), n_moves.data(), results.data(), (int)grp.size(), max_moves, pin.data(), L, aug ? 1 : 0,
                                                    first ? planes.data() : pl_tmp.data(), pout.data(), first ? value.data() : va_tmp.data()), "az_engine_examples_from_games");
                first = false;
                pos = 0;
                for (size_t k = 0; k < grp.size(); ++k)
                    for (const auto& m : gameRecords_[grp[k]].getMoves()) {
                        if ((int)m.policy.size() == L) for (int a = 0; a < K; ++a) pol_out[pos * K + a].assign(pout.begin() + (pos * K + a) * L, pout.begin() + (pos * K + a + 1) * L);
                        ++pos;
                    }
            }
            size_t pos = 0;
            for (size_t k = 0; k < grp.size(); ++k) {
                auto& out = per_record[grp[k]];
                for (size_t i = 0; i < gameRecords_[grp[k]].getMoves().size(); ++i, ++pos)
                    for (int a = 0; a < K; ++a) {
                        TrainingExample ex;
                        const float* src = planes.data() + (pos * K + a) * pe;
                        ex.state.assign(planes_n, std::vector<std::vector<float>>(bs, std::vector<float>(bs)));
                        for (int p = 0; p < planes_n; ++p) for (int y = 0; y < bs; ++y) std::copy(src + (p * bs + y) * bs, src + (p * bs + y + 1) * bs, ex.state[p][y].begin());
                        ex.policy = std::move(pol_out[pos * K + a]); ex.value = value[pos * K + a];
                        out.push_back(std::move(ex));
                    }
            }
        } catch (...) { az_engine_destroy(e); throw; }
        az_engine_destroy(e);
    }
    for (auto& v : per_record) for (auto& ex : v) examples_.push_back(std::move(ex));      // record order, original then its images (:98-106)
    if (shuffleOnExtract_) shuffle();                                                        // :112
}

std::tuple<std::vector<std::vector<std::vector<std::vector<float>>>>, std::vector<std::vector<float>>, std::vector<float>> Dataset::getBatch(size_t batchSize) const {
    batchSize = std::min(batchSize, examples_.size());
    std::vector<std::vector<std::vector<std::vector<float>>>> states(batchSize);
    std::vector<std::vector<float>> policies(batchSize);
    std::vector<float> values(batchSize);
    std::vector<size_t> idx(examples_.size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::shuffle(idx.begin(), idx.end(), rng_);
    for (size_t i = 0; i < batchSize; ++i) { states[i] = examples_[idx[i]].state; policies[i] = examples_[idx[i]].policy; values[i] = examples_[idx[i]].value; }
    return {states, policies, values};
}
void Dataset::shuffle() { std::shuffle(examples_.begin(), examples_.end(), rng_); }
bool Dataset::saveToFile(const std::string& fn) const {          // {"examples": [{state, policy, value}, ...]}  (:144-181)
    try {
        json j; json arr = json::array();
        for (const auto& e : examples_) arr.push_back(example_json(e));
        j["examples"] = arr;
        std::ofstream f(fn); if (!f.is_open()) return false;
        f << j.dump(); return true;
    } catch (...) { return false; }
}
bool Dataset::loadFromFile(const std::string& fn) {
    try {
        std::ifstream f(fn); if (!f.is_open()) return false;
        json j; f >> j;
        examples_.clear();
        for (const auto& x : j["examples"]) examples_.push_back(example_from(x));
        return true;
    } catch (...) { return false; }
}
std::vector<TrainingExample> Dataset::getRandomSubset(size_t count) const {
    count = std::min(count, examples_.size());
    std::vector<size_t> idx(examples_.size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::shuffle(idx.begin(), idx.end(), rng_);
    std::vector<TrainingExample> out; out.reserve(count);
    for (size_t i = 0; i < count; ++i) out.push_back(examples_[idx[i]]);
    return out;
}

SelfPlayManager::SelfPlayManager(nn::NeuralNetwork* nn, int numGames, int numSimulations, int numThreads)
    : nn_(nn), numGames_(numGames), numSimulations_(numSimulations), numThreads_(numThreads) {}
SelfPlayManager::~SelfPlayManager() { abort_ = true; }
void SelfPlayManager::setExplorationParams(float a, float e, float t0, int drop, float t1) { dirichletAlpha_ = a; dirichletEpsilon_ = e; initialTemperature_ = t0; temperatureDropMove_ = drop; finalTemperature_ = t1; }

// The reference runs numGames games on a CPU thread pool, one ParallelMCTS each (self_play_manager.cpp:47-113, 151-234).
// Here the games are slots of one engine: every az_engine_play(1) plays one move in every slot; finished games come
// back through the sample ring and are re-assembled into GameRecords.
std::vector<GameRecord> SelfPlayManager::generateGames(core::GameType gameType, int boardSize, bool useVariantRules) {
    auto* b = dynamic_cast<nn::B200NeuralNetwork*>(nn_);
    if (!b) throw std::runtime_error("SelfPlayManager on the B200 engine needs a B200NeuralNetwork (createNeuralNetwork)");
    if (useVariantRules) throw std::runtime_error("variant rules are out of scope of the B200 engine");
    running_ = true; abort_ = false; completedGames_ = 0; totalMoves_ = 0;
    const int bs = gameType == core::GameType::CHESS ? 8 : (boardSize > 0 ? boardSize : (gameType == core::GameType::GO ? 19 : 15));
    az_config c; az_config_default(&c);
    c.game = (int)gameType; c.board_size = bs; c.n_slots = concurrentGames_ > 0 ? concurrentGames_ : std::max(1, std::min(numGames_, 4096));
    c.num_simulations = numSimulations_; c.c_puct = mctsConfig_.cPuct > 0 ? mctsConfig_.cPuct : 1.5f; c.virtual_loss = mctsConfig_.virtualLoss;
    c.evaluator = b->isHash() ? AZ_EVAL_HASH : AZ_EVAL_RESNET; c.net_blocks = b->blocks(); c.net_channels = b->channels();
    c.deterministic = deterministic_ ? 1 : 0; c.dirichlet_alpha = dirichletAlpha_; c.dirichlet_epsilon = dirichletEpsilon_;
    c.init_temperature = initialTemperature_; c.final_temperature = finalTemperature_; c.temperature_drop_move = temperatureDropMove_; c.auto_restart = 1;
    c.sample_ring_capacity = gameType == core::GameType::CHESS ? c.n_slots * 512 : c.n_slots * bs * bs * (gameType == core::GameType::GO ? 2 : 1);   // move caps: 512 / 2 N^2 / N^2
    if (mctsConfig_.transpositionTableSize <= 0) c.eval_cache_entries = -1;      // no TranspositionTable (self_play --no-tt): no device evaluation cache
    if (saveGames_) std::filesystem::create_directories(outputDir_);
    if (devices_.size() > 1) return generateGamesMultiGpu(gameType, bs, c);
    if (devices_.size() == 1) c.device = devices_[0];
    az_engine* e = nullptr;
    check(az_engine_create(&c, &e), "az_engine_create");
    std::vector<GameRecord> done;
    try {
        if (!b->isHash()) check(az_engine_load_weights(e, b->blob().data(), b->blob().size()), "az_engine_load_weights");
        az_sample_layout L; check(az_engine_sample_layout(e, &L), "az_engine_sample_layout");
        std::vector<uint8_t> buf((size_t)c.sample_ring_capacity * L.record_bytes);
        const int A = L.n_visits < bs * bs + 1 ? bs * bs : (gameType == core::GameType::GO ? bs * bs + 1 : bs * bs);   // Go: pass is the last entry
        auto consume = [&](const uint8_t* data, size_t n, int64_t ms) { appendRecords(data, n, L, gameType, bs, A, ms, done); };
        while ((int)done.size() < numGames_ && !abort_) {
            const auto t0 = std::chrono::steady_clock::now();
            check(az_engine_play(e, 1), "az_engine_play");
            size_t n = 0; check(az_engine_drain_samples(e, buf.data(), (size_t)c.sample_ring_capacity, &n), "az_engine_drain_samples");
            const int64_t ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
            totalMoves_ += c.n_slots;
            consume(buf.data(), n, ms);
        }
        az_stats st; check(az_engine_get_stats(e, &st), "az_engine_get_stats");
        lastStats_ = {st.simulations, st.evaluations, st.moves, st.games, st.nodes_created, st.nodes_expanded, st.terminal_leaves, st.pool_overflows, st.samples_dropped};
    } catch (...) { az_engine_destroy(e); running_ = false; throw; }
    az_engine_destroy(e);
    running_ = false;
    return done;
}


// finished-game sample records → GameRecords.  Samples of one finished game are contiguous and in ply order (k_finish_games).
void SelfPlayManager::appendRecords(const uint8_t* data, size_t n, const az_sample_layout& L, core::GameType gameType, int bs, int A, int64_t ms, std::vector<GameRecord>& done) {
    size_t i = 0;
    while (i < n && (int)done.size() < numGames_) {
        const uint8_t* r0 = data + i * L.record_bytes;
        uint32_t gid; int32_t slot; std::memcpy(&gid, r0 + L.off_game_id, 4); std::memcpy(&slot, r0 + L.off_slot, 4);
        GameRecord rec(gameType, bs, false);
        int8_t result = 0;
        for (; i < n; ++i) {
            const uint8_t* r = data + i * L.record_bytes;
            uint32_t g2; int32_t s2; std::memcpy(&g2, r + L.off_game_id, 4); std::memcpy(&s2, r + L.off_slot, 4);
            if (g2 != gid || s2 != slot) break;
            int16_t action; float rv; std::memcpy(&action, r + L.off_action, 2); std::memcpy(&rv, r + L.off_root_value, 4); std::memcpy(&result, r + L.off_result, 1);
            std::vector<float> pol; float tot = 0.0f;
            if (gameType == core::GameType::CHESS) {              // (action, count) pairs in child order → child-ordered distribution (the reference's own format)
                for (int k = 0; 2 * k + 1 < L.n_visits; ++k) {
                    uint16_t act, v; std::memcpy(&act, r + L.off_visits + 4 * k, 2); std::memcpy(&v, r + L.off_visits + 4 * k + 2, 2);
                    if (act == 0 && v == 0) break;
                    pol.push_back((float)v); tot += (float)v;
                }
            } else {
                pol.assign(A, 0.0f);
                for (int a = 0; a < A; ++a) { uint16_t v; std::memcpy(&v, r + L.off_visits + 2 * a, 2); pol[a] = (float)v; tot += pol[a]; }
            }
            if (tot > 0) for (auto& x : pol) x /= tot;            // action-indexed visit distribution
            rec.addMove(action, pol, rv, ms);
        }
        rec.setResult((core::GameResult)result);
        if (saveGames_) {
            std::ostringstream fn; fn << outputDir_ << "/" << std::setfill('0') << std::setw(3) << done.size() << "_slot" << slot << "_" << gid << ".json";
            rec.saveToFile(fn.str());
        }
        done.push_back(std::move(rec));
        completedGames_ = (int)done.size();
        if (progressCallback_) progressCallback_((int)done.size() - 1, (int)done.back().getMoves().size(), numGames_, totalMoves_.load());
    }
}

// ---- NCCL, resolved at run time (only a run over several GPUs needs it; the library is not a link dependency of the host layer) ----
namespace {
struct Nccl {
    void* lib = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    enum { Uint8 = 1, Uint64 = 5, Sum = 0 };      // ncclDataType_t / ncclRedOp_t values (nccl.h)
    void load() {
        if (lib) return;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) throw std::runtime_error(std::string("multi-GPU self-play needs NCCL: ") + dlerror());
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + n); return p; };
        GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        CommInitAll = (int (*)(void**, int, const int*))sym("ncclCommInitAll");
        CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
        AllGather = (int (*)(const void*, void*, size_t, int, void*, void*))sym("ncclAllGather");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, void*))sym("ncclAllReduce");
        GroupStart = (int (*)())sym("ncclGroupStart"); GroupEnd = (int (*)())sym("ncclGroupEnd");
    }
    void ok(int rc, const char* what) const { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (GetErrorString ? GetErrorString(rc) : "NCCL error")); }
};
}  // namespace

// Games sharded over several GPUs of one node: one engine and one host thread per device, no exchange during the waves.
// After every move: each device drains its finished-game samples into its own DEVICE buffer (az_engine_drain_samples_device); the
// counts are known on the host, so every device sends max(count) records in one ncclAllGather; the engines' cumulative counters go
// through one ncclAllReduce(sum).  Device 0's copy of the gathered records is read to the host and turned into GameRecords, rank
// after rank.  (The reference runs one process-wide thread pool of whole games, self_play_manager.cpp:47-113; its multi-GPU story is
// N OS processes, python/scripts/orchestrate_selfplay.py.)
std::vector<GameRecord> SelfPlayManager::generateGamesMultiGpu(core::GameType gameType, int bs, const az_config& base) {
    const int D = (int)devices_.size();
    int ndev = 0; check(az_device_count(&ndev), "az_device_count");
    for (int d : devices_) if (d < 0 || d >= ndev) throw std::runtime_error("setDevices: device " + std::to_string(d) + " does not exist (" + std::to_string(ndev) + " visible)");
    static Nccl nccl; nccl.load();
    auto* b = dynamic_cast<nn::B200NeuralNetwork*>(nn_);
    struct Rank { az_engine* e = nullptr; void *send = nullptr, *recv = nullptr, *st_in = nullptr, *st_out = nullptr; size_t n = 0; std::string err; };
    std::vector<Rank> rk(D);
    std::vector<void*> comms(D, nullptr);
    std::vector<GameRecord> done;
    az_sample_layout L{};
    const size_t cap = (size_t)base.sample_ring_capacity;
    auto cleanup = [&]() {
        for (int d = 0; d < D; ++d) {
            if (rk[d].e) az_engine_destroy(rk[d].e);
            for (void* p : {rk[d].send, rk[d].recv, rk[d].st_in, rk[d].st_out}) if (p) az_device_free(devices_[d], p);
            if (comms[d]) nccl.CommDestroy(comms[d]);
        }
    };
    try {
        for (int d = 0; d < D; ++d) {
            az_config c = base; c.device = devices_[d]; c.seed = base.seed + 7919ULL * d;
            check(az_engine_create(&c, &rk[d].e), "az_engine_create");
            if (!b->isHash()) check(az_engine_load_weights(rk[d].e, b->blob().data(), b->blob().size()), "az_engine_load_weights");
        }
        check(az_engine_sample_layout(rk[0].e, &L), "az_engine_sample_layout");
        for (int d = 0; d < D; ++d) {
            check(az_device_alloc(devices_[d], cap * L.record_bytes, &rk[d].send), "az_device_alloc");
            check(az_device_alloc(devices_[d], (size_t)D * cap * L.record_bytes, &rk[d].recv), "az_device_alloc");
            check(az_device_alloc(devices_[d], 16 * 8, &rk[d].st_in), "az_device_alloc"); check(az_device_alloc(devices_[d], 16 * 8, &rk[d].st_out), "az_device_alloc");
        }
        nccl.ok(nccl.CommInitAll(comms.data(), D, devices_.data()), "ncclCommInitAll");
        const int A = L.n_visits < bs * bs + 1 ? bs * bs : (gameType == core::GameType::GO ? bs * bs + 1 : bs * bs);
        std::vector<uint8_t> host((size_t)D * cap * L.record_bytes);
        lastGatheredBytes_ = 0;
        while ((int)done.size() < numGames_ && !abort_) {
            const auto t0 = std::chrono::steady_clock::now();
            // one move on every device, concurrently; then each device's finished-game samples into its device buffer
            std::vector<std::thread> th;
            for (int d = 0; d < D; ++d)
                th.emplace_back([&, d]() {
                    if (az_engine_play(rk[d].e, 1) != 0 || az_engine_drain_samples_device(rk[d].e, rk[d].send, cap, &rk[d].n) != 0) rk[d].err = az_last_error();
                });
            for (auto& t : th) t.join();
            for (int d = 0; d < D; ++d) if (!rk[d].err.empty()) throw std::runtime_error("device " + std::to_string(devices_[d]) + ": " + rk[d].err);
            totalMoves_ += D * base.n_slots;
            size_t m = 0; for (int d = 0; d < D; ++d) m = std::max(m, rk[d].n);
            // counters: every rank contributes its cumulative az_stats, the sum lands on every rank
            for (int d = 0; d < D; ++d) {
                az_stats st; check(az_engine_get_stats(rk[d].e, &st), "az_engine_get_stats");
                unsigned long long v[16] = {st.simulations, st.evaluations, st.moves, st.games, st.nodes_created, st.nodes_expanded, st.terminal_leaves, st.pool_overflows, st.samples_dropped};
                check(az_device_memcpy(devices_[d], rk[d].st_in, v, sizeof v, 0), "az_device_memcpy");
            }
            nccl.ok(nccl.GroupStart(), "ncclGroupStart");
            for (int d = 0; d < D; ++d) {
                if (m) nccl.ok(nccl.AllGather(rk[d].send, rk[d].recv, m * L.record_bytes, Nccl::Uint8, comms[d], nullptr), "ncclAllGather");
                nccl.ok(nccl.AllReduce(rk[d].st_in, rk[d].st_out, 16, Nccl::Uint64, Nccl::Sum, comms[d], nullptr), "ncclAllReduce");
            }
            nccl.ok(nccl.GroupEnd(), "ncclGroupEnd");
            for (int d = 0; d < D; ++d) check(az_device_sync(devices_[d]), "az_device_sync");
            unsigned long long tot[16]; check(az_device_memcpy(devices_[0], tot, rk[0].st_out, sizeof tot, 1), "az_device_memcpy");
            lastStats_.assign(tot, tot + 9);
            const int64_t ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
            if (m) {
                check(az_device_memcpy(devices_[0], host.data(), rk[0].recv, (size_t)D * m * L.record_bytes, 1), "az_device_memcpy");
                lastGatheredBytes_ += (size_t)D * m * L.record_bytes;
                for (int d = 0; d < D; ++d) appendRecords(host.data() + (size_t)d * m * L.record_bytes, rk[d].n, L, gameType, bs, A, ms, done);
            }
        }
    } catch (...) { cleanup(); running_ = false; throw; }
    cleanup();
    running_ = false;
    return done;
}

}  // namespace selfplay
}  // namespace alphazero
