// oracle/az_oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product path (alphazero-multi-game_b200/csrc) never links or calls it.
//
// What is restated (reference file:line in each function):
//   * GomokuState   — src/games/gomoku/gomoku_state.cpp + gomoku_rules.cpp:39-115 (standard rules)
//   * GoState       — src/games/go/go_state.cpp + go_rules.cpp (capture / ko / superko / suicide / area score)
//   * serial search — src/mcts/parallel_mcts.cpp (numThreads=1, non-batched path) + mcts_node.cpp
//   * HashEvaluator — SURVEY.md Appendix C (stateless, exactly-rounded fp32 ops)
// Parity pin: checked in tests/test_oracle_vs_ref.py against oracle/_ref/libaz_ref.so (the patched
// reference itself, built by oracle/build_ref.sh) and against tests/golden/*.json generated from it.
//
// Build: g++ -std=c++17 -O2 -fPIC -ffp-contract=off -shared oracle/az_oracle.cpp -o oracle/libaz_oracle.so
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <queue>
#include <unordered_set>
#include <unordered_map>
#include <algorithm>
#include <memory>
#include <cstdlib>
#include <utility>

namespace orc {

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

enum { ONGOING = 0, DRAW = 1, WIN_P1 = 2, WIN_P2 = 3 };   // core::GameResult, igamestate.h:25-30

struct State {
    virtual ~State() {}
    virtual State* clone() const = 0;
    virtual std::vector<int> legal() const = 0;
    virtual bool make_move(int a) = 0;          // false = reference would throw
    virtual bool terminal() const = 0;
    virtual int result() const = 0;
    virtual int player() const = 0;
    virtual int action_space() const = 0;
    virtual int board_size() const = 0;
    virtual uint64_t key() const = 0;           // HashEvaluator key (Appendix C)
    // Granularity of the reference's getHash(), the TranspositionTable key (parallel_mcts.cpp:320-336, 851): Gomoku = stones + player
    // (gomoku_state.cpp:620-656), Go = stones + player + ko point (go_state.cpp:773-811) — what key() covers — but chess = the piece
    // PLACEMENT only (QUIRK C8, see Chess::tt_key).
    virtual uint64_t tt_key() const { return key(); }
    virtual int planes() const = 0;
    virtual void tensor(float* out) const = 0;  // [planes][N][N]
    virtual int game_type() const = 0;
};

// ------------------------------------------------------------------------------------------ Gomoku
struct Gomoku : State {
    int N, nw, cur;                       // gomoku_state.h:76-85; cur 1=BLACK first
    std::vector<uint64_t> bb[2];          // bit a = x*N + y
    std::vector<int> hist;
    // legal-move cache: the reference keeps a std::unordered_set<int> whose *iteration order*
    // is the legal-move order (gomoku_state.h:124, gomoku_state.cpp:531-578).  Same container,
    // same libstdc++, same fill discipline ⇒ same order.
    mutable std::unordered_set<int> cache;
    mutable bool dirty = true;

    explicit Gomoku(int n) : N(n), nw((n * n + 63) / 64), cur(1) { bb[0].assign(nw, 0); bb[1].assign(nw, 0); }
    State* clone() const override { return new Gomoku(*this); }
    bool bit(int p, int a) const { return (bb[p][a >> 6] >> (a & 63)) & 1; }
    bool occupied(int a) const { return ((bb[0][a >> 6] | bb[1][a >> 6]) >> (a & 63)) & 1; }
    int stones() const { int t = 0; for (int p = 0; p < 2; ++p) for (uint64_t w : bb[p]) t += __builtin_popcountll(w); return t; }

    void refill() const {                 // refresh_valid_moves_cache, gomoku_state.cpp:531-566
        cache.clear();
        for (int a = 0; a < N * N; ++a) if (!occupied(a)) cache.insert(a);
        dirty = false;
    }
    std::vector<int> legal() const override {            // get_valid_moves :518-529
        if (dirty) refill();
        return std::vector<int>(cache.begin(), cache.end());
    }
    int run(int x, int y, int dx, int dy, int p) const { // count_direction, gomoku_rules.cpp:97-115
        int c = 0;
        while (x >= 0 && x < N && y >= 0 && y < N && bit(p, x * N + y)) { ++c; x += dx; y += dy; }
        return c;
    }
    bool five(int player) const {         // is_five_in_a_row(-1, p) + check_line_for_five :39-95
        int p = player - 1;
        static const int D[4][2] = {{0, 1}, {1, 0}, {1, 1}, {1, -1}};
        for (int a = 0; a < N * N; ++a) {
            if (!bit(p, a)) continue;
            int x = a / N, y = a % N;
            for (auto& d : D) {
                int len = run(x, y, d[0], d[1], p) + run(x, y, -d[0], -d[1], p) - 1;
                if (player == 1 ? (len == 5) : (len >= 5)) return true;   // QUIRK G3: black exactly 5
            }
        }
        return false;
    }
    int winner() const { if (five(1)) return 1; if (five(2)) return 2; return 0; }   // :477-489
    bool stalemate() const {              // is_stalemate :506-521
        if (!dirty) return cache.empty();
        if (stones() >= N * N) return true;
        refill();
        return cache.empty();
    }
    bool terminal() const override { if (winner() != 0) return true; return stalemate(); }  // :491-504
    int result() const override {         // getGameResult :189-201
        int w = winner();
        if (w == 0) return stalemate() ? DRAW : ONGOING;
        return w == 1 ? WIN_P1 : WIN_P2;
    }
    bool make_move(int a) override {      // make_move :681-722
        if (a < 0 || a >= N * N || occupied(a)) return false;
        bb[cur - 1][a >> 6] |= 1ULL << (a & 63);
        cur = 3 - cur; hist.push_back(a); dirty = true;
        return true;
    }
    int player() const override { return cur; }
    int action_space() const override { return N * N; }
    int board_size() const override { return N; }
    int game_type() const override { return 0; }
    uint64_t key() const override {
        uint64_t h = 1469598103934665603ULL;
        for (int p = 0; p < 2; ++p) for (uint64_t w : bb[p]) h = mix64(h ^ w);
        return mix64(h ^ (uint64_t)cur);
    }
    int planes() const override { return 11; }
    std::vector<int> prev_moves(int player, int count) const {   // get_previous_moves :852-869 (QUIRK G5)
        std::vector<int> out(count, -1); int found = 0;
        for (int i = (int)hist.size() - 1; i >= 0 && found < count; --i) {
            int mp = ((hist.size() - i) % 2 == 1) ? cur : 3 - cur;
            if (mp == player) out[found++] = hist[i];
        }
        return out;
    }
    void tensor(float* t) const override {       // getEnhancedTensorRepresentation :207-258 + to_tensor :811-840
        int S = N * N; std::fill(t, t + 11 * S, 0.0f);
        int p = cur - 1;
        for (int a = 0; a < S; ++a) { if (bit(p, a)) t[a] = 1.0f; else if (bit(1 - p, a)) t[S + a] = 1.0f; }
        if (cur == 1) for (int a = 0; a < S; ++a) t[2 * S + a] = 1.0f;
        auto pb = prev_moves(1, 3), pw = prev_moves(2, 3);
        for (int i = 0; i < 3; ++i) { if (pb[i] != -1) t[(3 + i) * S + pb[i]] = 1.0f; if (pw[i] != -1) t[(6 + i) * S + pw[i]] = 1.0f; }
        for (int x = 0; x < N; ++x) for (int y = 0; y < N; ++y) {
            t[9 * S + x * N + y] = (float)x / (N - 1);
            t[10 * S + x * N + y] = (float)y / (N - 1);
        }
    }
};

// ---------------------------------------------------------------------------------------------- Go
struct Go : State {
    int N, cur = 1, ko = -1, passes = 0;      // go_state.h:204-219; pos = y*N + x
    float komi = 7.5f;
    std::vector<int> board; int captured[3] = {0, 0, 0};
    std::vector<uint64_t> pos_hist;            // position_history_
    std::vector<int> hist;

    explicit Go(int n) : N((n == 9 || n == 13 || n == 19) ? n : 19) { board.assign(N * N, 0); }
    State* clone() const override { return new Go(*this); }
    void adj(int pos, int* out, int& n) const {           // getAdjacentPositions go_state.cpp:773-790
        int x = pos % N, y = pos / N; n = 0;
        if (y > 0) out[n++] = pos - N; if (x < N - 1) out[n++] = pos + 1;
        if (y < N - 1) out[n++] = pos + N; if (x > 0) out[n++] = pos - 1;
    }
    // group of `pos` on board b; returns true if it has at least one liberty
    static bool group_has_liberty(const Go& g, const std::vector<int>& b, int pos, std::vector<int>* stones) {
        int col = b[pos]; std::vector<char> seen(g.N * g.N, 0); std::vector<int> st{pos}; seen[pos] = 1;
        bool lib = false; int nb[4], nn;
        for (size_t i = 0; i < st.size(); ++i) {
            g.adj(st[i], nb, nn);
            for (int k = 0; k < nn; ++k) {
                int q = nb[k];
                if (b[q] == 0) lib = true;
                else if (b[q] == col && !seen[q]) { seen[q] = 1; st.push_back(q); }
            }
        }
        if (stones) *stones = st;
        return lib;
    }
    bool suicidal(int a, int pl) const {        // isSuicidalMove go_rules.cpp:29-138
        std::vector<int> b = board; b[a] = pl; int opp = 3 - pl; int nb[4], nn; adj(a, nb, nn);
        for (int k = 0; k < nn; ++k) if (b[nb[k]] == opp && !group_has_liberty(*this, b, nb[k], nullptr)) return false;
        return !group_has_liberty(*this, b, a, nullptr);
    }
    bool valid(int a) const {                   // isValidMove go_state.cpp:814-835
        if (a < 0 || a >= N * N || board[a] != 0 || a == ko) return false;
        return !suicidal(a, cur);
    }
    // place + capture on a scratch board; returns captured positions, n groups, single-stone flag
    void place(std::vector<int>& b, int a, int pl, std::vector<int>& cap, int& ngroups, bool& single) const {
        b[a] = pl; int opp = 3 - pl; cap.clear(); ngroups = 0; single = false;
        std::vector<char> done(N * N, 0);
        for (int pos = 0; pos < N * N; ++pos) {  // findGroups(opponent) whole board, go_rules.cpp:144-181
            if (b[pos] != opp || done[pos]) continue;
            std::vector<int> st; bool lib = group_has_liberty(*this, b, pos, &st);
            for (int s : st) done[s] = 1;
            if (!lib) { ++ngroups; single = (st.size() == 1); for (int s : st) cap.push_back(s); }
        }
        for (int s : cap) b[s] = 0;
    }
    // Superko key (QUIRK Go3): Zobrist(board) ^ player(mover) ^ ko — here a collision-free-in-practice
    // 64-bit mix of the same three ingredients (the reference's keys are time-seeded, only equality matters).
    static uint64_t pos_key(const std::vector<int>& b, int mover, int kopt) {
        uint64_t h = 0;
        for (size_t i = 0; i < b.size(); ++i) if (b[i]) h ^= mix64(0x1234567ULL + i * 2 + (b[i] - 1));
        h ^= mix64(0xABCD0000ULL + mover);
        if (kopt >= 0) h ^= mix64(0xFEED0000ULL + kopt);
        return h;
    }
    bool superko(int a) const {                 // go_state.cpp:123-145, 837-845
        std::vector<int> b = board, cap; int ng; bool single; place(b, a, cur, cap, ng, single);
        uint64_t k = pos_key(b, cur, ko);       // candidate uses the OLD ko point
        for (uint64_t h : pos_hist) if (h == k) return true;
        return false;
    }
    std::vector<int> legal() const override {    // getLegalMoves :116-154 (QUIRK Go2: pass = -1 first)
        std::vector<int> m{-1};
        for (int pos = 0; pos < N * N; ++pos) if (valid(pos) && !superko(pos)) m.push_back(pos);
        return m;
    }
    bool is_legal(int a) const { if (a == -1) return true; return valid(a) && !superko(a); }
    bool make_move(int a) override {             // makeMove :190-261
        if (!is_legal(a)) return false;
        if (a == -1) { ++passes; ko = -1; hist.push_back(a); }
        else {
            passes = 0; std::vector<int> cap; int ng; bool single;
            place(board, a, cur, cap, ng, single);
            ko = (ng == 1 && single) ? cap[0] : -1;          // Go4
            captured[cur] += (int)cap.size();
            hist.push_back(a);
            pos_hist.push_back(pos_key(board, cur, ko));     // pushed before the player switch, NEW ko
        }
        cur = 3 - cur;
        return true;
    }
    bool terminal() const override { return passes >= 2; }   // :315-318
    void territory(std::vector<int>& terr) const {            // getTerritoryOwnership go_rules.cpp:196-243
        terr.assign(N * N, 0); std::vector<char> seen(N * N, 0); int nb[4], nn;
        for (int pos = 0; pos < N * N; ++pos) {
            if (board[pos] != 0 || seen[pos]) continue;
            std::vector<int> reg{pos}; seen[pos] = 1; bool tb = false, tw = false;
            for (size_t i = 0; i < reg.size(); ++i) {
                adj(reg[i], nb, nn);
                for (int k = 0; k < nn; ++k) {
                    int q = nb[k];
                    if (board[q] == 0) { if (!seen[q]) { seen[q] = 1; reg.push_back(q); } }
                    else if (board[q] == 1) tb = true; else tw = true;
                }
            }
            int col = (tb && !tw) ? 1 : (tw && !tb) ? 2 : 0;
            for (int s : reg) terr[s] = col;
        }
        for (int pos = 0; pos < N * N; ++pos) if (board[pos]) terr[pos] = board[pos];   // chinese rules
    }
    int result() const override {                              // getGameResult :320-335 + calculateScores
        if (!terminal()) return ONGOING;
        std::vector<int> terr; territory(terr);
        float bs = 0.0f, ws = 0.0f;
        for (int v : terr) { if (v == 1) bs += 1.0f; else if (v == 2) ws += 1.0f; }
        ws += komi;
        return bs > ws ? WIN_P1 : (ws > bs ? WIN_P2 : DRAW);
    }
    int player() const override { return cur; }
    int action_space() const override { return N * N + 1; }
    int board_size() const override { return N; }
    int game_type() const override { return 2; }
    uint64_t key() const override {
        uint64_t h = 1469598103934665603ULL;
        for (int pos = 0; pos < N * N; ++pos) h = mix64(h ^ (uint64_t)board[pos]);
        h = mix64(h ^ (uint64_t)cur);
        return mix64(h ^ (uint64_t)(int64_t)(ko + 1));
    }
    int planes() const override { return 8; }
    void tensor(float* t) const override {                     // go_state.cpp:349-445 (Go7)
        int S = N * N; std::fill(t, t + 8 * S, 0.0f);
        for (int pos = 0; pos < S; ++pos) { if (board[pos] == 1) t[pos] = 1.0f; else if (board[pos] == 2) t[S + pos] = 1.0f; }
        if (cur == 1) for (int pos = 0; pos < S; ++pos) t[2 * S + pos] = 1.0f;
        std::vector<char> done(S, 0); int nb[4], nn;
        for (int pos = 0; pos < S; ++pos) {
            if (!board[pos] || done[pos]) continue;
            std::vector<int> st; group_has_liberty(*this, board, pos, &st);
            std::unordered_set<int> libs;
            for (int s : st) { done[s] = 1; adj(s, nb, nn); for (int k = 0; k < nn; ++k) if (board[nb[k]] == 0) libs.insert(nb[k]); }
            float v = std::min(1.0f, (float)libs.size() / 10.0f);
            for (int s : st) t[(board[pos] == 1 ? 3 : 4) * S + s] = v;
        }
        if (ko >= 0) t[5 * S + ko] = 1.0f;
        for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
            t[6 * S + y * N + x] = (float)std::min(x, N - 1 - x) / (N / 2);
            t[7 * S + y * N + x] = (float)std::min(y, N - 1 - y) / (N / 2);
        }
    }
};

// -------------------------------------------------------------------------------------------- Chess
// The reference's chess is NOT runnable (makeMove → isLegalMove → moveExposesKing → cloneWithMove → makeMove … recurses
// without end), so this restatement cannot be pinned against oracle/_ref; it is pinned on the known answers
// the reference's own tests state (tests/games/chess/chess_state_test.cpp:27-65: 20 legal moves at the start, player
// alternation; :92-180 FEN positions) and on perft from the start position.  The recursion is cut where SURVEY says:
// moveExposesKing applies the move without a legality check.
struct Chess : State {
    enum { NONE = 0, PAWN = 1, KNIGHT = 2, BISHOP = 3, ROOK = 4, QUEEN = 5, KING = 6 };   // chess_state.h:20-36
    enum { WHITE = 1, BLACK = 2 };
    struct Pc { int8_t t = 0, c = 0; };
    struct Mv { int from, to, promo; };
    Pc b[64];                                   // square 0 = a8, rank = sq/8 counted from black's back rank (chess_rules.h:24-32)
    int cur = WHITE, ep = -1, half = 0;
    bool wk = true, wq = true, bk = true, bq = true;
    uint64_t h = 0;                             // QUIRK: hash_ is only ever updated by setPiece (chess_state.cpp:233-245); makeMove never
                                                // marks it dirty, so side to move / castling / en passant changes do not enter the
                                                // repetition key — it is the piece placement (plus constants that cancel)
    std::vector<std::pair<uint64_t, int>> pos_hist;   // position_history_ (hash → count)
    std::vector<int> hist;
    static bool& fide() { static bool f = false; return f; }   // false = literal isSquareAttacked pawn direction (QUIRK C5)

    Chess() { start(); }
    static uint64_t zk(const Pc& p, int sq) { return mix64(0xC0FFEEULL + (uint64_t)((p.t - 1) + (p.c == BLACK ? 6 : 0)) * 64 + sq); }
    void put(int sq, int t, int c) { if (b[sq].t) h ^= zk(b[sq], sq); b[sq].t = (int8_t)t; b[sq].c = (int8_t)(t ? c : 0); if (t) h ^= zk(b[sq], sq); }
    int& count(uint64_t k) { for (auto& e : pos_hist) if (e.first == k) return e.second; pos_hist.push_back({k, 0}); return pos_hist.back().second; }
    int seen(uint64_t k) const { for (auto& e : pos_hist) if (e.first == k) return e.second; return 0; }
    void clear() { for (auto& p : b) p = Pc(); h = 0; cur = WHITE; wk = wq = bk = bq = true; ep = -1; half = 0; pos_hist.clear(); hist.clear(); }
    void start() {                              // initializeStartingPosition chess_state.cpp:152-200
        clear();
        static const int back[8] = {ROOK, KNIGHT, BISHOP, QUEEN, KING, BISHOP, KNIGHT, ROOK};
        for (int f = 0; f < 8; ++f) { put(0 * 8 + f, back[f], BLACK); put(1 * 8 + f, PAWN, BLACK); put(6 * 8 + f, PAWN, WHITE); put(7 * 8 + f, back[f], WHITE); }
        count(h)++;                             // ctor: recordPosition() :66-67
    }
    bool set_fen(const char* fen) {             // test helper (setFromFEN): board / side / castling / ep / halfmove
        clear(); int r = 0, f = 0; const char* p = fen;
        for (; *p && *p != ' '; ++p) {
            if (*p == '/') { ++r; f = 0; continue; }
            if (*p >= '1' && *p <= '8') { f += *p - '0'; continue; }
            const char* names = "pnbrqk"; const char lc = (char)(*p | 32); const char* q = strchr(names, lc);
            if (!q || r > 7 || f > 7) return false;
            put(r * 8 + f, (int)(q - names) + 1, (*p & 32) ? BLACK : WHITE); ++f;
        }
        if (*p) ++p; cur = (*p == 'b') ? BLACK : WHITE; while (*p && *p != ' ') ++p; if (*p) ++p;
        wk = wq = bk = bq = false;
        for (; *p && *p != ' '; ++p) { if (*p == 'K') wk = true; if (*p == 'Q') wq = true; if (*p == 'k') bk = true; if (*p == 'q') bq = true; }
        if (*p) ++p;
        if (*p && *p != '-') { const int file = p[0] - 'a', rk = p[1] - '0'; if (file >= 0 && file < 8 && rk >= 1 && rk <= 8) ep = (8 - rk) * 8 + file; }
        while (*p && *p != ' ') ++p; if (*p) ++p;
        half = atoi(p);
        count(h)++;
        return true;
    }
    State* clone() const override { return new Chess(*this); }
    int player() const override { return cur; }
    int action_space() const override { return 64 * 64 * 5; }        // chess_state.h:117
    int board_size() const override { return 8; }
    int game_type() const override { return 1; }
    int planes() const override { return 18; }

    static bool on(int r, int f) { return r >= 0 && r < 8 && f >= 0 && f < 8; }
    bool attacked(int sq, int by) const {       // isSquareAttacked chess_rules.cpp:130-229: pawn, knight, king, diagonals, straights
        const int r = sq / 8, f = sq % 8;
        // QUIRK C5: the reference looks for `by`'s pawns at rank + pawnDir with pawnDir = (by == WHITE) ? -1 : +1, i.e. in the
        // direction those pawns MOVE — a white pawn really attacking (r, f) from (r+1, f±1) is not seen.  fide() flips it.
        const int pd = ((by == WHITE) ? -1 : 1) * (fide() ? -1 : 1);
        for (int df = -1; df <= 1; df += 2) if (on(r + pd, f + df)) { const Pc& a = b[(r + pd) * 8 + f + df]; if (a.t == PAWN && a.c == by) return true; }
        static const int KN[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}};
        static const int KG[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
        for (auto& d : KN) if (on(r + d[0], f + d[1])) { const Pc& a = b[(r + d[0]) * 8 + f + d[1]]; if (a.t == KNIGHT && a.c == by) return true; }
        for (auto& d : KG) if (on(r + d[0], f + d[1])) { const Pc& a = b[(r + d[0]) * 8 + f + d[1]]; if (a.t == KING && a.c == by) return true; }
        static const int BD[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}}, RD[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
        for (auto& d : BD) for (int k = 1; on(r + d[0] * k, f + d[1] * k); ++k) { const Pc& a = b[(r + d[0] * k) * 8 + f + d[1] * k]; if (a.t) { if (a.c == by && (a.t == BISHOP || a.t == QUEEN)) return true; break; } }
        for (auto& d : RD) for (int k = 1; on(r + d[0] * k, f + d[1] * k); ++k) { const Pc& a = b[(r + d[0] * k) * 8 + f + d[1] * k]; if (a.t) { if (a.c == by && (a.t == ROOK || a.t == QUEEN)) return true; break; } }
        return false;
    }
    int king_sq(int c) const { for (int s = 0; s < 64; ++s) if (b[s].t == KING && b[s].c == c) return s; return -1; }   // :1139-1147
    bool in_check(int c) const { const int k = king_sq(c); return k >= 0 && attacked(k, 3 - c); }                    // chess_rules.cpp:119-128
    void slide(std::vector<Mv>& m, int sq, const int (*dirs)[2], int nd) const {   // addSlidingMoves :559-590
        const int r = sq / 8, f = sq % 8;
        for (int i = 0; i < nd; ++i)
            for (int k = 1; on(r + dirs[i][0] * k, f + dirs[i][1] * k); ++k) {
                const int t = (r + dirs[i][0] * k) * 8 + f + dirs[i][1] * k;
                if (!b[t].t) m.push_back({sq, t, 0});
                else { if (b[t].c != cur) m.push_back({sq, t, 0}); break; }
            }
    }
    bool castle_ok(int from, int to) const {     // isValidCastle chess_rules.cpp:675-727 (standard chess: rook files 7 / 0)
        const int r = from / 8, ff = from % 8, tf = to % 8; const bool ks = tf > ff; const int rf = ks ? 7 : 0, rs = r * 8 + rf;
        if (b[rs].t != ROOK || b[rs].c != cur) return false;
        for (int f = std::min(ff, rf) + 1; f < std::max(ff, rf); ++f) if (b[r * 8 + f].t) return false;
        const int step = ks ? 1 : -1;
        for (int f = ff; f != tf + step; f += step) {
            const int s = r * 8 + f;
            if (s == from) continue;
            if (attacked(s, 3 - cur)) return false;
            if (s != rs && b[s].t) return false;
        }
        return true;
    }
    std::vector<Mv> pseudo() const {             // generatePseudoLegalMoves chess_rules.cpp:57-98: squares ascending, castling last
        std::vector<Mv> m;
        static const int KN[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}};
        static const int KG[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
        static const int BD[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}}, RD[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
        for (int sq = 0; sq < 64; ++sq) {
            if (b[sq].c != cur) continue;
            const int r = sq / 8, f = sq % 8;
            switch (b[sq].t) {
                case PAWN: {                     // addPawnMoves :470-528: push (promotions Q,R,B,N), double push, captures file-1 then +1, e.p.
                    const int d = cur == WHITE ? -1 : 1, nr = r + d;
                    if (nr >= 0 && nr < 8 && !b[nr * 8 + f].t) {
                        if (nr == 0 || nr == 7) for (int pr : {QUEEN, ROOK, BISHOP, KNIGHT}) m.push_back({sq, nr * 8 + f, pr});
                        else m.push_back({sq, nr * 8 + f, 0});
                        if ((cur == WHITE && r == 6) || (cur == BLACK && r == 1)) { const int t2 = (nr + d) * 8 + f; if (!b[t2].t) m.push_back({sq, t2, 0}); }
                    }
                    for (int df = -1; df <= 1; df += 2) {
                        const int nf = f + df;
                        if (!on(nr, nf)) continue;
                        const int t = nr * 8 + nf;
                        if (b[t].t && b[t].c != cur) {
                            if (nr == 0 || nr == 7) for (int pr : {QUEEN, ROOK, BISHOP, KNIGHT}) m.push_back({sq, t, pr});
                            else m.push_back({sq, t, 0});
                        }
                        if (ep == t) m.push_back({sq, t, 0});
                    }
                    break;
                }
                case KNIGHT: for (auto& dd : KN) if (on(r + dd[0], f + dd[1])) { const int t = (r + dd[0]) * 8 + f + dd[1]; if (!b[t].t || b[t].c != cur) m.push_back({sq, t, 0}); } break;
                case BISHOP: slide(m, sq, BD, 4); break;
                case ROOK: slide(m, sq, RD, 4); break;
                case QUEEN: slide(m, sq, KG, 8); break;       // QUEEN_DIRECTIONS == KING_MOVES order (chess_rules.cpp:15-30)
                case KING: for (auto& dd : KG) if (on(r + dd[0], f + dd[1])) { const int t = (r + dd[0]) * 8 + f + dd[1]; if (!b[t].t || b[t].c != cur) m.push_back({sq, t, 0}); } break;
                default: break;
            }
        }
        // addCastlingMoves :613-673: not in check, right still set, king target = file 4 +- 2 on the home rank
        if (!in_check(cur)) {
            const bool ck = cur == WHITE ? wk : bk, cq = cur == WHITE ? wq : bq;
            const int ks = king_sq(cur);
            if ((ck || cq) && ks >= 0) {
                const int home = cur == WHITE ? 7 : 0;
                if (ck && castle_ok(ks, home * 8 + 6)) m.push_back({ks, home * 8 + 6, 0});
                if (cq && castle_ok(ks, home * 8 + 2)) m.push_back({ks, home * 8 + 2, 0});
            }
        }
        return m;
    }
    void apply(const Mv& mv) {                   // makeMove(ChessMove) chess_state.cpp:976-1095 without the legality check
        Pc pc = b[mv.from]; const Pc cap = b[mv.to];
        half = (pc.t == PAWN || cap.t) ? 0 : half + 1;
        const int old_ep = ep; ep = -1;
        if (pc.t == PAWN) {
            const int fr = mv.from / 8, tr = mv.to / 8;
            if (std::abs(fr - tr) == 2) ep = ((fr + tr) / 2) * 8 + mv.from % 8;
            if (mv.to == old_ep) put((mv.from / 8) * 8 + mv.to % 8, 0, 0);
            if (mv.promo) pc.t = (int8_t)mv.promo;
        }
        if (pc.t == KING && std::abs(mv.from % 8 - mv.to % 8) == 2) {
            const int r = mv.from / 8; const bool ks = mv.to % 8 > mv.from % 8;
            const int rf = r * 8 + (ks ? 7 : 0), rt = r * 8 + (ks ? 5 : 3);
            const Pc rook = b[rf]; put(rf, 0, 0); put(rt, rook.t, rook.c);
        }
        // getUpdatedCastlingRights chess_rules.cpp:395-468
        if (pc.t == KING) { if (pc.c == WHITE) wk = wq = false; else bk = bq = false; }
        if (pc.t == ROOK) {
            const int f = mv.from % 8, r = mv.from / 8;
            if (pc.c == WHITE) { if (f == 7 && r == 7) wk = false; else if (f == 0 && r == 7) wq = false; }
            else { if (f == 7 && r == 0) bk = false; else if (f == 0 && r == 0) bq = false; }
        }
        if (cap.t == ROOK) {
            const int f = mv.to % 8, r = mv.to / 8;
            if (cap.c == WHITE) { if (f == 7 && r == 7) wk = false; else if (f == 0 && r == 7) wq = false; }
            else { if (f == 7 && r == 0) bk = false; else if (f == 0 && r == 0) bq = false; }
        }
        put(mv.from, 0, 0); put(mv.to, pc.t, pc.c);
        cur = 3 - cur;
        count(h)++;                              // recordPosition :1092-1095
    }
    std::vector<Mv> legal_moves() const {         // generateLegalMoves chess_rules.cpp:38-55 (+ the recursion cut in moveExposesKing)
        std::vector<Mv> out;
        for (const Mv& mv : pseudo()) { Chess t(*this); t.apply(mv); if (!t.in_check(cur)) out.push_back(mv); }
        return out;
    }
    static int code(const Mv& m) {                // chessMoveToAction chess_state.cpp:1214-1230 (C1)
        const int pc = m.promo == QUEEN ? 1 : m.promo == ROOK ? 2 : m.promo == BISHOP ? 3 : m.promo == KNIGHT ? 4 : 0;
        return (pc << 12) | (m.from << 6) | m.to;
    }
    static Mv decode(int a) { static const int P[5] = {0, QUEEN, ROOK, BISHOP, KNIGHT}; const int pc = (a >> 12) & 7; return {(a >> 6) & 63, a & 63, pc <= 4 ? P[pc] : 0}; }
    std::vector<int> legal() const override { std::vector<int> r; for (const Mv& m : legal_moves()) r.push_back(code(m)); return r; }
    bool make_move(int a) override {
        if (a < 0 || a >= action_space()) return false;
        const Mv want = decode(a);
        for (const Mv& m : legal_moves()) if (m.from == want.from && m.to == want.to && m.promo == want.promo) { apply(m); hist.push_back(a); return true; }
        return false;                              // "Illegal move attempted" (chess_state.cpp:977-979)
    }
    bool insufficient() const {                   // hasInsufficientMaterial chess_rules.cpp:231-389
        int n = 0, P[3] = {0, 0, 0}, N[3] = {0, 0, 0}, B[3] = {0, 0, 0}, R[3] = {0, 0, 0}, Q[3] = {0, 0, 0}; bool light[3] = {false, false, false}, dark[3] = {false, false, false};
        for (int s = 0; s < 64; ++s) {
            if (!b[s].t) continue; ++n; const int c = b[s].c; const bool lt = ((s / 8 + s % 8) % 2 == 0);
            switch (b[s].t) { case PAWN: ++P[c]; break; case KNIGHT: ++N[c]; break; case BISHOP: ++B[c]; (lt ? light[c] : dark[c]) = true; break; case ROOK: ++R[c]; break; case QUEEN: ++Q[c]; break; default: break; }
        }
        const int W = WHITE, K = BLACK;
        const bool noP = !P[W] && !P[K], noN = !N[W] && !N[K], noB = !B[W] && !B[K], noR = !R[W] && !R[K], noQ = !Q[W] && !Q[K];
        if (n == 2) return true;
        if (((N[W] == 1 && N[K] == 0) || (N[W] == 0 && N[K] == 1)) && noP && noB && noR && noQ) return true;
        if (((B[W] == 1 && B[K] == 0) || (B[W] == 0 && B[K] == 1)) && noP && noN && noR && noQ) return true;
        if (noP && noN && B[W] == 1 && B[K] == 1 && noR && noQ && ((light[W] && light[K]) || (dark[W] && dark[K]))) return true;
        if (N[W] == 2 && N[K] == 0 && noP && noB && noR && noQ) return true;
        if (N[W] == 0 && N[K] == 2 && noP && noB && noR && noQ) return true;
        if (N[W] == 1 && N[K] == 1 && noP && noB && noR && noQ) return true;
        if (((N[W] == 1 && B[K] == 1 && N[K] == 0 && B[W] == 0) || (N[K] == 1 && B[W] == 1 && N[W] == 0 && B[K] == 0)) && noP && noR && noQ) return true;
        return false;
    }
    int result() const override {                 // isTerminal chess_state.cpp:599-652: no moves, material, 50-move, threefold
        if (legal_moves().empty()) return in_check(cur) ? (cur == WHITE ? WIN_P2 : WIN_P1) : DRAW;
        if (insufficient() || half >= 100 || seen(h) >= 3) return DRAW;
        return ONGOING;
    }
    bool terminal() const override { return result() != ONGOING; }
    // QUIRK C8: ChessState::getHash() is recomputed in full only while hash_dirty_ (ctor / setFromFEN); afterwards setPiece updates it
    // incrementally (chess_state.cpp:230-241) and makeMove never marks it dirty, so side to move, castling rights and the e.p. square
    // of the CURRENT position never enter it.  The reference's TranspositionTable therefore returns the cached policy / value of
    // whichever position with the same piece placement was evaluated first in the game — also one with the other side to move.
    uint64_t tt_key() const override { return h; }
    uint64_t key() const override {               // HashEvaluator key for chess (this repo's definition; the survey probe had none)
        uint64_t k = 1469598103934665603ULL;
        for (int s = 0; s < 64; ++s) k = mix64(k ^ (uint64_t)(b[s].t + 8 * b[s].c));
        k = mix64(k ^ (uint64_t)cur);
        k = mix64(k ^ (uint64_t)((wk ? 1 : 0) | (wq ? 2 : 0) | (bk ? 4 : 0) | (bq ? 8 : 0)));
        return mix64(k ^ (uint64_t)(int64_t)(ep + 1));
    }
    void tensor(float* t) const override {        // getEnhancedTensorRepresentation chess_state.cpp:665-769 (C4), [plane][rank][file]
        std::fill(t, t + 18 * 64, 0.0f);
        for (int s = 0; s < 64; ++s) if (b[s].t) t[((b[s].t - 1) + (b[s].c == BLACK ? 6 : 0)) * 64 + s] = 1.0f;
        const float cast = (wk ? 0.25f : 0.0f) + (wq ? 0.25f : 0.0f) + (bk ? 0.25f : 0.0f) + (bq ? 0.25f : 0.0f);
        const float hm = std::min(1.0f, (float)half / 100.0f), rep = (float)seen(h) / 3.0f;
        for (int s = 0; s < 64; ++s) { t[12 * 64 + s] = cur == WHITE ? 1.0f : 0.0f; t[13 * 64 + s] = cast; t[15 * 64 + s] = hm; t[17 * 64 + s] = rep; }
        if (ep >= 0 && ep < 64) t[14 * 64 + ep] = 1.0f;
    }
    long perft(int d) const { if (d == 0) return 1; long n = 0; for (const Mv& m : legal_moves()) { Chess t(*this); t.apply(m); n += t.perft(d - 1); } return n; }
};

// ------------------------------------------------------------------------------------ HashEvaluator
// peaked (evaluator 2): the raw prior of action mix(h ^ 0x5EED) % A is multiplied by 4096 before the normalisation — a deterministic
// evaluator with a sharply peaked policy (deep, narrow trees: what a trained network produces), same exactly-rounded fp32 ops
static void hash_eval(uint64_t h, int A, float* policy, float* value, bool peaked = false) {
    float sum = 0.0f;
    const int peak = peaked ? (int)(mix64(h ^ 0x5EEDULL) % (uint64_t)A) : -1;
    for (int i = 0; i < A; ++i) {
        uint64_t r = mix64(h + (uint64_t)i * 0x9E3779B97F4A7C15ULL) >> 40;
        policy[i] = (float)(r + 1) / (float)(1 << 24);
        if (i == peak) policy[i] = policy[i] * 4096.0f;
        sum += policy[i];
    }
    for (int i = 0; i < A; ++i) policy[i] = policy[i] / sum;
    float v = ((float)(mix64(h ^ 0xABCDEFULL) >> 40) / (float)(1 << 24)) * 2.0f - 1.0f;
    *value = v * 0.5f;
}

// --------------------------------------------------------------------------------- serial search
typedef void (*eval_cb_t)(const float* planes, int C, int H, int W, int A, float* policy, float* value, void* user);

struct Node {                          // MCTSNode, include/alphazero/mcts/mcts_node.h:54-75
    int N = 0, VL = 0; float W = 0.0f, P = 0.0f;
    int action = -1, parent = -1, first = -1, nchild = 0;
    bool expanded = false, term = false; int result = ONGOING;
};

struct Search {
    std::unique_ptr<State> root_state; std::vector<Node> pool; int root = 0;
    int sims = 800; float cpuct = 1.5f; int vl = 3; int max_depth = 1000;
    eval_cb_t cb = nullptr; void* user = nullptr; long evals = 0; bool peaked = false;
    // TranspositionTable (transposition_table.cpp:44-84, 128-176) as the serial search sees it: lookup by full 64-bit hash, first store
    // wins ("existing-key store keeps the old policy").  Slot collisions (hash & (size - 1)) and the age-based replacement are not
    // modelled: with 1 M slots and a few thousand entries per game they do not occur in the pinned runs.  One table per game, kept
    // across moves (self_play_manager.cpp:159,175).  use_tt = false gives the TT-free search the device engine implements.
    bool use_tt = true; long tt_hits = 0;
    std::unordered_map<uint64_t, std::pair<std::vector<float>, float>> tt;
    bool tt_lookup(const State& s, std::vector<float>& pol, float& v) {
        if (!use_tt) return false;
        auto it = tt.find(s.tt_key()); if (it == tt.end()) return false;
        pol = it->second.first; v = it->second.second; ++tt_hits; return true;
    }
    void tt_store(const State& s, const std::vector<float>& pol, float v) { if (use_tt) tt.emplace(s.tt_key(), std::make_pair(pol, v)); }

    Search(const State& s, int sims_, float c, int vl_) : root_state(s.clone()), sims(sims_), cpuct(c), vl(vl_) {
        // root MCTSNode ctor evaluates state->isTerminal()/getGameResult() (mcts_node.cpp:24-25);
        // for Gomoku that call is also the first fill of the legal-move cache (QUIRK G2).
        Node r; r.term = root_state->terminal(); r.result = root_state->result(); pool.push_back(r);
    }
    static float to_value(int result, int pl) {            // convertToValue parallel_mcts.cpp:973-985
        if (result == WIN_P1) return pl == 1 ? 1.0f : -1.0f;
        if (result == WIN_P2) return pl == 2 ? 1.0f : -1.0f;
        return 0.0f;
    }
    void evaluate(const State& s, std::vector<float>& pol, float& v) {
        ++evals; int A = s.action_space(); pol.assign(A, 0.0f);
        if (cb) { int C = s.planes(), n = s.board_size(); std::vector<float> t((size_t)C * n * n); s.tensor(t.data()); cb(t.data(), C, n, n, A, pol.data(), &v, user); }
        else hash_eval(s.key(), A, pol.data(), &v, peaked);
    }
    void expand(int ni, const State& s, const std::vector<float>& pol) {   // expandNodeWithPolicy :681-745 (M5)
        if (pool[ni].expanded || pool[ni].term) return;
        std::vector<int> legal = s.legal();
        if (legal.empty()) { pool[ni].term = true; pool[ni].result = s.result(); pool[ni].expanded = true; return; }
        float sum = 0.0f; std::vector<float> lp(legal.size(), 0.0f);
        for (size_t i = 0; i < legal.size(); ++i) { int a = legal[i]; if (a >= 0 && a < (int)pol.size()) { lp[i] = pol[a]; sum += lp[i]; } }
        if (sum > 0.0f) for (auto& p : lp) p /= sum; else { float u = 1.0f / (float)legal.size(); for (auto& p : lp) p = u; }
        int first = (int)pool.size();
        for (size_t i = 0; i < legal.size(); ++i) { Node c; c.P = lp[i]; c.action = legal[i]; c.parent = ni; pool.push_back(c); }
        pool[ni].first = first; pool[ni].nchild = (int)legal.size(); pool[ni].expanded = true;
    }
    float score(const Node& c, int parentVisits) const {   // getPuctScore mcts_node.cpp:61-119 (M4)
        int visits = c.N; if (visits == 0) return FLT_MAX;
        float q = 0.0f; int actual = visits - c.VL; if (actual > 0) q = c.W / actual;
        const Node& par = pool[c.parent];
        if (par.parent >= 0 && pool[par.parent].parent < 0) q = -q;        // QUIRK: depth-2 children only
        float u = cpuct * c.P * std::sqrt((float)parentVisits) / (1.0f + visits);
        float d = 0.0f; if (visits < 5) d = 0.05f * (5 - visits);
        return q + u + d;
    }
    void add_vl(Node& n) { n.N += vl; n.VL += vl; n.W = n.W - (float)vl; }      // mcts_node.cpp:168-181 (M6)
    void rem_vl(Node& n) { n.N -= vl; n.VL -= vl; n.W = n.W + (float)vl; }      // :183-196
    void simulate() {                                       // runSingleSimulation :276-380 (M2)
        std::unique_ptr<State> st(root_state->clone()); std::vector<int> path;
        int ni = root; add_vl(pool[ni]); path.push_back(ni); int depth = 0;     // selectLeafWithPath :456-535 (M3)
        while (pool[ni].expanded && !pool[ni].term && depth < max_depth) {
            int pv = pool[ni].N; float best = -FLT_MAX; int bi = -1;            // selectChildPuct :537-563
            for (int i = 0; i < pool[ni].nchild; ++i) { float sc = score(pool[pool[ni].first + i], pv); if (sc > best) { best = sc; bi = pool[ni].first + i; } }
            if (bi < 0) break;
            if (!st->make_move(pool[bi].action)) { rem_vl(pool[ni]); path.pop_back(); break; }
            path.push_back(bi); ni = bi; ++depth;
        }
        for (int p : path) add_vl(pool[p]);                 // :293-295 (root gets it twice)
        float value = 0.0f;
        if (pool[ni].term || st->terminal()) {              // :300-313 (M10)
            if (pool[ni].term) value = to_value(pool[ni].result, st->player());
            else { value = to_value(st->result(), st->player()); pool[ni].term = true; pool[ni].result = st->result(); }
        } else {                                            // :316-358: TT first, else evaluate + store
            std::vector<float> pol;
            if (!tt_lookup(*st, pol, value)) { evaluate(*st, pol, value); tt_store(*st, pol, value); }
            expand(ni, *st, pol);
        }
        float cv = value;                                   // backpropagate :782-833 (M7)
        for (auto it = path.rbegin(); it != path.rend(); ++it) { Node& n = pool[*it]; rem_vl(n); n.N += 1; n.W = n.W + cv; cv = -cv; }
    }
    void search() {                                         // search :142-274 (M1), serial
        if (!pool[root].expanded && !root_state->terminal()) {      // :153-163: evaluateState (TT lookup inside, :848-856), store, expand
            std::vector<float> pol; float v;
            if (!tt_lookup(*root_state, pol, v)) { evaluate(*root_state, pol, v); tt_store(*root_state, pol, v); }
            expand(root, *root_state, pol);
        }
        for (int i = 0; i < sims; ++i) simulate();
    }
    std::vector<float> probs(float T) const {               // getVisitCountDistribution mcts_node.cpp:289-322 (M12)
        const Node& r = pool[root]; std::vector<float> d(r.nchild, 0.0f); if (!r.expanded || r.nchild == 0) return d;
        float total = 0.0f; std::vector<float> c(r.nchild);
        for (int i = 0; i < r.nchild; ++i) { c[i] = std::pow((float)pool[r.first + i].N, 1.0f / std::max(0.01f, T)); total += c[i]; }
        if (total > 0.0f) for (int i = 0; i < r.nchild; ++i) d[i] = c[i] / total; else for (auto& x : d) x = 1.0f / (float)r.nchild;
        return d;
    }
    int select_action(bool training, float T) {             // selectAction :987-1047 (M13), deterministic branches
        if (!pool[root].expanded) search();
        const Node& r = pool[root];
        if (r.term || r.nchild == 0) { auto m = root_state->legal(); return m.empty() ? -1 : m[0]; }
        if (training && T > 0.0f) { auto d = probs(T); int bi = (int)(std::max_element(d.begin(), d.end()) - d.begin()); return pool[r.first + bi].action; }
        int mx = 0; for (int i = 0; i < r.nchild; ++i) mx = std::max(mx, pool[r.first + i].N);
        for (int i = 0; i < r.nchild; ++i) if (pool[r.first + i].N == mx) return pool[r.first + i].action;
        return -1;
    }
    float root_value() const { const Node& r = pool[root]; if (r.nchild == 0) return 0.0f; return r.N == 0 ? 0.0f : r.W / r.N; }  // :1057-1063 (M15)
    void update_with_move(int a) {                          // updateWithMove :1065-1108 (M14)
        const Node& r = pool[root]; int child = -1;
        for (int i = 0; i < r.nchild; ++i) if (pool[r.first + i].action == a) { child = r.first + i; break; }
        if (!root_state->make_move(a)) return;
        if (child >= 0) { pool[child].parent = -1; root = child; }
        else { Node n; n.term = root_state->terminal(); n.result = root_state->result(); pool.push_back(n); root = (int)pool.size() - 1; }
    }
};

}  // namespace orc

using namespace orc;

extern "C" {
void* orc_state_new(int game_type, int board_size) {
    if (game_type == 0) return new Gomoku(board_size);
    if (game_type == 2) return new Go(board_size);
    if (game_type == 1) return new Chess();
    return nullptr;
}
void orc_state_free(void* h) { delete (State*)h; }
void* orc_state_clone(void* h) { return ((State*)h)->clone(); }
int orc_state_make_move(void* h, int a) { return ((State*)h)->make_move(a) ? 0 : -1; }
int orc_state_legal_moves(void* h, int* out, int cap) { auto m = ((State*)h)->legal(); int n = (int)m.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = m[i]; return n; }
int orc_state_is_terminal(void* h) { return ((State*)h)->terminal() ? 1 : 0; }
int orc_state_result(void* h) { return ((State*)h)->result(); }
int orc_state_current_player(void* h) { return ((State*)h)->player(); }
int orc_state_action_space(void* h) { return ((State*)h)->action_space(); }
int orc_state_board_size(void* h) { return ((State*)h)->board_size(); }
int orc_state_tensor(void* h, float* out) { State* s = (State*)h; if (out) s->tensor(out); return s->planes(); }
uint64_t orc_state_key(void* h) { return ((State*)h)->key(); }
void orc_hash_eval(void* h, float* policy, float* value) { State* s = (State*)h; hash_eval(s->key(), s->action_space(), policy, value); }
int orc_chess_set_fen(void* h, const char* fen) { return ((Chess*)h)->set_fen(fen) ? 0 : -1; }
void orc_chess_set_fide(int on) { Chess::fide() = on != 0; }
long orc_chess_perft(void* h, int depth) { return ((Chess*)h)->perft(depth); }
int orc_chess_piece(void* h, int sq) { const Chess* c = (Chess*)h; return c->b[sq].t + 8 * c->b[sq].c; }
int orc_chess_in_check(void* h) { const Chess* c = (Chess*)h; return c->in_check(c->cur) ? 1 : 0; }
int orc_go_stone(void* h, int pos) { return ((Go*)h)->board[pos]; }
int orc_go_ko(void* h) { return ((Go*)h)->ko; }
int orc_go_captured(void* h, int pl) { return ((Go*)h)->captured[pl]; }

void* orc_mcts_new(void* state, int sims, float cpuct, int vl, int evaluator, eval_cb_t cb, void* user) {
    Search* s = new Search(*(State*)state, sims, cpuct, vl);
    if (evaluator == 1) { s->cb = cb; s->user = user; }
    if (evaluator == 2) s->peaked = true;
    return s;
}
void orc_mcts_free(void* h) { delete (Search*)h; }
void orc_mcts_search(void* h) { ((Search*)h)->search(); }
void orc_mcts_set_sims(void* h, int sims) { ((Search*)h)->sims = sims; }
long orc_mcts_eval_calls(void* h) { return ((Search*)h)->evals; }
// 1 (default) = with the reference's TranspositionTable semantics; 0 = TT-free (every leaf evaluated on its own input: what the device engine does)
void orc_mcts_set_tt(void* h, int on) { ((Search*)h)->use_tt = on != 0; }
long orc_mcts_tt_hits(void* h) { return ((Search*)h)->tt_hits; }
int orc_mcts_root_stats(void* h, int* actions, int* N, float* W, float* P, int cap, int* rootN, float* rootW) {
    Search* s = (Search*)h; const Node& r = s->pool[s->root];
    for (int i = 0; i < r.nchild && i < cap; ++i) { const Node& c = s->pool[r.first + i]; actions[i] = c.action; N[i] = c.N; W[i] = c.W; P[i] = c.P; }
    if (rootN) *rootN = r.N; if (rootW) *rootW = r.W;
    return r.nchild;
}
int orc_mcts_select_action(void* h, int training, float T) { return ((Search*)h)->select_action(training != 0, T); }
int orc_mcts_action_probs(void* h, float T, float* out, int cap) { auto p = ((Search*)h)->probs(T); int n = (int)p.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = p[i]; return n; }
float orc_mcts_root_value(void* h) { return ((Search*)h)->root_value(); }
void orc_mcts_update_with_move(void* h, int a) { ((Search*)h)->update_with_move(a); }

// Dataset::augmentExample (src/selfplay/dataset.cpp:245-436): the 7 extra examples in the reference's order — rot90, rot180,
// rot270, flipH of the original, then flipH of rot90 / rot180 / rot270 — every plane and the first N*N policy entries moved
// by the same index map, policy entries past N*N copied unchanged.  planes [C][N][N], policy [A] → out [7][...].
static void orc_move(const float* pl, const float* po, int C, int N, int A, float* opl, float* opo, int kind) {
    for (int a = 0; a < A; ++a) opo[a] = po[a];                               // `TrainingExample t = example;` then overwrite
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            int i2, j2;
            if (kind == 0) { i2 = j; j2 = N - 1 - i; }                         // rot90   :264-283
            else if (kind == 1) { i2 = N - 1 - i; j2 = N - 1 - j; }            // rot180  :286-305
            else if (kind == 2) { i2 = N - 1 - j; j2 = i; }                    // rot270  :308-327
            else { i2 = i; j2 = N - 1 - j; }                                   // flipH   :330-349
            for (int p = 0; p < C; ++p) opl[(p * N + i2) * N + j2] = pl[(p * N + i) * N + j];
            const int oi = i * N + j, ni = i2 * N + j2;
            if (oi < A && ni < A) opo[ni] = po[oi];
        }
}
void orc_augment_example(const float* planes, int C, int N, const float* policy, int A, float* out_planes, float* out_policy) {
    const size_t ps = (size_t)C * N * N;
    for (int k = 0; k < 3; ++k) orc_move(planes, policy, C, N, A, out_planes + k * ps, out_policy + (size_t)k * A, k);
    orc_move(planes, policy, C, N, A, out_planes + 3 * ps, out_policy + (size_t)3 * A, 3);
    for (int k = 0; k < 3; ++k) orc_move(out_planes + k * ps, out_policy + (size_t)k * A, C, N, A, out_planes + (4 + k) * ps, out_policy + (size_t)(4 + k) * A, 3);   // :352-433
}

// First-fill legal order of a fresh Gomoku N×N state (QUIRK G2): literally std::unordered_set<int>
// with ascending inserts of the given empty cells.
int orc_first_fill_order(const int* empties, int n, int* out) {
    std::unordered_set<int> s; for (int i = 0; i < n; ++i) s.insert(empties[i]);
    int k = 0; for (int a : s) out[k++] = a; return k;
}
}  // extern "C"
