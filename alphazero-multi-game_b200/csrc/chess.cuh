// chess.cuh — chess rules (host + device) behind the per-warp game API of tree_kernels.cuh.
// Reference: src/games/chess/chess_rules.cpp + chess_state.cpp.  The reference's chess cannot run
// (makeMove → isLegalMove → moveExposesKing → cloneWithMove → makeMove … recurses without end); the checker is
// the CPU restatement in oracle/az_oracle.cpp (struct Chess), pinned on the reference tests' known answers and perft.
//
// Mailbox board (64 bytes, square 0 = a8, rank = sq / 8 counted from black's back rank).  Pseudo-legal moves are generated
// by lane 0 in the reference's order (squares ascending, per-piece direction tables, castling last); the legality filter —
// apply the move to a private copy, is the own king attacked — runs one move per lane; survivors are compacted in order.
// Reproduced quirks: C5 (the pawn part of isSquareAttacked looks in the direction the attacker's pawns MOVE), the
// repetition key (piece placement only: makeMove never marks hash_ dirty, chess_state.cpp:233-245 / :976-1095), en passant
// target set on every double push, castling target files 6 / 2 of the home rank.
#pragma once
#include "common.cuh"
#if defined(__CUDACC__)
#include <cuda_bf16.h>
#endif

namespace az {

struct Chess {
    static constexpr int N = 8;
    static constexpr int CELLS = 64;
    static constexpr int ACTIONS = 64 * 64 * 5;        // promo << 12 | from << 6 | to (chess_state.h:117, chess_state.cpp:1199-1230)
    static constexpr int MAX_CHILDREN = 256;           // >= 218, the most legal moves a chess position can have
    static constexpr int SAMPLE_VISITS = 2 * MAX_CHILDREN;   // (action, count) pairs in child order: 20480 counts would be 40 KB
    static constexpr int PLANES = 18;
    static constexpr bool FIRST_FILL = false;
    static constexpr bool LEGAL_POLICY = true;         // 20480-wide policy head: inside the waves only the legal moves' logits are computed
    static constexpr bool TT_COARSE = true;            // QUIRK C8: the reference's TT key is the piece placement only (tree.cuh EvalTT)
    static constexpr int MAX_GAME_MOVES = 512;         // engine cap: the game is drawn at this ply
    static constexpr int MAXH = MAX_GAME_MOVES + 2;
    static constexpr int EXTRA = 62;
    enum { NONE = 0, PAWN = 1, KNIGHT = 2, BISHOP = 3, ROOK = 4, QUEEN = 5, KING = 6 };
    enum { WHITE = 1, BLACK = 2 };
    enum { R_WK = 1, R_WQ = 2, R_BK = 4, R_BQ = 8 };

    struct Core {
        uint8_t b[64];             // type | colour << 3
        uint64_t key;              // repetition key: XOR of per-(piece, square) keys
        int16_t ep;                // en passant target square, -1 = none
        int16_t half;              // halfmove clock
        int16_t ply;
        int16_t hist_n;            // keys in the root lineage's history (every position reached, the initial one included)
        int8_t player;             // 1 = WHITE to move
        int8_t rights;             // R_* bits
        int8_t n_extra;
        int8_t pad_[5];
    };
    struct State { Core c; uint64_t hist[MAXH]; };
    struct Leaf { Core c; uint64_t extra[EXTRA]; };
    struct Snapshot { uint8_t b[64]; int16_t ep, half, ply; int8_t player, rights, reps, pad_; };   // reps: repetition count of the position (plane 17)

    AZ_HD static int T(uint8_t p) { return p & 7; }
    AZ_HD static int Cc(uint8_t p) { return p >> 3; }
    AZ_HD static bool on(int r, int f) { return r >= 0 && r < 8 && f >= 0 && f < 8; }
    AZ_HD static uint64_t zk(uint8_t p, int sq) { return mix64(0xC0FFEEULL + (uint64_t)((T(p) - 1) + (Cc(p) == BLACK ? 6 : 0)) * 64 + (uint64_t)sq); }
    AZ_HD static void put(Core& c, int sq, int t, int col) {
        if (c.b[sq]) c.key ^= zk(c.b[sq], sq);
        c.b[sq] = (uint8_t)(t ? (t | (col << 3)) : 0);
        if (t) c.key ^= zk(c.b[sq], sq);
    }
    AZ_HD static int code(int from, int to, int promo) { return ((promo == QUEEN ? 1 : promo == ROOK ? 2 : promo == BISHOP ? 3 : promo == KNIGHT ? 4 : 0) << 12) | (from << 6) | to; }
    AZ_HD static int promo_of(int a) { const int pc = (a >> 12) & 7; return pc == 1 ? QUEEN : pc == 2 ? ROOK : pc == 3 ? BISHOP : pc == 4 ? KNIGHT : 0; }
    AZ_HD static void dir8(int i, int& dr, int& df) {      // KING_MOVES == QUEEN_DIRECTIONS order (chess_rules.cpp:15-30)
        dr = i < 3 ? -1 : (i < 5 ? 0 : 1);
        df = i < 3 ? i - 1 : (i == 3 ? -1 : (i == 4 ? 1 : i - 6));
    }
    AZ_HD static void knight(int i, int& dr, int& df) {    // KNIGHT_MOVES order (chess_rules.cpp:10-12)
        dr = i < 2 ? -2 : (i < 4 ? -1 : (i < 6 ? 1 : 2));
        df = (i < 2 || i >= 6) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
    }
    AZ_HD static void bishop(int i, int& dr, int& df) { dr = i < 2 ? -1 : 1; df = (i & 1) ? 1 : -1; }                        // BISHOP_DIRECTIONS
    AZ_HD static void rook(int i, int& dr, int& df) { dr = i == 0 ? -1 : (i == 1 ? 1 : 0); df = i == 2 ? -1 : (i == 3 ? 1 : 0); }   // ROOK_DIRECTIONS

    AZ_HD static void init_core(Core& c) {               // initializeStartingPosition, chess_state.cpp:152-200
        for (int i = 0; i < 64; ++i) c.b[i] = 0;
        c.key = 0; c.ep = -1; c.half = 0; c.ply = 0; c.hist_n = 0; c.player = WHITE; c.rights = 15; c.n_extra = 0;
        for (int i = 0; i < 5; ++i) c.pad_[i] = 0;
        const int back[8] = {ROOK, KNIGHT, BISHOP, QUEEN, KING, BISHOP, KNIGHT, ROOK};
        for (int f = 0; f < 8; ++f) { put(c, f, back[f], BLACK); put(c, 8 + f, PAWN, BLACK); put(c, 48 + f, PAWN, WHITE); put(c, 56 + f, back[f], WHITE); }
    }
    // isSquareAttacked (chess_rules.cpp:130-229): pawn, knight, king, diagonal sliders, straight sliders
    AZ_HD static bool attacked(const Core& c, int sq, int by) {
        const int r = sq >> 3, f = sq & 7;
        const int pd = by == WHITE ? -1 : 1;              // QUIRK C5: the direction `by`'s pawns move
        for (int df = -1; df <= 1; df += 2) if (on(r + pd, f + df)) { const uint8_t a = c.b[(r + pd) * 8 + f + df]; if (T(a) == PAWN && Cc(a) == by) return true; }
        for (int i = 0; i < 8; ++i) { int dr, df; knight(i, dr, df); if (on(r + dr, f + df)) { const uint8_t a = c.b[(r + dr) * 8 + f + df]; if (T(a) == KNIGHT && Cc(a) == by) return true; } }
        for (int i = 0; i < 8; ++i) { int dr, df; dir8(i, dr, df); if (on(r + dr, f + df)) { const uint8_t a = c.b[(r + dr) * 8 + f + df]; if (T(a) == KING && Cc(a) == by) return true; } }
        for (int i = 0; i < 4; ++i) {
            int dr, df; bishop(i, dr, df);
            for (int k = 1; on(r + dr * k, f + df * k); ++k) { const uint8_t a = c.b[(r + dr * k) * 8 + f + df * k]; if (a) { if (Cc(a) == by && (T(a) == BISHOP || T(a) == QUEEN)) return true; break; } }
        }
        for (int i = 0; i < 4; ++i) {
            int dr, df; rook(i, dr, df);
            for (int k = 1; on(r + dr * k, f + df * k); ++k) { const uint8_t a = c.b[(r + dr * k) * 8 + f + df * k]; if (a) { if (Cc(a) == by && (T(a) == ROOK || T(a) == QUEEN)) return true; break; } }
        }
        return false;
    }
    AZ_HD static int king_sq(const Core& c, int col) {         // getKingSquare chess_state.cpp:1139-1147: the first square holding that king
#if defined(__CUDA_ARCH__)
        const uint32_t pat = 0x01010101u * (uint32_t)(KING | (col << 3));
        const uint32_t* wd = reinterpret_cast<const uint32_t*>(c.b);           // four squares per word, byte-wise compare
        for (int i = 0; i < 16; ++i) { const uint32_t m = __vcmpeq4(wd[i], pat); if (m) return i * 4 + ((__ffs(m) - 1) >> 3); }
        return -1;
#else
        for (int s = 0; s < 64; ++s) if (c.b[s] == (uint8_t)(KING | (col << 3))) return s;
        return -1;
#endif
    }
    AZ_HD static bool in_check(const Core& c, int col) { const int k = king_sq(c, col); return k >= 0 && attacked(c, k, 3 - col); }
    // isValidCastle (chess_rules.cpp:675-727), standard chess: rook files 7 / 0
    AZ_HD static bool castle_ok(const Core& c, int from, int to) {
        const int cur = c.player, r = from >> 3, ff = from & 7, tf = to & 7; const bool ks = tf > ff; const int rf = ks ? 7 : 0, rs = r * 8 + rf;
        if (c.b[rs] != (uint8_t)(ROOK | (cur << 3))) return false;
        for (int f = (ff < rf ? ff : rf) + 1; f < (ff > rf ? ff : rf); ++f) if (c.b[r * 8 + f]) return false;
        const int step = ks ? 1 : -1;
        for (int f = ff; f != tf + step; f += step) {
            const int s = r * 8 + f;
            if (s == from) continue;
            if (attacked(c, s, 3 - cur)) return false;
            if (s != rs && c.b[s]) return false;
        }
        return true;
    }
    // generatePseudoLegalMoves (chess_rules.cpp:57-98, :470-673) as action codes, in the reference's order: squares ascending (gen_square: the
    // moves of the piece on one square, at most 27), castling last (gen_castling)
    AZ_HD static int gen_square(const Core& c, int sq, int16_t* out) {
        int n = 0; const int cur = c.player;
        const uint8_t p = c.b[sq];
        if (!p || Cc(p) != cur) return 0;
        const int r = sq >> 3, f = sq & 7;
        switch (T(p)) {
            case PAWN: {
                const int d = cur == WHITE ? -1 : 1, nr = r + d;
                if (nr >= 0 && nr < 8 && !c.b[nr * 8 + f]) {
                    if (nr == 0 || nr == 7) { out[n++] = (int16_t)code(sq, nr * 8 + f, QUEEN); out[n++] = (int16_t)code(sq, nr * 8 + f, ROOK); out[n++] = (int16_t)code(sq, nr * 8 + f, BISHOP); out[n++] = (int16_t)code(sq, nr * 8 + f, KNIGHT); }
                    else out[n++] = (int16_t)code(sq, nr * 8 + f, 0);
                    if ((cur == WHITE && r == 6) || (cur == BLACK && r == 1)) { const int t2 = (nr + d) * 8 + f; if (!c.b[t2]) out[n++] = (int16_t)code(sq, t2, 0); }
                }
                for (int df = -1; df <= 1; df += 2) {
                    const int nf = f + df;
                    if (!on(nr, nf)) continue;
                    const int t = nr * 8 + nf;
                    if (c.b[t] && Cc(c.b[t]) != cur) {
                        if (nr == 0 || nr == 7) { out[n++] = (int16_t)code(sq, t, QUEEN); out[n++] = (int16_t)code(sq, t, ROOK); out[n++] = (int16_t)code(sq, t, BISHOP); out[n++] = (int16_t)code(sq, t, KNIGHT); }
                        else out[n++] = (int16_t)code(sq, t, 0);
                    }
                    if (c.ep == t) out[n++] = (int16_t)code(sq, t, 0);
                }
                break;
            }
            case KNIGHT:
            case KING:
                for (int i = 0; i < 8; ++i) {
                    int dr, df; if (T(p) == KNIGHT) knight(i, dr, df); else dir8(i, dr, df);
                    if (!on(r + dr, f + df)) continue;
                    const int t = (r + dr) * 8 + f + df;
                    if (!c.b[t] || Cc(c.b[t]) != cur) out[n++] = (int16_t)code(sq, t, 0);
                }
                break;
            default: {                                   // sliders (addSlidingMoves :559-590)
                const int nd = T(p) == QUEEN ? 8 : 4;
                for (int i = 0; i < nd; ++i) {
                    int dr, df; if (T(p) == QUEEN) dir8(i, dr, df); else if (T(p) == BISHOP) bishop(i, dr, df); else rook(i, dr, df);
                    for (int k = 1; on(r + dr * k, f + df * k); ++k) {
                        const int t = (r + dr * k) * 8 + f + df * k;
                        if (!c.b[t]) out[n++] = (int16_t)code(sq, t, 0);
                        else { if (Cc(c.b[t]) != cur) out[n++] = (int16_t)code(sq, t, 0); break; }
                    }
                }
            }
        }
        return n;
    }
    AZ_HD static int gen_castling(const Core& c, int16_t* out) {   // addCastlingMoves :613-673
        int n = 0; const int cur = c.player;
        if (!in_check(c, cur)) {
            const bool ck = c.rights & (cur == WHITE ? R_WK : R_BK), cq = c.rights & (cur == WHITE ? R_WQ : R_BQ);
            const int ks = king_sq(c, cur);
            if ((ck || cq) && ks >= 0) {
                const int home = cur == WHITE ? 7 : 0;
                if (ck && castle_ok(c, ks, home * 8 + 6)) out[n++] = (int16_t)code(ks, home * 8 + 6, 0);
                if (cq && castle_ok(c, ks, home * 8 + 2)) out[n++] = (int16_t)code(ks, home * 8 + 2, 0);
            }
        }
        return n;
    }
    AZ_HD static int gen_pseudo(const Core& c, int16_t* out) {
        int n = 0;
        for (int sq = 0; sq < 64; ++sq) n += gen_square(c, sq, out + n);
        return n + gen_castling(c, out + n);
    }
    // makeMove(ChessMove) without the legality check (chess_state.cpp:976-1095); returns the key recordPosition() stores
    AZ_HD static uint64_t apply_core(Core& c, int a) {
        const int from = (a >> 6) & 63, to = a & 63, promo = promo_of(a);
        uint8_t pc = c.b[from]; const uint8_t cap = c.b[to];
        const int col = Cc(pc);
        c.half = (int16_t)((T(pc) == PAWN || cap) ? 0 : c.half + 1);
        const int old_ep = c.ep; c.ep = -1;
        if (T(pc) == PAWN) {
            const int fr = from >> 3, tr = to >> 3;
            if (fr - tr == 2 || tr - fr == 2) c.ep = (int16_t)(((fr + tr) / 2) * 8 + (from & 7));
            if (to == old_ep) put(c, (from >> 3) * 8 + (to & 7), 0, 0);
            if (promo) pc = (uint8_t)(promo | (col << 3));
        }
        if (T(pc) == KING && ((from & 7) - (to & 7) == 2 || (to & 7) - (from & 7) == 2)) {
            const int r = from >> 3; const bool ks = (to & 7) > (from & 7);
            const int rf = r * 8 + (ks ? 7 : 0), rt = r * 8 + (ks ? 5 : 3);
            const uint8_t rk = c.b[rf]; put(c, rf, 0, 0); put(c, rt, T(rk), Cc(rk));
        }
        if (T(pc) == KING) c.rights &= (int8_t)(col == WHITE ? ~(R_WK | R_WQ) : ~(R_BK | R_BQ));       // getUpdatedCastlingRights :395-468
        if (T(pc) == ROOK) {
            const int f = from & 7, r = from >> 3;
            if (col == WHITE) { if (f == 7 && r == 7) c.rights &= ~R_WK; else if (f == 0 && r == 7) c.rights &= ~R_WQ; }
            else { if (f == 7 && r == 0) c.rights &= ~R_BK; else if (f == 0 && r == 0) c.rights &= ~R_BQ; }
        }
        if (T(cap) == ROOK) {
            const int f = to & 7, r = to >> 3;
            if (Cc(cap) == WHITE) { if (f == 7 && r == 7) c.rights &= ~R_WK; else if (f == 0 && r == 7) c.rights &= ~R_WQ; }
            else { if (f == 7 && r == 0) c.rights &= ~R_BK; else if (f == 0 && r == 0) c.rights &= ~R_BQ; }
        }
        put(c, from, 0, 0); put(c, to, T(pc), Cc(pc));
        c.player = (int8_t)(3 - c.player);
        c.ply = (int16_t)(c.ply + 1);
        return c.key;
    }
    // moveExposesKing with the recursion cut (chess_rules.cpp:745-751)
    AZ_HD static bool exposes(const Core& c, int a) { Core t = c; apply_core(t, a); return in_check(t, c.player); }
    AZ_HD static bool insufficient(const Core& c) {      // hasInsufficientMaterial chess_rules.cpp:231-389
        int n = 0, P[3] = {0, 0, 0}, Nn[3] = {0, 0, 0}, B[3] = {0, 0, 0}, R[3] = {0, 0, 0}, Q[3] = {0, 0, 0}; bool light[3] = {false, false, false}, dark[3] = {false, false, false};
        for (int s = 0; s < 64; ++s) {
            const uint8_t p = c.b[s]; if (!p) continue; ++n; const int col = Cc(p); const bool lt = (((s >> 3) + (s & 7)) % 2 == 0);
            switch (T(p)) { case PAWN: ++P[col]; break; case KNIGHT: ++Nn[col]; break; case BISHOP: ++B[col]; if (lt) light[col] = true; else dark[col] = true; break; case ROOK: ++R[col]; break; case QUEEN: ++Q[col]; break; default: break; }
        }
        const int W = WHITE, K = BLACK;
        const bool noP = !P[W] && !P[K], noN = !Nn[W] && !Nn[K], noB = !B[W] && !B[K], noR = !R[W] && !R[K], noQ = !Q[W] && !Q[K];
        if (n == 2) return true;
        if (((Nn[W] == 1 && Nn[K] == 0) || (Nn[W] == 0 && Nn[K] == 1)) && noP && noB && noR && noQ) return true;
        if (((B[W] == 1 && B[K] == 0) || (B[W] == 0 && B[K] == 1)) && noP && noN && noR && noQ) return true;
        if (noP && noN && B[W] == 1 && B[K] == 1 && noR && noQ && ((light[W] && light[K]) || (dark[W] && dark[K]))) return true;
        if (Nn[W] == 2 && Nn[K] == 0 && noP && noB && noR && noQ) return true;
        if (Nn[W] == 0 && Nn[K] == 2 && noP && noB && noR && noQ) return true;
        if (Nn[W] == 1 && Nn[K] == 1 && noP && noB && noR && noQ) return true;
        if (((Nn[W] == 1 && B[K] == 1 && Nn[K] == 0 && B[W] == 0) || (Nn[K] == 1 && B[W] == 1 && Nn[W] == 0 && B[K] == 0)) && noP && noR && noQ) return true;
        return false;
    }
    AZ_HD static int repetitions(const Core& c, const uint64_t* hist, const uint64_t* extra) {
        int n = 0;
        for (int i = 0; i < c.hist_n; ++i) n += hist[i] == c.key;
        for (int i = 0; i < c.n_extra; ++i) n += extra[i] == c.key;
        return n;
    }
    // isTerminal / getGameResult (chess_state.cpp:599-652) given the number of legal moves
    AZ_HD static int result_core(const Core& c, int n_legal, int reps) {
        if (n_legal == 0) return in_check(c, c.player) ? (c.player == WHITE ? RES_WIN_P2 : RES_WIN_P1) : RES_DRAW;
        if (insufficient(c) || c.half >= 100 || reps >= 3) return RES_DRAW;
        return RES_ONGOING;
    }
    AZ_HD static uint64_t key_core(const Core& c) {       // HashEvaluator key (oracle/az_oracle.cpp Chess::key)
        uint64_t k = 1469598103934665603ULL;
        for (int s = 0; s < 64; ++s) k = mix64(k ^ (uint64_t)(T(c.b[s]) + 8 * Cc(c.b[s])));
        k = mix64(k ^ (uint64_t)c.player);
        k = mix64(k ^ (uint64_t)(c.rights & 15));
        return mix64(k ^ (uint64_t)(int64_t)(c.ep + 1));
    }
    // getEnhancedTensorRepresentation (chess_state.cpp:665-769, C4), value of plane pl at square sq
    AZ_HD static float feature(const Core& c, int pl, int sq, int reps) {
        if (pl < 12) { const uint8_t p = c.b[sq]; return (p && (T(p) - 1) + (Cc(p) == BLACK ? 6 : 0) == pl) ? 1.0f : 0.0f; }
        switch (pl) {
            case 12: return c.player == WHITE ? 1.0f : 0.0f;
            case 13: return ((c.rights & R_WK) ? 0.25f : 0.0f) + ((c.rights & R_WQ) ? 0.25f : 0.0f) + ((c.rights & R_BK) ? 0.25f : 0.0f) + ((c.rights & R_BQ) ? 0.25f : 0.0f);
            case 14: return c.ep == sq ? 1.0f : 0.0f;
            case 15: return fminf(1.0f, (float)c.half / 100.0f);
            case 16: return 0.0f;
            default: return (float)reps / 3.0f;
        }
    }

    // ---------------------------------------------------------------------------------------------- host helpers (single thread)
    static void init(State& s) { init_core(s.c); s.hist[0] = s.c.key; s.c.hist_n = 1; }          // ctor: recordPosition()
    static int host_legal(const State& s, int16_t* out) {
        int16_t ps[MAX_CHILDREN]; const int np = gen_pseudo(s.c, ps); int n = 0;
        for (int i = 0; i < np; ++i) if (!exposes(s.c, ps[i])) out[n++] = ps[i];
        return n;
    }
    static bool host_apply(State& s, int a) {
        int16_t lg[MAX_CHILDREN]; const int n = host_legal(s, lg); bool ok = false;
        for (int i = 0; i < n; ++i) ok |= (lg[i] == a);
        if (!ok) return false;
        const uint64_t k = apply_core(s.c, a);
        if (s.c.hist_n < MAXH) s.hist[s.c.hist_n++] = k;
        return true;
    }
    static int host_root_result(const State& s) {
        int16_t lg[MAX_CHILDREN]; const int n = host_legal(s, lg);
        const int r = result_core(s.c, n, repetitions(s.c, s.hist, nullptr));
        return (r == RES_ONGOING && s.c.ply >= MAX_GAME_MOVES) ? RES_DRAW : r;
    }
    static int host_ply(const State& s) { return s.c.ply; }
    static int host_player(const State& s) { return s.c.player; }

#if defined(__CUDACC__)
    // ---------------------------------------------------------------------------------------------- warp API (tree_kernels.cuh)
    struct Warp { Leaf s; const uint64_t* hist; uint64_t* hist_rw; int16_t pseudo[MAX_CHILDREN]; int16_t legal[MAX_CHILDREN]; int n_legal; };
    struct EncTarget { __nv_bfloat16* ptr; int p_total, guard, board_pitch, f16; };

    __device__ static void copy_words(void* dst, const void* src, int bytes, int lane) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = lane; i < bytes / 4; i += 32) d[i] = s[i];
    }
    __device__ static void w_init(Warp& w, int lane) {
        if (lane == 0) { init_core(w.s.c); w.s.extra[0] = w.s.c.key; w.s.c.n_extra = 1; w.hist = nullptr; w.hist_rw = nullptr; }    // ctor: recordPosition()
        __syncwarp();
    }
    __device__ static void w_attach_history(Warp& w, uint64_t* buf, int lane) {
        if (lane == 0) { for (int i = 0; i < w.s.c.n_extra; ++i) buf[w.s.c.hist_n + i] = w.s.extra[i]; w.s.c.hist_n = (int16_t)(w.s.c.hist_n + w.s.c.n_extra); w.s.c.n_extra = 0; w.hist = buf; w.hist_rw = buf; }
        __syncwarp();
    }
    __device__ static void w_load_root(Warp& w, const State* g, int lane) {
        copy_words(&w.s.c, &g->c, sizeof(Core), lane);
        __syncwarp();
        if (lane == 0) { w.s.c.n_extra = 0; w.hist = g->hist; w.hist_rw = nullptr; }
        __syncwarp();
    }
    __device__ static void w_store_root(Warp& w, State* g, int lane) {
        __syncwarp();
        const int ne = w.s.c.n_extra, hn = w.s.c.hist_n;
        for (int i = lane; i < ne; i += 32) if (hn + i < MAXH) g->hist[hn + i] = w.s.extra[i];
        __syncwarp();
        if (lane == 0) { w.s.c.hist_n = (int16_t)min(hn + ne, MAXH); w.s.c.n_extra = 0; }
        __syncwarp();
        copy_words(&g->c, &w.s.c, sizeof(Core), lane);
    }
    __device__ static void w_store_leaf(Warp& w, Leaf* g, int lane) { __syncwarp(); copy_words(g, &w.s, sizeof(Core) + 8 * w.s.c.n_extra, lane); }
    __device__ static void w_load_leaf(Warp& w, const Leaf* g, const State* root, int lane) {
        copy_words(&w.s.c, &g->c, sizeof(Core), lane);
        __syncwarp();
        copy_words(w.s.extra, g->extra, 8 * w.s.c.n_extra, lane);
        if (lane == 0) { w.hist = root->hist; w.hist_rw = nullptr; }
        __syncwarp();
    }
    __device__ static void w_snapshot(const Warp& w, Snapshot* out, int lane) {
        for (int i = lane; i < 16; i += 32) reinterpret_cast<uint32_t*>(out->b)[i] = reinterpret_cast<const uint32_t*>(w.s.c.b)[i];
        if (lane == 0) { out->ep = w.s.c.ep; out->half = w.s.c.half; out->ply = w.s.c.ply; out->player = w.s.c.player; out->rights = w.s.c.rights; out->reps = (int8_t)repetitions(w.s.c, w.hist, w.s.extra); out->pad_ = 0; }
    }
    // legal moves of the warp's state into w.legal (reference order); pseudo-legal by lane 0, legality one move per lane
    __device__ static int w_gen_legal(Warp& w, int lane) {
        __syncwarp();
        // pseudo-legal moves: one square per lane (two rounds: squares 0-31, 32-63), each lane's moves placed behind those of the lower squares
        // (exclusive prefix sum of the per-square counts) — the reference's square-ascending order; castling last, by lane 0
        int np = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            int16_t mv[28];
            const int n = gen_square(w.s.c, half * 32 + lane, mv);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int off = np + incl - n;
            for (int i = 0; i < n; ++i) w.pseudo[off + i] = mv[i];
            np += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        int nc = 0;
        if (lane == 0) nc = gen_castling(w.s.c, w.pseudo + np);
        np += __shfl_sync(0xffffffffu, nc, 0);
        __syncwarp();
        int cnt = 0;
        for (int k = 0; k < np; k += 32) {
            const int i = k + lane;
            const bool ok = i < np && !exposes(w.s.c, w.pseudo[i]);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) w.legal[cnt + __popc(m & ((1u << lane) - 1))] = w.pseudo[i];
            cnt += __popc(m);
        }
        __syncwarp();
        if (lane == 0) w.n_legal = cnt;
        __syncwarp();
        return cnt;
    }
    __device__ static bool w_apply(Warp& w, int a, int lane, bool check = false) {
        bool ok = true;
        if (check) {
            ok = false;
            if (a >= 0 && a < ACTIONS) { const int n = w_gen_legal(w, lane); for (int i = 0; i < n; ++i) ok |= (w.legal[i] == a); }
        }
        __syncwarp();
        if (ok && lane == 0) {
            const uint64_t k = apply_core(w.s.c, a);
            if (w.hist_rw) w.hist_rw[w.s.c.hist_n++] = k;
            else if (w.s.c.n_extra < EXTRA) w.s.extra[w.s.c.n_extra++] = k;
        }
        __syncwarp();
        return ok;
    }
    // position count of the current placement key (isThreefoldRepetition / plane 17), the history scanned by the whole (converged) warp
    __device__ static int w_reps(Warp& w, int lane) {
        int n = 0; const uint64_t k = w.s.c.key;
        for (int i = lane; i < w.s.c.hist_n; i += 32) n += w.hist[i] == k;
        for (int i = lane; i < w.s.c.n_extra; i += 32) n += w.s.extra[i] == k;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        return n;
    }
    // hasInsufficientMaterial: every drawn pattern has at most four pieces on the board — count them with two ballots first
    __device__ static bool w_insufficient(Warp& w, int lane) {
        const int n = __popc(__ballot_sync(0xffffffffu, w.s.c.b[lane] != 0)) + __popc(__ballot_sync(0xffffffffu, w.s.c.b[lane + 32] != 0));
        return n <= 4 && insufficient(w.s.c);
    }
    __device__ static int w_result(Warp& w, int lane) {          // isTerminal / getGameResult chess_state.cpp:599-652, same order of tests as result_core
        const int n = w_gen_legal(w, lane);
        if (n == 0) return in_check(w.s.c, w.s.c.player) ? (w.s.c.player == WHITE ? RES_WIN_P2 : RES_WIN_P1) : RES_DRAW;
        if (w_insufficient(w, lane) || w.s.c.half >= 100 || w_reps(w, lane) >= 3) return RES_DRAW;
        return RES_ONGOING;
    }
    __device__ static int w_root_result(Warp& w, int lane) { const int r = w_result(w, lane); return (r == RES_ONGOING && w.s.c.ply >= MAX_GAME_MOVES) ? RES_DRAW : r; }
    __device__ static int w_player(const Warp& w) { return w.s.c.player; }
    __device__ static int w_ply(const Warp& w) { return w.s.c.ply; }
    // training sample: (action, visit count) pairs in child order
    __device__ static void record_visit(uint16_t* visits, int child, int action, int n) { visits[2 * child] = (uint16_t)action; visits[2 * child + 1] = (uint16_t)min(n, 65535); }
    __device__ static int w_enumerate(Warp& w, int lane, const int16_t*, int, int16_t* acts, float* raw, const float* pol) {
        const int n = w_gen_legal(w, lane);
        for (int i = lane; i < n; i += 32) { const int a = w.legal[i]; acts[i] = (int16_t)a; raw[i] = pol ? pol[a] : 0.0f; }   // pol == nullptr: priors come from the evaluation cache
        __syncwarp();
        return n;
    }
    // the legal list for the policy head; regen = false: w.legal is current (w_result / w_gen_legal ran on this state)
    __device__ static int w_store_legal(Warp& w, int lane, int16_t* out, bool regen) {
        const int n = regen ? w_gen_legal(w, lane) : w.n_legal;
        for (int i = lane; i < n; i += 32) out[i] = w.legal[i];
        return n;
    }
    __device__ static int w_enumerate_pre(int lane, const int16_t* pre, int n, int16_t* acts, float* raw, const float* pol) {
        for (int i = lane; i < n; i += 32) { const int a = pre[i]; acts[i] = (int16_t)a; raw[i] = pol ? pol[a] : 0.0f; }
        __syncwarp();
        return n;
    }
    __device__ static int w_legal(Warp& w, int lane, int32_t* out) {
        const int n = w_gen_legal(w, lane);
        for (int i = lane; i < n; i += 32) out[i] = w.legal[i];
        return n;
    }
    // feature planes straight into the conv trunk's input layout (bf16, 32 channels = 18 + 14 zero); row = rank * 9 + file
    __device__ static void w_encode(Warp& w, int lane, const EncTarget& enc, int slot) {
        const int reps = w_reps(w, lane);
        const size_t row0 = (size_t)enc.guard + (size_t)slot * enc.board_pitch;
        for (int sq = lane; sq < 64; sq += 32) {
            const size_t row = row0 + (size_t)(sq >> 3) * (N + 1) + (sq & 7);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                __align__(16) __nv_bfloat16 v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { const int pl = q * 8 + e; v[e] = net16(pl < PLANES ? feature(w.s.c, pl, sq, reps) : 0.0f, enc.f16 != 0); }
                *reinterpret_cast<uint4*>(enc.ptr + ((size_t)q * enc.p_total + row) * 8) = *reinterpret_cast<const uint4*>(v);
            }
        }
    }
    __device__ static void w_planes(Warp& w, int lane, float* out) {           // fp32 [18][rank][file]
        const int reps = w_reps(w, lane);
        for (int i = lane; i < PLANES * 64; i += 32) out[i] = feature(w.s.c, i / 64, i % 64, reps);
    }
    __device__ static uint64_t w_key(Warp& w, int) { return key_core(w.s.c); }
    __device__ static uint64_t w_tt_key(Warp& w) { return w.s.c.key; }           // ChessState::getHash() granularity: the placement
    // profiling (AZ_EVAL_DUP_STATS): the 18 planes = pieces, side, castling, e.p., min(1, halfmove / 100), repetitions / 3
    // (the placement part is the incrementally maintained Zobrist-style key: O(1) instead of key_core's 64-step chain)
    __device__ static uint64_t w_input_key(Warp& w, int lane) {
        const Core& c = w.s.c;
        const uint64_t rest = (uint64_t)c.player | ((uint64_t)(c.rights & 15) << 2) | ((uint64_t)(c.ep + 1) << 6) | ((uint64_t)min((int)c.half, 100) << 13) | ((uint64_t)w_reps(w, lane) << 20);
        return mix64(c.key ^ mix64(rest + 0x9E3779B97F4A7C15ULL));
    }
    __device__ static uint64_t w_ref_tt_key(Warp& w) { return w.s.c.key; }
    // training examples (az_engine_make_examples); the repetition count travels in the snapshot (w.n_legal doubles as its holder)
    __device__ static void w_from_snapshot(Warp& w, const Snapshot* g, int lane) {
        for (int i = lane; i < 16; i += 32) reinterpret_cast<uint32_t*>(w.s.c.b)[i] = reinterpret_cast<const uint32_t*>(g->b)[i];
        if (lane == 0) { w.s.c.key = 0; w.s.c.ep = g->ep; w.s.c.half = g->half; w.s.c.ply = g->ply; w.s.c.hist_n = 0; w.s.c.player = g->player; w.s.c.rights = g->rights; w.s.c.n_extra = 0; w.hist = nullptr; w.hist_rw = nullptr; w.n_legal = g->reps; }
        __syncwarp();
    }
    __device__ static float tensor_value(Warp& w, int c, int i, int j) { return feature(w.s.c, c, i * 8 + j, w.n_legal); }
    __device__ static int policy_total(const uint16_t* visits, int lane) {
        int t = 0; for (int k = lane; k < MAX_CHILDREN; k += 32) t += visits[2 * k + 1];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        return t;
    }
    template <class F> __device__ static void policy_for_each(const uint16_t* visits, int lane, F f) {
        for (int k = lane; k < MAX_CHILDREN; k += 32) if (visits[2 * k] != 0 || visits[2 * k + 1] != 0) f((int)visits[2 * k], (int)visits[2 * k + 1]);
    }
#endif
};

}  // namespace az
