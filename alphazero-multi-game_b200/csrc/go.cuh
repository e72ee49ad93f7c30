// go.cuh — Go rules on padded bitboards (host + device): capture, ko, positional superko, suicide, area score,
// 8-plane encoding.  Reference: src/games/go/go_state.cpp + go_rules.cpp.
//
// Replaces the reference's int board + whole-board BFS per move (go_rules.cpp:144-181) and its clone-per-candidate legal
// move generation (go_state.cpp:116-154) by flood fills on bitboards: cell pos = y*N + x (reference action index) sits
// at bit p = y*(N+1) + x, column N is a permanent hole, so a shift by 1 / N+1 is the 4-neighbourhood with no wrap.  The
// same p is the row of the cell inside the conv trunk's padded position stream (conv_trunk.cuh), as for Gomoku.
//
// One warp owns one tree (tree_kernels.cuh).  Moves are applied by lane 0 on the warp's shared-memory copy of the state;
// legal-move generation runs one candidate per lane (each lane floods the groups around its own candidate in registers).
// Positional superko (QUIRK Go3): candidate key = Zobrist(board after placement and captures) ^ mover ^ OLD ko point,
// compared with the keys pushed by earlier moves (board ^ mover ^ NEW ko point); passes push nothing.
#pragma once
#include "common.cuh"
#if defined(__CUDACC__)
#include <cuda_bf16.h>
#endif

namespace az {

template <int N_>
struct Go {
    static constexpr int N = N_;
    static constexpr int PITCH = N + 1;
    static constexpr int CELLS = N * N;
    static constexpr int PBITS = N * PITCH;
    static constexpr int NW = (PBITS + 63) / 64;
    static constexpr int ACTIONS = CELLS + 1;          // getActionSpaceSize (go_state.cpp:345-347); index N*N is never used
    static constexpr int MAX_CHILDREN = CELLS + 1;     // pass + every cell
    static constexpr int SAMPLE_VISITS = CELLS + 1;    // visit counts by action, pass last
    static constexpr int PLANES = 8;
    static constexpr bool LEGAL_POLICY = false;
    static constexpr bool TT_COARSE = false;          // the reference's TT key (stones + player + ko) covers the hash evaluator's input
    static constexpr bool FIRST_FILL = false;          // legal-move order is the same for every enumeration (QUIRK Go2)
    static constexpr int MAX_GAME_MOVES = 2 * CELLS;   // engine cap (the reference has none): the game is scored at this ply
    static constexpr int MAXH = MAX_GAME_MOVES + 2;    // superko keys of the root lineage
    static constexpr int EXTRA = 62;                   // superko keys added below the root inside one simulation (depth)
    static constexpr float KOMI = 7.5f;                // go_state.h:38 (Chinese rules, superko on)

    struct BB { uint64_t w[NW]; };

    struct Core {                  // everything but the superko history
        BB bb[2];                  // [0] = BLACK (player 1), [1] = WHITE (player 2)
        uint64_t key;              // Zobrist of the stones
        int16_t ko;                // reference pos of the ko point, -1 = none
        int16_t passes;            // consecutive_passes_
        int16_t ply;
        int16_t hist_n;            // number of keys in the root lineage's history
        int8_t player;             // 1 = BLACK to move
        int8_t n_extra;            // keys in `extra` (Leaf only)
        int8_t pad_[6];
    };
    struct State { Core c; uint64_t hist[MAXH]; };         // root of a slot
    struct Leaf { Core c; uint64_t extra[EXTRA]; };        // root + path moves of one simulation
    struct Snapshot { BB bb[2]; int16_t ko, passes, ply; int8_t player, pad_; };   // what a training sample keeps

    // ---------------------------------------------------------------------------------------------- bit helpers
    AZ_HD static int a2p(int a) { return (a / N) * PITCH + (a % N); }
    AZ_HD static int p2a(int p) { return (p / PITCH) * N + (p % PITCH); }
    AZ_HD static bool get(const BB& b, int p) { return (b.w[p >> 6] >> (p & 63)) & 1; }
    AZ_HD static void setb(BB& b, int p) { b.w[p >> 6] |= 1ULL << (p & 63); }
    AZ_HD static BB zero() { BB r; for (int i = 0; i < NW; ++i) r.w[i] = 0; return r; }
    AZ_HD static bool any(const BB& b) { uint64_t o = 0; for (int i = 0; i < NW; ++i) o |= b.w[i]; return o != 0; }
    AZ_HD static bool same(const BB& a, const BB& b) { uint64_t o = 0; for (int i = 0; i < NW; ++i) o |= a.w[i] ^ b.w[i]; return o == 0; }
    AZ_HD static int popc(const BB& b) {
        int t = 0;
        for (int i = 0; i < NW; ++i) {
#if defined(__CUDA_ARCH__)
            t += __popcll(b.w[i]);
#else
            t += __builtin_popcountll(b.w[i]);
#endif
        }
        return t;
    }
    AZ_HD static int lowest(const BB& b) {      // index of the lowest set bit (b must be non-empty)
        for (int i = 0; i < NW; ++i)
            if (b.w[i]) {
#if defined(__CUDA_ARCH__)
                return i * 64 + __ffsll((long long)b.w[i]) - 1;
#else
                return i * 64 + __builtin_ctzll(b.w[i]);
#endif
            }
        return -1;
    }
    static constexpr uint64_t valid_word(int i) {   // real cells of word i (pad column and bits >= PBITS cleared)
        uint64_t m = 0;
        for (int b = 0; b < 64; ++b) { const int p = i * 64 + b; if (p < PBITS && (p % PITCH) < N) m |= 1ULL << b; }
        return m;
    }
    AZ_HD static BB shl(const BB& b, int d) {
        BB r;
#pragma unroll
        for (int i = 0; i < NW; ++i) { uint64_t v = b.w[i] << d; if (i > 0) v |= b.w[i - 1] >> (64 - d); r.w[i] = v; }
        return r;
    }
    AZ_HD static BB shr(const BB& b, int d) {
        BB r;
#pragma unroll
        for (int i = 0; i < NW; ++i) { uint64_t v = b.w[i] >> d; if (i + 1 < NW) v |= b.w[i + 1] << (64 - d); r.w[i] = v; }
        return r;
    }
    // 4-neighbourhood of a set, restricted to real cells (getAdjacentPositions, go_state.cpp:773-790)
    AZ_HD static BB nb(const BB& b, const BB& valid) {
        const BB a = shl(b, 1), c = shr(b, 1), d = shl(b, PITCH), e = shr(b, PITCH);
        BB r;
#pragma unroll
        for (int i = 0; i < NW; ++i) r.w[i] = (a.w[i] | c.w[i] | d.w[i] | e.w[i]) & valid.w[i];
        return r;
    }
    template <int I> struct VW { static constexpr uint64_t v = valid_word(I < NW ? I : 0); };      // compile-time constants
    AZ_HD static BB valid_bb() {
        BB b;
        b.w[0] = VW<0>::v;
        if constexpr (NW > 1) b.w[1] = VW<1>::v;
        if constexpr (NW > 2) b.w[2] = VW<2>::v;
        if constexpr (NW > 3) b.w[3] = VW<3>::v;
        if constexpr (NW > 4) b.w[4] = VW<4>::v;
        if constexpr (NW > 5) b.w[5] = VW<5>::v;
        static_assert(NW <= 6, "boards up to 19x19");
        return b;
    }
    // connected component of `seed` inside `mask` (findGroups' BFS, go_rules.cpp:144-181)
    AZ_HD static BB flood(BB g, const BB& mask, const BB& valid) {
        for (int it = 0; it < CELLS; ++it) {
            const BB n = nb(g, valid);
            BB g2; uint64_t diff = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) { g2.w[i] = g.w[i] | (n.w[i] & mask.w[i]); diff |= g2.w[i] ^ g.w[i]; }
            g = g2;
            if (!diff) break;
        }
        return g;
    }
    AZ_HD static uint64_t zob(int a, int colour) { return mix64(0x1234567ULL + (uint64_t)a * 2 + (uint64_t)(colour - 1)); }
    AZ_HD static uint64_t zob_set(const BB& s, int colour) {
        uint64_t h = 0;
        for (int i = 0; i < NW; ++i) {
            uint64_t w = s.w[i];
            while (w) {
#if defined(__CUDA_ARCH__)
                const int b = __ffsll((long long)w) - 1;
#else
                const int b = __builtin_ctzll(w);
#endif
                w &= w - 1;
                h ^= zob(p2a(i * 64 + b), colour);
            }
        }
        return h;
    }
    // superko key (QUIRK Go3): stones ^ mover ^ ko point
    AZ_HD static uint64_t pos_key(uint64_t stones_key, int mover, int ko) {
        uint64_t h = stones_key ^ mix64(0xABCD0000ULL + (uint64_t)mover);
        if (ko >= 0) h ^= mix64(0xFEED0000ULL + (uint64_t)ko);
        return h;
    }

    // ---------------------------------------------------------------------------------------------- rules (one thread)
    AZ_HD static void init_core(Core& c) {
        c.bb[0] = zero(); c.bb[1] = zero(); c.key = 0; c.ko = -1; c.passes = 0; c.ply = 0; c.hist_n = 0; c.player = 1; c.n_extra = 0;
        for (int i = 0; i < 6; ++i) c.pad_[i] = 0;
    }
    // Place a stone of `mover` on the empty cell a: which opponent stones die, how many groups, is it suicide?
    // isSuicidalMove (go_rules.cpp:29-138) + the capture part of makeMove (go_state.cpp:190-261).
    AZ_HD static void place(const Core& c, int a, const BB& valid, BB& cap, int& ngroups, bool& suicide) {
        const int me = c.player - 1;
        const int p = a2p(a);
        BB own = c.bb[me]; setb(own, p);
        const BB& opp = c.bb[1 - me];
        BB empty;
#pragma unroll
        for (int i = 0; i < NW; ++i) empty.w[i] = valid.w[i] & ~(own.w[i] | opp.w[i]);
        cap = zero(); ngroups = 0;
        BB one = zero(); setb(one, p);
        BB adj = nb(one, valid);
#pragma unroll
        for (int i = 0; i < NW; ++i) adj.w[i] &= opp.w[i];
        while (any(adj)) {
            const int q = lowest(adj);
            BB seed = zero(); setb(seed, q);
            const BB g = flood(seed, opp, valid);
            const BB lib = nb(g, valid);
            uint64_t l = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) { l |= lib.w[i] & empty.w[i]; adj.w[i] &= ~g.w[i]; }
            if (!l) { ++ngroups; for (int i = 0; i < NW; ++i) cap.w[i] |= g.w[i]; }
        }
        suicide = false;
        if (ngroups == 0) {
            const BB g = flood(one, own, valid);
            const BB lib = nb(g, valid);
            uint64_t l = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) l |= lib.w[i] & empty.w[i];
            suicide = (l == 0);
        }
    }
    // getLegalMoves' per-cell test (go_state.cpp:116-154, 814-845): empty, not the ko point, not suicide, not a superko
    // repetition.  hist = keys of the root lineage, extra = keys pushed below the root in this simulation.
    AZ_HD static bool legal_cell(const Core& c, int a, const BB& valid, const uint64_t* hist, int hist_n, const uint64_t* extra, int n_extra) {
        if (a < 0 || a >= CELLS) return false;
        const int p = a2p(a);
        if (get(c.bb[0], p) || get(c.bb[1], p) || a == c.ko) return false;
        BB cap; int ng; bool suicide;
        place(c, a, valid, cap, ng, suicide);
        if (suicide) return false;
        const uint64_t k = pos_key(c.key ^ zob(a, c.player) ^ zob_set(cap, 3 - c.player), c.player, c.ko);   // OLD ko point
        bool rep = false;
        for (int i = 0; i < hist_n; ++i) rep |= (hist[i] == k);
        for (int i = 0; i < n_extra; ++i) rep |= (extra[i] == k);
        return !rep;
    }
    // makeMove (go_state.cpp:190-261) for a move already known to be legal.  Returns true and sets `pushed` when a superko
    // key has to be appended to the history (every move except a pass).
    AZ_HD static bool apply_core(Core& c, int a, const BB& valid, uint64_t& pushed) {
        bool push = false;
        if (a == -1) { c.passes = (int16_t)(c.passes + 1); c.ko = -1; }
        else {
            BB cap; int ng; bool suicide;
            place(c, a, valid, cap, ng, suicide);
            const int me = c.player - 1;
            setb(c.bb[me], a2p(a));
#pragma unroll
            for (int i = 0; i < NW; ++i) c.bb[1 - me].w[i] &= ~cap.w[i];
            c.key ^= zob(a, c.player) ^ zob_set(cap, 3 - c.player);
            c.ko = (ng == 1 && popc(cap) == 1) ? (int16_t)p2a(lowest(cap)) : (int16_t)-1;      // Go4
            c.passes = 0;
            pushed = pos_key(c.key, c.player, c.ko);                                           // before the player switch, NEW ko
            push = true;
        }
        c.player = (int8_t)(3 - c.player);
        c.ply = (int16_t)(c.ply + 1);
        return push;
    }
    // area score (getTerritoryOwnership / calculateScores, go_rules.cpp:196-359) → GameResult; caller checked terminality
    AZ_HD static int score(const Core& c, const BB& valid) {
        BB empty;
        for (int i = 0; i < NW; ++i) empty.w[i] = valid.w[i] & ~(c.bb[0].w[i] | c.bb[1].w[i]);
        int bs = popc(c.bb[0]), ws = popc(c.bb[1]);
        while (any(empty)) {
            BB seed = zero(); setb(seed, lowest(empty));
            const BB reg = flood(seed, empty, valid);
            const BB edge = nb(reg, valid);
            uint64_t tb = 0, tw = 0;
            for (int i = 0; i < NW; ++i) { tb |= edge.w[i] & c.bb[0].w[i]; tw |= edge.w[i] & c.bb[1].w[i]; empty.w[i] &= ~reg.w[i]; }
            if (tb && !tw) bs += popc(reg); else if (tw && !tb) ws += popc(reg);
        }
        const float b = (float)bs, w = (float)ws + KOMI;
        return b > w ? RES_WIN_P1 : (w > b ? RES_WIN_P2 : RES_DRAW);
    }
    AZ_HD static int result_core(const Core& c, const BB& valid) { return c.passes >= 2 ? score(c, valid) : RES_ONGOING; }   // go_state.cpp:315-335
    // HashEvaluator key: stones in reference pos order, then player, then ko + 1
    AZ_HD static uint64_t key_core(const Core& c) {
        uint64_t h = 1469598103934665603ULL;
        for (int a = 0; a < CELLS; ++a) {
            const int p = a2p(a);
            const uint64_t st = get(c.bb[0], p) ? 1 : (get(c.bb[1], p) ? 2 : 0);
            h = mix64(h ^ st);
        }
        h = mix64(h ^ (uint64_t)c.player);
        return mix64(h ^ (uint64_t)(int64_t)(c.ko + 1));
    }
    // liberty count of the group containing p (for planes 3/4), go_state.cpp:349-445
    AZ_HD static int group_libs(const Core& c, int p, const BB& valid, BB& group) {
        const int col = get(c.bb[0], p) ? 0 : 1;
        BB seed = zero(); setb(seed, p);
        group = flood(seed, c.bb[col], valid);
        const BB edge = nb(group, valid);
        BB lib;
        for (int i = 0; i < NW; ++i) lib.w[i] = edge.w[i] & ~(c.bb[0].w[i] | c.bb[1].w[i]);
        return popc(lib);
    }
    // all 8 planes of one cell given its group's liberty count (libs < 0: empty cell)
    AZ_HD static float feature(const Core& c, int plane, int x, int y, int libs) {
        const int p = y * PITCH + x;
        switch (plane) {
            case 0: return get(c.bb[0], p) ? 1.0f : 0.0f;
            case 1: return get(c.bb[1], p) ? 1.0f : 0.0f;
            case 2: return c.player == 1 ? 1.0f : 0.0f;
            case 3: return get(c.bb[0], p) ? fminf(1.0f, (float)libs / 10.0f) : 0.0f;
            case 4: return get(c.bb[1], p) ? fminf(1.0f, (float)libs / 10.0f) : 0.0f;
            case 5: return (c.ko >= 0 && c.ko == y * N + x) ? 1.0f : 0.0f;
            case 6: return (float)(x < N - 1 - x ? x : N - 1 - x) / (float)(N / 2);
            default: return (float)(y < N - 1 - y ? y : N - 1 - y) / (float)(N / 2);
        }
    }

    // host-side helpers used by az_engine_set_root / slot_state (single thread)
    static void init(State& s) { init_core(s.c); }
    static bool host_apply(State& s, int a) {
        const BB valid = valid_bb();
        if (a != -1 && !legal_cell(s.c, a, valid, s.hist, s.c.hist_n, nullptr, 0)) return false;
        uint64_t k;
        if (apply_core(s.c, a, valid, k) && s.c.hist_n < MAXH) s.hist[s.c.hist_n++] = k;
        return true;
    }
    static int host_root_result(const State& s) {
        const BB valid = valid_bb();
        if (s.c.passes >= 2 || s.c.ply >= MAX_GAME_MOVES) return score(s.c, valid);
        return RES_ONGOING;
    }
    static int host_ply(const State& s) { return s.c.ply; }
    static int host_player(const State& s) { return s.c.player; }

#if defined(__CUDACC__)
    // ---------------------------------------------------------------------------------------------- warp API (tree_kernels.cuh)
    struct Warp { Leaf s; const uint64_t* hist; uint64_t* hist_rw; uint8_t libs[(CELLS + 15) / 16 * 16]; };
    struct EncTarget { __nv_bfloat16* ptr; int p_total, guard, board_pitch, f16; };

    __device__ static void copy_words(void* dst, const void* src, int bytes, int lane) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = lane; i < bytes / 4; i += 32) d[i] = s[i];
    }
    __device__ static void w_init(Warp& w, int lane) { if (lane == 0) { init_core(w.s.c); w.hist = nullptr; w.hist_rw = nullptr; } __syncwarp(); }
    // state API (az_rules_replay): an unbounded move list needs a writable history of its own
    __device__ static void w_attach_history(Warp& w, uint64_t* buf, int lane) { if (lane == 0) { w.hist = buf; w.hist_rw = buf; } __syncwarp(); }
    __device__ static void w_load_root(Warp& w, const State* g, int lane) {
        copy_words(&w.s.c, &g->c, sizeof(Core), lane);
        __syncwarp();
        if (lane == 0) { w.s.c.n_extra = 0; w.hist = g->hist; w.hist_rw = nullptr; }
        __syncwarp();
    }
    // after the root advanced (k_choose_move): keys pushed since the load go to the end of the slot's history
    __device__ static void w_store_root(Warp& w, State* g, int lane) {
        __syncwarp();
        const int ne = w.s.c.n_extra, hn = w.s.c.hist_n;
        for (int i = lane; i < ne; i += 32) if (hn + i < MAXH) g->hist[hn + i] = w.s.extra[i];
        __syncwarp();
        if (lane == 0) { w.s.c.hist_n = (int16_t)min(hn + ne, MAXH); w.s.c.n_extra = 0; }
        __syncwarp();
        copy_words(&g->c, &w.s.c, sizeof(Core), lane);
    }
    __device__ static void w_store_leaf(Warp& w, Leaf* g, int lane) {
        __syncwarp();
        copy_words(g, &w.s, sizeof(Core) + 8 * w.s.c.n_extra, lane);
    }
    __device__ static void w_load_leaf(Warp& w, const Leaf* g, const State* root, int lane) {
        copy_words(&w.s.c, &g->c, sizeof(Core), lane);
        __syncwarp();
        copy_words(w.s.extra, g->extra, 8 * w.s.c.n_extra, lane);
        if (lane == 0) { w.hist = root->hist; w.hist_rw = nullptr; }
        __syncwarp();
    }
    __device__ static void w_snapshot(const Warp& w, Snapshot* out, int lane) {
        if (lane == 0) { out->bb[0] = w.s.c.bb[0]; out->bb[1] = w.s.c.bb[1]; out->ko = w.s.c.ko; out->passes = w.s.c.passes; out->ply = w.s.c.ply; out->player = w.s.c.player; out->pad_ = 0; }
    }
    // one candidate per lane; `check` = validate first (state API: the reference's makeMove throws on illegal moves)
    __device__ static bool w_apply(Warp& w, int a, int lane, bool check = false) {
        bool ok = true;
        if (check && a != -1) ok = legal_cell(w.s.c, a, valid_bb(), w.hist, w.s.c.hist_n, w.s.extra, w.s.c.n_extra);   // all lanes, same answer
        __syncwarp();
        if (ok && lane == 0) {
            uint64_t k;
            if (apply_core(w.s.c, a, valid_bb(), k)) {
                if (w.hist_rw) w.hist_rw[w.s.c.hist_n++] = k;
                else if (w.s.c.n_extra < EXTRA) w.s.extra[w.s.c.n_extra++] = k;
            }
        }
        __syncwarp();
        return ok;
    }
    __device__ static int w_result(Warp& w, int) { return result_core(w.s.c, valid_bb()); }
    // root-level result: also ends the game at the engine's move cap (area score)
    __device__ static int w_root_result(Warp& w, int) {
        if (w.s.c.passes >= 2 || w.s.c.ply >= MAX_GAME_MOVES) return score(w.s.c, valid_bb());
        return RES_ONGOING;
    }
    __device__ static int w_player(const Warp& w) { return w.s.c.player; }
    __device__ static int w_ply(const Warp& w) { return w.s.c.ply; }
    __device__ static int visit_index(int action) { return action < 0 ? CELLS : action; }
    __device__ static void record_visit(uint16_t* visits, int, int action, int n) { visits[visit_index(action)] = (uint16_t)min(n, 65535); }
    // legal moves in the reference's order (QUIRK Go2): pass (-1) first, then cells ascending.  `emit(i, a)` per move.
    template <class F>
    __device__ static int w_for_legal(Warp& w, int lane, F emit) {
        const BB valid = valid_bb();
        if (lane == 0) emit(0, -1);
        int cnt = 1;
        for (int k = 0; k < CELLS; k += 32) {
            const int a = k + lane;
            const bool ok = a < CELLS && legal_cell(w.s.c, a, valid, w.hist, w.s.c.hist_n, w.s.extra, w.s.c.n_extra);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) emit(cnt + __popc(m & ((1u << lane) - 1)), a);
            cnt += __popc(m);
        }
        __syncwarp();
        return cnt;
    }
    // children + raw priors: p_i = policy[a_i] if 0 <= a_i < len(policy) else 0 (expandNodeWithPolicy, parallel_mcts.cpp:690-711)
    __device__ static int w_enumerate(Warp& w, int lane, const int16_t*, int, int16_t* acts, float* raw, const float* pol) {
        return w_for_legal(w, lane, [&](int i, int a) { acts[i] = (int16_t)a; raw[i] = a >= 0 ? pol[a] : 0.0f; });
    }
    __device__ static int w_legal(Warp& w, int lane, int32_t* out) {
        return w_for_legal(w, lane, [&](int i, int a) { out[i] = a; });
    }
    // liberties of every stone's group into w.libs (one group at a time; every lane computes the same flood)
    __device__ static void w_group_libs(Warp& w, int lane) {
        const BB valid = valid_bb();
        BB todo;
        for (int i = 0; i < NW; ++i) todo.w[i] = w.s.c.bb[0].w[i] | w.s.c.bb[1].w[i];
        while (any(todo)) {
            BB g;
            const int n = group_libs(w.s.c, lowest(todo), valid, g);
            for (int a = lane; a < CELLS; a += 32) if (get(g, a2p(a))) w.libs[a] = (uint8_t)min(n, 255);
            for (int i = 0; i < NW; ++i) todo.w[i] &= ~g.w[i];
        }
        __syncwarp();
    }
    // feature planes straight into the conv trunk's input layout (bf16, 16 channels = 8 + 8 zero)
    __device__ static void w_encode(Warp& w, int lane, const EncTarget& enc, int slot) {
        w_group_libs(w, lane);
        const size_t row0 = (size_t)enc.guard + (size_t)slot * enc.board_pitch;
        for (int p = lane; p < N * PITCH; p += 32) {
            const int y = p / PITCH, x = p % PITCH;
            if (x >= N) continue;                             // hole column stays zero
            const int libs = w.libs[y * N + x];
            __align__(16) __nv_bfloat16 v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = net16(c < PLANES ? feature(w.s.c, c, x, y, libs) : 0.0f, enc.f16 != 0);
            *reinterpret_cast<uint4*>(enc.ptr + ((size_t)0 * enc.p_total + row0 + p) * 8) = *reinterpret_cast<const uint4*>(&v[0]);
            *reinterpret_cast<uint4*>(enc.ptr + ((size_t)1 * enc.p_total + row0 + p) * 8) = *reinterpret_cast<const uint4*>(&v[8]);
        }
    }
    __device__ static void w_planes(Warp& w, int lane, float* out) {      // fp32 [PLANES][N][N], index [plane][y][x] (state API)
        w_group_libs(w, lane);
        for (int i = lane; i < PLANES * CELLS; i += 32) { const int c = i / CELLS, a = i % CELLS; out[i] = feature(w.s.c, c, a % N, a / N, w.libs[a]); }
    }
    __device__ static uint64_t w_key(Warp& w, int) { return key_core(w.s.c); }
    // profiling (AZ_EVAL_DUP_STATS): the 8 planes are functions of stones, side and ko point — the same key as the reference's TT (go_state.cpp:773-811)
    __device__ static uint64_t w_input_key(Warp& w, int) { return key_core(w.s.c); }
    __device__ static uint64_t w_ref_tt_key(Warp& w) { return key_core(w.s.c); }
    // training examples (az_engine_make_examples)
    __device__ static void w_from_snapshot(Warp& w, const Snapshot* g, int lane) {
        if (lane == 0) {
            init_core(w.s.c); w.hist = nullptr; w.hist_rw = nullptr;
            w.s.c.bb[0] = g->bb[0]; w.s.c.bb[1] = g->bb[1]; w.s.c.ko = g->ko; w.s.c.passes = g->passes; w.s.c.ply = g->ply; w.s.c.player = g->player;
        }
        __syncwarp();
        w_group_libs(w, lane);
    }
    __device__ static float tensor_value(Warp& w, int c, int i, int j) { return feature(w.s.c, c, j, i, w.libs[i * N + j]); }     // planes are [c][y][x]
    __device__ static int policy_total(const uint16_t* visits, int lane) {
        int t = 0; for (int a = lane; a <= CELLS; a += 32) t += visits[a];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        return t;
    }
    // dense policy index: cell a, pass at N*N (the reference's action -1 has no slot in a [0, A) vector; A = N*N + 1 leaves the last free)
    template <class F> __device__ static void policy_for_each(const uint16_t* visits, int lane, F f) { for (int a = lane; a <= CELLS; a += 32) f(a, (int)visits[a]); }
#endif
};

}  // namespace az
