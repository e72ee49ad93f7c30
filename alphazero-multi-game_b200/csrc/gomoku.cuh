// gomoku.cuh — Gomoku rules on padded bitboards (host + device).
//
// Replaces the per-cell std::function scans of the reference (src/games/gomoku/gomoku_rules.cpp:39-115,
// src/games/gomoku/gomoku_state.cpp:477-521, :681-722) with shift-AND run detection on a bitboard whose
// row pitch is N+1: cell (x,y) — reference action a = x*N + y (gomoku_state.cpp:444-450) — sits at bit
// p = x*(N+1) + y, column N is a permanent hole, so no line can wrap across a board edge.  For 15x15 the
// board is exactly 4 x u64 and p is also the row index of the position inside the conv trunk's padded
// activation layout (conv_trunk.cu), so encode needs no index translation.
#pragma once
#include "common.cuh"
#if defined(__CUDACC__)
#include <cuda_bf16.h>
#endif

namespace az {

template <int N_>
struct Gomoku {
    static constexpr int N = N_;
    static constexpr int PITCH = N + 1;
    static constexpr int CELLS = N * N;            // action space (getActionSpaceSize)
    static constexpr int PBITS = N * PITCH;        // padded bit count
    static constexpr int NW = (PBITS + 63) / 64;   // words per colour
    static constexpr int PLANES = 11;              // getEnhancedTensorRepresentation (QUIRK G5)
    static constexpr int PACKED_NW = (CELLS + 63) / 64;
    static constexpr int ACTIONS = CELLS;          // length of the policy vector / visit-count vector
    static constexpr int MAX_CHILDREN = CELLS;
    static constexpr int SAMPLE_VISITS = CELLS;
    static constexpr bool LEGAL_POLICY = false;       // narrow policy head: all A logits are computed
    static constexpr bool TT_COARSE = false;          // the reference's TT key (stones + player) covers the hash evaluator's input
    static constexpr bool FIRST_FILL = true;       // QUIRK G2: the first enumeration of a lineage has its own order
    static constexpr int MAX_GAME_MOVES = CELLS;

    struct BB { uint64_t w[NW]; };

    struct State {
        BB bb[2];              // [0] = BLACK (player 1), [1] = WHITE (player 2)
        int16_t last[6];       // most recent move first (reference action index), -1 = none
        int16_t ply;
        int8_t player;         // 1 = BLACK to move, 2 = WHITE
        int8_t pad_;
    };

    using Leaf = State;
    using Snapshot = State;

    AZ_HD static int a2p(int a) { return (a / N) * PITCH + (a % N); }
    AZ_HD static int p2a(int p) { return (p / PITCH) * N + (p % PITCH); }

    AZ_HD static void init(State& s) {
        for (int c = 0; c < 2; ++c) for (int i = 0; i < NW; ++i) s.bb[c].w[i] = 0;
        for (int i = 0; i < 6; ++i) s.last[i] = -1;
        s.ply = 0; s.player = 1; s.pad_ = 0;
    }
    AZ_HD static bool get(const BB& b, int p) { return (b.w[p >> 6] >> (p & 63)) & 1; }
    AZ_HD static bool occupied(const State& s, int a) {
        int p = a2p(a); return get(s.bb[0], p) || get(s.bb[1], p);
    }
    // GomokuState::make_move (gomoku_state.cpp:681-722), standard rules: caller guarantees legality.
    AZ_HD static void apply(State& s, int a) {
        int p = a2p(a);
        s.bb[s.player - 1].w[p >> 6] |= 1ULL << (p & 63);
        for (int i = 5; i > 0; --i) s.last[i] = s.last[i - 1];
        s.last[0] = (int16_t)a;
        s.player = (int8_t)(3 - s.player);
        s.ply++;
    }
    AZ_HD static BB shr(const BB& b, int d) {
        BB r;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint64_t v = b.w[i] >> d;
            if (i + 1 < NW) v |= b.w[i + 1] << (64 - d);
            r.w[i] = v;
        }
        return r;
    }
    AZ_HD static BB shl(const BB& b, int d) {
        BB r;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint64_t v = b.w[i] << d;
            if (i > 0) v |= b.w[i - 1] >> (64 - d);
            r.w[i] = v;
        }
        return r;
    }
    // Is there a run of >=5 (exact == false) or a maximal run of exactly 5 (exact == true)?
    // check_line_for_five (gomoku_rules.cpp:60-95): BLACK needs length == 5, WHITE length >= 5 (QUIRK G3).
    AZ_HD static bool five(const BB& b, bool exact) {
        const int dirs[4] = {1, PITCH, PITCH + 1, PITCH - 1};
        bool hit = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int d = dirs[k];
            BB t = shr(b, d), r = b;
#pragma unroll
            for (int i = 0; i < NW; ++i) r.w[i] &= t.w[i];                 // 2 in a row starting here
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                t = shr(t, d);
#pragma unroll
                for (int i = 0; i < NW; ++i) r.w[i] &= t.w[i];             // 3, 4, 5
            }
            if (exact) {
                BB after = shr(t, d), before = shl(b, d);                    // stone at s+5d / s-d
#pragma unroll
                for (int i = 0; i < NW; ++i) r.w[i] &= ~after.w[i] & ~before.w[i];
            }
            uint64_t any = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) any |= r.w[i];
            hit = hit || (any != 0);
        }
        return hit;
    }
    // refresh_winner_cache (gomoku_state.cpp:477-489): BLACK checked first.
    AZ_HD static int winner(const State& s) {
        if (five(s.bb[0], true)) return 1;
        if (five(s.bb[1], false)) return 2;
        return 0;
    }
    AZ_HD static int stones(const State& s) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
#if defined(__CUDA_ARCH__)
            t += __popcll(s.bb[0].w[i]) + __popcll(s.bb[1].w[i]);
#else
            t += __builtin_popcountll(s.bb[0].w[i]) + __builtin_popcountll(s.bb[1].w[i]);
#endif
        }
        return t;
    }
    // getGameResult / is_terminal (gomoku_state.cpp:189-201, 491-521): winner, else draw iff board full.
    AZ_HD static int result(const State& s) {
        int w = winner(s);
        if (w == 1) return RES_WIN_P1;
        if (w == 2) return RES_WIN_P2;
        return stones(s) >= CELLS ? RES_DRAW : RES_ONGOING;
    }
    // HashEvaluator key over the reference's *packed* words (bit a = x*N + y).
    AZ_HD static uint64_t key(const State& s) {
        uint64_t h = 1469598103934665603ULL;
        for (int c = 0; c < 2; ++c) {
            uint64_t packed[PACKED_NW];
            for (int i = 0; i < PACKED_NW; ++i) packed[i] = 0;
            for (int x = 0; x < N; ++x) {
                int p0 = x * PITCH;   // N bits starting at p0
                uint64_t row = s.bb[c].w[p0 >> 6] >> (p0 & 63);
                if ((p0 & 63) + N > 64) row |= s.bb[c].w[(p0 >> 6) + 1] << (64 - (p0 & 63));
                row &= (1ULL << N) - 1;
                int a0 = x * N;
                packed[a0 >> 6] |= row << (a0 & 63);
                if ((a0 & 63) + N > 64) packed[(a0 >> 6) + 1] |= row >> (64 - (a0 & 63));
            }
            for (int i = 0; i < PACKED_NW; ++i) h = mix64(h ^ packed[i]);
        }
        return mix64(h ^ (uint64_t)s.player);
    }
    // One input feature value, plane c (0..10) at cell (x,y): getEnhancedTensorRepresentation
    // (gomoku_state.cpp:207-258) + to_tensor (:811-840) + get_previous_moves (:852-869, QUIRK G5: the most
    // recent move is attributed to the side TO MOVE, so "black history" holds the other colour's stones).
    AZ_HD static float feature(const State& s, int c, int x, int y) {
        int p = x * PITCH + y, a = x * N + y;
        int me = s.player - 1;
        switch (c) {
            case 0: return get(s.bb[me], p) ? 1.0f : 0.0f;
            case 1: return get(s.bb[1 - me], p) ? 1.0f : 0.0f;
            case 2: return s.player == 1 ? 1.0f : 0.0f;
            case 9: return (float)x / (float)(N - 1);
            case 10: return (float)y / (float)(N - 1);
            default: {
                // planes 3-5: get_previous_moves(BLACK,3); 6-8: get_previous_moves(WHITE,3)
                int colour = c < 6 ? 1 : 2, k = c < 6 ? c - 3 : c - 6;
                // moves attributed to `player` are last[0], last[2], last[4]; to the other last[1], [3], [5]
                int idx = (colour == s.player) ? 2 * k : 2 * k + 1;
                return s.last[idx] == a ? 1.0f : 0.0f;
            }
        }
    }

    // host-side helpers used by az_engine_set_root / slot_state (single thread)
    static bool host_apply(State& s, int a) { if (a < 0 || a >= CELLS || occupied(s, a)) return false; apply(s, a); return true; }   // IllegalMove, gomoku_state.cpp:681-689
    static int host_root_result(const State& s) { return result(s); }
    static int host_ply(const State& s) { return s.ply; }
    static int host_player(const State& s) { return s.player; }

#if defined(__CUDACC__)
    // ---------------------------------------------------------------------------------------------
    // Warp API used by the tree kernels (tree_kernels.cuh): the game state of the tree a warp owns lives in that
    // warp's shared-memory workspace; mutating calls are made by the whole warp (lane 0 writes, __syncwarp after).
    struct Warp { State s; };
    struct EncTarget { __nv_bfloat16* ptr; int p_total, guard, board_pitch, f16; };

    __device__ static void w_load(Warp& w, const State* g, int lane) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&w.s);
        for (int i = lane; i < (int)(sizeof(State) / 4); i += 32) dst[i] = src[i];
        __syncwarp();
    }
    __device__ static void w_store(const Warp& w, State* g, int lane) {
        __syncwarp();
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&w.s);
        uint32_t* dst = reinterpret_cast<uint32_t*>(g);
        for (int i = lane; i < (int)(sizeof(State) / 4); i += 32) dst[i] = src[i];
    }
    // root / leaf states are the same 88-byte record for Gomoku (Go keeps its superko history out of the leaf, go.cuh)
    __device__ static void w_load_root(Warp& w, const State* g, int lane) { w_load(w, g, lane); }
    __device__ static void w_store_root(Warp& w, State* g, int lane) { w_store(w, g, lane); }
    __device__ static void w_store_leaf(Warp& w, Leaf* g, int lane) { w_store(w, g, lane); }
    __device__ static void w_load_leaf(Warp& w, const Leaf* g, const State*, int lane) { w_load(w, g, lane); }
    __device__ static void w_snapshot(const Warp& w, Snapshot* out, int lane) { w_store(w, out, lane); }
    __device__ static int w_root_result(Warp& w, int) { return result(w.s); }
    __device__ static void w_init(Warp& w, int lane) { if (lane == 0) init(w.s); __syncwarp(); }
    __device__ static void w_attach_history(Warp&, uint64_t*, int) {}
    // returns false if the reference's makeMove would throw (gomoku_state.cpp:681-689)
    __device__ static bool w_apply(Warp& w, int a, int lane, bool = false) {
        const bool ok = a >= 0 && a < CELLS && !occupied(w.s, a);
        __syncwarp();
        if (ok && lane == 0) apply(w.s, a);
        __syncwarp();
        return ok;
    }
    __device__ static int w_result(Warp& w, int) { return result(w.s); }
    __device__ static int w_player(const Warp& w) { return w.s.player; }
    __device__ static int w_ply(const Warp& w) { return w.s.ply; }
    __device__ static int visit_index(int action) { return action; }
    __device__ static void record_visit(uint16_t* visits, int, int action, int n) { visits[visit_index(action)] = (uint16_t)min(n, 65535); }
    // Legal moves in the reference's order (QUIRK G2): descending action index — std::unordered_set iteration order
    // after any refill — except the first enumeration of a lineage (`root_order`, computed on the host with the real
    // container).  Writes acts[i] and raw[i] = policy[action] (expandNodeWithPolicy, parallel_mcts.cpp:705-711).
    __device__ static int w_enumerate(Warp& w, int lane, const int16_t* root_order, int root_order_n, int16_t* acts, float* raw, const float* pol) {
        if (root_order != nullptr) {
            for (int i = lane; i < root_order_n; i += 32) { const int a = root_order[i]; acts[i] = (int16_t)a; raw[i] = pol[a]; }
            __syncwarp();
            return root_order_n;
        }
        int cnt = 0;
        for (int k = 0; k < CELLS; k += 32) {
            const int a = CELLS - 1 - (k + lane);
            const bool empty = a >= 0 && !occupied(w.s, a);
            const unsigned m = __ballot_sync(0xffffffffu, empty);
            if (empty) { const int i = cnt + __popc(m & ((1u << lane) - 1)); acts[i] = (int16_t)a; raw[i] = pol[a]; }
            cnt += __popc(m);
        }
        __syncwarp();
        return cnt;
    }
    // legal moves for the state API (az_rules_replay): the order of every enumeration after the first
    __device__ static int w_legal(Warp& w, int lane, int32_t* out) {
        int cnt = 0;
        for (int k = 0; k < CELLS; k += 32) {
            const int a = CELLS - 1 - (k + lane);
            const bool empty = a >= 0 && !occupied(w.s, a);
            const unsigned m = __ballot_sync(0xffffffffu, empty);
            if (empty) out[cnt + __popc(m & ((1u << lane) - 1))] = a;
            cnt += __popc(m);
        }
        return cnt;
    }
    // feature planes straight into the conv trunk's input layout (bf16, 16 channels = 11 + 5 zero)
    __device__ static void w_encode(Warp& w, int lane, const EncTarget& enc, int slot) {
        const size_t row0 = (size_t)enc.guard + (size_t)slot * enc.board_pitch;
        for (int p = lane; p < N * PITCH; p += 32) {
            const int x = p / PITCH, y = p % PITCH;
            if (y >= N) continue;                             // hole column stays zero
            __align__(16) __nv_bfloat16 v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = net16(c < PLANES ? feature(w.s, c, x, y) : 0.0f, enc.f16 != 0);
            *reinterpret_cast<uint4*>(enc.ptr + ((size_t)0 * enc.p_total + row0 + p) * 8) = *reinterpret_cast<const uint4*>(&v[0]);
            *reinterpret_cast<uint4*>(enc.ptr + ((size_t)1 * enc.p_total + row0 + p) * 8) = *reinterpret_cast<const uint4*>(&v[8]);
        }
    }
    __device__ static void w_planes(Warp& w, int lane, float* out) {      // fp32 [PLANES][N][N] (state API)
        for (int i = lane; i < PLANES * CELLS; i += 32) { const int c = i / CELLS, a = i % CELLS; out[i] = feature(w.s, c, a / N, a % N); }
    }
    __device__ static uint64_t w_key(Warp& w, int) { return key(w.s); }
    // profiling (AZ_EVAL_DUP_STATS): key over the whole network input (stones, side, the six history moves of planes 3-8) and over what the
    // reference's TranspositionTable distinguishes (stones + side, gomoku_state.cpp:620-656)
    __device__ static uint64_t w_input_key(Warp& w, int) { uint64_t h = key(w.s); for (int i = 0; i < 6; ++i) h = mix64(h ^ (uint64_t)(uint16_t)w.s.last[i]); return h; }
    __device__ static uint64_t w_ref_tt_key(Warp& w) { return key(w.s); }
    // training examples (az_engine_make_examples): state from a sample's snapshot, plane value at tensor index [c][i][j], dense policy
    __device__ static void w_from_snapshot(Warp& w, const Snapshot* g, int lane) { w_load(w, g, lane); }
    __device__ static float tensor_value(Warp& w, int c, int i, int j) { return feature(w.s, c, i, j); }
    __device__ static int policy_total(const uint16_t* visits, int lane) {
        int t = 0; for (int a = lane; a < CELLS; a += 32) t += visits[a];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        return t;
    }
    template <class F> __device__ static void policy_for_each(const uint16_t* visits, int lane, F f) { for (int a = lane; a < CELLS; a += 32) f(a, (int)visits[a]); }
#endif
};

}  // namespace az
