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

    struct BB { uint64_t w[NW]; };

    struct State {
        BB bb[2];              // [0] = BLACK (player 1), [1] = WHITE (player 2)
        int16_t last[6];       // most recent move first (reference action index), -1 = none
        int16_t ply;
        int8_t player;         // 1 = BLACK to move, 2 = WHITE
        int8_t pad_;
    };

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
};

}  // namespace az
