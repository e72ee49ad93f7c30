// gemm_tc.cu — small K-pipelined GEMM on tcgen05 for the network heads (1x1 convs and the FC layers):
//   D[m][n] = act( sum_k A[m][k] * B[n][k] + bias[n] ),  bf16 operands, fp32 accumulation in TMEM.
//
// Operand format is the conv trunk's (conv_trunk.cuh): A lives in HBM as A[k/8][rows][8] bf16, so a 128-row x
// 64-channel stage is 8 TMA bulk copies of 2 KB and already IS the K-major SWIZZLE_NONE shared-memory image
// (core matrix = 8 rows x 16 bytes; LBO = 2048 B between K chunks, SBO = 128 B between 8-row groups).  B (weights) is
// pre-arranged on the host as one contiguous 8 KB image per (n-tile of 64, K stage of 64).
// Work item = (M tile of 128 rows, N tile of 64 columns); 8-stage TMA ring; 64-column accumulators, double-buffered.
// Warp roles as in the conv kernel: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer / TMEM allocator.
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace az { namespace nn {

using namespace az::ptx;

namespace {

constexpr int G_BM = 128, G_BN = 64, G_BK = 64;
constexpr int G_A_STAGE = (G_BK / 8) * G_BM * 16;   // 16 KB
constexpr int G_B_STAGE = (G_BK / 8) * G_BN * 16;   // 8 KB per 64-column weight sub-tile
constexpr int G_THREADS = 192;
// NSUB = 64-column weight sub-tiles per work item (1, 2 or 4): the 128-row A stage is loaded once and multiplied with NSUB weight
// sub-tiles (NSUB MMAs of N = 64 per K step into adjacent TMEM columns) — the head GEMMs are bound by L2 -> SM traffic (ncu: 5.8 TB/s of
// the ~9 TB/s LTS cap with NSUB = 1), and a wider item divides the A re-reads by NSUB.  The weight image keeps its per-64-column
// layout, so a stage's B part is NSUB separate 8 KB bulk copies.
template <int NSUB> struct GCfg {
    static constexpr int STAGE = G_A_STAGE + NSUB * G_B_STAGE;                 // 24 / 32 / 48 KB
    static constexpr int NST = NSUB == 1 ? 8 : (NSUB == 2 ? 6 : 4);           // ~192 KB in flight per SM
    static constexpr int OFF_BARS = NST * STAGE;
    static constexpr int OFF_TSLOT = OFF_BARS + (2 * NST + 4) * 8;
    static constexpr int SMEM = OFF_TSLOT + 16;
    static constexpr int ACC_COLS = NSUB * G_BN;                               // accumulator columns per buffer (double-buffered)
};

template <int NSUB>
__global__ void __launch_bounds__(G_THREADS, 1) k_gemm_tc(const GemmParams p) {
    using C = GCfg<NSUB>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TSLOT);
    uint64_t* full = bars;                         // [NST] TMA → MMA
    uint64_t* empty = bars + C::NST;               // [NST] MMA → TMA
    uint64_t* acc_full = bars + 2 * C::NST;        // [2]
    uint64_t* acc_empty = bars + 2 * C::NST + 2;   // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work decomposition: `units` row groups (GEMM1: the 64 pooled cells; FC: 1), each with m_tiles tiles of 128 rows; the n_tiles
    // weight tiles of 64 columns are taken NSUB at a time
    const int m_valid = p.m_valid_dev ? *p.m_valid_dev : p.m_valid;            // valid rows inside one unit
    const int m_tiles = (m_valid + G_BM - 1) / G_BM;
    const int KS = p.k_splits > 1 ? p.k_splits : 1;          // split-K: an item covers k_stages consecutive K stages of one output tile
    const int n_groups = (p.n_tiles + NSUB - 1) / NSUB;
    const int n_items = p.units * m_tiles * n_groups * KS;
    const int k_stages_total = p.K / G_BK, k_stages = k_stages_total / KS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tslot, 2 * C::ACC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (warp == 4) {
        // TMA producer (one thread).  No integer division inside the stage loop: the A plane index advances by 8 per stage and
        // wraps at a_plane_mod (a multiple of 8), addresses are running pointers.
        if (lane == 0) {
            uint32_t it = 0;
            const size_t plane_stride = (size_t)p.a_rows * 8;                 // elements between two 8-channel K planes of A
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int split = item % KS, tile = item / KS;
                const int ng = tile % n_groups, mi = tile / n_groups;
                const int unit = mi / m_tiles, mt = mi % m_tiles;
                const int nsub = min(NSUB, p.n_tiles - ng * NSUB);
                const size_t row0 = (size_t)unit * p.unit_rows + (size_t)mt * G_BM;
                const __nv_bfloat16* a_base = p.A + row0 * 8;
                const uint8_t* b_src = reinterpret_cast<const uint8_t*>(p.B) + ((size_t)(ng * NSUB) * k_stages_total + (size_t)split * k_stages) * G_B_STAGE;
                const size_t b_tile_stride = (size_t)k_stages_total * G_B_STAGE;      // bytes between two 64-column weight tiles
                int plane = (split * k_stages * (G_BK / 8)) % p.a_plane_mod;
                for (int ks = 0; ks < k_stages; ++ks, ++it, b_src += G_B_STAGE) {
                    const uint32_t s = it % C::NST, ph = (it / C::NST) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], G_A_STAGE + nsub * G_B_STAGE);
                    uint8_t* dst = smem + s * C::STAGE;
                    const __nv_bfloat16* a_src = a_base + (size_t)plane * plane_stride;
#pragma unroll
                    for (int j = 0; j < G_BK / 8; ++j) bulk_g2s(dst + j * (G_BM * 16), a_src + (size_t)j * plane_stride, G_BM * 16, &full[s]);
#pragma unroll
                    for (int t = 0; t < NSUB; ++t) if (t < nsub) bulk_g2s(dst + G_A_STAGE + t * G_B_STAGE, b_src + (size_t)t * b_tile_stride, G_B_STAGE, &full[s]);
                    plane += G_BK / 8; if (plane >= p.a_plane_mod) plane -= p.a_plane_mod;
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        constexpr uint32_t IDESC = idesc_bf16(G_BM, G_BN);
        // converged warp, tcgen05 under elect.sync, descriptors `base + constant` (see conv_trunk.cu)
        {
            const uint64_t a_desc0 = smem_desc(smem_u32(smem), G_BM * 16, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(smem) + G_A_STAGE, G_BN * 16, 128);
            uint32_t it = 0, ait = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
                const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
                const int ng = (item / KS) % n_groups;
                const int nsub = min(NSUB, p.n_tiles - ng * NSUB);
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + as * C::ACC_COLS;
                uint32_t accumulate = 0;
                for (int ks = 0; ks < k_stages; ++ks, ++it) {
                    const uint32_t s = it % C::NST, ph = (it / C::NST) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t so = (uint64_t)(s * (C::STAGE >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < G_BK / 16; ++kk)
#pragma unroll
                            for (int t = 0; t < NSUB; ++t)
                                if (t < nsub)
                                    umma_bf16(acc + t * G_BN, a_desc0 + so + (uint64_t)(2 * kk * G_BM), b_desc0 + so + (uint64_t)(t * (G_B_STAGE >> 4) + 2 * kk * G_BN), IDESC,
                                              kk == 0 ? accumulate : 1u);
                        umma_commit(&empty[s]);
                        if (ks == k_stages - 1) umma_commit(&acc_full[as]);
                    }
                    __syncwarp();
                    accumulate = 1;
                }
            }
        }
    } else {
        uint32_t ait = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
            const int split = item % KS, tile = item / KS;
            const int ng = tile % n_groups, mi = tile / n_groups;
            const int unit = mi / m_tiles, mt = mi % m_tiles;
            const int nsub = min(NSUB, p.n_tiles - ng * NSUB);
            const int r = mt * G_BM + warp * 32 + lane;          // row inside the unit
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * C::ACC_COLS;
#pragma unroll 1
            for (int c0 = 0; c0 < nsub * G_BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (r < m_valid) {
                    if (p.mode == GEMM_OUT_FEAT) {
                        // 1x1-conv output → the FC layers' A layout: feature k' = unit*32 + channel, plane = k'/8, row = board.
                        // columns 0-31 = policy head, 32-63 = value head.  (NSUB = 1 only.)
                        __nv_bfloat16* dst = (c0 == 0) ? p.out_feat0 : p.out_feat1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 o, ol;
                            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
                            __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&ol);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float a = __uint_as_float(v[q * 8 + 2 * e]) + p.bias[c0 + q * 8 + 2 * e];
                                float b = __uint_as_float(v[q * 8 + 2 * e + 1]) + p.bias[c0 + q * 8 + 2 * e + 1];
                                if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                                ob[e] = __floats2bfloat162_rn(a, b);
                                const float2 hi = __bfloat1622float2(ob[e]);
                                lb[e] = __floats2bfloat162_rn(a - hi.x, b - hi.y);          // low half of the hi/lo split
                            }
                            *reinterpret_cast<uint4*>(dst + ((size_t)(unit * 4 + q) * p.feat_rows + r) * 8) = o;
                            *reinterpret_cast<uint4*>(dst + ((size_t)(p.feat_lo_plane + unit * 4 + q) * p.feat_rows + r) * 8) = ol;
                        }
                    } else {
                        // split-K: raw partial sums into slab `split` (bias / ReLU are applied by the consumer, k_policy_value)
                        float* dst = p.out_rows + (size_t)split * p.split_stride + (size_t)r * p.ldo;
                        const int n0 = ng * NSUB * G_BN + c0;
                        if (p.k_splits >= 1 && (p.ldo & 3) == 0 && n0 + 32 <= p.ldo) {
                            // raw partial sums, 16-byte stores (a thread owns 32 consecutive columns of its row; the row pitch is a multiple of 4 floats):
                            // 4x fewer store instructions than scalar stores — with one row per lane every store instruction is 32 separate sectors
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<uint4*>(dst + n0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = n0 + j;
                                if (n < p.n_valid) {
                                    float a = __uint_as_float(v[j]);
                                    if (p.k_splits < 1) { a += p.bias[n]; if (p.relu) a = fmaxf(a, 0.0f); }      // k_splits >= 1: raw sums, the consumer adds bias / ReLU
                                    dst[n] = a;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * C::ACC_COLS);
}

// adaptive_avg_pool2d (H x W → PH x PW, windows [floor(i*H/PH), ceil((i+1)*H/PH))) of the trunk output, written as the
// 1x1-conv GEMM's A operand: pooled[c/8][cell * boards_cap + board][8] as a bf16 hi/lo pair (planes [0,C/8) hi,
// [C/8, 2C/8) lo) so the heads see the fp32 average.
// One block = POOL_NB consecutive boards x one 8-channel chunk: the boards' rows of that chunk are one contiguous run of the
// position stream (coalesced 16-byte loads into shared memory, every trunk byte read exactly once); the outputs of one
// pooled cell for the POOL_NB boards are contiguous in `pooled` (full 128-byte lines).
constexpr int POOL_NB = 8;
__global__ void __launch_bounds__(256) k_pool(PoolParams p) {
    extern __shared__ __align__(16) uint4 srows[];                    // [POOL_NB][board_pitch + 1]: +1 row so the boards start in different banks
    const int KCH = p.channels / 8;
    const int PH = p.H < 8 ? p.H : 8, PW = p.W < 8 ? p.W : 8, cells = PH * PW;
    const int n = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int groups = (n + POOL_NB - 1) / POOL_NB;
    for (int item = blockIdx.x; item < groups * KCH; item += gridDim.x) {
        const int kc = item % KCH, b0 = (item / KCH) * POOL_NB;
        const int nb = min(POOL_NB, n - b0);
        const uint4* src = reinterpret_cast<const uint4*>(p.act) + (size_t)kc * p.p_total + (size_t)p.guard + (size_t)b0 * p.board_pitch;
        __syncthreads();                                              // previous item's readers are done
        for (int i = threadIdx.x; i < nb * p.board_pitch; i += blockDim.x) srows[i + i / p.board_pitch] = __ldg(src + i);
        __syncthreads();
        for (int o = threadIdx.x; o < POOL_NB * cells; o += blockDim.x) {
            const int bl = o % POOL_NB, cell = o / POOL_NB;
            if (bl >= nb) continue;
            const int oy = cell / PW, ox = cell % PW;
            const int y0 = (oy * p.H) / PH, y1 = ((oy + 1) * p.H + PH - 1) / PH;
            const int x0 = (ox * p.W) / PW, x1 = ((ox + 1) * p.W + PW - 1) / PW;
            const uint4* brow = srows + bl * (p.board_pitch + 1);
            float s[8] = {};
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    const uint4 u = brow[y * p.row_pitch + x];
                    const uint32_t* h = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 f = p.f16 ? unpack2_16<true>(h[e]) : unpack2_16<false>(h[e]); s[2 * e] += f.x; s[2 * e + 1] += f.y; }
                }
            const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
            uint4 ov, ol;
            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&ov);
            __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&ol);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = s[2 * e] * inv, c = s[2 * e + 1] * inv;
                ob[e] = __floats2bfloat162_rn(a, c);
                const float2 hi = __bfloat1622float2(ob[e]);
                lb[e] = __floats2bfloat162_rn(a - hi.x, c - hi.y);
            }
            const size_t row = (size_t)cell * p.boards_cap + b0 + bl;
            *reinterpret_cast<uint4*>(p.pooled + ((size_t)kc * p.pooled_rows + row) * 8) = ov;
            *reinterpret_cast<uint4*>(p.pooled + ((size_t)(KCH + kc) * p.pooled_rows + row) * 8) = ol;
        }
    }
}

}  // namespace

template <int NSUB> static int gemm_tc_launch_n(const GemmParams& p, int grid, cudaStream_t s) {
    if (cudaError_t e = smem_opt_in((const void*)k_gemm_tc<NSUB>, (int)GCfg<NSUB>::SMEM)) return (int)e;
    k_gemm_tc<NSUB><<<grid, G_THREADS, GCfg<NSUB>::SMEM, s>>>(p);
    return (int)cudaGetLastError();
}
int gemm_tc_launch(const GemmParams& p, int grid, cudaStream_t s) {
    static const int force = getenv("AZ_GEMM_NSUB") ? atoi(getenv("AZ_GEMM_NSUB")) : 0;      // profiling switch (1 = the narrow item everywhere)
    int nsub = p.mode == GEMM_OUT_FEAT ? 1 : (p.n_tiles >= 4 ? 4 : (p.n_tiles >= 2 ? 2 : 1));
    if (force == 1 || force == 2 || force == 4) nsub = p.mode == GEMM_OUT_FEAT ? 1 : std::min(force, nsub);
    if (nsub == 4) return gemm_tc_launch_n<4>(p, grid, s);
    if (nsub == 2) return gemm_tc_launch_n<2>(p, grid, s);
    return gemm_tc_launch_n<1>(p, grid, s);
}
int pool_launch(const PoolParams& p, int grid, cudaStream_t s) {
    const size_t smem = (size_t)POOL_NB * (p.board_pitch + 1) * 16;
    if (cudaError_t e = smem_opt_in((const void*)k_pool, 96 * 1024)) return (int)e;
    if (smem > 96 * 1024) return (int)cudaErrorInvalidValue;
    k_pool<<<grid, 256, smem, s>>>(p);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
