// head_conv.cu — the heads' two 1x1 convolutions + adaptive average pooling + BatchNorm shift + ReLU in ONE pass over the trunk output
// (reference head shape: python/alphazero/models/ddw_randwire.py:203-235, evaluated through torch::jit in
// src/nn/torch_neural_network.cpp:224-363).
//
// The network computes  relu(bn(conv1x1(avgpool(x)))).  Pooling and a 1x1 convolution are both linear and the pooling weights sum to
// one, so  bn(conv1x1(avgpool(x))) = avgpool(scale * conv1x1(x)) + shift : the 1x1 convs can run on the FULL-resolution trunk output
// and the pooling can happen on their 64 output channels afterwards.  That is what this kernel does, because the trunk output already
// IS a tcgen05 A operand (act[c/8][row][8]: every 8 rows of a channel chunk are one SWIZZLE_NONE core matrix, conv_trunk.cuh):
//   * work item = one board = up to 256 consecutive rows of the position stream; 16 TMA bulk copies (one per channel chunk) bring its
//     128 channels into shared memory, double-buffered;
//   * 2 M tiles x 8 K steps x {W_hi, W_lo} tcgen05.mma (128 x 64 x 16): the folded weights are split into a bf16 hi/lo pair so that the
//     heads see fp32-accurate weights, the activations are exact (they are bf16 already) — no rounding of the heads' own;
//   * the 256 epilogue threads move the fp32 [rows][64] result from TMEM to shared memory (row-major, 16-byte chunks XOR-swizzled by the
//     row so that both the row-per-lane writes and the window reads spread over the banks), pool it with 16-byte reads
//     (adaptive windows [floor(i H / 8), ceil((i+1) H / 8)) ), add the BatchNorm shift, ReLU, and write the bf16 hi/lo feature pairs
//     straight into the FC GEMMs' operand layout (feature = cell * 32 + channel).
// It replaces k_pool + the 1x1 GEMM (gemm_tc.cu): the trunk output is read once and the 134 MB pooled intermediate (write + read)
// disappears.  Built for 128-channel trunks and boards of at most 256 padded rows (Gomoku 15, Go 9 / 13, chess); everything else keeps
// the two-kernel path.
#include "head_conv.cuh"
#include "ptx.cuh"

namespace az { namespace nn {

using namespace az::ptx;

namespace {

constexpr int HC_THREADS = 320;                        // warps 0-7 epilogue / pooling, warp 8 TMA producer, warp 9 MMA issuer
constexpr int HC_ROWS = 256;                          // rows of an A stage (2 M tiles)
constexpr int HC_PLANE = HC_ROWS * 16;                // 4 KB per 8-channel chunk
constexpr int HC_KCH = 16;                            // 128 channels
constexpr int HC_A_STAGE = HC_KCH * HC_PLANE;         // 64 KB
constexpr int HC_WPLANE = 64 * 16;                    // 1 KB: 64 output channels x 8 input channels
constexpr int HC_W_HALF = HC_KCH * HC_WPLANE;         // 16 KB (hi or lo image)
constexpr int HC_OFF_W = 2 * HC_A_STAGE;
constexpr int HC_OFF_T = HC_OFF_W + 2 * HC_W_HALF;
constexpr int HC_OFF_BIAS = HC_OFF_T + HC_ROWS * 64 * 4;      // pooling tile: [256 rows][64 channels] fp32, 64 KB
constexpr int HC_OFF_BARS = HC_OFF_BIAS + 64 * 4;
constexpr int HC_OFF_TSLOT = HC_OFF_BARS + 16 * 8;
constexpr int HC_SMEM = HC_OFF_TSLOT + 16;            // 229,776 B

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // the 8 epilogue warps only
// pooling tile address (in floats) of 16-byte chunk c4 (4 channels) of a row: chunks XOR-swizzled by the row
__device__ __forceinline__ int tile_off(int row, int c4) { return row * 64 + ((c4 ^ (row & 15)) << 2); }

__global__ void __launch_bounds__(HC_THREADS, 1) k_head_conv_pool(const HeadConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + HC_OFF_W;
    float* sT = reinterpret_cast<float*>(smem + HC_OFF_T);
    float* sBias = reinterpret_cast<float*>(smem + HC_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HC_OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + HC_OFF_TSLOT);
    uint64_t* a_full = bars;            // [2] TMA → MMA
    uint64_t* a_empty = bars + 2;       // [2] MMA → TMA
    uint64_t* acc_full = bars + 4;      // [2] MMA → epilogue
    uint64_t* acc_empty = bars + 6;     // [2] epilogue → MMA
    uint64_t* w_full = bars + 8;        // [1]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_boards = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int m_tiles = (p.board_pitch + 127) / 128;              // 1 or 2

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 64) sBias[threadIdx.x] = p.bias[threadIdx.x];
    if (warp == 9) tmem_alloc(tslot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, 2 * HC_W_HALF);
            bulk_g2s(sW, p.w, 2 * HC_W_HALF, w_full);
            uint32_t ait = 0;
            for (int b = blockIdx.x; b < n_boards; b += gridDim.x, ++ait) {
                const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                const int bb = p.reverse ? n_boards - 1 - b : b;
                const size_t row0 = (size_t)p.guard + (size_t)bb * p.board_pitch;
                // never read past the end of a channel plane (the last boards of the stream): rows past board_pitch are never pooled
                const int rows = (int)min((size_t)(m_tiles * 128), (size_t)p.p_total - row0);
                mbar_wait(&a_empty[as], ph ^ 1);
                mbar_arrive_expect_tx(&a_full[as], (uint32_t)(HC_KCH * rows * 16));
                for (int kc = 0; kc < HC_KCH; ++kc)
                    bulk_g2s(sA + as * HC_A_STAGE + kc * HC_PLANE, p.act + ((size_t)kc * p.p_total + row0) * 8, (uint32_t)(rows * 16), &a_full[as]);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===================== MMA issuer (converged warp, tcgen05 under elect.sync) =====================
        const uint32_t IDESC = idesc_16(128, 64, p.f16 != 0);
        const uint64_t a_desc0 = smem_desc(smem_u32(sA), HC_PLANE, 128);
        const uint64_t b_desc0 = smem_desc(smem_u32(sW), HC_WPLANE, 128);
        mbar_wait(w_full, 0);
        uint32_t ait = 0;
        for (int b = blockIdx.x; b < n_boards; b += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
            mbar_wait(&a_full[as], ph);
            mbar_wait(&acc_empty[as], ph ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + as * 128;
            const uint64_t a_st = a_desc0 + (uint64_t)(as * (HC_A_STAGE >> 4));
            if (elect_one()) {
                for (int mt = 0; mt < m_tiles; ++mt) {
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int kk = 0; kk < HC_KCH / 2; ++kk)
                            umma_bf16(acc + mt * 64, a_st + (uint64_t)(mt * 128 + 2 * kk * (HC_PLANE >> 4)),
                                      b_desc0 + (uint64_t)(s * (HC_W_HALF >> 4) + 2 * kk * (HC_WPLANE >> 4)), IDESC, (s == 0 && kk == 0) ? 0u : 1u);
                }
                umma_commit(&a_empty[as]);
                umma_commit(&acc_full[as]);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 0-7): TMEM → pooling tile → pooled features =====================
        const int t = threadIdx.x;                           // 0..255
        const int ox = t & 7, cg = (t >> 3) & 7, oyq = t >> 6;
        const int x0 = (ox * p.W) / 8, x1 = ((ox + 1) * p.W + 7) / 8;
        const int mt = warp >> 2, wq = warp & 3;             // M tile this warp copies out of TMEM, TMEM lane quarter
        __nv_bfloat16* dst = cg < 4 ? p.featP : p.featV;
        const int q = cg & 3;
        float bias[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bias[e] = sBias[cg * 8 + e];
        uint32_t ait = 0;
        for (int b = blockIdx.x; b < n_boards; b += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
            mbar_wait(&acc_full[as], ph);
            tc_fence_after();
            if (mt < m_tiles) {
                const int row = mt * 128 + wq * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * 128 + mt * 64;
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
                tmem_ld32(taddr + 32, vb);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    *reinterpret_cast<uint4*>(sT + tile_off(row, j)) = make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]);
                    *reinterpret_cast<uint4*>(sT + tile_off(row, 8 + j)) = make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);                     // accumulators drained: the issuer may start the board after next
            epi_bar();                                        // the tile is complete
#pragma unroll 1
            for (int oy = oyq; oy < 8; oy += 4) {
                const int y0 = (oy * p.H) / 8, y1 = ((oy + 1) * p.H + 7) / 8;
                float s[8] = {};
                // adaptive windows of boards 8..16 wide are at most 3 x 3 cells (launch check): a fixed, fully unrolled 3 x 3 walk with the cells
                // outside the window predicated off — all shared-memory loads of a window in flight together, same summation order (y, then x)
                float4 u[9][2];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {         // loads first, unconditionally (cells outside the window: clamped to the board, discarded below)
                        const int row = min(y0 + dy, p.H - 1) * p.row_pitch + min(x0 + dx, p.W - 1);
                        u[dy * 3 + dx][0] = *reinterpret_cast<const float4*>(sT + tile_off(row, 2 * cg));
                        u[dy * 3 + dx][1] = *reinterpret_cast<const float4*>(sT + tile_off(row, 2 * cg + 1));
                    }
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (y0 + dy < y1 && x0 + dx < x1) {
                            const float4 u0 = u[dy * 3 + dx][0], u1 = u[dy * 3 + dx][1];
                            s[0] += u0.x; s[1] += u0.y; s[2] += u0.z; s[3] += u0.w; s[4] += u1.x; s[5] += u1.y; s[6] += u1.z; s[7] += u1.w;
                        }
                    }
                const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
                uint4 ov, ol;
                __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&ov);
                __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&ol);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a = fmaxf(s[2 * e] * inv + bias[2 * e], 0.0f), c = fmaxf(s[2 * e + 1] * inv + bias[2 * e + 1], 0.0f);
                    ob[e] = __floats2bfloat162_rn(a, c);
                    const float2 hi = __bfloat1622float2(ob[e]);
                    lb[e] = __floats2bfloat162_rn(a - hi.x, c - hi.y);          // low half of the hi/lo split
                }
                const int cell = oy * 8 + ox;
                const int bb = p.reverse ? n_boards - 1 - b : b;
                *reinterpret_cast<uint4*>(dst + ((size_t)(cell * 4 + q) * p.feat_rows + bb) * 8) = ov;
                *reinterpret_cast<uint4*>(dst + ((size_t)(p.feat_lo_plane + cell * 4 + q) * p.feat_rows + bb) * 8) = ol;
            }
            epi_bar();                                        // everyone is done reading the tile before the next board overwrites it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 256);
}

}  // namespace

// (H, W <= 16: every adaptive pooling window [floor(i H / 8), ceil((i + 1) H / 8)) is at most 3 cells wide — the epilogue's fixed 3 x 3 walk)
bool head_conv_supported(int channels, int board_pitch, int H, int W) { return channels == 128 && board_pitch <= HC_ROWS && H >= 8 && W >= 8 && H <= 16 && W <= 16; }

int head_conv_launch(const HeadConvParams& p, int grid, cudaStream_t s) {
    if (cudaError_t e = smem_opt_in((const void*)k_head_conv_pool, (int)HC_SMEM)) return (int)e;
    k_head_conv_pool<<<grid, HC_THREADS, HC_SMEM, s>>>(p);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
