// conv_trunk.cu — see conv_trunk.cuh for the design.  sm_100a only (tcgen05 / TMEM / TMA bulk copy).
#include "conv_trunk.cuh"
#include "ptx.cuh"

namespace az { namespace nn {

using namespace az::ptx;

namespace {

template <int CIN>
struct Cfg {
    static constexpr int COUT = CONV_COUT;
    static constexpr int KCH = CIN / 8;                        // channel chunks (planes) of the input
    static constexpr int WK = CIN < 64 ? CIN : 64;             // K channels per weight stage
    static constexpr int STAGES_PER_TAP = CIN / WK;
    static constexpr int KSTEPS = WK / 16;                     // UMMA_K = 16 steps per weight stage
    static constexpr int ROWS = CONV_BM + 2 * CONV_HALO;       // rows of one A tile incl. halo
    static constexpr int PLANE = ROWS * 16;                    // bytes per channel-chunk plane in smem
    static constexpr int A_STAGE = KCH * PLANE;
    static constexpr int NAS = 2;                              // A stages
    static constexpr int WPLANE = COUT * 16;                   // bytes per 8-channel K chunk of the weights
    static constexpr int W_STAGE = (WK / 8) * WPLANE;
    static constexpr int NWS = 4;                              // weight stages in the ring
    static constexpr int OFF_W = NAS * A_STAGE;
    static constexpr int OFF_BIAS = OFF_W + NWS * W_STAGE;
    static constexpr int OFF_BARS = OFF_BIAS + COUT * 4;
    static constexpr int OFF_TSLOT = OFF_BARS + 16 * 8;
    static constexpr int SMEM = OFF_TSLOT + 16;
};

template <int CIN>
__global__ void __launch_bounds__(CONV_THREADS, 1) k_conv3x3(const ConvParams p) {
    using C = Cfg<CIN>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + C::OFF_W;
    float* sBias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TSLOT);
    uint64_t* a_full = bars;          // [2] TMA → MMA
    uint64_t* a_empty = bars + 2;     // [2] MMA → TMA
    uint64_t* w_full = bars + 4;      // [4]
    uint64_t* w_empty = bars + 8;     // [4]
    uint64_t* acc_full = bars + 12;   // [2] MMA → epilogue
    uint64_t* acc_empty = bars + 14;  // [2] epilogue → MMA

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rows = p.n_boards_dev ? (*p.n_boards_dev) * p.board_pitch : p.n_rows;
    const int n_items = (n_rows + CONV_BM - 1) / CONV_BM;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        for (int i = 0; i < C::NWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < C::COUT; i += CONV_THREADS) sBias[i] = p.bias[i];
    if (warp == 5) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (warp == 4) {
        // ===================== TMA producer =====================
        uint32_t wit = 0, ait = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
            if ((p.dbg & 8) && ait >= 2) { if (!(p.dbg & 4)) goto weights; continue; }
            mbar_wait(&a_empty[as], aph ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&a_full[as], C::A_STAGE);
                const size_t row0 = (size_t)CONV_GUARD + (size_t)item * CONV_BM - CONV_HALO;
                for (int kc = 0; kc < C::KCH; ++kc)
                    bulk_g2s(sA + as * C::A_STAGE + kc * C::PLANE, p.in + ((size_t)kc * p.p_total + row0) * 8, C::PLANE, &a_full[as]);
            }
        weights:
            for (int st = 0; st < 9 * C::STAGES_PER_TAP; ++st, ++wit) {
                const uint32_t ws = wit % C::NWS, wph = (wit / C::NWS) & 1;
                if ((p.dbg & 4) && wit >= C::NWS) continue;
                mbar_wait(&w_empty[ws], wph ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&w_full[ws], C::W_STAGE);
                    bulk_g2s(sW + ws * C::W_STAGE, reinterpret_cast<const uint8_t*>(p.w) + (size_t)st * C::W_STAGE, C::W_STAGE, &w_full[ws]);
                }
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC = idesc_bf16(128, C::COUT);
        const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
        uint32_t wit = 0, ait = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
            if (!((p.dbg & 8) && ait >= 2)) mbar_wait(&a_full[as], ph);
            mbar_wait(&acc_empty[as], ph ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + as * 256;
            for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * p.row_pitch + (tap % 3 - 1);
                for (int h = 0; h < C::STAGES_PER_TAP; ++h, ++wit) {
                    const uint32_t ws = wit % C::NWS, wph = (wit / C::NWS) & 1;
                    if (!((p.dbg & 4) && wit >= C::NWS)) mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    if (lane == 0) {
#pragma unroll
                        for (int it = 0; it < 2 * C::KSTEPS; ++it) {
                            {
                                const int mt = (p.dbg & 1) ? (it & 1) : (it / C::KSTEPS);
                                const int kk = (p.dbg & 1) ? (it >> 1) : (it % C::KSTEPS);
                                const int kc = h * (C::WK / 8) + 2 * kk;
                                const uint32_t a_addr = sA_u + as * C::A_STAGE + kc * C::PLANE + (CONV_HALO + mt * 128 + shift) * 16;
                                const uint32_t b_addr = sW_u + ws * C::W_STAGE + (2 * kk) * C::WPLANE;
                                umma_bf16(acc + mt * 128, smem_desc(a_addr, C::PLANE, 128), smem_desc(b_addr, C::WPLANE, 128), IDESC,
                                          (tap | h | kk) != 0 ? 1u : 0u);
                            }
                        }
                        if (!(p.dbg & 4)) umma_commit(&w_empty[ws]);     // weight stage reusable once these MMAs retire
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) { if (!(p.dbg & 8)) umma_commit(&a_empty[as]); umma_commit(&acc_full[as]); }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 0-3) =====================
        uint32_t ait = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
            mbar_wait(&acc_full[as], ph);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                const int row = item * CONV_BM + mt * 128 + warp * 32 + lane;
                const size_t grow = (size_t)CONV_GUARD + row;
                const bool valid = (row < n_rows) && (p.rowvalid[grow] != 0);
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 256 + mt * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < C::COUT; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    uint4 res[4];
                    if (p.dbg & 2) { tmem_ld_wait(); continue; }
                    if (p.resid != nullptr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            res[q] = *reinterpret_cast<const uint4*>(p.resid + ((size_t)(c0 / 8 + q) * p.p_total + grow) * 8);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[q * 8 + e]) + sBias[c0 + q * 8 + e];
                        if (p.resid != nullptr) {
                            const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&res[q]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(rb[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
                        }
                        uint4 o;
                        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float a = v[2 * e], b = v[2 * e + 1];
                            if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                            if (!valid) { a = 0.0f; b = 0.0f; }
                            ob[e] = __floats2bfloat162_rn(a, b);
                        }
                        *reinterpret_cast<uint4*>(p.out + ((size_t)(c0 / 8 + q) * p.p_total + grow) * 8) = o;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t conv_smem_bytes(int cin) { return cin == 16 ? Cfg<16>::SMEM : Cfg<128>::SMEM; }

size_t conv_weight_elems(int cin_total) { return (size_t)9 * cin_total * CONV_COUT; }

// Weight image = the exact shared-memory picture of each weight stage, stages in the order the kernel
// consumes them: [tap][K stage h][8-channel chunk j][cout n][8 channels e].
size_t conv_weight_index(int cin_total, int tap, int ci, int co) {
    const int wk = cin_total < 64 ? cin_total : 64;
    const int spt = cin_total / wk;
    const int h = ci / wk, j = (ci % wk) / 8, e = ci % 8;
    return ((((size_t)tap * spt + h) * (wk / 8) + j) * CONV_COUT + co) * 8 + e;
}

int conv3x3_launch(const ConvParams& p, int cin, int grid, cudaStream_t stream) {
    static bool attr_done[2] = {false, false};
    cudaError_t err;
    if (cin == 16) {
        if (!attr_done[0]) { err = cudaFuncSetAttribute(k_conv3x3<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<16>::SMEM); if (err) return (int)err; attr_done[0] = true; }
        k_conv3x3<16><<<grid, CONV_THREADS, Cfg<16>::SMEM, stream>>>(p);
    } else if (cin == 128) {
        if (!attr_done[1]) { err = cudaFuncSetAttribute(k_conv3x3<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::SMEM); if (err) return (int)err; attr_done[1] = true; }
        k_conv3x3<128><<<grid, CONV_THREADS, Cfg<128>::SMEM, stream>>>(p);
    } else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
