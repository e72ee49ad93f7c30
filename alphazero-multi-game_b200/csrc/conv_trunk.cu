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
    // weights: a ring of 4 stages streamed per item — or, when all nine taps fit (the stems: 36 KB at 16 input channels, 72 KB at 32), loaded
    // once and kept: re-streaming 9 stages through a 4-deep ring for every item (18 MMAs) made the stem wait on TMA round trips, not on stores
    static constexpr bool RESIDENT = CIN < 64;
    static constexpr int NWS = RESIDENT ? 9 * STAGES_PER_TAP : 4;
    static constexpr int OFF_W = NAS * A_STAGE;
    static constexpr int OFF_BIAS = OFF_W + NWS * W_STAGE;
    static constexpr int OFF_BARS = OFF_BIAS + COUT * 4;
    static constexpr int OFF_TSLOT = OFF_BARS + (8 + 2 * NWS) * 8;
    static constexpr int SMEM = OFF_TSLOT + 16;
};

// one 32-channel chunk of the epilogue: + bias (+ residual) → ReLU → pad-row mask → bf16 / fp16 → four 16-byte stores.  The epilogue warps
// share the SM's issue slots with the MMA issuer, so the per-pair work is kept to: 2 FADD (bias), 1-2 unpack + 2 FADD (residual), ONE
// F2FP conversion that also does the ReLU and the fp16 saturation, one select for the pad-row mask.
template <bool F16, bool RELU>
__device__ __forceinline__ void epi_chunk_t(const uint32_t* r, const uint4* res, bool has_res, const float* sBias, int c0, bool valid,
                                            __nv_bfloat16* out, size_t p_total, size_t grow, bool no_store, bool skip_store, uint64_t pol = 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0 + q * 8), b1 = *reinterpret_cast<const float4*>(sBias + c0 + q * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[q * 8 + e]) + bb[e];
        if (has_res) {
            const uint32_t* rb = reinterpret_cast<const uint32_t*>(&res[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = unpack2_16<F16>(rb[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
        }
        uint4 o;
        uint32_t* ob = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const uint32_t pk = pack2_16<F16, RELU>(v[2 * e], v[2 * e + 1]); ob[e] = valid ? pk : 0u; }
        if (no_store && o.x != 0x7fc17fc1u) continue;       // profiling experiment (dbg & 32): keep the math, drop the store
        if (skip_store) continue;                           // k_trunk_pair: a row past the end of this pair's board group belongs to another pair
        __nv_bfloat16* dst = out + ((size_t)(c0 / 8 + q) * p_total + grow) * 8;
        if (pol) asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w), "l"(pol) : "memory");
        else *reinterpret_cast<uint4*>(dst) = o;
    }
}
__device__ __forceinline__ void pair_epi_chunk(bool f16, const uint32_t* r, const uint4* res, bool has_res, const float* sBias, int c0, bool relu, bool valid,
                                               __nv_bfloat16* out, size_t p_total, size_t grow, bool no_store = false, bool skip_store = false, uint64_t pol = 0) {
    if (f16) { if (relu) epi_chunk_t<true, true>(r, res, has_res, sBias, c0, valid, out, p_total, grow, no_store, skip_store, pol);
               else epi_chunk_t<true, false>(r, res, has_res, sBias, c0, valid, out, p_total, grow, no_store, skip_store, pol); }
    else { if (relu) epi_chunk_t<false, true>(r, res, has_res, sBias, c0, valid, out, p_total, grow, no_store, skip_store, pol);
           else epi_chunk_t<false, false>(r, res, has_res, sBias, c0, valid, out, p_total, grow, no_store, skip_store, pol); }
}

// 8 epilogue warps (warps 0-3: rows 0-127 of the item, warps 4-7: rows 128-255; a warp reads TMEM lanes 32 (w % 4) ...), warp 8 = TMA
// producer, warp 9 = MMA issuer.  With 4 epilogue warps the stem (K = 16 per tap: 18 MMAs per item) was bound by the epilogue
// draining 256 x 128 outputs per item.
constexpr int CONV1_THREADS = 320;
template <int CIN>
__global__ void __launch_bounds__(CONV1_THREADS, 1) k_conv3x3(const ConvParams p) {
    using C = Cfg<CIN>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + C::OFF_W;
    float* sBias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TSLOT);
    uint64_t* a_full = bars;          // [2] TMA → MMA
    uint64_t* a_empty = bars + 2;     // [2] MMA → TMA
    uint64_t* acc_full = bars + 4;    // [2] MMA → epilogue
    uint64_t* acc_empty = bars + 6;   // [2] epilogue → MMA
    uint64_t* w_full = bars + 8;      // [NWS]
    uint64_t* w_empty = bars + 8 + C::NWS;   // [NWS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rows = p.n_boards_dev ? (*p.n_boards_dev) * p.board_pitch : p.n_rows;
    const int n_items = (n_rows + CONV_BM - 1) / CONV_BM;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
        for (int i = 0; i < C::NWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < C::COUT; i += CONV1_THREADS) sBias[i] = p.bias[i];
    if (warp == 9) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (warp == 8) {
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
                if (C::RESIDENT && wit >= C::NWS) continue;             // resident weights: loaded with the first item only
                if ((p.dbg & 4) && wit >= C::NWS) continue;
                mbar_wait(&w_empty[ws], wph ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&w_full[ws], C::W_STAGE);
                    bulk_g2s(sW + ws * C::W_STAGE, reinterpret_cast<const uint8_t*>(p.w) + (size_t)st * C::W_STAGE, C::W_STAGE, &w_full[ws]);
                }
            }
            __syncwarp();
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        // One thread issues every tcgen05.mma of the CTA, so the issue loop itself is on the critical path: a 128x128x16
        // MMA occupies the tensor pipe for 64 cycles and the loop must spend fewer instructions than that per MMA.
        // Descriptors are therefore built once; inside the loop a descriptor is `base + constant` (the start-address
        // field is the low 14 bits in 16-byte units and never carries into the LBO field for addresses < 256 KB).
        const uint32_t IDESC = idesc_16(128, C::COUT, p.f16 != 0);
        {
            const bool skip_a = (p.dbg & 8) != 0, skip_w = (p.dbg & 4) != 0;
            const uint64_t a_desc0 = smem_desc(smem_u32(sA) + CONV_HALO * 16, C::PLANE, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(sW), C::WPLANE, 128);
            const int row_pitch = p.row_pitch;
            uint32_t wit = 0, ait = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
                const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                if (!(skip_a && ait >= 2)) mbar_wait(&a_full[as], ph);
                mbar_wait(&acc_empty[as], ph ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + as * 256;
                const uint64_t a_item = a_desc0 + (uint64_t)(as * (C::A_STAGE >> 4));
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const int shift = (tap / 3 - 1) * row_pitch + (tap % 3 - 1);      // rows == 16-byte units
                    const uint64_t a_tap = a_item + (uint64_t)(int64_t)shift;
#pragma unroll
                    for (int h = 0; h < C::STAGES_PER_TAP; ++h, ++wit) {
                        const uint32_t ws = wit % C::NWS, wph = (wit / C::NWS) & 1;
                        if (!((skip_w || C::RESIDENT) && wit >= C::NWS)) mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        const uint64_t b_st = b_desc0 + (uint64_t)(ws * (C::W_STAGE >> 4));
                        const uint64_t a_h = a_tap + (uint64_t)(h * (C::WK / 8) * (C::PLANE >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                                for (int kk = 0; kk < C::KSTEPS; ++kk) {
                                    umma_bf16(acc + mt * 128, a_h + (uint64_t)(2 * kk * (C::PLANE >> 4) + mt * 128),
                                              b_st + (uint64_t)(2 * kk * (C::WPLANE >> 4)), IDESC, kk == 0 ? accumulate : 1u);
                                }
                            }
                            if (!skip_w && !C::RESIDENT) umma_commit(&w_empty[ws]);     // weight stage reusable once these MMAs retire
                            if (tap == 8 && h == C::STAGES_PER_TAP - 1) { if (!skip_a) umma_commit(&a_empty[as]); umma_commit(&acc_full[as]); }
                        }
                        __syncwarp();
                        accumulate = 1;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 0-7) =====================
        const bool has_res = p.resid != nullptr, relu = p.relu != 0, f16 = p.f16 != 0;
        const size_t p_total = (size_t)p.p_total;
        const int mt = warp >> 2, wq = warp & 3;             // M tile of the item, TMEM lane quarter
        uint32_t ait = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ait) {
            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
            bool waited = false;
            {
                const int row = item * CONV_BM + mt * 128 + wq * 32 + lane;
                const size_t grow = (size_t)CONV_GUARD + row;
                const bool valid = (row < n_rows) && (p.rowvalid[grow] != 0);
                uint4 res[16];                                    // the residual does not depend on the MMAs: fetch it first
                if (has_res && !(p.dbg & 2)) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) res[q] = *reinterpret_cast<const uint4*>(p.resid + ((size_t)q * p_total + grow) * 8);
                }
                if (!waited) { mbar_wait(&acc_full[as], ph); tc_fence_after(); waited = true; }
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * 256 + mt * 128;
                uint32_t ra[32], rb[32];                          // TMEM → registers, software-pipelined
                tmem_ld32(taddr, ra);
                tmem_ld_wait();
                tmem_ld32(taddr + 32, rb);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, ra, res, has_res, sBias, 0, relu, valid, p.out, p_total, grow);
                tmem_ld_wait();
                tmem_ld32(taddr + 64, ra);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, rb, res + 4, has_res, sBias, 32, relu, valid, p.out, p_total, grow);
                tmem_ld_wait();
                tmem_ld32(taddr + 96, rb);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, ra, res + 8, has_res, sBias, 64, relu, valid, p.out, p_total, grow);
                tmem_ld_wait();
                tc_fence_before(); mbar_arrive(&acc_empty[as]);                        // this thread's part of the stage's two accumulators is drained (256 arrivals)
                if (!(p.dbg & 2)) pair_epi_chunk(f16, rb, res + 12, has_res, sBias, 96, relu, valid, p.out, p_total, grow);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// k_conv3x3_pair — the 128 -> 128 layer as a WEIGHT-STATIONARY CTA-PAIR kernel (cta_group::2).
//
// Why: with one CTA per tile the single-CTA kernel re-streams the whole 295 KB weight image from L2 into shared memory
// for every 256 rows (1.5 GB per launch) and each 128x128x16 MMA reads 8 KB of operands from shared memory in 64 cycles
// — exactly the shared-memory bandwidth, with the TMA fills on top.  A CTA pair splits B by N: each CTA holds HALF the
// weights (64 output channels x 1152 K = 147 KB), which fits in shared memory next to two activation stages, so the
// weights are loaded ONCE per launch and an MMA (M = 256: one 128-row tile per CTA, N = 128, K = 16) reads 4 KB of A and
// 2 KB of B per CTA per 64 cycles.  L2 -> SM traffic drops from 1.8 GB to the activations alone.
//
// Work item = 256 consecutive rows of the padded position stream, 128 per CTA (rank r takes rows 256 i + 128 r).
// Roles per CTA (192 threads): warps 0-3 epilogue, warp 4 TMA producer, warp 5 = MMA issuer in the leader (rank 0) /
// relay in the peer (forwards "my stage is full" to the leader's barriers with a remote mbarrier arrive; bulk copies can
// only signal a barrier of the CTA they write to).  tcgen05.commit is multicast to both CTAs' barriers.
constexpr int PAIR_HALO = 17;                                   // max |tap shift| = row_pitch + 1 <= 17  (W <= 15)
struct PairCfg {
    static constexpr int ROWS = 128 + 2 * PAIR_HALO;            // 162 rows per A tile incl. halo
    static constexpr int PLANE = ROWS * 16;                     // 2592 B per 8-channel chunk
    static constexpr int A_STAGE = 16 * PLANE;                  // 41,472 B
    static constexpr int WPLANE = 64 * 16;                      // 1 KB: 64 output channels x 8 input channels
    static constexpr int WTAP = 16 * WPLANE;                    // 16 KB per tap
    static constexpr int W_BYTES = 9 * WTAP;                    // 147,456 B: this CTA's half of the layer
    static constexpr int OFF_A = W_BYTES;
    static constexpr int OFF_BIAS = OFF_A + 2 * A_STAGE;
    static constexpr int OFF_BARS = OFF_BIAS + CONV_COUT * 4;
    static constexpr int OFF_TSLOT = OFF_BARS + 24 * 8;
    static constexpr int SMEM = OFF_TSLOT + 16;                 // 231,120 B
};

__global__ void __launch_bounds__(CONV_THREADS, 1) k_conv3x3_pair(const ConvParams p) {
    using C = PairCfg;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW = smem;
    uint8_t* sA = smem + C::OFF_A;
    float* sBias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TSLOT);
    uint64_t* a_full = bars;            // [2] own TMA (+ in the leader: the peer's relay) → leader: MMA issuer | peer: relay
    uint64_t* a_empty = bars + 2;       // [2] MMA commit (multicast) → own TMA producer
    uint64_t* acc_full = bars + 4;      // [2] MMA commit (multicast) → own epilogue
    uint64_t* acc_empty = bars + 6;     // [2] leader only: 4 local + 4 remote epilogue warps → MMA issuer
    uint64_t* w_full = bars + 8;        // [9] own weight half loaded (+ in the leader: the peer's), one barrier per tap: the first item's MMAs start
                                        //     behind tap 0's 16 KB instead of behind all 147 KB — the rest streams in under them

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_rows = p.n_boards_dev ? (*p.n_boards_dev) * p.board_pitch : p.n_rows;
    const int n_items = (n_rows + 255) / 256;
    const int first_item = (int)cluster_id_x(), item_step = (int)n_clusters_x();

    if (threadIdx.x == 0) {
        const uint32_t full_count = rank == 0 ? 2 : 1;      // the leader's "full" barriers also take the peer relay's arrival
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], full_count); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        for (int i = 0; i < 9; ++i) mbar_init(&w_full[i], full_count);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < CONV_COUT; i += CONV_THREADS) sBias[i] = p.bias[i];
    if (warp == 5) tmem_alloc2(tslot, 256);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (first_item < n_items) {
        if (warp == 4) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * C::W_BYTES;
                auto load_tap = [&](int tap) {
                    mbar_arrive_expect_tx(&w_full[tap], C::WTAP);
                    bulk_g2s(sW + tap * C::WTAP, wsrc + (size_t)tap * C::WTAP, C::WTAP, &w_full[tap]);
                };
                // issue order = consumption order: tap 0, the first item's activations, taps 1-8, then the steady-state stage loop.
                // PDL: the weights do not depend on the previous layer and go out before griddepcontrol.wait; activations after it.
                load_tap(0);
                if (p.dbg & 64) for (int tap = 1; tap < 9; ++tap) load_tap(tap);        // profiling experiment: all weights ahead of the first stage
                grid_dep_wait();
                grid_dep_launch();          // (after the wait, so a dependent grid can only start once this grid's own prerequisites are complete)
                uint32_t ait = 0;
                for (int item = first_item; item < n_items; item += item_step, ++ait) {
                    const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
                    if ((p.dbg & 8) && ait >= 2) continue;
                    mbar_wait(&a_empty[as], aph ^ 1);
                    mbar_arrive_expect_tx(&a_full[as], C::A_STAGE);
                    const int item_eff = p.reverse ? n_items - 1 - item : item;
                    const size_t row0 = (size_t)CONV_GUARD + (size_t)item_eff * 256 + rank * 128 - PAIR_HALO;
                    for (int kc = 0; kc < 16; ++kc)
                        bulk_g2s(sA + as * C::A_STAGE + kc * C::PLANE, p.in + ((size_t)kc * p.p_total + row0) * 8, C::PLANE, &a_full[as]);
                    if (ait == 0 && !(p.dbg & 64)) for (int tap = 1; tap < 9; ++tap) load_tap(tap);
                }
            }
            __syncwarp();
        } else if (warp == 5 && rank != 0) {
            // ===================== relay (peer CTA): my stage is full → second arrival on the leader's barrier =====================
            if (lane == 0) {
                // forward "landed" to the leader in the order the leader consumes: tap 0, first stage, taps 1-8, then the stages
                mbar_wait(&w_full[0], 0);
                mbar_arrive_cluster(&w_full[0], 0);
                uint32_t ait = 0;
                for (int item = first_item; item < n_items; item += item_step, ++ait) {
                    const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                    if ((p.dbg & 8) && ait >= 2) continue;
                    mbar_wait(&a_full[as], ph);
                    mbar_arrive_cluster(&a_full[as], 0);
                    if (ait == 0) for (int tap = 1; tap < 9; ++tap) { mbar_wait(&w_full[tap], 0); mbar_arrive_cluster(&w_full[tap], 0); }
                }
            }
            __syncwarp();
        } else if (warp == 5) {
            // ===================== MMA issuer (leader CTA, one thread) =====================
            // The tensor pipe accepts only ~2 MMAs ahead of execution (measured: the commit fires ~90 cycles after the last
            // issue), so this thread runs in lock-step with the pipe: every instruction between two MMAs beyond a handful, and
            // every barrier wait at an item boundary, is pipe idle time.  Hence: all 72 MMAs of an item unrolled with
            // descriptors that are `register + constant`, per-tap bases computed once per launch, one fused wait per item.
            const uint32_t IDESC = idesc_16(256, CONV_COUT, p.f16 != 0);
            {
                // the whole warp runs the loop converged; the tcgen05 instructions sit under elect_one()
                const bool skip_a = (p.dbg & 8) != 0;
                const uint64_t a_desc0 = smem_desc(smem_u32(sA) + PAIR_HALO * 16, C::PLANE, 128);
                const uint64_t b_desc0 = smem_desc(smem_u32(sW), C::WPLANE, 128);
                uint64_t a_tap0[9];
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
                    a_tap0[tap] = a_desc0 + (uint64_t)(int64_t)((tap / 3 - 1) * p.row_pitch + (tap % 3 - 1));     // tap shift in rows == 16-byte units
                const long long t_begin = p.trace ? clock64() : 0;
                long long wait_all = 0;
                mbar_wait_cluster(&w_full[0], 0);
                if (p.trace && lane == 0) p.trace[1024 + 2 * first_item] = clock64() - t_begin;      // cycles until tap 0 of both weight halves is resident
                uint32_t ait = 0;
                for (int item = first_item; item < n_items; item += item_step, ++ait) {
                    const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                    const bool tr = p.trace != nullptr && first_item == 0 && ait < 64 && lane == 0;
                    const long long tw0 = p.trace ? clock64() : 0;
                    if (tr) p.trace[ait * 8 + 0] = tw0;
                    if (skip_a && ait >= 2) mbar_wait_cluster(&acc_empty[as], ph ^ 1);
                    else mbar_wait2_cluster(&a_full[as], ph, &acc_empty[as], ph ^ 1);
                    if (p.trace) wait_all += clock64() - tw0;
                    if (tr) p.trace[ait * 8 + 3] = clock64();
                    tc_fence_after();
                    const uint32_t acc = tmem_base + as * 128;
                    const uint64_t a_off = (uint64_t)(as * (C::A_STAGE >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (ait == 0 && tap > 0) mbar_wait_cluster(&w_full[tap], 0);      // first item only: the weights stream in tap by tap behind it
                            const uint64_t a_tap = a_tap0[tap] + a_off;
                            const uint64_t b_tap = b_desc0 + (uint64_t)(tap * (C::WTAP >> 4));
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma2_bf16(acc, a_tap + (uint64_t)(2 * kk * (C::PLANE >> 4)), b_tap + (uint64_t)(2 * kk * (C::WPLANE >> 4)), IDESC,
                                           (kk == 0 && tap == 0) ? 0u : 1u);
                        }
                        if (!skip_a) umma2_commit_both(&a_empty[as]);
                        umma2_commit_both(&acc_full[as]);
                    }
                    __syncwarp();
                    if (tr) p.trace[ait * 8 + 4] = clock64();
                }
                if (p.trace && lane == 0) { p.trace[1024 + 2 * first_item + 1] = clock64() - t_begin; p.trace[1280 + 2 * first_item + 1] = wait_all; }
            }
        } else {
            // ===================== epilogue (warps 0-3): TMEM lanes 32w..32w+31 = rows of this CTA's tile =====================
            const bool has_res = p.resid != nullptr, relu = p.relu != 0, f16 = p.f16 != 0;
            const size_t p_total = (size_t)p.p_total;
            uint32_t ait = 0;
            grid_dep_wait();                // PDL: the residual is read, and the output written, only after the previous layer has completed
            for (int item = first_item; item < n_items; item += item_step, ++ait) {
                const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                const int row = (p.reverse ? n_items - 1 - item : item) * 256 + (int)rank * 128 + warp * 32 + lane;
                const size_t grow = (size_t)CONV_GUARD + row;
                const bool valid = (row < n_rows) && (p.rowvalid[grow] != 0);
                // the residual does not depend on the MMAs: fetch all of it before waiting for the accumulator
                uint4 res[16];
                if (has_res && !(p.dbg & 2)) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) res[q] = *reinterpret_cast<const uint4*>(p.resid + ((size_t)q * p_total + grow) * 8);
                }
                mbar_wait(&acc_full[as], ph);
                tc_fence_after();
                const bool tr = p.trace != nullptr && first_item == 0 && ait < 64 && warp == 0 && lane == 0;
                if (tr) p.trace[ait * 8 + 5 + rank * 512] = clock64();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 128;
                // TMEM → registers, software-pipelined: chunk c+1 is in flight while chunk c is converted and stored
                uint32_t ra[32], rb[32];
                tmem_ld32(taddr, ra);
                tmem_ld_wait();
                tmem_ld32(taddr + 32, rb);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, ra, res, has_res, sBias, 0, relu, valid, p.out, p_total, grow, (p.dbg & 32) != 0);
                tmem_ld_wait();
                tmem_ld32(taddr + 64, ra);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, rb, res + 4, has_res, sBias, 32, relu, valid, p.out, p_total, grow, (p.dbg & 32) != 0);
                tmem_ld_wait();
                tmem_ld32(taddr + 96, rb);
                if (!(p.dbg & 2)) pair_epi_chunk(f16, ra, res + 8, has_res, sBias, 64, relu, valid, p.out, p_total, grow, (p.dbg & 32) != 0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);      // accumulator drained: the leader's barrier (also from the leader itself)
                if (!(p.dbg & 2)) pair_epi_chunk(f16, rb, res + 12, has_res, sBias, 96, relu, valid, p.out, p_total, grow, (p.dbg & 32) != 0);
                if (tr) p.trace[ait * 8 + 6 + rank * 512] = clock64();
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                 // the peer's shared memory / TMEM / barriers stay alive until both CTAs are done
    if (warp == 5) tmem_dealloc2(tmem_base, 256);
}


// ------------------------------------------------------------------------------------------------------------------
// k_conv3x3_pair_wide — the same weight-stationary CTA-pair kernel for boards up to 19 wide (tap shift <= 21 rows).
// With a 21-row halo two full activation stages no longer fit next to the resident half-layer of weights (147 KB + 2 x
// 43.5 KB > 227 KB), so a stage is split by K: channel planes 0-7 are double-buffered (P0[2]), planes 8-15 single-buffered
// (P1).  An item runs its 36 MMAs over planes 0-7 first (all taps), then the 36 over planes 8-15; P1 of the NEXT item is
// loaded while that item's first half computes, P0 of the next item while the current item computes.  Same roles, barriers
// and epilogue as k_conv3x3_pair.
constexpr int WIDE_HALO = 21;                                   // row_pitch + 1 <= 21  (W <= 19)
struct WideCfg {
    static constexpr int ROWS = 128 + 2 * WIDE_HALO;            // 170
    static constexpr int PLANE = ROWS * 16;                     // 2720 B
    static constexpr int HALF = 8 * PLANE;                      // 21,760 B: 8 channel planes
    static constexpr int WPLANE = 64 * 16, WTAP = 16 * WPLANE, W_BYTES = 9 * WTAP;
    static constexpr int OFF_P0 = W_BYTES;                      // [2][HALF]
    static constexpr int OFF_P1 = OFF_P0 + 2 * HALF;            // [HALF]
    static constexpr int OFF_BIAS = OFF_P1 + HALF;
    static constexpr int OFF_BARS = OFF_BIAS + CONV_COUT * 4;
    static constexpr int OFF_TSLOT = OFF_BARS + 16 * 8;
    static constexpr int SMEM = OFF_TSLOT + 16;                 // 213,392 B
};

__global__ void __launch_bounds__(CONV_THREADS, 1) k_conv3x3_pair_wide(const ConvParams p) {
    using C = WideCfg;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW = smem;
    uint8_t* sP0 = smem + C::OFF_P0;
    uint8_t* sP1 = smem + C::OFF_P1;
    float* sBias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TSLOT);
    uint64_t* p0_full = bars;           // [2]
    uint64_t* p0_empty = bars + 2;      // [2]
    uint64_t* acc_full = bars + 4;      // [2]
    uint64_t* acc_empty = bars + 6;     // [2] leader only
    uint64_t* w_full = bars + 8;        // [1]
    uint64_t* p1_full = bars + 9;       // [1]
    uint64_t* p1_empty = bars + 10;     // [1]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_rows = p.n_boards_dev ? (*p.n_boards_dev) * p.board_pitch : p.n_rows;
    const int n_items = (n_rows + 255) / 256;
    const int first_item = (int)cluster_id_x(), item_step = (int)n_clusters_x();

    if (threadIdx.x == 0) {
        const uint32_t full_count = rank == 0 ? 2 : 1;
        for (int i = 0; i < 2; ++i) { mbar_init(&p0_full[i], full_count); mbar_init(&p0_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        mbar_init(w_full, full_count); mbar_init(p1_full, full_count); mbar_init(p1_empty, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < CONV_COUT; i += CONV_THREADS) sBias[i] = p.bias[i];
    if (warp == 5) tmem_alloc2(tslot, 256);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (first_item < n_items) {
        if (warp == 4) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                mbar_arrive_expect_tx(w_full, C::W_BYTES);
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * C::W_BYTES;
                for (int tap = 0; tap < 9; ++tap) bulk_g2s(sW + tap * C::WTAP, wsrc + (size_t)tap * C::WTAP, C::WTAP, w_full);
                grid_dep_wait();            // PDL: everything above ran under the previous layer's tail; its activations are needed from here on
                grid_dep_launch();          // (after the wait, so a dependent grid can only start once this grid's own prerequisites are complete)
                uint64_t pol_stream = 0;
                if (p.l2_hints & 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
                uint32_t ait = 0;
                for (int item = first_item; item < n_items; item += item_step, ++ait) {
                    const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
                    const int item_eff = p.reverse ? n_items - 1 - item : item;
                    const size_t row0 = (size_t)CONV_GUARD + (size_t)item_eff * 256 + rank * 128 - WIDE_HALO;
                    mbar_wait(&p0_empty[as], aph ^ 1);
                    mbar_arrive_expect_tx(&p0_full[as], C::HALF);
                    if (pol_stream) { for (int kc = 0; kc < 8; ++kc) bulk_g2s_hint(sP0 + as * C::HALF + kc * C::PLANE, p.in + ((size_t)kc * p.p_total + row0) * 8, C::PLANE, &p0_full[as], pol_stream); }
                    else for (int kc = 0; kc < 8; ++kc) bulk_g2s(sP0 + as * C::HALF + kc * C::PLANE, p.in + ((size_t)kc * p.p_total + row0) * 8, C::PLANE, &p0_full[as]);
                    mbar_wait(p1_empty, (ait & 1) ^ 1);
                    mbar_arrive_expect_tx(p1_full, C::HALF);
                    if (pol_stream) { for (int kc = 0; kc < 8; ++kc) bulk_g2s_hint(sP1 + kc * C::PLANE, p.in + ((size_t)(8 + kc) * p.p_total + row0) * 8, C::PLANE, p1_full, pol_stream); }
                    else for (int kc = 0; kc < 8; ++kc) bulk_g2s(sP1 + kc * C::PLANE, p.in + ((size_t)(8 + kc) * p.p_total + row0) * 8, C::PLANE, p1_full);
                }
            }
            __syncwarp();
        } else if (warp == 5 && rank != 0) {
            // ===================== relay (peer CTA) =====================
            if (lane == 0) {
                mbar_wait(w_full, 0);
                mbar_arrive_cluster(w_full, 0);
                uint32_t ait = 0;
                for (int item = first_item; item < n_items; item += item_step, ++ait) {
                    const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                    mbar_wait(&p0_full[as], ph);
                    mbar_arrive_cluster(&p0_full[as], 0);
                    mbar_wait(p1_full, ait & 1);
                    mbar_arrive_cluster(p1_full, 0);
                }
            }
            __syncwarp();
        } else if (warp == 5) {
            // ===================== MMA issuer (leader CTA) =====================
            const uint32_t IDESC = idesc_16(256, CONV_COUT, p.f16 != 0);
            const uint64_t p0_desc0 = smem_desc(smem_u32(sP0) + WIDE_HALO * 16, C::PLANE, 128);
            const uint64_t p1_desc0 = smem_desc(smem_u32(sP1) + WIDE_HALO * 16, C::PLANE, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(sW), C::WPLANE, 128);
            int64_t shift[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) shift[tap] = (int64_t)((tap / 3 - 1) * p.row_pitch + (tap % 3 - 1));
            mbar_wait_cluster(w_full, 0);
            uint32_t ait = 0;
            for (int item = first_item; item < n_items; item += item_step, ++ait) {
                const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                mbar_wait2_cluster(&p0_full[as], ph, &acc_empty[as], ph ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + as * 128;
                const uint64_t a0 = p0_desc0 + (uint64_t)(as * (C::HALF >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint64_t a_tap = a0 + (uint64_t)shift[tap];
                        const uint64_t b_tap = b_desc0 + (uint64_t)(tap * (C::WTAP >> 4));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma2_bf16(acc, a_tap + (uint64_t)(2 * kk * (C::PLANE >> 4)), b_tap + (uint64_t)(2 * kk * (C::WPLANE >> 4)), IDESC, (kk == 0 && tap == 0) ? 0u : 1u);
                    }
                    umma2_commit_both(&p0_empty[as]);
                }
                __syncwarp();
                mbar_wait_cluster(p1_full, ait & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint64_t a_tap = p1_desc0 + (uint64_t)shift[tap];
                        const uint64_t b_tap = b_desc0 + (uint64_t)(tap * (C::WTAP >> 4) + 8 * (C::WPLANE >> 4));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma2_bf16(acc, a_tap + (uint64_t)(2 * kk * (C::PLANE >> 4)), b_tap + (uint64_t)(2 * kk * (C::WPLANE >> 4)), IDESC, 1u);
                    }
                    umma2_commit_both(p1_empty);
                    umma2_commit_both(&acc_full[as]);
                }
                __syncwarp();
            }
        } else {
            // ===================== epilogue (warps 0-3) =====================
            const bool has_res = p.resid != nullptr, relu = p.relu != 0, f16 = p.f16 != 0;
            const size_t p_total = (size_t)p.p_total;
            uint32_t ait = 0;
            uint64_t pol = 0;
            if (p.l2_hints & 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            grid_dep_wait();                // PDL: the residual is read, and the output written, only after the previous layer has completed
            for (int item = first_item; item < n_items; item += item_step, ++ait) {
                const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                const int row = (p.reverse ? n_items - 1 - item : item) * 256 + (int)rank * 128 + warp * 32 + lane;
                const size_t grow = (size_t)CONV_GUARD + row;
                const bool valid = (row < n_rows) && (p.rowvalid[grow] != 0);
                uint4 res[16];
                if (has_res) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) res[q] = *reinterpret_cast<const uint4*>(p.resid + ((size_t)q * p_total + grow) * 8);
                }
                mbar_wait(&acc_full[as], ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 128;
                uint32_t ra[32], rb[32];
                tmem_ld32(taddr, ra);
                tmem_ld_wait();
                tmem_ld32(taddr + 32, rb);
                pair_epi_chunk(f16, ra, res, has_res, sBias, 0, relu, valid, p.out, p_total, grow, false, false, pol);
                tmem_ld_wait();
                tmem_ld32(taddr + 64, ra);
                pair_epi_chunk(f16, rb, res + 4, has_res, sBias, 32, relu, valid, p.out, p_total, grow, false, false, pol);
                tmem_ld_wait();
                tmem_ld32(taddr + 96, rb);
                pair_epi_chunk(f16, ra, res + 8, has_res, sBias, 64, relu, valid, p.out, p_total, grow, false, false, pol);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
                pair_epi_chunk(f16, rb, res + 12, has_res, sBias, 96, relu, valid, p.out, p_total, grow, false, false, pol);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 5) tmem_dealloc2(tmem_base, 256);
}


// ------------------------------------------------------------------------------------------------------------------
// k_trunk_pair — ALL residual-block layers of the trunk in ONE persistent launch of the weight-stationary CTA-pair kernel, for
// 256-row boards (Gomoku 15x15: work item = exactly one board).
//
// Boards are independent: a conv layer of board b reads only board b's 256 rows (its halo rows above are the structural zero rows
// of the position stream, the ones below only feed masked padding outputs).  So a CTA pair can carry ITS OWN boards — groups of
// TRUNK_GROUP = 7 work items — through every layer with no grid-wide synchronisation: layer l+1 of item j needs layer l of item j,
// which the same pair produced.  What changes against k_conv3x3_pair:
//   * the pair walks (group, layer, item); activations ping-pong between X and Y in global memory, but a group is small enough
//     (74 pairs x 7 boards x 2 x 64 KB = 68 MB) to stay in the 126 MB L2 for all 20 layers: the trunk stops touching HBM
//     (per-layer launches: 670 MB per layer, the residual layers run 10 % slower because of it);
//   * weights: a ROLLING per-tap reload — the issuer commits w_empty[tap] behind the last item of a layer, the producer refills
//     that tap with the next layer's weights while the remaining taps of the old layer are still being multiplied;
//   * the epilogue publishes "these items of this layer are in memory" on two out_ready barriers in both CTAs (cluster-scope
//     release, twice per layer; the next layer's TMA loads of either CTA read rows written by both), the producer waits for
//     them (+ proxy fence) before loading;
//   * the per-layer bias lives in shared memory, double-buffered by layer parity (read from global in the epilogue it cost 45 %).
// Group size: the largest number of boards per pair whose X + Y stay L2-resident over all pairs; TRUNK_GROUP = 7 also divides the
// 55-56 boards a pair owns at 4096 boards into full groups.
constexpr int TRUNK_GROUP = 7;
constexpr int TRUNK_BATCHED_MIN = 5;        // groups of at least this many items publish their outputs twice per layer (see the epilogue), shorter ones once
// taps [T0, T1) of one item: 8 MMAs per tap, straight-line; COMMIT_W: a tcgen05.commit on w_empty[tap] behind each tap
template <int T0, int T1, bool COMMIT_W>
__device__ __forceinline__ void trunk_issue_taps(uint32_t acc, const uint64_t* a_tap0, uint64_t a_off, uint64_t b_desc0, uint32_t idesc, uint64_t* w_empty) {
#pragma unroll
    for (int tap = T0; tap < T1; ++tap) {
        const uint64_t a_tap = a_tap0[tap] + a_off;
        const uint64_t b_tap = b_desc0 + (uint64_t)(tap * (PairCfg::WTAP >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            umma2_bf16(acc, a_tap + (uint64_t)(2 * kk * (PairCfg::PLANE >> 4)), b_tap + (uint64_t)(2 * kk * (PairCfg::WPLANE >> 4)), idesc, (kk == 0 && tap == 0) ? 0u : 1u);
        if (COMMIT_W) umma2_commit_both(&w_empty[tap]);
    }
}
struct TrunkCfg : PairCfg {
    static constexpr int OFF_TBIAS = PairCfg::OFF_BIAS;         // [2][128] fp32: the current layer's folded BatchNorm shifts, double-buffered by layer parity
    static constexpr int OFF_TBARS = OFF_TBIAS + 2 * CONV_COUT * 4;
    static constexpr int OFF_TTSLOT = OFF_TBARS + 40 * 8;
    static constexpr int SMEM = OFF_TTSLOT + 16;
};

template <bool CONTIG>      // CONTIG: groups of consecutive boards of any size (rows past a group's end are not stored); else 256-row boards, strided groups
__global__ void __launch_bounds__(CONV_THREADS, 1) k_trunk_pair(const TrunkParams p) {
    using C = TrunkCfg;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW = smem;
    uint8_t* sA = smem + C::OFF_A;
    float* sBias2 = reinterpret_cast<float*>(smem + C::OFF_TBIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_TBARS);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::OFF_TTSLOT);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* acc_full = bars + 4;      // [2]
    uint64_t* acc_empty = bars + 6;     // [2] leader only
    uint64_t* w_full = bars + 8;        // [9] tap landed (leader: + the peer's relay)
    uint64_t* w_empty = bars + 17;      // [9] the layer's last MMAs on this tap have completed (multicast commit): the tap may be refilled
    uint64_t* out_ready = bars + 26;    // [2] the current layer's first nj-2 items / last 2 items are in memory: one arrival per CTA of the pair

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // Work: groups of `group_boards` consecutive boards; group gi belongs to pair gi % #pairs.  Inside a group the 256-row work items start
    // at the group's first row (a pair-local item grid: boards of any padded size), the last one may reach past the group's end — those
    // rows belong to another pair and are computed but never stored.
    const int n_boards = p.n_boards_dev ? *p.n_boards_dev : p.n_rows / p.board_pitch;
    const int first_group = (int)cluster_id_x(), group_step = (int)n_clusters_x();
    const int GB = (CONTIG && p.balance) ? max(1, min(p.group_boards, (n_boards + group_step - 1) / group_step)) : p.group_boards;
    const int L = p.n_layers;                                    // even: 2 per residual block
    // Two ways to form a group.  Boards that are exactly one work item (256 rows, Gomoku 15x15): the group's GB boards are STRIDED by the
    // number of pairs, so at any time the pairs stream through consecutive boards together (3 % faster than contiguous groups: the
    // write-backs stay page-local).  Any other board size: GB consecutive boards on the pair-local item grid.
    constexpr bool strided = !CONTIG;
    const int n_groups_total = strided ? group_step * (((n_boards + group_step - 1) / group_step + GB - 1) / GB) : (n_boards + GB - 1) / GB;
    auto group_rows = [&](int gi) { return min(GB, n_boards - gi * GB) * p.board_pitch; };                                   // contiguous mode
    auto group_items = [&](int gi) {
        if (!strided) return (group_rows(gi) + 255) / 256;
        const int left = n_boards - gi % group_step - group_step * GB * (gi / group_step);                               // boards from this group's first one on
        return left <= 0 ? 0 : min(GB, (left + group_step - 1) / group_step);
    };
    auto item_base = [&](int gi, int j) {                                                                                // first row of item j
        return strided ? (gi % group_step + group_step * (GB * (gi / group_step) + j)) * 256 : gi * GB * p.board_pitch + j * 256;
    };
    const int item_stride = strided ? group_step * 256 : 256;                                                            // rows between two items of a group

    if (threadIdx.x == 0) {
        const uint32_t full_count = rank == 0 ? 2 : 1;
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], full_count); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        for (int i = 0; i < 9; ++i) { mbar_init(&w_full[i], full_count); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) mbar_init(&out_ready[i], 2);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc2(tslot, 256);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (first_group < n_groups_total) {
        if (warp == 4) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                uint32_t ait = 0, wl = 0;                                 // activation stage counter, layer-instance counter
                for (int g = first_group; g < n_groups_total; g += group_step) {
                    const int nj = group_items(g);
                    if (nj == 0) continue;
                    const int gbase = item_base(g, 0);
                    for (int l = 0; l < L; ++l, ++wl) {
                        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w[l]) + (size_t)rank * C::W_BYTES;
                        const __nv_bfloat16* in = (l & 1) ? p.Y : p.X;
                        auto load_tap = [&](int tap) {
                            mbar_wait(&w_empty[tap], (wl & 1) ^ 1);       // the previous layer is done with this tap (first layer: passes at once)
                            mbar_arrive_expect_tx(&w_full[tap], C::WTAP);
                            bulk_g2s(sW + tap * C::WTAP, wsrc + (size_t)tap * C::WTAP, C::WTAP, &w_full[tap]);
                        };
                        load_tap(0);
                        for (int j = 0; j < nj; ++j, ++ait) {
                            const uint32_t as = ait & 1, aph = (ait >> 1) & 1;
                            if (l > 0 && (j == 0 || (nj >= TRUNK_BATCHED_MIN && j == nj - 2))) {
                                // both CTAs have written layer l-1 of the boards loaded from here on: out_ready[0] covers the items before the
                                // last two (all items in a short group), out_ready[1] the last two
                                mbar_wait_cluster(&out_ready[j == 0 ? 0 : 1], (uint32_t)((l - 1) & 1));
                                if (!(p.dbg & 2)) asm volatile("fence.proxy.async.global;" ::: "memory");   // their generic-proxy stores → this thread's TMA loads
                            }
                            mbar_wait(&a_empty[as], aph ^ 1);
                            mbar_arrive_expect_tx(&a_full[as], C::A_STAGE);
                            const size_t row0 = (size_t)CONV_GUARD + (size_t)(gbase + j * item_stride) + rank * 128 - PAIR_HALO;
                            for (int kc = 0; kc < 16; ++kc)
                                bulk_g2s(sA + as * C::A_STAGE + kc * C::PLANE, in + ((size_t)kc * p.p_total + row0) * 8, C::PLANE, &a_full[as]);
                            if (j == 0) for (int tap = 1; tap < 9; ++tap) load_tap(tap);
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 5 && rank != 0) {
            // ===================== relay (peer CTA) =====================
            if (lane == 0) {
                uint32_t ait = 0, wl = 0;
                for (int g = first_group; g < n_groups_total; g += group_step) {
                    const int nj = group_items(g);
                    if (nj == 0) continue;
                    for (int l = 0; l < L; ++l, ++wl) {
                        mbar_wait(&w_full[0], wl & 1);
                        mbar_arrive_cluster(&w_full[0], 0);
                        for (int j = 0; j < nj; ++j, ++ait) {
                            const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                            mbar_wait(&a_full[as], ph);
                            mbar_arrive_cluster(&a_full[as], 0);
                            if (j == 0) for (int tap = 1; tap < 9; ++tap) { mbar_wait(&w_full[tap], wl & 1); mbar_arrive_cluster(&w_full[tap], 0); }
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 5) {
            // ===================== MMA issuer (leader CTA; converged warp, tcgen05 under elect.sync) =====================
            const uint32_t IDESC = idesc_16(256, CONV_COUT, p.f16 != 0);
            const uint64_t a_desc0 = smem_desc(smem_u32(sA) + PAIR_HALO * 16, C::PLANE, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(sW), C::WPLANE, 128);
            uint64_t a_tap0[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) a_tap0[tap] = a_desc0 + (uint64_t)(int64_t)((tap / 3 - 1) * p.row_pitch + (tap % 3 - 1));
            uint32_t ait = 0, wl = 0;
            for (int g = first_group; g < n_groups_total; g += group_step) {
                const int nj = group_items(g);
                    if (nj == 0) continue;
                for (int l = 0; l < L; ++l, ++wl) {
                    for (int j = 0; j < nj; ++j, ++ait) {
                        const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                        const bool first = j == 0, last = j == nj - 1;
                        mbar_wait2_cluster(&a_full[as], ph, &acc_empty[as], ph ^ 1);
                        tc_fence_after();
                        const uint32_t acc = tmem_base + as * 128;
                        const uint64_t a_off = (uint64_t)(as * (C::A_STAGE >> 4));
                        // The issue loop must stay a handful of instructions per MMA (see k_conv3x3_pair), so the layer-boundary work is kept out
                        // of it: three straight-line variants of the same 72 MMAs.
                        if (!first && !last) {
                            if (elect_one()) { trunk_issue_taps<0, 9, false>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); umma2_commit_both(&a_empty[as]); umma2_commit_both(&acc_full[as]); }
                            __syncwarp();
                        } else if (!first) {         // last item of the layer: hand each tap back to the producer as soon as its MMAs are issued
                            if (elect_one()) { trunk_issue_taps<0, 9, true>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); umma2_commit_both(&a_empty[as]); umma2_commit_both(&acc_full[as]); }
                            __syncwarp();
                        } else {                     // first item: the layer's weights stream in behind it; wait for them three taps at a time (converged warp)
                            for (int t = 0; t < 3; ++t) mbar_wait_cluster(&w_full[t], wl & 1);
                            if (elect_one()) { if (last) trunk_issue_taps<0, 3, true>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); else trunk_issue_taps<0, 3, false>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); }
                            __syncwarp();
                            for (int t = 3; t < 6; ++t) mbar_wait_cluster(&w_full[t], wl & 1);
                            if (elect_one()) { if (last) trunk_issue_taps<3, 6, true>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); else trunk_issue_taps<3, 6, false>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); }
                            __syncwarp();
                            for (int t = 6; t < 9; ++t) mbar_wait_cluster(&w_full[t], wl & 1);
                            if (elect_one()) {
                                if (last) trunk_issue_taps<6, 9, true>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty); else trunk_issue_taps<6, 9, false>(acc, a_tap0, a_off, b_desc0, IDESC, w_empty);
                                umma2_commit_both(&a_empty[as]); umma2_commit_both(&acc_full[as]);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        } else {
            // ===================== epilogue (warps 0-3) =====================
            const size_t p_total = (size_t)p.p_total;
            const bool f16 = p.f16 != 0;
            uint32_t ait = 0;
            bool pending_b = false;
            // X — the skip connection, live across two layers (input of the next layer, residual of the one after) — is stored with an L2 evict_last
            // policy: with the dead Y rows discarded this brings the launch to 1.12 GB written (ncu), profiles/r2_summary.md 17
            uint64_t pol_last = 0;
            if (p.discard) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            // "these items of the current layer are in memory", to both CTAs' producers.  One publication per CTA: a named barrier orders the
            // four epilogue warps' stores before thread 0's cluster-scope release fence; the generic → async proxy fence sits on the consumer
            // side, one thread, right before the TMA loads.  The fence costs an L2 round trip, so a layer publishes only twice: its first
            // nj-2 items when the epilogue reaches item nj-2 (out_ready[0]: the next layer's first load needs them about then), its last two
            // at item 1 of the NEXT layer (out_ready[1]: needed by that layer's load of item nj-2).  Both points are a full MMA phase behind
            // the stores they cover, so the fence finds them drained.  Groups shorter than TRUNK_BATCHED_MIN items (the tail of the board
            // list) publish everything on out_ready[0] right behind their last item.
            auto publish = [&](int which) {
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (threadIdx.x == 0) {
                    if (!(p.dbg & 1)) asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    // default-semantics arrives behind the explicit cluster-scope fence (as CUTLASS' ClusterBarrier does for its cross-CTA
                    // consumer release); arrive.release.cluster on both would cost two more memory barriers: measured 3.26 -> 3.53 ms per forward
                    mbar_arrive(&out_ready[which]);
                    mbar_arrive_cluster(&out_ready[which], rank ^ 1);
                }
            };
            for (int g = first_group; g < n_groups_total; g += group_step) {
                const int nj = group_items(g);
                    if (nj == 0) continue;
                    const int gbase = item_base(g, 0);
                const int lim = strided ? 0x7fffffff : g * GB * p.board_pitch + group_rows(g);      // strided boards are whole items: every row is ours
                for (int l = 0; l < L; ++l) {
                    const bool has_res = (l & 1) != 0;                                   // second conv of a block: + the block's input (in place on X)
                    __nv_bfloat16* out = (l & 1) ? p.X : p.Y;
                    // this layer's bias → shared memory (the epilogue reads all 128 per row); the barrier also keeps the four warps within one
                    // layer of each other, which is what makes the two-deep buffer safe
                    float* bias = sBias2 + (l & 1) * CONV_COUT;
                    bias[threadIdx.x] = p.bias[l][threadIdx.x];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int j = 0; j < nj; ++j, ++ait) {
                        const uint32_t as = ait & 1, ph = (ait >> 1) & 1;
                        const int row = gbase + j * item_stride + (int)rank * 128 + warp * 32 + lane;
                        const size_t grow = (size_t)CONV_GUARD + row;
                        const bool mine = CONTIG ? row < lim : true;                     // rows past the group's end: computed, never stored
                        const bool valid = mine && (p.rowvalid[grow] != 0);
                        uint4 res[16];
                        if (has_res) {                                                   // (uniform per layer; rows that are not ours are readable, just not ours to write)
#pragma unroll
                            for (int q = 0; q < 16; ++q) res[q] = *reinterpret_cast<const uint4*>(p.X + ((size_t)q * p_total + grow) * 8);
                        }
                        mbar_wait(&acc_full[as], ph);
                        tc_fence_after();
                        if constexpr (!CONTIG) {
                            // The MMAs of this item are complete, so its input rows are consumed.  On odd layers the input is Y, which nobody reads again
                            // before the next even layer overwrites it: drop those lines from L2 (discard.global.L2) instead of letting them be written
                            // back — ncu: 3.12 -> 1.76 GB written, 776 -> 531 MB read per 20-layer launch (profiles/r2_summary.md 17).  Rows 232-255 of a
                            // board stay (the next board's halo reads them as zeros); X is never dead (skip connection, overwritten in place).
                            if (p.discard && has_res) {
                                const int r = (int)rank * 128 + warp * 32 + lane;               // row inside the 256-row board
                                if ((r & 7) == 0 && r < 232) {
                                    const char* y0 = reinterpret_cast<const char*>(p.Y) + ((size_t)CONV_GUARD + row) * 16;
#pragma unroll
                                    for (int q = 0; q < 16; ++q) asm volatile("discard.global.L2 [%0], 128;" :: "l"(y0 + (size_t)q * p_total * 16) : "memory");
                                }
                            }
                        } else {
                            // Board-aligned groups: items are consecutive 256-row windows of the group, so the last 17 rows of item j are still the top
                            // halo of item j + 1 — every item drops the window shifted back by 24 rows.  Never the group's last 17 rows (the next group's
                            // halo reads them as zeros) and nothing outside [gbase, lim) (another pair's rows).
                            if (p.discard && has_res) {
                                const int rp = row - 24;
                                if ((rp & 7) == 0 && rp >= gbase && rp + 8 <= lim - 17) {
                                    const char* y0 = reinterpret_cast<const char*>(p.Y) + ((size_t)CONV_GUARD + rp) * 16;
#pragma unroll
                                    for (int q = 0; q < 16; ++q) asm volatile("discard.global.L2 [%0], 128;" :: "l"(y0 + (size_t)q * p_total * 16) : "memory");
                                }
                            }
                        }
                        // deferred publication of the PREVIOUS item: its stores were issued a whole MMA phase ago, so the cluster-scope release fence
                        // (which waits for this thread's outstanding stores) finds them drained
                        if (pending_b && (j == 1 || nj < TRUNK_BATCHED_MIN)) { publish(1); pending_b = false; }
                        if (nj >= TRUNK_BATCHED_MIN && j == nj - 2) publish(0);
                        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 128;
                        uint32_t ra[32], rb[32];
                        tmem_ld32(taddr, ra);
                        tmem_ld_wait();
                        tmem_ld32(taddr + 32, rb);
                        pair_epi_chunk(f16, ra, res, has_res, bias, 0, true, valid, out, p_total, grow, false, CONTIG && !mine, has_res ? pol_last : 0);
                        tmem_ld_wait();
                        tmem_ld32(taddr + 64, ra);
                        pair_epi_chunk(f16, rb, res + 4, has_res, bias, 32, true, valid, out, p_total, grow, false, CONTIG && !mine, has_res ? pol_last : 0);
                        tmem_ld_wait();
                        tmem_ld32(taddr + 96, rb);
                        pair_epi_chunk(f16, ra, res + 8, has_res, bias, 64, true, valid, out, p_total, grow, false, CONTIG && !mine, has_res ? pol_last : 0);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
                        pair_epi_chunk(f16, rb, res + 12, has_res, bias, 96, true, valid, out, p_total, grow, false, CONTIG && !mine, has_res ? pol_last : 0);
                        if (j == nj - 1) { if (nj >= TRUNK_BATCHED_MIN) pending_b = true; else publish(0); }
                    }
                }
            }
            if (pending_b) publish(1);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 5) tmem_dealloc2(tmem_base, 256);
}

}  // namespace

size_t conv_smem_bytes(int cin) { return cin == 16 ? Cfg<16>::SMEM : (cin == 32 ? Cfg<32>::SMEM : Cfg<128>::SMEM); }

bool conv_uses_pair(int cin, int row_pitch) { return cin == CONV_COUT && row_pitch + 1 <= WIDE_HALO; }     // same weight image for both pair kernels

int conv3x3_launch(const ConvParams& p, int cin, int grid, cudaStream_t stream) {
    cudaError_t err;
    if (cin == 32) {     // chess stem: 18 planes padded to 32 channels
        err = smem_opt_in((const void*)k_conv3x3<32>, (int)Cfg<32>::SMEM); if (err) return (int)err;
        k_conv3x3<32><<<grid, CONV1_THREADS, Cfg<32>::SMEM, stream>>>(p);
    } else if (cin == 16) {
        err = smem_opt_in((const void*)k_conv3x3<16>, (int)Cfg<16>::SMEM); if (err) return (int)err;
        k_conv3x3<16><<<grid, CONV1_THREADS, Cfg<16>::SMEM, stream>>>(p);
    } else if (conv_uses_pair(cin, p.row_pitch) && p.row_pitch + 1 > PAIR_HALO) {          // boards 16..19 wide: K-split stages
        err = smem_opt_in((const void*)k_conv3x3_pair_wide, (int)WideCfg::SMEM); if (err) return (int)err;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(grid & ~1)); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = WideCfg::SMEM; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = p.pdl ? 2 : 1;
        err = cudaLaunchKernelEx(&cfg, k_conv3x3_pair_wide, p);
        if (err) return (int)err;
    } else if (conv_uses_pair(cin, p.row_pitch)) {
        err = smem_opt_in((const void*)k_conv3x3_pair, (int)PairCfg::SMEM); if (err) return (int)err;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(grid & ~1)); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = PairCfg::SMEM; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = p.pdl ? 2 : 1;
        err = cudaLaunchKernelEx(&cfg, k_conv3x3_pair, p);
        if (err) return (int)err;
    } else if (cin == 128) {
        err = smem_opt_in((const void*)k_conv3x3<128>, (int)Cfg<128>::SMEM); if (err) return (int)err;
        k_conv3x3<128><<<grid, CONV1_THREADS, Cfg<128>::SMEM, stream>>>(p);
    } else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}

bool trunk_fused_supported(int channels, int board_pitch, int row_pitch, int n_layers) {
    return channels == CONV_COUT && row_pitch + 1 <= PAIR_HALO && board_pitch <= TRUNK_GROUP * 256 / TRUNK_BATCHED_MIN && n_layers >= 2 && n_layers % 2 == 0 && n_layers <= TRUNK_MAX_LAYERS;
}
// boards per group: as many whole boards as fit in TRUNK_GROUP work items (Gomoku 15x15: 7 boards = 7 items; Go 9x9: 17 boards = 6.6 items)
int trunk_group_boards(int board_pitch) { return TRUNK_GROUP * 256 / board_pitch; }

int trunk_launch(const TrunkParams& p, int grid, cudaStream_t stream) {
    cudaError_t err;
    err = smem_opt_in((const void*)k_trunk_pair<false>, (int)TrunkCfg::SMEM); if (err) return (int)err;
    err = smem_opt_in((const void*)k_trunk_pair<true>, (int)TrunkCfg::SMEM); if (err) return (int)err;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(grid & ~1)); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = TrunkCfg::SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    err = p.board_pitch == 256 ? cudaLaunchKernelEx(&cfg, k_trunk_pair<false>, p) : cudaLaunchKernelEx(&cfg, k_trunk_pair<true>, p);
    if (err) return (int)err;
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
