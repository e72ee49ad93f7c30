// engine.cu — host side of the batched self-play engine + its C ABI (include/az_b200.h).
//
// The host owns no search logic: it allocates the pools, launches the wave kernels in order
// (select → evaluator → expand/backup) and the per-move kernels (choose → re-root → game turnover),
// and moves results across the ABI.  Reference counterparts: ParallelMCTS (src/mcts/parallel_mcts.cpp),
// SelfPlayManager (src/selfplay/self_play_manager.cpp), TorchNeuralNetwork (src/nn/torch_neural_network.cpp).
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <memory>
#include <cmath>
#include <algorithm>
#include <unordered_set>
#include <set>
#include <mutex>
#include <cuda.h>

#include "../../include/az_b200.h"
#include "common.cuh"

namespace az {
static thread_local std::string g_error;
void set_error(const std::string& s) { g_error = s; }

cudaError_t smem_opt_in(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;      // (kernel, device)
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kernel, dev})) return cudaSuccess;
    if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) return e;
    done.insert({kernel, dev});
    return cudaSuccess;
}
}  // namespace az

#include "gomoku.cuh"
#include "go.cuh"
#include "chess.cuh"
#include "tree_kernels.cuh"
#include "conv_trunk.cuh"
#include "heads.cuh"
#include "gemm_tc.cuh"
#include "head_conv.cuh"

namespace az {

#define AZ_CHECK(cond, msg) do { if (!(cond)) { set_error(std::string(msg)); return -1; } } while (0)
#define AZ_LAUNCH_CHECK() AZ_CUDA_CHECK(cudaGetLastError())

template <class T>
static int dev_alloc(T** p, size_t n) { AZ_CUDA_CHECK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T))); return 0; }

// ------------------------------------------------------------------------------------------------ weight preparation
// az_engine_load_weights hands over the trainer's fp32 tensors (AZW1 blob).  Folding eval-mode BatchNorm into the conv
// weights, converting to bf16 and laying the weights out as the kernels' shared-memory images runs on the device: the
// host only uploads the blob.  scale = gamma / sqrt(var + 1e-5), shift = beta - mean * scale (bn = [gamma|beta|mean|var]).
AZ_D float bn_scale(const float* bn, int n, int i) { return bn[i] / sqrtf(bn[3 * n + i] + 1e-5f); }

// One 128-output-channel x `cin`-input-channel sub-block (co0.., ci0..) of a layer whose full weight tensor is
// w[C][cin_total][9]: the conv kernels are built for 128 output channels, wider trunks run as channel slices (Net::forward).
AZ_D __nv_bfloat16 w16(float v, int f16) { return net16(v, f16 != 0); }      // a conv weight in the network's 16-bit type (fp16 subnormals below 6e-5: absolute error <= 3e-8)
__global__ void k_prep_conv(const float* __restrict__ w, const float* __restrict__ bn, __nv_bfloat16* img, float* bias,
                            int C, int cin_total, int co0, int ci0, int cin_real, int cin, int pair, int f16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int CO = nn::CONV_COUT;
    if (idx < CO && bias) bias[idx] = bn[C + co0 + idx] - bn[2 * C + co0 + idx] * bn_scale(bn, C, co0 + idx);
    if (idx >= CO * cin * 9) return;
    const int t = idx % 9, ci = (idx / 9) % cin, co = idx / (9 * cin);
    const float v = ci < cin_real ? w[((size_t)(co0 + co) * cin_total + ci0 + ci) * 9 + t] * bn_scale(bn, C, co0 + co) : 0.0f;
    img[nn::conv_weight_index(cin, pair != 0, t, ci, co)] = w16(v, f16);
}
// Head GEMM weight images: three-term bf16 split  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (K' = 3K), image = [W_hi | W_hi | W_lo]
AZ_D void put3(__nv_bfloat16* img, int K, int n, int k, float v) {
    const __nv_bfloat16 hi = __float2bfloat16(v), lo = __float2bfloat16(v - __bfloat162float(hi));
    img[nn::gemm_weight_index(3 * K, n, k)] = hi; img[nn::gemm_weight_index(3 * K, n, K + k)] = hi; img[nn::gemm_weight_index(3 * K, n, 2 * K + k)] = lo;
}
// both 1x1 convs as one [64 x C] matrix (rows 0-31 policy, 32-63 value), BatchNorm scale folded in; b1 = BN shifts
__global__ void k_prep_1x1(const float* __restrict__ pcw, const float* __restrict__ pbn, const float* __restrict__ vcw, const float* __restrict__ vbn,
                           __nv_bfloat16* g1, float* b1, int C, __nv_bfloat16* h1, int f16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 64) { const float* bn = idx < 32 ? pbn : vbn; const int o = idx & 31; b1[idx] = bn[32 + o] - bn[64 + o] * bn_scale(bn, 32, o); }
    if (idx >= 64 * C) return;
    const int c = idx % C, n = idx / C, o = n & 31;
    const float* cw = n < 32 ? pcw : vcw; const float* bn = n < 32 ? pbn : vbn;
    const float wf = cw[(size_t)o * C + c] * bn_scale(bn, 32, o);
    put3(g1, C, n, c, wf);
    if (h1 && C == 128) {       // the fused heads kernel's image (head_conv.cuh): bf16 hi / lo pair of the same folded weight
        // (fp16 mode: an fp16 hi / lo pair — the MMA's two operands share the type of the trunk output it multiplies)
        const __nv_bfloat16 hi = w16(wf, f16);
        const float hif = f16 ? __half2float(*reinterpret_cast<const __half*>(&hi)) : __bfloat162float(hi);
        h1[nn::head_conv_weight_index(0, n, c)] = hi;
        h1[nn::head_conv_weight_index(1, n, c)] = w16(wf - hif, f16);
    }
}
// FC weights with K re-ordered from torch's flatten order (ch*64 + cell) to the feature order the 1x1-conv GEMM writes
// (cell*32 + ch); rows >= n_rows stay zero (N padded to 256)
// split = 0: plain bf16 image (K' = K) for very wide heads (chess' 20480 actions: the three-term image would be 252 MB and its
// GEMM 3x the L2 traffic; the A operand then uses only the hi planes)
// policy FC weights row-major [n_rows][feat] bf16 with K in feature order (cell * 32 + channel): the legal-moves-only policy kernel reads one row per legal action
__global__ void k_prep_fc_rows(const float* __restrict__ w /*[n_rows][feat]*/, __nv_bfloat16* rows, int n_rows, int feat) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_rows * feat) return;
    const int kt = (int)(idx % feat), n = (int)(idx / feat);
    rows[(size_t)n * feat + (kt % 64) * 32 + kt / 64] = __float2bfloat16(w[idx]);
}
__global__ void k_prep_fc(const float* __restrict__ w /*[n_rows][feat]*/, __nv_bfloat16* img, int n_rows, int feat, int split) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_rows * feat) return;
    const int kt = (int)(idx % feat), n = (int)(idx / feat);
    const int ch = kt / 64, cell = kt % 64;
    if (split) put3(img, feat, n, cell * 32 + ch, w[idx]);
    else img[nn::gemm_weight_index(feat, n, cell * 32 + ch)] = __float2bfloat16(w[idx]);
}

// ------------------------------------------------------------------------------------------------ network
// ------------------------------------------------------------------------------------------------ SM partitions
// Green contexts (CUDA 12.4+ driver API, resolved at run time — the library does not link libcuda): the device's SMs are
// split into a large partition that runs nothing but the trunk's 128 -> 128 conv launches and a small one for everything
// else (tree kernels, stem, heads).  The small kernels are latency- / HBM-bound and cost far fewer SM-seconds on a dozen SMs
// than on 148, and with two stream groups they run in the shadow of the other group's tensor-core pass.
struct SmPartitions {
    CUgreenCtx conv_ctx = nullptr, small_ctx = nullptr;
    int conv_sms = 0, small_sms = 0;
    CUresult (*p_stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    CUresult (*p_ctx_destroy)(CUgreenCtx) = nullptr;
    template <class F> static bool sym(const char* name, F& f) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion(name, &fn, 12090, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            set_error(std::string("driver entry point not available: ") + name + " (SM partitioning needs a CUDA 12.5+ driver)"); return false;
        }
        f = reinterpret_cast<F>(fn); return true;
    }
    // conv partition >= want_conv SMs (rounded up by the driver to its granularity), the remainder is the small partition
    int create(int device, int want_conv) {
        CUresult (*p_dev_get)(CUdevice*, int) = nullptr;
        CUresult (*p_get_res)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
        CUresult (*p_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
        CUresult (*p_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
        CUresult (*p_create)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
        if (!(sym("cuDeviceGet", p_dev_get) && sym("cuDeviceGetDevResource", p_get_res) && sym("cuDevSmResourceSplitByCount", p_split) &&
              sym("cuDevResourceGenerateDesc", p_desc) && sym("cuGreenCtxCreate", p_create) && sym("cuGreenCtxStreamCreate", p_stream_create) &&
              sym("cuGreenCtxDestroy", p_ctx_destroy))) return -1;
        CUdevice dev; AZ_CHECK(p_dev_get(&dev, device) == CUDA_SUCCESS, "cuDeviceGet failed");
        CUdevResource all{}, big{}, rest{};
        AZ_CHECK(p_get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS, "cuDeviceGetDevResource failed");
        unsigned int nb = 1;
        AZ_CHECK(p_split(&big, &nb, &all, &rest, 0, (unsigned)want_conv) == CUDA_SUCCESS && nb == 1, "cuDevSmResourceSplitByCount failed");
        AZ_CHECK(rest.type == CU_DEV_RESOURCE_TYPE_SM && rest.sm.smCount >= 2, "SM split left no SMs for the small partition");
        CUdevResourceDesc d_big = nullptr, d_rest = nullptr;
        AZ_CHECK(p_desc(&d_big, &big, 1) == CUDA_SUCCESS && p_desc(&d_rest, &rest, 1) == CUDA_SUCCESS, "cuDevResourceGenerateDesc failed");
        AZ_CHECK(p_create(&conv_ctx, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS, "cuGreenCtxCreate (conv partition) failed");
        AZ_CHECK(p_create(&small_ctx, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS, "cuGreenCtxCreate (small partition) failed");
        conv_sms = (int)big.sm.smCount; small_sms = (int)rest.sm.smCount;
        return 0;
    }
    int stream(bool conv, cudaStream_t* out) {
        CUstream st = nullptr;
        AZ_CHECK(p_stream_create(&st, conv ? conv_ctx : small_ctx, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS, "cuGreenCtxStreamCreate failed");
        *out = (cudaStream_t)st; return 0;
    }
    void destroy() {
        if (p_ctx_destroy) { if (conv_ctx) p_ctx_destroy(conv_ctx); if (small_ctx) p_ctx_destroy(small_ctx); }
        conv_ctx = small_ctx = nullptr;
    }
};

struct NetWeights {            // device images
    std::vector<__nv_bfloat16*> conv_w;   // stem + 2*blocks
    std::vector<float*> conv_b;
    float *b1x1 = nullptr, *pfc_b = nullptr, *vfc1_b = nullptr, *vfc2_w = nullptr, *vfc2_b = nullptr;   // fp32 biases / tiny last layer
    float* zero_bias = nullptr;                                                                          // [128] for the accumulating channel-slice launches
    __nv_bfloat16 *g1_w = nullptr, *pfc_img = nullptr, *vfc1_img = nullptr;                              // tcgen05 GEMM weight images
    __nv_bfloat16* pfc_rows = nullptr;                                                                    // wide heads: policy FC weights row-major [A][feat] (legal-moves-only policy, heads.cu)
    __nv_bfloat16* h1_w = nullptr;                                                                        // fused heads kernel: hi / lo image of the folded 1x1 weights (head_conv.cuh)
    float* blob = nullptr; size_t blob_bytes = 0;                                                        // device copy of the last AZW1 blob
    int blocks = -1, in_planes = -1;                                                                     // shape the images above were allocated for
};

// Net = one set of activation buffers (one per stream group) + a pointer to the shared weights.
struct Net {
    int blocks = 0, C = 0, in_planes = 0, H = 0, W = 0, A = 0, PH = 0, PW = 0, feat = 0;
    int cin_pad = 16;          // stem input channels after zero padding: 16, or 32 for chess' 18 planes
    int p_tiles = 4;           // policy FC N tiles of 64: ceil(A / 64)
    int p_split = 1;           // policy FC as the three-term hi/lo bf16 split (1) or plain bf16 (0: heads wider than 1024 actions)
    int f16 = 1;               // 16-bit storage type of activations and conv weights: 1 = fp16 (default), 0 = bf16 (az_config.net_precision)
    int NS = 1;                // channel slices of 128: a C-channel layer runs as NS x NS launches of the 128 -> 128 kernel
    int wi(int l, int co, int ci) const { return l == 0 ? co : NS + (l - 1) * NS * NS + co * NS + ci; }   // weight image index
    int bi(int l, int co) const { return l * NS + co; }
    int row_pitch = 0, board_pitch = 0, p_total = 0, max_boards = 0;
    bool loaded = false;
    NetWeights w;              // owned by group 0's Net; other groups hold a shallow copy (share())
    bool owns_weights = true;
    __nv_bfloat16 *in16 = nullptr, *X = nullptr, *Y = nullptr;
    uint8_t* rowvalid = nullptr;
    __nv_bfloat16 *pooled = nullptr, *featP = nullptr, *featV = nullptr;   // head GEMM operands (bf16, chunk-plane layout)
    float *logits = nullptr, *hidden = nullptr;                            // final logits [n][A]; `hidden` unused since split-K (kept for the destroy list)
    float *logits_part = nullptr, *hidden_part = nullptr;                  // split-K partial sums of the two FC GEMMs: [fc_splits][max_boards][A | 256]
    int fc_splits = 4;                                                     // split-K of the FC GEMMs: 4, 8 or 16 (divisors of the 96 K stages), enough for ~one item per SM
    int ld_part = 0;                                                       // row pitch of the policy partial-sum slabs: p_tiles * 64 (16-byte stores in the GEMM epilogue)
    int p_splits = 4;                                                      // policy FC splits: 1 for very wide heads (the partial slabs would cost more than they save)
    int boards_cap = 0;
    int n_sms = 148;
    // SM partitions (SmPartitions): the trunk's 128 -> 128 layers go to `conv_stream` (conv partition, n_sms_conv SMs), everything else stays on
    // the caller's stream (small partition, n_sms SMs); ev_in / ev_out order the two.  conv_stream == nullptr: one stream, all SMs.
    cudaStream_t conv_stream = nullptr; cudaEvent_t ev_in = nullptr, ev_out = nullptr; int n_sms_conv = 0;
    unsigned long long launches = 0;
    // live timing of the trunk's 128 -> 128 conv launches inside production waves (every 64th forward): CUDA events on the launching stream
    cudaEvent_t tev[2] = {nullptr, nullptr}; unsigned long long fwd_count = 0, conv_sampled = 0; double conv_ms = 0.0;
    // AZ_WAVE_TIMING (profiling): when fe_on is set for a forward, events after stem / trunk / pool / 1x1 GEMM / policy FC / value FC / softmax
    cudaEvent_t fe[7] = {}; bool fe_on = false;
    void fe_rec(int i, cudaStream_t s) { if (fe_on) { if (!fe[i]) cudaEventCreate(&fe[i]); cudaEventRecord(fe[i], s); } }

    int init(int H_, int W_, int A_, int max_boards_, int channels, int planes) {
        H = H_; W = W_; A = A_; max_boards = max_boards_; C = channels;
        cin_pad = planes <= 16 ? 16 : 32; p_tiles = (A + 63) / 64; p_split = A <= 1024 ? 1 : 0; 
        {   // work items of an FC GEMM = 128-row tiles x groups of up to 4 weight tiles (gemm_tc.cu NSUB) x K splits
            const int m_tiles = (max_boards_ + 127) / 128;
            fc_splits = 4; while (fc_splits < 16 && m_tiles * fc_splits < 120) fc_splits *= 2;
            if (const char* f = getenv("AZ_FC_SPLITS")) { const int v = atoi(f); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) fc_splits = v; }
        }
        p_splits = A <= 1024 ? fc_splits : 1;
        ld_part = p_tiles * 64;
        AZ_CHECK(planes <= 32, "at most 32 input planes");
        row_pitch = W + 1; board_pitch = (H + 1) * (W + 1);
        AZ_CHECK(W + 2 <= nn::CONV_HALO, "board too wide for the conv halo");
        AZ_CHECK(channels == 128 || channels == 256, "conv trunk is built for 128 or 256 channels");
        NS = channels / nn::CONV_COUT;
        const size_t rows = (size_t)max_boards * board_pitch;
        p_total = (int)(nn::CONV_GUARD + (rows + nn::CONV_BM - 1) / nn::CONV_BM * nn::CONV_BM + nn::CONV_GUARD + (board_pitch % nn::CONV_BM == 0 ? 0 : nn::TRUNK_TAIL_ROWS));      // (boards of whole work items never overhang)
        PH = std::min(8, H); PW = std::min(8, W); feat = 32 * PH * PW;
        if (dev_alloc(&in16, (size_t)(cin_pad / 8) * p_total * 8)) return -1;
        if (dev_alloc(&X, (size_t)(C / 8) * p_total * 8)) return -1;
        if (dev_alloc(&Y, (size_t)(C / 8) * p_total * 8)) return -1;
        AZ_CUDA_CHECK(cudaMemset(in16, 0, (size_t)(cin_pad / 8) * p_total * 16));
        AZ_CUDA_CHECK(cudaMemset(X, 0, (size_t)(C / 8) * p_total * 16));
        AZ_CUDA_CHECK(cudaMemset(Y, 0, (size_t)(C / 8) * p_total * 16));
        std::vector<uint8_t> rv(p_total, 0);
        for (int b = 0; b < max_boards; ++b)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) rv[nn::CONV_GUARD + (size_t)b * board_pitch + y * row_pitch + x] = 1;
        if (dev_alloc(&rowvalid, (size_t)p_total)) return -1;
        AZ_CUDA_CHECK(cudaMemcpy(rowvalid, rv.data(), p_total, cudaMemcpyHostToDevice));
        AZ_CHECK(PH == 8 && PW == 8, "heads are built for an 8x8 pooled map (board >= 8)");
        boards_cap = (max_boards + 127) / 128 * 128;
        if (dev_alloc(&pooled, (size_t)2 * (C / 8) * 64 * boards_cap * 8) || dev_alloc(&featP, (size_t)512 * boards_cap * 8) || dev_alloc(&featV, (size_t)512 * boards_cap * 8)) return -1;
        AZ_CUDA_CHECK(cudaMemset(pooled, 0, (size_t)2 * (C / 8) * 64 * boards_cap * 16));
        AZ_CUDA_CHECK(cudaMemset(featP, 0, (size_t)512 * boards_cap * 16));
        AZ_CUDA_CHECK(cudaMemset(featV, 0, (size_t)512 * boards_cap * 16));
        if (dev_alloc(&logits, (size_t)max_boards * A)) return -1;
        if (dev_alloc(&logits_part, (size_t)p_splits * max_boards * ld_part) || dev_alloc(&hidden_part, (size_t)fc_splits * max_boards * 256)) return -1;
        if (dev_alloc(&hidden, (size_t)max_boards * 256)) return -1;
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
        return 0;
    }

    // AZW1 blob: header + fp32 tensors (net.py:export_weights).  The blob is uploaded as is; BatchNorm folding, bf16
    // conversion and the shared-memory image layouts are produced by the k_prep_* kernels above.  Device buffers are
    // allocated on the first load and reused while the network shape stays the same.
    int load(const void* blob, size_t bytes, cudaStream_t st) {
        struct Hdr { char magic[4]; int32_t version, blocks, channels, in_planes, H, W, actions; };
        AZ_CHECK(bytes >= sizeof(Hdr), "weight blob too small");
        const Hdr* h = (const Hdr*)blob;
        AZ_CHECK(std::memcmp(h->magic, "AZW1", 4) == 0 && h->version == 1, "bad weight blob magic/version");
        AZ_CHECK(h->channels == C && h->H == H && h->W == W && h->actions == A, "weight blob does not match engine config");
        AZ_CHECK(h->in_planes <= cin_pad && h->in_planes >= 1 && h->blocks >= 0, "weight blob has more input planes than the engine's game");
        const int nb = h->blocks, ip = h->in_planes, nconv = 1 + 2 * nb;
        // tensor offsets (in floats) inside the blob
        size_t off = 0;
        auto take = [&](size_t n) { const size_t o = off; off += n; return o; };
        std::vector<size_t> cw(nconv), cbn(nconv);
        for (int l = 0; l < nconv; ++l) { cw[l] = take((size_t)C * (l == 0 ? ip : C) * 9); cbn[l] = take((size_t)4 * C); }
        const size_t pcw = take((size_t)32 * C), pbn = take(128), pfw = take((size_t)A * feat), pfb = take(A);
        const size_t vcw = take((size_t)32 * C), vbn = take(128), v1w = take((size_t)256 * feat), v1b = take(256), v2w = take(256), v2b = take(1);
        AZ_CHECK(sizeof(Hdr) + off * 4 <= bytes, "weight blob truncated");
        if (w.blocks != nb || w.in_planes != ip) {          // (re)allocate the images for this shape
            free_weights();
            for (int l = 0; l < nconv; ++l) {
                for (int k = 0; k < (l == 0 ? NS : NS * NS); ++k) { __nv_bfloat16* dw; if (dev_alloc(&dw, nn::conv_weight_elems(l == 0 ? cin_pad : nn::CONV_COUT))) return -1; w.conv_w.push_back(dw); }
                for (int k = 0; k < NS; ++k) { float* db; if (dev_alloc(&db, (size_t)nn::CONV_COUT)) return -1; w.conv_b.push_back(db); }
            }
            if (dev_alloc(&w.zero_bias, (size_t)nn::CONV_COUT)) return -1;
            AZ_CUDA_CHECK(cudaMemsetAsync(w.zero_bias, 0, nn::CONV_COUT * 4, st));
            if (dev_alloc(&w.b1x1, 64) || dev_alloc(&w.pfc_b, (size_t)p_tiles * 64) || dev_alloc(&w.vfc1_b, 256) || dev_alloc(&w.vfc2_w, 256) || dev_alloc(&w.vfc2_b, 1) ||
                dev_alloc(&w.g1_w, nn::gemm_weight_elems(64, 3 * C)) || dev_alloc(&w.pfc_img, nn::gemm_weight_elems(p_tiles * 64, (p_split ? 3 : 1) * feat)) ||
                dev_alloc(&w.vfc1_img, nn::gemm_weight_elems(256, 3 * feat)) || dev_alloc(&w.h1_w, nn::head_conv_weight_elems())) return -1;
            if (!p_split && dev_alloc(&w.pfc_rows, (size_t)A * feat)) return -1;
            AZ_CUDA_CHECK(cudaMemsetAsync(w.pfc_img, 0, nn::gemm_weight_elems(p_tiles * 64, (p_split ? 3 : 1) * feat) * 2, st));      // rows >= A stay zero
            AZ_CUDA_CHECK(cudaMemsetAsync(w.pfc_b, 0, (size_t)p_tiles * 64 * 4, st));
            w.blocks = nb; w.in_planes = ip;
        }
        if (w.blob_bytes < off * 4) { cudaFree(w.blob); w.blob = nullptr; if (dev_alloc(&w.blob, off)) return -1; w.blob_bytes = off * 4; }
        AZ_CUDA_CHECK(cudaMemcpyAsync(w.blob, (const char*)blob + sizeof(Hdr), off * 4, cudaMemcpyHostToDevice, st));
        const float* d = w.blob;
        for (int l = 0; l < nconv; ++l) {
            const int cin_total = l == 0 ? ip : C, cin = l == 0 ? cin_pad : nn::CONV_COUT, cin_real = l == 0 ? ip : nn::CONV_COUT;
            const int n = nn::CONV_COUT * cin * 9;
            for (int co = 0; co < NS; ++co)
                for (int ci = 0; ci < (l == 0 ? 1 : NS); ++ci)
                    k_prep_conv<<<(n + 255) / 256, 256, 0, st>>>(d + cw[l], d + cbn[l], w.conv_w[wi(l, co, ci)], ci == 0 ? w.conv_b[bi(l, co)] : nullptr, C, cin_total,
                                                                 co * nn::CONV_COUT, ci * nn::CONV_COUT, cin_real, cin, nn::conv_uses_pair(cin, row_pitch) ? 1 : 0, f16);
        }
        k_prep_1x1<<<(64 * C + 255) / 256, 256, 0, st>>>(d + pcw, d + pbn, d + vcw, d + vbn, w.g1_w, w.b1x1, C, w.h1_w, f16);
        k_prep_fc<<<(unsigned)(((size_t)A * feat + 255) / 256), 256, 0, st>>>(d + pfw, w.pfc_img, A, feat, p_split);
        k_prep_fc<<<(unsigned)(((size_t)256 * feat + 255) / 256), 256, 0, st>>>(d + v1w, w.vfc1_img, 256, feat, 1);
        if (w.pfc_rows) k_prep_fc_rows<<<(unsigned)(((size_t)A * feat + 255) / 256), 256, 0, st>>>(d + pfw, w.pfc_rows, A, feat);
        AZ_CUDA_CHECK(cudaGetLastError());
        AZ_CUDA_CHECK(cudaMemcpyAsync(w.pfc_b, d + pfb, (size_t)A * 4, cudaMemcpyDeviceToDevice, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(w.vfc1_b, d + v1b, 256 * 4, cudaMemcpyDeviceToDevice, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(w.vfc2_w, d + v2w, 256 * 4, cudaMemcpyDeviceToDevice, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(w.vfc2_b, d + v2b, 4, cudaMemcpyDeviceToDevice, st));
        AZ_CUDA_CHECK(cudaStreamSynchronize(st));            // the caller's blob may go away after this call returns
        launches += (size_t)NS + (size_t)(nconv - 1) * NS * NS + 3;
        blocks = nb; in_planes = ip;
        loaded = true;
        return 0;
    }
    void share(const Net& o) { f16 = o.f16; w = o.w; blocks = o.blocks; in_planes = o.in_planes; loaded = o.loaded; owns_weights = false; }
    void free_weights() {
        if (!owns_weights) { w = NetWeights(); loaded = false; return; }
        for (auto p : w.conv_w) cudaFree(p);
        for (auto p : w.conv_b) cudaFree(p);
        w.conv_w.clear(); w.conv_b.clear();
        for (float** p : {&w.b1x1, &w.pfc_b, &w.vfc1_b, &w.vfc2_w, &w.vfc2_b, &w.blob, &w.zero_bias}) { cudaFree(*p); *p = nullptr; }
        w.blob_bytes = 0; w.blocks = -1; w.in_planes = -1;
        for (__nv_bfloat16** p : {&w.g1_w, &w.pfc_img, &w.vfc1_img, &w.h1_w, &w.pfc_rows}) { cudaFree(*p); *p = nullptr; }
        loaded = false;
    }
    void destroy() {
        free_weights();
        for (auto e : {ev_in, ev_out, tev[0], tev[1]}) if (e) cudaEventDestroy(e);
        for (auto& e : fe) if (e) { cudaEventDestroy(e); e = nullptr; }
        ev_in = ev_out = tev[0] = tev[1] = nullptr;
        if (conv_stream) { cudaStreamDestroy(conv_stream); conv_stream = nullptr; }
        for (void* p : {(void*)in16, (void*)X, (void*)Y, (void*)rowvalid, (void*)pooled, (void*)featP, (void*)featV, (void*)logits, (void*)hidden, (void*)logits_part, (void*)hidden_part}) cudaFree(p);
    }
    // in16 (already filled) → policy[n][A], value[n].  n from n_dev (device) or n_fixed.
    // legal / n_legal (device, per evaluation slot) != nullptr on a wide head: logits of the legal moves only (k_policy_legal_value)
    int forward(const int* n_dev, int n_fixed, float* policy, float* value, cudaStream_t s, const int16_t* legal = nullptr, const int32_t* n_legal = nullptr, int legal_pitch = 0,
                const int32_t* slot_tree = nullptr) {
        static const bool alt_order = getenv("AZ_CONV_NO_ALT") == nullptr;      // profiling switch for the alternating item order
        static const bool pdl = getenv("AZ_CONV_NO_PDL") == nullptr;            // profiling switch for programmatic dependent launch of the trunk layers
        static const int conv_dbg = getenv("AZ_CONV_WEIGHTS_FIRST") ? 64 : 0;   // profiling switch: all nine weight taps ahead of the first activation stage
        AZ_CHECK(loaded, "no network weights loaded (az_engine_load_weights)");
        cudaStream_t cs = conv_stream ? conv_stream : s;                          // stream / SM count of the trunk layers
        const int cs_sms = conv_stream ? n_sms_conv : n_sms;
        nn::ConvParams cp{};
        cp.rowvalid = rowvalid; cp.n_boards_dev = n_dev; cp.n_rows = n_fixed * board_pitch; cp.board_pitch = board_pitch;
        cp.p_total = p_total; cp.row_pitch = row_pitch; cp.relu = 1; cp.dbg = conv_dbg; cp.f16 = f16;
        const size_t slice = (size_t)(nn::CONV_COUT / 8) * p_total * 8;          // elements per 128-channel slice of an activation buffer
        for (int co = 0; co < NS; ++co) {                                         // stem: one launch per 128 output channels
            cp.in = in16; cp.out = X + co * slice; cp.resid = nullptr; cp.w = w.conv_w[wi(0, co, 0)]; cp.bias = w.conv_b[bi(0, co)];
            AZ_CHECK(nn::conv3x3_launch(cp, cin_pad, n_sms, s) == 0, "stem conv launch failed"); ++launches;
        }
        // A C -> C layer = NS x NS launches of the 128 -> 128 kernel: output slice co accumulates over the input slices, the first
        // launch adds bias (+ the block's skip connection), the following ones add the partial sum through the residual path (in
        // place: a thread re-reads only the rows it writes), ReLU on the last.  NS = 1 is the plain single launch.
        int slice_launch = 0;
        auto layer = [&](const __nv_bfloat16* in, __nv_bfloat16* out, const __nv_bfloat16* skip, int l, int reverse) -> int {
            for (int co = 0; co < NS; ++co)
                for (int ci = 0; ci < NS; ++ci) {
                    cp.in = in + ci * slice; cp.out = out + co * slice;
                    cp.w = w.conv_w[wi(l, co, ci)]; cp.bias = ci == 0 ? w.conv_b[bi(l, co)] : w.zero_bias;
                    cp.resid = ci == 0 ? (skip ? skip + co * slice : nullptr) : cp.out;
                    cp.relu = ci == NS - 1 ? 1 : 0; cp.pdl = pdl ? 1 : 0;
                    // item order: every launch starts on the rows its predecessor touched last (still in L2).  Slice launches: the partial sum a launch
                    // wrote is the next one's residual, so consecutive launches alternate direction (AZ_CONV_NO_ALT: all forward)
                    cp.reverse = NS == 1 ? reverse : (alt_order ? ((slice_launch++) & 1) : 0);
                    // slice launches: a partial sum (ci < NS - 1) is read back by the next launch — keep it in L2; the activations stream through once
                    static const int hint_bits = getenv("AZ_SLICE_HINTS") ? atoi(getenv("AZ_SLICE_HINTS")) : 3;      // profiling switch: 0 = no hints
                    cp.l2_hints = NS > 1 ? ((ci < NS - 1 ? (hint_bits & 1) : 0) | (hint_bits & 2)) : 0;
                    AZ_CHECK(nn::conv3x3_launch(cp, 128, cs_sms, cs) == 0, "conv launch failed"); ++launches;
                }
            return 0;
        };
        fe_rec(0, s);
        const bool sample = false;      // (live trunk timing now comes from the wave-timing events of EngineT::wave)
        if (conv_stream && blocks > 0) { AZ_CUDA_CHECK(cudaEventRecord(ev_in, s)); AZ_CUDA_CHECK(cudaStreamWaitEvent(cs, ev_in, 0)); }
        if (sample) { for (auto& e : tev) if (!e) cudaEventCreate(&e); cudaEventRecord(tev[0], cs); }
        static const bool trunk_fused = getenv("AZ_TRUNK_LAYERED") == nullptr;    // default: the whole trunk as one persistent launch (k_trunk_pair) where it applies
        if (trunk_fused && NS == 1 && nn::trunk_fused_supported(C, board_pitch, row_pitch, 2 * blocks)) {
            nn::TrunkParams tp{};
            tp.X = X; tp.Y = Y; tp.rowvalid = rowvalid; tp.n_boards_dev = n_dev; tp.n_rows = n_fixed * board_pitch;
            tp.f16 = f16; tp.n_layers = 2 * blocks; tp.p_total = p_total; tp.row_pitch = row_pitch; tp.board_pitch = board_pitch; tp.group_boards = nn::trunk_group_boards(board_pitch);
            // small batches (Go 9x9 at 2048 boards, chess at 1024): X + Y fit in L2 as a whole, so one group per CTA pair balances the pairs better
            // than groups of 7 work items (2048 Go boards = 120 such groups on 74 pairs)
            if (board_pitch != 256 && (size_t)2 * max_boards * board_pitch * 256 <= ((size_t)110 << 20)) { tp.group_boards = std::max(1, (max_boards + cs_sms / 2 - 1) / (cs_sms / 2)); tp.balance = 1; }
            if (const char* d = getenv("AZ_TRUNK_DBG")) tp.dbg = atoi(d);
            static const bool no_discard = getenv("AZ_TRUNK_NO_DISCARD") != nullptr;
            tp.discard = no_discard ? 0 : 1;
            if (const char* d = getenv("AZ_TRUNK_GROUP")) tp.group_boards = std::max(1, atoi(d));           // profiling switch: boards per group
            for (int l = 0; l < 2 * blocks; ++l) { tp.w[l] = w.conv_w[wi(1 + l, 0, 0)]; tp.bias[l] = w.conv_b[bi(1 + l, 0)]; }
            AZ_CHECK(nn::trunk_launch(tp, cs_sms, cs) == 0, "fused trunk launch failed"); ++launches;
        } else
        for (int b = 0; b < blocks; ++b) {
            if (layer(X, Y, nullptr, 1 + 2 * b, alt_order ? 1 : 0)) return -1;
            if (layer(Y, X, X, 2 + 2 * b, 0)) return -1;
        }
        if (conv_stream && blocks > 0) { AZ_CUDA_CHECK(cudaEventRecord(ev_out, cs)); AZ_CUDA_CHECK(cudaStreamWaitEvent(s, ev_out, 0)); }
        if (sample) {
            cudaEventRecord(tev[1], cs); cudaEventSynchronize(tev[1]);
            float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[1]);
            conv_ms += ms; conv_sampled += (unsigned long long)(2 * blocks * NS * NS);
        }
        cp.relu = 1; cp.reverse = 0; cp.pdl = 0;
        fe_rec(1, s);
        // heads: pool → 1x1 convs (GEMM, bf16 features in the FC operand layout) → policy FC / value FC1 (GEMMs, fp32 out)
        static const bool fuse_heads = getenv("AZ_NO_HEAD_FUSION") == nullptr;      // profiling / parity switch: the two-kernel path
        if (fuse_heads && nn::head_conv_supported(C, board_pitch, H, W)) {
            // 1x1 convs on the full-resolution trunk output, pooling in the epilogue (head_conv.cu): one pass over X, no pooled intermediate
            nn::HeadConvParams hp{X, w.h1_w, w.b1x1, n_dev, n_fixed, H, W, row_pitch, board_pitch, p_total, nn::CONV_GUARD, featP, featV, boards_cap, 256, f16, alt_order ? 1 : 0};
            AZ_CHECK(nn::head_conv_launch(hp, n_sms, s) == 0, "fused heads launch failed"); ++launches;
            fe_rec(2, s); fe_rec(3, s);
        } else {
            nn::PoolParams pp{X, pooled, n_dev, n_fixed, C, H, W, row_pitch, board_pitch, p_total, nn::CONV_GUARD, boards_cap, 64 * boards_cap, f16};
            AZ_CHECK(nn::pool_launch(pp, n_sms * 8, s) == 0, "pool launch failed"); ++launches;
            fe_rec(2, s);
            nn::GemmParams g1{}; g1.A = pooled; g1.B = w.g1_w; g1.bias = w.b1x1; g1.a_rows = 64 * boards_cap; g1.a_plane_mod = 2 * (C / 8); g1.K = 3 * C; g1.n_tiles = 1; g1.n_valid = 64;
            g1.units = 64; g1.unit_rows = boards_cap; g1.m_valid_dev = n_dev; g1.m_valid = n_fixed; g1.relu = 1; g1.mode = nn::GEMM_OUT_FEAT;
            g1.out_feat0 = featP; g1.out_feat1 = featV; g1.feat_rows = boards_cap; g1.feat_lo_plane = 256;
            AZ_CHECK(nn::gemm_tc_launch(g1, n_sms, s) == 0, "1x1 conv gemm launch failed"); ++launches;
            fe_rec(3, s);
        }
        // the two FC GEMMs run split-K (fc_splits x more work items: a 4096 x 256 x 6144 GEMM is only 128 tiles); the raw partial
        // sums are added, with bias / ReLU, by k_policy_value
        nn::GemmParams g2{}; g2.A = featP; g2.B = w.pfc_img; g2.bias = w.pfc_b; g2.a_rows = boards_cap; g2.a_plane_mod = 512; g2.K = (p_split ? 3 : 1) * feat; g2.n_tiles = p_tiles; g2.n_valid = A;
        g2.units = 1; g2.unit_rows = 0; g2.m_valid_dev = n_dev; g2.m_valid = n_fixed; g2.relu = 0; g2.mode = nn::GEMM_OUT_ROWS; g2.out_rows = logits_part; g2.ldo = ld_part;
        g2.k_splits = p_splits; g2.split_stride = (size_t)max_boards * ld_part;
        const bool legal_only = legal != nullptr && w.pfc_rows != nullptr && feat == 2048;
        if (!legal_only) { AZ_CHECK(nn::gemm_tc_launch(g2, n_sms, s) == 0, "policy fc gemm launch failed"); ++launches; }
        fe_rec(4, s);
        nn::GemmParams g3 = g2; g3.A = featV; g3.B = w.vfc1_img; g3.bias = w.vfc1_b; g3.K = 3 * feat; g3.n_tiles = 4; g3.n_valid = 256; g3.relu = 1; g3.out_rows = hidden_part; g3.ldo = 256;
        g3.k_splits = fc_splits; g3.split_stride = (size_t)max_boards * 256;
        AZ_CHECK(nn::gemm_tc_launch(g3, n_sms, s) == 0, "value fc gemm launch failed"); ++launches;
        fe_rec(5, s);
        if (legal_only) {
            nn::LegalPolicyParams lp{featP, boards_cap, 256, w.pfc_rows, w.pfc_b, legal, n_legal, legal_pitch, slot_tree, hidden_part, (size_t)max_boards * 256, fc_splits, w.vfc1_b, w.vfc2_w, w.vfc2_b, 256,
                                     policy, value, n_dev, n_fixed, A};
            AZ_CHECK(nn::policy_legal_value_launch(lp, max_boards, s) == 0, "legal policy launch failed"); ++launches;
            fe_rec(6, s);
            return 0;
        }
        nn::OutParams op{logits_part, hidden_part, (size_t)max_boards * ld_part, (size_t)max_boards * 256, p_splits, fc_splits, w.pfc_b, w.vfc1_b, w.vfc2_w, w.vfc2_b, logits, policy, value, n_dev, n_fixed, A, 256, n_dev == nullptr ? 1 : 0, ld_part};
        AZ_CHECK(nn::policy_value_launch(op, max_boards, s) == 0, "output launch failed"); ++launches;
        fe_rec(6, s);
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------ rules replay
// State API on the device (one warp per game): replay a move list, then report what the reference's IGameState would:
// getLegalMoves (in order), isTerminal, getGameResult, getCurrentPlayer, getEnhancedTensorRepresentation.
template <class G>
__global__ void __launch_bounds__(128) k_rules_replay(const int32_t* moves, const int32_t* n_moves, int n_games, int max_moves, int32_t* legal,
                                                     int32_t* n_legal, int32_t* terminal, int32_t* result, int32_t* player, float* planes,
                                                     uint64_t* hist_scratch /*[n_games][max_moves + 1]*/) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n_games) return;
    typename G::Warp& w = warp_ws<G>(smem);
    G::w_init(w, lane);
    G::w_attach_history(w, hist_scratch + (size_t)g * (max_moves + 1), lane);     // initial position + one key per move
    int bad = 0;
    for (int i = 0; i < n_moves[g]; ++i) {
        if (!G::w_apply(w, moves[(size_t)g * max_moves + i], lane, true)) { bad = 1; break; }   // makeMove throws
    }
    const int res = G::w_result(w, lane);
    const int n = G::w_legal(w, lane, legal + (size_t)g * G::MAX_CHILDREN);
    if (lane == 0) { n_legal[g] = bad ? -1 : n; terminal[g] = res != RES_ONGOING; result[g] = res; player[g] = G::w_player(w); }
    if (planes) G::w_planes(w, lane, planes + (size_t)g * G::PLANES * G::CELLS);
}

// createGameState + fresh ParallelMCTS for every slot (self_play_manager.cpp:157-175)
__global__ void k_fill_i32(int32_t* p, int32_t v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

template <class G>
__global__ void __launch_bounds__(128) k_reset_all(TreePools tp, typename G::State* root_state, const int16_t* default_order, int n_order, int16_t* root_order,
                                                  int32_t* root_order_n, int noise, int T, EvalTT tt, long long pool_nodes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    typename G::Warp& w = warp_ws<G>(smem);
    const long long per = pool_nodes / T;               // fresh trees: equal regions
    const size_t base = (size_t)((long long)t * per);
    G::w_init(w, lane);
    G::w_store_root(w, root_state + t, lane);
    if (lane == 0) {
        tp.base[t] = (long long)t * per; tp.limit[t] = (int32_t)per;
        tp.N[base] = 0; tp.W[base] = 0.0f; tp.P[base] = 0.0f; tp.first[base] = -1; tp.sub[base] = 0; tp.act[base] = -1; tp.nchild[base] = 0; tp.flags[base] = 0;
        tp.root[t] = 0; tp.alloc[t] = 1; tp.root_vl[t] = 0; tp.move_num[t] = 0; tp.game_id[t] = 0;
        tp.tflags[t] = (uint8_t)(TF_ACTIVE | (G::FIRST_FILL ? TF_FIRST_FILL : 0) | (noise ? TF_NEED_NOISE : 0));
        root_order_n[t] = n_order;
    }
    for (int i = lane; i < n_order; i += 32) root_order[(size_t)t * G::MAX_CHILDREN + i] = default_order[i];
    if (tt.keys) { for (int i = lane; i < tt.cap; i += 32) tt.keys[(size_t)t * tt.cap + i] = 0; if (lane == 0) tt.count[t] = 0; }
}

// ------------------------------------------------------------------------------------------------ engine
struct EngineBase {
    int device = 0;                 // every C-ABI call makes this the calling thread's current device (engines on several GPUs in one process)
    virtual ~EngineBase() {}
    virtual int set_search_params(float c_puct, int virtual_loss) = 0;
    virtual int set_num_simulations(int sims) = 0;
    virtual int get_timing(az_timing* out) = 0;
    virtual int set_external(az_eval_fn fn, void* user) = 0;
    virtual int node_stats(int slot, const int32_t* path, int n_path, int32_t* actions, int32_t* visits, float* wsum, float* priors, int32_t* n,
                           int32_t* node_visits, float* node_wsum, float* node_prior, int32_t* node_flags) = 0;
    virtual int load_weights(const void* blob, size_t bytes) = 0;
    virtual int reset_games() = 0;
    virtual int set_root(int slot, const int32_t* moves, int n, const int32_t* order, int n_order) = 0;
    virtual int search(int sims) = 0;
    virtual int root_stats(int slot, int32_t* actions, int32_t* visits, float* wsum, float* priors, int32_t* n, int32_t* rn, float* rw) = 0;
    virtual int advance(const int32_t* actions, int n) = 0;
    virtual int play(int n_moves) = 0;
    virtual int add_noise(float alpha, float eps) = 0;
    virtual int last_actions(int32_t* out, int n) = 0;
    virtual int slot_state(int slot, int32_t* result, int32_t* ply, int32_t* player) = 0;
    virtual int sample_layout(az_sample_layout* out) = 0;
    virtual int drain(void* buf, size_t cap, size_t* n, bool device) = 0;
    virtual int get_stats(az_stats* out) = 0;
    virtual int make_examples(const void* samples, size_t n, int augment, float* planes, float* policy, float* value) = 0;
    virtual int examples_from_games(const int32_t* moves, const int32_t* n_moves, const int8_t* results, int n_games, int max_moves, const float* policy_in, int P,
                                    int augment, float* planes, float* policy, float* value) = 0;
    virtual int sync() = 0;
    virtual int nn_forward(const float* planes, int n, float* policy, float* value, float* logits) = 0;
    virtual int nn_bench(int n_boards, int reps, float* ms) = 0;
    virtual int conv_bench(int n_boards, int reps, float* ms) = 0;
    virtual int conv_sampled(double* ms_sum, unsigned long long* launches_n) = 0;
    virtual int event_record(int idx) = 0;
    virtual int event_elapsed(int i, int j, float* ms) = 0;
    virtual int rules_replay(const int32_t* moves, const int32_t* n_moves, int n_games, int max_moves, int32_t* legal, int32_t* n_legal,
                             int32_t* terminal, int32_t* result, int32_t* player, float* planes) = 0;
};

template <class G>
struct EngineT : EngineBase {
    using State = typename G::State;
    using Leaf = typename G::Leaf;
    using SampleT = Sample<G>;
    static constexpr int A = G::ACTIONS;          // policy length
    static constexpr int MC = G::MAX_CHILDREN;
    // Slots are split into `NG` stream groups; each group runs its own wave sequence (select → network → expand) on
    // its own stream with its own wave / activation buffers, so the tree kernels of one group overlap the tensor-core
    // pass of the other.  Move-commit kernels run once for all slots on the main stream.
    struct Group { int t0 = 0, n = 0; cudaStream_t stream = nullptr; cudaEvent_t ev = nullptr; WaveBuffers wb{}; TreePools tp{}; EvalTT tt{}; EvalCache ec{}; Net net;
                   // the simulation wave as a CUDA graph, one instance per pool-buffer parity (the node arrays swap at every move commit); replays add `n` launches
                   struct WaveGraph { cudaGraphExec_t exec = nullptr; const void* key = nullptr; unsigned long long n = 0, n_net = 0; } wg[2]; };
    // AZ_EVAL_EXTERNAL: the caller's evaluator + staging (device: leaf paths by evaluation slot; host: the same + its answers)
    az_eval_fn ext_fn = nullptr; void* ext_user = nullptr;
    int32_t *ext_paths = nullptr, *ext_plen = nullptr, *ext_slot_tree = nullptr;
    std::vector<int32_t> h_paths, h_plen, h_slot; std::vector<float> h_policy, h_value;
    DupStats dup{};                               // AZ_EVAL_DUP_STATS (profiling): see tree.cuh
    EvalTT tt{};                                  // model of the reference's TranspositionTable (chess + hash evaluators, tree.cuh)
    bool hash_eval() const { return cfg.evaluator == AZ_EVAL_HASH || cfg.evaluator == AZ_EVAL_HASH_PEAKED; }
    az_config cfg;
    int T = 0, NG = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_main = nullptr;
    std::vector<Group> groups;
    SmPartitions parts; bool partitioned = false;
    TreePools tp{};                               // current pool buffer + per-tree arrays
    TreePools alt{};                              // the other pool buffer (node arrays, base, limit); per-tree arrays shared with tp
    long long pool_nodes = 0;                     // nodes per pool buffer
    int32_t* kept = nullptr; RegionPlan* plan = nullptr;
    int budget_sims = 0, sims_since_plan = 0;     // simulations every region is provisioned for / run since the regions were cut
    unsigned long long overflows_seen = 0;
    State* root_state = nullptr; Leaf* leaf_state = nullptr;
    int16_t *root_order = nullptr, *default_order = nullptr; int32_t* root_order_n = nullptr;
    int32_t *chosen_child = nullptr, *chosen_action = nullptr, *forced = nullptr;
    SampleT *game_buf = nullptr, *ring = nullptr; int32_t* ring_count = nullptr; int ring_cap = 0; int max_moves = 0;
    float* noise_scratch = nullptr;
    Stats* dstats = nullptr;
    unsigned long long launches = 0, waves = 0;
    std::vector<int16_t> h_default_order;

    ~EngineT() override { cudaSetDevice(device); destroy(); }

    void destroy() {
        cudaDeviceSynchronize();
        if (dup.counters) {
            unsigned long long c[8] = {}; cudaMemcpy(c, dup.counters, 64, cudaMemcpyDeviceToHost);
            const double ev = c[0] ? (double)c[0] : 1.0;
            fprintf(stderr, "az eval-duplicate stats: %llu leaf evaluations; exact network input: %.3f %% duplicate inside their wave, %.3f %% seen before in the run; "
                            "reference TT key: %.3f %% / %.3f %%; run-table inserts refused: %llu\n", c[0], 100.0 * c[1] / ev, 100.0 * c[2] / ev, 100.0 * c[3] / ev, 100.0 * c[4] / ev, c[5]);
            for (void* p : {(void*)dup.wave_keys, (void*)dup.run_keys, (void*)dup.wave_keys_ref, (void*)dup.run_keys_ref, (void*)dup.counters}) cudaFree(p);
            dup = DupStats{};
        }
        if (wave_timing && wt_n) {
            fprintf(stderr, "az wave timing over %ld sampled waves: select %.3f ms, dedup + encode %.3f ms, evaluator %.3f ms, expand/backup %.3f ms; move commit %.3f ms over %ld moves\n", wt_n, wt_ms[0] / wt_n,
                    wt_ms[1] / wt_n, wt_ms[2] / wt_n, wt_ms[3] / wt_n, cm_n ? cm_ms / cm_n : 0.0, cm_n);
            fprintf(stderr, "  evaluator: stem %.3f, trunk %.3f, pool (or fused 1x1 conv + pool) %.3f, 1x1 gemm %.3f, policy fc %.3f, value fc %.3f, softmax/tanh %.3f ms\n", wt_fwd[0] / wt_n, wt_fwd[1] / wt_n,
                    wt_fwd[2] / wt_n, wt_fwd[3] / wt_n, wt_fwd[4] / wt_n, wt_fwd[5] / wt_n, wt_fwd[6] / wt_n);
        }
        for (auto& g : groups) {
            for (void* p : {(void*)g.wb.path, (void*)g.wb.path_len, (void*)g.wb.leaf_node, (void*)g.wb.leaf_kind, (void*)g.wb.leaf_value, (void*)g.wb.policy,
                            (void*)g.wb.value, (void*)g.wb.eval_slot, (void*)g.wb.n_eval, (void*)g.wb.eval_key, (void*)g.wb.legal, (void*)g.wb.n_legal, (void*)g.wb.slot_tree, (void*)g.wb.dd_keys, (void*)g.wb.dd_owner, (void*)g.wb.dd_idx, (void*)g.wb.cache_entry,
                            (void*)g.ec.keys, (void*)g.ec.stamp, (void*)g.ec.value, (void*)g.ec.policy, (void*)g.ec.wave}) cudaFree(p);
            drop_graphs(g);
            g.net.destroy();
            if (g.ev) cudaEventDestroy(g.ev);
            if (g.stream) cudaStreamDestroy(g.stream);
        }
        groups.clear();
        parts.destroy(); partitioned = false;
        if (ev_main) { cudaEventDestroy(ev_main); ev_main = nullptr; }
        for (void* p : {(void*)tp.N, (void*)tp.W, (void*)tp.P, (void*)tp.first, (void*)tp.sub, (void*)tp.act, (void*)tp.nchild, (void*)tp.flags, (void*)tp.base, (void*)tp.limit,
                        (void*)tp.root, (void*)tp.alloc, (void*)tp.root_vl, (void*)tp.tflags, (void*)tp.move_num, (void*)tp.game_id,
                        (void*)alt.N, (void*)alt.W, (void*)alt.P, (void*)alt.first, (void*)alt.sub, (void*)alt.act, (void*)alt.nchild, (void*)alt.flags, (void*)alt.base, (void*)alt.limit,
                        (void*)kept, (void*)plan,
                        (void*)root_state, (void*)leaf_state, (void*)root_order, (void*)default_order, (void*)root_order_n, (void*)chosen_child,
                        (void*)chosen_action, (void*)forced, (void*)game_buf, (void*)ring, (void*)ring_count, (void*)noise_scratch, (void*)dstats,
                        (void*)tt.keys, (void*)tt.vals, (void*)tt.count, (void*)ext_paths, (void*)ext_plen, (void*)ext_slot_tree})
            cudaFree(p);
        if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    }

    int init(const az_config& c) {
        cfg = c; T = c.n_slots; device = c.device;
        AZ_CUDA_CHECK(cudaSetDevice(c.device));
        cudaDeviceProp prop; AZ_CUDA_CHECK(cudaGetDeviceProperties(&prop, c.device));
        AZ_CHECK(prop.major >= 10, "az_b200 needs an sm_100-class GPU (no fallback path exists)");
        AZ_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        // Node pool (tree.cuh): two buffers of pool_nodes nodes (25 bytes each) shared by all trees.  max_nodes_per_tree is the AVERAGE per
        // tree; default: what fits in ~55 % of the free device memory, between 1.25 x and 4 x one search's worst-case growth.
        budget_sims = std::max(c.num_simulations, 1);
        const long long growth = (long long)(budget_sims + 1) * MC + 1;
        long long per_tree = c.max_nodes_per_tree;
        if (per_tree <= 0) {
            size_t free_b = 0, total_b = 0; AZ_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
            per_tree = (long long)(0.55 * (double)free_b / (50.0 * T));
            per_tree = std::max(growth * 5 / 4, std::min(growth * 4, per_tree));
        }
        AZ_CHECK(per_tree >= MC + 1, "max_nodes_per_tree must hold at least one expansion");
        pool_nodes = per_tree * T;
        const size_t tn = (size_t)pool_nodes;
        for (TreePools* b : {&tp, &alt})
            if (dev_alloc(&b->N, tn) || dev_alloc(&b->W, tn) || dev_alloc(&b->P, tn) || dev_alloc(&b->first, tn) || dev_alloc(&b->sub, tn) || dev_alloc(&b->act, tn) ||
                dev_alloc(&b->nchild, tn) || dev_alloc(&b->flags, tn) || dev_alloc(&b->base, T) || dev_alloc(&b->limit, T)) return -1;
        if (dev_alloc(&tp.root, T) || dev_alloc(&tp.alloc, T) || dev_alloc(&tp.root_vl, T) ||
            dev_alloc(&tp.tflags, T) || dev_alloc(&tp.move_num, T) || dev_alloc(&tp.game_id, T) || dev_alloc(&kept, T) || dev_alloc(&plan, 1)) return -1;
        alt.root = tp.root; alt.alloc = tp.alloc; alt.root_vl = tp.root_vl; alt.tflags = tp.tflags; alt.move_num = tp.move_num; alt.game_id = tp.game_id;
        if (dev_alloc(&root_state, T) || dev_alloc(&leaf_state, T) || dev_alloc(&root_order, (size_t)T * MC) || dev_alloc(&default_order, MC) ||
            dev_alloc(&root_order_n, T) || dev_alloc(&chosen_child, T) || dev_alloc(&chosen_action, T) || dev_alloc(&forced, T)) return -1;
        max_moves = G::MAX_GAME_MOVES;
        ring_cap = c.sample_ring_capacity > 0 ? c.sample_ring_capacity : std::max(32 * T, 4096);    // a finished game appends up to MAX_GAME_MOVES records at once
        if (dev_alloc(&game_buf, (size_t)T * max_moves) || dev_alloc(&ring, (size_t)ring_cap) || dev_alloc(&ring_count, 1)) return -1;
        AZ_CUDA_CHECK(cudaMemset(ring_count, 0, 4));
        if (dev_alloc(&noise_scratch, (size_t)T * MC) || dev_alloc(&dstats, 1)) return -1;
        AZ_CUDA_CHECK(cudaMemset(dstats, 0, sizeof(Stats)));
        // QUIRK G2: legal-move order of the first-ever enumeration of a fresh state = iteration order of a
        // libstdc++ std::unordered_set<int> filled with 0..A-1 ascending (include/alphazero/games/gomoku/
        // gomoku_state.h:124, gomoku_state.cpp:531-566).  Computed with the real container, uploaded once.
        if (G::FIRST_FILL) {
            std::unordered_set<int> us; for (int a = 0; a < G::CELLS; ++a) us.insert(a);
            h_default_order.assign(us.begin(), us.end());
            AZ_CUDA_CHECK(cudaMemcpy(default_order, h_default_order.data(), h_default_order.size() * 2, cudaMemcpyHostToDevice));
        }
        if (G::TT_COARSE && hash_eval() && c.tt_entries >= 0) {
            // per game: one entry per evaluation; sized for 64 moves' worth of simulations unless the caller says otherwise (inserts stop at half load)
            long want = c.tt_entries > 0 ? (long)c.tt_entries : 2L * 64 * (std::max(c.num_simulations, 1) + 1);
            int capn = 1024; while (capn < want && capn < (1 << 22)) capn <<= 1;
            tt.cap = capn;
            if (dev_alloc(&tt.keys, (size_t)T * capn) || dev_alloc(&tt.vals, (size_t)T * capn) || dev_alloc(&tt.count, T)) return -1;
        }
        if (c.evaluator == AZ_EVAL_EXTERNAL) {
            if (dev_alloc(&ext_paths, (size_t)T * MAX_DEPTH) || dev_alloc(&ext_plen, T) || dev_alloc(&ext_slot_tree, T)) return -1;
            h_paths.resize((size_t)T * MAX_DEPTH); h_plen.resize(T); h_slot.resize(T); h_policy.resize((size_t)T * A); h_value.resize(T);
        }
        if (getenv("AZ_EVAL_DUP_STATS")) {
            dup.wave_mask = (1u << 20) - 1; dup.run_mask = (1u << 28) - 1;                 // 8 MB per wave set, 2 GB per run set
            if (dev_alloc(&dup.wave_keys, (size_t)dup.wave_mask + 1) || dev_alloc(&dup.wave_keys_ref, (size_t)dup.wave_mask + 1) ||
                dev_alloc(&dup.run_keys, (size_t)dup.run_mask + 1) || dev_alloc(&dup.run_keys_ref, (size_t)dup.run_mask + 1) || dev_alloc(&dup.counters, 8)) return -1;
            AZ_CUDA_CHECK(cudaMemset(dup.run_keys, 0, ((size_t)dup.run_mask + 1) * 8)); AZ_CUDA_CHECK(cudaMemset(dup.run_keys_ref, 0, ((size_t)dup.run_mask + 1) * 8));
            AZ_CUDA_CHECK(cudaMemset(dup.counters, 0, 64));
        }
        AZ_CUDA_CHECK(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming));
        NG = std::max(1, std::min(c.n_streams > 0 ? c.n_streams : 1, T));
        if (c.evaluator == AZ_EVAL_EXTERNAL) NG = 1;
        const int per = (T + NG - 1) / NG;
        NG = (T + per - 1) / per;                       // groups of `per` slots; a count that does not divide T leaves no empty trailing group
        groups.resize(NG);
        // SM partitions: AZ_SM_SPLIT=<SMs of the conv partition> (0 / unset: off); only useful with >= 2 stream groups and the network evaluator
        if (const char* sp = getenv("AZ_SM_SPLIT")) {
            const int want = atoi(sp);
            if (want > 0 && c.evaluator == AZ_EVAL_RESNET) { if (parts.create(c.device, want)) return -1; partitioned = true; }
        }
        for (int gi = 0; gi < NG; ++gi) {
            Group& g = groups[gi];
            g.t0 = gi * per; g.n = std::min(per, T - g.t0);
            if (partitioned) { if (parts.stream(false, &g.stream)) return -1; }
            else AZ_CUDA_CHECK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
            AZ_CUDA_CHECK(cudaEventCreateWithFlags(&g.ev, cudaEventDisableTiming));
            const int n = g.n;
            if (dev_alloc(&g.wb.path, (size_t)n * MAX_DEPTH) || dev_alloc(&g.wb.path_len, n) || dev_alloc(&g.wb.leaf_node, n) || dev_alloc(&g.wb.leaf_kind, n) ||
                dev_alloc(&g.wb.leaf_value, n) || dev_alloc(&g.wb.policy, (size_t)n * A) || dev_alloc(&g.wb.value, n) || dev_alloc(&g.wb.eval_slot, n) ||
                dev_alloc(&g.wb.n_eval, 1)) return -1;
            if (G::LEGAL_POLICY && c.evaluator == AZ_EVAL_RESNET && !c.dense_policy) {
                if (dev_alloc(&g.wb.legal, (size_t)n * MC) || dev_alloc(&g.wb.n_legal, n)) return -1;
            }
            // evaluation cache across waves (EvalCache, tree.cuh): on by default behind the ResNet evaluator; behind a hash evaluator only when asked
            // for (parity runs: the search must equal the reference's with the cache on) and not next to the chess table model, whose
            // evaluations are not a function of the leaf's own input
            const bool hash_cache = hash_eval() && c.eval_cache_entries > 0 && !tt.keys;
            const bool want_cache = c.eval_dedup >= 0 && c.eval_cache_entries >= 0 && (c.evaluator == AZ_EVAL_RESNET || hash_cache);
            if (c.evaluator == AZ_EVAL_RESNET && dev_alloc(&g.wb.slot_tree, n)) return -1;
            if ((c.evaluator == AZ_EVAL_RESNET && c.eval_dedup >= 0) || want_cache) {
                unsigned int cap_dd = 1024; while (cap_dd < 4u * (unsigned)n) cap_dd <<= 1;
                g.wb.dd_mask = cap_dd - 1;
                if (dev_alloc(&g.wb.dd_keys, cap_dd) || dev_alloc(&g.wb.dd_owner, cap_dd) || dev_alloc(&g.wb.dd_idx, n)) return -1;
            }
            if (want_cache) {
                // entries are split over the groups (Gomoku 15x15: 916 B per entry, 4 M = 3.8 GB); never more than ~1/16 of the device memory
                const int pw = G::LEGAL_POLICY ? MC : A;
                size_t free_b = 0, total_b = 0; AZ_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
                // default: room for four moves' worth of evaluations of every slot, between 64 K and 4 M entries (chess: 8.8 / 11.5 / 13.0 / 13.1 % hits at
                // 256 K / 1 M / 4 M / 16 M entries in moves 6-11 of the bench)
                const long long four_moves = 4LL * T * (std::max(c.num_simulations, 1) + 1);
                long long want = c.eval_cache_entries > 0 ? (long long)c.eval_cache_entries
                                                          : std::max<long long>(1 << 16, std::min<long long>(four_moves, (long long)std::min<size_t>((size_t)1 << 22, total_b / 16 / (size_t)(16 + 4 * pw))));
                want = std::max<long long>(want / NG, 64);
                unsigned int capn = 64; while ((long long)capn * 2 <= want && capn < (1u << 30)) capn <<= 1;
                // the cache is an optimisation: on a device that is nearly full (a caller-sized node pool) it shrinks, down to nothing, instead of failing the engine
                while (capn > 64 && (size_t)capn * (16 + 4 * (size_t)pw) + ((size_t)1 << 30) > free_b) capn >>= 1;
                if ((size_t)capn * (16 + 4 * (size_t)pw) + ((size_t)1 << 30) <= free_b) {
                    g.ec.mask = capn - 1; g.ec.pw = pw;
                    if (dev_alloc(&g.ec.wave, 2)) return -1;
                    { const uint32_t w0[2] = {1u, 1u}; AZ_CUDA_CHECK(cudaMemcpy(g.ec.wave, w0, 8, cudaMemcpyHostToDevice)); }
                    if (dev_alloc(&g.ec.keys, capn) || dev_alloc(&g.ec.stamp, capn) || dev_alloc(&g.ec.value, capn) || dev_alloc(&g.ec.policy, (size_t)capn * pw) ||
                        dev_alloc(&g.wb.cache_entry, n)) return -1;
                    AZ_CUDA_CHECK(cudaMemset(g.ec.keys, 0, (size_t)capn * 8)); AZ_CUDA_CHECK(cudaMemset(g.ec.stamp, 0, (size_t)capn * 4));
                }
            }
            if (tt.keys) {
                if (dev_alloc(&g.wb.eval_key, n)) return -1;
                g.tt = tt; g.tt.keys += (size_t)g.t0 * tt.cap; g.tt.vals += (size_t)g.t0 * tt.cap; g.tt.count += g.t0;
            }
            refresh_group_view(g);
            if (c.evaluator == AZ_EVAL_RESNET) {
                g.net.f16 = c.net_precision == AZ_NET_BF16 ? 0 : 1;
                if (g.net.init(G::N, G::N, A, per, c.net_channels, G::PLANES)) return -1;
                if (partitioned) {
                    if (parts.stream(true, &g.net.conv_stream)) return -1;
                    AZ_CUDA_CHECK(cudaEventCreateWithFlags(&g.net.ev_in, cudaEventDisableTiming));
                    AZ_CUDA_CHECK(cudaEventCreateWithFlags(&g.net.ev_out, cudaEventDisableTiming));
                    g.net.n_sms = parts.small_sms; g.net.n_sms_conv = parts.conv_sms;
                }
            }
        }
        return reset_games();
    }

    // view of the pools for a group's slots: per-tree arrays start at the group's first slot (node arrays are addressed through base[])
    void refresh_group_view(Group& g) {
        g.tp = tp;
        g.tp.base += g.t0; g.tp.limit += g.t0;
        g.tp.root += g.t0; g.tp.alloc += g.t0; g.tp.root_vl += g.t0; g.tp.tflags += g.t0; g.tp.move_num += g.t0; g.tp.game_id += g.t0;
    }
    SearchParams sparams() const { return SearchParams{cfg.c_puct, cfg.virtual_loss, MAX_DEPTH}; }
    int blocks_for_warps(int n) const { return (n * 32 + 127) / 128; }

    // main stream waits for every group stream (and vice versa): brackets the per-move kernels and host reads
    int join_groups() {
        for (auto& g : groups) { AZ_CUDA_CHECK(cudaEventRecord(g.ev, g.stream)); AZ_CUDA_CHECK(cudaStreamWaitEvent(stream, g.ev, 0)); }
        return 0;
    }
    int fork_groups() {
        AZ_CUDA_CHECK(cudaEventRecord(ev_main, stream));
        for (auto& g : groups) AZ_CUDA_CHECK(cudaStreamWaitEvent(g.stream, ev_main, 0));
        return 0;
    }
    int sync_all() {
        for (auto& g : groups) AZ_CUDA_CHECK(cudaStreamSynchronize(g.stream));
        AZ_CUDA_CHECK(cudaStreamSynchronize(stream));
        return 0;
    }
    unsigned long long net_launches() const { unsigned long long n = 0; for (auto& g : groups) n += g.net.launches; return n; }

    int load_weights(const void* blob, size_t bytes) override {
        AZ_CHECK(cfg.evaluator == AZ_EVAL_RESNET, "engine was created with the hash evaluator");
        if (sync_all()) return -1;
        if (groups[0].net.load(blob, bytes, stream)) return -1;
        for (size_t gi = 1; gi < groups.size(); ++gi) groups[gi].net.share(groups[0].net);
        for (auto& g : groups) drop_graphs(g);                       // weight images may have moved
        for (auto& g : groups)      // a new network: cached evaluations of the old one are void
            if (g.ec.keys) { AZ_CUDA_CHECK(cudaMemsetAsync(g.ec.keys, 0, ((size_t)g.ec.mask + 1) * 8, stream)); AZ_CUDA_CHECK(cudaStreamSynchronize(stream)); }
        return 0;
    }

    int write_fresh(int slot, const State& s, const int16_t* order, int n_order, bool first_fill) {
        if (sync_all()) return -1;
        long long base_ll = 0; AZ_CUDA_CHECK(cudaMemcpy(&base_ll, tp.base + slot, 8, cudaMemcpyDeviceToHost));
        const size_t base = (size_t)base_ll;
        const int res = G::host_root_result(s);
        int32_t zero = 0, one = 1, m1 = -1; float fz = 0.0f; int16_t a16 = -1, z16 = 0;
        uint8_t nf = res != RES_ONGOING ? (uint8_t)(NF_TERMINAL | (res << NF_RESULT_SHIFT)) : 0;
        uint8_t tf = TF_ACTIVE | (first_fill ? TF_FIRST_FILL : 0) | (res != RES_ONGOING ? TF_GAME_OVER : 0) | (cfg.deterministic ? 0 : TF_NEED_NOISE);
        uint32_t gid = 0;
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.N + base, &zero, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.W + base, &fz, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.P + base, &fz, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.first + base, &m1, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.sub + base, &zero, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.act + base, &a16, 2, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.nchild + base, &z16, 2, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.flags + base, &nf, 1, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.root + slot, &zero, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.alloc + slot, &one, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.root_vl + slot, &zero, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.tflags + slot, &tf, 1, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.move_num + slot, &zero, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(tp.game_id + slot, &gid, 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(root_state + slot, &s, sizeof(State), cudaMemcpyHostToDevice, stream));
        if (tt.keys) {      // ParallelMCTS ctor: a table of its own
            AZ_CUDA_CHECK(cudaMemsetAsync(tt.keys + (size_t)slot * tt.cap, 0, (size_t)tt.cap * 8, stream));
            AZ_CUDA_CHECK(cudaMemsetAsync(tt.count + slot, 0, 4, stream));
        }
        if (order && n_order > 0) AZ_CUDA_CHECK(cudaMemcpyAsync(root_order + (size_t)slot * MC, order, (size_t)n_order * 2, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(root_order_n + slot, &n_order, 4, cudaMemcpyHostToDevice, stream));
        if (sync_all()) return -1;   // host temporaries above go out of scope
        return 0;
    }

    int reset_games() override {
        k_reset_all<G><<<blocks_for_warps(T), 128, warp_ws_bytes<G>(), stream>>>(tp, root_state, default_order, G::FIRST_FILL ? G::CELLS : 0, root_order, root_order_n,
                                                                                 cfg.deterministic ? 0 : 1, T, tt, pool_nodes);
        AZ_LAUNCH_CHECK(); ++launches;
        sims_since_plan = 0; budget_sims = (int)std::min<long long>(std::max(cfg.num_simulations, 1), (pool_nodes / T - 1) / MC - 1);
        AZ_CUDA_CHECK(cudaMemsetAsync(ring_count, 0, 4, stream));
        if (sync_all()) return -1;
        return 0;
    }

    int set_root(int slot, const int32_t* moves, int n, const int32_t* order, int n_order) override {
        AZ_CHECK(slot >= 0 && slot < T, "slot out of range");
        std::unique_ptr<State> sp(new State()); State& s = *sp; G::init(s);
        for (int i = 0; i < n; ++i) AZ_CHECK(G::host_apply(s, moves[i]), "illegal move in az_engine_set_root");   // IllegalMove (igamestate.h:36-52)
        AZ_CHECK(!order || G::FIRST_FILL, "a root child order can only be given for Gomoku (QUIRK G2)");
        std::vector<int16_t> ord;
        if (order) {
            // the caller's first-fill order must be exactly the root's legal moves, each once (k_expand_backup trusts it)
            AZ_CHECK(n_order >= 0 && n_order <= MC, "az_engine_set_root: first_fill_order longer than the action space");
            std::vector<char> seen(G::CELLS, 0);
            int n_empty = 0;
            for (int a = 0; a < G::CELLS; ++a) { State t = s; if (G::host_apply(t, a)) ++n_empty; }
            for (int i = 0; i < n_order; ++i) {
                const int a = order[i];
                AZ_CHECK(a >= 0 && a < G::CELLS && !seen[a], "az_engine_set_root: first_fill_order entry out of range or repeated");
                State t = s; AZ_CHECK(G::host_apply(t, a), "az_engine_set_root: first_fill_order names an occupied cell");
                seen[a] = 1;
            }
            AZ_CHECK(n_order == n_empty, "az_engine_set_root: first_fill_order must list every legal move of the root once");
            ord.resize(n_order); for (int i = 0; i < n_order; ++i) ord[i] = (int16_t)order[i];
        }
        return write_fresh(slot, s, order ? ord.data() : nullptr, order ? n_order : 0, order != nullptr);
    }

    // one wave of one group: select → evaluator → expand/backup, on the group's stream
    // AZ_WAVE_TIMING=1 (profiling): CUDA events around select / evaluator / expand of every 64th wave, printed at destroy
    // Every 64th wave of a ResNet engine is bracketed kernel by kernel with CUDA events on its stream (one host sync per 64 waves): the
    // per-step budget bench.py reports (az_engine_get_timing) and the live trunk time of its roofline.  AZ_WAVE_TIMING=1 also prints it.
    bool wave_timing = getenv("AZ_WAVE_TIMING") != nullptr;
    cudaEvent_t wt_ev[5] = {}; double wt_ms[4] = {0, 0, 0, 0}, wt_fwd[7] = {0, 0, 0, 0, 0, 0, 0}; long wt_n = 0, wt_seen = 0;
    cudaEvent_t cm_ev[2] = {}; bool cm_pending = false; double cm_ms = 0; long cm_n = 0;
    void resolve_commit_timing() {
        if (!cm_pending || cudaEventQuery(cm_ev[1]) != cudaSuccess) return;
        float ms = 0; if (cudaEventElapsedTime(&ms, cm_ev[0], cm_ev[1]) == cudaSuccess) { cm_ms += ms; ++cm_n; }
        cm_pending = false;
    }
    // The simulation wave (mode 0) is the same launch sequence 800 times per move — its sizes travel in device counters — so it is captured
    // once per pool-buffer parity into a CUDA graph and replayed (AZ_NO_WAVE_GRAPH=1: plain launches).  Root-expansion waves, the sampled
    // (event-bracketed) waves, the external evaluator and the profiling modes launch directly.
    bool use_graphs = getenv("AZ_NO_WAVE_GRAPH") == nullptr;
    void drop_graphs(Group& g) { for (auto& w : g.wg) { if (w.exec) cudaGraphExecDestroy(w.exec); w = typename Group::WaveGraph{}; } }
    int wave(Group& g, int mode) {
        const bool timed = (wave_timing || cfg.evaluator == AZ_EVAL_RESNET) && (wt_seen++ % 64) == 63;
        if (!use_graphs || mode != 0 || timed || cfg.evaluator == AZ_EVAL_EXTERNAL || dup.counters != nullptr) return enqueue_wave(g, mode, timed);
        typename Group::WaveGraph* w = nullptr;
        for (auto& c : g.wg) if (c.exec && c.key == (const void*)g.tp.N) w = &c;
        if (!w) {
            w = g.wg[0].exec ? &g.wg[1] : &g.wg[0];
            if (w->exec) { cudaGraphExecDestroy(w->exec); *w = typename Group::WaveGraph{}; }
            const unsigned long long l0 = launches, n0 = g.net.launches;
            cudaGraph_t graph = nullptr;
            AZ_CUDA_CHECK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeRelaxed));
            const int rc = enqueue_wave(g, 0, false);
            const cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
            if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) set_error(std::string("wave graph capture failed: ") + cudaGetErrorString(ce)); return -1; }
            const cudaError_t ie = cudaGraphInstantiate(&w->exec, graph, 0);
            cudaGraphDestroy(graph);
            AZ_CUDA_CHECK(ie);
            w->key = (const void*)g.tp.N; w->n = launches - l0; w->n_net = g.net.launches - n0;
            launches = l0; g.net.launches = n0;                      // the capture launched nothing
        }
        AZ_CUDA_CHECK(cudaGraphLaunch(w->exec, g.stream));
        launches += w->n; g.net.launches += w->n_net;
        return 0;
    }
    int enqueue_wave(Group& g, int mode, bool timed) {
        cudaStream_t st = g.stream;
        if (timed) { for (auto& e : wt_ev) if (!e) cudaEventCreate(&e); cudaEventRecord(wt_ev[0], st); }
        AZ_CUDA_CHECK(cudaMemsetAsync(g.wb.n_eval, 0, 4, st));
        if (g.wb.dd_keys) {
            AZ_CUDA_CHECK(cudaMemsetAsync(g.wb.dd_keys, 0, ((size_t)g.wb.dd_mask + 1) * 8, st));
            AZ_CUDA_CHECK(cudaMemsetAsync(g.wb.dd_owner, 0x7f, ((size_t)g.wb.dd_mask + 1) * 4, st));      // 0x7f7f7f7f > any tree index
        }
        if (dup.counters) {
            AZ_CUDA_CHECK(cudaMemsetAsync(dup.wave_keys, 0, ((size_t)dup.wave_mask + 1) * 8, st));
            AZ_CUDA_CHECK(cudaMemsetAsync(dup.wave_keys_ref, 0, ((size_t)dup.wave_mask + 1) * 8, st));
        }
        typename G::EncTarget enc{nullptr, 0, 0, 0, 0};
        if (cfg.evaluator == AZ_EVAL_RESNET) enc = typename G::EncTarget{g.net.in16, g.net.p_total, nn::CONV_GUARD, g.net.board_pitch, g.net.f16};
        k_select<G><<<blocks_for_warps(g.n), 128, warp_ws_bytes<G>(), st>>>(g.tp, root_state + g.t0, leaf_state + g.t0, g.wb, sparams(), enc, g.tt, g.n, mode, dup, g.ec);
        AZ_LAUNCH_CHECK(); ++launches;
        if (timed) cudaEventRecord(wt_ev[1], st);
        if (cfg.evaluator == AZ_EVAL_EXTERNAL) {
            if (external_eval(g)) return -1;
        } else if (hash_eval()) {
            if (g.wb.dd_keys) {
                k_dedup_encode<G><<<blocks_for_warps(g.n), 128, warp_ws_bytes<G>(), st>>>(leaf_state + g.t0, root_state + g.t0, g.wb, enc, g.n, dstats, g.ec.wave);
                AZ_LAUNCH_CHECK(); ++launches;
            }
            k_hash_eval<G><<<blocks_for_warps(g.n), 128, warp_ws_bytes<G>((A < HASH_EVAL_CHUNK ? A : HASH_EVAL_CHUNK) * 4), st>>>(leaf_state + g.t0, root_state + g.t0, g.wb, g.n,
                                                                                                                              cfg.evaluator == AZ_EVAL_HASH_PEAKED ? 1 : 0);
            AZ_LAUNCH_CHECK(); ++launches;
        } else {
            g.net.fe_on = timed;
            if (g.wb.dd_keys) {
                k_dedup_encode<G><<<blocks_for_warps(g.n), 128, warp_ws_bytes<G>(), st>>>(leaf_state + g.t0, root_state + g.t0, g.wb, enc, g.n, dstats, g.ec.wave);
                AZ_LAUNCH_CHECK(); ++launches;
            }
            if (timed) cudaEventRecord(wt_ev[2], st);
            if (g.net.forward(g.wb.n_eval, 0, g.wb.policy, g.wb.value, st, g.wb.legal, g.wb.n_legal, MC, g.wb.slot_tree)) return -1;
            g.net.fe_on = false;
        }
        if (timed) { if (hash_eval()) cudaEventRecord(wt_ev[2], st); cudaEventRecord(wt_ev[3], st); }
        k_expand_backup<G><<<blocks_for_warps(g.n), 128, warp_ws_bytes<G>(MC * 4 + (MC * 2 + 15) / 16 * 16), st>>>(g.tp, leaf_state + g.t0, root_state + g.t0, g.wb, root_order + (size_t)g.t0 * MC,
                                                                                    root_order_n + g.t0, sparams(), g.n, dstats, g.ec);
        AZ_LAUNCH_CHECK(); ++launches;
        if (timed) {
            cudaEventRecord(wt_ev[4], st); cudaEventSynchronize(wt_ev[4]);
            for (int i = 0; i < 4; ++i) { float ms = 0; cudaEventElapsedTime(&ms, wt_ev[i], wt_ev[i + 1]); wt_ms[i] += ms; }     // select, dedup + encode, evaluator, expand
            if (cfg.evaluator == AZ_EVAL_RESNET && g.net.fe[6]) {       // dedup end → stem → trunk → pool → 1x1 → policy FC → value FC → softmax
                cudaEvent_t seq[8] = {wt_ev[2], g.net.fe[0], g.net.fe[1], g.net.fe[2], g.net.fe[3], g.net.fe[4], g.net.fe[5], g.net.fe[6]};
                for (int i = 0; i < 7; ++i) { float ms = 0; cudaEventElapsedTime(&ms, seq[i], seq[i + 1]); wt_fwd[i] += ms; }
                float tr = 0; cudaEventElapsedTime(&tr, g.net.fe[0], g.net.fe[1]);                 // the trunk of this wave: the live roofline sample
                if (g.net.blocks > 0 && mode == 0) { g.net.conv_ms += tr; g.net.conv_sampled += (unsigned long long)(2 * g.net.blocks * g.net.NS * g.net.NS); }
            }
            ++wt_n;
            resolve_commit_timing();
        }
        return 0;
    }

    int search(int sims) override {
        if (sims <= 0) sims = cfg.num_simulations;
        // every region is provisioned for budget_sims more simulations since it was last cut: re-cut (in place, no move) before running past that
        if (sims_since_plan + sims > budget_sims && recut_regions(nullptr, std::max(sims, cfg.num_simulations))) return -1;
        sims_since_plan += sims;
        if (fork_groups()) return -1;
        for (auto& g : groups) {
            if (wave(g, 1)) return -1;                  // search() preamble: expand unexpanded roots
            if (!cfg.deterministic) {
                k_dirichlet<<<blocks_for_warps(g.n), 128, 0, g.stream>>>(g.tp, g.n, g.t0, cfg.dirichlet_alpha, cfg.dirichlet_epsilon, cfg.seed,
                                                                        noise_scratch + (size_t)g.t0 * MC, MC);
                AZ_LAUNCH_CHECK(); ++launches;
            }
        }
        for (int i = 0; i < sims; ++i)
            for (auto& g : groups) if (wave(g, 0)) return -1;
        waves += sims + 1;
        if (join_groups()) return -1;
        if (cfg.deterministic) return check_overflow();          // parity mode: a failed expansion is an error, not a statistic
        return 0;
    }
    // a failed expansion (node pool exhausted) changes the search: report it as an error
    int check_overflow() {
        if (sync_all()) return -1;
        Stats s; AZ_CUDA_CHECK(cudaMemcpy(&s, dstats, sizeof(Stats), cudaMemcpyDeviceToHost));
        if (s.pool_overflows > overflows_seen) {
            const unsigned long long n = s.pool_overflows - overflows_seen; overflows_seen = s.pool_overflows;
            set_error("node pool exhausted: " + std::to_string(n) + " expansion(s) failed — raise az_config.max_nodes_per_tree (average nodes per tree; now " +
                      std::to_string(pool_nodes / T) + ")");
            return -4;
        }
        return 0;
    }
    // k_region_need / k_region_plan / k_reroot_copy (tree_kernels.cuh) with the per-slot chosen children in `cc_dev` (nullptr: no slot moves), then swap the buffers
    int recut_regions(const int32_t* cc_dev, int sims_budget) {
        if (!cc_dev) {
            k_fill_i32<<<(T + 255) / 256, 256, 0, stream>>>(chosen_child, -2, T);
            AZ_LAUNCH_CHECK(); ++launches;
            cc_dev = chosen_child;
        }
        const int budget = (sims_budget + 1) * MC;
        k_region_need<<<(T + 255) / 256, 256, 0, stream>>>(tp, cc_dev, kept, T);
        AZ_LAUNCH_CHECK(); ++launches;
        k_region_plan<<<1, 1024, 0, stream>>>(kept, alt.base, alt.limit, T, budget, pool_nodes, plan);
        AZ_LAUNCH_CHECK(); ++launches;
        k_reroot_copy<G><<<blocks_for_warps(T), 128, warp_ws_bytes<G>(), stream>>>(tp, alt, root_state, cc_dev, T, dstats);
        AZ_LAUNCH_CHECK(); ++launches;
        for (auto f : {&TreePools::N, &TreePools::first, &TreePools::sub}) std::swap(tp.*f, alt.*f);
        std::swap(tp.W, alt.W); std::swap(tp.P, alt.P); std::swap(tp.act, alt.act); std::swap(tp.nchild, alt.nchild); std::swap(tp.flags, alt.flags);
        std::swap(tp.base, alt.base); std::swap(tp.limit, alt.limit);
        for (auto& g : groups) refresh_group_view(g);
        // the plan guarantees min(budget, slack share) growth per tree; budget_sims = what that covers
        budget_sims = sims_budget; sims_since_plan = 0;
        return 0;
    }

    int commit_moves(const int32_t* forced_dev) {
        MoveParams mp{cfg.deterministic, cfg.init_temperature, cfg.final_temperature, cfg.temperature_drop_move, cfg.seed};
        k_choose_move<G><<<blocks_for_warps(T), 128, warp_ws_bytes<G>(), stream>>>(tp, root_state, mp, game_buf, max_moves, forced_dev, chosen_child, chosen_action, T, dstats);
        AZ_LAUNCH_CHECK(); ++launches;
        if (recut_regions(chosen_child, cfg.num_simulations)) return -1;
        k_finish_games<G><<<blocks_for_warps(T), 128, warp_ws_bytes<G>(), stream>>>(tp, root_state, game_buf, max_moves, ring, ring_cap, ring_count, default_order, G::FIRST_FILL ? G::CELLS : 0,
                                                                   root_order, root_order_n, cfg.auto_restart, cfg.deterministic ? 0 : 1, T, dstats, tt);
        AZ_LAUNCH_CHECK(); ++launches;
        return 0;
    }

    int advance(const int32_t* actions, int n) override {
        AZ_CHECK(n == T, "az_engine_advance needs one action per slot");
        AZ_CUDA_CHECK(cudaMemcpyAsync(forced, actions, 4 * T, cudaMemcpyHostToDevice, stream));
        if (commit_moves(forced)) return -1;
        if (sync_all()) return -1;
        return 0;
    }

    int add_noise(float alpha, float eps) override {
        if (fork_groups()) return -1;
        for (auto& g : groups) {
            k_flag_noise<<<(g.n + 127) / 128, 128, 0, g.stream>>>(g.tp, g.n);
            AZ_LAUNCH_CHECK(); ++launches;
            if (wave(g, 1)) return -1;
            k_dirichlet<<<blocks_for_warps(g.n), 128, 0, g.stream>>>(g.tp, g.n, g.t0, alpha, eps, cfg.seed ^ (0x9E37ULL * ++noise_calls), noise_scratch + (size_t)g.t0 * MC, MC);
            AZ_LAUNCH_CHECK(); ++launches;
        }
        if (join_groups()) return -1;
        return sync_all();
    }
    unsigned long long noise_calls = 0;

    int play(int n_moves) override {
        for (int m = 0; m < n_moves; ++m) {
            if (search(cfg.num_simulations)) return -1;
            resolve_commit_timing();
            const bool tm = !cm_pending;
            if (tm) { for (auto& e : cm_ev) if (!e) cudaEventCreate(&e); cudaEventRecord(cm_ev[0], stream); }
            if (commit_moves(nullptr)) return -1;
            if (tm) { cudaEventRecord(cm_ev[1], stream); cm_pending = true; }
        }
        return 0;
    }

    int last_actions(int32_t* out, int n) override {
        AZ_CHECK(n == T, "need one entry per slot");
        if (sync_all()) return -1;
        AZ_CUDA_CHECK(cudaMemcpy(out, chosen_action, 4 * T, cudaMemcpyDeviceToHost));
        return 0;
    }

    int slot_state(int slot, int32_t* result, int32_t* ply, int32_t* player) override {
        AZ_CHECK(slot >= 0 && slot < T, "slot out of range");
        if (sync_all()) return -1;
        std::unique_ptr<State> sp(new State());
        AZ_CUDA_CHECK(cudaMemcpy(sp.get(), root_state + slot, sizeof(State), cudaMemcpyDeviceToHost));
        if (result) *result = G::host_root_result(*sp);
        if (ply) *ply = G::host_ply(*sp);
        if (player) *player = G::host_player(*sp);
        return 0;
    }

    int set_external(az_eval_fn fn, void* user) override {
        AZ_CHECK(cfg.evaluator == AZ_EVAL_EXTERNAL, "engine was not created with AZ_EVAL_EXTERNAL");
        ext_fn = fn; ext_user = user;
        return 0;
    }
    // one wave's leaves → the caller's evaluator → policy / value on the device (the group's stream is drained: one host round trip per wave)
    int external_eval(Group& g) {
        AZ_CHECK(ext_fn != nullptr, "no external evaluator set (az_engine_set_external_evaluator)");
        AZ_CHECK(NG == 1, "AZ_EVAL_EXTERNAL runs with one stream group");
        cudaStream_t st = g.stream;
        int32_t n = 0;
        AZ_CUDA_CHECK(cudaMemcpyAsync(&n, g.wb.n_eval, 4, cudaMemcpyDeviceToHost, st));
        k_leaf_paths<<<(g.n + 127) / 128, 128, 0, st>>>(g.tp, g.wb, ext_slot_tree, ext_paths, ext_plen, g.n);
        AZ_LAUNCH_CHECK(); ++launches;
        AZ_CUDA_CHECK(cudaStreamSynchronize(st));
        if (n <= 0) return 0;
        AZ_CUDA_CHECK(cudaMemcpyAsync(h_paths.data(), ext_paths, (size_t)n * MAX_DEPTH * 4, cudaMemcpyDeviceToHost, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(h_plen.data(), ext_plen, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(h_slot.data(), ext_slot_tree, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        AZ_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int i = 0; i < n; ++i) h_slot[i] += g.t0;
        std::fill(h_policy.begin(), h_policy.begin() + (size_t)n * A, 0.0f);
        AZ_CHECK(ext_fn(n, h_slot.data(), h_paths.data(), h_plen.data(), MAX_DEPTH, A, h_policy.data(), h_value.data(), ext_user) == 0, "the external evaluator reported an error");
        AZ_CUDA_CHECK(cudaMemcpyAsync(g.wb.policy, h_policy.data(), (size_t)n * A * 4, cudaMemcpyHostToDevice, st));
        AZ_CUDA_CHECK(cudaMemcpyAsync(g.wb.value, h_value.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        AZ_CUDA_CHECK(cudaStreamSynchronize(st));              // (pageable host vectors: the copies above are staged; keep it simple and ordered)
        return 0;
    }
    int get_timing(az_timing* o) override {
        if (sync_all()) return -1;
        resolve_commit_timing();
        std::memset(o, 0, sizeof(*o));
        o->waves_sampled = (uint64_t)wt_n; o->moves_sampled = (uint64_t)cm_n;
        o->select_ms = wt_ms[0]; o->dedup_encode_ms = wt_ms[1]; o->evaluator_ms = wt_ms[2]; o->expand_backup_ms = wt_ms[3]; o->commit_ms = cm_ms;
        o->stem_ms = wt_fwd[0]; o->trunk_ms = wt_fwd[1]; o->head_conv_ms = wt_fwd[2]; o->conv1x1_gemm_ms = wt_fwd[3]; o->policy_fc_ms = wt_fwd[4]; o->value_fc_ms = wt_fwd[5]; o->policy_value_ms = wt_fwd[6];
        return 0;
    }
    int set_search_params(float c_puct, int virtual_loss) override {
        AZ_CHECK(c_puct > 0.0f && virtual_loss >= 0, "bad search parameters");
        cfg.c_puct = c_puct; cfg.virtual_loss = virtual_loss;
        if (sync_all()) return -1;
        for (auto& g : groups) drop_graphs(g);                       // the search parameters are kernel arguments of the captured waves
        return 0;
    }
    // children of the node reached from the root by `path` (actions): MCTSNode::children / actions of any node (mcts_node.h:54-75)
    int set_num_simulations(int sims) override {
        AZ_CHECK(sims >= 1, "num_simulations must be at least 1");
        // the node pool was sized at creation (az_config.num_simulations / max_nodes_per_tree): a search must fit its worst-case growth
        AZ_CHECK(((long long)sims + 1) * MC + 1 <= pool_nodes / T, "num_simulations exceeds what the node pool was sized for (create the engine with the largest count, or a larger max_nodes_per_tree)");
        cfg.num_simulations = sims;                                // az_engine_play / the default of az_engine_search; regions are re-cut on demand (search())
        return 0;
    }
    int node_stats(int slot, const int32_t* path, int n_path, int32_t* actions, int32_t* visits, float* wsum, float* priors, int32_t* n,
                   int32_t* node_visits, float* node_wsum, float* node_prior, int32_t* node_flags) override {
        AZ_CHECK(slot >= 0 && slot < T, "slot out of range");
        AZ_CHECK(n_path >= 0 && (n_path == 0 || path), "bad path");
        if (sync_all()) return -1;
        long long base_ll = 0; AZ_CUDA_CHECK(cudaMemcpy(&base_ll, tp.base + slot, 8, cudaMemcpyDeviceToHost));
        const size_t base = (size_t)base_ll;
        int32_t node; AZ_CUDA_CHECK(cudaMemcpy(&node, tp.root + slot, 4, cudaMemcpyDeviceToHost));
        std::vector<int16_t> a16;
        for (int d = 0; d <= n_path; ++d) {
            int32_t f; int16_t nc;
            AZ_CUDA_CHECK(cudaMemcpy(&f, tp.first + base + node, 4, cudaMemcpyDeviceToHost));
            AZ_CUDA_CHECK(cudaMemcpy(&nc, tp.nchild + base + node, 2, cudaMemcpyDeviceToHost));
            const int count = f >= 0 ? nc : 0;
            a16.resize(count);
            if (count) AZ_CUDA_CHECK(cudaMemcpy(a16.data(), tp.act + base + f, 2 * count, cudaMemcpyDeviceToHost));
            if (d == n_path) {
                AZ_CHECK(*n >= count, "node_stats: buffers too small");
                *n = count;
                for (int i = 0; i < count; ++i) actions[i] = a16[i];
                if (count) {
                    AZ_CUDA_CHECK(cudaMemcpy(visits, tp.N + base + f, 4 * count, cudaMemcpyDeviceToHost));
                    AZ_CUDA_CHECK(cudaMemcpy(wsum, tp.W + base + f, 4 * count, cudaMemcpyDeviceToHost));
                    AZ_CUDA_CHECK(cudaMemcpy(priors, tp.P + base + f, 4 * count, cudaMemcpyDeviceToHost));
                }
                break;
            }
            int found = -1;
            for (int i = 0; i < count; ++i) if (a16[i] == path[d]) { found = i; break; }
            AZ_CHECK(found >= 0, "node_stats: action not found among the node's children");
            node = f + found;
        }
        uint8_t fl = 0;
        if (node_visits) AZ_CUDA_CHECK(cudaMemcpy(node_visits, tp.N + base + node, 4, cudaMemcpyDeviceToHost));
        if (node_wsum) AZ_CUDA_CHECK(cudaMemcpy(node_wsum, tp.W + base + node, 4, cudaMemcpyDeviceToHost));
        if (node_prior) AZ_CUDA_CHECK(cudaMemcpy(node_prior, tp.P + base + node, 4, cudaMemcpyDeviceToHost));
        if (node_flags) { AZ_CUDA_CHECK(cudaMemcpy(&fl, tp.flags + base + node, 1, cudaMemcpyDeviceToHost)); *node_flags = fl; }
        return 0;
    }

    int root_stats(int slot, int32_t* actions, int32_t* visits, float* wsum, float* priors, int32_t* n, int32_t* rn, float* rw) override {
        AZ_CHECK(slot >= 0 && slot < T, "slot out of range");
        if (sync_all()) return -1;
        long long base_ll = 0; AZ_CUDA_CHECK(cudaMemcpy(&base_ll, tp.base + slot, 8, cudaMemcpyDeviceToHost));
        const size_t base = (size_t)base_ll;
        int32_t root; AZ_CUDA_CHECK(cudaMemcpy(&root, tp.root + slot, 4, cudaMemcpyDeviceToHost));
        int32_t f; int16_t nc;
        AZ_CUDA_CHECK(cudaMemcpy(&f, tp.first + base + root, 4, cudaMemcpyDeviceToHost));
        AZ_CUDA_CHECK(cudaMemcpy(&nc, tp.nchild + base + root, 2, cudaMemcpyDeviceToHost));
        if (rn) AZ_CUDA_CHECK(cudaMemcpy(rn, tp.N + base + root, 4, cudaMemcpyDeviceToHost));
        if (rw) AZ_CUDA_CHECK(cudaMemcpy(rw, tp.W + base + root, 4, cudaMemcpyDeviceToHost));
        int count = f >= 0 ? nc : 0;
        AZ_CHECK(*n >= count, "root_stats: buffers too small");
        *n = count;
        if (count > 0) {
            std::vector<int16_t> a16(count);
            AZ_CUDA_CHECK(cudaMemcpy(a16.data(), tp.act + base + f, 2 * count, cudaMemcpyDeviceToHost));
            for (int i = 0; i < count; ++i) actions[i] = a16[i];
            AZ_CUDA_CHECK(cudaMemcpy(visits, tp.N + base + f, 4 * count, cudaMemcpyDeviceToHost));
            AZ_CUDA_CHECK(cudaMemcpy(wsum, tp.W + base + f, 4 * count, cudaMemcpyDeviceToHost));
            AZ_CUDA_CHECK(cudaMemcpy(priors, tp.P + base + f, 4 * count, cudaMemcpyDeviceToHost));
        }
        return 0;
    }

    int sample_layout(az_sample_layout* o) override {
        SampleT* z = nullptr;
        o->record_bytes = sizeof(SampleT);
        o->off_game_id = (int)(size_t)&z->game_id; o->off_slot = (int)(size_t)&z->slot; o->off_ply = (int)(size_t)&z->ply;
        o->off_action = (int)(size_t)&z->action; o->off_player = (int)(size_t)&z->player; o->off_z = (int)(size_t)&z->z;
        o->off_result = (int)(size_t)&z->result; o->off_root_value = (int)(size_t)&z->root_value; o->off_root_visits = (int)(size_t)&z->root_visits;
        o->off_state = (int)(size_t)&z->state; o->state_bytes = sizeof(typename G::Snapshot); o->off_visits = (int)(size_t)&z->visits[0];
        o->n_visits = (int)(sizeof(z->visits) / 2);
        return 0;
    }

    int drain(void* buf, size_t cap, size_t* n, bool device) override {
        if (sync_all()) return -1;
        int32_t cnt = 0; AZ_CUDA_CHECK(cudaMemcpy(&cnt, ring_count, 4, cudaMemcpyDeviceToHost));
        const size_t have = std::min<size_t>(cnt, ring_cap), take = std::min(have, cap);
        if (take) AZ_CUDA_CHECK(cudaMemcpy(buf, ring, take * sizeof(SampleT), device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
        // a caller buffer smaller than the ring content keeps the rest for the next call (records move to the front of the ring)
        const int32_t left = (int32_t)(have - take);
        if (left) {
            SampleT* tmp; if (dev_alloc(&tmp, (size_t)left)) return -1;
            AZ_CUDA_CHECK(cudaMemcpy(tmp, ring + take, (size_t)left * sizeof(SampleT), cudaMemcpyDeviceToDevice));
            AZ_CUDA_CHECK(cudaMemcpy(ring, tmp, (size_t)left * sizeof(SampleT), cudaMemcpyDeviceToDevice));
            cudaFree(tmp);
        }
        AZ_CUDA_CHECK(cudaMemcpy(ring_count, &left, 4, cudaMemcpyHostToDevice));
        *n = take;
        return 0;
    }

    int make_examples(const void* samples, size_t n, int augment, float* planes, float* policy, float* value) override {
        if (n == 0) return 0;
        AZ_CHECK(samples && planes && policy && value, "null buffer");
        const int k = (augment && cfg.game != AZ_GAME_CHESS) ? 8 : 1;              // dataset.cpp:250-253: no augmentation for chess
        const size_t chunk = 4096;                                                 // records per launch (bounds the staging buffers)
        SampleT* ds; float *dp, *dq, *dv;
        const size_t pe = (size_t)G::PLANES * G::CELLS;
        if (dev_alloc(&ds, chunk) || dev_alloc(&dp, chunk * k * pe) || dev_alloc(&dq, chunk * k * A) || dev_alloc(&dv, chunk * k)) return -1;
        for (size_t o = 0; o < n; o += chunk) {
            const size_t c = std::min(chunk, n - o);
            AZ_CUDA_CHECK(cudaMemcpyAsync(ds, (const SampleT*)samples + o, c * sizeof(SampleT), cudaMemcpyHostToDevice, stream));
            k_make_examples<G><<<blocks_for_warps((int)(c * k)), 128, warp_ws_bytes<G>(), stream>>>(ds, (int)c, k, dp, dq, dv);
            AZ_LAUNCH_CHECK(); ++launches;
            AZ_CUDA_CHECK(cudaMemcpyAsync(planes + o * k * pe, dp, c * k * pe * 4, cudaMemcpyDeviceToHost, stream));
            AZ_CUDA_CHECK(cudaMemcpyAsync(policy + o * k * A, dq, c * k * A * 4, cudaMemcpyDeviceToHost, stream));
            AZ_CUDA_CHECK(cudaMemcpyAsync(value + o * k, dv, c * k * 4, cudaMemcpyDeviceToHost, stream));
            AZ_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        for (void* p : {(void*)ds, (void*)dp, (void*)dq, (void*)dv}) cudaFree(p);
        return 0;
    }

    // Dataset::extractExamples on game records (move lists): one example per recorded move, k images each; chunked over positions
    int examples_from_games(const int32_t* moves, const int32_t* n_moves, const int8_t* results, int n_games, int max_mv, const float* policy_in, int P,
                            int augment, float* planes, float* policy, float* value) override {
        AZ_CHECK(n_games >= 0 && max_mv >= 1 && P >= 0, "bad sizes");
        if (n_games == 0) return 0;
        AZ_CHECK(moves && n_moves && results && planes && value && (P == 0 || (policy_in && policy)), "null buffer");
        const int k = (augment && cfg.game != AZ_GAME_CHESS) ? 8 : 1;              // dataset.cpp:250-253: no augmentation for chess
        std::vector<int32_t> pg, pp;
        for (int g = 0; g < n_games; ++g) {
            AZ_CHECK(n_moves[g] >= 0 && n_moves[g] <= max_mv, "n_moves out of range");
            AZ_CHECK(results[g] >= RES_ONGOING && results[g] <= RES_WIN_P2, "bad game result code");
            for (int i = 0; i < n_moves[g]; ++i) { pg.push_back(g); pp.push_back(i); }
        }
        const size_t n_pos = pg.size();
        if (n_pos == 0) return 0;
        const size_t chunk = 2048, pe = (size_t)G::PLANES * G::CELLS, Pz = (size_t)std::max(P, 1);
        int32_t *dm, *dg, *dpl, *derr; int8_t* dr; float *dpi, *dp, *dq, *dv; uint64_t* dh;
        if (dev_alloc(&dm, (size_t)n_games * max_mv) || dev_alloc(&dg, chunk) || dev_alloc(&dpl, chunk) || dev_alloc(&derr, 1) || dev_alloc(&dr, n_games) ||
            dev_alloc(&dpi, chunk * Pz) || dev_alloc(&dp, chunk * k * pe) || dev_alloc(&dq, chunk * k * Pz) || dev_alloc(&dv, chunk * k) ||
            dev_alloc(&dh, chunk * (size_t)(max_mv + 1))) return -1;
        AZ_CUDA_CHECK(cudaMemcpyAsync(dm, moves, (size_t)n_games * max_mv * 4, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemcpyAsync(dr, results, (size_t)n_games, cudaMemcpyHostToDevice, stream));
        AZ_CUDA_CHECK(cudaMemsetAsync(derr, 0, 4, stream));
        int rc = 0;
        for (size_t o = 0; o < n_pos && rc == 0; o += chunk) {
            const size_t c = std::min(chunk, n_pos - o);
            AZ_CUDA_CHECK(cudaMemcpyAsync(dg, pg.data() + o, c * 4, cudaMemcpyHostToDevice, stream));
            AZ_CUDA_CHECK(cudaMemcpyAsync(dpl, pp.data() + o, c * 4, cudaMemcpyHostToDevice, stream));
            if (P > 0) AZ_CUDA_CHECK(cudaMemcpyAsync(dpi, policy_in + o * P, c * P * 4, cudaMemcpyHostToDevice, stream));
            k_examples_from_games<G><<<blocks_for_warps((int)c), 128, warp_ws_bytes<G>(), stream>>>(dm, dg, dpl, (int)c, max_mv, dr, dpi, P, k, dp, dq, dv, dh, derr);
            AZ_LAUNCH_CHECK(); ++launches;
            AZ_CUDA_CHECK(cudaMemcpyAsync(planes + o * k * pe, dp, c * k * pe * 4, cudaMemcpyDeviceToHost, stream));
            if (P > 0) AZ_CUDA_CHECK(cudaMemcpyAsync(policy + o * k * P, dq, c * k * P * 4, cudaMemcpyDeviceToHost, stream));
            AZ_CUDA_CHECK(cudaMemcpyAsync(value + o * k, dv, c * k * 4, cudaMemcpyDeviceToHost, stream));
            int32_t err = 0; AZ_CUDA_CHECK(cudaMemcpyAsync(&err, derr, 4, cudaMemcpyDeviceToHost, stream));
            AZ_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (err) { set_error("illegal move in game record " + std::to_string(err - 1) + " (IGameState::makeMove throws, dataset.cpp:76)"); rc = -1; }
        }
        for (void* p : {(void*)dm, (void*)dg, (void*)dpl, (void*)derr, (void*)dr, (void*)dpi, (void*)dp, (void*)dq, (void*)dv, (void*)dh}) cudaFree(p);
        return rc;
    }

    int get_stats(az_stats* o) override {
        if (sync_all()) return -1;
        Stats s; AZ_CUDA_CHECK(cudaMemcpy(&s, dstats, sizeof(Stats), cudaMemcpyDeviceToHost));
        o->simulations = s.simulations; o->evaluations = s.evaluations; o->terminal_leaves = s.terminal_leaves; o->nodes_created = s.nodes_created;
        o->nodes_expanded = s.nodes_expanded; o->pool_overflows = s.pool_overflows; o->moves = s.moves; o->games = s.games; o->samples_dropped = s.samples_dropped; o->eval_shared = s.eval_shared; o->eval_cached = s.eval_cached;
        o->kernel_launches = launches + net_launches(); o->waves = waves;
        return 0;
    }
    int sync() override { return sync_all(); }

    int nn_forward(const float* planes, int n, float* policy, float* value, float* logits) override {
        AZ_CHECK(cfg.evaluator == AZ_EVAL_RESNET, "engine was created with the hash evaluator");
        AZ_CHECK(n >= 1, "n must be >= 1");
        if (sync_all()) return -1;
        Group& g = groups[0];
        Net& net = g.net;
        AZ_CHECK(net.loaded, "no network weights loaded (az_engine_load_weights)");
        const int cap = net.max_boards;
        float* dpl; if (dev_alloc(&dpl, (size_t)cap * net.in_planes * G::CELLS)) return -1;
        for (int o = 0; o < n; o += cap) {               // group 0's buffers, `cap` boards at a time
            const int c = std::min(cap, n - o);
            AZ_CUDA_CHECK(cudaMemcpyAsync(dpl, planes + (size_t)o * net.in_planes * G::CELLS, (size_t)c * net.in_planes * G::CELLS * 4, cudaMemcpyHostToDevice, g.stream));
            AZ_CHECK(nn::pack_planes_launch(dpl, net.in16, c, net.in_planes, net.cin_pad, G::N, G::N, net.row_pitch, net.board_pitch, net.p_total, nn::CONV_GUARD, net.f16, g.stream) == 0, "pack launch failed");
            ++launches;
            if (net.forward(nullptr, c, g.wb.policy, g.wb.value, g.stream)) { cudaFree(dpl); return -1; }
            AZ_CUDA_CHECK(cudaMemcpyAsync(policy + (size_t)o * A, g.wb.policy, (size_t)c * A * 4, cudaMemcpyDeviceToHost, g.stream));
            AZ_CUDA_CHECK(cudaMemcpyAsync(value + o, g.wb.value, (size_t)c * 4, cudaMemcpyDeviceToHost, g.stream));
            if (logits) AZ_CUDA_CHECK(cudaMemcpyAsync(logits + (size_t)o * A, net.logits, (size_t)c * A * 4, cudaMemcpyDeviceToHost, g.stream));
            AZ_CUDA_CHECK(cudaStreamSynchronize(g.stream));
        }
        cudaFree(dpl);
        return 0;
    }

    // whole-network forward over n_boards (in group-capacity chunks on group 0's buffers), `reps` times
    int nn_bench(int n_boards, int reps, float* ms) override {
        AZ_CHECK(cfg.evaluator == AZ_EVAL_RESNET, "engine was created with the hash evaluator");
        AZ_CHECK(n_boards >= 1, "n_boards must be >= 1");
        if (sync_all()) return -1;
        Group& g = groups[0];
        const int cap = g.net.max_boards;
        auto pass = [&]() -> int {
            for (int o = 0; o < n_boards; o += cap) if (g.net.forward(nullptr, std::min(cap, n_boards - o), g.wb.policy, g.wb.value, g.stream)) return -1;
            return 0;
        };
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (pass()) return -1;   // warm-up
        AZ_CUDA_CHECK(cudaEventRecord(e0, g.stream));
        for (int i = 0; i < reps; ++i) if (pass()) return -1;
        AZ_CUDA_CHECK(cudaEventRecord(e1, g.stream));
        AZ_CUDA_CHECK(cudaEventSynchronize(e1));
        float t = 0; cudaEventElapsedTime(&t, e0, e1); *ms = t / reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return 0;
    }

    // one production-shaped launch of the 128->128 conv (n_boards <= one group's capacity), timed alone
    int conv_bench(int n_boards, int reps, float* ms) override {
        if (sync_all()) return -1;
        Group& g = groups[0];
        Net& net = g.net;
        AZ_CHECK(cfg.evaluator == AZ_EVAL_RESNET && net.loaded && net.blocks >= 1, "conv_bench needs a loaded ResNet with >= 1 block");
        AZ_CHECK(n_boards >= 1 && n_boards <= net.max_boards, "n_boards must be in [1, slots per stream group]");
        nn::ConvParams cp{};
        cp.rowvalid = net.rowvalid; cp.n_boards_dev = nullptr; cp.n_rows = n_boards * net.board_pitch; cp.board_pitch = net.board_pitch;
        cp.p_total = net.p_total; cp.row_pitch = net.row_pitch; cp.relu = 1; cp.f16 = net.f16;
        cp.in = net.X; cp.out = net.Y; cp.resid = nullptr; cp.w = net.w.conv_w[net.wi(1, 0, 0)]; cp.bias = net.w.conv_b[net.bi(1, 0)];
        if (const char* d = getenv("AZ_CONV_DBG")) cp.dbg = atoi(d);      // profiling experiments (conv_trunk.cu)
        if (getenv("AZ_CONV_RESID")) cp.resid = net.X;                    // time the residual variant (second conv of a block)
        cp.pdl = getenv("AZ_CONV_NO_PDL") ? 0 : 1;
        long long* trace = nullptr;
        if (getenv("AZ_CONV_TRACE")) { if (dev_alloc(&trace, 2048)) return -1; AZ_CUDA_CHECK(cudaMemset(trace, 0, 2048 * 8)); cp.trace = trace; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaStream_t cs = net.conv_stream ? net.conv_stream : g.stream;
        const int cs_sms = net.conv_stream ? net.n_sms_conv : net.n_sms;
        for (int i = 0; i < 3; ++i) AZ_CHECK(nn::conv3x3_launch(cp, 128, cs_sms, cs) == 0, "conv launch failed");
        AZ_CUDA_CHECK(cudaEventRecord(e0, cs));
        for (int i = 0; i < reps; ++i) AZ_CHECK(nn::conv3x3_launch(cp, 128, cs_sms, cs) == 0, "conv launch failed");
        AZ_CUDA_CHECK(cudaEventRecord(e1, cs));
        AZ_CUDA_CHECK(cudaEventSynchronize(e1));
        float t = 0; cudaEventElapsedTime(&t, e0, e1); *ms = t / reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (trace) {   // stamps of the last launch: cluster 0's issuer (top, waits done, issued) + epilogue (start, end) per item; totals per cluster
            std::vector<long long> h(2048); AZ_CUDA_CHECK(cudaMemcpy(h.data(), trace, 2048 * 8, cudaMemcpyDeviceToHost)); cudaFree(trace);
            const long long t0 = h[0];
            for (int i = 0; i < 56; i += (i < 8 ? 1 : 8))
                fprintf(stderr, "item %2d: top %7lld ready %7lld issued %7lld | epi0 %7lld..%7lld | epi1 (own clock) %7lld..%7lld\n", i, h[i * 8] - t0, h[i * 8 + 3] - t0,
                        h[i * 8 + 4] - t0, h[i * 8 + 5] - t0, h[i * 8 + 6] - t0, h[i * 8 + 5 + 512] - h[5 + 512], h[i * 8 + 6 + 512] - h[5 + 512]);
            long long mn = 1LL << 60, mx = 0, sum = 0, wmx = 0; int n = 0;
            for (int c = 0; c < 74; ++c) { const long long v = h[1024 + 2 * c + 1]; if (v > 0) { mn = std::min(mn, v); mx = std::max(mx, v); sum += v; ++n; wmx = std::max(wmx, h[1024 + 2 * c]); } }
            for (int c = 0; c < 74; c += 6) fprintf(stderr, "cluster %2d: total %lld  wait a_full %lld  wait all %lld\n", c, h[1024 + 2 * c + 1], h[1280 + 2 * c], h[1280 + 2 * c + 1]);
            if (n) fprintf(stderr, "issuer cycles per cluster: min %lld avg %lld max %lld (%d clusters); weights resident after <= %lld cycles; %.1f us/launch => %.0f MHz\n", mn, sum / n, mx, n, wmx, *ms * 1e3, mx / (*ms * 1e3));
        }
        launches += reps + 3;
        return 0;
    }
    int conv_sampled(double* ms_sum, unsigned long long* launches_n) override {
        double m = 0; unsigned long long n = 0;
        for (auto& g : groups) { m += g.net.conv_ms; n += g.net.conv_sampled; }
        if (ms_sum) *ms_sum = m; if (launches_n) *launches_n = n;
        return 0;
    }
    cudaEvent_t events[8] = {};
    int event_record(int idx) override {
        AZ_CHECK(idx >= 0 && idx < 8, "event index out of range");
        if (!events[idx]) AZ_CUDA_CHECK(cudaEventCreate(&events[idx]));
        if (join_groups()) return -1;                    // the main stream is ordered after all group streams
        AZ_CUDA_CHECK(cudaEventRecord(events[idx], stream));
        return 0;
    }
    int event_elapsed(int i, int j, float* ms) override {
        AZ_CHECK(i >= 0 && i < 8 && j >= 0 && j < 8 && events[i] && events[j], "events not recorded");
        AZ_CUDA_CHECK(cudaEventSynchronize(events[j]));
        AZ_CUDA_CHECK(cudaEventElapsedTime(ms, events[i], events[j]));
        return 0;
    }

    int rules_replay(const int32_t* moves, const int32_t* n_moves, int n_games, int max_mv, int32_t* legal, int32_t* n_legal,
                     int32_t* terminal, int32_t* result, int32_t* player, float* planes) override {
        AZ_CHECK(n_games >= 1 && max_mv >= 1 && moves && n_moves, "az_rules_replay: bad sizes / null buffer");
        for (int g = 0; g < n_games; ++g) AZ_CHECK(n_moves[g] >= 0 && n_moves[g] <= max_mv, "az_rules_replay: n_moves out of range");
        int32_t *dm, *dn, *dl, *dnl, *dt, *dr, *dp; float* dpl = nullptr; uint64_t* dh;
        if (dev_alloc(&dh, (size_t)n_games * (max_mv + 1))) return -1;
        if (dev_alloc(&dm, (size_t)n_games * max_mv) || dev_alloc(&dn, n_games) || dev_alloc(&dl, (size_t)n_games * MC) || dev_alloc(&dnl, n_games) ||
            dev_alloc(&dt, n_games) || dev_alloc(&dr, n_games) || dev_alloc(&dp, n_games)) return -1;
        if (planes && dev_alloc(&dpl, (size_t)n_games * G::PLANES * G::CELLS)) return -1;
        AZ_CUDA_CHECK(cudaMemcpy(dm, moves, (size_t)n_games * max_mv * 4, cudaMemcpyHostToDevice));
        AZ_CUDA_CHECK(cudaMemcpy(dn, n_moves, (size_t)n_games * 4, cudaMemcpyHostToDevice));
        k_rules_replay<G><<<blocks_for_warps(n_games), 128, warp_ws_bytes<G>(), stream>>>(dm, dn, n_games, max_mv, dl, dnl, dt, dr, dp, dpl, dh);
        AZ_LAUNCH_CHECK(); ++launches;
        AZ_CUDA_CHECK(cudaStreamSynchronize(stream));
        AZ_CUDA_CHECK(cudaMemcpy(legal, dl, (size_t)n_games * MC * 4, cudaMemcpyDeviceToHost));
        AZ_CUDA_CHECK(cudaMemcpy(n_legal, dnl, (size_t)n_games * 4, cudaMemcpyDeviceToHost));
        AZ_CUDA_CHECK(cudaMemcpy(terminal, dt, (size_t)n_games * 4, cudaMemcpyDeviceToHost));
        AZ_CUDA_CHECK(cudaMemcpy(result, dr, (size_t)n_games * 4, cudaMemcpyDeviceToHost));
        AZ_CUDA_CHECK(cudaMemcpy(player, dp, (size_t)n_games * 4, cudaMemcpyDeviceToHost));
        if (planes) AZ_CUDA_CHECK(cudaMemcpy(planes, dpl, (size_t)n_games * G::PLANES * G::CELLS * 4, cudaMemcpyDeviceToHost));
        for (void* p : {(void*)dm, (void*)dn, (void*)dl, (void*)dnl, (void*)dt, (void*)dr, (void*)dp, (void*)dpl, (void*)dh}) cudaFree(p);
        return 0;
    }
};

}  // namespace az

// ================================================================================================ C ABI
struct az_engine { std::unique_ptr<az::EngineBase> impl; };

extern "C" {

AZ_API void az_config_default(az_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->game = AZ_GAME_GOMOKU; c->board_size = 15; c->n_slots = 4096; c->num_simulations = 800; c->c_puct = 1.5f; c->virtual_loss = 3;
    c->evaluator = AZ_EVAL_RESNET; c->net_blocks = 10; c->net_channels = 128; c->max_nodes_per_tree = 0; c->deterministic = 0;
    c->dirichlet_alpha = 0.03f; c->dirichlet_epsilon = 0.25f; c->init_temperature = 1.0f; c->final_temperature = 0.0f; c->temperature_drop_move = 30;
    c->auto_restart = 1; c->sample_ring_capacity = 0; c->device = 0; c->seed = 1234; c->n_streams = 1; c->net_precision = AZ_NET_FP16; c->tt_entries = 0; c->dense_policy = 0; c->eval_dedup = 0; c->eval_cache_entries = 0;
}

AZ_API const char* az_last_error(void) { return az::g_error.c_str(); }

AZ_API int az_engine_create(const az_config* cfg, az_engine** out) {
    if (!cfg || !out) { az::set_error("null argument"); return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { az::set_error("no CUDA device: az_b200 has no CPU path"); return -2; }
    if (cfg->n_slots < 1) { az::set_error("n_slots must be >= 1"); return -1; }
    std::unique_ptr<az::EngineBase> impl;
    int rc = -1;
    if (cfg->game == AZ_GAME_GOMOKU && cfg->board_size == 15) { auto* e = new az::EngineT<az::Gomoku<15>>(); impl.reset(e); rc = e->init(*cfg); }
    else if (cfg->game == AZ_GAME_GOMOKU && cfg->board_size == 9) { auto* e = new az::EngineT<az::Gomoku<9>>(); impl.reset(e); rc = e->init(*cfg); }
    else if (cfg->game == AZ_GAME_GO && cfg->board_size == 9) { auto* e = new az::EngineT<az::Go<9>>(); impl.reset(e); rc = e->init(*cfg); }
    else if (cfg->game == AZ_GAME_GO && cfg->board_size == 13) { auto* e = new az::EngineT<az::Go<13>>(); impl.reset(e); rc = e->init(*cfg); }
    else if (cfg->game == AZ_GAME_GO && cfg->board_size == 19) { auto* e = new az::EngineT<az::Go<19>>(); impl.reset(e); rc = e->init(*cfg); }
    else if (cfg->game == AZ_GAME_CHESS && (cfg->board_size == 8 || cfg->board_size == 0)) { auto* e = new az::EngineT<az::Chess>(); impl.reset(e); rc = e->init(*cfg); }
    else { az::set_error("unsupported game / board size (built: Gomoku 15x15, 9x9; Go 9x9, 13x13, 19x19; chess)"); return -3; }
    if (rc != 0) return rc;
    *out = new az_engine{std::move(impl)};
    return 0;
}
AZ_API int az_engine_destroy(az_engine* e) { delete e; return 0; }
#define AZ_FWD(call) do { if (!e) { az::set_error("null engine"); return -1; } \
                          if (cudaSetDevice(e->impl->device) != cudaSuccess) { az::set_error("cudaSetDevice failed"); return -1; } return e->impl->call; } while (0)
AZ_API int az_engine_load_weights(az_engine* e, const void* blob, size_t bytes) { AZ_FWD(load_weights(blob, bytes)); }
AZ_API int az_engine_reset_games(az_engine* e) { AZ_FWD(reset_games()); }
AZ_API int az_engine_set_root(az_engine* e, int slot, const int32_t* moves, int n, const int32_t* order, int n_order) { AZ_FWD(set_root(slot, moves, n, order, n_order)); }
AZ_API int az_engine_search(az_engine* e, int sims) { AZ_FWD(search(sims)); }
AZ_API int az_engine_root_stats(az_engine* e, int slot, int32_t* a, int32_t* v, float* w, float* p, int32_t* n, int32_t* rn, float* rw) { AZ_FWD(root_stats(slot, a, v, w, p, n, rn, rw)); }
AZ_API int az_engine_advance(az_engine* e, const int32_t* actions, int n) { AZ_FWD(advance(actions, n)); }
AZ_API int az_engine_play(az_engine* e, int n_moves) { AZ_FWD(play(n_moves)); }
AZ_API int az_engine_add_dirichlet_noise(az_engine* e, float alpha, float eps) { AZ_FWD(add_noise(alpha, eps)); }
AZ_API int az_engine_last_actions(az_engine* e, int32_t* actions, int n) { AZ_FWD(last_actions(actions, n)); }
AZ_API int az_engine_slot_state(az_engine* e, int slot, int32_t* r, int32_t* ply, int32_t* pl) { AZ_FWD(slot_state(slot, r, ply, pl)); }
AZ_API int az_engine_sample_layout(az_engine* e, az_sample_layout* out) { AZ_FWD(sample_layout(out)); }
AZ_API int az_engine_drain_samples(az_engine* e, void* buf, size_t cap, size_t* n) { AZ_FWD(drain(buf, cap, n, false)); }
AZ_API int az_engine_drain_samples_device(az_engine* e, void* buf, size_t cap, size_t* n) { AZ_FWD(drain(buf, cap, n, true)); }
AZ_API int az_engine_get_stats(az_engine* e, az_stats* out) { AZ_FWD(get_stats(out)); }
AZ_API int az_engine_examples_from_games(az_engine* e, const int32_t* moves, const int32_t* n_moves, const int8_t* results, int n_games, int max_moves, const float* policy_in,
                                         int policy_len, int augment, float* planes, float* policy, float* value) {
    AZ_FWD(examples_from_games(moves, n_moves, results, n_games, max_moves, policy_in, policy_len, augment, planes, policy, value));
}
AZ_API int az_engine_make_examples(az_engine* e, const void* samples, size_t n, int augment, float* planes, float* policy, float* value) { AZ_FWD(make_examples(samples, n, augment, planes, policy, value)); }
AZ_API int az_engine_sync(az_engine* e) { AZ_FWD(sync()); }
AZ_API int az_engine_set_search_params(az_engine* e, float c_puct, int virtual_loss) { AZ_FWD(set_search_params(c_puct, virtual_loss)); }
AZ_API int az_engine_set_num_simulations(az_engine* e, int sims) { AZ_FWD(set_num_simulations(sims)); }
AZ_API int az_engine_set_external_evaluator(az_engine* e, az_eval_fn fn, void* user) { AZ_FWD(set_external(fn, user)); }
AZ_API int az_engine_get_timing(az_engine* e, az_timing* out) { if (!out) { az::set_error("null argument"); return -1; } AZ_FWD(get_timing(out)); }
AZ_API int az_engine_node_stats(az_engine* e, int slot, const int32_t* path, int n_path, int32_t* a, int32_t* v, float* w, float* p, int32_t* n, int32_t* nv, float* nw,
                                float* np_, int32_t* nf) { AZ_FWD(node_stats(slot, path, n_path, a, v, w, p, n, nv, nw, np_, nf)); }
// plain device memory for the host layer's multi-GPU sample exchange (the host mirror links no CUDA runtime of its own)
#define AZ_DEV(dev) do { if (cudaSetDevice(dev) != cudaSuccess) { az::set_error("cudaSetDevice failed"); return -1; } } while (0)
AZ_API int az_device_count(int* n) { if (!n) return -1; if (cudaGetDeviceCount(n) != cudaSuccess) { *n = 0; cudaGetLastError(); } return 0; }
AZ_API int az_device_alloc(int device, size_t bytes, void** out) { AZ_DEV(device); AZ_CUDA_CHECK(cudaMalloc(out, bytes ? bytes : 1)); return 0; }
AZ_API int az_device_free(int device, void* p) { AZ_DEV(device); AZ_CUDA_CHECK(cudaFree(p)); return 0; }
AZ_API int az_device_memcpy(int device, void* dst, const void* src, size_t bytes, int to_host) {
    AZ_DEV(device); AZ_CUDA_CHECK(cudaMemcpy(dst, src, bytes, to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice)); return 0;
}
AZ_API int az_device_sync(int device) { AZ_DEV(device); AZ_CUDA_CHECK(cudaDeviceSynchronize()); return 0; }
AZ_API int az_engine_nn_forward(az_engine* e, const float* planes, int n, float* policy, float* value, float* logits) { AZ_FWD(nn_forward(planes, n, policy, value, logits)); }
AZ_API int az_engine_nn_bench(az_engine* e, int n_boards, int reps, float* ms) { AZ_FWD(nn_bench(n_boards, reps, ms)); }
AZ_API int az_engine_conv_bench(az_engine* e, int n_boards, int reps, float* ms) { AZ_FWD(conv_bench(n_boards, reps, ms)); }
AZ_API int az_engine_conv_sampled(az_engine* e, double* ms_sum, unsigned long long* launches_n) { AZ_FWD(conv_sampled(ms_sum, launches_n)); }
AZ_API int az_engine_event_record(az_engine* e, int idx) { AZ_FWD(event_record(idx)); }
AZ_API int az_engine_event_elapsed(az_engine* e, int i, int j, float* ms) { AZ_FWD(event_elapsed(i, j, ms)); }
AZ_API int az_rules_replay(az_engine* e, const int32_t* moves, const int32_t* n_moves, int n_games, int max_moves, int32_t* legal, int32_t* n_legal,
                    int32_t* terminal, int32_t* result, int32_t* player, float* planes) {
    AZ_FWD(rules_replay(moves, n_moves, n_games, max_moves, legal, n_legal, terminal, result, player, planes));
}

}  // extern "C"
