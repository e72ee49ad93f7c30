// heads.cu — the last step of the policy / value heads (reference head shape: SURVEY.md §8a N1,
// python/alphazero/models/ddw_randwire.py:203-235).  Pooling, the 1x1 convs and the two big FC layers run on tcgen05
// (gemm_tc.cu); what is left here is elementwise: policy = softmax over the A logits (TorchNeuralNetwork::predictBatch,
// src/nn/torch_neural_network.cpp:298-316), value = tanh(FC 256→1), plus the fp32-planes → bf16 input packer used by
// az_engine_nn_forward.
#include "heads.cuh"
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace az { namespace nn {

namespace {

// One warp per board: policy = softmax(logits[0..A)), value = tanh(hidden . w2 + b2).
__global__ void __launch_bounds__(128) k_policy_value(OutParams p) {
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nb) return;
    // logits = bias + the split-K partial sums, added in slab order (deterministic)
    float* lg = p.logits + (size_t)b * p.A;
    float mx = -3.4e38f;
    for (int i = lane; i < p.A; i += 32) {
        float v = p.bias_p[i];
#pragma unroll 4
        for (int sidx = 0; sidx < p.n_split_p; ++sidx) v += p.logits_part[(size_t)sidx * p.logits_stride + (size_t)b * p.ld_part + i];
        lg[i] = v; mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.0f;
    for (int i = lane; i < p.A; i += 32) s += expf(lg[i] - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
    for (int i = lane; i < p.A; i += 32) p.policy[(size_t)b * p.A + i] = expf(lg[i] - mx) * inv;
    float d = 0.0f;
    for (int i = lane; i < p.hidden_n; i += 32) {
        float h = p.bias_h[i];
#pragma unroll 4
        for (int sidx = 0; sidx < p.n_split_h; ++sidx) h += p.hidden_part[(size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + i];
        d = fmaf(fmaxf(h, 0.0f), p.w2[i], d);
    }
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) p.value[b] = tanhf(d + p.b2[0]);
}

// The same for A <= 32 * PER (Gomoku 225, Go 82 / 170 / 362) with a board's logits in registers: the kernel is a chain of memory round
// trips (one warp per board), so all PER loads of a split-K slab are issued together, four slabs deep, instead of one element's slabs at
// a time, and the logits are not written to memory and read back between the max / sum / store passes.  The arithmetic per element and
// the order of every sum are those of k_policy_value: bit-identical output.
template <int PER>
__global__ void __launch_bounds__(128) k_policy_value_reg(OutParams p) {
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nb) return;
    constexpr int HP = 8;                                   // hidden units per lane and pass (hidden_n = 256: one pass)
    float v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int i = lane + 32 * k; v[k] = i < p.A ? __ldg(p.bias_p + i) : 0.0f; }
    const float* part = p.logits_part + (size_t)b * p.ld_part;
#pragma unroll 4
    for (int sidx = 0; sidx < p.n_split_p; ++sidx) {
        float tk[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) { const int i = lane + 32 * k; tk[k] = i < p.A ? part[(size_t)sidx * p.logits_stride + i] : 0.0f; }
#pragma unroll
        for (int k = 0; k < PER; ++k) if (lane + 32 * k < p.A) v[k] += tk[k];
    }
    float mx = -3.4e38f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = lane + 32 * k;
        if (i < p.A) { mx = fmaxf(mx, v[k]); if (p.want_logits) p.logits[(size_t)b * p.A + i] = v[k]; }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) if (lane + 32 * k < p.A) { v[k] = expf(v[k] - mx); s += v[k]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int i = lane + 32 * k; if (i < p.A) p.policy[(size_t)b * p.A + i] = v[k] * inv; }
    float d = 0.0f;
    for (int i0 = 0; i0 < p.hidden_n; i0 += 32 * HP) {
        float h[HP];
#pragma unroll
        for (int k = 0; k < HP; ++k) { const int i = i0 + lane + 32 * k; h[k] = i < p.hidden_n ? __ldg(p.bias_h + i) : 0.0f; }
        const float* hpart = p.hidden_part + (size_t)b * p.hidden_n + i0;
#pragma unroll 4
        for (int sidx = 0; sidx < p.n_split_h; ++sidx) {
            float tk[HP];
#pragma unroll
            for (int k = 0; k < HP; ++k) { const int i = lane + 32 * k; tk[k] = i0 + i < p.hidden_n ? hpart[(size_t)sidx * p.hidden_stride + i] : 0.0f; }
#pragma unroll
            for (int k = 0; k < HP; ++k) if (i0 + lane + 32 * k < p.hidden_n) h[k] += tk[k];
        }
#pragma unroll
        for (int k = 0; k < HP; ++k) { const int i = i0 + lane + 32 * k; if (i < p.hidden_n) d = fmaf(fmaxf(h[k], 0.0f), __ldg(p.w2 + i), d); }
    }
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) p.value[b] = tanhf(d + p.b2[0]);
}

// Wide heads (chess: A = 20480): one block of 1024 threads per board, the board's logits stay in registers between the max /
// sum / write passes (one read of the partial sums, one write of logits and policy).
constexpr int PV_WIDE_THREADS = 1024, PV_WIDE_MAX_PER_THREAD = 24;     // covers A <= 24576; many threads with few elements each: the passes are latency-bound
__global__ void __launch_bounds__(PV_WIDE_THREADS) k_policy_value_wide(OutParams p) {
    __shared__ float red_m[PV_WIDE_THREADS / 32], red_s[PV_WIDE_THREADS / 32];
    __shared__ float bcast[2];
    __shared__ float hid[1024];
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b >= nb) return;
    if (tid < p.hidden_n) {      // value head, first half: one hidden unit per thread, ahead of the softmax passes
        float h = p.bias_h[tid];
        for (int sidx = 0; sidx < p.n_split_h; ++sidx) h += __ldg(p.hidden_part + (size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + tid);
        hid[tid] = fmaxf(h, 0.0f);
    }
    // Online softmax: one pass gives every thread (max, sum of exp(x - max)) over its own logits, ONE block reduction combines the pairs
    // ((m1, s1) + (m2, s2) = (M, s1 e^(m1 - M) + s2 e^(m2 - M))), one expf per logit.  The kernel is latency-bound (one 1024-thread
    // block per SM, its phases do not overlap), so a reduction and an exp pass less is time saved.
    float v[PV_WIDE_MAX_PER_THREAD];
    // all loads first (a thread's ~20 partial sums and biases in flight together), the max afterwards
    const float* part0 = p.logits_part + (size_t)b * p.ld_part;
#pragma unroll
    for (int k = 0; k < PV_WIDE_MAX_PER_THREAD; ++k) {
        const int i = tid + k * PV_WIDE_THREADS;
        v[k] = i < p.A ? __ldg(p.bias_p + i) + __ldg(part0 + i) : 0.0f;
    }
    for (int sidx = 1; sidx < p.n_split_p; ++sidx) {
#pragma unroll
        for (int k = 0; k < PV_WIDE_MAX_PER_THREAD; ++k) {
            const int i = tid + k * PV_WIDE_THREADS;
            if (i < p.A) v[k] += p.logits_part[(size_t)sidx * p.logits_stride + (size_t)b * p.ld_part + i];
        }
    }
    float mx = -3.4e38f;
#pragma unroll
    for (int k = 0; k < PV_WIDE_MAX_PER_THREAD; ++k) {
        const int i = tid + k * PV_WIDE_THREADS;
        if (i < p.A) {
            mx = fmaxf(mx, v[k]);
            if (p.want_logits) p.logits[(size_t)b * p.A + i] = v[k];   // API path only (az_engine_nn_forward)
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < PV_WIDE_MAX_PER_THREAD; ++k) { const int i = tid + k * PV_WIDE_THREADS; if (i < p.A) { v[k] = expf(v[k] - mx); sum += v[k]; } }
    float M = mx, S = sum;
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, M, o), s2 = __shfl_xor_sync(0xffffffffu, S, o);
        const float mm = fmaxf(M, m2);
        S = S * expf(M - mm) + s2 * expf(m2 - mm); M = mm;
    }
    if (lane == 0) { red_m[warp] = M; red_s[warp] = S; }
    __syncthreads();
    if (warp == 0) {
        M = red_m[lane]; S = red_s[lane];                              // PV_WIDE_THREADS / 32 == 32 partial pairs
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, M, o), s2 = __shfl_xor_sync(0xffffffffu, S, o);
            const float mm = fmaxf(M, m2);
            S = S * expf(M - mm) + s2 * expf(m2 - mm); M = mm;
        }
        if (lane == 0) { bcast[0] = M; bcast[1] = S; }
    }
    __syncthreads();
    const float scale = expf(mx - bcast[0]) / bcast[1];              // this thread's exp(x - mx) values → exp(x - M) / S
#pragma unroll
    for (int k = 0; k < PV_WIDE_MAX_PER_THREAD; ++k) {
        const int i = tid + k * PV_WIDE_THREADS;
        if (i < p.A) p.policy[(size_t)b * p.A + i] = v[k] * scale;
    }
    if (warp == 0) {      // value head, second half: tanh(relu(hidden) . w2 + b2)
        float d = 0.0f;
        for (int i = lane; i < p.hidden_n; i += 32) d = fmaf(hid[i], p.w2[i], d);
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) p.value[b] = tanhf(d + p.b2[0]);
    }
}


// Policy over the LEGAL moves only (chess: A = 20480 logits, ~30 legal moves) + the value head: one 256-thread block per board.
// softmax over all A followed by the expansion's renormalisation over the legal moves (expandNodeWithPolicy, parallel_mcts.cpp:705-724) is
// algebraically the softmax over the legal logits alone: prior_a = e^{l_a} / sum_{legal} e^{l}.  So the 20480 x 2048 policy FC shrinks to
// one 2048-long dot product per legal move: the board's pooled features (fp32 from the bf16 hi / lo pair) sit in shared memory, a warp
// takes a legal action at a time and reads that action's weight row (row-major bf16 [A][2048], 4 KB, coalesced; the 84 MB matrix
// lives in L2), the block softmaxes the <= 256 logits and scatters the priors into policy[b][action] — exactly the entries the
// expansion reads.  Replaces 0.126 ms of GEMM + 0.119 ms of 20480-wide softmax per wave of 1024 boards.
constexpr int PL_THREADS = 256, PL_FEAT = 2048, PL_MAX_LEGAL = 256;
__global__ void __launch_bounds__(PL_THREADS) k_policy_legal_value(LegalPolicyParams p) {
    __shared__ __align__(16) float feat[PL_FEAT];
    __shared__ float logit[PL_MAX_LEGAL];
    __shared__ float red[PL_THREADS / 32];
    __shared__ float hid[1024];
    __shared__ int acts_s[PL_MAX_LEGAL];
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b >= nb) return;
    // The kernel is a chain of dependent memory round trips (one block per board, ~7 blocks per SM): every load that does not depend on
    // another one is issued before the first value is used — slot -> tree, the feature planes, and ALL split-K slabs of the value head's
    // hidden units (chess at 1024 boards runs the FC split 16 ways: added one behind the other they were 16 round trips).
    const int tree = p.slot_tree[b];
    const uint4 fhi = *reinterpret_cast<const uint4*>(p.featP + ((size_t)tid * p.feat_rows + b) * 8);                       // feature plane `tid` (8 features): hi + lo
    const uint4 flo = *reinterpret_cast<const uint4*>(p.featP + ((size_t)(p.feat_lo_plane + tid) * p.feat_rows + b) * 8);
    constexpr int MAX_SPLIT = 16;
    for (int i = tid; i < p.hidden_n; i += PL_THREADS) {      // value head, first half: hidden unit per thread (split-K slabs added in slab order)
        float hp[MAX_SPLIT];
#pragma unroll
        for (int sidx = 0; sidx < MAX_SPLIT; ++sidx) hp[sidx] = sidx < p.n_split_h ? __ldg(p.hidden_part + (size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + i) : 0.0f;
        float h = p.bias_h[i];
#pragma unroll
        for (int sidx = 0; sidx < MAX_SPLIT; ++sidx) if (sidx < p.n_split_h) h += hp[sidx];
        for (int sidx = MAX_SPLIT; sidx < p.n_split_h; ++sidx) h += __ldg(p.hidden_part + (size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + i);
        hid[i] = fmaxf(h, 0.0f);
    }
    const int n = min(p.n_legal[tree], PL_MAX_LEGAL);
    const int16_t* lg = p.legal + (size_t)tree * p.legal_pitch;
    if (tid < p.legal_pitch) acts_s[tid] = (int)(uint16_t)lg[tid];      // the leaf's legal actions (entries beyond n are not used), next to the load of n
    {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&fhi); const __nv_bfloat162* l = reinterpret_cast<const __nv_bfloat162*>(&flo);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 a = __bfloat1622float2(h[e]), c = __bfloat1622float2(l[e]); feat[tid * 8 + 2 * e] = a.x + c.x; feat[tid * 8 + 2 * e + 1] = a.y + c.y; }
    }
    __syncthreads();
    // one legal move per warp at a time, two moves' weight rows in flight
#pragma unroll 2
    for (int i = warp; i < n; i += PL_THREADS / 32) {
        const int a = acts_s[i];
        const uint4* wr = reinterpret_cast<const uint4*>(p.w_rows + (size_t)a * PL_FEAT);
        const float bp = __ldg(p.bias_p + a);                    // with the weight row, not behind the reduction
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < PL_FEAT / 256; ++j) {
            const uint4 wv = __ldg(wr + j * 32 + lane);
            const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
            const float4 f0 = *reinterpret_cast<const float4*>(feat + (j * 32 + lane) * 8), f1 = *reinterpret_cast<const float4*>(feat + (j * 32 + lane) * 8 + 4);
            const float2 w0 = __bfloat1622float2(wh[0]), w1 = __bfloat1622float2(wh[1]), w2 = __bfloat1622float2(wh[2]), w3 = __bfloat1622float2(wh[3]);
            acc = fmaf(w0.x, f0.x, acc); acc = fmaf(w0.y, f0.y, acc); acc = fmaf(w1.x, f0.z, acc); acc = fmaf(w1.y, f0.w, acc);
            acc = fmaf(w2.x, f1.x, acc); acc = fmaf(w2.y, f1.y, acc); acc = fmaf(w3.x, f1.z, acc); acc = fmaf(w3.y, f1.w, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) logit[i] = acc + bp;
    }
    __syncthreads();
    // softmax over the n legal logits (thread i owns logit i)
    const float x = tid < n ? logit[tid] : -3.4e38f;
    float mx = x;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int k = 1; k < PL_THREADS / 32; ++k) mx = fmaxf(mx, red[k]);
    __syncthreads();
    const float e = tid < n ? expf(x - mx) : 0.0f;
    float sm = e;
    for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    if (lane == 0) red[warp] = sm;
    __syncthreads();
    sm = 0.0f;
#pragma unroll
    for (int k = 0; k < PL_THREADS / 32; ++k) sm += red[k];
    if (tid < n) p.policy[(size_t)b * p.A + acts_s[tid]] = e / sm;
    if (warp == 0) {      // value head, second half: tanh(relu(hidden) . w2 + b2) (hid[] was published by the barriers above)
        float d = 0.0f;
        for (int i = lane; i < p.hidden_n; i += 32) d = fmaf(hid[i], p.w2[i], d);
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) p.value[b] = tanhf(d + p.b2[0]);
    }
}

// fp32 NCHW planes (host-supplied, az_engine_nn_forward) → the trunk's bf16 input layout
__global__ void k_pack_planes(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch, int board_pitch, int p_total, int guard, int f16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cells = H * W;
    if (idx >= n * cells) return;
    const int b = idx / cells, cell = idx % cells, y = cell / W, x = cell % W;
    const size_t row = (size_t)guard + (size_t)b * board_pitch + y * row_pitch + x;
    for (int c = 0; c < cin_pad; ++c) {
        const float v = c < Cp ? planes[((size_t)b * Cp + c) * cells + cell] : 0.0f;
        __nv_bfloat16 o = __float2bfloat16_rn(v);
        if (f16) { const __half h = __float2half_rn(v); o = *reinterpret_cast<const __nv_bfloat16*>(&h); }
        in[((size_t)(c >> 3) * p_total + row) * 8 + (c & 7)] = o;
    }
}

}  // namespace

int policy_value_launch(const OutParams& p, int max_boards, cudaStream_t s) {
    if (p.A > 1024) {
        if (p.A > PV_WIDE_THREADS * PV_WIDE_MAX_PER_THREAD) return (int)cudaErrorInvalidValue;
        k_policy_value_wide<<<max_boards, PV_WIDE_THREADS, 0, s>>>(p);
        return (int)cudaGetLastError();
    }
    static const bool old_kernel = getenv("AZ_PV_OLD") != nullptr;      // parity switch: the memory-resident variant (tools/pv_bits.py compares the two bit for bit)
    if (old_kernel) k_policy_value<<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    else if (p.A <= 256) k_policy_value_reg<8><<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    else if (p.A <= 384) k_policy_value_reg<12><<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    else k_policy_value<<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}
int policy_legal_value_launch(const LegalPolicyParams& p, int max_boards, cudaStream_t s) {
    if (p.feat_lo_plane != 256 || p.hidden_n > 1024) return (int)cudaErrorInvalidValue;        // built for 32 x 8 x 8 = 2048 pooled features
    k_policy_legal_value<<<max_boards, PL_THREADS, 0, s>>>(p);
    return (int)cudaGetLastError();
}
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch, int board_pitch, int p_total, int guard, int f16, cudaStream_t s) {
    const int total = n * H * W;
    k_pack_planes<<<(total + 255) / 256, 256, 0, s>>>(planes, in, n, Cp, cin_pad, H, W, row_pitch, board_pitch, p_total, guard, f16);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
