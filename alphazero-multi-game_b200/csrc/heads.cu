// heads.cu — the last step of the policy / value heads (reference head shape: SURVEY.md §8a N1,
// python/alphazero/models/ddw_randwire.py:203-235).  Pooling, the 1x1 convs and the two big FC layers run on tcgen05
// (gemm_tc.cu); what is left here is elementwise: policy = softmax over the A logits (TorchNeuralNetwork::predictBatch,
// src/nn/torch_neural_network.cpp:298-316), value = tanh(FC 256→1), plus the fp32-planes → bf16 input packer used by
// az_engine_nn_forward.
#include "heads.cuh"
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
        for (int sidx = 0; sidx < p.n_split_h; ++sidx) h += p.hidden_part[(size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + i];
        d = fmaf(fmaxf(h, 0.0f), p.w2[i], d);
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
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b >= nb) return;
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
    if (warp == 0) {      // value head: tanh(relu(hidden) . w2 + b2)
        float d = 0.0f;
        for (int i = lane; i < p.hidden_n; i += 32) {
            float h = p.bias_h[i];
            for (int sidx = 0; sidx < p.n_split_h; ++sidx) h += p.hidden_part[(size_t)sidx * p.hidden_stride + (size_t)b * p.hidden_n + i];
            d = fmaf(fmaxf(h, 0.0f), p.w2[i], d);
        }
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
    k_policy_value<<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch, int board_pitch, int p_total, int guard, int f16, cudaStream_t s) {
    const int total = n * H * W;
    k_pack_planes<<<(total + 255) / 256, 256, 0, s>>>(planes, in, n, Cp, cin_pad, H, W, row_pitch, board_pitch, p_total, guard, f16);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
