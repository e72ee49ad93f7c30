// heads.cu — the last step of the policy / value heads (reference head shape: SURVEY.md §8a N1,
// python/alphazero/models/ddw_randwire.py:203-235).  Pooling, the 1x1 convs and the two big FC layers run on tcgen05
// (gemm_tc.cu); what is left here is elementwise: policy = softmax over the A logits (TorchNeuralNetwork::predictBatch,
// src/nn/torch_neural_network.cpp:298-316), value = tanh(FC 256→1), plus the fp32-planes → bf16 input packer used by
// az_engine_nn_forward.
#include "heads.cuh"
#include <cuda_bf16.h>

namespace az { namespace nn {

namespace {

// One warp per board: policy = softmax(logits[0..A)), value = tanh(hidden . w2 + b2).
__global__ void __launch_bounds__(128) k_policy_value(OutParams p) {
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nb) return;
    const float* lg = p.logits + (size_t)b * p.A;
    float mx = -3.4e38f;
    for (int i = lane; i < p.A; i += 32) mx = fmaxf(mx, lg[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.0f;
    for (int i = lane; i < p.A; i += 32) s += expf(lg[i] - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
    for (int i = lane; i < p.A; i += 32) p.policy[(size_t)b * p.A + i] = expf(lg[i] - mx) * inv;
    float d = 0.0f;
    for (int i = lane; i < p.hidden_n; i += 32) d = fmaf(p.hidden[(size_t)b * p.hidden_n + i], p.w2[i], d);
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) p.value[b] = tanhf(d + p.b2[0]);
}

// fp32 NCHW planes (host-supplied, az_engine_nn_forward) → the trunk's bf16 input layout
__global__ void k_pack_planes(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch, int board_pitch, int p_total, int guard) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cells = H * W;
    if (idx >= n * cells) return;
    const int b = idx / cells, cell = idx % cells, y = cell / W, x = cell % W;
    const size_t row = (size_t)guard + (size_t)b * board_pitch + y * row_pitch + x;
    for (int c = 0; c < cin_pad; ++c) {
        const float v = c < Cp ? planes[((size_t)b * Cp + c) * cells + cell] : 0.0f;
        in[((size_t)(c >> 3) * p_total + row) * 8 + (c & 7)] = __float2bfloat16_rn(v);
    }
}

}  // namespace

int policy_value_launch(const OutParams& p, int max_boards, cudaStream_t s) {
    k_policy_value<<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch, int board_pitch, int p_total, int guard, cudaStream_t s) {
    const int total = n * H * W;
    k_pack_planes<<<(total + 255) / 256, 256, 0, s>>>(planes, in, n, Cp, cin_pad, H, W, row_pitch, board_pitch, p_total, guard);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
