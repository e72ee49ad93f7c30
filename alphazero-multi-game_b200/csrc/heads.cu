// heads.cu — policy / value heads behind the trunk (reference head shape: SURVEY.md §8a N1,
// python/alphazero/models/ddw_randwire.py:203-235): adaptive_avg_pool2d to 8x8 → 1x1 conv (C→32, no bias)
// + BatchNorm + ReLU per head → policy FC(2048→A) → softmax (TorchNeuralNetwork::predictBatch,
// src/nn/torch_neural_network.cpp:298-316); value FC(2048→256)+ReLU → FC(256→1) → tanh.
// These are <1 % of the network's FLOPs; they run on CUDA cores in fp32 from the trunk's bf16 output.
#include "heads.cuh"
#include <cuda_bf16.h>

namespace az { namespace nn {

namespace {

// adaptive_avg_pool2d window of output index i: [floor(i*in/out), ceil((i+1)*in/out))
__device__ __forceinline__ void pool_window(int i, int in, int out, int& lo, int& hi) {
    lo = (i * in) / out;
    hi = ((i + 1) * in + out - 1) / out;
}

// One block per board: stage the board's trunk output in shared memory, pool to PHxPW, then both heads'
// 1x1 convolutions (BatchNorm scale folded into the weights) + shift + ReLU.
// feat[b][head][ch*PH*PW + cell]  (NCHW flatten order of torch .view(B, -1)).
__global__ void __launch_bounds__(256) k_head_pool_conv(HeadParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int C = p.channels, KCH = C / 8, BP = p.board_pitch;
    __nv_bfloat16* sAct = reinterpret_cast<__nv_bfloat16*>(smem);                 // [KCH][BP][8]
    float* sPool = reinterpret_cast<float*>(smem + (size_t)KCH * BP * 16);        // [C][cells]
    const int PH = p.H < 8 ? p.H : 8, PW = p.W < 8 ? p.W : 8, cells = PH * PW;
    const int n = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const size_t row0 = (size_t)p.guard + (size_t)b * BP;
        for (int i = threadIdx.x; i < KCH * BP; i += blockDim.x) {
            const int kc = i / BP, r = i % BP;
            reinterpret_cast<uint4*>(sAct)[i] = *reinterpret_cast<const uint4*>(p.act + ((size_t)kc * p.p_total + row0 + r) * 8);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells * C; i += blockDim.x) {
            const int c = i / cells, cell = i % cells;
            int y0, y1, x0, x1;
            pool_window(cell / PW, p.H, PH, y0, y1);
            pool_window(cell % PW, p.W, PW, x0, x1);
            float s = 0.0f;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) s += __bfloat162float(sAct[((size_t)(c >> 3) * BP + y * p.row_pitch + x) * 8 + (c & 7)]);
            sPool[c * cells + cell] = s / (float)((y1 - y0) * (x1 - x0));
        }
        __syncthreads();
        // 64 output channels (32 policy + 32 value) x cells
        for (int i = threadIdx.x; i < 64 * cells; i += blockDim.x) {
            const int oc = i / cells, cell = i % cells;
            const float* w = p.w1x1 + (size_t)oc * C;
            const float* x = sPool + cell;
            float s = 0.0f;
#pragma unroll 8
            for (int c = 0; c < C; ++c) s = fmaf(w[c], x[c * cells], s);
            s = fmaxf(s + p.b1x1[oc], 0.0f);
            const int head = oc >> 5, ch = oc & 31;
            p.feat[((size_t)b * 2 + head) * (32 * cells) + ch * cells + cell] = s;
        }
        __syncthreads();
    }
}

// out[b][n] = act(sum_k x[b*ldx + k] * w[n][k] + bias[n]); 64x64 block tile, 4x4 per thread, K tile 16.
__global__ void __launch_bounds__(256) k_fc(FcParams p) {
    __shared__ float sX[16][64 + 4];
    __shared__ float sWt[16][64 + 4];
    const int nb = p.n_boards_dev ? *p.n_boards_dev : p.n_boards;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int tiles_n = (p.N + 63) / 64, tiles_b = (nb + 63) / 64;
    for (int tile = blockIdx.x; tile < tiles_n * tiles_b; tile += gridDim.x) {
        const int b0 = (tile / tiles_n) * 64, n0 = (tile % tiles_n) * 64;
        float acc[4][4] = {};
        for (int k0 = 0; k0 < p.K; k0 += 16) {
            for (int i = threadIdx.x; i < 64 * 16; i += 256) {
                const int r = i >> 4, k = i & 15;
                const int b = b0 + r, nn = n0 + r;
                sX[k][r] = (b < nb) ? p.x[(size_t)b * p.ldx + k0 + k] : 0.0f;
                sWt[k][r] = (nn < p.N) ? p.w[(size_t)nn * p.K + k0 + k] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float xa[4], wa[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { xa[i] = sX[k][ty * 4 + i]; wa[i] = sWt[k][tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], wa[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = b0 + ty * 4 + i;
            if (b >= nb) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nn = n0 + tx * 4 + j;
                if (nn >= p.N) continue;
                float v = acc[i][j] + p.bias[nn];
                if (p.relu) v = fmaxf(v, 0.0f);
                p.out[(size_t)b * p.ldo + nn] = v;
            }
        }
    }
}

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
__global__ void k_pack_planes(const float* planes, __nv_bfloat16* in, int n, int Cp, int H, int W, int row_pitch, int board_pitch, int p_total, int guard) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cells = H * W;
    if (idx >= n * cells) return;
    const int b = idx / cells, cell = idx % cells, y = cell / W, x = cell % W;
    const size_t row = (size_t)guard + (size_t)b * board_pitch + y * row_pitch + x;
    for (int c = 0; c < 16; ++c) {
        const float v = c < Cp ? planes[((size_t)b * Cp + c) * cells + cell] : 0.0f;
        in[((size_t)(c >> 3) * p_total + row) * 8 + (c & 7)] = __float2bfloat16_rn(v);
    }
}

}  // namespace

size_t head_pool_smem(int channels, int board_pitch, int H, int W) {
    const int PH = H < 8 ? H : 8, PW = W < 8 ? W : 8;
    return (size_t)(channels / 8) * board_pitch * 16 + (size_t)PH * PW * channels * 4;
}
int head_pool_conv_launch(const HeadParams& p, int grid, cudaStream_t s) {
    const size_t sm = head_pool_smem(p.channels, p.board_pitch, p.H, p.W);
    static bool done = false;
    if (!done) { cudaError_t e = cudaFuncSetAttribute(k_head_pool_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e) return (int)e; done = true; }
    k_head_pool_conv<<<grid, 256, sm, s>>>(p);
    return (int)cudaGetLastError();
}
int fc_launch(const FcParams& p, int grid, cudaStream_t s) { k_fc<<<grid, 256, 0, s>>>(p); return (int)cudaGetLastError(); }
int policy_value_launch(const OutParams& p, int max_boards, cudaStream_t s) {
    k_policy_value<<<(max_boards * 32 + 127) / 128, 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int H, int W, int row_pitch, int board_pitch, int p_total, int guard, cudaStream_t s) {
    const int total = n * H * W;
    k_pack_planes<<<(total + 255) / 256, 256, 0, s>>>(planes, in, n, Cp, H, W, row_pitch, board_pitch, p_total, guard);
    return (int)cudaGetLastError();
}

}}  // namespace az::nn
