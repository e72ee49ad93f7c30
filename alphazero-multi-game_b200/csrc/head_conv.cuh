// head_conv.cuh — launch interface of the fused heads kernel (head_conv.cu): 1x1 convs on the full-resolution trunk output + adaptive
// average pooling + BatchNorm shift + ReLU → the FC GEMMs' bf16 hi/lo feature operands.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace az { namespace nn {

struct HeadConvParams {
    const __nv_bfloat16* act;   // trunk output [16][p_total][8] (128 channels)
    const __nv_bfloat16* w;     // folded 1x1 weights as a bf16 hi / lo pair: [2][16 k-chunks][64 out][8]; out 0-31 policy head, 32-63 value head
    const float* bias;          // [64] BatchNorm shifts
    const int* n_boards_dev; int n_boards;
    int H, W, row_pitch, board_pitch, p_total, guard;
    // features in the FC layers' A layout (gemm_tc.cuh GEMM_OUT_FEAT): plane = cell*4 + ch/8, row = board; lo half at plane feat_lo_plane + ...
    __nv_bfloat16* featP; __nv_bfloat16* featV; int feat_rows; int feat_lo_plane;
    int f16;                    // 1: the trunk output and the hi / lo weight image are fp16 (else bf16); the pooled features stay bf16 hi / lo
    int reverse;                // walk the boards from the last to the first: the trunk's last layer wrote them first to last, so the last ones are still in L2
};

// element index of weight (half s, out n, in k) inside the image
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t head_conv_weight_index(int s, int n, int k) { return ((((size_t)s * 16 + k / 8) * 64 + n) * 8) + k % 8; }
inline size_t head_conv_weight_elems() { return (size_t)2 * 16 * 64 * 8; }

bool head_conv_supported(int channels, int board_pitch, int H, int W);
int head_conv_launch(const HeadConvParams& p, int grid, cudaStream_t s);

}}  // namespace az::nn
