// heads.cuh — launch interfaces of the head / packing kernels (heads.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace az { namespace nn {

struct HeadParams {
    const __nv_bfloat16* act;   // trunk output [C/8][p_total][8]
    const float* w1x1;          // [64][C]  rows 0-31 policy conv, 32-63 value conv, BatchNorm scale folded in
    const float* b1x1;          // [64]     BatchNorm shift
    float* feat;                // [n][2][32*PH*PW]
    const int* n_boards_dev; int n_boards;
    int channels, H, W, row_pitch, board_pitch, p_total, guard;
};
struct FcParams {
    const float* x; const float* w; const float* bias; float* out;
    const int* n_boards_dev; int n_boards;
    int K, N, ldx, ldo, relu;
};
struct OutParams {
    const float* logits; const float* hidden; const float* w2; const float* b2;
    float* policy; float* value;
    const int* n_boards_dev; int n_boards;
    int A, hidden_n;
};

size_t head_pool_smem(int channels, int board_pitch, int H, int W);
int head_pool_conv_launch(const HeadParams& p, int grid, cudaStream_t s);
int fc_launch(const FcParams& p, int grid, cudaStream_t s);
int policy_value_launch(const OutParams& p, int max_boards, cudaStream_t s);
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int H, int W, int row_pitch,
                       int board_pitch, int p_total, int guard, cudaStream_t s);

}}  // namespace az::nn
