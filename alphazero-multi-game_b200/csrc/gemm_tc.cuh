// gemm_tc.cuh — launch interface of the tcgen05 head GEMM and the pooling kernel (gemm_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace az { namespace nn {

enum { GEMM_OUT_FEAT = 0, GEMM_OUT_ROWS = 1 };

struct GemmParams {
    const __nv_bfloat16* A;     // [K/8][a_rows][8]
    const __nv_bfloat16* B;     // weight image, gemm_weight_index()
    const float* bias;          // [n_tiles*64]
    int a_rows;                 // rows per K-chunk plane of A
    int a_plane_mod;            // A plane index wraps modulo this (hi/lo split: [A_hi | A_lo | A_hi] without a third copy)
    int K;                      // multiple of 64
    int n_tiles;                // N = n_tiles * 64
    int n_valid;                // columns actually written (GEMM_OUT_ROWS)
    int units;                  // independent row groups (1x1 conv: pooled cells; FC: 1)
    int unit_rows;              // row stride between units inside A
    const int* m_valid_dev; int m_valid;   // valid rows per unit (boards), device counter or fixed
    int relu;
    int mode;
    // GEMM_OUT_FEAT: bf16 features in the FC layers' A layout, feature = unit*32 + channel: out[(unit*4 + ch/8)][feat_rows][8]
    // as a bf16 hi/lo pair: hi at plane index p, lo at plane feat_lo_plane + p
    __nv_bfloat16* out_feat0; __nv_bfloat16* out_feat1; int feat_rows; int feat_lo_plane;
    // GEMM_OUT_ROWS: fp32 row-major out[row][ldo]
    float* out_rows; int ldo;
    // split-K (GEMM_OUT_ROWS only): k_splits >= 1 → item = (tile, split), RAW partial sums (no bias / ReLU) go to out_rows + split * split_stride;
    // 0 = classic epilogue
    int k_splits; size_t split_stride;
};

struct PoolParams {
    const __nv_bfloat16* act;   // trunk output [C/8][p_total][8]
    __nv_bfloat16* pooled;      // [2*C/8][pooled_rows][8] (hi planes then lo planes), row = cell*boards_cap + board
    const int* n_boards_dev; int n_boards;
    int channels, H, W, row_pitch, board_pitch, p_total, guard, boards_cap, pooled_rows;
    int f16;                    // 1: the trunk output is fp16 (else bf16); the pooled hi / lo planes are bf16 either way
};

inline size_t gemm_weight_elems(int n_total, int k_total) { return (size_t)((n_total + 63) / 64) * 64 * k_total; }
// image[ntile][k stage][k chunk j][n in tile][e]
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t gemm_weight_index(int k_total, int n, int k) {
    const int nt = n / 64, ni = n % 64, ks = k / 64, j = (k % 64) / 8, e = k % 8;
    return ((((size_t)nt * (k_total / 64) + ks) * 8 + j) * 64 + ni) * 8 + e;
}
int gemm_tc_launch(const GemmParams& p, int grid, cudaStream_t s);
int pool_launch(const PoolParams& p, int grid, cudaStream_t s);

}}  // namespace az::nn
