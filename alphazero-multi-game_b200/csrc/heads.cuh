// heads.cuh — launch interfaces of the head / packing kernels (heads.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace az { namespace nn {

struct OutParams {
    const float* logits_part; const float* hidden_part;     // split-K partial sums [n_split][stride]: raw GEMM outputs
    size_t logits_stride, hidden_stride; int n_split_p, n_split_h;
    const float* bias_p; const float* bias_h;              // policy FC bias [A], value FC1 bias [hidden_n] (ReLU after it)
    const float* w2; const float* b2;
    float* logits;                                           // final logits [n][A] (kept for az_engine_nn_forward)
    float* policy; float* value;
    const int* n_boards_dev; int n_boards;
    int A, hidden_n;
    int want_logits;                                         // wide heads: also store the final logits (az_engine_nn_forward); the search needs the policy only
    int ld_part;                                             // row pitch (floats) of the policy partial-sum slabs (>= A, a multiple of 4: the GEMM epilogue stores float4)
};

int policy_value_launch(const OutParams& p, int max_boards, cudaStream_t s);

// policy over the legal moves only + value head (heads.cu k_policy_legal_value); wide heads (chess) inside the search waves
struct LegalPolicyParams {
    const __nv_bfloat16* featP; int feat_rows, feat_lo_plane;   // pooled policy features, bf16 hi / lo planes (the policy FC's A operand)
    const __nv_bfloat16* w_rows;                                // policy FC weights row-major [A][2048] bf16, K in feature order (cell * 32 + channel)
    const float* bias_p;
    const int16_t* legal; const int32_t* n_legal; int legal_pitch;   // per TREE: the leaf's legal actions in child order
    const int32_t* slot_tree;                                        // evaluation slot (= board of the batch) → tree
    const float* hidden_part; size_t hidden_stride; int n_split_h; const float* bias_h; const float* w2; const float* b2; int hidden_n;
    float* policy; float* value;                                // policy[b][A]: only the legal entries are written
    const int* n_boards_dev; int n_boards; int A;
};
int policy_legal_value_launch(const LegalPolicyParams& p, int max_boards, cudaStream_t s);
int pack_planes_launch(const float* planes, __nv_bfloat16* in, int n, int Cp, int cin_pad, int H, int W, int row_pitch,
                       int board_pitch, int p_total, int guard, int f16, cudaStream_t s);

}}  // namespace az::nn
