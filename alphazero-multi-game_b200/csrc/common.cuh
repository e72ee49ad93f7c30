// common.cuh — shared helpers for the B200 (sm_100a) self-play engine.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <string>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define AZ_HD __host__ __device__ __forceinline__
#define AZ_D __device__ __forceinline__
#else
#define AZ_HD inline
#define AZ_D inline
#endif

namespace az {

void set_error(const std::string& s);

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (kernel, device) pair: engines may live on several GPUs of one
// process (az_config.device), so the opt-in is made once per kernel AND device, under a lock (engine.cu).
cudaError_t smem_opt_in(const void* kernel, int bytes);

// Exactly-rounded fp32 ops with no FMA contraction: the search arithmetic must reproduce the
// reference's x86 single-op sequence bit for bit (SURVEY.md §7 "Hard parts").
AZ_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
AZ_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
AZ_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
AZ_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
AZ_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

AZ_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

#if defined(__CUDACC__)
}  // namespace az
#include <cuda_bf16.h>
#include <cuda_fp16.h>
namespace az {
// a network input value in the engine's 16-bit storage type: fp16 or bf16 bits in the same container (az_config.net_precision)
AZ_D __nv_bfloat16 net16(float x, bool f16) {
    if (f16) { const __half h = __float2half_rn(x); return *reinterpret_cast<const __nv_bfloat16*>(&h); }
    return __float2bfloat16_rn(x);
}
#endif

// core::GameResult (reference include/alphazero/core/igamestate.h:25-30)
enum : int { RES_ONGOING = 0, RES_DRAW = 1, RES_WIN_P1 = 2, RES_WIN_P2 = 3 };

// ParallelMCTS::convertToValue (reference src/mcts/parallel_mcts.cpp:973-985)
AZ_HD float result_to_value(int result, int player_to_move) {
    if (result == RES_WIN_P1) return player_to_move == 1 ? 1.0f : -1.0f;
    if (result == RES_WIN_P2) return player_to_move == 2 ? 1.0f : -1.0f;
    return 0.0f;
}

}  // namespace az

#define AZ_CUDA_CHECK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            az::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)
