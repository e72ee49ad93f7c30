// tools/mma_rate.cu — microbenchmark (not product code): issue rate of tcgen05.mma kind::f16 (bf16, K = 16) from shared-memory
// operands in the SWIZZLE_NONE layout the conv trunk uses, for cta_group::1 / ::2, several N, one or two accumulators and
// 16-byte-shifted A start addresses.  Prints cycles per MMA measured with clock64 around a commit + wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_rate tools/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../alphazero-multi-game_b200/csrc/ptx.cuh"
using namespace az::ptx;

struct Args { int cta_group, N, n_acc, a_shift, iters, lbo_a, unroll_taps; long long* out; };

__global__ void __launch_bounds__(192, 1) k_rate(Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = a.cta_group == 2 ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 5) { if (a.cta_group == 2) tmem_alloc2(&tslot, 512); else tmem_alloc(&tslot, 512); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    if (a.cta_group == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (warp == 5 && lane == 0 && rank == 0) {
        const int M = a.cta_group == 2 ? 256 : 128;
        const uint32_t idesc = idesc_bf16(M, a.N);
        const int rowsB = a.cta_group == 2 ? a.N / 2 : a.N;
        const uint64_t a0 = smem_desc(smem_u32(smem) + 32 * 16, a.lbo_a, 128);
        const uint64_t b0 = smem_desc(smem_u32(smem) + 64 * 1024, rowsB * 16, 128);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            // 8 MMAs per trip with descriptors that are `base + constant`; the tap shift advances once per trip
            const uint64_t astep = (uint64_t)(2 * (a.lbo_a >> 4)), bstep = (uint64_t)(2 * rowsB);
            uint64_t at = a0;
            if (a.cta_group == 2) {
                for (int it = 0; it < a.iters; it += 8) {
                    const uint32_t acc = tb + ((a.n_acc == 2 && (it & 8)) ? 256 : 0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma2_bf16(acc, at + k * astep, b0 + k * bstep, idesc, 1u);
                    at = (it & 64) ? a0 : at + a.a_shift;
                }
            } else {
                for (int it = 0; it < a.iters; it += 8) {
                    const uint32_t acc = tb + ((a.n_acc == 2 && (it & 8)) ? 256 : 0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16(acc, at + k * astep, b0 + k * bstep, idesc, 1u);
                    at = (it & 64) ? a0 : at + a.a_shift;
                }
            }
            if (a.cta_group == 2) umma2_commit_both(&bar); else umma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            const long long t1 = clock64();
            if (blockIdx.x == 0) a.out[rep] = t1 - t0;
        }
    }
    if (a.cta_group == 2 && rank == 1 && warp == 5 && lane == 0) { for (int rep = 0; rep < 3; ++rep) mbar_wait(&bar, rep & 1); }
    tc_fence_before();
    if (a.cta_group == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 5) { if (a.cta_group == 2) tmem_dealloc2(tb, 512); else tmem_dealloc(tb, 512); }
}

int main() {
    long long* out; cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int grids[2] = {2, 148};
    printf("cta_group  N  n_acc a_shift  lbo_a grid   cyc/MMA  (ideal = 128*N/256 per cta_group::1, M=256: same)\n");
    for (int g = 0; g < 2; ++g)
        for (int cg = 1; cg <= 2; ++cg)
            for (int N : {64, 128, 256})
                for (int nacc = 1; nacc <= 2; ++nacc)
                    for (int sh : {0, 1}) {
                        if (nacc == 2 && N == 256 && false) continue;
                        Args a{cg, N, nacc, sh, 2048, 2592, 0, out};
                        cudaLaunchConfig_t cfg{};
                        cfg.gridDim = dim3(grids[g]); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = 200 * 1024;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        cudaError_t e = cudaLaunchKernelEx(&cfg, k_rate, a);
                        if (e == cudaSuccess) e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s (cg %d N %d)\n", cudaGetErrorString(e), cg, N); return 1; }
                        printf("%9d %3d %5d %7d %6d %4d  %8.1f\n", cg, N, nacc, sh, 2592, grids[g], (double)out[2] / a.iters);
                    }
    return 0;
}
