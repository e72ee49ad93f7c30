// conv_trunk.cuh — 3x3 convolution (+ folded BatchNorm bias, residual add, ReLU) as an implicit GEMM on
// tcgen05 tensor cores, operands staged in shared memory by TMA bulk copies, accumulators in TMEM.
//
// Replaces: the cuDNN convolutions behind torch::jit forward (reference src/nn/torch_neural_network.cpp:187,273)
// for the residual trunk of SURVEY.md §8a N1b.
//
// Layout ("padded position stream"): activations live in HBM as  act[C/8][p_total][8]  bf16.  A row is one
// board cell; a board of H x W cells occupies (H+1)*(W+1) consecutive rows with row pitch W+1 — column W and
// row H are permanent zeros — and boards follow each other back to back, with `guard` zero rows before the
// first and after the last.  The zero column/row give every cell its 3x3 zero padding for free: tap (ky,kx) of
// output row r reads input row r + (ky-1)*(W+1) + (kx-1).  So for one tap the A operand of the GEMM
// (M = 128 consecutive cells, K = 16 channels) is the SAME shared-memory tile at a shifted start address:
// no im2col, no duplicated data, one TMA load per (channel chunk, work item).
//
// Shared-memory operand format: K-major, SWIZZLE_NONE core matrices (8 rows x 16 bytes contiguous).  With
// act[c/8][row][c%8] every 8 consecutive rows of one channel chunk ARE a core matrix, wherever the start row
// is, which is what makes the shifted-address trick legal (SBO = 128 B, LBO = bytes per channel-chunk plane).
//
// Work item: BM = 256 consecutive rows (two UMMA M=128 tiles; for 15x15 Gomoku exactly one board).
// Per item: 9 taps x CIN/16 K-steps x 2 M-tiles tcgen05.mma (128x128x16) into two 128-column TMEM accumulators;
// accumulators are double-buffered (4 x 128 = 512 TMEM columns) so the epilogue of item i overlaps the MMAs
// of item i+1.  Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..32w+31), warp 4 TMA producer,
// warp 5 MMA issuer + TMEM allocator.  Persistent: grid = #SMs, items strided by gridDim.x.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace az { namespace nn {

constexpr int CONV_BM = 256;        // rows per work item
constexpr int CONV_HALO = 24;       // max |tap shift| supported = W + 2 <= 24  (W <= 22)
constexpr int CONV_GUARD = 32;      // zero rows before/after the board stream (>= CONV_HALO)
constexpr int CONV_COUT = 128;
constexpr int CONV_THREADS = 192;

struct ConvParams {
    const __nv_bfloat16* in;        // [CIN/8][p_total][8]
    __nv_bfloat16* out;             // [COUT/8][p_total][8]
    const __nv_bfloat16* resid;     // same layout as out, or nullptr
    const __nv_bfloat16* w;         // weight image, see conv_weight_image()
    const float* bias;              // [COUT] folded BatchNorm shift
    const uint8_t* rowvalid;        // [p_total] 1 = real board cell, 0 = padding row
    const int* n_boards_dev;        // optional: rows to process = *n_boards_dev * board_pitch
    int n_rows;                     // rows to process when n_boards_dev == nullptr
    int board_pitch;                // rows per board (H+1)*(W+1)
    int p_total;                    // rows per channel-chunk plane
    int row_pitch;                  // W + 1
    int relu;
    int reverse;                    // pair kernel: walk the work items from the last to the first (alternated layer by layer so a layer starts on the
                                    // rows its predecessor wrote last, which are still in L2)
    long long* trace;               // profiling only: per-item clock64 stamps of cluster 0 (conv_bench, AZ_CONV_TRACE), else nullptr
    int dbg;                        // profiling experiments only (conv_bench): see conv_trunk.cu; 0 in production
    int f16;                        // 1: activations, residual and weights are fp16 (else bf16) in the same 16-bit containers
    int l2_hints;                   // wide pair kernel (slice launches of 256-channel trunks): bit 0 = stores evict_last (a partial sum the next launch reads back),
                                    // bit 1 = activation loads evict_first (streamed once per launch)
    int pdl;                        // pair kernels: launch with programmatic stream serialization (the prologue — barriers, TMEM, the resident
                                    // weight half — runs while the previous kernel of the stream drains; activations are touched after griddepcontrol.wait)
};

// the whole residual trunk in one persistent launch (k_trunk_pair, conv_trunk.cu): layer l (0-based) reads X (l even) or Y (l odd) and writes
// the other; odd layers add the block's input (in place on X); ReLU everywhere
constexpr int TRUNK_MAX_LAYERS = 40;
struct TrunkParams {
    __nv_bfloat16* X; __nv_bfloat16* Y;          // [16][p_total][8] each
    const __nv_bfloat16* w[TRUNK_MAX_LAYERS];    // pair-kernel weight images, one per layer
    const float* bias[TRUNK_MAX_LAYERS];         // [128] folded BatchNorm shifts, one per layer
    const uint8_t* rowvalid;
    const int* n_boards_dev; int n_rows;
    int n_layers, p_total, row_pitch;
    int board_pitch;                             // rows per board (H+1)*(W+1)
    int f16;                                     // 1: fp16 activations / weights (else bf16)
    int group_boards;                            // boards a CTA pair takes through all layers at a time (trunk_group_boards): 74 pairs x 7 items x 128 KB stay in L2
    int balance;                                 // 1 (boards of any size, one group per pair): the group size follows the number of boards of THIS launch (device counter), so that
                                                 // every pair gets a group when in-wave sharing / the evaluation cache leave fewer boards than the engine has slots
    int discard;                                 // 1: L2 management — consumed Y rows are dropped from L2 instead of written back, X is stored evict_last (AZ_TRUNK_NO_DISCARD=1: off)
    int dbg;                                     // profiling experiments only (AZ_TRUNK_DBG): 1 = no cluster-scope release fence, 2 = no proxy fence, 4 = publish every item at once instead of one item later; 0 in production
};
bool trunk_fused_supported(int channels, int board_pitch, int row_pitch, int n_layers);
int trunk_group_boards(int board_pitch);
constexpr int TRUNK_TAIL_ROWS = 512;          // extra zero rows the activation planes need after the last board: a pair-local item grid may overhang the stream's end
int trunk_launch(const TrunkParams& p, int grid, cudaStream_t stream);

size_t conv_smem_bytes(int cin);
// launches on `stream`; cin in {16, 128}; returns cudaError_t as int
int conv3x3_launch(const ConvParams& p, int cin, int grid, cudaStream_t stream);
// element index of weight (tap, cin, cout) inside the image for a layer with `cin` input channels
// Weight image = the exact shared-memory picture the kernel consumes.
//   single-CTA kernel: [tap][K stage h][8-channel chunk j][cout n][8 channels e], stages in consumption order;
//   pair kernel:       [CTA rank r = cout / 64][tap][8-channel chunk j (16)][cout % 64][8 channels e].
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t conv_weight_index(int cin_total, bool pair, int tap, int ci, int co) {
    if (pair) return ((((size_t)(co / 64) * 9 + tap) * 16 + ci / 8) * 64 + co % 64) * 8 + ci % 8;
    const int wk = cin_total < 64 ? cin_total : 64;
    const int spt = cin_total / wk;
    const int h = ci / wk, j = (ci % wk) / 8, e = ci % 8;
    return ((((size_t)tap * spt + h) * (wk / 8) + j) * CONV_COUT + co) * 8 + e;
}
// true when the launch for (cin, row_pitch) goes to the weight-stationary CTA-pair kernel (its own weight image layout)
bool conv_uses_pair(int cin, int row_pitch);
inline size_t conv_weight_elems(int cin_total) { return (size_t)9 * cin_total * CONV_COUT; }

}}  // namespace az::nn
