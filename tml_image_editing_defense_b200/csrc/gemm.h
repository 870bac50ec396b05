// Host-side description of one implicit-GEMM launch (convolution taps over an NHWC bf16 tensor,
// plain GEMM, or batched GEMM) and the launchers for the tcgen05 kernel and the SIMT debug kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tml {

constexpr int kMaxTaps = 9;
extern thread_local bool g_dry_run;   // see kernels.h: launchers return without launching
extern thread_local bool g_dry_validate;   // with g_dry_run: gemm_launch still plans the op and reports shape errors

// D[b, oh, ow, n] = alpha * sum_t sum_c A[b, oh*stride + dh[t], ow*stride + dw[t], c] * Bm[batch][n][t*A_C + c]
//                   + bias[n] + resid[b, oh, ow, n]
// Out-of-range A coordinates read as zero (that is the convolution padding).
struct GemmOp {
    // A operand: bf16, channel-contiguous (NHWC-like); strides in elements.
    const void* A = nullptr;
    int A_C = 0, A_W = 0, A_H = 0, A_B = 0;
    int64_t A_sW = 0, A_sH = 0, A_sB = 0;
    // a_trans = 1: A is stored TRANSPOSED, [batch][k][m] with the output-row index m = oh*OW + ow contiguous (row stride
    // A_sK, batch stride A_sB; A_C = K).  The kernel reads it as an MN-major UMMA operand, so products such as dS^T Q or
    // P^T dO need no transposed copy of the token x token matrix.  One tap, stride 1, tiles of whole rows only.
    int a_trans = 0;
    int64_t A_sK = 0;
    int stride = 1;  // 1 or 2
    int ntaps = 1;
    int dh[kMaxTaps] = {0}, dw[kMaxTaps] = {0};
    int OW = 0, OH = 0;  // output pixels per image (M = A_B * OH * OW)
    // B operand: bf16 [batch][N][Ktot], Ktot = ntaps * A_C, K contiguous.
    const void* Bm = nullptr;
    int N = 0;
    int64_t B_sN = 0;      // row stride (elements)
    int64_t B_sBatch = 0;  // 0: shared by all images (weights); else per-image operand
    // Epilogue.
    float alpha = 1.0f;
    const float* bias = nullptr;   // [N] or null
    const void* resid = nullptr;   // bf16, n contiguous; strides below
    int64_t R_sB = 0, R_sH = 0, R_sW = 0;
    void* D = nullptr;
    int out_fp32 = 0;  // 0: bf16, 1: fp32
    int64_t D_sB = 0, D_sH = 0, D_sW = 0, D_sN = 1;
    int n_store = 0;  // store only columns n < n_store (0 = N)
    float beta = 0.f;  // fp32 outputs only: D = beta * D + result (grad_reps accumulation)
    // Fused GroupNorm reductions over the bf16 output (dense NHWC, D_sN == 1, N % 32 == 0):
    //   gn_mode 1: partial[img][tile][g] = (sum, sumsq) of the output      -> statistics of the next GroupNorm
    //   gn_mode 2: partial[img][tile][g] = (sum dxh, sum dxh*xh), dxh = out * act'(x*sc+sh) * gamma,
    //              xh = (x-mean)*rstd                                       -> reductions of its backward
    // partial is [A_B][gemm_gn_tiles_per_image(op)][32][2] floats.
    int gn_mode = 0;
    float* gn_partial = nullptr;
    const void* gn_x = nullptr;        // bf16, same layout as D
    const float2* gn_ss = nullptr;     // [A_B][N]
    const float2* gn_mr = nullptr;     // [A_B][32]
    const float* gn_gamma = nullptr;   // [N]
    int gn_silu = 0;
    // Row-wise epilogues of the attention products (thread = output row = query token; dense bf16 output, lean path):
    //   epi_mode 1: no output; row_part[row][2*n_tiles] = max_j acc[row][j] per (column tile, column half)
    //   epi_mode 2: D = bf16(exp2((acc - row_a[row]) * exp_scale)); row_part[...] = sum of the stored values
    //   epi_mode 3: D = bf16(resid[row][col] * (acc - row_a[row]) * alpha * row_b[row])     (softmax backward)
    // row = image * OH*OW + pixel.  row_scale (any epilogue): result row multiplied by row_scale[row] (after alpha).
    int epi_mode = 0;
    const float* row_a = nullptr;
    const float* row_b = nullptr;
    float* row_part = nullptr;
    float exp_scale = 0.f;
    const float* row_scale = nullptr;
    // Fused input normalisation (operand-swapped CTA-pair kernel only, see gemm_fuses_input_gn): A is the RAW GroupNorm
    // input and every staged input row is rewritten in shared memory as silu(x * scale + shift) before the MMAs read it;
    // in_gn_ss = [A_B][A_C] (scale, shift) per image and input channel (gn_finalize's output).
    const float2* in_gn_ss = nullptr;
    // hardware experiment (tests only): A tile loaded `dbg_shift` pixels early into a (TW+8)-row box and
    // consumed through a row-shifted UMMA descriptor; dbg_bo = 1 also sets the descriptor base_offset field
    int dbg_shift = 0, dbg_bo = 0;
    const char* name = "";
};

// Derived tiling, shared by both kernels (the debug kernel ignores the tile fields).
struct GemmTiling {
    int TW, TH, rows_valid, tiles_w, tiles_h, BN, n_tiles, kchunks, stages, mt, stage_bytes, halo, halo_bytes, pair;
    int h66;         // halo mode on rows of 64 pixels (pitch-66 tile, 128-slot M tiles; tiles_h = tiles per image)
    int out_bytes;   // > 0: the epilogue stages output tiles in shared memory and writes them with TMA stores
    size_t smem_bytes;
};

// Returns 0 and fills `t`, or a negative value and sets the thread-local error message.
int gemm_plan(const GemmOp& op, GemmTiling* t);
// number of 128-row tiles per image (= chunk count of the fused GroupNorm partial buffer, gn_mode 1)
int gemm_gn_tiles_per_image(int OH, int OW);
// partial-sum entries per image this op's fused reduction writes (gn_mode set); at most 2 x tiles per image
int gemm_gn_chunks_per_image(const GemmOp& op);

// entries per output row of `row_part` for an epi_mode 1 / 2 op (2 x column tiles), or a negative error code
int gemm_row_partials(const GemmOp& op);

// true when gemm_launch_tc runs this op on the operand-swapped 3x3 kernel (channels as M, 256 pixels of a row as N),
// whose fused GroupNorm reductions are cheap enough to use at any K
bool gemm_swapped_shape(const GemmOp& op);

// true when this op (shape only) runs on the CTA-pair form of the operand-swapped kernel, which can apply the GroupNorm +
// SiLU of its input on the operand path (GemmOp::in_gn_ss) instead of reading a separately normalised tensor
bool gemm_fuses_input_gn(const GemmOp& op);

// tcgen05/TMEM/TMA implicit-GEMM kernel (the product path).
int gemm_launch_tc(const GemmOp& op, int num_sms, cudaStream_t stream);
// Straightforward SIMT kernel with identical semantics: test/debug aid only (selected through
// tml_debug_set_gemm_impl(1)); never used by default.
int gemm_launch_simt(const GemmOp& op, cudaStream_t stream);

int gemm_launch(const GemmOp& op, int num_sms, cudaStream_t stream);  // dispatches on the debug switch
void gemm_set_impl(int impl);
int gemm_get_impl();
long gemm_launch_count();  // number of tcgen05 GEMM launches since process start

void gemm_timing_enable(int max_launches);  // 0 disables and recycles the events
void gemm_timing_collect(double* total_ms, double* total_flops, long* launches, long* dropped);
size_t gemm_timing_report(char* buf, size_t cap);
int gemm_last_hang();  // id of the mbarrier wait that timed out in the last failed launch (0 = none)

void set_error(const char* fmt, ...);
const char* last_error();

}  // namespace tml
