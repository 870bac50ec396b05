// Launchers for the bandwidth-bound kernels of the PGD hot path (everything that is not a GEMM).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tml {

typedef __nv_bfloat16 bf16;

// Dry run (host only): while set, every launcher of this library returns without touching the GPU.  The network
// walks are replayed this way when a layout is built, so the scratch size handed to the caller is the peak the real
// walk will reach -- measured, not hand-accounted -- before any kernel has been enqueued.
extern thread_local bool g_dry_run;

// conv_in (3 -> C0, 3x3 s1 p1), step 1: fp32 NCHW image -> bf16 NHWC [B,H,W,64]
// (channels [0,3) hi = bf16(x), [3,6) lo = bf16(x - hi), rest 0); step 2 is a 3x3 convolution on the GEMM kernels.
void launch_conv_in_pack(const float* x, bf16* a, int B, int H, int W, cudaStream_t s);
// conv_in input gradient, step 2 (col2im): y [B][H][W][32] fp32 tap products (entry (r*3+s)*3+ci) ->
// dx[b][ci][h][w] = beta*dx + sum_{r,s} y[b][h-r+1][w-s+1][(r*3+s)*3+ci]   (zero outside the image)
void launch_conv_in_col2im(const float* y, float* dx, int B, int H, int W, float beta, cudaStream_t s);

// GroupNorm(32 groups) over bf16 [B, HW, C].
//   partial: [B][chunks][32][2] fp32 scratch;  ss: [B][C] float2 (scale, shift);  mr: [B][32] float2 (mean, rstd)
int gn_num_chunks(int HW, int C);
void launch_gn_stats(const bf16* x, float* partial, int B, int HW, int C, cudaStream_t s);
// nchunks: number of partial entries per image (gn_num_chunks(HW, C) for launch_gn_stats, the tile count for
// partials written by a GEMM epilogue)
void launch_gn_finalize(const float* partial, const float* gamma, const float* beta, float2* ss, float2* mr, int B,
                        int HW, int C, float eps, int nchunks, cudaStream_t s);
void launch_gn_apply(const bf16* x, const float2* ss, bf16* y, int B, int HW, int C, int silu, cudaStream_t s);
// backward of y = act(GN(x)):  dx = rstd*(dxh - mean(dxh) - xh*mean(dxh*xh)) (+ resid)
void launch_gn_bwd_partial(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float* gamma,
                           float* partial, int B, int HW, int C, int silu, cudaStream_t s);
void launch_gn_bwd_finalize(const float* partial, float2* mm, int B, int HW, int C, int nchunks, cudaStream_t s);
void launch_gn_bwd_apply(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float2* mm,
                         const float* gamma, const bf16* resid, bf16* dx, int B, int HW, int C, int silu,
                         cudaStream_t s);

// attention helpers
// out[b][c][r] = in[b][r][c]; in row stride ld_in, batch strides in elements
void launch_transpose(const bf16* in, bf16* out, int batch, int R, int C, long long ld_in, long long bs_in,
                      long long ld_out, long long bs_out, cudaStream_t s, const float* row_scale = nullptr);
// softmax row helpers (the exp / normalisation / backward live in GEMM epilogues, see GemmOp::epi_mode)
void launch_row_reduce(const float* part, float* out, long long rows, int n, int op /*0 max, 1 reciprocal sum*/, cudaStream_t s);
void launch_row_dot(const bf16* a, const bf16* b, float* out, long long rows, int C /*multiple of 8, dense rows*/, cudaStream_t s);

// posterior sample + latent loss + d(loss)/d(moments)          (main.py:162,191; losses.py:39-41)
void launch_latent_loss(int kind, const float* moments, const float* noise, const float* target, int B, int h, int w,
                        float grad_scale, float* z, float* loss, float* dmoments, cudaStream_t s);
// fp32 NCHW [B,8,h,w] -> bf16 NHWC [B,h,w,64] (channels 8..63 zero): the A operand of conv_out's dgrad
void launch_dmoments_pack(const float* dm, bf16* out, int B, int h, int w, cudaStream_t s);

// PGD updates                                                   (main.py:248-276)
void launch_pgd_linf(float* x_adv, const float* grad, const float* x, float eps, float step, float lo, float hi,
                     long long n, cudaStream_t s);
size_t pgd_l2_workspace_bytes(int B, long long per_image);
void launch_pgd_l2(float* x_adv, const float* grad, const float* x, const float* mask, float eps, float step, float lo,
                   float hi, int B, int C, long long hw, void* ws, cudaStream_t s);
// universal perturbation                                        (old/train_noise.py:127-185)
void launch_add_delta(const float* x, const float* delta, float* out, int B, long long per_image, cudaStream_t s);
void launch_batch_sum(const float* g, float* out, int B, long long per_image, float scale, cudaStream_t s);
void launch_universal_step(float* delta, const float* grad, const float* source, float eps, float step, float lo,
                           float hi, long long n, void* ws, cudaStream_t s);
void launch_universal_project(float* delta, const float* sources, int nsrc, float lo, float hi, long long n, cudaStream_t s);

// ---- decoder-side helpers (vae.decode, main.py:156; image-space losses main.py:160,168) ----
// z fp32 NCHW [B,4,hw] -> post_quant_conv (4x4 1x1 conv + bias, fp32) -> bf16 NHWC [B,hw,64], channels 4..63 zero
void launch_latent_pack(const float* z, const float* wpq, const float* bpq, bf16* out, int B, int hw, cudaStream_t s);
// its backward: d bf16 NHWC [B,hw,64] (first 4 channels) -> dz fp32 NCHW = Wpq^T d
void launch_latent_unpack_bwd(const bf16* d, const float* wpq, float* dz, int B, int hw, cudaStream_t s);
// nearest-neighbour 2x upsample of bf16 NHWC [B,h,w,C] -> [B,2h,2w,C], and its backward (sum of the 4 copies)
void launch_upsample2x(const bf16* in, bf16* out, int B, int h, int w, int C, cudaStream_t s);
void launch_upsample2x_bwd(const bf16* dout, bf16* din, int B, int h, int w, int C, cudaStream_t s);
// fp32 NCHW [B,3,HW] -> bf16 NHWC [B,HW,64] (channels 3..63 zero): the A operand of decoder.conv_out's dgrad
void launch_image_pack(const float* dimg, bf16* out, int B, long long hw, cudaStream_t s);
// per image: rec = ||out - target||_2, pert = mean((out - source)^2), dout = rec_l*d rec + pert_l*d pert
size_t image_loss_workspace_bytes(int B);
void launch_image_loss(const float* out, const float* target, const float* source, int B, long long per_image,
                       float rec_l, float pert_l, float* rec, float* pert, float* dout, void* ws, cudaStream_t s);
// z = mean + exp(.5*clamp(logvar))*noise and its backward w.r.t. the moments
void launch_posterior_sample(const float* moments, const float* noise, float* z, int B, int hw, cudaStream_t s);
void launch_posterior_sample_bwd(const float* moments, const float* noise, const float* dz, float* dmoments, int B,
                                 int hw, cudaStream_t s);

// ---- UNet helpers (unet_kernels.cu; diffusion attack, main.py:229-243) ----
// GroupNorm(32 groups) for any channel count with C % 32 == 0 and C % 8 == 0 (partials / finalize as above)
int gng_num_chunks(int HW, int C);   // partial entries per image of the gng_* kernels (a function of HW and C only)
void launch_gng_stats(const bf16* x, float* partial, int B, int HW, int C, cudaStream_t s);
void launch_gng_apply(const bf16* x, const float2* ss, bf16* y, int B, int HW, int C, int silu, cudaStream_t s);
void launch_gng_bwd_partial(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float* gamma,
                            float* partial, int B, int HW, int C, int silu, cudaStream_t s);
void launch_gng_bwd_apply(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float2* mm,
                          const bf16* resid, bf16* dx, int B, int HW, int C, int silu, cudaStream_t s);
// LayerNorm over C of [rows][C] (C <= 2048); stats[row] = (mean, rstd)
void launch_ln_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float2* stats, long long rows, int C,
                   float eps, cudaStream_t s);
void launch_ln_bwd(const bf16* x, const bf16* dy, const float* gamma, const float2* stats, const bf16* resid, bf16* dx,
                   long long rows, int C, cudaStream_t s);
// GEGLU: h [rows][2I] = [x | gate] -> out [rows][I] = x * gelu(gate); backward dh [rows][2I]
void launch_geglu_fwd(const bf16* h, bf16* out, long long rows, int I, cudaStream_t s);
void launch_geglu_bwd(const bf16* h, const bf16* dout, bf16* dh, long long rows, int I, cudaStream_t s);
// [B][tok_src][ld_in] columns col0 + h*d .. -> [(B*heads)][tok_out][dpad] (zero padded; fill 1: query slot d = 1,
// fill 2: key slot d = -29952 on rows >= tok_valid), and the inverse
void launch_head_split(const bf16* in, long long ld_in, long long bs_in, int col0, bf16* out, int B, int heads,
                       int tok_src, int tok_valid, int tok_out, int d, int dpad, int fill, cudaStream_t s);
void launch_head_merge(const bf16* in, bf16* out, long long ld_out, long long bs_out, int col0, int B, int heads,
                       int tok, int d, int dpad, cudaStream_t s);
void launch_copy_cols(const bf16* in, long long ld_in, int ic0, bf16* out, long long ld_out, int oc0, int ncols,
                      long long rows, cudaStream_t s);
void launch_add_bf16(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t s);
void launch_timestep_embed(float t, float* out, int dim, cudaStream_t s);
void launch_small_linear(const float* x, const float* W, const float* bias, float* y, int N, int K, int silu_in,
                         cudaStream_t s);
void launch_nchw_pack64(const float* x, bf16* out, int B, int C, int hw, cudaStream_t s);
void launch_nhwc64_unpack(const bf16* in, float* out, int B, int C, int hw, cudaStream_t s);

// fused attention forward (attn_fused.cu): O = softmax(Q K^T * scale) V per batch given the row maxima of Q K^T;
// Q [nb][tq][dp], K / V [nb][tkv][dp] dense bf16, dp in {64, 128}, tq and tkv multiples of 128
bool attn_fused_supported(int tq, int tkv, int dp);
// rmax: known row maxima, or null = online softmax (the reference maxima it settles on go to rmax_out)
int launch_attn_fused_fwd(const bf16* Q, const bf16* K, const bf16* V, const float* rmax, float* rmax_out, float* inv_l,
                          bf16* O, int nb, int tq, int tkv, int dp, int lcol /* channel of V holding 1.0 */, float scale,
                          cudaStream_t s);

int launch_attn_fused_bwd(const bf16* Q, const bf16* K, const bf16* V, const bf16* O, const bf16* dO, const float* rmax,
                          const float* inv_l, bf16* dOs, float* Dp, bf16* dQ, bf16* dK, bf16* dV, int nb, int tq, int tkv,
                          int dp, float scale, cudaStream_t s);

long kernel_launch_count();

}  // namespace tml
