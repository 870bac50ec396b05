/* tml_b200 — C ABI of the B200-native PGD-immunization hot path.
 *
 * Plain C, opaque handle, caller-owned device buffers, explicit cudaStream_t (passed as void*),
 * fully asynchronous (no internal synchronisation), int status (0 = ok, <0 = error, message from
 * tml_last_error()).  No C++ exceptions and no torch types cross this boundary.
 * One handle per device; a handle is not thread-safe, distinct handles are.
 * Devices: a handle-based call (tml_encoder_*, tml_decoder_*) makes the handle's device current for its duration and
 * restores the caller's device before it returns; the stream must belong to that device.  The pointer-only calls
 * launch on the CURRENT device: make the buffers' device current before calling them.
 *
 * The reference (OrLichter/tml_image_editing_defense) is pure Python with no FFI of its own; each
 * entry point below names the reference call site it replaces (file:line in /root/reference).
 * The ctypes binding a maintainer of the reference would add is in INTEGRATION.md and implemented
 * in tml_image_editing_defense_b200/_lib.py.
 */
#ifndef TML_B200_H
#define TML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TmlEncoder TmlEncoder;

/* AutoencoderKL encoder configuration (diffusers config.json fields). */
typedef struct TmlEncoderCfg {
    int in_channels;            /* 3 */
    int latent_channels;        /* 4 */
    int num_blocks;             /* 4 */
    int block_out_channels[8];  /* 128,256,512,512 */
    int layers_per_block;       /* 2 */
    int norm_num_groups;        /* 32 */
    float norm_eps;             /* 1e-6 */
    int mid_block_add_attention;/* 1 */
} TmlEncoderCfg;

enum { TML_DTYPE_F32 = 0, TML_DTYPE_BF16 = 1, TML_DTYPE_F16 = 2 };
enum { TML_LOSS_L2NORM = 0, TML_LOSS_MSE = 1 };

const char* tml_last_error(void);
int tml_version(void);

/* ---- encoder: replaces pipeline.vae.encode(...) (main.py:75,191; old/train_noise.py:133;
 *      pipelines/pipeline_stable_diffusion_img2img.py:751,756) and its autograd backward
 *      (torch.autograd.grad(loss,[cur_image]), main.py:176). ---- */
int tml_encoder_create(const TmlEncoderCfg* cfg, int device, TmlEncoder** out);
void tml_encoder_destroy(TmlEncoder* enc);
/* Register one tensor of the diffusers state dict ("encoder.conv_in.weight", "quant_conv.bias", ...;
 * legacy attention names query/key/value/proj_attn are accepted).  ptr may be host or device memory. */
int tml_encoder_set_weight(TmlEncoder* enc, const char* diffusers_key, const void* ptr, int dtype,
                           const int64_t* shape, int ndim);
/* Repack to bf16 K-major GEMM operands (forward and input-gradient forms), fold quant_conv into
 * conv_out, upload.  Must be called once after all weights are set. */
int tml_encoder_finalize(TmlEncoder* enc, void* stream);
/* Bytes of scratch (`ws`) and of forward->backward state (`saved`) for a [B,3,H,W] batch.  The scratch size is not an
 * estimate: the call replays the forward and backward walks with every kernel launch disabled and reports the peak of
 * the same arena the real walks use (so it also validates the shape: a shape the walks cannot run fails here). */
int tml_encoder_query(TmlEncoder* enc, int B, int H, int W, size_t* workspace_bytes, size_t* saved_bytes);
/* x: fp32 NCHW [B,3,H,W] -> moments: fp32 NCHW [B, 2*latent, H/8, W/8] (quant_conv output). */
int tml_encoder_forward(TmlEncoder* enc, const float* x_nchw, int B, int H, int W, float* moments_nchw, void* saved,
                        void* ws, void* stream);
/* dmoments: fp32 NCHW [B,2*latent,H/8,W/8] -> dx = beta*dx + dLoss/dx, fp32 NCHW [B,3,H,W].
 * beta = 1 accumulates the grad_reps of main.py:88-102 in place. */
int tml_encoder_backward(TmlEncoder* enc, const float* dmoments_nchw, int B, int H, int W, const void* saved,
                         float* dx_nchw, float beta, void* ws, void* stream);

/* ---- decoder: replaces pipeline.vae.decode(output_latent).sample (main.py:156) and its backward.  Uses the
 *      same handle: register the "decoder.*" and "post_quant_conv.*" tensors with tml_encoder_set_weight before
 *      tml_encoder_finalize.  z: fp32 NCHW [B,4,h,w] -> image fp32 NCHW [B,3,8h,8w]. ---- */
int tml_decoder_query(TmlEncoder* vae, int B, int h, int w, size_t* workspace_bytes, size_t* saved_bytes);
int tml_decoder_forward(TmlEncoder* vae, const float* z_nchw, int B, int h, int w, float* image_nchw, void* saved,
                        void* ws, void* stream);
int tml_decoder_backward(TmlEncoder* vae, const float* dimage_nchw, int B, int h, int w, const void* saved,
                         float* dz_nchw, void* ws, void* stream);
/* image-space losses per image: rec = ||out - target||_2 (main.py:160), pert = mean((out - source)^2)
 * (losses/losses.py:39-41, main.py:168); dout = rec_lambda * d rec + pert_lambda * d pert (main.py:169).
 * source / rec / pert / dout may be NULL; ws of tml_image_loss_workspace(B) bytes. */
size_t tml_image_loss_workspace(int B);
int tml_image_loss(const float* out, const float* target, const float* source, int B, int64_t per_image,
                   float rec_lambda, float pert_lambda, float* rec, float* pert, float* dout, void* ws, void* stream);
/* latent_dist.sample() / mode() (noise NULL) as a stand-alone op and its backward w.r.t. the moments */
int tml_posterior_sample(const float* moments, const float* noise, float* z, int B, int h, int w, void* stream);
int tml_posterior_sample_backward(const float* moments, const float* noise, const float* dz, float* dmoments, int B,
                                  int h, int w, void* stream);

/* ---- posterior sample + latent loss + gradient: replaces latent_dist.sample() (main.py:191),
 *      (output_latent - target_latent).norm(p=2) (main.py:162, kind 0) and F.mse_loss
 *      (losses/losses.py:39-41, kind 1).  noise may be NULL (-> latent_dist.mode()).
 *      z_out, loss_per_image, dmoments may be NULL.  dmoments = grad_scale * dloss_b/dmoments. ---- */
int tml_latent_loss(int kind, const float* moments, const float* noise, const float* target, int B, int h, int w,
                    float grad_scale, float* z_out, float* loss_per_image, float* dmoments, void* stream);

/* ---- PGD updates: replace Trainer.perturbation_step (main.py:248-276). ---- */
/* linf branch, main.py:272-274; in place on x_adv; bit-exact with the ATen sequence. */
int tml_pgd_step_linf(float* x_adv, const float* grad, const float* x, float eps, float step, float lo, float hi,
                      int64_t n, void* stream);
/* l2 branch, main.py:254-268; mask [B,1,H,W] or NULL; ws of tml_pgd_l2_workspace(B) bytes.  The per-image norms are
 * reduced in a fixed order (reproducible for any batch split), not in ATen's: <= 2e-6 absolute on the result
 * ("Stated tolerances"). */
size_t tml_pgd_l2_workspace(int B);
int tml_pgd_step_l2(float* x_adv, const float* grad, const float* x, const float* mask, float eps, float step,
                    float lo, float hi, int B, int C, int64_t hw, void* ws, void* stream);

/* ---- universal perturbation (old/train_noise.py:127-185) ---- */
/* out[b] = x[b] + delta                                   (:132) */
int tml_add_delta(const float* x, const float* delta, float* out, int B, int64_t per_image, void* stream);
/* out = scale * sum_b g[b]   (fixed summation order)      (gradient of the shared delta) */
int tml_batch_sum(const float* g, float* out, int B, int64_t per_image, float scale, void* stream);
/* L2-normalised step, clamp to +-eps, optional image-range projection (:173-185); ws >= 1 KiB. */
int tml_universal_step(float* delta, const float* grad, const float* source, float eps, float step, float lo,
                       float hi, int64_t n, void* ws, void* stream);
/* Image-range re-projection alone (:183-185): for each of the nsrc source images [nsrc][n] in order,
 * delta = clamp(source + delta, lo, hi) - source.  One source = the reference statement; a sharded step
 * passes the per-pixel (min, max) images of the whole dataset so every replica applies the same bounds. */
int tml_universal_project(float* delta, const float* sources, int nsrc, float lo, float hi, int64_t n, void* stream);

/* ---- UNet: replaces self.pipeline.unet(latent_model_input, t, encoder_hidden_states=prompt_embeds).sample
 *      (main.py:233-238; diffusers UNet2DConditionModel, SD-1.5 topology) and its backward w.r.t. the sample
 *      (torch.autograd.grad(loss, [cur_image]) through the denoising loop, main.py:176,229-243). ---- */
typedef struct TmlUnet TmlUnet;
typedef struct TmlUnetCfg {
    int in_channels;            /* 4 */
    int out_channels;           /* 4 */
    int num_blocks;             /* 4 */
    int block_out_channels[8];  /* 320,640,1280,1280 (multiples of 64) */
    int layers_per_block;       /* 2 */
    int cross_attention_dim;    /* 768 */
    int num_heads;              /* 8 (diffusers' `attention_head_dim` of SD-1.5 is the head COUNT) */
    int norm_num_groups;        /* 32 */
    int down_has_attn[8];       /* 1,1,1,0 */
    int up_has_attn[8];         /* 0,1,1,1 */
} TmlUnetCfg;
int tml_unet_create(const TmlUnetCfg* cfg, int device, TmlUnet** out);
void tml_unet_destroy(TmlUnet* unet);
/* diffusers state-dict keys ("conv_in.weight", "down_blocks.0.resnets.0.norm1.weight", ...), host or device memory */
int tml_unet_set_weight(TmlUnet* unet, const char* diffusers_key, const void* ptr, int dtype, const int64_t* shape, int ndim);
int tml_unet_finalize(TmlUnet* unet, void* stream);
/* scratch / forward->backward state for a [B,4,h,w] sample with ctx_tokens prompt tokens (dry run of both walks) */
int tml_unet_query(TmlUnet* unet, int B, int h, int w, int ctx_tokens, size_t* workspace_bytes, size_t* saved_bytes);
/* sample fp32 NCHW [B,4,h,w]; one scalar timestep for the batch (main.py:233); ctx fp32 [B,ctx_tokens,cross_attention_dim]
 * -> out fp32 NCHW [B,4,h,w] (the predicted noise). */
int tml_unet_forward(TmlUnet* unet, const float* sample_nchw, float timestep, const float* ctx, int B, int h, int w,
                     int ctx_tokens, float* out_nchw, void* saved, void* ws, void* stream);
/* dout fp32 NCHW -> dsample fp32 NCHW; `saved` from the forward of the same inputs */
int tml_unet_backward(TmlUnet* unet, const float* dout_nchw, int B, int h, int w, int ctx_tokens, const void* saved,
                      float* dsample_nchw, void* ws, void* stream);
/* saved bf16 NHWC activations: "conv_in", "resnet_h1"/"resnet_out", "tf_t0"/"tf_x1"/"tf_x2"/"tf_out" (transformer index),
 * "down_out", "up_out"; "count_resnets"/"count_tf" return the count in *offset */
int tml_debug_unet_saved_tensor(TmlUnet* unet, const char* name, int index, size_t* offset, int dims[4]);
/* Tests on a machine without a GPU: while on, tml_unet_create skips the device checks and finalize uploads nothing, so
 * tml_unet_query still replays both walks as a dry run (layout, scratch size, shape validation of every GEMM). */
void tml_debug_set_host_only(int on);
/* The fused attention kernels on their own (kernel unit tests): Q [nb][tq][dp], K / V [nb][tkv][dp] dense bf16, dp = 64 or
 * 128, tq and tkv multiples of 128; channel `lcol` of V must be 1.0 on every key row (the softmax denominator column).
 * Forward (online softmax): O [nb][tq][dp] bf16, rmax / inv_l [nb][tq] fp32.  With dO (bf16, same shape as O): dQ, and dK / dV
 * unless both are NULL; ws >= nb*tq*(2*dp + 12) + 512 bytes. */
int tml_debug_attention(const void* Q, const void* K, const void* V, int nb, int tq, int tkv, int dp, int lcol, float scale,
                        void* O, float* rmax, float* inv_l, const void* dO, void* dQ, void* dK, void* dV, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- introspection / test hooks ---- */
/* number of kernels launched by this library since load: [0] tcgen05 GEMMs, [1] all other kernels */
void tml_launch_counts(int64_t out[2]);
/* Per-launch CUDA-event timing of the tcgen05 GEMM kernel on its launch stream (bench.py roofline).
 * enable(n): record up to n launches (0 = off).  collect (after a synchronize): out = {total ms,
 * total algorithmic flops (2*M*N*K, conv padding not discounted), launches timed, launches dropped}. */
void tml_gemm_timing_enable(int max_launches);
void tml_gemm_timing_collect(double out[4]);
/* per-shape table of the timed launches, one "name|M|N|K|mode|count|ms|flops" line each; returns bytes written */
size_t tml_gemm_timing_report(char* buf, size_t cap);
/* id of the mbarrier wait that timed out in the last failed GEMM launch (0 = none): 1 weight-tile slot,
 * 2 halo slot, 3 accumulator free, 4 halo ready, 5 weight tile ready, 6 accumulator ready; +100 on the peer CTA */
int tml_debug_last_hang(void);
/* 0 = tcgen05 kernel (default, the product path), 1 = SIMT debug kernel (tests only) */
void tml_debug_set_gemm_impl(int impl);
/* Generic implicit-GEMM entry used by the kernel unit tests (same operation the encoder issues). */
typedef struct TmlGemmDesc {
    const void* A; int A_C, A_W, A_H, A_B; int64_t A_sW, A_sH, A_sB;
    int stride, ntaps; int dh[9], dw[9]; int OW, OH;
    const void* Bm; int N; int64_t B_sN, B_sBatch;
    float alpha; const float* bias; const void* resid; int64_t R_sB, R_sH, R_sW;
    void* D; int out_fp32; int64_t D_sB, D_sH, D_sW, D_sN; int n_store; float beta;
    /* fused GroupNorm reductions of the output (see csrc/gemm.h): mode 1 = (sum, sumsq), 2 = backward sums */
    int gn_mode; float* gn_partial; const void* gn_x; const void* gn_ss; const void* gn_mr; const float* gn_gamma; int gn_silu;
    int dbg_shift, dbg_bo; /* hardware experiment: row-shifted UMMA descriptor (tests only) */
    /* fused input normalisation (CTA-pair 3x3 kernel): A is the raw GroupNorm input, in_gn_ss = [A_B][A_C] float2
     * (scale, shift); the kernel convolves silu(A * scale + shift) */
    const void* in_gn_ss;
    /* transposed A: stored [batch][k][m] with m = oh*OW + ow contiguous, row stride A_sK (csrc/gemm.h) */
    int a_trans; int64_t A_sK;
} TmlGemmDesc;
int tml_debug_gemm(const TmlGemmDesc* d, void* stream);
/* entries per image of the partial buffer a gn_mode GEMM writes: [B][tiles][32][2] floats */
int tml_debug_gn_tiles_per_image(int OH, int OW);
/* the same for one specific op (the operand-swapped 3x3 kernel reduces gn_mode 2 per 64-pixel segment) */
int tml_debug_gn_chunks_per_image(const TmlGemmDesc* d);
/* partial entries per image of the UNet's general GroupNorm kernels (gng_*): a function of the pixel count and the
   channel count only -- never of the batch -- so an image's statistics do not depend on the batch it shares */
int tml_debug_unet_gn_chunks_per_image(int HW, int C);
/* Offsets (bytes into `saved`) and [B,H,W,C] dims of the bf16 NHWC activations the forward keeps:
 * "conv_in", "resnet_h1"/"resnet_out" (index = resnet in forward order), "down_out", "attn_qkv",
 * "attn_P" ([B,tok,tok,1]), "attn_out". */
int tml_debug_saved_tensor(TmlEncoder* enc, const char* name, int index, size_t* offset, int dims[4]);
int tml_debug_decoder_saved_tensor(TmlEncoder* vae, const char* name, int index, size_t* offset, int dims[4]);
/* Every backward stage copies its output gradient (bf16 NHWC) into slot k of dev_buffer (NULL = off). */
void tml_debug_set_grad_dump(void* dev_buffer, size_t slot_bytes, int slots);
/* Host-only helpers (no CUDA calls) used by the CPU tests of the weight packing. */
int tml_debug_pack_conv3x3(const float* w /*[Co][Ci][3][3]*/, int Co, int Ci, int mode /*0 fwd s1, 1 dgrad s1, 2 fwd s2, 3..6 dgrad s2 parity (ph,pw)=(0,0),(0,1),(1,0),(1,1) for pad (0,1,0,1); 7 fwd s2 pad 1, 8..11 its parity dgrads*/,
                           uint16_t* out_bf16, int* ntaps, int* dh, int* dw);

#ifdef __cplusplus
}
#endif
#endif /* TML_B200_H */
