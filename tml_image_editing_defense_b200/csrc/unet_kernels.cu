// Bandwidth-bound kernels of the diffusion attack's UNet (main.py:229-243 -> diffusers UNet2DConditionModel):
// GroupNorm for any channel count (10 / 20 / 30 / 40 / 60 / 80 channels per group at 320 .. 2560 channels),
// LayerNorm, GEGLU, the head split / merge around the multi-head attention products, channel concatenation of
// the skip connections, and the timestep embedding.  Same rules as elementwise.cu: bf16 NHWC activations, fp32
// statistics, every reduction in a fixed order (bitwise reproducible, independent of the batch an image shares).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace tml {

void count_launch();   // elementwise.cu
#define COUNT_LAUNCH() count_launch()

__device__ __forceinline__ float u_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float u_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t u_pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void u_unpack8(const uint4& u, float (&f)[8]) {
    f[0] = u_lo(u.x); f[1] = u_hi(u.x); f[2] = u_lo(u.y); f[3] = u_hi(u.y);
    f[4] = u_lo(u.z); f[5] = u_hi(u.z); f[6] = u_lo(u.w); f[7] = u_hi(u.w);
}
__device__ __forceinline__ uint4 u_pack8(const float (&f)[8]) {
    return make_uint4(u_pack2(f[0], f[1]), u_pack2(f[2], f[3]), u_pack2(f[4], f[5]), u_pack2(f[6], f[7]));
}
__device__ __forceinline__ float u_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sigmoid(u) = 0.5 + 0.5 tanh(u / 2): one MUFU op (tanh.approx, relative error ~2^-11, far below the bf16 rounding of
// the results) instead of exp + reciprocal
__device__ __forceinline__ float u_sigmoid(float u) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float u_silu(float u) { return u * u_sigmoid(u); }
__device__ __forceinline__ float u_dsilu(float u) {
    const float sg = u_sigmoid(u);
    return fmaf(u * sg, 1.f - sg, sg);
}

// ================================================================================================
// GroupNorm (32 groups) for any C % 8 == 0 with C / 32 an integer: diffusers ResnetBlock2D.norm1/norm2 at the
// UNet's widths, Transformer2DModel.norm, conv_norm_out.
// A block owns a chunk of pixels of one image.  Thread = (channel octet, pixel lane): `span` = min(C/8, 256)
// octets side by side (consecutive threads read consecutive 16-byte vectors), 256 / span pixel lanes; wider tensors
// (C/8 > 256) walk the octets in steps of 256.  Per-channel sums go through shared memory and thread g < 32 adds
// the channels of its group in a fixed order, so a group may start anywhere inside an octet.
// ================================================================================================
// Pixels per block.  The UNet's GroupNorms run on small tensors (8 x 8 ... 64 x 64 latents at 320 ... 2560 channels), so
// what matters is how many blocks and how many loads are in flight, not the streaming rate of one block: a thread owns
// 4 ... 16 pixels of its channel octet (at least ~64 blocks per image), and the loops below issue kU independent
// 16-byte loads per tensor before they touch the data (a plain `#pragma unroll` over the guarded pixel loop keeps ONE
// load in flight per thread: measured 0.2 - 2 TB/s).  The chunking depends on (HW, C) only -- never on the batch --
// so an image's statistics are bit-identical whatever batch it shares.
static int gng_pl(int C) { const int C8 = C >> 3, span = C8 < 256 ? C8 : 256; return 256 / span; }
static int gng_ppc(int HW, int C) {
    const int PL = gng_pl(C);
    int per_thread = HW / (PL * 64);
    per_thread = per_thread < 4 ? 4 : (per_thread > 16 ? 16 : per_thread);
    return PL * per_thread;
}
int gng_num_chunks(int HW, int C) { const int p = gng_ppc(HW, C); return (HW + p - 1) / p; }
static size_t gng_smem(int C) { return (size_t)gng_pl(C) * C * 2 * sizeof(float); }
constexpr int kU = 4;   // independent 16-byte loads in flight per thread and tensor

// Sum of the 8 lanes that share a group (lanes 8k .. 8k+7), fixed tree.
__device__ __forceinline__ float u_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__global__ void __launch_bounds__(256) gng_stats_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int HW,
                                                        int C, int ppc) {
    extern __shared__ float sm[];   // [PL][C][2]
    const int C8 = C >> 3, span = C8 < 256 ? C8 : 256, PL = 256 / span;
    const int o0 = threadIdx.x % span, l = threadIdx.x / span;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
    if (l < PL) {
        for (int o = o0; o < C8; o += span) {
            float s[8], q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
            const bf16* base = x + (size_t)b * HW * C + (size_t)o * 8;
            for (int p = p0 + l; p < p1; p += PL * kU) {
                uint4 u[kU];
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (p + k * PL < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + k * PL) * C));
#pragma unroll
                for (int k = 0; k < kU; ++k) {
                    if (p + k * PL >= p1) break;
                    float f[8];
                    u_unpack8(u[k], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sm[((size_t)l * C + o * 8 + j) * 2] = s[j];
                sm[((size_t)l * C + o * 8 + j) * 2 + 1] = q[j];
            }
        }
    }
    __syncthreads();
    {   // eight threads per group: thread (g, sub) adds channels g*cpg + sub, + 8, ... in order, then a fixed tree
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7, cpg = C / 32;
        float s = 0.f, q = 0.f;
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += 8)
            for (int ll = 0; ll < PL; ++ll) {
                s += sm[((size_t)ll * C + c) * 2];
                q += sm[((size_t)ll * C + c) * 2 + 1];
            }
        s = u_sum8(s);
        q = u_sum8(q);
        if (sub == 0) {
            float* out = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
            out[0] = s;
            out[1] = q;
        }
    }
}

void launch_gng_stats(const bf16* x, float* partial, int B, int HW, int C, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gng_num_chunks(HW, C), B);
    gng_stats_kernel<<<grid, 256, gng_smem(C), s>>>(x, partial, HW, C, gng_ppc(HW, C));
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) gng_apply_kernel(const bf16* __restrict__ x, const float2* __restrict__ ss,
                                                        bf16* __restrict__ y, int HW, int C, int ppc, int silu) {
    const int C8 = C >> 3, span = C8 < 256 ? C8 : 256, PL = 256 / span;
    const int o0 = threadIdx.x % span, l = threadIdx.x / span;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
    if (l >= PL) return;
    for (int o = o0; o < C8; o += span) {
        float sc[8], sh[8];
        {
            const float4* sp = reinterpret_cast<const float4*>(ss + (size_t)b * C + o * 8);   // 8 (scale, shift) pairs
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = __ldg(sp + j);
                sc[2 * j] = v.x; sh[2 * j] = v.y; sc[2 * j + 1] = v.z; sh[2 * j + 1] = v.w;
            }
        }
        const size_t base = (size_t)b * HW * C + (size_t)o * 8;
        for (int p = p0 + l; p < p1; p += PL * kU) {
            uint4 u[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (p + k * PL < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + k * PL) * C));
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                if (p + k * PL >= p1) break;
                float f[8];
                u_unpack8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = fmaf(f[j], sc[j], sh[j]);
                    f[j] = silu ? u_silu(v) : v;
                }
                *reinterpret_cast<uint4*>(y + base + (size_t)(p + k * PL) * C) = u_pack8(f);
            }
        }
    }
}

void launch_gng_apply(const bf16* x, const float2* ss, bf16* y, int B, int HW, int C, int silu, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gng_num_chunks(HW, C), B);
    gng_apply_kernel<<<grid, 256, 0, s>>>(x, ss, y, HW, C, gng_ppc(HW, C), silu);
    COUNT_LAUNCH();
}

// partial[b][chunk][g] = (sum dxh, sum dxh*xh), dxh = dy*act'(u)*gamma, xh = (x-mean)*rstd.  Accumulated per channel
// as S1 = sum dy*act'(u), S2 = sum dy*act'(u)*x; the group pass forms gamma*S1 and gamma*rstd*(S2 - mean*S1).
__global__ void __launch_bounds__(256) gng_bwd_partial_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              const float2* __restrict__ ss,
                                                              const float2* __restrict__ mr,
                                                              const float* __restrict__ gamma,
                                                              float* __restrict__ partial, int HW, int C, int ppc,
                                                              int silu) {
    extern __shared__ float sm[];   // [PL][C][2]
    const int C8 = C >> 3, span = C8 < 256 ? C8 : 256, PL = 256 / span;
    const int o0 = threadIdx.x % span, l = threadIdx.x / span;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
    if (l < PL) {
        for (int o = o0; o < C8; o += span) {
            float sc[8], sh[8], s1[8], s2[8];
            {
                const float4* sp = reinterpret_cast<const float4*>(ss + (size_t)b * C + o * 8);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = __ldg(sp + j);
                    sc[2 * j] = v.x; sh[2 * j] = v.y; sc[2 * j + 1] = v.z; sh[2 * j + 1] = v.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
            const size_t base = (size_t)b * HW * C + (size_t)o * 8;
            for (int p = p0 + l; p < p1; p += PL * kU) {
                uint4 ux[kU], ud[kU];
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (p + k * PL < p1) {
                        ux[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + k * PL) * C));
                        ud[k] = __ldg(reinterpret_cast<const uint4*>(dy + base + (size_t)(p + k * PL) * C));
                    }
#pragma unroll
                for (int k = 0; k < kU; ++k) {
                    if (p + k * PL >= p1) break;
                    float fx[8], fd[8];
                    u_unpack8(ux[k], fx);
                    u_unpack8(ud[k], fd);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float u = fmaf(fx[j], sc[j], sh[j]);
                        const float d = silu ? fd[j] * u_dsilu(u) : fd[j];
                        s1[j] += d;
                        s2[j] = fmaf(d, fx[j], s2[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sm[((size_t)l * C + o * 8 + j) * 2] = s1[j];
                sm[((size_t)l * C + o * 8 + j) * 2 + 1] = s2[j];
            }
        }
    }
    __syncthreads();
    {   // eight threads per group (see gng_stats_kernel)
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7, cpg = C / 32;
        const float2 m = __ldg(&mr[(size_t)b * 32 + g]);
        float a = 0.f, q = 0.f;
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += 8) {
            float s1 = 0.f, s2 = 0.f;
            for (int ll = 0; ll < PL; ++ll) {
                s1 += sm[((size_t)ll * C + c) * 2];
                s2 += sm[((size_t)ll * C + c) * 2 + 1];
            }
            const float gm = __ldg(&gamma[c]);
            a = fmaf(gm, s1, a);
            q = fmaf(gm * m.y, s2 - m.x * s1, q);
        }
        a = u_sum8(a);
        q = u_sum8(q);
        if (sub == 0) {
            float* out = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
            out[0] = a;
            out[1] = q;
        }
    }
}

void launch_gng_bwd_partial(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float* gamma,
                            float* partial, int B, int HW, int C, int silu, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gng_num_chunks(HW, C), B);
    gng_bwd_partial_kernel<<<grid, 256, gng_smem(C), s>>>(x, dy, ss, mr, gamma, partial, HW, C, gng_ppc(HW, C), silu);
    COUNT_LAUNCH();
}

// dx = rstd*(dxh - m1 - xh*m2) [+ resid] = sc * act'(u) * dy + cx * x + c0 [+ resid]   (sc = rstd*gamma,
// cx = -rstd^2 m2, c0 = rstd (rstd*mean*m2 - m1); mm = (m1, m2) from gn_bwd_finalize)
__global__ void __launch_bounds__(256) gng_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            const float2* __restrict__ ss,
                                                            const float2* __restrict__ mr,
                                                            const float2* __restrict__ mm,
                                                            const bf16* __restrict__ resid, bf16* __restrict__ dx,
                                                            int HW, int C, int ppc, int silu) {
    const int C8 = C >> 3, span = C8 < 256 ? C8 : 256, PL = 256 / span, cpg = C / 32;
    const int o0 = threadIdx.x % span, l = threadIdx.x / span;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
    if (l >= PL) return;
    for (int o = o0; o < C8; o += span) {
        float sc[8], sh[8], cx[8], c0[8];
        {
            const float4* sp = reinterpret_cast<const float4*>(ss + (size_t)b * C + o * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = __ldg(sp + j);
                sc[2 * j] = v.x; sh[2 * j] = v.y; sc[2 * j + 1] = v.z; sh[2 * j + 1] = v.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = o * 8 + j;
            const float2 m = __ldg(&mr[(size_t)b * 32 + c / cpg]);
            const float2 k = __ldg(&mm[(size_t)b * 32 + c / cpg]);
            cx[j] = -m.y * m.y * k.y;
            c0[j] = m.y * (m.y * m.x * k.y - k.x);
        }
        const size_t base = (size_t)b * HW * C + (size_t)o * 8;
        for (int p = p0 + l; p < p1; p += PL * kU) {
            uint4 ux[kU], ud[kU], ur[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (p + k * PL < p1) {
                    ux[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + k * PL) * C));
                    ud[k] = __ldg(reinterpret_cast<const uint4*>(dy + base + (size_t)(p + k * PL) * C));
                    if (resid != nullptr)
                        ur[k] = __ldg(reinterpret_cast<const uint4*>(resid + base + (size_t)(p + k * PL) * C));
                }
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                if (p + k * PL >= p1) break;
                float fx[8], fd[8], fr[8], out[8];
                u_unpack8(ux[k], fx);
                u_unpack8(ud[k], fd);
                if (resid != nullptr) u_unpack8(ur[k], fr);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float u = fmaf(fx[j], sc[j], sh[j]);
                    const float d = silu ? fd[j] * u_dsilu(u) : fd[j];
                    float v = fmaf(sc[j], d, fmaf(cx[j], fx[j], c0[j]));
                    if (resid != nullptr) v += fr[j];
                    out[j] = v;
                }
                *reinterpret_cast<uint4*>(dx + base + (size_t)(p + k * PL) * C) = u_pack8(out);
            }
        }
    }
}

void launch_gng_bwd_apply(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float2* mm,
                          const bf16* resid, bf16* dx, int B, int HW, int C, int silu, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gng_num_chunks(HW, C), B);
    gng_bwd_apply_kernel<<<grid, 256, 0, s>>>(x, dy, ss, mr, mm, resid, dx, HW, C, gng_ppc(HW, C), silu);
    COUNT_LAUNCH();
}

// ================================================================================================
// LayerNorm over the channel dimension of [rows][C] (BasicTransformerBlock.norm1/2/3, eps 1e-5, affine).
// One warp per row; a lane keeps its octets (C/8 of them over 32 lanes, at most 8 each: C <= 2048) in registers,
// mean and variance are two passes over the registers.  stats[row] = (mean, rstd) for the backward.
// ================================================================================================
constexpr int kLnMaxOct = 8;   // widest instantiation: C <= 2048 (the register arrays are sized per instantiation)

// R rows per warp: all of a warp's loads (R rows x MAXO octets per lane, times the tensors) are issued before the first
// reduction, so a lane keeps R x MAXO (x 2-3) 16-byte loads in flight instead of one row's worth.
template <int MAXO, int R>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float2* __restrict__ stats, long long rows, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
    if (row0 >= rows) return;
    const int C8 = C >> 3;
    uint4 v[R][MAXO];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8 && row0 + r < rows) v[r][i] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(row0 + r) * C + (size_t)o * 8));
        }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long row = row0 + r;
        if (row >= rows) break;   // (warp-uniform)
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8) {
                float f[8];
                u_unpack8(v[r][i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[j];
            }
        }
        const float mean = u_warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8) {
                float f[8];
                u_unpack8(v[r][i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; q = fmaf(d, d, q); }
            }
        }
        const float rstd = rsqrtf(u_warp_sum(q) / (float)C + eps);
        if (lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8) {
                float f[8];
                u_unpack8(v[r][i], f);
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8) + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + o * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + o * 8) + 1);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf((f[j] - mean) * rstd, gg[j], bb[j]);
                *reinterpret_cast<uint4*>(y + (size_t)row * C + (size_t)o * 8) = u_pack8(f);
            }
        }
    }
}

void launch_ln_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float2* stats, long long rows, int C,
                   float eps, cudaStream_t s) {
    if (g_dry_run) return;
    const int per_lane = ((C >> 3) + 31) / 32;   // octets per lane
    // (one row per warp: with a single input tensor two rows per warp measured no faster -- 3.27 vs 3.08 ms per step under ncu)
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (per_lane <= 2) ln_fwd_kernel<2, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, C, eps);
    else if (per_lane <= 3) ln_fwd_kernel<3, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, C, eps);
    else if (per_lane <= 5) ln_fwd_kernel<5, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, C, eps);
    else ln_fwd_kernel<kLnMaxOct, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, C, eps);
    COUNT_LAUNCH();
}

// dx = rstd * (g - mean(g) - xh * mean(g * xh)) [+ resid],  g = dy * gamma,  xh = (x - mean) * rstd
template <int MAXO, int R>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                     const float* __restrict__ gamma, const float2* __restrict__ stats,
                                                     const bf16* __restrict__ resid, bf16* __restrict__ dx,
                                                     long long rows, int C) {
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
    if (row0 >= rows) return;
    const int C8 = C >> 3;
    // x, dy (and the residual) stay packed in registers between the two passes (gamma is re-read as two float4 per
    // octet: L1 hits)
    uint4 vx[R][MAXO], vd[R][MAXO], vr[R][MAXO];
    float2 st[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row0 + r < rows) st[r] = stats[row0 + r];
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8 && row0 + r < rows) {
                const size_t at = (size_t)(row0 + r) * C + (size_t)o * 8;
                vx[r][i] = __ldg(reinterpret_cast<const uint4*>(x + at));
                vd[r][i] = __ldg(reinterpret_cast<const uint4*>(dy + at));
                if (resid != nullptr) vr[r][i] = __ldg(reinterpret_cast<const uint4*>(resid + at));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long row = row0 + r;
        if (row >= rows) break;   // (warp-uniform)
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8) + 1);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                float fx[8], fd[8];
                u_unpack8(vx[r][i], fx);
                u_unpack8(vd[r][i], fd);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float g = fd[j] * gg[j];
                    a += g;
                    b = fmaf(g, (fx[j] - st[r].x) * st[r].y, b);
                }
            }
        }
        const float m1 = u_warp_sum(a) / (float)C, m2 = u_warp_sum(b) / (float)C;
#pragma unroll
        for (int i = 0; i < MAXO; ++i) {
            const int o = lane + 32 * i;
            if (o < C8) {
                float fx[8], fd[8], fr[8], out[8];
                u_unpack8(vx[r][i], fx);
                u_unpack8(vd[r][i], fd);
                if (resid != nullptr) u_unpack8(vr[r][i], fr);
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8) + 1);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float g = fd[j] * gg[j];
                    float v = st[r].y * (g - m1 - (fx[j] - st[r].x) * st[r].y * m2);
                    if (resid != nullptr) v += fr[j];
                    out[j] = v;
                }
                *reinterpret_cast<uint4*>(dx + (size_t)row * C + (size_t)o * 8) = u_pack8(out);
            }
        }
    }
}

void launch_ln_bwd(const bf16* x, const bf16* dy, const float* gamma, const float2* stats, const bf16* resid, bf16* dx,
                   long long rows, int C, cudaStream_t s) {
    if (g_dry_run) return;
    const int per_lane = ((C >> 3) + 31) / 32;
    const unsigned grid = (unsigned)((rows + 7) / 8), grid2 = (unsigned)((rows + 15) / 16);
    if (per_lane <= 2) ln_bwd_kernel<2, 2><<<grid2, 256, 0, s>>>(x, dy, gamma, stats, resid, dx, rows, C);
    else if (per_lane <= 3) ln_bwd_kernel<3, 2><<<grid2, 256, 0, s>>>(x, dy, gamma, stats, resid, dx, rows, C);
    else if (per_lane <= 5) ln_bwd_kernel<5, 1><<<grid, 256, 0, s>>>(x, dy, gamma, stats, resid, dx, rows, C);
    else ln_bwd_kernel<kLnMaxOct, 1><<<grid, 256, 0, s>>>(x, dy, gamma, stats, resid, dx, rows, C);
    COUNT_LAUNCH();
}

// ================================================================================================
// GEGLU (FeedForward.net[0]):  h = [x | gate] of width 2I;  out = x * gelu(gate), exact (erf) gelu as F.gelu.
// ================================================================================================
__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float g) {
    return 0.5f * (1.f + erff(g * 0.70710678118654752f)) + g * 0.3989422804014327f * __expf(-0.5f * g * g);
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ out,
                                                        long long rows, int I) {
    const unsigned I8 = unsigned(I) >> 3;
    const long long total = rows * I8;
    uint4 ux[2], ug[2];
    size_t at[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {   // two (value, gate) pairs of loads in flight per thread
        const long long t = (long long)blockIdx.x * 512 + k * 256 + threadIdx.x;
        if (t >= total) continue;
        // (row index by 32-bit division whenever the vector count allows it)
        const long long row = total < (1ll << 32) ? (long long)(unsigned(t) / I8) : t / I8;
        const unsigned o = unsigned(t - row * I8);
        ux[k] = __ldg(reinterpret_cast<const uint4*>(h + (size_t)row * 2 * I + (size_t)o * 8));
        ug[k] = __ldg(reinterpret_cast<const uint4*>(h + (size_t)row * 2 * I + I + (size_t)o * 8));
        at[k] = (size_t)row * I + (size_t)o * 8;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long t = (long long)blockIdx.x * 512 + k * 256 + threadIdx.x;
        if (t >= total) continue;
        float fx[8], fg[8];
        u_unpack8(ux[k], fx);
        u_unpack8(ug[k], fg);
#pragma unroll
        for (int j = 0; j < 8; ++j) fx[j] *= gelu_f(fg[j]);
        *reinterpret_cast<uint4*>(out + at[k]) = u_pack8(fx);
    }
}

void launch_geglu_fwd(const bf16* h, bf16* out, long long rows, int I, cudaStream_t s) {
    if (g_dry_run) return;
    const long long n = rows * (I >> 3);
    geglu_fwd_kernel<<<(unsigned)((n + 511) / 512), 256, 0, s>>>(h, out, rows, I);
    COUNT_LAUNCH();
}

// dh[:, :I] = dout * gelu(gate);  dh[:, I:] = dout * x * gelu'(gate)
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ dout,
                                                        bf16* __restrict__ dh, long long rows, int I) {
    const int I8 = I >> 3;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= rows * I8) return;
    const long long row = t / I8;
    const int o = (int)(t - row * I8);
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(h + (size_t)row * 2 * I + (size_t)o * 8));
    const uint4 ug = __ldg(reinterpret_cast<const uint4*>(h + (size_t)row * 2 * I + I + (size_t)o * 8));
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)row * I + (size_t)o * 8));
    float fx[8], fg[8], fd[8], dx[8], dg[8];
    u_unpack8(ux, fx);
    u_unpack8(ug, fg);
    u_unpack8(ud, fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        dx[j] = fd[j] * gelu_f(fg[j]);
        dg[j] = fd[j] * fx[j] * dgelu_f(fg[j]);
    }
    *reinterpret_cast<uint4*>(dh + (size_t)row * 2 * I + (size_t)o * 8) = u_pack8(dx);
    *reinterpret_cast<uint4*>(dh + (size_t)row * 2 * I + I + (size_t)o * 8) = u_pack8(dg);
}

void launch_geglu_bwd(const bf16* h, const bf16* dout, bf16* dh, long long rows, int I, cudaStream_t s) {
    if (g_dry_run) return;
    const long long n = rows * (I >> 3);
    geglu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h, dout, dh, rows, I);
    COUNT_LAUNCH();
}

// ================================================================================================
// Multi-head attention plumbing.  The attention products run on the tcgen05 GEMM kernels as batched GEMMs over
// (image, head) pairs, which want every head as a dense [tokens][dpad] matrix with dpad a multiple of 64 (one
// 128-byte swizzle row per 64 channels): head_split copies head h's d channels out of a [B][tok][ld] projection and
// pads them with zeros; head_merge is the inverse (drops the padding).
// Masking of padded key tokens (cross attention: 77 prompt tokens in a tile of 128) needs no kernel support: the
// first padding channel of every QUERY row is 1 and that of a padded KEY row is -30000, so a padded logit is
// -30000 and its softmax numerator underflows to exactly 0; real key rows carry 0 there.
//   fill: 0 = zeros, 1 = query rows (slot d = 1), 2 = key rows (slot d = -30000 on rows >= tok_valid)
// ================================================================================================
// grid = (vectors of one (image, head) / 1024, heads, images): no 64-bit division per thread, four independent
// 16-byte loads in flight per thread (one per thread left these copies at 1.2 - 2.1 TB/s)
__global__ void __launch_bounds__(256) head_split_kernel(const bf16* __restrict__ in, long long ld_in, long long bs_in,
                                                         int col0, bf16* __restrict__ out, int tok_src,
                                                         int tok_valid, int tok_out, int d, int dpad, int fill) {
    const unsigned P8 = unsigned(dpad) >> 3, total = unsigned(tok_out) * P8;
    const int h = blockIdx.y, heads = gridDim.y;
    const size_t b = blockIdx.z;
    const bf16* src = in + b * bs_in + col0 + h * d;
    bf16* dst = out + ((b * heads + h) * tok_out) * dpad;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned t = blockIdx.x * 1024u + k * 256u + threadIdx.x;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (t >= total) continue;
        const unsigned tk = t / P8, oj = t - tk * P8;
        if (int(oj * 8) < d) {
            if (int(tk) < tok_valid && int(tk) < tok_src)
                v[k] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)tk * ld_in + oj * 8));
        } else if (int(oj * 8) == d) {
            if (fill == 1) v[k].x = 0x3F80u;                               // bf16 1.0 in the low half
            else if (fill == 2 && int(tk) >= tok_valid) v[k].x = 0xC6EAu;   // bf16 -29952
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned t = blockIdx.x * 1024u + k * 256u + threadIdx.x;
        if (t < total) *reinterpret_cast<uint4*>(dst + (size_t)t * 8) = v[k];
    }
}

void launch_head_split(const bf16* in, long long ld_in, long long bs_in, int col0, bf16* out, int B, int heads,
                       int tok_src, int tok_valid, int tok_out, int d, int dpad, int fill, cudaStream_t s) {
    if (g_dry_run) return;
    const unsigned per = (unsigned)tok_out * (unsigned)(dpad >> 3);
    head_split_kernel<<<dim3((per + 1023) / 1024, heads, B), 256, 0, s>>>(in, ld_in, bs_in, col0, out, tok_src, tok_valid,
                                                                         tok_out, d, dpad, fill);
    COUNT_LAUNCH();
}

// out[b][t][col0 + h*d + j] = in[(b*heads + h)][t][j], j < d.   grid = (vectors of one image / 256, images); the
// writes of consecutive threads are consecutive (octets of a head, then the next head)
__global__ void __launch_bounds__(256) head_merge_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                         long long ld_out, long long bs_out, int col0, int heads,
                                                         int tok, int d, int dpad) {
    const unsigned D8 = unsigned(d) >> 3, total = unsigned(tok) * unsigned(heads) * D8;
    const size_t b = blockIdx.y;
    uint4 v[4];
    size_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned t = blockIdx.x * 1024u + k * 256u + threadIdx.x;
        if (t >= total) continue;
        const unsigned r = t / D8, oj = t - r * D8;
        const unsigned tk = r / unsigned(heads), h = r - tk * unsigned(heads);
        v[k] = __ldg(reinterpret_cast<const uint4*>(in + ((b * heads + h) * tok + tk) * dpad + oj * 8));
        off[k] = b * bs_out + (size_t)tk * ld_out + col0 + h * d + oj * 8;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned t = blockIdx.x * 1024u + k * 256u + threadIdx.x;
        if (t < total) *reinterpret_cast<uint4*>(out + off[k]) = v[k];
    }
}

void launch_head_merge(const bf16* in, bf16* out, long long ld_out, long long bs_out, int col0, int B, int heads,
                       int tok, int d, int dpad, cudaStream_t s) {
    if (g_dry_run) return;
    const unsigned per = (unsigned)tok * (unsigned)heads * (unsigned)(d >> 3);
    head_merge_kernel<<<dim3((per + 1023) / 1024, B), 256, 0, s>>>(in, out, ld_out, bs_out, col0, heads, tok, d, dpad);
    COUNT_LAUNCH();
}

// ================================================================================================
// Column-block copy (skip-connection concatenation and its split in the backward), elementwise add.
// ================================================================================================
// out[r][oc0 + j] = in[r][ic0 + j],  j < ncols (ncols, offsets, strides multiples of 8).  Index = 32-bit when the
// vector count allows it (always, at the UNet's sizes): a 64-bit division per 16-byte copy is most of the kernel.
template <typename I>
__global__ void __launch_bounds__(256) copy_cols_kernel(const bf16* __restrict__ in, long long ld_in, int ic0,
                                                        bf16* __restrict__ out, long long ld_out, int oc0, int ncols,
                                                        long long total) {
    const I N8 = (I)(ncols >> 3);
    uint4 v[4];
    size_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // four independent loads in flight per thread
        const I t = (I)blockIdx.x * 1024 + k * 256 + threadIdx.x;
        if ((long long)t >= total) continue;
        const I r = t / N8;
        const I o = t - r * N8;
        v[k] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)r * ld_in + ic0 + (size_t)o * 8));
        off[k] = (size_t)r * ld_out + oc0 + (size_t)o * 8;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const I t = (I)blockIdx.x * 1024 + k * 256 + threadIdx.x;
        if ((long long)t < total) *reinterpret_cast<uint4*>(out + off[k]) = v[k];
    }
}

void launch_copy_cols(const bf16* in, long long ld_in, int ic0, bf16* out, long long ld_out, int oc0, int ncols,
                      long long rows, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = rows * (ncols >> 3);
    const unsigned blocks = (unsigned)((total + 1023) / 1024);
    if (total < (1ll << 31) - 1024)
        copy_cols_kernel<unsigned><<<blocks, 256, 0, s>>>(in, ld_in, ic0, out, ld_out, oc0, ncols, total);
    else
        copy_cols_kernel<long long><<<blocks, 256, 0, s>>>(in, ld_in, ic0, out, ld_out, oc0, ncols, total);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                       bf16* __restrict__ out, long long n8) {
    uint4 va[2], vb[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long t = (long long)blockIdx.x * 512 + k * 256 + threadIdx.x;
        if (t < n8) { va[k] = __ldg(reinterpret_cast<const uint4*>(a) + t); vb[k] = __ldg(reinterpret_cast<const uint4*>(b) + t); }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long t = (long long)blockIdx.x * 512 + k * 256 + threadIdx.x;
        if (t >= n8) continue;
        float fa[8], fb[8];
        u_unpack8(va[k], fa);
        u_unpack8(vb[k], fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] += fb[j];
        reinterpret_cast<uint4*>(out)[t] = u_pack8(fa);
    }
}

void launch_add_bf16(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t s) {
    if (g_dry_run) return;
    const long long n8 = n >> 3;
    add_bf16_kernel<<<(unsigned)((n8 + 511) / 512), 256, 0, s>>>(a, b, out, n8);
    COUNT_LAUNCH();
}

// ================================================================================================
// Timestep embedding.  diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin] of
// t * 10000^(-i/half); TimestepEmbedding = linear -> SiLU -> linear; every ResnetBlock2D adds
// time_emb_proj(SiLU(temb)) to conv1's output -- one value per channel, the same for every image (the reference
// passes one scalar t, main.py:233-238), so it is folded into conv1's bias vector here.
// small_linear: y[n] = bias[n] + sum_k W[n][k] * act(x[k]), fp32, one warp per output.
// ================================================================================================
__global__ void timestep_embed_kernel(float t, float* __restrict__ out, int dim) {
    const int half = dim >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const float f = expf(-9.210340371976184f * (float)i / (float)half);
    const float a = t * f;
    out[i] = cosf(a);
    out[half + i] = sinf(a);
}

void launch_timestep_embed(float t, float* out, int dim, cudaStream_t s) {
    if (g_dry_run) return;
    timestep_embed_kernel<<<(dim / 2 + 127) / 128, 128, 0, s>>>(t, out, dim);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int N, int K, int silu_in) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        float v = __ldg(&x[k]);
        if (silu_in) v = v / (1.f + expf(-v));
        acc = fmaf(__ldg(&W[(size_t)n * K + k]), v, acc);
    }
    acc = u_warp_sum(acc);
    if (lane == 0) y[n] = acc + (bias ? bias[n] : 0.f);
}

void launch_small_linear(const float* x, const float* W, const float* bias, float* y, int N, int K, int silu_in,
                         cudaStream_t s) {
    if (g_dry_run) return;
    small_linear_kernel<<<(N + 7) / 8, 256, 0, s>>>(x, W, bias, y, N, K, silu_in);
    COUNT_LAUNCH();
}

// fp32 NCHW [B][C][hw] (C <= 64) -> bf16 NHWC [B][hw][64], channels >= C zero; and back (first C channels).
__global__ void __launch_bounds__(256) nchw_pack64_kernel(const float* __restrict__ x, bf16* __restrict__ out, int C,
                                                          int hw, long long total_px) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long px = t >> 3;
    if (px >= total_px) return;
    const int oct = (int)(t & 7);
    const long long b = px / hw, p = px - b * hw;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = oct * 8 + j;
        f[j] = c < C ? __ldg(x + ((size_t)b * C + c) * hw + p) : 0.f;
    }
    *reinterpret_cast<uint4*>(out + (size_t)px * 64 + oct * 8) = u_pack8(f);
}

void launch_nchw_pack64(const float* x, bf16* out, int B, int C, int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total_px = (long long)B * hw;
    nchw_pack64_kernel<<<(unsigned)((total_px * 8 + 255) / 256), 256, 0, s>>>(x, out, C, hw, total_px);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) nhwc64_unpack_kernel(const bf16* __restrict__ in, float* __restrict__ out, int C,
                                                            int hw, long long total) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;   // (b, c, p), p fastest
    if (t >= total) return;
    const long long p = t % hw;
    const long long bc = t / hw;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    out[t] = __bfloat162float(in[((size_t)b * hw + p) * 64 + c]);
}

void launch_nhwc64_unpack(const bf16* in, float* out, int B, int C, int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * C * hw;
    nhwc64_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, C, hw, total);
    COUNT_LAUNCH();
}

}  // namespace tml
