// Bandwidth-bound kernels of the PGD hot path: conv_in pack (hi/lo bf16 split of the fp32 image) / col2im, GroupNorm
// (+SiLU) forward / backward, attention row helpers and transposes, posterior sample + latent loss
// gradient, and the fused PGD updates.  All reductions are two-stage with a fixed order, so results
// are bitwise reproducible run to run and independent of how images are sharded over GPUs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include <atomic>

#include "kernels.h"

namespace tml {

thread_local bool g_dry_run = false;
static std::atomic<long> g_launches{0};
long kernel_launch_count() { return g_launches.load(); }
#define COUNT_LAUNCH() g_launches.fetch_add(1)
void count_launch() { g_launches.fetch_add(1); }

constexpr int kNumSMs = 148;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf_lo(u.x); f[1] = bf_hi(u.x); f[2] = bf_lo(u.y); f[3] = bf_hi(u.y);
    f[4] = bf_lo(u.z); f[5] = bf_hi(u.z); f[6] = bf_lo(u.w); f[7] = bf_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum broadcast to every thread (blockDim.x multiple of 32, <= 1024); fixed order.
__device__ __forceinline__ double block_sum_d(double v, double* red /*[33]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ================================================================================================
// conv_in forward, step 1: fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64]   (diffusers encoder.conv_in,
// reached from main.py:191).  Channels: [0,3) hi = bf16(x), [3,6) lo = bf16(x - hi), [6,64) zero.
// The 3x3 tensor-core convolution that follows carries the same bf16 weights on channels c and c+3, so the image
// enters with ~16 mantissa bits: a PGD step of 2/255 on a pixel near 1.0 is one bf16 ulp and would vanish with hi
// alone.  64 channels = one 128-byte swizzle row = the minimum K chunk of the GEMM kernels.
// Thread = (pixel, 16-byte octet): octet 0 carries the data, octets 1..7 are zeros; a warp writes 512 contiguous bytes.
// ================================================================================================
__global__ void __launch_bounds__(256) conv_in_pack_kernel(const float* __restrict__ x, bf16* __restrict__ a, int HW,
                                                           long long total_px) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long px = t >> 3;
    if (px >= total_px) return;
    const int oct = (int)(t & 7);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (oct == 0) {
        const long long b = px / HW, p = px - b * HW;
        const float* src = x + (size_t)b * 3 * HW + p;
        uint32_t hi[3], lo[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(src + (size_t)c * HW);
            const bf16 h = __float2bfloat16_rn(v);
            hi[c] = __bfloat16_as_ushort(h);
            lo[c] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
        }
        o = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (lo[0] << 16), lo[1] | (lo[2] << 16), 0u);
    }
    *reinterpret_cast<uint4*>(a + (size_t)px * 64 + oct * 8) = o;
}

void launch_conv_in_pack(const float* x, bf16* a, int B, int H, int W, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total_px = (long long)B * H * W;
    conv_in_pack_kernel<<<(unsigned)((total_px * 8 + 255) / 256), 256, 0, s>>>(x, a, H * W, total_px);
    COUNT_LAUNCH();
}

// conv_in input gradient, step 2 (col2im): y [B][H][W][32] fp32 tap products, entry (r*3+s)*3+ci of pixel p is
// sum_co dy[p][co] * W[co][ci][r][s];  dx[ci][h][w] = sum_{r,s} y[(h-r+1, w-s+1)][(r*3+s)*3+ci].
// Block = 64 pixels of one row: the 3 x 66 neighbouring lines are staged in shared memory with coalesced 16-byte
// loads, then thread = (pixel, channel) sums its nine entries.
__global__ void __launch_bounds__(256) conv_in_col2im_kernel(const float* __restrict__ y, float* __restrict__ dx, int H,
                                                             int W, float beta) {
    __shared__ float sy[3][66][33];   // [row h-1..h+1][pixel w0-1..w0+64][32 floats + 1 pad: conflict-free column reads]
    const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 3 * 66 * 8; i += 256) {
        const int v = i & 7, col = (i >> 3) % 66, rr = i / (66 * 8);
        const int hh = h + rr - 1, ww = w0 + col - 1;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            t = __ldg(reinterpret_cast<const float4*>(y + (((size_t)b * H + hh) * W + ww) * 32) + v);
        float* dst = &sy[rr][col][4 * v];
        dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
    }
    __syncthreads();
    const int px = threadIdx.x & 63, ci = threadIdx.x >> 6;   // 64 pixels x (3 channels + one idle quarter)
    if (ci >= 3 || w0 + px >= W) return;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            // source pixel (h - r + 1, w - s + 1) -> staged row (2 - r), column px + 1 - s + 1
            acc += sy[2 - r][px + 2 - s][(r * 3 + s) * 3 + ci];
        }
    float* d = dx + ((size_t)(b * 3 + ci) * H + h) * W + w0 + px;
    *d = beta != 0.f ? fmaf(beta, *d, acc) : acc;
}

void launch_conv_in_col2im(const float* y, float* dx, int B, int H, int W, float beta, cudaStream_t s) {
    if (g_dry_run) return;
    conv_in_col2im_kernel<<<dim3((W + 63) / 64, H, B), 256, 0, s>>>(y, dx, H, W, beta);
    COUNT_LAUNCH();
}

// ================================================================================================
// GroupNorm (32 groups, affine) [+ SiLU] over bf16 NHWC.   diffusers ResnetBlock2D.norm1/norm2,
// Attention.group_norm, Encoder.conv_norm_out.
// Thread layout shared by the stats / apply / backward kernels: a block owns a contiguous chunk of
// pixels of one image; thread = (channel octet, pixel lane); consecutive threads read consecutive
// 16-byte vectors, so every access is a full 128-byte line.
// ================================================================================================
// Pixels per block: every thread owns ~16 pixels of one channel octet, whatever C is, so the small
// 64x64 layers still launch ~1000 blocks (a fixed 512-pixel chunk left them at < 1 block per SM).
__host__ __device__ inline int gn_pix_per_chunk(int C) { return 32768 / C; }
int gn_num_chunks(int HW, int C) { const int p = gn_pix_per_chunk(C); return (HW + p - 1) / p; }

// partial[b][chunk][g] = (sum, sumsq) over the chunk's pixels of group g
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int HW,
                                                       int C) {
    __shared__ float red[256][4];
    const int C8 = C >> 3, PL = 256 / C8;
    const int oct = threadIdx.x % C8, pl = threadIdx.x / C8;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int p0 = chunk * gn_pix_per_chunk(C), p1 = min(HW, p0 + gn_pix_per_chunk(C));
    float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
    const bf16* base = x + (size_t)b * HW * C + (size_t)oct * 8;
    constexpr int U = 4;
    for (int p = p0 + pl; p < p1; p += PL * U) {
        uint4 u[U];
#pragma unroll
        for (int k = 0; k < U; ++k)
            u[k] = (p + k * PL < p1) ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + k * PL) * C))
                                     : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < U; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_lo += f[j]; q_lo = fmaf(f[j], f[j], q_lo); }
#pragma unroll
            for (int j = 4; j < 8; ++j) { s_hi += f[j]; q_hi = fmaf(f[j], f[j], q_hi); }
        }
    }
    red[threadIdx.x][0] = s_lo; red[threadIdx.x][1] = q_lo; red[threadIdx.x][2] = s_hi; red[threadIdx.x][3] = q_hi;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x, cpg = C / 32;
        float s = 0.f, q = 0.f;
        // half-octets (4 channels each) belonging to group g: [g*cpg/4, (g+1)*cpg/4)
        for (int hq = g * cpg / 4; hq < (g + 1) * cpg / 4; ++hq) {
            const int o = hq >> 1, half = hq & 1;
            for (int l = 0; l < PL; ++l) {
                s += red[l * C8 + o][half * 2];
                q += red[l * C8 + o][half * 2 + 1];
            }
        }
        float* out = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
        out[0] = s;
        out[1] = q;
    }
}

void launch_gn_stats(const bf16* x, float* partial, int B, int HW, int C, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gn_num_chunks(HW, C), B);
    gn_stats_kernel<<<grid, 256, 0, s>>>(x, partial, HW, C);
    COUNT_LAUNCH();
}

// per (image, group): mean / rstd, then per channel scale = rstd*gamma, shift = beta - mean*scale.
// One block per (group, image): the partial list (up to 4096 entries per image for 512^2 layers) is summed by 256
// threads in double precision and a fixed order (thread t takes entries t, t+256, ...; then a fixed tree).
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partial,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float2* __restrict__ ss,
                                                          float2* __restrict__ mr, int nchunks, int HW, int C,
                                                          float eps) {
    __shared__ double red[33];
    __shared__ float2 smr;
    const int g = blockIdx.x, b = blockIdx.y;
    double s = 0.0, q = 0.0;
    for (int c = threadIdx.x; c < nchunks; c += 256) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(partial + (((size_t)b * nchunks + c) * 32 + g) * 2));
        s += (double)v.x;
        q += (double)v.y;
    }
    s = block_sum_d(s, red);
    q = block_sum_d(q, red);
    const int cpg = C / 32;
    if (threadIdx.x == 0) {
        const double n = (double)HW * cpg;
        const double mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        smr = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
        mr[(size_t)b * 32 + g] = smr;
    }
    __syncthreads();
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        const float sc = smr.y * gamma[c];
        ss[(size_t)b * C + c] = make_float2(sc, beta[c] - smr.x * sc);
    }
}

void launch_gn_finalize(const float* partial, const float* gamma, const float* beta, float2* ss, float2* mr, int B,
                        int HW, int C, float eps, int nchunks, cudaStream_t s) {
    if (g_dry_run) return;
    gn_finalize_kernel<<<dim3(32, B), 256, 0, s>>>(partial, gamma, beta, ss, mr, nchunks, HW, C, eps);
    COUNT_LAUNCH();
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_f(float u) { return __fdividef(u, 1.f + __expf(-u)); }

__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, const float2* __restrict__ ss,
                                                       bf16* __restrict__ y, int HW, int C, int silu) {
    const int C8 = C >> 3, PL = 256 / C8;
    const int oct = threadIdx.x % C8, pl = threadIdx.x / C8;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * gn_pix_per_chunk(C), p1 = min(HW, p0 + gn_pix_per_chunk(C));
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 v = __ldg(&ss[(size_t)b * C + oct * 8 + j]);
        sc[j] = v.x; sh[j] = v.y;
    }
    const size_t base = (size_t)b * HW * C + (size_t)oct * 8;
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (int p = p0 + pl; p < p1; p += PL * U) {
        uint4 u[U];
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (p + k * PL < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + k * PL) * C));
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (p + k * PL >= p1) break;
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = fmaf(f[j], sc[j], sh[j]);
                f[j] = silu ? silu_f(v) : v;
            }
            *reinterpret_cast<uint4*>(y + base + (size_t)(p + k * PL) * C) = pack8(f);
        }
    }
}

void launch_gn_apply(const bf16* x, const float2* ss, bf16* y, int B, int HW, int C, int silu, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gn_num_chunks(HW, C), B);
    gn_apply_kernel<<<grid, 256, 0, s>>>(x, ss, y, HW, C, silu);
    COUNT_LAUNCH();
}

// d(act(u))/du for act = SiLU (u*sigmoid(u)) or identity
__device__ __forceinline__ float dact(float u, int silu) {
    if (!silu) return 1.f;
    const float sg = __fdividef(1.f, 1.f + __expf(-u));
    return sg * fmaf(u, 1.f - sg, 1.f);
}

// partial[b][chunk][g] = (sum dxh, sum dxh*xh), dxh = dy*act'(u)*gamma, xh = (x-mean)*rstd
__global__ void __launch_bounds__(256) gn_bwd_partial_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                             const float2* __restrict__ ss,
                                                             const float2* __restrict__ mr,
                                                             const float* __restrict__ gamma,
                                                             float* __restrict__ partial, int HW, int C, int silu) {
    __shared__ float red[256][4];
    const int C8 = C >> 3, PL = 256 / C8, cpg = C / 32;
    const int oct = threadIdx.x % C8, pl = threadIdx.x / C8;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int p0 = chunk * gn_pix_per_chunk(C), p1 = min(HW, p0 + gn_pix_per_chunk(C));
    float sc[8], sh[8], gm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 v = __ldg(&ss[(size_t)b * C + oct * 8 + j]);
        sc[j] = v.x; sh[j] = v.y;
        gm[j] = __ldg(&gamma[oct * 8 + j]);
    }
    const float2 m_lo = __ldg(&mr[(size_t)b * 32 + (oct * 8) / cpg]);
    const float2 m_hi = __ldg(&mr[(size_t)b * 32 + (oct * 8 + 4) / cpg]);
    float a_lo = 0.f, b_lo = 0.f, a_hi = 0.f, b_hi = 0.f;
    const size_t base = (size_t)b * HW * C + (size_t)oct * 8;
    constexpr int U = 4;  // pixels per iteration: 8 independent 16-byte loads in flight per thread
    for (int p = p0 + pl; p < p1; p += PL * U) {
        uint4 ux[U], ud[U];
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (p + k * PL < p1) {
                ux[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + k * PL) * C));
                ud[k] = __ldg(reinterpret_cast<const uint4*>(dy + base + (size_t)(p + k * PL) * C));
            }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (p + k * PL >= p1) break;
            float fx[8], fd[8];
            unpack8(ux[k], fx);
            unpack8(ud[k], fd);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float u = fmaf(fx[j], sc[j], sh[j]);
                const float dxh = fd[j] * dact(u, silu) * gm[j];
                const float2 m = j < 4 ? m_lo : m_hi;
                const float xh = (fx[j] - m.x) * m.y;
                if (j < 4) { a_lo += dxh; b_lo = fmaf(dxh, xh, b_lo); }
                else { a_hi += dxh; b_hi = fmaf(dxh, xh, b_hi); }
            }
        }
    }
    red[threadIdx.x][0] = a_lo; red[threadIdx.x][1] = b_lo; red[threadIdx.x][2] = a_hi; red[threadIdx.x][3] = b_hi;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x;
        float s = 0.f, q = 0.f;
        for (int hq = g * cpg / 4; hq < (g + 1) * cpg / 4; ++hq) {
            const int o = hq >> 1, half = hq & 1;
            for (int l = 0; l < PL; ++l) {
                s += red[l * C8 + o][half * 2];
                q += red[l * C8 + o][half * 2 + 1];
            }
        }
        float* out = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
        out[0] = s;
        out[1] = q;
    }
}

void launch_gn_bwd_partial(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float* gamma,
                           float* partial, int B, int HW, int C, int silu, cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gn_num_chunks(HW, C), B);
    gn_bwd_partial_kernel<<<grid, 256, 0, s>>>(x, dy, ss, mr, gamma, partial, HW, C, silu);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(const float* __restrict__ partial,
                                                              float2* __restrict__ mm, int nchunks, int HW, int C) {
    __shared__ double red[33];
    const int g = blockIdx.x, b = blockIdx.y;   // one block per (group, image), same scheme as gn_finalize_kernel
    double s = 0.0, q = 0.0;
    for (int c = threadIdx.x; c < nchunks; c += 256) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(partial + (((size_t)b * nchunks + c) * 32 + g) * 2));
        s += (double)v.x;
        q += (double)v.y;
    }
    s = block_sum_d(s, red);
    q = block_sum_d(q, red);
    if (threadIdx.x == 0) {
        const double n = (double)HW * (C / 32);
        mm[(size_t)b * 32 + g] = make_float2((float)(s / n), (float)(q / n));
    }
}

void launch_gn_bwd_finalize(const float* partial, float2* mm, int B, int HW, int C, int nchunks, cudaStream_t s) {
    if (g_dry_run) return;
    gn_bwd_finalize_kernel<<<dim3(32, B), 256, 0, s>>>(partial, mm, nchunks, HW, C);
    COUNT_LAUNCH();
}

// dx = rstd * (dxh - mean(dxh) - xh * mean(dxh*xh)) [+ resid],  dxh = dy * act'(u) * gamma,  u = x*sc + sh,
// xh = (x - mean) * rstd.  The kernel is instruction-bound before it is bandwidth-bound (3-4 streams of 2 B per
// element), so the affine parts are folded into per-thread constants:
//   dx = (rstd*gamma) * (dy * act'(u)) + (cx * x + c0),   cx = -rstd^2 * k.y,   c0 = rstd * (rstd*mean*k.y - k.x)
// and the SiLU derivative is sg + u * sg * (1 - sg) with exp2 taking a pre-scaled argument.
__global__ void __launch_bounds__(256, 3) gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                           const float2* __restrict__ ss,
                                                           const float2* __restrict__ mr,
                                                           const float2* __restrict__ mm,
                                                           const float* __restrict__ gamma,
                                                           const bf16* __restrict__ resid, bf16* __restrict__ dx,
                                                           int HW, int C, int silu) {
    const int C8 = C >> 3, PL = 256 / C8, cpg = C / 32;
    const int oct = threadIdx.x % C8, pl = threadIdx.x / C8;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * gn_pix_per_chunk(C), p1 = min(HW, p0 + gn_pix_per_chunk(C));
    const int g_lo = (oct * 8) / cpg, g_hi = (oct * 8 + 4) / cpg;
    const float2 m_lo = __ldg(&mr[(size_t)b * 32 + g_lo]), m_hi = __ldg(&mr[(size_t)b * 32 + g_hi]);
    const float2 k_lo = __ldg(&mm[(size_t)b * 32 + g_lo]), k_hi = __ldg(&mm[(size_t)b * 32 + g_hi]);
    float sc2[8], sh2[8], ag[8];
    constexpr float kNegLog2e = -1.4426950408889634f, kNegLn2 = -0.6931471805599453f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 v = __ldg(&ss[(size_t)b * C + oct * 8 + j]);
        sc2[j] = v.x * kNegLog2e; sh2[j] = v.y * kNegLog2e;   // t = -u*log2(e): exp(-u) = exp2(t), u = -ln2 * t
        ag[j] = (j < 4 ? m_lo.y : m_hi.y) * __ldg(&gamma[oct * 8 + j]);
    }
    const float cx_lo = -m_lo.y * m_lo.y * k_lo.y, c0_lo = m_lo.y * (m_lo.y * m_lo.x * k_lo.y - k_lo.x);
    const float cx_hi = -m_hi.y * m_hi.y * k_hi.y, c0_hi = m_hi.y * (m_hi.y * m_hi.x * k_hi.y - k_hi.x);
    const size_t base = (size_t)b * HW * C + (size_t)oct * 8;
    constexpr int U = 2;
    for (int p = p0 + pl; p < p1; p += PL * U) {
        uint4 ux[U], ud[U], ur[U];
#pragma unroll
        for (int q = 0; q < U; ++q)
            if (p + q * PL < p1) {
                ux[q] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(p + q * PL) * C));
                ud[q] = __ldg(reinterpret_cast<const uint4*>(dy + base + (size_t)(p + q * PL) * C));
                if (resid != nullptr) ur[q] = __ldg(reinterpret_cast<const uint4*>(resid + base + (size_t)(p + q * PL) * C));
            }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (p + q * PL >= p1) break;
            float fx[8], fd[8], fr[8], o[8];
            unpack8(ux[q], fx);
            unpack8(ud[q], fd);
            if (resid != nullptr) unpack8(ur[q], fr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d = fd[j];
                if (silu) {
                    const float t = fmaf(fx[j], sc2[j], sh2[j]);
                    const float sg = __fdividef(1.f, 1.f + ex2_approx(t));
                    d *= fmaf(t * kNegLn2, fmaf(-sg, sg, sg), sg);   // sg + u*sg*(1-sg)
                }
                float v = fmaf(ag[j], d, fmaf(j < 4 ? cx_lo : cx_hi, fx[j], j < 4 ? c0_lo : c0_hi));
                if (resid != nullptr) v += fr[j];
                o[j] = v;
            }
            *reinterpret_cast<uint4*>(dx + base + (size_t)(p + q * PL) * C) = pack8(o);
        }
    }
}

void launch_gn_bwd_apply(const bf16* x, const bf16* dy, const float2* ss, const float2* mr, const float2* mm,
                         const float* gamma, const bf16* resid, bf16* dx, int B, int HW, int C, int silu,
                         cudaStream_t s) {
    if (g_dry_run) return;
    dim3 grid(gn_num_chunks(HW, C), B);
    gn_bwd_apply_kernel<<<grid, 256, 0, s>>>(x, dy, ss, mr, mm, gamma, resid, dx, HW, C, silu);
    COUNT_LAUNCH();
}

// ================================================================================================
// Attention helpers.  (The softmax itself -- row max, exp, normalisation, backward -- runs in the epilogues of the
// QK^T / dP GEMMs, see GemmOp::epi_mode; what is left here are the transposes and the tiny row kernels.)
// ================================================================================================
// out[b][c][r] = in[b][r][c].  64x64 tiles; each thread moves bf16 pairs (4-byte accesses, 128-byte
// warp transactions on both sides); the +2 padding keeps the column reads bank-conflict free.
// row_scale (optional): input row r of batch b is multiplied by row_scale[b * R + r] on the way.
__global__ void __launch_bounds__(256) transpose_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R,
                                                        int C, long long ld_in, long long bs_in, long long ld_out,
                                                        long long bs_out, const float* __restrict__ row_scale) {
    __shared__ bf16 tile[64][66];
    const int b = blockIdx.z;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool full = (r0 + 64 <= R) && (c0 + 64 <= C);
    if (full) {
#pragma unroll
        for (int i = ty; i < 64; i += 8) {
            uint32_t v = *reinterpret_cast<const uint32_t*>(in + (size_t)b * bs_in + (size_t)(r0 + i) * ld_in + c0 + 2 * tx);
            if (row_scale) {
                const float sc = row_scale[(size_t)b * R + r0 + i];
                v = pack2(bf_lo(v) * sc, bf_hi(v) * sc);
            }
            *reinterpret_cast<uint32_t*>(&tile[i][2 * tx]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 64; i += 8) {
            __nv_bfloat162 v;
            v.x = tile[2 * tx][i];
            v.y = tile[2 * tx + 1][i];
            *reinterpret_cast<__nv_bfloat162*>(out + (size_t)b * bs_out + (size_t)(c0 + i) * ld_out + r0 + 2 * tx) = v;
        }
    } else {
        for (int i = ty; i < 64; i += 8)
            for (int j = tx; j < 64; j += 32) {
                const int r = r0 + i, c = c0 + j;
                if (r < R && c < C) {
                    bf16 v = in[(size_t)b * bs_in + (size_t)r * ld_in + c];
                    if (row_scale) v = __float2bfloat16_rn(__bfloat162float(v) * row_scale[(size_t)b * R + r]);
                    tile[i][j] = v;
                }
            }
        __syncthreads();
        for (int i = ty; i < 64; i += 8)
            for (int j = tx; j < 64; j += 32) {
                const int c = c0 + i, r = r0 + j;
                if (r < R && c < C) out[(size_t)b * bs_out + (size_t)c * ld_out + r] = tile[j][i];
            }
    }
}

void launch_transpose(const bf16* in, bf16* out, int batch, int R, int C, long long ld_in, long long bs_in,
                      long long ld_out, long long bs_out, cudaStream_t s, const float* row_scale) {
    if (g_dry_run) return;
    dim3 grid((C + 63) / 64, (R + 63) / 64, batch);
    transpose_kernel<<<grid, 256, 0, s>>>(in, out, R, C, ld_in, bs_in, ld_out, bs_out, row_scale);
    COUNT_LAUNCH();
}

// ================================================================================================
// Row helpers of the attention softmax, whose exp / normalisation / backward run in GEMM epilogues:
//   row_reduce: out[r] = max_k part[r][k]  (op 0)   or   1 / sum_k part[r][k]  (op 1), fixed order
//   row_dot:    out[r] = sum_c a[r][c] * b[r][c]    (D = rowsum(dO * O) of the softmax backward)
// ================================================================================================
__global__ void __launch_bounds__(256) row_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                         long long rows, int n, int op) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* pr = part + r * n;
    float acc = op == 0 ? -INFINITY : 0.f;
    for (int k = 0; k < n; ++k) acc = op == 0 ? fmaxf(acc, pr[k]) : acc + pr[k];
    out[r] = op == 0 ? acc : __fdiv_rn(1.f, acc);
}
void launch_row_reduce(const float* part, float* out, long long rows, int n, int op, cudaStream_t s) {
    if (g_dry_run) return;
    row_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(part, out, rows, n, op);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) row_dot_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                      float* __restrict__ out, long long rows, int C) {
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;   // one warp per row
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const uint4* pa = reinterpret_cast<const uint4*>(a + r * C);
    const uint4* pb = reinterpret_cast<const uint4*>(b + r * C);
    float acc = 0.f;
    for (int i = lane; i < C / 8; i += 32) {
        float fa[8], fb[8];
        unpack8(pa[i], fa);
        unpack8(pb[i], fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(fa[j], fb[j], acc);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) out[r] = acc;
}
void launch_row_dot(const bf16* a, const bf16* b, float* out, long long rows, int C, cudaStream_t s) {
    if (g_dry_run) return;
    row_dot_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(a, b, out, rows, C);
    COUNT_LAUNCH();
}

// ================================================================================================
// Posterior sample + latent loss + gradient w.r.t. the moments  
//   mean, logvar = chunk(moments); logvar = clamp(logvar,-30,20); z = mean + exp(.5 logvar)*eps
//   kind 0: loss_b = ||z_b - t_b||_2          (main.py:162, per image)
//   kind 1: loss_b = mean((z_b - t_b)^2)      (losses/losses.py:39-41)
// One block per image; the per-image sum uses a fixed-order block reduction in double.
// ================================================================================================
__global__ void __launch_bounds__(256) latent_loss_kernel(int kind, const float* __restrict__ moments,
                                                          const float* __restrict__ noise,
                                                          const float* __restrict__ target, int hw, float grad_scale,
                                                          float* __restrict__ z_out, float* __restrict__ loss,
                                                          float* __restrict__ dmoments) {
    __shared__ double red[33];
    const int b = blockIdx.x;
    const int n = 4 * hw;
    const float* mom = moments + (size_t)b * 8 * hw;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float mu = mom[i];
        const float lv = fminf(fmaxf(mom[n + i], -30.f), 20.f);
        const float e = noise ? noise[(size_t)b * n + i] : 0.f;
        const float z = fmaf(expf(0.5f * lv), e, mu);
        if (z_out) z_out[(size_t)b * n + i] = z;
        const float d = z - target[(size_t)b * n + i];
        acc += (double)d * (double)d;
    }
    const double tot = block_sum_d(acc, red);
    float coef, l;
    if (kind == 0) {
        l = (float)sqrt(tot);
        coef = l > 0.f ? grad_scale / l : 0.f;
    } else {
        l = (float)(tot / (double)n);
        coef = grad_scale * 2.f / (float)n;
    }
    if (threadIdx.x == 0 && loss) loss[b] = l;
    if (dmoments == nullptr) return;
    float* dm = dmoments + (size_t)b * 8 * hw;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float mu = mom[i];
        const float lv_raw = mom[n + i];
        const float lv = fminf(fmaxf(lv_raw, -30.f), 20.f);
        const float e = noise ? noise[(size_t)b * n + i] : 0.f;
        const float sd = expf(0.5f * lv);
        const float z = fmaf(sd, e, mu);
        const float dz = coef * (z - target[(size_t)b * n + i]);
        dm[i] = dz;
        dm[n + i] = (lv_raw >= -30.f && lv_raw <= 20.f) ? dz * e * sd * 0.5f : 0.f;
    }
}

void launch_latent_loss(int kind, const float* moments, const float* noise, const float* target, int B, int h, int w,
                        float grad_scale, float* z, float* loss, float* dmoments, cudaStream_t s) {
    if (g_dry_run) return;
    latent_loss_kernel<<<B, 256, 0, s>>>(kind, moments, noise, target, h * w, grad_scale, z, loss, dmoments);
    COUNT_LAUNCH();
}

__global__ void dmoments_pack_kernel(const float* __restrict__ dm, bf16* __restrict__ out, int hw, long long total) {
    // one thread per (pixel, octet) of the 64-channel padded NHWC tensor
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int oct = int(i & 7);
    const long long pix = i >> 3;  // b*hw + p
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (oct == 0) {
        const long long b = pix / hw, p = pix % hw;
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = dm[(b * 8 + c) * hw + p];
    }
    *reinterpret_cast<uint4*>(out + pix * 64 + oct * 8) = pack8(f);
}

void launch_dmoments_pack(const float* dm, bf16* out, int B, int h, int w, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * h * w * 8;
    dmoments_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dm, out, h * w, total);
    COUNT_LAUNCH();
}

// ================================================================================================
// Decoder-side helpers (vae.decode at main.py:156; image-space losses main.py:160,168)
// ================================================================================================
__global__ void latent_pack_kernel(const float* __restrict__ z, const float* __restrict__ wpq,
                                   const float* __restrict__ bpq, bf16* __restrict__ out, int hw, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (pixel, octet)
    if (i >= total) return;
    const int oct = int(i & 7);
    const long long pix = i >> 3;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (oct == 0) {
        const long long b = pix / hw, p = pix % hw;
        float zi[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) zi[c] = z[(b * 4 + c) * hw + p];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float a = bpq[o];
#pragma unroll
            for (int c = 0; c < 4; ++c) a = fmaf(wpq[o * 4 + c], zi[c], a);
            f[o] = a;
        }
    }
    *reinterpret_cast<uint4*>(out + pix * 64 + oct * 8) = pack8(f);
}
void launch_latent_pack(const float* z, const float* wpq, const float* bpq, bf16* out, int B, int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * hw * 8;
    latent_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, wpq, bpq, out, hw, total);
    COUNT_LAUNCH();
}

__global__ void latent_unpack_bwd_kernel(const bf16* __restrict__ d, const float* __restrict__ wpq,
                                         float* __restrict__ dz, int hw, long long total) {
    const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const uint2 u = *reinterpret_cast<const uint2*>(d + pix * 64);
    const float g[4] = {bf_lo(u.x), bf_hi(u.x), bf_lo(u.y), bf_hi(u.y)};
    const long long b = pix / hw, p = pix % hw;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float a = 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) a = fmaf(wpq[o * 4 + c], g[o], a);
        dz[(b * 4 + c) * hw + p] = a;
    }
}
void launch_latent_unpack_bwd(const bf16* d, const float* wpq, float* dz, int B, int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * hw;
    latent_unpack_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d, wpq, dz, hw, total);
    COUNT_LAUNCH();
}

// one thread per (low-res pixel, channel octet): read once, write the four copies
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int h,
                                                         int w, int C8, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int oct = int(i % C8);
        long long pix = i / C8;
        const int x = int(pix % w); pix /= w;
        const int y = int(pix % h);
        const long long b = pix / h;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
        const long long W2 = 2LL * w;
        uint4* o = reinterpret_cast<uint4*>(out) + ((b * 2 * h + 2 * y) * W2 + 2 * x) * C8 + oct;
        o[0] = v; o[C8] = v; o[W2 * C8] = v; o[W2 * C8 + C8] = v;
    }
}
void launch_upsample2x(const bf16* in, bf16* out, int B, int h, int w, int C, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * h * w * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    upsample2x_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, h, w, C / 8, total);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ din,
                                                             int h, int w, int C8, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int oct = int(i % C8);
        long long pix = i / C8;
        const int x = int(pix % w); pix /= w;
        const int y = int(pix % h);
        const long long b = pix / h;
        const long long W2 = 2LL * w;
        const uint4* o = reinterpret_cast<const uint4*>(dout) + ((b * 2 * h + 2 * y) * W2 + 2 * x) * C8 + oct;
        float a[8], t[8];
        unpack8(__ldg(o), a);
        unpack8(__ldg(o + C8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += t[j];
        unpack8(__ldg(o + W2 * C8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += t[j];
        unpack8(__ldg(o + W2 * C8 + C8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += t[j];
        reinterpret_cast<uint4*>(din)[i] = pack8(a);
    }
}
void launch_upsample2x_bwd(const bf16* dout, bf16* din, int B, int h, int w, int C, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * h * w * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    upsample2x_bwd_kernel<<<(unsigned)blocks, 256, 0, s>>>(dout, din, h, w, C / 8, total);
    COUNT_LAUNCH();
}

__global__ void image_pack_kernel(const float* __restrict__ dimg, bf16* __restrict__ out, long long hw, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (pixel, octet)
    if (i >= total) return;
    const int oct = int(i & 7);
    const long long pix = i >> 3;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (oct == 0) {
        const long long b = pix / hw, p = pix % hw;
#pragma unroll
        for (int c = 0; c < 3; ++c) f[c] = dimg[(b * 3 + c) * hw + p];
    }
    *reinterpret_cast<uint4*>(out + pix * 64 + oct * 8) = pack8(f);
}
void launch_image_pack(const float* dimg, bf16* out, int B, long long hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * hw * 8;
    image_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dimg, out, hw, total);
    COUNT_LAUNCH();
}

// image-space losses, per image (main.py:160 rec_loss = (output_image - target_image).norm(p=2);
// :168 pert_loss = F.mse_loss(output_image, source_image)), two-stage fixed-order reduction
constexpr int kImgChunks = 64;
size_t image_loss_workspace_bytes(int B) { return (size_t)B * kImgChunks * 2 * sizeof(double); }

__global__ void __launch_bounds__(256) image_loss_partial_kernel(const float* __restrict__ out,
                                                                 const float* __restrict__ target,
                                                                 const float* __restrict__ source,
                                                                 double* __restrict__ part, long long per_image) {
    __shared__ double red[33];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const long long c0 = per_image * chunk / kImgChunks, c1 = per_image * (chunk + 1) / kImgChunks;
    const size_t off = (size_t)b * per_image;
    double a = 0.0, q = 0.0;
    for (long long i = c0 + threadIdx.x; i < c1; i += 256) {
        const float o = out[off + i];
        const float dt = o - target[off + i];
        a += (double)dt * dt;
        if (source) { const float ds = o - source[off + i]; q += (double)ds * ds; }
    }
    a = block_sum_d(a, red);
    q = block_sum_d(q, red);
    if (threadIdx.x == 0) {
        part[((size_t)b * kImgChunks + chunk) * 2] = a;
        part[((size_t)b * kImgChunks + chunk) * 2 + 1] = q;
    }
}
__global__ void __launch_bounds__(256) image_loss_grad_kernel(const float* __restrict__ out,
                                                              const float* __restrict__ target,
                                                              const float* __restrict__ source,
                                                              const double* __restrict__ part, float rec_l,
                                                              float pert_l, float* __restrict__ rec,
                                                              float* __restrict__ pert, float* __restrict__ dout,
                                                              long long per_image) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    double a = 0.0, q = 0.0;
    for (int c = 0; c < kImgChunks; ++c) {
        a += part[((size_t)b * kImgChunks + c) * 2];
        q += part[((size_t)b * kImgChunks + c) * 2 + 1];
    }
    const float nrm = (float)sqrt(a);
    const float mse = (float)(q / (double)per_image);
    if (chunk == 0 && threadIdx.x == 0) {
        if (rec) rec[b] = nrm;
        if (pert) pert[b] = mse;
    }
    if (!dout) return;
    const float cr = nrm > 0.f ? rec_l / nrm : 0.f;
    const float cp = (source && pert_l != 0.f) ? pert_l * 2.f / (float)per_image : 0.f;
    const long long c0 = per_image * chunk / kImgChunks, c1 = per_image * (chunk + 1) / kImgChunks;
    const size_t off = (size_t)b * per_image;
    for (long long i = c0 + threadIdx.x; i < c1; i += 256) {
        const float o = out[off + i];
        float g = cr * (o - target[off + i]);
        if (cp != 0.f) g = fmaf(cp, o - source[off + i], g);
        dout[off + i] = g;
    }
}
void launch_image_loss(const float* out, const float* target, const float* source, int B, long long per_image,
                       float rec_l, float pert_l, float* rec, float* pert, float* dout, void* ws, cudaStream_t s) {
    if (g_dry_run) return;
    double* part = reinterpret_cast<double*>(ws);
    dim3 grid(kImgChunks, B);
    image_loss_partial_kernel<<<grid, 256, 0, s>>>(out, target, source, part, per_image);
    image_loss_grad_kernel<<<grid, 256, 0, s>>>(out, target, source, part, rec_l, pert_l, rec, pert, dout, per_image);
    COUNT_LAUNCH(); COUNT_LAUNCH();
}

__global__ void posterior_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                        float* __restrict__ z, int n4, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / n4, k = i % n4;
    const float mu = moments[b * 2 * n4 + k];
    const float lv = fminf(fmaxf(moments[b * 2 * n4 + n4 + k], -30.f), 20.f);
    z[i] = noise ? fmaf(expf(0.5f * lv), noise[i], mu) : mu;
}
void launch_posterior_sample(const float* moments, const float* noise, float* z, int B, int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * 4 * hw;
    posterior_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(moments, noise, z, 4 * hw, total);
    COUNT_LAUNCH();
}
__global__ void posterior_sample_bwd_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                            const float* __restrict__ dz, float* __restrict__ dm, int n4,
                                            long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / n4, k = i % n4;
    const float lv_raw = moments[b * 2 * n4 + n4 + k];
    const float lv = fminf(fmaxf(lv_raw, -30.f), 20.f);
    const float g = dz[i];
    dm[b * 2 * n4 + k] = g;
    dm[b * 2 * n4 + n4 + k] = (noise && lv_raw >= -30.f && lv_raw <= 20.f) ? g * noise[i] * expf(0.5f * lv) * 0.5f : 0.f;
}
void launch_posterior_sample_bwd(const float* moments, const float* noise, const float* dz, float* dmoments, int B,
                                 int hw, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * 4 * hw;
    posterior_sample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(moments, noise, dz, dmoments, 4 * hw, total);
    COUNT_LAUNCH();
}

// ================================================================================================
// PGD updates (main.py:248-276).  L-inf: one fused, 16-byte vectorised, grid-stride
// kernel — 16 B/element of HBM traffic (read X_adv, grad, X; write X_adv) instead of the 8 ATen
// launches of the reference.  Bit-exact with the ATen sequence including its special values:
// sign(+-0) = sign(NaN) = 0, minimum / maximum / clamp propagate NaN.
// ================================================================================================
__device__ __forceinline__ float sgn(float g) { return float(g > 0.f) - float(g < 0.f); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); }
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) { return (v != v) ? v : fminf(fmaxf(v, lo), hi); }

__device__ __forceinline__ float linf_one(float xa, float g, float x, float eps, float step, float lo, float hi) {
    float v = __fsub_rn(xa, __fmul_rn(sgn(g), step));       // X_adv - sign(grad)*step  (no fma contraction)
    v = min_nan(max_nan(v, __fsub_rn(x, eps)), __fadd_rn(x, eps));
    return clamp_nan(v, lo, hi);
}

__global__ void __launch_bounds__(256) pgd_linf_kernel(float* __restrict__ x_adv, const float* __restrict__ grad,
                                                       const float* __restrict__ x, float eps, float step, float lo,
                                                       float hi, long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<const float4*>(x_adv)[i];
        const float4 g = __ldcs(reinterpret_cast<const float4*>(grad) + i);
        const float4 s = __ldg(reinterpret_cast<const float4*>(x) + i);
        a.x = linf_one(a.x, g.x, s.x, eps, step, lo, hi);
        a.y = linf_one(a.y, g.y, s.y, eps, step, lo, hi);
        a.z = linf_one(a.z, g.z, s.z, eps, step, lo, hi);
        a.w = linf_one(a.w, g.w, s.w, eps, step, lo, hi);
        reinterpret_cast<float4*>(x_adv)[i] = a;
    }
    // tail (n not a multiple of 4)
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        x_adv[i] = linf_one(x_adv[i], grad[i], x[i], eps, step, lo, hi);
}

void launch_pgd_linf(float* x_adv, const float* grad, const float* x, float eps, float step, float lo, float hi,
                     long long n, cudaStream_t s) {
    if (g_dry_run) return;
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pgd_linf_kernel<<<(unsigned)blocks, 256, 0, s>>>(x_adv, grad, x, eps, step, lo, hi, n);
    COUNT_LAUNCH();
}

// ---- L2: per-image gradient normalisation, optional mask, step, renorm projection, clamp ----
constexpr int kL2Chunks = 64;
size_t pgd_l2_workspace_bytes(int B, long long) { return (size_t)B * kL2Chunks * sizeof(double) * 2; }

__global__ void __launch_bounds__(256) l2_sumsq_kernel(const float* __restrict__ g, double* __restrict__ part,
                                                       long long per_image) {
    __shared__ double red[33];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const long long c0 = per_image * chunk / kL2Chunks, c1 = per_image * (chunk + 1) / kL2Chunks;
    const float* p = g + (size_t)b * per_image;
    double acc = 0.0;
    for (long long i = c0 + threadIdx.x; i < c1; i += 256) acc += (double)p[i] * (double)p[i];
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0) part[(size_t)b * kL2Chunks + chunk] = acc;
}

__device__ __forceinline__ float l2_norm_from_parts(const double* part, int b) {
    double s = 0.0;
    for (int c = 0; c < kL2Chunks; ++c) s += part[(size_t)b * kL2Chunks + c];
    return (float)sqrt(s);
}

// x_adv <- x_adv - g/(||g||+1e-10) [*mask] * step ; part2 <- partial ||x_adv - x||^2
__global__ void __launch_bounds__(256) l2_step_kernel(float* __restrict__ x_adv, const float* __restrict__ g,
                                                      const float* __restrict__ x, const float* __restrict__ mask,
                                                      const double* __restrict__ part, double* __restrict__ part2,
                                                      float step, long long per_image, long long hw) {
    __shared__ double red[33];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float denom = l2_norm_from_parts(part, b) + 1e-10f;
    const long long c0 = per_image * chunk / kL2Chunks, c1 = per_image * (chunk + 1) / kL2Chunks;
    const size_t off = (size_t)b * per_image;
    double acc = 0.0;
    for (long long i = c0 + threadIdx.x; i < c1; i += 256) {
        float gn = __fdiv_rn(g[off + i], denom);
        if (mask) gn = __fmul_rn(gn, mask[(size_t)b * hw + (i % hw)]);
        const float v = __fsub_rn(x_adv[off + i], __fmul_rn(gn, step));
        x_adv[off + i] = v;
        const float d = __fsub_rn(v, x[off + i]);
        acc += (double)d * (double)d;
    }
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0) part2[(size_t)b * kL2Chunks + chunk] = acc;
}

// torch.renorm(d, 2, 0, eps): rows with norm > eps are scaled by eps/(norm+1e-7); then clamp(X + d)
__global__ void __launch_bounds__(256) l2_project_kernel(float* __restrict__ x_adv, const float* __restrict__ x,
                                                         const double* __restrict__ part2, float eps, float lo,
                                                         float hi, long long per_image) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float nrm = l2_norm_from_parts(part2, b);
    const bool scale = nrm > eps;
    const float factor = scale ? __fdiv_rn(eps, __fadd_rn(nrm, 1e-7f)) : 1.f;
    const long long c0 = per_image * chunk / kL2Chunks, c1 = per_image * (chunk + 1) / kL2Chunks;
    const size_t off = (size_t)b * per_image;
    for (long long i = c0 + threadIdx.x; i < c1; i += 256) {
        float d = __fsub_rn(x_adv[off + i], x[off + i]);
        if (scale) d = __fmul_rn(d, factor);
        x_adv[off + i] = clamp_nan(__fadd_rn(x[off + i], d), lo, hi);
    }
}

void launch_pgd_l2(float* x_adv, const float* grad, const float* x, const float* mask, float eps, float step, float lo,
                   float hi, int B, int C, long long hw, void* ws, cudaStream_t s) {
    if (g_dry_run) return;
    const long long per_image = (long long)C * hw;
    double* part = reinterpret_cast<double*>(ws);
    double* part2 = part + (size_t)B * kL2Chunks;
    dim3 grid(kL2Chunks, B);
    l2_sumsq_kernel<<<grid, 256, 0, s>>>(grad, part, per_image);
    l2_step_kernel<<<grid, 256, 0, s>>>(x_adv, grad, x, mask, part, part2, step, per_image, hw);
    l2_project_kernel<<<grid, 256, 0, s>>>(x_adv, x, part2, eps, lo, hi, per_image);
    COUNT_LAUNCH(); COUNT_LAUNCH(); COUNT_LAUNCH();
}

// ================================================================================================
// Universal perturbation helpers (old/train_noise.py:127-185)
// ================================================================================================
__global__ void add_delta_kernel(const float* __restrict__ x, const float* __restrict__ delta, float* __restrict__ out,
                                 long long per_image, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride)
        out[i] = __fadd_rn(x[i], delta[i % per_image]);   // :132  source_image + perturbation
}
void launch_add_delta(const float* x, const float* delta, float* out, int B, long long per_image, cudaStream_t s) {
    if (g_dry_run) return;
    const long long total = (long long)B * per_image;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    add_delta_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, delta, out, per_image, total);
    COUNT_LAUNCH();
}

// out[i] = scale * sum_b g[b][i], summed in image order (fixed order -> identical on every rank)
__global__ void batch_sum_kernel(const float* __restrict__ g, float* __restrict__ out, int B, long long per_image,
                                 float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_image; i += stride) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += g[(size_t)b * per_image + i];
        out[i] = acc * scale;
    }
}
void launch_batch_sum(const float* g, float* out, int B, long long per_image, float scale, cudaStream_t s) {
    if (g_dry_run) return;
    long long blocks = (per_image + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    batch_sum_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, out, B, per_image, scale);
    COUNT_LAUNCH();
}

// delta <- clamp(delta - g/(||g||+1e-10)*step, -eps, eps); optional image-range re-projection
__global__ void __launch_bounds__(256) universal_step_kernel(float* __restrict__ delta, const float* __restrict__ g,
                                                             const float* __restrict__ source,
                                                             const double* __restrict__ part, float eps, float step,
                                                             float lo, float hi, long long n) {
    const float denom = l2_norm_from_parts(part, 0) + 1e-10f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float d = __fsub_rn(delta[i], __fmul_rn(__fdiv_rn(g[i], denom), step));   // :173-177
        d = clamp_nan(d, -eps, eps);                                              // :180
        if (source) d = __fsub_rn(clamp_nan(__fadd_rn(source[i], d), lo, hi), source[i]);  // :183-185
        delta[i] = d;
    }
}
void launch_universal_step(float* delta, const float* grad, const float* source, float eps, float step, float lo,
                           float hi, long long n, void* ws, cudaStream_t s) {
    if (g_dry_run) return;
    double* part = reinterpret_cast<double*>(ws);
    dim3 grid(kL2Chunks, 1);
    l2_sumsq_kernel<<<grid, 256, 0, s>>>(grad, part, n);
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    universal_step_kernel<<<(unsigned)blocks, 256, 0, s>>>(delta, grad, source, part, eps, step, lo, hi, n);
    COUNT_LAUNCH(); COUNT_LAUNCH();
}


// delta <- for each source image s in order: clamp(source_s + delta, lo, hi) - source_s      (old/train_noise.py:183-185,
// one image per reference step; several sources = the same statement applied once per image of the step's batch)
__global__ void __launch_bounds__(256) universal_project_kernel(float* __restrict__ delta, const float* __restrict__ sources,
                                                                int nsrc, float lo, float hi, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float d = delta[i];
        for (int s = 0; s < nsrc; ++s) {
            const float x = sources[(size_t)s * n + i];
            d = __fsub_rn(clamp_nan(__fadd_rn(x, d), lo, hi), x);
        }
        delta[i] = d;
    }
}
void launch_universal_project(float* delta, const float* sources, int nsrc, float lo, float hi, long long n, cudaStream_t s) {
    if (g_dry_run) return;
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    universal_project_kernel<<<(unsigned)blocks, 256, 0, s>>>(delta, sources, nsrc, lo, hi, n);
    COUNT_LAUNCH();
}

}  // namespace tml
