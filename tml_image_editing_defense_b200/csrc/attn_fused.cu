// Fused multi-head attention for the UNet's self-attention layers (diffusers Attention / SDPA call reached from
// main.py:233-238, and its backward through torch.autograd.grad at main.py:176): O = softmax(Q K^T * scale) V for every
// (sample, head) without the token x token matrix ever leaving the SM.
//
// Forward, mh_attn_fwd_kernel: CTA = one (sample, head) x 128 query rows, looping over key tiles of 64:
//     warp 0   TMA producer: Q once, then K and V tiles into a two-stage ring (128-byte swizzle)
//     warp 1   tcgen05.mma issuer: S = Q K^T into 64 TMEM columns, and, one tile behind, O += P~ V with P~ read from
//              shared memory and the V tile [key][channel] read as an MN-major operand (no transposed copy)
//     warps 2-9  softmax, thread = (query row, 32-key half): tcgen05.ld of S into registers (S is released at once, so one
//              TMEM buffer suffices), P~ = exp2((s - rowmax) * scale * log2 e) with the row maxima of the preceding max
//              pass (no accumulator rescaling is ever needed), bf16 P~ into the swizzled K-major operand tile of the PV MMA
//   TMEM: S (64 columns) | O (DP columns) = one 256-column allocation; with 80 KB of shared memory two CTAs share an SM
//   and fill each other's pipeline bubbles (measured: 1 CTA/SM with 128-key tiles 1.20 ms per layer, this form 1.01 ms).
//   The softmax denominator is a column of O: V carries 1.0 in its first padding channel.
//   Epilogue: O / l -> bf16, 1 / l -> fp32 (kept for the backward).
// Backward: attn_bwd_dq_kernel / attn_bwd_dkv_kernel, see the comment above them.
//
// The unfused GEMM-epilogue path of unet.cu stays the verified baseline (TML_NO_FUSED_ATTN=1) and still serves head widths
// above 128 and the cross attention.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gemm.h"
#include "kernels.h"
#include "ptx.cuh"

namespace tml {

void count_launch();
int encode_map_bf16_sw128(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                          const cuuint32_t* box, const char* what);   // gemm_tc.cu

namespace {

constexpr int kFaThreads = 320;   // TMA warp, MMA warp, 8 softmax warps (two per SM sub-partition: column halves of a row)
constexpr int kSmThreads = 256;
constexpr uint32_t kUmmaBMajorMN = 1u << 16;

struct FaParams {
    const float* rmax;    // [nb][tq]  known row maxima (mh_attn_fwd_kernel)
    float* rmax_out;      // [nb][tq]  reference maxima chosen by the online kernel (kept for the backward)
    float* inv_l;         // [nb][tq]
    __nv_bfloat16* O;     // [nb][tq][DP]
    int tq, tkv;
    int lcol;             // channel of V that holds 1.0 for every key: O[:, lcol] accumulates the softmax denominator
    float exp_scale;      // scale * log2(e)
};

__device__ __forceinline__ float fa_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t fa_pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// explicit shared-space accesses (the tile pointers are derived from a rounded-up integer address, so the compiler would
// otherwise emit generic ST.E / LD.E for them)
__device__ __forceinline__ void fa_sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 fa_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float fa_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float fa_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

template <int DP>
__global__ void __launch_bounds__(kFaThreads, 2) mh_attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                   const __grid_constant__ CUtensorMap mapK,
                                                                   const __grid_constant__ CUtensorMap mapV,
                                                                   const FaParams p) {
    constexpr int NC = DP / 64;                 // 64-channel chunks of the head width
    constexpr int kBig = 128 * 128, kSmall = 64 * 128;   // [128 rows][128 B], [64 rows][128 B]
    constexpr int kQBytes = NC * kBig, kKBytes = NC * kSmall;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;                 // 2 stages of [64 keys][DP]
    uint8_t* sV = sK + 2 * kKBytes;             // 2 stages
    uint8_t* sP = sV + 2 * kKBytes;             // 2 buffers of [128 rows][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kBig);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;               // [2]
    uint64_t* kv_empty = bars + 3;              // [2]
    uint64_t* s_full = bars + 5;                // [1]
    uint64_t* s_empty = bars + 7;               // [1]
    uint64_t* p_full = bars + 9;                // [2]
    uint64_t* p_empty = bars + 11;              // [2]
    uint64_t* o_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, batch = blockIdx.y;
    const int ntiles = p.tkv / 64;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], kSmThreads);
            mbar_init(&p_full[i], kSmThreads);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    // 256 TMEM columns (S: 64, freed by the softmax warps as soon as it is in registers; O: DP) and 80 KB of shared
    // memory at DP = 64: two CTAs share an SM and fill each other's pipeline bubbles (prologue, barrier round trips)
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kOCol = 64;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, kQBytes);
            for (int c = 0; c < NC; ++c) tma_load_3d(sQ + c * kBig, &mapQ, q_full, c * 64, q0, batch);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], (((j >> 1) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[s], 2 * kKBytes);
                for (int c = 0; c < NC; ++c) {
                    tma_load_3d(sK + s * kKBytes + c * kSmall, &mapK, &kv_full[s], c * 64, j * 64, batch);
                    tma_load_3d(sV + s * kKBytes + c * kSmall, &mapV, &kv_full[s], c * 64, j * 64, batch);
                }
            }
        }
    } else if (warp == 1) {
        // S = Q K^T: M = 128 queries, N = 64 keys, K = DP (both operands K-major)
        const uint32_t idesc_s = umma_idesc_bf16(128, 64);
        // O += P V: M = 128 queries, N = DP channels, K = 64 keys; V tile [key][channel] is the MN-major B operand
        const uint32_t idesc_o = umma_idesc_bf16(128, DP) | kUmmaBMajorMN;
        mbar_wait(q_full, 0);
        tc_fence_after();
        auto issue_pv = [&](int j) {
            const int s = j & 1, b = j & 1;
            mbar_wait(&p_full[b], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t pa = smem_u32(sP + b * kBig), va = smem_u32(sV + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)   // 4 x 16 keys
                    // MN-major: channel atoms (64) kSmall apart, groups of 8 keys 1 KB apart; 16 keys = 2 KB per step
                    umma_bf16(tmem_base + kOCol, umma_desc_sw128(pa) + uint64_t(kk * 2),
                              umma_desc_sw128_mn(va + kk * 2048, kSmall, 1024), idesc_o, (j | kk) != 0 ? 1u : 0u);
                umma_commit(&kv_empty[s]);
                umma_commit(&p_empty[b]);
            }
            __syncwarp();
        };
        for (int j = 0; j < ntiles; ++j) {
            const int s = j & 1;
            mbar_wait(&kv_full[s], (j >> 1) & 1);
            mbar_wait(&s_empty[0], ((j & 1) ^ 1));
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base, umma_desc_sw128(qa + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(ka + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                umma_commit(&s_full[0]);
            }
            __syncwarp();
            if (j > 0) issue_pv(j - 1);
        }
        issue_pv(ntiles - 1);
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
    } else {
        // ===================================================================== softmax warps
        // thread = (query row, 32-key half of the tile): warps w and w + 4 share a TMEM lane quadrant
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const long long grow = (long long)batch * p.tq + q0 + row;
        const float c = p.exp_scale;
        const float ra = -__ldg(p.rmax + grow) * c;
        const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16);
        for (int j = 0; j < ntiles; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[0], j & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(t_row + uint32_t(half * 32), v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_empty[0]);   // S is in registers: the next tile's logits may overwrite it
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fa_ex2(fmaf(__uint_as_float(v[i]), c, ra));
            mbar_wait(&p_empty[b], (((j >> 1) & 1) ^ 1));
            uint8_t* dst = sP + b * kBig + row * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 o = make_uint4(fa_pack(f[8 * u], f[8 * u + 1]), fa_pack(f[8 * u + 2], f[8 * u + 3]),
                                           fa_pack(f[8 * u + 4], f[8 * u + 5]), fa_pack(f[8 * u + 6], f[8 * u + 7]));
                const int unit = half * 4 + u;
                fa_sts128(smem_u32(dst) + ((unit ^ (row & 7)) << 4), o);
            }
            fence_proxy_async();
            mbar_arrive(&p_full[b]);
        }
        // The softmax denominator is a column of O: channel `lcol` of V is 1.0 for every key (a padding channel of the
        // head), so the tensor cores sum the probabilities exactly as they are stored (bf16), in fp32, for free.
        mbar_wait(o_full, 0);
        tc_fence_after();
        float l;
        {
            uint32_t v8[8];
            tmem_ld8(t_row + kOCol + uint32_t(p.lcol & ~7), v8);
            tmem_ld_wait();
            l = __uint_as_float(v8[p.lcol & 7]);
        }
        const float il = 1.f / l;
        if (half == 0) p.inv_l[grow] = il;
        __nv_bfloat16* orow = p.O + grow * DP;
#pragma unroll 1
        for (int ch = half * (DP / 64); ch < (half + 1) * (DP / 64); ++ch) {
            uint32_t v[32];
            tmem_ld32(t_row + kOCol + uint32_t(ch * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 o = make_uint4(fa_pack(__uint_as_float(v[8 * u]) * il, __uint_as_float(v[8 * u + 1]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 2]) * il, __uint_as_float(v[8 * u + 3]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 4]) * il, __uint_as_float(v[8 * u + 5]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 6]) * il, __uint_as_float(v[8 * u + 7]) * il));
                *reinterpret_cast<uint4*>(orow + ch * 32 + u * 8) = o;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// The same forward with an ONLINE softmax: no preceding row-max pass over Q K^T.  Thread = query row (four softmax warps,
// 64 keys per tile).  The exponent reference m_ref of a row only moves when a tile's maximum exceeds it by more than 8 in
// log2 units, so stored probabilities stay below 2^8 (bf16 keeps its relative precision there) and the accumulator --
// including its denominator column -- is rescaled in TMEM (tcgen05.ld / st) only on those rare tiles, after the previous
// tile's P V product has retired.  The final m_ref is written out: the backward recomputes P~ against it.
template <int DP>
__global__ void __launch_bounds__(192, 2) mh_attn_fwd_online_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                    const __grid_constant__ CUtensorMap mapK,
                                                                    const __grid_constant__ CUtensorMap mapV,
                                                                    const FaParams p) {
    constexpr int NC = DP / 64;
    constexpr int KS = DP == 64 ? 3 : 2;        // K / V ring depth: the TMA round trip is longer than one 64-key tile
    constexpr int kBig = 128 * 128, kSmall = 64 * 128;
    constexpr int kQBytes = NC * kBig, kKBytes = NC * kSmall;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;
    uint8_t* sV = sK + KS * kKBytes;
    uint8_t* sP = sV + KS * kKBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kBig);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 16;              // [KS]
    uint64_t* kv_empty = bars + 20;             // [KS]
    uint64_t* s_full = bars + 5;
    uint64_t* s_empty = bars + 7;
    uint64_t* p_full = bars + 9;
    uint64_t* p_empty = bars + 11;
    uint64_t* o_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, batch = blockIdx.y;
    const int ntiles = p.tkv / 64;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < KS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kOCol = 64;
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, kQBytes);
            for (int c = 0; c < NC; ++c) tma_load_3d(sQ + c * kBig, &mapQ, q_full, c * 64, q0, batch);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % KS;
                mbar_wait(&kv_empty[s], (((j / KS) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[s], 2 * kKBytes);
                for (int c = 0; c < NC; ++c) {
                    tma_load_3d(sK + s * kKBytes + c * kSmall, &mapK, &kv_full[s], c * 64, j * 64, batch);
                    tma_load_3d(sV + s * kKBytes + c * kSmall, &mapV, &kv_full[s], c * 64, j * 64, batch);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(128, 64);
        const uint32_t idesc_o = umma_idesc_bf16(128, DP) | kUmmaBMajorMN;
        mbar_wait(q_full, 0);
        tc_fence_after();
        auto issue_pv = [&](int j) {
            const int s = j % KS, b = j & 1;
            mbar_wait(&p_full[b], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t pa = smem_u32(sP + b * kBig), va = smem_u32(sV + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base + kOCol, umma_desc_sw128(pa) + uint64_t(kk * 2),
                              umma_desc_sw128_mn(va + kk * 2048, kSmall, 1024), idesc_o, (j | kk) != 0 ? 1u : 0u);
                umma_commit(&kv_empty[s]);
                umma_commit(&p_empty[b]);
            }
            __syncwarp();
        };
        for (int j = 0; j < ntiles; ++j) {
            const int s = j % KS;
            mbar_wait(&kv_full[s], (j / KS) & 1);
            mbar_wait(&s_empty[0], ((j & 1) ^ 1));
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base, umma_desc_sw128(qa + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(ka + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                umma_commit(&s_full[0]);
            }
            __syncwarp();
            if (j > 0) issue_pv(j - 1);
        }
        issue_pv(ntiles - 1);
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const long long grow = (long long)batch * p.tq + q0 + row;
        const float c = p.exp_scale;
        const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16);
        float m_ref = 0.f;
        for (int j = 0; j < ntiles; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[0], j & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld32(t_row, v0);
            tmem_ld32(t_row + 32u, v1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_empty[0]);
            float mt = __uint_as_float(v0[0]);
#pragma unroll
            for (int i = 1; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(v0[i]));
#pragma unroll
            for (int i = 0; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(v1[i]));
            if (j == 0) {
                m_ref = mt;
            } else {
                const bool need = (mt - m_ref) * c > 8.f;
                if (__any_sync(0xffffffffu, need)) {
                    // rescale this warp's 32 accumulator rows once the previous tile's P V has retired
                    mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
                    tc_fence_after();
                    const float alpha = need ? fa_ex2((m_ref - mt) * c) : 1.f;
                    if (need) m_ref = mt;
#pragma unroll 1
                    for (int ch = 0; ch < DP / 32; ++ch) {
                        uint32_t o[32];
                        tmem_ld32(t_row + kOCol + uint32_t(ch * 32), o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(t_row + kOCol + uint32_t(ch * 32), o);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                }
            }
            const float ra = -m_ref * c;
            mbar_wait(&p_empty[b], (((j >> 1) & 1) ^ 1));
            uint8_t* dst = sP + b * kBig + row * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t* vv = u < 4 ? v0 : v1;
                const int k = (u & 3) * 8;
                const uint4 o = make_uint4(
                    fa_pack(fa_ex2(fmaf(__uint_as_float(vv[k]), c, ra)), fa_ex2(fmaf(__uint_as_float(vv[k + 1]), c, ra))),
                    fa_pack(fa_ex2(fmaf(__uint_as_float(vv[k + 2]), c, ra)), fa_ex2(fmaf(__uint_as_float(vv[k + 3]), c, ra))),
                    fa_pack(fa_ex2(fmaf(__uint_as_float(vv[k + 4]), c, ra)), fa_ex2(fmaf(__uint_as_float(vv[k + 5]), c, ra))),
                    fa_pack(fa_ex2(fmaf(__uint_as_float(vv[k + 6]), c, ra)), fa_ex2(fmaf(__uint_as_float(vv[k + 7]), c, ra))));
                fa_sts128(smem_u32(dst) + ((u ^ (row & 7)) << 4), o);
            }
            fence_proxy_async();
            mbar_arrive(&p_full[b]);
        }
        mbar_wait(o_full, 0);
        tc_fence_after();
        float l;
        {
            uint32_t v8[8];
            tmem_ld8(t_row + kOCol + uint32_t(p.lcol & ~7), v8);
            tmem_ld_wait();
            l = __uint_as_float(v8[p.lcol & 7]);
        }
        const float il = 1.f / l;
        p.inv_l[grow] = il;
        p.rmax_out[grow] = m_ref;
        __nv_bfloat16* orow = p.O + grow * DP;
#pragma unroll 1
        for (int ch = 0; ch < DP / 32; ++ch) {
            uint32_t v[32];
            tmem_ld32(t_row + kOCol + uint32_t(ch * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 o = make_uint4(fa_pack(__uint_as_float(v[8 * u]) * il, __uint_as_float(v[8 * u + 1]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 2]) * il, __uint_as_float(v[8 * u + 3]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 4]) * il, __uint_as_float(v[8 * u + 5]) * il),
                                           fa_pack(__uint_as_float(v[8 * u + 6]) * il, __uint_as_float(v[8 * u + 7]) * il));
                *reinterpret_cast<uint4*>(orow + ch * 32 + u * 8) = o;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

template <int DP>
int launch_fa(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, const float* rmax, float* rmax_out,
              float* inv_l, __nv_bfloat16* O, int nb, int tq, int tkv, int lcol, float scale, cudaStream_t st) {
    CUtensorMap mq, mk, mv;
    int rc;
    cuuint32_t box[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    {
        cuuint64_t dims[3] = {(cuuint64_t)DP, (cuuint64_t)tq, (cuuint64_t)nb};
        cuuint64_t str[2] = {(cuuint64_t)DP * 2, (cuuint64_t)tq * DP * 2};
        if ((rc = encode_map_bf16_sw128(&mq, Q, 3, dims, str, box, "attn.fused.Q"))) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)DP, (cuuint64_t)tkv, (cuuint64_t)nb};
        cuuint64_t str[2] = {(cuuint64_t)DP * 2, (cuuint64_t)tkv * DP * 2};
        if ((rc = encode_map_bf16_sw128(&mk, K, 3, dims, str, box64, "attn.fused.K"))) return rc;
        if ((rc = encode_map_bf16_sw128(&mv, V, 3, dims, str, box64, "attn.fused.V"))) return rc;
    }
    constexpr int kBig = 128 * 128, kSmall = 64 * 128;
    constexpr int KS = DP == 64 ? 3 : 2;   // (both forward kernels are given the online kernel's deeper K / V ring)
    constexpr size_t smem = size_t(DP / 64) * (kBig + 2 * KS * kSmall) + 2 * kBig + 256 + 1024;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(mh_attn_fwd_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mh_attn_fwd_online_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("attn.fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -4; }
        attr_set[dev & 63] = true;
    }
    FaParams p;
    p.rmax = rmax; p.rmax_out = rmax_out; p.inv_l = inv_l; p.O = O; p.tq = tq; p.tkv = tkv; p.lcol = lcol;
    p.exp_scale = scale * 1.4426950408889634f;
    if (rmax == nullptr) mh_attn_fwd_online_kernel<DP><<<dim3(tq / 128, nb), 192, smem, st>>>(mq, mk, mv, p);
    else mh_attn_fwd_kernel<DP><<<dim3(tq / 128, nb), kFaThreads, smem, st>>>(mq, mk, mv, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("attn.fused: launch failed: %s", cudaGetErrorString(e)); return -5; }
    count_launch();
    return 0;
}


// ================================================================================================
// Fused backward.  Inputs: Q, K, V [nb][tok][DP]; dO' = dO / l (rows pre-scaled, bf16) and D' = (dO . O) / l from
// attn_bwd_prep_kernel; the saved row maxima.  With P~ = exp2((s - m) * scale * log2 e) rounded to bf16 as in the forward:
//     dP' = dO' V^T,   dS = P~ * (dP' - D') * scale,   dQ = dS K,   dK = dS^T Q,   dV = P~^T dO'
// Two kernels, each recomputing S and dP' on the tensor cores so that nothing token x token is ever written:
//   attn_bwd_dq_kernel   CTA = 128 query rows, loop over 64-key tiles:   S | dP' -> dS (smem) -> dQ += dS K
//   attn_bwd_dkv_kernel  CTA = 128 key rows, loop over 64-query tiles:   S^T | dP'^T -> P~^T, dS^T (smem)
//                                                                         -> dV += P~^T dO',  dK += dS^T Q
// The K (resp. Q, dO') tile that serves as the K-major B operand of the logits is read a second time as the MN-major
// B operand of the accumulating product: same bytes in shared memory, two descriptors.
// ================================================================================================
struct FbParams {
    const float* rmax;    // [nb][tq]
    const float* Dp;      // [nb][tq]  D' = (dO . O) / l
    const float2* cq;     // [nb][tq]  (-rmax * scale * log2 e, -D' * scale): the dK/dV kernel's per-query constants
    __nv_bfloat16* dQ;    // [nb][tq][DP]            (dq kernel)
    __nv_bfloat16* dK;    // [nb][tkv][DP]           (dkv kernel)
    __nv_bfloat16* dV;
    int tq, tkv;
    float exp_scale, scale;
};

// eight lanes per row (DP / 8 = 8 or 16 octets): D'[row] = il * sum_c dO[row][c] * O[row][c];  dOs[row][c] = dO[row][c] * il;
// cq[row] = (-rmax * scale * log2 e, -D' * scale)
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dO,
                                                            const __nv_bfloat16* __restrict__ O,
                                                            const float* __restrict__ inv_l, const float* __restrict__ rmax,
                                                            __nv_bfloat16* __restrict__ dOs, float* __restrict__ Dp,
                                                            float2* __restrict__ cq, float exp_scale, float scale,
                                                            long long rows, int DP) {
    const int sub = threadIdx.x & 7;
    const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 3;
    const bool live = row < rows;
    const float il = live ? __ldg(inv_l + row) : 0.f;
    float acc = 0.f;
    if (live)
        for (int o = sub; o < (DP >> 3); o += 8) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(dO + row * DP) + o);
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(O + row * DP) + o);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc = fmaf(fa_lo(aw[k]), fa_lo(bw[k]), acc);
                acc = fmaf(fa_hi(aw[k]), fa_hi(bw[k]), acc);
                ow[k] = fa_pack(fa_lo(aw[k]) * il, fa_hi(aw[k]) * il);
            }
            reinterpret_cast<uint4*>(dOs + row * DP)[o] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && sub == 0) {
        Dp[row] = acc * il;
        cq[row] = make_float2(-__ldg(rmax + row) * exp_scale, -acc * il * scale);
    }
}

template <int DP>
__global__ void __launch_bounds__(kFaThreads, 2) attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                   const __grid_constant__ CUtensorMap mapDO,
                                                                   const __grid_constant__ CUtensorMap mapK,
                                                                   const __grid_constant__ CUtensorMap mapV,
                                                                   const FbParams p) {
    constexpr int NC = DP / 64;
    constexpr int KS = DP == 64 ? 3 : 2;        // K / V ring depth: the TMA round trip is longer than one 64-key tile
    constexpr int kBig = 128 * 128, kSmall = 64 * 128;    // [128 rows][128 B], [64 rows][128 B]
    constexpr int kQBytes = NC * kBig, kKBytes = NC * kSmall;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sDO = sQ + kQBytes;
    uint8_t* sK = sDO + kQBytes;                // 2 stages
    uint8_t* sV = sK + KS * kKBytes;             // 2 stages
    uint8_t* sDS = sV + KS * kKBytes;           // 2 buffers of [128 rows][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * kBig);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 16;              // [KS]
    uint64_t* kv_empty = bars + 20;             // [KS]
    uint64_t* sd_full = bars + 5;
    uint64_t* sd_empty = bars + 7;
    uint64_t* ds_full = bars + 9;
    uint64_t* ds_empty = bars + 11;
    uint64_t* acc_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, batch = blockIdx.y;
    const int ntiles = p.tkv / 64;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < KS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], kSmThreads);
            mbar_init(&ds_full[i], kSmThreads); mbar_init(&ds_empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    // 256 TMEM columns (S | dP' single-buffered: the softmax warps free them as soon as they are in registers, + dQ) and
    // ~100 KB of shared memory: two CTAs share an SM and fill each other's pipeline bubbles
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    if (DP == 64 && smem != smem_raw) __trap();   // the launch requests no alignment slack at DP = 64 (112 KB + barriers)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kAccCol = 128;
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 2 * kQBytes);
            for (int c = 0; c < NC; ++c) {
                tma_load_3d(sQ + c * kBig, &mapQ, q_full, c * 64, q0, batch);
                tma_load_3d(sDO + c * kBig, &mapDO, q_full, c * 64, q0, batch);
            }
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % KS;
                mbar_wait(&kv_empty[s], (((j / KS) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[s], 2 * kKBytes);
                for (int c = 0; c < NC; ++c) {
                    tma_load_3d(sK + s * kKBytes + c * kSmall, &mapK, &kv_full[s], c * 64, j * 64, batch);
                    tma_load_3d(sV + s * kKBytes + c * kSmall, &mapV, &kv_full[s], c * 64, j * 64, batch);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(128, 64);
        const uint32_t idesc_a = umma_idesc_bf16(128, DP) | kUmmaBMajorMN;
        mbar_wait(q_full, 0);
        tc_fence_after();
        auto issue_acc = [&](int j) {
            const int s = j % KS, b = j & 1;
            mbar_wait(&ds_full[b], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t da = smem_u32(sDS + b * kBig), ka = smem_u32(sK + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)   // dQ += dS K: 4 x 16 keys; K tile [key][channel] read MN-major
                    umma_bf16(tmem_base + kAccCol, umma_desc_sw128(da) + uint64_t(kk * 2),
                              umma_desc_sw128_mn(ka + kk * 2048, kSmall, 1024), idesc_a, (j | kk) != 0 ? 1u : 0u);
                umma_commit(&kv_empty[s]);
                umma_commit(&ds_empty[b]);
            }
            __syncwarp();
        };
        for (int j = 0; j < ntiles; ++j) {
            const int s = j % KS;
            mbar_wait(&kv_full[s], (j / KS) & 1);
            mbar_wait(&sd_empty[0], ((j & 1) ^ 1));
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), oa = smem_u32(sDO), ka = smem_u32(sK + s * kKBytes), va = smem_u32(sV + s * kKBytes);
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_bf16(tmem_base, umma_desc_sw128(qa + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(ka + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                        umma_bf16(tmem_base + 64u, umma_desc_sw128(oa + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(va + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                    }
                umma_commit(&sd_full[0]);
            }
            __syncwarp();
            if (j > 0) issue_acc(j - 1);
        }
        issue_acc(ntiles - 1);
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;   // thread = (query row, 32-key half of the tile)
        const int row = quad * 32 + lane;
        const long long grow = (long long)batch * p.tq + q0 + row;
        const float c = p.exp_scale, sc = p.scale;
        const float ra = -__ldg(p.rmax + grow) * c;
        const float dpr = -__ldg(p.Dp + grow) * sc;   // dS = P~ * (dP' * scale - D' * scale)
        const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16);
        for (int j = 0; j < ntiles; ++j) {
            const int b = j & 1;
            mbar_wait(&sd_full[0], j & 1);
            tc_fence_after();
            uint8_t* drow = sDS + b * kBig + row * 128;
            uint32_t vs[32], vd[32];
            tmem_ld32(t_row + uint32_t(half * 32), vs);
            tmem_ld32(t_row + uint32_t(64 + half * 32), vd);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sd_empty[0]);   // the accumulators are in registers: the next S | dP' may overwrite them
            mbar_wait(&ds_empty[b], (((j >> 1) & 1) ^ 1));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = 8 * u + 2 * k;
                    w[k] = fa_pack(fa_ex2(fmaf(__uint_as_float(vs[i]), c, ra)) * fmaf(__uint_as_float(vd[i]), sc, dpr),
                                   fa_ex2(fmaf(__uint_as_float(vs[i + 1]), c, ra)) * fmaf(__uint_as_float(vd[i + 1]), sc, dpr));
                }
                const int unit = half * 4 + u;
                fa_sts128(smem_u32(drow) + ((unit ^ (row & 7)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            }
            fence_proxy_async();
            mbar_arrive(&ds_full[b]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        __nv_bfloat16* orow = p.dQ + grow * DP;
#pragma unroll 1
        for (int ch = half * (DP / 64); ch < (half + 1) * (DP / 64); ++ch) {
            uint32_t v[32];
            tmem_ld32(t_row + kAccCol + uint32_t(ch * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(orow + ch * 32 + u * 8) =
                    make_uint4(fa_pack(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])),
                               fa_pack(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                               fa_pack(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])),
                               fa_pack(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

template <int DP>
__global__ void __launch_bounds__(kFaThreads, 2) attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mapK,
                                                                    const __grid_constant__ CUtensorMap mapV,
                                                                    const __grid_constant__ CUtensorMap mapQ,
                                                                    const __grid_constant__ CUtensorMap mapDO,
                                                                    const FbParams p) {
    constexpr int NC = DP / 64;
    constexpr int QS = 2;                       // Q / dO' / constants ring (a third stage measured no gain here)
    constexpr int kBig = 128 * 128, kSmall = 64 * 128;
    constexpr int kKBytes = NC * kBig, kQBytes = NC * kSmall;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;
    uint8_t* sV = sK + kKBytes;
    uint8_t* sQ = sV + kKBytes;                 // 2 stages of [64 queries][DP]
    uint8_t* sDO = sQ + QS * kQBytes;
    uint8_t* sPT = sDO + QS * kQBytes;          // [128 keys][64 queries]
    uint8_t* sDST = sPT + kBig;
    float2* cvec = reinterpret_cast<float2*>(sDST + kBig);   // [QS][64] (-m*c, -D'*scale) of a tile's queries, part of the ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(cvec + QS * 64);
    uint64_t* kv_full = bars;
    uint64_t* qd_full = bars + 16;              // [QS]
    uint64_t* qd_empty = bars + 20;             // [QS]
    uint64_t* sd_full = bars + 5;
    uint64_t* sd_empty = bars + 7;
    uint64_t* pd_full = bars + 9;
    uint64_t* pd_empty = bars + 11;
    uint64_t* acc_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, batch = blockIdx.y;
    const int ntiles = p.tq / 64;
    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < QS; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], kSmThreads);
            mbar_init(&pd_full[i], kSmThreads); mbar_init(&pd_empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    // DP = 64: 256 TMEM columns (S^T | dP'^T single-buffered + dV + dK) and 96 KB of shared memory -> two CTAs per SM
    constexpr uint32_t kTmemCols = DP == 64 ? 256 : 512;
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    if (DP == 64 && smem != smem_raw) __trap();   // the launch requests no alignment slack at DP = 64 (112 KB + constants + barriers)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kDvCol = 128, kDkCol = 128 + DP;
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * kKBytes);
            for (int c = 0; c < NC; ++c) {
                tma_load_3d(sK + c * kBig, &mapK, kv_full, c * 64, k0, batch);
                tma_load_3d(sV + c * kBig, &mapV, kv_full, c * 64, k0, batch);
            }
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % QS;
                mbar_wait(&qd_empty[s], (((i / QS) & 1) ^ 1));
                mbar_arrive_expect_tx(&qd_full[s], 2 * kQBytes + 512);
                bulk_load_1d(cvec + s * 64, p.cq + (long long)batch * p.tq + i * 64, 512, &qd_full[s]);
                for (int c = 0; c < NC; ++c) {
                    tma_load_3d(sQ + s * kQBytes + c * kSmall, &mapQ, &qd_full[s], c * 64, i * 64, batch);
                    tma_load_3d(sDO + s * kQBytes + c * kSmall, &mapDO, &qd_full[s], c * 64, i * 64, batch);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(128, 64);
        const uint32_t idesc_a = umma_idesc_bf16(128, DP) | kUmmaBMajorMN;
        mbar_wait(kv_full, 0);
        tc_fence_after();
        auto issue_acc = [&](int i) {
            const int s = i % QS;
            mbar_wait(&pd_full[0], i & 1);
            tc_fence_after();
            const uint32_t pa = smem_u32(sPT), da = smem_u32(sDST);
            const uint32_t qa = smem_u32(sQ + s * kQBytes), oa = smem_u32(sDO + s * kQBytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {   // 4 x 16 queries; dO' and Q tiles [query][channel] read MN-major
                    umma_bf16(tmem_base + kDvCol, umma_desc_sw128(pa) + uint64_t(kk * 2),
                              umma_desc_sw128_mn(oa + kk * 2048, kSmall, 1024), idesc_a, (i | kk) != 0 ? 1u : 0u);
                    umma_bf16(tmem_base + kDkCol, umma_desc_sw128(da) + uint64_t(kk * 2),
                              umma_desc_sw128_mn(qa + kk * 2048, kSmall, 1024), idesc_a, (i | kk) != 0 ? 1u : 0u);
                }
                umma_commit(&qd_empty[s]);
                umma_commit(&pd_empty[0]);
            }
            __syncwarp();
        };
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % QS;
            mbar_wait(&qd_full[s], (i / QS) & 1);
            mbar_wait(&sd_empty[0], ((i & 1) ^ 1));
            tc_fence_after();
            const uint32_t ka = smem_u32(sK), va = smem_u32(sV), qa = smem_u32(sQ + s * kQBytes), oa = smem_u32(sDO + s * kQBytes);
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        // S^T = K Q^T and dP'^T = V dO'^T: 128 keys x 64 queries
                        umma_bf16(tmem_base, umma_desc_sw128(ka + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(qa + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                        umma_bf16(tmem_base + 64u, umma_desc_sw128(va + c * kBig) + uint64_t(kk * 2),
                                  umma_desc_sw128(oa + c * kSmall) + uint64_t(kk * 2), idesc_s, (c | kk) != 0 ? 1u : 0u);
                    }
                umma_commit(&sd_full[0]);
            }
            __syncwarp();
            if (i > 0) issue_acc(i - 1);
        }
        issue_acc(ntiles - 1);
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;   // thread = (key row, 32-query half of the tile)
        const int row = quad * 32 + lane;
        const long long gkey = (long long)batch * p.tkv + k0 + row;
        const float c = p.exp_scale, sc = p.scale;
        const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16);
        for (int i = 0; i < ntiles; ++i) {
            mbar_wait(&qd_full[i % QS], (i / QS) & 1);   // (long complete: the logits below were issued after it) the constants landed
            mbar_wait(&sd_full[0], i & 1);
            tc_fence_after();
            uint8_t* prow = sPT + row * 128;
            uint8_t* drow = sDST + row * 128;
            const uint32_t cv_addr = smem_u32(cvec + (i % QS) * 64 + half * 32);
            uint32_t vs[32], vd[32];
            tmem_ld32(t_row + uint32_t(half * 32), vs);
            tmem_ld32(t_row + uint32_t(64 + half * 32), vd);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sd_empty[0]);
            mbar_wait(&pd_empty[0], ((i & 1) ^ 1));   // the previous tile's dV / dK products have read the operand tiles
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t wp[4], wd[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i2 = 8 * u + 2 * k;
                    const float4 cc = fa_lds128(cv_addr + uint32_t(i2) * 8u);   // two queries' (-m*c, -D'*scale)
                    const float2 c0 = make_float2(cc.x, cc.y), c1 = make_float2(cc.z, cc.w);
                    const float p0 = fa_ex2(fmaf(__uint_as_float(vs[i2]), c, c0.x));
                    const float p1 = fa_ex2(fmaf(__uint_as_float(vs[i2 + 1]), c, c1.x));
                    wp[k] = fa_pack(p0, p1);
                    wd[k] = fa_pack(p0 * fmaf(__uint_as_float(vd[i2]), sc, c0.y), p1 * fmaf(__uint_as_float(vd[i2 + 1]), sc, c1.y));
                }
                const int unit = half * 4 + u;
                const int off = (unit ^ (row & 7)) << 4;
                fa_sts128(smem_u32(prow) + off, make_uint4(wp[0], wp[1], wp[2], wp[3]));
                fa_sts128(smem_u32(drow) + off, make_uint4(wd[0], wd[1], wd[2], wd[3]));
            }
            fence_proxy_async();
            mbar_arrive(&pd_full[0]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        {   // half 0 writes dV, half 1 writes dK
            __nv_bfloat16* orow = (half ? p.dK : p.dV) + gkey * DP;
            const uint32_t col0 = half ? kDkCol : kDvCol;
#pragma unroll 1
            for (int ch = 0; ch < DP / 32; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_row + col0 + uint32_t(ch * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(orow + ch * 32 + u * 8) =
                        make_uint4(fa_pack(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])),
                                   fa_pack(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                                   fa_pack(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])),
                                   fa_pack(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

template <int DP>
int launch_fb(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, const __nv_bfloat16* dOs,
              const float* rmax, const float* Dp, const float2* cq, __nv_bfloat16* dQ, __nv_bfloat16* dK, __nv_bfloat16* dV, int nb, int tq,
              int tkv, float scale, cudaStream_t st) {
    CUtensorMap mq128, mo128, mk64, mv64, mk128, mv128, mq64, mo64;
    int rc;
    cuuint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    cuuint64_t dq[3] = {(cuuint64_t)DP, (cuuint64_t)tq, (cuuint64_t)nb}, sq[2] = {(cuuint64_t)DP * 2, (cuuint64_t)tq * DP * 2};
    cuuint64_t dk[3] = {(cuuint64_t)DP, (cuuint64_t)tkv, (cuuint64_t)nb}, sk[2] = {(cuuint64_t)DP * 2, (cuuint64_t)tkv * DP * 2};
    if ((rc = encode_map_bf16_sw128(&mq128, Q, 3, dq, sq, box128, "attn.bwd.Q"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mo128, dOs, 3, dq, sq, box128, "attn.bwd.dO"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mq64, Q, 3, dq, sq, box64, "attn.bwd.Q64"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mo64, dOs, 3, dq, sq, box64, "attn.bwd.dO64"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mk128, K, 3, dk, sk, box128, "attn.bwd.K"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mv128, V, 3, dk, sk, box128, "attn.bwd.V"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mk64, K, 3, dk, sk, box64, "attn.bwd.K64"))) return rc;
    if ((rc = encode_map_bf16_sw128(&mv64, V, 3, dk, sk, box64, "attn.bwd.V64"))) return rc;
    constexpr int NC = DP / 64, kBig = 128 * 128, kSmall = 64 * 128;
    constexpr int KS = DP == 64 ? 3 : 2;
    constexpr size_t smem_dq = size_t(2 * NC) * kBig + size_t(2 * KS * NC) * kSmall + 2 * kBig + 256 + (DP == 64 ? 0 : 1024);
    constexpr size_t smem_dkv = size_t(2 * NC) * kBig + size_t(4 * NC) * kSmall + 2 * kBig + 1024 + 256 + 1024;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv);
        if (e != cudaSuccess) { set_error("attn.bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -4; }
        attr_set[dev & 63] = true;
    }
    FbParams p;
    p.rmax = rmax; p.Dp = Dp; p.cq = cq; p.dQ = dQ; p.dK = dK; p.dV = dV; p.tq = tq; p.tkv = tkv;
    p.scale = scale; p.exp_scale = scale * 1.4426950408889634f;
    attn_bwd_dq_kernel<DP><<<dim3(tq / 128, nb), kFaThreads, smem_dq, st>>>(mq128, mo128, mk64, mv64, p);
    if (dK != nullptr && dV != nullptr)   // (cross attention: keys / values are constants of the attack)
        attn_bwd_dkv_kernel<DP><<<dim3(tkv / 128, nb), kFaThreads, smem_dkv, st>>>(mk128, mv128, mq64, mo64, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("attn.bwd: launch failed: %s", cudaGetErrorString(e)); return -5; }
    count_launch();
    count_launch();
    return 0;
}

}  // namespace

bool attn_fused_supported(int tq, int tkv, int dp) {
    static const bool off = getenv("TML_NO_FUSED_ATTN") && getenv("TML_NO_FUSED_ATTN")[0] == '1';   // A/B switch
    return !off && (dp == 64 || dp == 128) && tq % 128 == 0 && tkv % 128 == 0 && tkv >= 128;
}

// O[nb][tq][dp] = softmax(Q K^T * scale) V given the row maxima of Q K^T; inv_l[nb][tq] = 1 / row sums of P~.
// Channel `lcol` (< dp) of V must hold 1.0 for every key (a padding channel of the head): O[:, lcol] is then the row sum.
// rmax != null: known row maxima (a preceding max pass); rmax == null: online softmax, the reference maxima the kernel
// settled on are written to rmax_out (what the backward must recompute P~ against).
int launch_attn_fused_fwd(const bf16* Q, const bf16* K, const bf16* V, const float* rmax, float* rmax_out, float* inv_l,
                          bf16* O, int nb, int tq, int tkv, int dp, int lcol, float scale, cudaStream_t st) {
    if (g_dry_run) return 0;
    if (!attn_fused_supported(tq, tkv, dp) || lcol < 0 || lcol >= dp) {
        set_error("attn.fused: unsupported shape tq=%d tkv=%d dp=%d lcol=%d", tq, tkv, dp, lcol);
        return -1;
    }
    if (rmax == nullptr && rmax_out == nullptr) { set_error("attn.fused: no row-maximum buffer"); return -1; }
    if (dp == 64) return launch_fa<64>(Q, K, V, rmax, rmax_out, inv_l, O, nb, tq, tkv, lcol, scale, st);
    return launch_fa<128>(Q, K, V, rmax, rmax_out, inv_l, O, nb, tq, tkv, lcol, scale, st);
}

// Fused backward of the same attention: dO [nb][tq][dp], O, inv_l, rmax from the forward; dOs (bf16 [nb][tq][dp]) and
// Dp (fp32, 3 * nb * tq + 4 floats: D' then the dK/dV kernel's per-query constants) are scratch.  Writes dQ [nb][tq][dp], dK and dV [nb][tkv][dp] (both null: dQ only).
int launch_attn_fused_bwd(const bf16* Q, const bf16* K, const bf16* V, const bf16* O, const bf16* dO, const float* rmax,
                          const float* inv_l, bf16* dOs, float* Dp, bf16* dQ, bf16* dK, bf16* dV, int nb, int tq, int tkv,
                          int dp, float scale, cudaStream_t st) {
    if (g_dry_run) return 0;
    if (!attn_fused_supported(tq, tkv, dp)) { set_error("attn.bwd: unsupported shape tq=%d tkv=%d dp=%d", tq, tkv, dp); return -1; }
    const long long rows = (long long)nb * tq;
    // Dp scratch: [rows] D' followed by [rows] float2 per-query constants of the dK/dV kernel (3 floats per row)
    float2* cq = reinterpret_cast<float2*>(Dp + ((rows + 3) & ~3LL));
    attn_bwd_prep_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>(dO, O, inv_l, rmax, dOs, Dp, cq, scale * 1.4426950408889634f,
                                                                    scale, rows, dp);
    count_launch();
    if (dp == 64) return launch_fb<64>(Q, K, V, dOs, rmax, Dp, cq, dQ, dK, dV, nb, tq, tkv, scale, st);
    return launch_fb<128>(Q, K, V, dOs, rmax, Dp, cq, dQ, dK, dV, nb, tq, tkv, scale, st);
}

}  // namespace tml
