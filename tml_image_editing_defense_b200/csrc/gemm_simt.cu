// SIMT implicit-GEMM with the same operation semantics as gemm_tc.cu.  TEST/DEBUG AID ONLY: it lets
// the GPU tests separate "is the tcgen05 pipeline right" from "is the layer plan right".  It is
// never selected unless a test calls tml_debug_set_gemm_impl(1).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gemm.h"

namespace tml {

struct SimtParams {
    const __nv_bfloat16* A;
    int A_C, A_W, A_H, A_B;
    long long A_sW, A_sH, A_sB, A_sK;
    int a_trans;
    int stride, ntaps;
    int dh[kMaxTaps], dw[kMaxTaps];
    int OW, OH;
    const __nv_bfloat16* Bm;
    int N;
    long long B_sN, B_sBatch;
    float alpha;
    const float* bias;
    const __nv_bfloat16* resid;
    long long R_sB, R_sH, R_sW;
    void* D;
    int out_fp32;
    long long D_sB, D_sH, D_sW, D_sN;
    int n_store;
    float beta;
};

__global__ void conv_gemm_simt_kernel(const SimtParams p) {
    const long long total = (long long)p.A_B * p.OH * p.OW * p.N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = int(idx % p.N);
        long long m = idx / p.N;
        const int ow = int(m % p.OW); m /= p.OW;
        const int oh = int(m % p.OH);
        const int b = int(m / p.OH);
        if (n >= p.n_store) continue;
        const __nv_bfloat16* brow = p.Bm + (long long)b * p.B_sBatch + (long long)n * p.B_sN;
        float acc = 0.f;
        if (p.a_trans) {   // A stored [batch][k][m], m = oh*OW + ow
            const __nv_bfloat16* acol = p.A + (long long)b * p.A_sB + (long long)oh * p.OW + ow;
            for (int c = 0; c < p.A_C; ++c)
                acc = fmaf(__bfloat162float(acol[(long long)c * p.A_sK]), __bfloat162float(brow[c]), acc);
        } else
        for (int t = 0; t < p.ntaps; ++t) {
            const int ih = oh * p.stride + p.dh[t], iw = ow * p.stride + p.dw[t];
            if (ih < 0 || ih >= p.A_H || iw < 0 || iw >= p.A_W) continue;
            const __nv_bfloat16* arow = p.A + (long long)b * p.A_sB + (long long)ih * p.A_sH + (long long)iw * p.A_sW;
            const __nv_bfloat16* bk = brow + (long long)t * p.A_C;
            for (int c = 0; c < p.A_C; ++c) acc = fmaf(__bfloat162float(arow[c]), __bfloat162float(bk[c]), acc);
        }
        float v = acc * p.alpha;
        if (p.bias) v += p.bias[n];
        if (p.resid) v += __bfloat162float(p.resid[(long long)b * p.R_sB + (long long)oh * p.R_sH + (long long)ow * p.R_sW + n]);
        const long long o = (long long)b * p.D_sB + (long long)oh * p.D_sH + (long long)ow * p.D_sW + (long long)n * p.D_sN;
        if (p.out_fp32) reinterpret_cast<float*>(p.D)[o] = p.beta != 0.f ? fmaf(p.beta, reinterpret_cast<float*>(p.D)[o], v) : v;
        else reinterpret_cast<__nv_bfloat16*>(p.D)[o] = __float2bfloat16_rn(v);
    }
}

int gemm_launch_simt(const GemmOp& op, cudaStream_t stream) {
    SimtParams p;
    p.A = reinterpret_cast<const __nv_bfloat16*>(op.A);
    p.A_C = op.A_C; p.A_W = op.A_W; p.A_H = op.A_H; p.A_B = op.A_B;
    p.A_sW = op.A_sW; p.A_sH = op.A_sH; p.A_sB = op.A_sB; p.A_sK = op.A_sK; p.a_trans = op.a_trans;
    p.stride = op.stride; p.ntaps = op.ntaps;
    for (int i = 0; i < kMaxTaps; ++i) { p.dh[i] = op.dh[i]; p.dw[i] = op.dw[i]; }
    p.OW = op.OW; p.OH = op.OH;
    p.Bm = reinterpret_cast<const __nv_bfloat16*>(op.Bm);
    p.N = op.N; p.B_sN = op.B_sN; p.B_sBatch = op.B_sBatch;
    p.alpha = op.alpha; p.bias = op.bias;
    p.resid = reinterpret_cast<const __nv_bfloat16*>(op.resid);
    p.R_sB = op.R_sB; p.R_sH = op.R_sH; p.R_sW = op.R_sW;
    p.D = op.D; p.out_fp32 = op.out_fp32;
    p.D_sB = op.D_sB; p.D_sH = op.D_sH; p.D_sW = op.D_sW; p.D_sN = op.D_sN;
    p.n_store = op.n_store > 0 ? op.n_store : op.N;
    p.beta = op.out_fp32 ? op.beta : 0.f;
    const long long total = (long long)op.A_B * op.OH * op.OW * op.N;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (blocks < 1) blocks = 1;
    conv_gemm_simt_kernel<<<(int)blocks, 256, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s: simt launch failed: %s", op.name, cudaGetErrorString(e)); return -5; }
    return 0;
}

}  // namespace tml
