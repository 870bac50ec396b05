// AutoencoderKL encoder forward + input-gradient backward as a sequence of sm_100a kernels, and the
// C ABI declared in include/tml_b200.h.
//
// Layout in HBM: activations and activation gradients are bf16 NHWC (channel-contiguous: the K
// dimension of every implicit GEMM is contiguous, which is what TMA + the K-major UMMA descriptor
// want); the image, its gradient, the moments and the PGD iterate stay fp32 NCHW exactly as the
// reference holds them (main.py:33).  `saved` holds what the backward needs: the input of every
// GroupNorm (pre-norm conv outputs), the per-image GroupNorm statistics, qkv and the attention
// probabilities.  There is no weight gradient anywhere: torch.autograd.grad(loss, [cur_image])
// (main.py:176) only asks for the input gradient, so every backward conv is a dgrad-only GEMM.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/tml_b200.h"
#include "gemm.h"
#include "kernels.h"

using namespace tml;

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
static uint16_t f2bf(float f) {  // round-to-nearest-even, NaN preserved
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float half2f(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, u;
    if (e == 0) {
        if (m == 0) u = s << 31;
        else { int sh = 0; while (!(m & 1024)) { m <<= 1; ++sh; } m &= 1023; u = (s << 31) | ((113 - sh) << 23) | (m << 13); }
    } else if (e == 31) u = (s << 31) | 0x7F800000u | (m << 13);
    else u = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

#define CUDA_OK(call)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -10;                                                                      \
        }                                                                                    \
    } while (0)
#define RC(call)              \
    do {                      \
        int _r = (call);      \
        if (_r) return _r;    \
    } while (0)

struct HostTensor {
    std::vector<float> v;
    std::vector<int64_t> shape;
};

// 3x3 tap tables + packed matrices -------------------------------------------------------------
// mode 0: forward s1 p1      B[co][t*Ci+ci] = W[co][ci][r][s], t=r*3+s, (dh,dw) = (r-1, s-1)
// mode 1: dgrad of s1 p1     B[ci][t*Co+co] = W[co][ci][r][s],           (dh,dw) = (1-r, 1-s)
// mode 2: forward s2, pad (0,1,0,1): same matrix as mode 0,             (dh,dw) = (r, s)
// mode 3..6: dgrad of s2 for output parity (ph,pw): dX[2i+ph, 2j+pw] = sum over taps with
//            r = ph (mod 2), s = pw (mod 2) of dY[i + (ph-r)/2, j + (pw-s)/2] * W[co][ci][r][s]
static int pack_conv3x3(const float* w, int Co, int Ci, int mode, std::vector<uint16_t>& out, int* ntaps, int* dh,
                        int* dw) {
    auto W = [&](int co, int ci, int r, int s) { return w[(((size_t)co * Ci + ci) * 3 + r) * 3 + s]; };
    if (mode == 0 || mode == 2) {
        *ntaps = 9;
        out.assign((size_t)Co * 9 * Ci, 0);
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                const int t = r * 3 + s;
                dh[t] = mode == 0 ? r - 1 : r;
                dw[t] = mode == 0 ? s - 1 : s;
                for (int co = 0; co < Co; ++co)
                    for (int ci = 0; ci < Ci; ++ci) out[(size_t)co * 9 * Ci + (size_t)t * Ci + ci] = f2bf(W(co, ci, r, s));
            }
        return 0;
    }
    if (mode == 1) {
        *ntaps = 9;
        out.assign((size_t)Ci * 9 * Co, 0);
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                const int t = r * 3 + s;
                dh[t] = 1 - r;
                dw[t] = 1 - s;
                for (int ci = 0; ci < Ci; ++ci)
                    for (int co = 0; co < Co; ++co) out[(size_t)ci * 9 * Co + (size_t)t * Co + co] = f2bf(W(co, ci, r, s));
            }
        return 0;
    }
    if (mode >= 3 && mode <= 6) {
        const int ph = (mode - 3) >> 1, pw = (mode - 3) & 1;
        int rs[2], ss[2], nr = 0, ns = 0;
        for (int r = 0; r < 3; ++r) if ((r & 1) == ph) rs[nr++] = r;
        for (int s = 0; s < 3; ++s) if ((s & 1) == pw) ss[ns++] = s;
        *ntaps = nr * ns;
        out.assign((size_t)Ci * (*ntaps) * Co, 0);
        int t = 0;
        for (int a = 0; a < nr; ++a)
            for (int b = 0; b < ns; ++b, ++t) {
                const int r = rs[a], s = ss[b];
                dh[t] = (ph - r) / 2;  // 0 or -1
                dw[t] = (pw - s) / 2;
                for (int ci = 0; ci < Ci; ++ci)
                    for (int co = 0; co < Co; ++co)
                        out[(size_t)ci * (*ntaps) * Co + (size_t)t * Co + co] = f2bf(W(co, ci, r, s));
            }
        return 0;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// device-side layer parameters
// ------------------------------------------------------------------------------------------------
struct Packed {
    bf16* w = nullptr;  // [N][ntaps*K]
    int ntaps = 1;
    int dh[kMaxTaps] = {0}, dw[kMaxTaps] = {0};
};
struct Conv3 {
    int ci = 0, co = 0, stride = 1;
    Packed fwd, bwd, bwd_par[4];
    float* bias = nullptr;
};
struct Lin {  // 1x1 conv or linear: fwd [co][ci], bwd [ci][co]
    int ci = 0, co = 0;
    bf16* fwd = nullptr;
    bf16* bwd = nullptr;
    float* bias = nullptr;
};
struct Norm {
    int C = 0;
    float* gamma = nullptr;
    float* beta = nullptr;
};
struct Resnet {
    int ci = 0, co = 0;
    Norm n1, n2;
    Conv3 c1, c2;
    bool has_sc = false;
    Lin sc;
};
struct Attn {
    int C = 0;
    Norm gn;
    Lin qkv;  // fwd [3C][C], bwd [C][3C]
    Lin out;
};

// saved-state record of one GroupNorm: scale/shift per (b,c) and mean/rstd per (b,g)
struct GnSaved { size_t ss = 0, mr = 0; };
struct ResnetRec { size_t x = 0, h1 = 0, out = 0; GnSaved g1, g2; int h = 0, w = 0; };
struct DownRec { size_t x = 0, out = 0; int h = 0, w = 0; };
struct AttnRec { size_t x = 0, qkv = 0, P = 0, out = 0; GnSaved g; int h = 0, w = 0; };

struct Arena {
    size_t off = 0, peak = 0;
    size_t alloc(size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) & ~size_t(255);
        if (off > peak) peak = off;
        return o;
    }
    size_t mark() const { return off; }
    void reset(size_t m) { off = m; }
};

struct Layout {
    int B = 0, H = 0, W = 0;
    size_t saved_bytes = 0, ws_bytes = 0;
    size_t x0 = 0;  // conv_in output
    std::vector<ResnetRec> res;
    std::vector<DownRec> down;
    AttnRec attn;
    GnSaved gout;
    size_t xlast = 0;
    int hl = 0, wl = 0;
};

struct TmlEncoder {
    TmlEncoderCfg cfg;
    int device = 0;
    int num_sms = 148;
    bool finalized = false;
    std::map<std::string, HostTensor> host;
    std::vector<void*> dev_allocs;
    // parameters
    float* conv_in_w = nullptr;  // [27][C0] fp32
    float* conv_in_b = nullptr;
    Packed conv_in_bwd;          // dgrad as a GEMM: B[ci (3, padded to 16)][t*C0 + co]
    std::vector<Resnet> resnets;          // in forward order (down blocks then mid[0], mid[1])
    std::vector<Conv3> downs;
    bool has_attn = false;
    Attn attn;
    Norm norm_out;
    Conv3 conv_out;  // folded with quant_conv, N padded to 16; bwd has K = 64 (8 real)
    Layout lay;
};

// ------------------------------------------------------------------------------------------------
// weight upload
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(TmlEncoder* e, const std::vector<T>& h, T** out) {
    void* d = nullptr;
    CUDA_OK(cudaMalloc(&d, h.size() * sizeof(T) + 256));
    CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    e->dev_allocs.push_back(d);
    *out = reinterpret_cast<T*>(d);
    return 0;
}
static int upload_bf16(TmlEncoder* e, const std::vector<uint16_t>& h, bf16** out) {
    uint16_t* d = nullptr;
    RC(upload<uint16_t>(e, h, &d));
    *out = reinterpret_cast<bf16*>(d);
    return 0;
}

static const HostTensor* find(TmlEncoder* e, const std::string& key, size_t expect_numel) {
    auto it = e->host.find(key);
    if (it == e->host.end()) { set_error("missing weight '%s'", key.c_str()); return nullptr; }
    if (it->second.v.size() != expect_numel) {
        set_error("weight '%s' has %zu elements, expected %zu", key.c_str(), it->second.v.size(), expect_numel);
        return nullptr;
    }
    return &it->second;
}

static int make_norm(TmlEncoder* e, const std::string& key, int C, Norm* n) {
    const HostTensor* g = find(e, key + ".weight", C);
    const HostTensor* b = find(e, key + ".bias", C);
    if (!g || !b) return -20;
    n->C = C;
    RC(upload<float>(e, g->v, &n->gamma));
    RC(upload<float>(e, b->v, &n->beta));
    return 0;
}

static int make_packed(TmlEncoder* e, const float* w, int Co, int Ci, int mode, Packed* p) {
    std::vector<uint16_t> h;
    if (pack_conv3x3(w, Co, Ci, mode, h, &p->ntaps, p->dh, p->dw)) { set_error("pack mode %d", mode); return -21; }
    return upload_bf16(e, h, &p->w);
}

static int make_conv3(TmlEncoder* e, const std::string& key, int Ci, int Co, int stride, Conv3* c) {
    const HostTensor* w = find(e, key + ".weight", (size_t)Co * Ci * 9);
    const HostTensor* b = find(e, key + ".bias", Co);
    if (!w || !b) return -20;
    c->ci = Ci; c->co = Co; c->stride = stride;
    RC(make_packed(e, w->v.data(), Co, Ci, stride == 1 ? 0 : 2, &c->fwd));
    if (stride == 1) RC(make_packed(e, w->v.data(), Co, Ci, 1, &c->bwd));
    else for (int q = 0; q < 4; ++q) RC(make_packed(e, w->v.data(), Co, Ci, 3 + q, &c->bwd_par[q]));
    RC(upload<float>(e, b->v, &c->bias));
    return 0;
}

// W: [co][ci] row-major
static int make_lin_from(TmlEncoder* e, const std::vector<float>& W, const std::vector<float>& bias, int Ci, int Co,
                         Lin* l) {
    l->ci = Ci; l->co = Co;
    std::vector<uint16_t> f((size_t)Co * Ci), t((size_t)Ci * Co);
    for (int o = 0; o < Co; ++o)
        for (int i = 0; i < Ci; ++i) {
            const uint16_t v = f2bf(W[(size_t)o * Ci + i]);
            f[(size_t)o * Ci + i] = v;
            t[(size_t)i * Co + o] = v;
        }
    RC(upload_bf16(e, f, &l->fwd));
    RC(upload_bf16(e, t, &l->bwd));
    RC(upload<float>(e, bias, &l->bias));
    return 0;
}

static int make_resnet(TmlEncoder* e, const std::string& key, int Ci, int Co, Resnet* r) {
    r->ci = Ci; r->co = Co;
    RC(make_norm(e, key + ".norm1", Ci, &r->n1));
    RC(make_conv3(e, key + ".conv1", Ci, Co, 1, &r->c1));
    RC(make_norm(e, key + ".norm2", Co, &r->n2));
    RC(make_conv3(e, key + ".conv2", Co, Co, 1, &r->c2));
    r->has_sc = Ci != Co;
    if (r->has_sc) {
        const HostTensor* w = find(e, key + ".conv_shortcut.weight", (size_t)Co * Ci);
        const HostTensor* b = find(e, key + ".conv_shortcut.bias", Co);
        if (!w || !b) return -20;
        RC(make_lin_from(e, w->v, b->v, Ci, Co, &r->sc));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout (what lives where in `saved`; how much scratch the walks need)
// ------------------------------------------------------------------------------------------------
static size_t act_bytes(int B, int h, int w, int c) { return (size_t)B * h * w * c * sizeof(bf16); }
// partial sums written by a GEMM epilogue: one entry per (image, 128-row tile, group)
static size_t fused_partial_bytes(int B, int h, int w) { return (size_t)B * gemm_gn_tiles_per_image(h, w) * 32 * 2 * sizeof(float); }
// sized for the smallest channel count (largest chunk count) so one bound covers every layer
static size_t gn_partial_bytes(int B, int hw) { return (size_t)B * gn_num_chunks(hw, 512) * 32 * 2 * sizeof(float); }

static GnSaved alloc_gn(Arena& a, int B, int C) {
    GnSaved g;
    g.ss = a.alloc((size_t)B * C * sizeof(float2));
    g.mr = a.alloc((size_t)B * 32 * sizeof(float2));
    return g;
}

static int build_layout(TmlEncoder* e, int B, int H, int W) {
    Layout& L = e->lay;
    if (L.B == B && L.H == H && L.W == W && L.saved_bytes) return 0;
    const TmlEncoderCfg& c = e->cfg;
    const int nb = c.num_blocks;
    const int down_factor = 1 << (nb - 1);
    if (H % down_factor || W % down_factor) { set_error("H,W must be multiples of %d", down_factor); return -30; }
    if ((W / down_factor) % 8) { set_error("W/%d must be a multiple of 8 (got W=%d)", down_factor, W); return -30; }
    L = Layout();
    L.B = B; L.H = H; L.W = W;
    Arena S, WS;
    int h = H, w = W;
    L.x0 = S.alloc(act_bytes(B, h, w, c.block_out_channels[0]));
    size_t cur = L.x0;
    size_t ws_peak = 0;
    auto resnet_ws = [&](int ci, int co, int hh, int ww) {
        // forward: partial + a + a2 + sc ; backward: d_a2 + partial + mm + d_h1 + d_a1 + tmp
        const size_t f = gn_partial_bytes(B, hh * ww) + act_bytes(B, hh, ww, ci) + 2 * act_bytes(B, hh, ww, co) + 4096;
        const size_t b = 2 * act_bytes(B, hh, ww, co) + 2 * act_bytes(B, hh, ww, ci) + gn_partial_bytes(B, hh * ww) +
                         (size_t)B * 32 * sizeof(float2) + 8192;
        return f > b ? f : b;
    };
    int cin = c.block_out_channels[0];
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[i];
        for (int j = 0; j < c.layers_per_block; ++j) {
            ResnetRec r;
            r.h = h; r.w = w; r.x = cur;
            r.g1 = alloc_gn(S, B, cin);
            r.h1 = S.alloc(act_bytes(B, h, w, cout));
            r.g2 = alloc_gn(S, B, cout);
            r.out = S.alloc(act_bytes(B, h, w, cout));
            ws_peak = std::max(ws_peak, resnet_ws(cin, cout, h, w));
            L.res.push_back(r);
            cur = r.out;
            cin = cout;
        }
        if (i != nb - 1) {
            DownRec d;
            d.h = h; d.w = w; d.x = cur;
            h /= 2; w /= 2;
            d.out = S.alloc(act_bytes(B, h, w, cout));
            L.down.push_back(d);
            cur = d.out;
        }
    }
    // mid block
    for (int m = 0; m < 2; ++m) {
        ResnetRec r;
        r.h = h; r.w = w; r.x = cur;
        r.g1 = alloc_gn(S, B, cin);
        r.h1 = S.alloc(act_bytes(B, h, w, cin));
        r.g2 = alloc_gn(S, B, cin);
        r.out = S.alloc(act_bytes(B, h, w, cin));
        ws_peak = std::max(ws_peak, resnet_ws(cin, cin, h, w));
        L.res.push_back(r);
        cur = r.out;
        if (m == 0 && c.mid_block_add_attention) {
            AttnRec a;
            a.h = h; a.w = w; a.x = cur;
            const size_t tok = (size_t)h * w;
            a.g = alloc_gn(S, B, cin);
            a.qkv = S.alloc((size_t)B * tok * 3 * cin * sizeof(bf16));
            a.P = S.alloc((size_t)B * tok * tok * sizeof(bf16));
            a.out = S.alloc(act_bytes(B, h, w, cin));
            const size_t act = act_bytes(B, h, w, cin);
            const size_t fwd_ws = gn_partial_bytes(B, (int)tok) + act /*t*/ + (size_t)B * tok * tok * 4 /*S*/ + act /*Vt*/ + act /*a*/ + 8192;
            const size_t bwd_ws = 2 * act /*da, daT*/ + (size_t)B * tok * tok * 4 /*dP*/ + 3 * (size_t)B * tok * tok * 2 /*dS,dST,PT*/ +
                                  2 * act /*Kt,Qt*/ + 3 * act /*dqkv*/ + act /*dt*/ + gn_partial_bytes(B, (int)tok) + 16384;
            ws_peak = std::max(ws_peak, std::max(fwd_ws, bwd_ws));
            L.attn = a;
            cur = a.out;
        }
    }
    L.gout = alloc_gn(S, B, cin);
    L.xlast = cur;
    L.hl = h; L.wl = w;
    // final: partial + a (fwd); dm64 + d_a + partial + mm (bwd)
    ws_peak = std::max(ws_peak, gn_partial_bytes(B, h * w) + 2 * act_bytes(B, h, w, cin) + act_bytes(B, h, w, 64) + 16384);
    // two ping-pong gradient buffers of the largest activation
    size_t gmax = 0;
    {
        int hh = H, ww = W;
        for (int i = 0; i < nb; ++i) {
            gmax = std::max(gmax, act_bytes(B, hh, ww, c.block_out_channels[i]));
            if (i) gmax = std::max(gmax, act_bytes(B, hh, ww, c.block_out_channels[i - 1]));
            if (i != nb - 1) { hh /= 2; ww /= 2; }
        }
    }
    L.saved_bytes = S.peak + 256;
    L.ws_bytes = ws_peak + 2 * (gmax + 256) + 3 * (fused_partial_bytes(B, H, W) + 256) + (64 << 10);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// op builders
// ------------------------------------------------------------------------------------------------
static GemmOp dense_conv_op(const char* name, const bf16* A, int B, int ih, int iw, int Ci, const Packed& pk, int N,
                            int stride, int oh, int ow, const float* bias, const bf16* resid, bf16* D) {
    GemmOp o;
    o.name = name;
    o.A = A; o.A_C = Ci; o.A_W = iw; o.A_H = ih; o.A_B = B;
    o.A_sW = Ci; o.A_sH = (int64_t)iw * Ci; o.A_sB = (int64_t)ih * iw * Ci;
    o.stride = stride; o.ntaps = pk.ntaps;
    for (int t = 0; t < pk.ntaps; ++t) { o.dh[t] = pk.dh[t]; o.dw[t] = pk.dw[t]; }
    o.OW = ow; o.OH = oh;
    o.Bm = pk.w; o.N = N; o.B_sN = (int64_t)pk.ntaps * Ci; o.B_sBatch = 0;
    o.bias = bias;
    o.resid = resid; o.R_sW = N; o.R_sH = (int64_t)ow * N; o.R_sB = (int64_t)oh * ow * N;
    o.D = D; o.D_sW = N; o.D_sH = (int64_t)ow * N; o.D_sB = (int64_t)oh * ow * N; o.D_sN = 1;
    return o;
}
static GemmOp dense_lin_op(const char* name, const bf16* A, int B, int h, int w, int K, const bf16* Wm, int N,
                           const float* bias, const bf16* resid, bf16* D) {
    Packed pk;
    pk.w = const_cast<bf16*>(Wm);
    pk.ntaps = 1;
    return dense_conv_op(name, A, B, h, w, K, pk, N, 1, h, w, bias, resid, D);
}

// test hook: copy every backward stage's output gradient into consecutive slots of a caller buffer
static char* g_dump_base = nullptr;
static size_t g_dump_slot = 0;
static int g_dump_slots = 0, g_dump_next = 0;
static void dump_grad(const void* p, size_t bytes, cudaStream_t st) {
    if (!g_dump_base || g_dump_next >= g_dump_slots) return;
    cudaMemcpyAsync(g_dump_base + (size_t)g_dump_next * g_dump_slot, p, bytes < g_dump_slot ? bytes : g_dump_slot,
                    cudaMemcpyDeviceToDevice, st);
    ++g_dump_next;
}

// Partial GroupNorm sums already produced by a GEMM epilogue (null = run the reduction kernel).
struct Partials {
    float* p = nullptr;
    int nchunks = 0;
};

struct Run {
    TmlEncoder* e;
    char* saved;
    char* ws;
    Arena wsa;
    cudaStream_t st;
    int B;
    float* statbuf[2] = {nullptr, nullptr};  // forward: ping-pong buffers for epilogue-fused GroupNorm statistics
    Partials pending;                         // statistics of the current activation, if its producer fused them
    template <typename T> T* S(size_t off) const { return reinterpret_cast<T*>(saved + off); }
    template <typename T> T* Walloc(size_t bytes) { return reinterpret_cast<T*>(ws + wsa.alloc(bytes)); }
};


static int gn_forward(Run& r, const bf16* x, const Norm& n, const GnSaved& g, bf16* y, int hw, int silu,
                      const Partials& pre = Partials()) {
    const size_t m = r.wsa.mark();
    const float* part = pre.p;
    int nchunks = pre.nchunks;
    if (!part) {
        float* own = r.Walloc<float>(gn_partial_bytes(r.B, hw));
        launch_gn_stats(x, own, r.B, hw, n.C, r.st);
        part = own;
        nchunks = gn_num_chunks(hw, n.C);
    }
    launch_gn_finalize(part, n.gamma, n.beta, r.S<float2>(g.ss), r.S<float2>(g.mr), r.B, hw, n.C, r.e->cfg.norm_eps,
                       nchunks, r.st);
    launch_gn_apply(x, r.S<float2>(g.ss), y, r.B, hw, n.C, silu, r.st);
    r.wsa.reset(m);
    return 0;
}
static int gn_backward(Run& r, const bf16* x, const bf16* dy, const Norm& n, const GnSaved& g, const bf16* resid,
                       bf16* dx, int hw, int silu, const Partials& pre = Partials()) {
    const size_t m = r.wsa.mark();
    const float* part = pre.p;
    int nchunks = pre.nchunks;
    if (!part) {
        float* own = r.Walloc<float>(gn_partial_bytes(r.B, hw));
        launch_gn_bwd_partial(x, dy, r.S<float2>(g.ss), r.S<float2>(g.mr), n.gamma, own, r.B, hw, n.C, silu, r.st);
        part = own;
        nchunks = gn_num_chunks(hw, n.C);
    }
    float2* mm = r.Walloc<float2>((size_t)r.B * 32 * sizeof(float2));
    launch_gn_bwd_finalize(part, mm, r.B, hw, n.C, nchunks, r.st);
    launch_gn_bwd_apply(x, dy, r.S<float2>(g.ss), r.S<float2>(g.mr), mm, n.gamma, resid, dx, r.B, hw, n.C, silu, r.st);
    r.wsa.reset(m);
    return 0;
}
// Ask a GEMM to also reduce (sum, sumsq) of its output per (image, tile, group).
static bool env_off(const char* name) {
    const char* v = getenv(name);
    return v && v[0] == '1';
}
static Partials fuse_stats(GemmOp& o, float* buf, int oh, int ow) {
    static const bool off = env_off("TML_NO_FUSE_STATS");     // tuning switch
    if (off || gemm_get_impl() != 0) return Partials();  // the SIMT debug kernel has no fused reductions
    o.gn_mode = 1;
    o.gn_partial = buf;
    Partials p;
    p.p = buf;
    p.nchunks = gemm_gn_tiles_per_image(oh, ow);
    return p;
}
// Ask a dgrad GEMM to also reduce the GroupNorm-backward sums of the norm whose output it differentiates.
static Partials fuse_gn_bwd(Run& r, GemmOp& o, const bf16* x, const Norm& n, const GnSaved& g, int silu, float* buf,
                            int oh, int ow) {
    static const bool off = env_off("TML_NO_FUSE_GNBWD");     // tuning switch
    // Measured on B200: for K = 9*128 the main loop of a tile is too short to hide this epilogue (exp + rcp per
    // element and an extra read of x), so 128-channel layers keep the standalone reduction kernel.
    static const int min_k = getenv("TML_GNBWD_MIN_K") ? atoi(getenv("TML_GNBWD_MIN_K")) : 2000;
    if (off || gemm_get_impl() != 0 || o.ntaps * o.A_C < min_k) return Partials();
    o.gn_mode = 2;
    o.gn_partial = buf;
    o.gn_x = x;
    o.gn_ss = r.S<float2>(g.ss);
    o.gn_mr = r.S<float2>(g.mr);
    o.gn_gamma = n.gamma;
    o.gn_silu = silu;
    Partials p;
    p.p = buf;
    p.nchunks = gemm_gn_tiles_per_image(oh, ow);
    return p;
}

static int resnet_forward(Run& r, const Resnet& p, const ResnetRec& rec) {
    const int B = r.B, h = rec.h, w = rec.w, hw = h * w;
    const int ns = r.e->num_sms;
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* a = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
    RC(gn_forward(r, x, p.n1, rec.g1, a, hw, 1, r.pending));
    bf16* h1 = r.S<bf16>(rec.h1);
    GemmOp c1 = dense_conv_op("resnet.conv1", a, B, h, w, p.ci, p.c1.fwd, p.co, 1, h, w, p.c1.bias, nullptr, h1);
    const Partials s1 = fuse_stats(c1, r.statbuf[0], h, w);     // statistics of h1 for norm2, from the epilogue
    RC(gemm_launch(c1, ns, r.st));
    bf16* a2 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    RC(gn_forward(r, h1, p.n2, rec.g2, a2, hw, 1, s1));
    const bf16* resid = x;
    if (p.has_sc) {
        bf16* sc = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
        RC(gemm_launch(dense_lin_op("resnet.shortcut", x, B, h, w, p.ci, p.sc.fwd, p.co, p.sc.bias, nullptr, sc), ns, r.st));
        resid = sc;
    }
    GemmOp c2 = dense_conv_op("resnet.conv2", a2, B, h, w, p.co, p.c2.fwd, p.co, 1, h, w, p.c2.bias, resid, r.S<bf16>(rec.out));
    r.pending = fuse_stats(c2, r.statbuf[1], h, w);             // statistics of the block output for the next norm
    RC(gemm_launch(c2, ns, r.st));
    r.wsa.reset(m);
    return 0;
}

// dout -> dx (both dense [B,h,w,*]); dx must not alias dout
static int resnet_backward(Run& r, const Resnet& p, const ResnetRec& rec, const bf16* dout, bf16* dx) {
    const int B = r.B, h = rec.h, w = rec.w, hw = h * w;
    const int ns = r.e->num_sms;
    const size_t m = r.wsa.mark();
    float* pbuf = r.Walloc<float>(fused_partial_bytes(B, h, w));
    bf16* d_a2 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    GemmOp g2 = dense_conv_op("resnet.conv2.dgrad", dout, B, h, w, p.co, p.c2.bwd, p.co, 1, h, w, nullptr, nullptr, d_a2);
    const Partials p2 = fuse_gn_bwd(r, g2, r.S<bf16>(rec.h1), p.n2, rec.g2, 1, pbuf, h, w);
    RC(gemm_launch(g2, ns, r.st));
    bf16* d_h1 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    RC(gn_backward(r, r.S<bf16>(rec.h1), d_a2, p.n2, rec.g2, nullptr, d_h1, hw, 1, p2));
    bf16* d_a1 = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
    GemmOp g1 = dense_conv_op("resnet.conv1.dgrad", d_h1, B, h, w, p.co, p.c1.bwd, p.ci, 1, h, w, nullptr, nullptr, d_a1);
    const Partials p1 = fuse_gn_bwd(r, g1, r.S<bf16>(rec.x), p.n1, rec.g1, 1, pbuf, h, w);
    RC(gemm_launch(g1, ns, r.st));
    if (p.has_sc) {
        bf16* tmp = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
        RC(gn_backward(r, r.S<bf16>(rec.x), d_a1, p.n1, rec.g1, nullptr, tmp, hw, 1, p1));
        RC(gemm_launch(dense_lin_op("resnet.shortcut.dgrad", dout, B, h, w, p.co, p.sc.bwd, p.ci, nullptr, tmp, dx), ns, r.st));
    } else {
        RC(gn_backward(r, r.S<bf16>(rec.x), d_a1, p.n1, rec.g1, dout, dx, hw, 1, p1));
    }
    r.wsa.reset(m);
    return 0;
}

static int attn_forward(Run& r, const Attn& p, const AttnRec& rec) {
    const int B = r.B, h = rec.h, w = rec.w, C = p.C, ns = r.e->num_sms;
    const int tok = h * w;
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* t = r.Walloc<bf16>(act_bytes(B, h, w, C));
    RC(gn_forward(r, x, p.gn, rec.g, t, tok, 0, r.pending));
    bf16* qkv = r.S<bf16>(rec.qkv);
    RC(gemm_launch(dense_lin_op("attn.qkv", t, B, h, w, C, p.qkv.fwd, 3 * C, p.qkv.bias, nullptr, qkv), ns, r.st));
    // S = QK^T / sqrt(C)   (fp32, [B][tok][tok])
    float* S = r.Walloc<float>((size_t)B * tok * tok * sizeof(float));
    {
        GemmOp o;
        o.name = "attn.qk";
        o.A = qkv; o.A_C = C; o.A_W = w; o.A_H = h; o.A_B = B;
        o.A_sW = 3 * C; o.A_sH = (int64_t)w * 3 * C; o.A_sB = (int64_t)tok * 3 * C;
        o.OW = w; o.OH = h;
        o.Bm = qkv + C; o.N = tok; o.B_sN = 3 * C; o.B_sBatch = (int64_t)tok * 3 * C;
        o.alpha = 1.0f / sqrtf((float)C);
        o.D = S; o.out_fp32 = 1; o.D_sW = tok; o.D_sH = (int64_t)w * tok; o.D_sB = (int64_t)tok * tok; o.D_sN = 1;
        RC(gemm_launch(o, ns, r.st));
    }
    bf16* P = r.S<bf16>(rec.P);
    launch_softmax_rows(S, P, (long long)B * tok, tok, r.st);
    bf16* Vt = r.Walloc<bf16>(act_bytes(B, h, w, C));  // [B][C][tok]
    launch_transpose(qkv + 2 * C, Vt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* a = r.Walloc<bf16>(act_bytes(B, h, w, C));
    {
        GemmOp o;
        o.name = "attn.pv";
        o.A = P; o.A_C = tok; o.A_W = w; o.A_H = h; o.A_B = B;
        o.A_sW = tok; o.A_sH = (int64_t)w * tok; o.A_sB = (int64_t)tok * tok;
        o.OW = w; o.OH = h;
        o.Bm = Vt; o.N = C; o.B_sN = tok; o.B_sBatch = (int64_t)C * tok;
        o.D = a; o.D_sW = C; o.D_sH = (int64_t)w * C; o.D_sB = (int64_t)tok * C; o.D_sN = 1;
        RC(gemm_launch(o, ns, r.st));
    }
    GemmOp oo = dense_lin_op("attn.out", a, B, h, w, C, p.out.fwd, C, p.out.bias, x, r.S<bf16>(rec.out));
    r.pending = fuse_stats(oo, r.statbuf[1], h, w);
    RC(gemm_launch(oo, ns, r.st));
    r.wsa.reset(m);
    return 0;
}

static int attn_backward(Run& r, const Attn& p, const AttnRec& rec, const bf16* dout, bf16* dx) {
    const int B = r.B, h = rec.h, w = rec.w, C = p.C, ns = r.e->num_sms;
    const int tok = h * w;
    const float scale = 1.0f / sqrtf((float)C);
    const bf16* qkv = r.S<bf16>(rec.qkv);
    const bf16* P = r.S<bf16>(rec.P);
    const size_t m = r.wsa.mark();
    const size_t act = act_bytes(B, h, w, C);
    const size_t tt = (size_t)B * tok * tok;
    bf16* da = r.Walloc<bf16>(act);
    RC(gemm_launch(dense_lin_op("attn.out.dgrad", dout, B, h, w, C, p.out.bwd, C, nullptr, nullptr, da), ns, r.st));
    bf16* daT = r.Walloc<bf16>(act);
    launch_transpose(da, daT, B, tok, C, C, (long long)tok * C, tok, (long long)C * tok, r.st);
    float* dP = r.Walloc<float>(tt * sizeof(float));
    {
        GemmOp o;  // dP = da V^T
        o.name = "attn.dP";
        o.A = da; o.A_C = C; o.A_W = w; o.A_H = h; o.A_B = B;
        o.A_sW = C; o.A_sH = (int64_t)w * C; o.A_sB = (int64_t)tok * C;
        o.OW = w; o.OH = h;
        o.Bm = qkv + 2 * C; o.N = tok; o.B_sN = 3 * C; o.B_sBatch = (int64_t)tok * 3 * C;
        o.D = dP; o.out_fp32 = 1; o.D_sW = tok; o.D_sH = (int64_t)w * tok; o.D_sB = (int64_t)tok * tok; o.D_sN = 1;
        RC(gemm_launch(o, ns, r.st));
    }
    bf16* dS = r.Walloc<bf16>(tt * 2);
    launch_softmax_bwd_rows(P, dP, dS, scale, (long long)B * tok, tok, r.st);
    bf16* dST = r.Walloc<bf16>(tt * 2);
    launch_transpose(dS, dST, B, tok, tok, tok, (long long)tok * tok, tok, (long long)tok * tok, r.st);
    bf16* PT = r.Walloc<bf16>(tt * 2);
    launch_transpose(P, PT, B, tok, tok, tok, (long long)tok * tok, tok, (long long)tok * tok, r.st);
    bf16* Kt = r.Walloc<bf16>(act);
    launch_transpose(qkv + C, Kt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* Qt = r.Walloc<bf16>(act);
    launch_transpose(qkv, Qt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* dqkv = r.Walloc<bf16>(3 * act);
    auto tok_gemm = [&](const char* name, const bf16* A, const bf16* Bt, bf16* D) {
        GemmOp o;  // D[tok, C] (row stride 3C) = A[tok, tok'] * Bt[C, tok']^T
        o.name = name;
        o.A = A; o.A_C = tok; o.A_W = w; o.A_H = h; o.A_B = B;
        o.A_sW = tok; o.A_sH = (int64_t)w * tok; o.A_sB = (int64_t)tok * tok;
        o.OW = w; o.OH = h;
        o.Bm = Bt; o.N = C; o.B_sN = tok; o.B_sBatch = (int64_t)C * tok;
        o.D = D; o.D_sW = 3 * C; o.D_sH = (int64_t)w * 3 * C; o.D_sB = (int64_t)tok * 3 * C; o.D_sN = 1;
        return gemm_launch(o, ns, r.st);
    };
    RC(tok_gemm("attn.dQ", dS, Kt, dqkv));
    RC(tok_gemm("attn.dK", dST, Qt, dqkv + C));
    RC(tok_gemm("attn.dV", PT, daT, dqkv + 2 * C));
    bf16* dt = r.Walloc<bf16>(act);
    GemmOp gq = dense_lin_op("attn.qkv.dgrad", dqkv, B, h, w, 3 * C, p.qkv.bwd, C, nullptr, nullptr, dt);
    float* pbuf = r.Walloc<float>(fused_partial_bytes(B, h, w));
    const Partials pq = fuse_gn_bwd(r, gq, r.S<bf16>(rec.x), p.gn, rec.g, 0, pbuf, h, w);
    RC(gemm_launch(gq, ns, r.st));
    RC(gn_backward(r, r.S<bf16>(rec.x), dt, p.gn, rec.g, dout, dx, tok, 0, pq));
    r.wsa.reset(m);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* tml_last_error(void) { return last_error(); }
int tml_version(void) { return 100; }

int tml_encoder_create(const TmlEncoderCfg* cfg, int device, TmlEncoder** out) {
    if (!cfg || !out) { set_error("null argument"); return -1; }
    if (cfg->in_channels != 3 || cfg->norm_num_groups != 32 || cfg->num_blocks < 1 || cfg->num_blocks > 8 ||
        cfg->block_out_channels[0] != 128 || cfg->latent_channels != 4) {
        set_error("unsupported encoder config (need in=3, groups=32, block_out_channels[0]=128, latent=4)");
        return -1;
    }
    for (int i = 0; i < cfg->num_blocks; ++i)
        if (cfg->block_out_channels[i] % 128) { set_error("block_out_channels must be multiples of 128"); return -1; }
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return -1; }
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("tml_b200 needs an sm_100a device (Blackwell B200); device %d is sm_%d%d — there is no fallback path",
                  device, prop.major, prop.minor);
        return -2;
    }
    TmlEncoder* e = new TmlEncoder();
    e->cfg = *cfg;
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    *out = e;
    return 0;
}

void tml_encoder_destroy(TmlEncoder* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->dev_allocs) cudaFree(p);
    delete e;
}

int tml_encoder_set_weight(TmlEncoder* e, const char* key, const void* ptr, int dtype, const int64_t* shape, int ndim) {
    if (!e || !key || !ptr) { set_error("null argument"); return -1; }
    size_t n = 1;
    HostTensor t;
    for (int i = 0; i < ndim; ++i) { n *= (size_t)shape[i]; t.shape.push_back(shape[i]); }
    const size_t esz = dtype == TML_DTYPE_F32 ? 4 : 2;
    std::vector<unsigned char> raw(n * esz);
    cudaPointerAttributes attr;
    cudaError_t pe = cudaPointerGetAttributes(&attr, ptr);
    if (pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
        CUDA_OK(cudaMemcpy(raw.data(), ptr, n * esz, cudaMemcpyDeviceToHost));
    } else {
        cudaGetLastError();
        memcpy(raw.data(), ptr, n * esz);
    }
    t.v.resize(n);
    if (dtype == TML_DTYPE_F32) memcpy(t.v.data(), raw.data(), n * 4);
    else {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(raw.data());
        for (size_t i = 0; i < n; ++i) {
            if (dtype == TML_DTYPE_BF16) { uint32_t u = (uint32_t)h[i] << 16; memcpy(&t.v[i], &u, 4); }
            else t.v[i] = half2f(h[i]);
        }
    }
    std::string k(key);
    // legacy (pre-0.2x diffusers) attention names
    const char* legacy[][2] = {{".query.", ".to_q."}, {".key.", ".to_k."}, {".value.", ".to_v."}, {".proj_attn.", ".to_out.0."}};
    for (auto& lg : legacy) {
        size_t pos = k.find(lg[0]);
        if (pos != std::string::npos) k.replace(pos, strlen(lg[0]), lg[1]);
    }
    e->host[k] = std::move(t);
    e->finalized = false;
    return 0;
}

int tml_encoder_finalize(TmlEncoder* e, void* stream) {
    (void)stream;
    if (!e) { set_error("null handle"); return -1; }
    CUDA_OK(cudaSetDevice(e->device));
    const TmlEncoderCfg& c = e->cfg;
    const int C0 = c.block_out_channels[0];
    {   // conv_in: [C0][3][3][3] -> [27][C0] fp32
        const HostTensor* w = find(e, "encoder.conv_in.weight", (size_t)C0 * 27);
        const HostTensor* b = find(e, "encoder.conv_in.bias", C0);
        if (!w || !b) return -20;
        std::vector<float> kc((size_t)27 * C0);
        for (int co = 0; co < C0; ++co)
            for (int k = 0; k < 27; ++k) kc[(size_t)k * C0 + co] = w->v[(size_t)co * 27 + k];
        RC(upload<float>(e, kc, &e->conv_in_w));
        RC(upload<float>(e, b->v, &e->conv_in_b));
        // input gradient runs on the tensor cores: N = 3 padded to 16 (UMMA needs N % 16 == 0 at M = 128)
        std::vector<float> w16((size_t)C0 * 16 * 9, 0.f);
        for (int co = 0; co < C0; ++co)
            for (int ci = 0; ci < 3; ++ci)
                for (int k = 0; k < 9; ++k) w16[((size_t)co * 16 + ci) * 9 + k] = w->v[((size_t)co * 3 + ci) * 9 + k];
        RC(make_packed(e, w16.data(), C0, 16, 1, &e->conv_in_bwd));
    }
    e->resnets.clear();
    e->downs.clear();
    int cin = C0;
    char key[256];
    for (int i = 0; i < c.num_blocks; ++i) {
        const int cout = c.block_out_channels[i];
        for (int j = 0; j < c.layers_per_block; ++j) {
            snprintf(key, sizeof(key), "encoder.down_blocks.%d.resnets.%d", i, j);
            Resnet r;
            RC(make_resnet(e, key, cin, cout, &r));
            e->resnets.push_back(r);
            cin = cout;
        }
        if (i != c.num_blocks - 1) {
            snprintf(key, sizeof(key), "encoder.down_blocks.%d.downsamplers.0.conv", i);
            Conv3 d;
            RC(make_conv3(e, key, cout, cout, 2, &d));
            e->downs.push_back(d);
        }
    }
    for (int m = 0; m < 2; ++m) {
        snprintf(key, sizeof(key), "encoder.mid_block.resnets.%d", m);
        Resnet r;
        RC(make_resnet(e, key, cin, cin, &r));
        e->resnets.push_back(r);
    }
    e->has_attn = c.mid_block_add_attention != 0;
    if (e->has_attn) {
        const std::string a = "encoder.mid_block.attentions.0";
        Attn& at = e->attn;
        at.C = cin;
        RC(make_norm(e, a + ".group_norm", cin, &at.gn));
        const char* names[3] = {".to_q", ".to_k", ".to_v"};
        std::vector<float> Wqkv((size_t)3 * cin * cin), bqkv((size_t)3 * cin);
        for (int q = 0; q < 3; ++q) {
            const HostTensor* w = find(e, a + names[q] + ".weight", (size_t)cin * cin);
            const HostTensor* b = find(e, a + names[q] + ".bias", cin);
            if (!w || !b) return -20;
            memcpy(&Wqkv[(size_t)q * cin * cin], w->v.data(), (size_t)cin * cin * 4);
            memcpy(&bqkv[(size_t)q * cin], b->v.data(), (size_t)cin * 4);
        }
        RC(make_lin_from(e, Wqkv, bqkv, cin, 3 * cin, &at.qkv));
        const HostTensor* wo = find(e, a + ".to_out.0.weight", (size_t)cin * cin);
        const HostTensor* bo = find(e, a + ".to_out.0.bias", cin);
        if (!wo || !bo) return -20;
        RC(make_lin_from(e, wo->v, bo->v, cin, cin, &at.out));
    }
    RC(make_norm(e, "encoder.conv_norm_out", cin, &e->norm_out));
    {   // conv_out (cin -> 2L, 3x3) folded with quant_conv (2L -> 2L, 1x1):  W' = Wq * Wout, b' = Wq*b_out + b_q
        const int L2 = 2 * c.latent_channels;  // 8
        const HostTensor* wo = find(e, "encoder.conv_out.weight", (size_t)L2 * cin * 9);
        const HostTensor* bo = find(e, "encoder.conv_out.bias", L2);
        const HostTensor* wq = find(e, "quant_conv.weight", (size_t)L2 * L2);
        const HostTensor* bq = find(e, "quant_conv.bias", L2);
        if (!wo || !bo || !wq || !bq) return -20;
        const int NP = 16;  // UMMA needs N % 16 == 0 at M = 128
        std::vector<float> wf((size_t)NP * cin * 9, 0.f), bf((size_t)NP, 0.f);
        for (int o = 0; o < L2; ++o) {
            double bb = bq->v[o];
            for (int m = 0; m < L2; ++m) bb += (double)wq->v[(size_t)o * L2 + m] * bo->v[m];
            bf[o] = (float)bb;
            for (size_t k = 0; k < (size_t)cin * 9; ++k) {
                double s = 0.0;
                for (int m = 0; m < L2; ++m) s += (double)wq->v[(size_t)o * L2 + m] * wo->v[(size_t)m * cin * 9 + k];
                wf[(size_t)o * cin * 9 + k] = (float)s;
            }
        }
        Conv3& co = e->conv_out;
        co.ci = cin; co.co = NP; co.stride = 1;
        RC(make_packed(e, wf.data(), NP, cin, 0, &co.fwd));
        RC(upload<float>(e, bf, &co.bias));
        // dgrad: A = dmoments padded to 64 channels; B[ci][t*64 + m] = W'[m][ci][r][s]
        std::vector<float> w64((size_t)64 * cin * 9, 0.f);
        memcpy(w64.data(), wf.data(), (size_t)NP * cin * 9 * 4);
        RC(make_packed(e, w64.data(), 64, cin, 1, &co.bwd));
    }
    e->host.clear();
    e->finalized = true;
    e->lay = Layout();
    return 0;
}

int tml_encoder_query(TmlEncoder* e, int B, int H, int W, size_t* workspace_bytes, size_t* saved_bytes) {
    if (!e) { set_error("null handle"); return -1; }
    RC(build_layout(e, B, H, W));
    if (workspace_bytes) *workspace_bytes = e->lay.ws_bytes;
    if (saved_bytes) *saved_bytes = e->lay.saved_bytes;
    return 0;
}

int tml_encoder_forward(TmlEncoder* e, const float* x, int B, int H, int W, float* moments, void* saved, void* ws,
                        void* stream) {
    if (!e || !e->finalized) { set_error("encoder not finalized"); return -1; }
    if (!x || !moments || !saved || !ws) { set_error("null buffer"); return -1; }
    RC(build_layout(e, B, H, W));
    const Layout& L = e->lay;
    Run r{e, reinterpret_cast<char*>(saved), reinterpret_cast<char*>(ws), Arena(), reinterpret_cast<cudaStream_t>(stream), B};
    const int C0 = e->cfg.block_out_channels[0];
    r.statbuf[0] = r.Walloc<float>(fused_partial_bytes(B, H, W));
    r.statbuf[1] = r.Walloc<float>(fused_partial_bytes(B, H, W));
    launch_conv_in_fwd(x, e->conv_in_w, e->conv_in_b, r.S<bf16>(L.x0), B, H, W, C0, r.st);
    size_t ri = 0, di = 0;
    for (int i = 0; i < e->cfg.num_blocks; ++i) {
        for (int j = 0; j < e->cfg.layers_per_block; ++j, ++ri) RC(resnet_forward(r, e->resnets[ri], L.res[ri]));
        if (i != e->cfg.num_blocks - 1) {
            const DownRec& d = L.down[di];
            const Conv3& c = e->downs[di];
            GemmOp dn = dense_conv_op("downsample", r.S<bf16>(d.x), B, d.h, d.w, c.ci, c.fwd, c.co, 2, d.h / 2, d.w / 2,
                                      c.bias, nullptr, r.S<bf16>(d.out));
            r.pending = fuse_stats(dn, r.statbuf[1], d.h / 2, d.w / 2);
            RC(gemm_launch(dn, e->num_sms, r.st));
            ++di;
        }
    }
    RC(resnet_forward(r, e->resnets[ri], L.res[ri])); ++ri;
    if (e->has_attn) RC(attn_forward(r, e->attn, L.attn));
    RC(resnet_forward(r, e->resnets[ri], L.res[ri])); ++ri;
    {   // conv_norm_out + SiLU + (conv_out o quant_conv) -> fp32 NCHW moments
        const int h = L.hl, w = L.wl, C = e->norm_out.C, L2 = 2 * e->cfg.latent_channels;
        bf16* a = r.Walloc<bf16>(act_bytes(B, h, w, C));
        RC(gn_forward(r, r.S<bf16>(L.xlast), e->norm_out, L.gout, a, h * w, 1, r.pending));
        GemmOp o = dense_conv_op("conv_out", a, B, h, w, C, e->conv_out.fwd, 16, 1, h, w, e->conv_out.bias, nullptr, nullptr);
        o.D = moments; o.out_fp32 = 1; o.n_store = L2;
        o.D_sB = (int64_t)L2 * h * w; o.D_sH = w; o.D_sW = 1; o.D_sN = (int64_t)h * w;
        RC(gemm_launch(o, e->num_sms, r.st));
    }
    if (r.wsa.peak > L.ws_bytes) { set_error("internal: workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_encoder_backward(TmlEncoder* e, const float* dmoments, int B, int H, int W, const void* saved, float* dx,
                         float beta, void* ws, void* stream) {
    if (!e || !e->finalized) { set_error("encoder not finalized"); return -1; }
    if (!dmoments || !saved || !ws || !dx) { set_error("null buffer"); return -1; }
    RC(build_layout(e, B, H, W));
    const Layout& L = e->lay;
    Run r{e, const_cast<char*>(reinterpret_cast<const char*>(saved)), reinterpret_cast<char*>(ws), Arena(),
          reinterpret_cast<cudaStream_t>(stream), B};
    // ping-pong gradient buffers
    size_t gmax = 0;
    {
        int hh = H, ww = W;
        for (int i = 0; i < e->cfg.num_blocks; ++i) {
            gmax = std::max(gmax, act_bytes(B, hh, ww, e->cfg.block_out_channels[i]));
            if (i) gmax = std::max(gmax, act_bytes(B, hh, ww, e->cfg.block_out_channels[i - 1]));
            if (i != e->cfg.num_blocks - 1) { hh /= 2; ww /= 2; }
        }
    }
    bf16* G[2] = {r.Walloc<bf16>(gmax), r.Walloc<bf16>(gmax)};
    int cur = 0;
    g_dump_next = 0;
    {   // d(moments) -> d(conv_norm_out input)
        const int h = L.hl, w = L.wl, C = e->norm_out.C;
        const size_t m = r.wsa.mark();
        bf16* dm64 = r.Walloc<bf16>(act_bytes(B, h, w, 64));
        launch_dmoments_pack(dmoments, dm64, B, h, w, r.st);
        bf16* d_a = r.Walloc<bf16>(act_bytes(B, h, w, C));
        GemmOp go = dense_conv_op("conv_out.dgrad", dm64, B, h, w, 64, e->conv_out.bwd, C, 1, h, w, nullptr, nullptr, d_a);
        float* pbuf = r.Walloc<float>(fused_partial_bytes(B, h, w));
        const Partials po = fuse_gn_bwd(r, go, r.S<bf16>(L.xlast), e->norm_out, L.gout, 1, pbuf, h, w);
        RC(gemm_launch(go, e->num_sms, r.st));
        RC(gn_backward(r, r.S<bf16>(L.xlast), d_a, e->norm_out, L.gout, nullptr, G[cur], h * w, 1, po));
        r.wsa.reset(m);
        dump_grad(G[cur], act_bytes(B, h, w, C), r.st);
    }
    size_t ri = e->resnets.size();
    const size_t mid_bytes = act_bytes(B, L.hl, L.wl, e->norm_out.C);
    RC(resnet_backward(r, e->resnets[ri - 1], L.res[ri - 1], G[cur], G[cur ^ 1])); cur ^= 1; --ri;
    dump_grad(G[cur], mid_bytes, r.st);
    if (e->has_attn) { RC(attn_backward(r, e->attn, L.attn, G[cur], G[cur ^ 1])); cur ^= 1; dump_grad(G[cur], mid_bytes, r.st); }
    RC(resnet_backward(r, e->resnets[ri - 1], L.res[ri - 1], G[cur], G[cur ^ 1])); cur ^= 1; --ri;
    dump_grad(G[cur], mid_bytes, r.st);
    size_t di = e->downs.size();
    for (int i = e->cfg.num_blocks - 1; i >= 0; --i) {
        if (i != e->cfg.num_blocks - 1) {
            --di;
            const DownRec& d = L.down[di];
            const Conv3& c = e->downs[di];
            const int oh = d.h / 2, ow = d.w / 2;
            for (int q = 0; q < 4; ++q) {
                const int ph = q >> 1, pw = q & 1;
                GemmOp o = dense_conv_op("downsample.dgrad", G[cur], B, oh, ow, c.co, c.bwd_par[q], c.ci, 1, oh, ow, nullptr,
                                         nullptr, G[cur ^ 1] + ((size_t)ph * d.w + pw) * c.ci);
                o.D_sW = 2 * c.ci; o.D_sH = (int64_t)2 * d.w * c.ci; o.D_sB = (int64_t)d.h * d.w * c.ci;
                RC(gemm_launch(o, e->num_sms, r.st));
            }
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, d.h, d.w, c.ci), r.st);
        }
        for (int j = e->cfg.layers_per_block - 1; j >= 0; --j) {
            --ri;
            RC(resnet_backward(r, e->resnets[ri], L.res[ri], G[cur], G[cur ^ 1]));
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, L.res[ri].h, L.res[ri].w, e->resnets[ri].ci), r.st);
        }
    }
    {   // conv_in input gradient -> fp32 NCHW image gradient (the tensor PGD consumes, main.py:176);
        // beta = 1 accumulates the grad_reps of main.py:88-102 in place.
        const int C0 = e->cfg.block_out_channels[0];
        GemmOp o = dense_conv_op("conv_in.dgrad", G[cur], B, H, W, C0, e->conv_in_bwd, 16, 1, H, W, nullptr, nullptr, nullptr);
        o.D = dx; o.out_fp32 = 1; o.n_store = 3; o.beta = beta;
        o.D_sB = (int64_t)3 * H * W; o.D_sH = W; o.D_sW = 1; o.D_sN = (int64_t)H * W;
        RC(gemm_launch(o, e->num_sms, r.st));
    }
    if (r.wsa.peak > L.ws_bytes) { set_error("internal: workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_latent_loss(int kind, const float* moments, const float* noise, const float* target, int B, int h, int w,
                    float grad_scale, float* z_out, float* loss, float* dmoments, void* stream) {
    if (!moments || !target) { set_error("null buffer"); return -1; }
    if (kind != TML_LOSS_L2NORM && kind != TML_LOSS_MSE) { set_error("unknown loss kind %d", kind); return -1; }
    launch_latent_loss(kind, moments, noise, target, B, h, w, grad_scale, z_out, loss, dmoments,
                       reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_pgd_step_linf(float* x_adv, const float* grad, const float* x, float eps, float step, float lo, float hi,
                      int64_t n, void* stream) {
    if (n <= 0) return 0;
    if (!x_adv || !grad || !x) { set_error("null buffer"); return -1; }
    if (((uintptr_t)x_adv | (uintptr_t)grad | (uintptr_t)x) & 15) { set_error("buffers must be 16-byte aligned"); return -1; }
    if (n <= 0) return 0;
    launch_pgd_linf(x_adv, grad, x, eps, step, lo, hi, n, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

size_t tml_pgd_l2_workspace(int B) { return pgd_l2_workspace_bytes(B, 0); }

int tml_pgd_step_l2(float* x_adv, const float* grad, const float* x, const float* mask, float eps, float step, float lo,
                    float hi, int B, int C, int64_t hw, void* ws, void* stream) {
    if (!x_adv || !grad || !x || !ws) { set_error("null buffer"); return -1; }
    if (B <= 0) return 0;
    launch_pgd_l2(x_adv, grad, x, mask, eps, step, lo, hi, B, C, hw, ws, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_add_delta(const float* x, const float* delta, float* out, int B, int64_t per_image, void* stream) {
    if (!x || !delta || !out) { set_error("null buffer"); return -1; }
    launch_add_delta(x, delta, out, B, per_image, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}
int tml_batch_sum(const float* g, float* out, int B, int64_t per_image, float scale, void* stream) {
    if (!g || !out) { set_error("null buffer"); return -1; }
    launch_batch_sum(g, out, B, per_image, scale, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}
int tml_universal_step(float* delta, const float* grad, const float* source, float eps, float step, float lo, float hi,
                       int64_t n, void* ws, void* stream) {
    if (!delta || !grad || !ws) { set_error("null buffer"); return -1; }
    launch_universal_step(delta, grad, source, eps, step, lo, hi, n, ws, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

void tml_launch_counts(int64_t out[2]) {
    out[0] = gemm_launch_count();
    out[1] = kernel_launch_count();
}
void tml_debug_set_gemm_impl(int impl) { gemm_set_impl(impl); }

int tml_debug_gemm(const TmlGemmDesc* d, void* stream) {
    if (!d) { set_error("null desc"); return -1; }
    GemmOp o;
    o.name = "debug_gemm";
    o.A = d->A; o.A_C = d->A_C; o.A_W = d->A_W; o.A_H = d->A_H; o.A_B = d->A_B;
    o.A_sW = d->A_sW; o.A_sH = d->A_sH; o.A_sB = d->A_sB;
    o.stride = d->stride; o.ntaps = d->ntaps;
    if (d->ntaps < 1 || d->ntaps > kMaxTaps) { set_error("ntaps"); return -1; }
    for (int t = 0; t < d->ntaps; ++t) { o.dh[t] = d->dh[t]; o.dw[t] = d->dw[t]; }
    o.OW = d->OW; o.OH = d->OH;
    o.Bm = d->Bm; o.N = d->N; o.B_sN = d->B_sN; o.B_sBatch = d->B_sBatch;
    o.alpha = d->alpha; o.bias = d->bias; o.resid = d->resid;
    o.R_sB = d->R_sB; o.R_sH = d->R_sH; o.R_sW = d->R_sW;
    o.D = d->D; o.out_fp32 = d->out_fp32;
    o.D_sB = d->D_sB; o.D_sH = d->D_sH; o.D_sW = d->D_sW; o.D_sN = d->D_sN; o.n_store = d->n_store;
    o.beta = d->beta;
    o.gn_mode = d->gn_mode; o.gn_partial = d->gn_partial; o.gn_x = d->gn_x;
    o.gn_ss = reinterpret_cast<const float2*>(d->gn_ss); o.gn_mr = reinterpret_cast<const float2*>(d->gn_mr);
    o.gn_gamma = d->gn_gamma; o.gn_silu = d->gn_silu;
    o.dbg_shift = d->dbg_shift; o.dbg_bo = d->dbg_bo;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return gemm_launch(o, sms, reinterpret_cast<cudaStream_t>(stream));
}

void tml_gemm_timing_enable(int max_launches) { gemm_timing_enable(max_launches); }
void tml_gemm_timing_collect(double out[4]) {
    long n = 0, d = 0;
    gemm_timing_collect(&out[0], &out[1], &n, &d);
    out[2] = (double)n; out[3] = (double)d;
}

size_t tml_gemm_timing_report(char* buf, size_t cap) { return gemm_timing_report(buf, cap); }
int tml_debug_last_hang(void) { return gemm_last_hang(); }

void tml_debug_set_grad_dump(void* dev_buffer, size_t slot_bytes, int slots) {
    g_dump_base = reinterpret_cast<char*>(dev_buffer);
    g_dump_slot = slot_bytes;
    g_dump_slots = slots;
    g_dump_next = 0;
}

int tml_debug_saved_tensor(TmlEncoder* e, const char* name, int index, size_t* offset, int dims[4]) {
    if (!e || !name || !offset || !dims) { set_error("null argument"); return -1; }
    const Layout& L = e->lay;
    if (!L.saved_bytes) { set_error("no layout yet: call tml_encoder_query/forward first"); return -1; }
    const std::string n(name);
    const int C0 = e->cfg.block_out_channels[0];
    auto set = [&](size_t off, int h, int w, int c) { *offset = off; dims[0] = L.B; dims[1] = h; dims[2] = w; dims[3] = c; return 0; };
    if (n == "conv_in") return set(L.x0, L.H, L.W, C0);
    if (n == "resnet_h1" || n == "resnet_out") {
        if (index < 0 || index >= (int)L.res.size()) { set_error("index"); return -1; }
        const ResnetRec& r = L.res[index];
        return set(n == "resnet_h1" ? r.h1 : r.out, r.h, r.w, e->resnets[index].co);
    }
    if (n == "down_out") {
        if (index < 0 || index >= (int)L.down.size()) { set_error("index"); return -1; }
        return set(L.down[index].out, L.down[index].h / 2, L.down[index].w / 2, e->downs[index].co);
    }
    if (n == "attn_qkv") return set(L.attn.qkv, L.attn.h, L.attn.w, 3 * e->attn.C);
    if (n == "attn_out") return set(L.attn.out, L.attn.h, L.attn.w, e->attn.C);
    if (n == "attn_P") { *offset = L.attn.P; dims[0] = L.B; dims[1] = L.attn.h * L.attn.w; dims[2] = L.attn.h * L.attn.w; dims[3] = 1; return 0; }
    set_error("unknown tensor '%s'", name);
    return -1;
}

int tml_debug_gn_tiles_per_image(int OH, int OW) { return gemm_gn_tiles_per_image(OH, OW); }

int tml_debug_pack_conv3x3(const float* w, int Co, int Ci, int mode, uint16_t* out, int* ntaps, int* dh, int* dw) {
    std::vector<uint16_t> h;
    if (pack_conv3x3(w, Co, Ci, mode, h, ntaps, dh, dw)) { set_error("bad pack mode %d", mode); return -1; }
    memcpy(out, h.data(), h.size() * 2);
    return 0;
}

}  // extern "C"
