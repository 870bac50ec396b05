// Shared implementation pieces of the VAE encoder and decoder walks: host-side weight packing, the layer
// parameter / saved-state records, the scratch arena, and the ResnetBlock2D / GroupNorm / attention
// forward and input-gradient sequences that both networks are built from.  Included by encoder.cu and
// decoder.cu (everything here has internal linkage).
#pragma once
// (decoder.cu) builds the decoder parameters from the registered host tensors; called by tml_encoder_finalize
struct TmlEncoder;
int decoder_finalize(TmlEncoder* e);
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

// Makes the handle's device current for the duration of a C-ABI call and restores the caller's device afterwards
// (the caller may be a framework that tracks the current device itself).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

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
    // mode 7: forward s2, symmetric pad 1 (the UNet's Downsample2D): same matrix as mode 0, (dh,dw) = (r-1, s-1)
    // mode 8..11: its dgrad per output parity (ph,pw): input row i = 2o + r - 1, so r = ph + 1 (mod 2) and
    //             dY row o = i' + (ph + 1 - r)/2 for i = 2i' + ph
    if (mode == 7) {
        const int rc = pack_conv3x3(w, Co, Ci, 0, out, ntaps, dh, dw);
        return rc;
    }
    if ((mode >= 3 && mode <= 6) || (mode >= 8 && mode <= 11)) {
        const int pad = mode >= 8 ? 1 : 0;
        const int q = mode >= 8 ? mode - 8 : mode - 3;
        const int ph = q >> 1, pw = q & 1;
        int rs[2], ss[2], nr = 0, ns = 0;
        for (int r = 0; r < 3; ++r) if ((r & 1) == ((ph + pad) & 1)) rs[nr++] = r;
        for (int s = 0; s < 3; ++s) if ((s & 1) == ((pw + pad) & 1)) ss[ns++] = s;
        *ntaps = nr * ns;
        out.assign((size_t)Ci * (*ntaps) * Co, 0);
        int t = 0;
        for (int a = 0; a < nr; ++a)
            for (int b = 0; b < ns; ++b, ++t) {
                const int r = rs[a], s = ss[b];
                dh[t] = (ph + pad - r) / 2;  // pad 0: 0 or -1; pad 1: 0 or +1 (exact divisions)
                dw[t] = (pw + pad - s) / 2;
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
// P: unnormalised softmax numerators exp(s - rowmax) (bf16 [B][tok][tok]); inv_l: 1 / row sums (fp32 [B][tok]);
// a: attention output before the projection (bf16 [B][tok][C])
struct AttnRec { size_t x = 0, qkv = 0, P = 0, a = 0, inv_l = 0, out = 0; GnSaved g; int h = 0, w = 0; };

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
    bool measured = false;   // ws_bytes has been replaced by the peak of a dry run of both walks
    size_t x0 = 0;  // conv_in output
    std::vector<ResnetRec> res;
    std::vector<DownRec> down;
    AttnRec attn;
    GnSaved gout;
    size_t xlast = 0;
    int hl = 0, wl = 0;
};

struct UpRec { size_t out = 0; int h = 0, w = 0; };   // Upsample2D: (h, w) = input size, out = conv output at (2h, 2w)
struct DecLayout {
    int B = 0, h = 0, w = 0;             // latent size
    size_t saved_bytes = 0, ws_bytes = 0;
    bool measured = false;               // as in Layout
    size_t x0 = 0;                       // decoder.conv_in output
    std::vector<ResnetRec> res;
    AttnRec attn;
    std::vector<UpRec> ups;
    GnSaved gout;
    size_t xlast = 0;
    int Hl = 0, Wl = 0;                  // output image size
};

struct TmlEncoder {
    TmlEncoderCfg cfg;
    int device = 0;
    int num_sms = 148;
    bool finalized = false;
    std::map<std::string, HostTensor> host;
    std::vector<void*> dev_allocs;
    // parameters
    Conv3 conv_in_fwd;           // forward as a 3x3 conv over 64 channels [hi(3) | lo(3) | 0]: weights on c and c+3
    Lin conv_in_bwd;             // input gradient, step 1: [32][C0], row (r*3+s)*3+ci = W[:, ci, r, s] (27 used)
    std::vector<Resnet> resnets;          // in forward order (down blocks then mid[0], mid[1])
    std::vector<Conv3> downs;
    bool has_attn = false;
    Attn attn;
    Norm norm_out;
    Conv3 conv_out;  // folded with quant_conv, N padded to 16; bwd has K = 64 (8 real)
    Layout lay;
    // ---- decoder (optional: present when "decoder.*" / "post_quant_conv.*" weights were registered) ----
    bool has_decoder = false;
    float* pq_w = nullptr;                // post_quant_conv [4][4] fp32
    float* pq_b = nullptr;
    Conv3 dec_conv_in;                    // 4 (padded to 64) -> C: fwd [C][9*64], bwd [64][9*C]
    std::vector<Resnet> dec_resnets;      // mid[0], mid[1], then the up blocks' resnets in forward order
    bool dec_has_attn = false;
    Attn dec_attn;
    std::vector<Conv3> ups;               // Upsample2D convs
    Norm dec_norm_out;
    Conv3 dec_conv_out;                   // C0 -> 3: fwd N padded to 16; bwd A = d(image) padded to 64 channels
    DecLayout dlay;
};

// ------------------------------------------------------------------------------------------------
// weight upload
// ------------------------------------------------------------------------------------------------
// tests without a GPU (tml_debug_set_host_only): parameters keep their shapes but nothing is uploaded, so the walks
// can be replayed as dry runs (layout, scratch size, shape validation of every GEMM) on a CPU-only machine
inline bool g_host_only = false;
template <typename T>
static int upload(TmlEncoder* e, const std::vector<T>& h, T** out) {
    if (g_host_only) { *out = nullptr; return 0; }
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
// partial sums written by a GEMM epilogue: one entry per (image, 128-row tile or 64-pixel segment, group)
static size_t fused_partial_bytes(int B, int h, int w) { return (size_t)B * 2 * gemm_gn_tiles_per_image(h, w) * 32 * 2 * sizeof(float); }
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
    Arena S;   // (the scratch size is measured by a dry run of the walks: enc_layout)
    int h = H, w = W;
    L.x0 = S.alloc(act_bytes(B, h, w, c.block_out_channels[0]));
    size_t cur = L.x0;
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
        L.res.push_back(r);
        cur = r.out;
        if (m == 0 && c.mid_block_add_attention) {
            AttnRec a;
            a.h = h; a.w = w; a.x = cur;
            const size_t tok = (size_t)h * w;
            a.g = alloc_gn(S, B, cin);
            a.qkv = S.alloc((size_t)B * tok * 3 * cin * sizeof(bf16));
            a.P = S.alloc((size_t)B * tok * tok * sizeof(bf16));
            a.a = S.alloc(act_bytes(B, h, w, cin));
            a.inv_l = S.alloc((size_t)B * tok * sizeof(float));
            a.out = S.alloc(act_bytes(B, h, w, cin));
            // (scratch of the attention walks is measured by the dry run, see enc_layout)
            L.attn = a;
            cur = a.out;
        }
    }
    L.gout = alloc_gn(S, B, cin);
    L.xlast = cur;
    L.hl = h; L.wl = w;
    L.saved_bytes = S.peak + 256;
    L.ws_bytes = 0;   // set by enc_layout from the dry run
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
inline char* g_dump_base = nullptr;
inline size_t g_dump_slot = 0;
inline int g_dump_slots = 0, g_dump_next = 0;
inline void dump_grad(const void* p, size_t bytes, cudaStream_t st) {
    if (g_dry_run || !g_dump_base || g_dump_next >= g_dump_slots) return;
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
    // y == nullptr: statistics only -- the consumer applies scale / shift / SiLU on its operand path (GemmOp::in_gn_ss)
    if (y != nullptr) launch_gn_apply(x, r.S<float2>(g.ss), y, r.B, hw, n.C, silu, r.st);
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
    p.nchunks = gemm_gn_chunks_per_image(o);
    return p;
}
// Ask a dgrad GEMM to also reduce the GroupNorm-backward sums of the norm whose output it differentiates.
static Partials fuse_gn_bwd(Run& r, GemmOp& o, const bf16* x, const Norm& n, const GnSaved& g, int silu, float* buf,
                            int oh, int ow) {
    static const bool off = env_off("TML_NO_FUSE_GNBWD");     // tuning switch
    // Measured on B200: in the pixel-major kernel a K = 9*128 main loop is too short to hide this epilogue (exp + rcp
    // per element, an extra read of x, cross-lane sums), so short-K layers keep the standalone reduction kernel there.
    // The operand-swapped kernel (thread = channel, 16 epilogue warps) hides it at any K: 128-channel dgrads go
    // 7.1 -> 8.0 ms per 64 images and the 2.2 ms/16-image reduction kernel disappears.
    static const int min_k = getenv("TML_GNBWD_MIN_K") ? atoi(getenv("TML_GNBWD_MIN_K")) : 2000;
    if (off || gemm_get_impl() != 0 || (o.ntaps * o.A_C < min_k && !gemm_swapped_shape(o))) return Partials();
    o.gn_mode = 2;
    o.gn_partial = buf;
    o.gn_x = x;
    o.gn_ss = r.S<float2>(g.ss);
    o.gn_mr = r.S<float2>(g.mr);
    o.gn_gamma = n.gamma;
    o.gn_silu = silu;
    Partials p;
    p.p = buf;
    p.nchunks = gemm_gn_chunks_per_image(o);
    return p;
}

static int resnet_forward(Run& r, const Resnet& p, const ResnetRec& rec) {
    const int B = r.B, h = rec.h, w = rec.w, hw = h * w;
    const int ns = r.e->num_sms;
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* h1 = r.S<bf16>(rec.h1);
    // Where the convolution runs on the CTA-pair form of the operand-swapped kernel (the 256^2 and 128^2 stages), the
    // GroupNorm + SiLU of its input CAN be applied on the operand path (the normalised activation is then never written);
    // measured slower than the separate apply pass, so gemm_fuses_input_gn is off unless TML_FUSE_INGN=1.
    GemmOp c1 = dense_conv_op("resnet.conv1", x, B, h, w, p.ci, p.c1.fwd, p.co, 1, h, w, p.c1.bias, nullptr, h1);
    if (gemm_fuses_input_gn(c1)) {
        RC(gn_forward(r, x, p.n1, rec.g1, nullptr, hw, 1, r.pending));
        c1.in_gn_ss = r.S<float2>(rec.g1.ss);
    } else {
        bf16* a = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
        RC(gn_forward(r, x, p.n1, rec.g1, a, hw, 1, r.pending));
        c1.A = a;
    }
    const Partials s1 = fuse_stats(c1, r.statbuf[0], h, w);     // statistics of h1 for norm2, from the epilogue
    RC(gemm_launch(c1, ns, r.st));
    const bf16* resid = x;
    bf16* sc = nullptr;
    if (p.has_sc) sc = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    GemmOp c2 = dense_conv_op("resnet.conv2", h1, B, h, w, p.co, p.c2.fwd, p.co, 1, h, w, p.c2.bias, resid, r.S<bf16>(rec.out));
    if (gemm_fuses_input_gn(c2)) {
        RC(gn_forward(r, h1, p.n2, rec.g2, nullptr, hw, 1, s1));
        c2.in_gn_ss = r.S<float2>(rec.g2.ss);
    } else {
        bf16* a2 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
        RC(gn_forward(r, h1, p.n2, rec.g2, a2, hw, 1, s1));
        c2.A = a2;
    }
    if (p.has_sc) {
        RC(gemm_launch(dense_lin_op("resnet.shortcut", x, B, h, w, p.ci, p.sc.fwd, p.co, p.sc.bias, nullptr, sc), ns, r.st));
        c2.resid = sc;
    }
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

// Attention (1 head, d = C) without fp32 logits in memory: the softmax lives in the epilogues of the GEMMs around it.
//   pass 1  QK^T, epilogue keeps only the row maxima            (GemmOp::epi_mode 1)
//   pass 2  QK^T again, epilogue stores P~ = exp(scale*(s - max)) as bf16 and the row sums of what it stored (mode 2)
//   PV      rows scaled by 1 / sum in the epilogue               (row_scale)
// The second QK^T costs 2*tok^2*C flops per image (0.7 % of the step at 512^2) and replaces a 4-byte-per-logit write,
// the softmax kernel's 4-byte read and its 2-byte write.
// The attention products see tokens, not pixels: they are tiled over a virtual [tok / gw][gw] grid (gw = 128 or 64) so that
// every tile is 128 consecutive tokens and every epilogue warp owns 32 of them, whatever the latent's height and width.
static int attn_grid(int tok, int* gh, int* gw) {
    if (tok % 64 != 0) { set_error("attention needs (H/8)*(W/8) %% 64 == 0 tokens, got %d", tok); return -31; }
    *gw = tok % 128 == 0 ? 128 : 64;
    *gh = tok / *gw;
    return 0;
}

static GemmOp attn_logits_op(const char* name, const bf16* A, const bf16* Bk, int B, int h, int w, int C, int ldA, int ldB) {
    const int tok = h * w;
    GemmOp o;   // D[tok, tok'] = A[tok, C] * Bk[tok', C]^T, both operands rows of a [B][tok][ld] tensor
    o.name = name;
    o.A = A; o.A_C = C; o.A_W = w; o.A_H = h; o.A_B = B;
    o.A_sW = ldA; o.A_sH = (int64_t)w * ldA; o.A_sB = (int64_t)tok * ldA;
    o.OW = w; o.OH = h;
    o.Bm = Bk; o.N = tok; o.B_sN = ldB; o.B_sBatch = (int64_t)tok * ldB;
    o.D_sW = tok; o.D_sH = (int64_t)w * tok; o.D_sB = (int64_t)tok * tok; o.D_sN = 1;
    return o;
}

static int attn_forward(Run& r, const Attn& p, const AttnRec& rec) {
    const int B = r.B, h = rec.h, w = rec.w, C = p.C, ns = r.e->num_sms;
    const int tok = h * w;
    const long long rows = (long long)B * tok;
    const float scale = 1.0f / sqrtf((float)C);
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* t = r.Walloc<bf16>(act_bytes(B, h, w, C));
    RC(gn_forward(r, x, p.gn, rec.g, t, tok, 0, r.pending));
    bf16* qkv = r.S<bf16>(rec.qkv);
    RC(gemm_launch(dense_lin_op("attn.qkv", t, B, h, w, C, p.qkv.fwd, 3 * C, p.qkv.bias, nullptr, qkv), ns, r.st));
    bf16* P = r.S<bf16>(rec.P);
    float* inv_l = r.S<float>(rec.inv_l);
    int gh = 0, gw = 0;
    RC(attn_grid(tok, &gh, &gw));
    {
        GemmOp o = attn_logits_op("attn.qk.max", qkv, qkv + C, B, gh, gw, C, 3 * C, 3 * C);
        o.epi_mode = 1;
        const int np = gemm_row_partials(o);
        if (np < 0) return np;
        float* part = r.Walloc<float>((size_t)rows * np * sizeof(float));
        float* rmax = r.Walloc<float>((size_t)rows * sizeof(float));
        o.row_part = part;
        RC(gemm_launch(o, ns, r.st));
        launch_row_reduce(part, rmax, rows, np, 0, r.st);
        GemmOp e = attn_logits_op("attn.qk.exp", qkv, qkv + C, B, gh, gw, C, 3 * C, 3 * C);
        e.epi_mode = 2;
        e.row_a = rmax; e.row_part = part;
        e.exp_scale = scale * 1.4426950408889634f;   // exp(scale * (s - max)) = exp2((s - max) * scale * log2 e)
        e.D = P;
        RC(gemm_launch(e, ns, r.st));
        launch_row_reduce(part, inv_l, rows, np, 1, r.st);
    }
    bf16* Vt = r.Walloc<bf16>(act_bytes(B, h, w, C));  // [B][C][tok]
    launch_transpose(qkv + 2 * C, Vt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* a = r.S<bf16>(rec.a);
    {
        GemmOp o;
        o.name = "attn.pv";
        o.A = P; o.A_C = tok; o.A_W = gw; o.A_H = gh; o.A_B = B;
        o.A_sW = tok; o.A_sH = (int64_t)gw * tok; o.A_sB = (int64_t)tok * tok;
        o.OW = gw; o.OH = gh;
        o.Bm = Vt; o.N = C; o.B_sN = tok; o.B_sBatch = (int64_t)C * tok;
        o.row_scale = inv_l;
        o.D = a; o.D_sW = C; o.D_sH = (int64_t)gw * C; o.D_sB = (int64_t)tok * C; o.D_sN = 1;
        RC(gemm_launch(o, ns, r.st));
    }
    GemmOp oo = dense_lin_op("attn.out", a, B, h, w, C, p.out.fwd, C, p.out.bias, x, r.S<bf16>(rec.out));
    r.pending = fuse_stats(oo, r.statbuf[1], h, w);
    RC(gemm_launch(oo, ns, r.st));
    r.wsa.reset(m);
    return 0;
}

// Backward: dP = da V^T never leaves the tensor memory -- the epilogue of that GEMM turns it into
// dS = P * (dP - D) * scale with P = P~ / l and D = rowsum(dP * P) = da . a (mode 3); dV = P^T da uses the
// unnormalised P~ with the rows of da pre-scaled by 1 / l.
static int attn_backward(Run& r, const Attn& p, const AttnRec& rec, const bf16* dout, bf16* dx) {
    const int B = r.B, h = rec.h, w = rec.w, C = p.C, ns = r.e->num_sms;
    const int tok = h * w;
    const long long rows = (long long)B * tok;
    const float scale = 1.0f / sqrtf((float)C);
    const bf16* qkv = r.S<bf16>(rec.qkv);
    const bf16* P = r.S<bf16>(rec.P);
    const float* inv_l = r.S<float>(rec.inv_l);
    const size_t m = r.wsa.mark();
    const size_t act = act_bytes(B, h, w, C);
    const size_t tt = (size_t)B * tok * tok;
    bf16* da = r.Walloc<bf16>(act);
    RC(gemm_launch(dense_lin_op("attn.out.dgrad", dout, B, h, w, C, p.out.bwd, C, nullptr, nullptr, da), ns, r.st));
    float* Drow = r.Walloc<float>((size_t)rows * sizeof(float));
    launch_row_dot(da, r.S<bf16>(rec.a), Drow, rows, C, r.st);
    bf16* daT = r.Walloc<bf16>(act);   // (da / l)^T
    launch_transpose(da, daT, B, tok, C, C, (long long)tok * C, tok, (long long)C * tok, r.st, inv_l);
    int gh = 0, gw = 0;
    RC(attn_grid(tok, &gh, &gw));
    bf16* dS = r.Walloc<bf16>(tt * 2);
    {
        GemmOp o = attn_logits_op("attn.dS", da, qkv + 2 * C, B, gh, gw, C, C, 3 * C);   // dP = da V^T
        o.epi_mode = 3;
        o.alpha = scale;
        o.row_a = Drow; o.row_b = inv_l;
        o.resid = P; o.R_sW = tok; o.R_sH = (int64_t)gw * tok; o.R_sB = (int64_t)tok * tok;
        o.D = dS;
        RC(gemm_launch(o, ns, r.st));
    }
    bf16* Kt = r.Walloc<bf16>(act);
    launch_transpose(qkv + C, Kt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* Qt = r.Walloc<bf16>(act);
    launch_transpose(qkv, Qt, B, tok, C, 3 * C, (long long)tok * 3 * C, tok, (long long)C * tok, r.st);
    bf16* dqkv = r.Walloc<bf16>(3 * act);
    // D[tok, C] (row stride 3C) = A[tok, tok'] * Bt[C, tok']^T; transposed = true reads A[tok', tok] (the stored dS / P~)
    // as an MN-major operand instead (GemmOp::a_trans): dK = dS^T Q and dV = P~^T (da / l) without a transposed copy
    auto tok_gemm = [&](const char* name, const bf16* A, const bf16* Bt, bf16* D, bool transposed) {
        GemmOp o;
        o.name = name;
        o.A = A; o.A_C = tok; o.A_W = gw; o.A_H = gh; o.A_B = B;
        o.A_sW = tok; o.A_sH = (int64_t)gw * tok; o.A_sB = (int64_t)tok * tok;
        if (transposed) { o.a_trans = 1; o.A_sK = tok; }
        o.OW = gw; o.OH = gh;
        o.Bm = Bt; o.N = C; o.B_sN = tok; o.B_sBatch = (int64_t)C * tok;
        o.D = D; o.D_sW = 3 * C; o.D_sH = (int64_t)gw * 3 * C; o.D_sB = (int64_t)tok * 3 * C; o.D_sN = 1;
        return gemm_launch(o, ns, r.st);
    };
    static const bool no_atrans = env_off("TML_NO_ATRANS");   // A/B switch: transposed copies of dS and P~ as in round 1
    RC(tok_gemm("attn.dQ", dS, Kt, dqkv, false));
    if (no_atrans || gemm_get_impl() != 0) {
        bf16* dST = r.Walloc<bf16>(tt * 2);
        launch_transpose(dS, dST, B, tok, tok, tok, (long long)tok * tok, tok, (long long)tok * tok, r.st);
        bf16* PT = r.Walloc<bf16>(tt * 2);
        launch_transpose(P, PT, B, tok, tok, tok, (long long)tok * tok, tok, (long long)tok * tok, r.st);
        RC(tok_gemm("attn.dK", dST, Qt, dqkv + C, false));
        RC(tok_gemm("attn.dV", PT, daT, dqkv + 2 * C, false));
    } else {
        RC(tok_gemm("attn.dK", dS, Qt, dqkv + C, true));
        RC(tok_gemm("attn.dV", P, daT, dqkv + 2 * C, true));
    }
    bf16* dt = r.Walloc<bf16>(act);
    GemmOp gq = dense_lin_op("attn.qkv.dgrad", dqkv, B, h, w, 3 * C, p.qkv.bwd, C, nullptr, nullptr, dt);
    float* pbuf = r.Walloc<float>(fused_partial_bytes(B, h, w));
    const Partials pq = fuse_gn_bwd(r, gq, r.S<bf16>(rec.x), p.gn, rec.g, 0, pbuf, h, w);
    RC(gemm_launch(gq, ns, r.st));
    RC(gn_backward(r, r.S<bf16>(rec.x), dt, p.gn, rec.g, dout, dx, tok, 0, pq));
    r.wsa.reset(m);
    return 0;
}

