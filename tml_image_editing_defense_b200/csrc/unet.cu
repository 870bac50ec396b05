// SD-1.5 UNet2DConditionModel forward + input-gradient backward as a sequence of sm_100a kernels: the denoiser the
// reference calls once per DDIM step inside the attack (main.py:229-243, `self.pipeline.unet(latent_model_input, t,
// encoder_hidden_states=prompt_embeds).sample`) and differentiates with torch.autograd.grad(loss, [cur_image])
// (main.py:176).  Built from the same pieces as the VAE walks: the tcgen05 implicit-GEMM kernel for every
// convolution, linear and attention product; GroupNorm / LayerNorm / GEGLU / head split kernels of unet_kernels.cu.
//
//   sample [B,4,h,w] fp32 -> conv_in -> 4 down blocks (ResnetBlock2D + Transformer2DModel, stride-2 conv between)
//     -> mid (resnet, transformer, resnet) -> 4 up blocks (skip concatenation, nearest-2x + conv between)
//     -> GroupNorm + SiLU -> conv_out -> [B,4,h,w] fp32
//
// Only the gradient w.r.t. `sample` exists (the timestep, the prompt embeddings and the weights are constants of the
// attack), so every backward linear / convolution is a dgrad-only GEMM, the time embedding folds into conv1's bias
// vector, and the cross-attention keys / values need no backward at all.
#include "vae_impl.cuh"

struct TmlUnet;

namespace {

struct ULin {          // linear / 1x1 conv: fwd [co][ci], bwd [ci][co], optional bias
    int ci = 0, co = 0;
    bf16* fwd = nullptr;
    bf16* bwd = nullptr;
    float* bias = nullptr;
};
struct ULn { float* g = nullptr; float* b = nullptr; };
struct UResnet {
    int ci = 0, co = 0;
    Norm n1, n2;
    Conv3 c1, c2;
    bool has_sc = false;
    ULin sc;
    int tb_off = 0;    // offset of this block's conv1 bias (+ time embedding projection) in the per-call bias vector
};
struct UTf {
    int C = 0, heads = 0, d = 0, dp = 0, dp2 = 0;   // dp / dp2: head width padded to 64 (self) / with a mask slot (cross)
    Norm gn;
    ULin proj_in, proj_out;
    ULn ln1, ln2, ln3;
    ULin qkv, out1;       // self attention: [3C][C] without bias, to_out.0
    ULin q2, kv2, out2;   // cross attention: to_q [C][C], [to_k; to_v] [2C][Dctx], to_out.0
    ULin ff1, ff2;        // GEGLU proj [8C][C], net.2 [C][4C]
};

struct ResRec { size_t x = 0, h1 = 0, out = 0; GnSaved g1, g2; int h = 0, w = 0; };
struct TfRec {
    int h = 0, w = 0;
    size_t x = 0, out = 0;
    GnSaved g;
    size_t t0 = 0, st1 = 0, Qh = 0, Kh = 0, Vh = 0, rmax = 0, Oh = 0, invl = 0;   // self attention: P~ is recomputed in the backward
    size_t x1 = 0, st2 = 0, Q2 = 0, K2 = 0, V2 = 0, P2 = 0, rmax2 = 0, O2 = 0, invl2 = 0;   // P2 only on the unfused path
    size_t x2 = 0, st3 = 0, hff = 0;
};
struct SampRec { size_t x = 0, out = 0; int h = 0, w = 0; };   // down / up sampler: (h, w) = input size
struct CatRec { size_t out = 0; int ca = 0, cb = 0, skip = 0, h = 0, w = 0; };

struct Tape {
    std::vector<ResRec> res;
    std::vector<TfRec> tf;
    std::vector<SampRec> down, up;
    std::vector<CatRec> cat;
    std::vector<size_t> skip_off;      // activations pushed as skip connections, in push order
    std::vector<int> skip_c, skip_h, skip_w;
    size_t x0 = 0, xlast = 0;
    GnSaved gout;
    size_t saved_bytes = 0, ws_bytes = 0;
    int B = 0, h = 0, w = 0, T = 0;
    bool valid = false;
};

}  // namespace

struct TmlUnet {
    TmlEncoder base;   // host tensor map, device allocations, device id, SM count (helpers of vae_impl.cuh)
    TmlUnetCfg cfg;
    bool finalized = false;
    Conv3 conv_in, conv_out;
    float* te_w1 = nullptr; float* te_b1 = nullptr; float* te_w2 = nullptr; float* te_b2 = nullptr;
    float* tp_w = nullptr; float* tp_b = nullptr;   // all time_emb_proj stacked [Ntot][temb], bias = proj bias + conv1 bias
    int temb = 0, tb_total = 0;
    std::vector<UResnet> res;    // forward order
    std::vector<UTf> tf;
    std::vector<Conv3> downs, ups;
    Norm norm_out;
    Tape tape;
};

namespace {

struct URun {
    TmlUnet* u;
    char* saved;
    char* ws;
    Arena wsa, sva;
    cudaStream_t st;
    int B;
    template <typename T> T* S(size_t off) const { return reinterpret_cast<T*>(saved + off); }
    template <typename T> T* Walloc(size_t bytes) { return reinterpret_cast<T*>(ws + wsa.alloc(bytes)); }
};

int make_ulin(TmlUnet* u, const std::string& key, int Ci, int Co, bool bias, ULin* l) {
    TmlEncoder* e = &u->base;
    const HostTensor* w = find(e, key + ".weight", (size_t)Co * Ci);
    if (!w) return -20;
    l->ci = Ci; l->co = Co;
    std::vector<uint16_t> f((size_t)Co * Ci), t((size_t)Ci * Co);
    for (int o = 0; o < Co; ++o)
        for (int i = 0; i < Ci; ++i) {
            const uint16_t v = f2bf(w->v[(size_t)o * Ci + i]);
            f[(size_t)o * Ci + i] = v;
            t[(size_t)i * Co + o] = v;
        }
    RC(upload_bf16(e, f, &l->fwd));
    RC(upload_bf16(e, t, &l->bwd));
    if (bias) {
        const HostTensor* b = find(e, key + ".bias", Co);
        if (!b) return -20;
        RC(upload<float>(e, b->v, &l->bias));
    }
    return 0;
}

// rows of several [co_i][ci] matrices stacked (qkv, [k; v]); no bias
int make_ulin_stacked(TmlUnet* u, const std::vector<std::string>& keys, int Ci, int Co_each, ULin* l) {
    TmlEncoder* e = &u->base;
    const int n = (int)keys.size(), Co = n * Co_each;
    l->ci = Ci; l->co = Co;
    std::vector<uint16_t> f((size_t)Co * Ci), t((size_t)Ci * Co);
    for (int q = 0; q < n; ++q) {
        const HostTensor* w = find(e, keys[q] + ".weight", (size_t)Co_each * Ci);
        if (!w) return -20;
        for (int o = 0; o < Co_each; ++o)
            for (int i = 0; i < Ci; ++i) {
                const uint16_t v = f2bf(w->v[(size_t)o * Ci + i]);
                f[((size_t)q * Co_each + o) * Ci + i] = v;
                t[(size_t)i * Co + (size_t)q * Co_each + o] = v;
            }
    }
    RC(upload_bf16(e, f, &l->fwd));
    RC(upload_bf16(e, t, &l->bwd));
    return 0;
}

int make_uln(TmlUnet* u, const std::string& key, int C, ULn* n) {
    TmlEncoder* e = &u->base;
    const HostTensor* g = find(e, key + ".weight", C);
    const HostTensor* b = find(e, key + ".bias", C);
    if (!g || !b) return -20;
    RC(upload<float>(e, g->v, &n->g));
    RC(upload<float>(e, b->v, &n->b));
    return 0;
}

// ResnetBlock2D with time embedding; tp_w / tp_b collect time_emb_proj (+ conv1 bias) of every block
int make_uresnet(TmlUnet* u, const std::string& key, int Ci, int Co, std::vector<float>& tp_w, std::vector<float>& tp_b,
                 UResnet* r) {
    TmlEncoder* e = &u->base;
    r->ci = Ci; r->co = Co;
    RC(make_norm(e, key + ".norm1", Ci, &r->n1));
    RC(make_conv3(e, key + ".conv1", Ci, Co, 1, &r->c1));
    RC(make_norm(e, key + ".norm2", Co, &r->n2));
    RC(make_conv3(e, key + ".conv2", Co, Co, 1, &r->c2));
    r->has_sc = Ci != Co;
    if (r->has_sc) RC(make_ulin(u, key + ".conv_shortcut", Ci, Co, true, &r->sc));
    const HostTensor* tw = find(e, key + ".time_emb_proj.weight", (size_t)Co * u->temb);
    const HostTensor* tb = find(e, key + ".time_emb_proj.bias", Co);
    const HostTensor* cb = find(e, key + ".conv1.bias", Co);
    if (!tw || !tb || !cb) return -20;
    r->tb_off = (int)tp_b.size();
    tp_w.insert(tp_w.end(), tw->v.begin(), tw->v.end());
    for (int i = 0; i < Co; ++i) tp_b.push_back(tb->v[i] + cb->v[i]);
    return 0;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

int make_utf(TmlUnet* u, const std::string& key, int C, UTf* t) {
    TmlEncoder* e = &u->base;
    const int heads = u->cfg.num_heads, Dc = u->cfg.cross_attention_dim;
    if (C % heads || (C / heads) % 8) { set_error("unet: head width %d/%d must be a multiple of 8", C, heads); return -1; }
    t->C = C; t->heads = heads; t->d = C / heads;
    t->dp = round_up(t->d, 64);
    t->dp2 = round_up(t->d + 1, 64);
    RC(make_norm(e, key + ".norm", C, &t->gn));
    RC(make_ulin(u, key + ".proj_in", C, C, true, &t->proj_in));
    RC(make_ulin(u, key + ".proj_out", C, C, true, &t->proj_out));
    const std::string b = key + ".transformer_blocks.0";
    RC(make_uln(u, b + ".norm1", C, &t->ln1));
    RC(make_uln(u, b + ".norm2", C, &t->ln2));
    RC(make_uln(u, b + ".norm3", C, &t->ln3));
    RC(make_ulin_stacked(u, {b + ".attn1.to_q", b + ".attn1.to_k", b + ".attn1.to_v"}, C, C, &t->qkv));
    RC(make_ulin(u, b + ".attn1.to_out.0", C, C, true, &t->out1));
    RC(make_ulin(u, b + ".attn2.to_q", C, C, false, &t->q2));
    RC(make_ulin_stacked(u, {b + ".attn2.to_k", b + ".attn2.to_v"}, Dc, C, &t->kv2));
    RC(make_ulin(u, b + ".attn2.to_out.0", C, C, true, &t->out2));
    RC(make_ulin(u, b + ".ff.net.0.proj", C, 8 * C, true, &t->ff1));
    RC(make_ulin(u, b + ".ff.net.2", 4 * C, C, true, &t->ff2));
    return 0;
}

int make_down(TmlUnet* u, const std::string& key, int C, Conv3* c) {   // 3x3 stride 2, symmetric pad 1
    TmlEncoder* e = &u->base;
    const HostTensor* w = find(e, key + ".weight", (size_t)C * C * 9);
    const HostTensor* b = find(e, key + ".bias", C);
    if (!w || !b) return -20;
    c->ci = C; c->co = C; c->stride = 2;
    RC(make_packed(e, w->v.data(), C, C, 7, &c->fwd));
    for (int q = 0; q < 4; ++q) RC(make_packed(e, w->v.data(), C, C, 8 + q, &c->bwd_par[q]));
    RC(upload<float>(e, b->v, &c->bias));
    return 0;
}

GemmOp lin_op(const char* name, const bf16* A, int B, int h, int w, int K, const bf16* Wm, int N, const float* bias,
              const bf16* resid, bf16* D) {
    return dense_lin_op(name, A, B, h, w, K, Wm, N, bias, resid, D);
}

size_t gn_part_bytes(int B, int hw, int C) { return (size_t)B * gng_num_chunks(hw, C) * 32 * 2 * sizeof(float); }

int ugn_forward(URun& r, const bf16* x, const Norm& n, const GnSaved& g, bf16* y, int hw, int silu, float eps) {
    const size_t m = r.wsa.mark();
    float* part = r.Walloc<float>(gn_part_bytes(r.B, hw, n.C));
    launch_gng_stats(x, part, r.B, hw, n.C, r.st);
    launch_gn_finalize(part, n.gamma, n.beta, r.S<float2>(g.ss), r.S<float2>(g.mr), r.B, hw, n.C, eps,
                       gng_num_chunks(hw, n.C), r.st);
    launch_gng_apply(x, r.S<float2>(g.ss), y, r.B, hw, n.C, silu, r.st);
    r.wsa.reset(m);
    return 0;
}
int ugn_backward(URun& r, const bf16* x, const bf16* dy, const Norm& n, const GnSaved& g, const bf16* resid, bf16* dx,
                 int hw, int silu) {
    const size_t m = r.wsa.mark();
    float* part = r.Walloc<float>(gn_part_bytes(r.B, hw, n.C));
    launch_gng_bwd_partial(x, dy, r.S<float2>(g.ss), r.S<float2>(g.mr), n.gamma, part, r.B, hw, n.C, silu, r.st);
    float2* mm = r.Walloc<float2>((size_t)r.B * 32 * sizeof(float2));
    launch_gn_bwd_finalize(part, mm, r.B, hw, n.C, gng_num_chunks(hw, n.C), r.st);
    launch_gng_bwd_apply(x, dy, r.S<float2>(g.ss), r.S<float2>(g.mr), mm, resid, dx, r.B, hw, n.C, silu, r.st);
    r.wsa.reset(m);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ResnetBlock2D
// ------------------------------------------------------------------------------------------------
int ures_forward(URun& r, const UResnet& p, ResRec& rec, size_t x_off, int h, int w, const float* tbias) {
    const int B = r.B, hw = h * w, ns = r.u->base.num_sms;
    rec.h = h; rec.w = w; rec.x = x_off;
    rec.g1 = alloc_gn(r.sva, B, p.ci);
    rec.h1 = r.sva.alloc(act_bytes(B, h, w, p.co));
    rec.g2 = alloc_gn(r.sva, B, p.co);
    rec.out = r.sva.alloc(act_bytes(B, h, w, p.co));
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* a = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
    RC(ugn_forward(r, x, p.n1, rec.g1, a, hw, 1, 1e-5f));
    // conv1 + (bias + time_emb_proj(silu(temb))): one vector per call, the same for every image
    RC(gemm_launch(dense_conv_op("unet.resnet.conv1", a, B, h, w, p.ci, p.c1.fwd, p.co, 1, h, w,
                                 tbias ? tbias + p.tb_off : nullptr, nullptr, r.S<bf16>(rec.h1)), ns, r.st));
    bf16* a2 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    RC(ugn_forward(r, r.S<bf16>(rec.h1), p.n2, rec.g2, a2, hw, 1, 1e-5f));
    const bf16* resid = x;
    if (p.has_sc) {
        bf16* sc = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
        RC(gemm_launch(lin_op("unet.resnet.shortcut", x, B, h, w, p.ci, p.sc.fwd, p.co, p.sc.bias, nullptr, sc), ns, r.st));
        resid = sc;
    }
    RC(gemm_launch(dense_conv_op("unet.resnet.conv2", a2, B, h, w, p.co, p.c2.fwd, p.co, 1, h, w, p.c2.bias, resid,
                                 r.S<bf16>(rec.out)), ns, r.st));
    r.wsa.reset(m);
    return 0;
}

int ures_backward(URun& r, const UResnet& p, const ResRec& rec, const bf16* dout, bf16* dx) {
    const int B = r.B, h = rec.h, w = rec.w, hw = h * w, ns = r.u->base.num_sms;
    const size_t m = r.wsa.mark();
    bf16* d_a2 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    RC(gemm_launch(dense_conv_op("unet.resnet.conv2.dgrad", dout, B, h, w, p.co, p.c2.bwd, p.co, 1, h, w, nullptr, nullptr,
                                 d_a2), ns, r.st));
    bf16* d_h1 = r.Walloc<bf16>(act_bytes(B, h, w, p.co));
    RC(ugn_backward(r, r.S<bf16>(rec.h1), d_a2, p.n2, rec.g2, nullptr, d_h1, hw, 1));
    bf16* d_a1 = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
    RC(gemm_launch(dense_conv_op("unet.resnet.conv1.dgrad", d_h1, B, h, w, p.co, p.c1.bwd, p.ci, 1, h, w, nullptr, nullptr,
                                 d_a1), ns, r.st));
    if (p.has_sc) {
        bf16* tmp = r.Walloc<bf16>(act_bytes(B, h, w, p.ci));
        RC(ugn_backward(r, r.S<bf16>(rec.x), d_a1, p.n1, rec.g1, nullptr, tmp, hw, 1));
        RC(gemm_launch(lin_op("unet.resnet.shortcut.dgrad", dout, B, h, w, p.co, p.sc.bwd, p.ci, nullptr, tmp, dx), ns, r.st));
    } else {
        RC(ugn_backward(r, r.S<bf16>(rec.x), d_a1, p.n1, rec.g1, dout, dx, hw, 1));
    }
    r.wsa.reset(m);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// multi-head attention on the GEMM kernels: batch = (image, head), every head a dense [tokens][dp] matrix
// ------------------------------------------------------------------------------------------------
GemmOp mh_logits_op(const char* name, const bf16* A, const bf16* Bk, int nb, int tq, int tkv, int dp, int gh, int gw) {
    GemmOp o;   // D[tq, tkv] = A[tq, dp] * Bk[tkv, dp]^T per batch
    o.name = name;
    o.A = A; o.A_C = dp; o.A_W = gw; o.A_H = gh; o.A_B = nb;
    o.A_sW = dp; o.A_sH = (int64_t)gw * dp; o.A_sB = (int64_t)tq * dp;
    o.OW = gw; o.OH = gh;
    o.Bm = Bk; o.N = tkv; o.B_sN = dp; o.B_sBatch = (int64_t)tkv * dp;
    o.D_sW = tkv; o.D_sH = (int64_t)gw * tkv; o.D_sB = (int64_t)tq * tkv; o.D_sN = 1;
    return o;
}
// D[tq, dp] = A[tq, tk] * Bt[dp, tk]^T per batch (dense output rows of dp)
GemmOp mh_tok_op(const char* name, const bf16* A, const bf16* Bt, bf16* D, int nb, int tq, int tk, int dp, int gh, int gw) {
    GemmOp o;
    o.name = name;
    o.A = A; o.A_C = tk; o.A_W = gw; o.A_H = gh; o.A_B = nb;
    o.A_sW = tk; o.A_sH = (int64_t)gw * tk; o.A_sB = (int64_t)tq * tk;
    o.OW = gw; o.OH = gh;
    o.Bm = Bt; o.N = dp; o.B_sN = tk; o.B_sBatch = (int64_t)dp * tk;
    o.D = D; o.D_sW = dp; o.D_sH = (int64_t)gw * dp; o.D_sB = (int64_t)tq * dp; o.D_sN = 1;
    return o;
}

// P~ = exp2((Q K^T - rowmax) * scale * log2 e) as bf16 [nb][tq][tkv] from known row maxima; `part` receives the row sums
// per column tile (scratch of rows x gemm_row_partials floats).  Bitwise reproducible: the backward calls it again
// instead of keeping 2 bytes per logit alive (4.3 GB per layer at 64 x 64 latents and batch 16).
int mh_probs(URun& r, const bf16* Q, const bf16* K, const float* rmax, float* part, bf16* P, int nb, int tq, int tkv, int dp,
             int gh, int gw, float scale, const char* name) {
    GemmOp e = mh_logits_op(name, Q, K, nb, tq, tkv, dp, gh, gw);
    e.epi_mode = 2;
    e.row_a = rmax; e.row_part = part;
    e.exp_scale = scale * 1.4426950408889634f;
    e.D = P;
    return gemm_launch(e, r.u->base.num_sms, r.st);
}

// softmax(Q K^T * scale) V for nb batches: the row maxima (rmax) and 1/l are kept for the backward, P~ only if the
// caller passes a saved buffer (cross attention: 128 key slots); O dense [nb][tq][dp]
int mh_attention_forward(URun& r, const bf16* Q, const bf16* K, const bf16* V, bf16* P, float* rmax, float* inv_l, bf16* O,
                         int nb, int tq, int tkv, int dp, float scale, int lcol = -1) {
    const int ns = r.u->base.num_sms;
    const long long rows = (long long)nb * tq;
    int gh = 0, gw = 0;
    RC(attn_grid(tq, &gh, &gw));
    const size_t m = r.wsa.mark();
    static const bool known_max = getenv("TML_ATTN_KNOWN_MAX") && getenv("TML_ATTN_KNOWN_MAX")[0] == '1';   // A/B switch
    const bool fused = P == nullptr && lcol >= 0 && attn_fused_supported(tq, tkv, dp);
    if (fused && !known_max && rmax != nullptr) {
        // fused, online softmax (attn_fused.cu): no max pass; S, P~ and O never leave the SM
        RC(launch_attn_fused_fwd(Q, K, V, nullptr, rmax, inv_l, O, nb, tq, tkv, dp, lcol, scale, r.st));
        r.wsa.reset(m);
        return 0;
    }
    GemmOp o = mh_logits_op("unet.attn.qk.max", Q, K, nb, tq, tkv, dp, gh, gw);
    o.epi_mode = 1;
    const int np = gemm_row_partials(o);   // host-only planning: also valid during a dry run
    if (np < 0) return np;
    float* part = r.Walloc<float>((size_t)rows * np * sizeof(float));
    if (rmax == nullptr) rmax = r.Walloc<float>((size_t)rows * sizeof(float));
    o.row_part = part;
    RC(gemm_launch(o, ns, r.st));
    launch_row_reduce(part, rmax, rows, np, 0, r.st);
    if (fused) {
        // fused with the row maxima of the pass above
        RC(launch_attn_fused_fwd(Q, K, V, rmax, nullptr, inv_l, O, nb, tq, tkv, dp, lcol, scale, r.st));
        r.wsa.reset(m);
        return 0;
    }
    if (P == nullptr) P = r.Walloc<bf16>((size_t)rows * tkv * sizeof(bf16));
    RC(mh_probs(r, Q, K, rmax, part, P, nb, tq, tkv, dp, gh, gw, scale, "unet.attn.qk.exp"));
    launch_row_reduce(part, inv_l, rows, np, 1, r.st);
    bf16* Vt = r.Walloc<bf16>((size_t)nb * tkv * dp * sizeof(bf16));   // [nb][dp][tkv]
    launch_transpose(V, Vt, nb, tkv, dp, dp, (long long)tkv * dp, tkv, (long long)dp * tkv, r.st);
    GemmOp pv = mh_tok_op("unet.attn.pv", P, Vt, O, nb, tq, tkv, dp, gh, gw);
    pv.row_scale = inv_l;
    RC(gemm_launch(pv, ns, r.st));
    r.wsa.reset(m);
    return 0;
}

// dO [nb][tq][dp] -> dQ (always), dK / dV (self attention only; null for cross attention)
int mh_attention_backward(URun& r, const bf16* Q, const bf16* K, const bf16* V, const bf16* P, const float* rmax,
                          const float* inv_l, const bf16* O, const bf16* dO, bf16* dQ, bf16* dK, bf16* dV, int nb, int tq,
                          int tkv, int dp, float scale) {
    const int ns = r.u->base.num_sms;
    const long long rows = (long long)nb * tq;
    int gh = 0, gw = 0;
    RC(attn_grid(tq, &gh, &gw));
    const size_t m = r.wsa.mark();
    static const bool no_fused_bwd = getenv("TML_NO_FUSED_ATTN_BWD") && getenv("TML_NO_FUSED_ATTN_BWD")[0] == '1';   // A/B switch
    if (P == nullptr && !no_fused_bwd && attn_fused_supported(tq, tkv, dp)) {
        // fused (attn_fused.cu): S and dP are recomputed on the tensor cores inside the two backward kernels
        bf16* dOs = r.Walloc<bf16>((size_t)rows * dp * sizeof(bf16));
        float* Dp = r.Walloc<float>(((size_t)rows * 3 + 8) * sizeof(float));   // D' + per-query constants
        RC(launch_attn_fused_bwd(Q, K, V, O, dO, rmax, inv_l, dOs, Dp, dQ, dK, dV, nb, tq, tkv, dp, scale, r.st));
        r.wsa.reset(m);
        return 0;
    }
    if (P == nullptr) {   // not kept by the forward: the same GEMM + epilogue gives the same bits again
        GemmOp o = mh_logits_op("unet.attn.qk.exp", Q, K, nb, tq, tkv, dp, gh, gw);
        o.epi_mode = 2;
        const int np = gemm_row_partials(o);
        if (np < 0) return np;
        bf16* Pn = r.Walloc<bf16>((size_t)rows * tkv * sizeof(bf16));
        float* part = r.Walloc<float>((size_t)rows * np * sizeof(float));
        RC(mh_probs(r, Q, K, rmax, part, Pn, nb, tq, tkv, dp, gh, gw, scale, "unet.attn.qk.exp.recompute"));
        P = Pn;
    }
    float* Drow = r.Walloc<float>((size_t)rows * sizeof(float));
    launch_row_dot(dO, O, Drow, rows, dp, r.st);
    bf16* dS = r.Walloc<bf16>((size_t)nb * tq * tkv * sizeof(bf16));
    {
        GemmOp o = mh_logits_op("unet.attn.dS", dO, V, nb, tq, tkv, dp, gh, gw);   // dP = dO V^T
        o.epi_mode = 3;
        o.alpha = scale;
        o.row_a = Drow; o.row_b = inv_l;
        o.resid = P; o.R_sW = tkv; o.R_sH = (int64_t)gw * tkv; o.R_sB = (int64_t)tq * tkv;
        o.D = dS;
        RC(gemm_launch(o, ns, r.st));
    }
    bf16* Kt = r.Walloc<bf16>((size_t)nb * tkv * dp * sizeof(bf16));   // [nb][dp][tkv]
    launch_transpose(K, Kt, nb, tkv, dp, dp, (long long)tkv * dp, tkv, (long long)dp * tkv, r.st);
    RC(gemm_launch(mh_tok_op("unet.attn.dQ", dS, Kt, dQ, nb, tq, tkv, dp, gh, gw), ns, r.st));
    if (dK != nullptr && dV != nullptr) {
        // dK = dS^T Q and dV = P~^T (dO / l): the token x token matrices are read TRANSPOSED by the tensor cores
        // (GemmOp::a_trans, MN-major operand), so no transposed copy of dS or P~ is ever written
        int kh = 0, kw = 0;
        RC(attn_grid(tkv, &kh, &kw));
        bf16* Qt = r.Walloc<bf16>((size_t)nb * tq * dp * sizeof(bf16));
        launch_transpose(Q, Qt, nb, tq, dp, dp, (long long)tq * dp, tq, (long long)dp * tq, r.st);
        auto trans_op = [&](const char* name, const bf16* A, const bf16* Bt, bf16* D) {
            GemmOp o = mh_tok_op(name, A, Bt, D, nb, tkv, tq, dp, kh, kw);
            o.a_trans = 1; o.A_sK = tkv; o.A_sB = (int64_t)tq * tkv;   // A[b][k = query][m = key]
            return gemm_launch(o, ns, r.st);
        };
        RC(trans_op("unet.attn.dK", dS, Qt, dK));
        bf16* dOT = Qt;   // (dO / l)^T (Qt has been consumed by the launch above, same stream)
        launch_transpose(dO, dOT, nb, tq, dp, dp, (long long)tq * dp, tq, (long long)dp * tq, r.st, inv_l);
        RC(trans_op("unet.attn.dV", P, dOT, dV));
    }
    r.wsa.reset(m);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Transformer2DModel (GroupNorm -> proj_in -> BasicTransformerBlock -> proj_out + residual)
// ------------------------------------------------------------------------------------------------
int utf_forward(URun& r, const UTf& p, TfRec& rec, size_t x_off, int h, int w, const bf16* ctxp, int T, int Tp) {
    const int B = r.B, C = p.C, H = p.heads, d = p.d, dp = p.dp, dp2 = p.dp2, ns = r.u->base.num_sms;
    const int tok = h * w, nb = B * H;
    const long long rows = (long long)B * tok;
    const float scale = 1.0f / sqrtf((float)d);
    const size_t act = act_bytes(B, h, w, C);
    rec.h = h; rec.w = w; rec.x = x_off;
    rec.g = alloc_gn(r.sva, B, C);
    rec.t0 = r.sva.alloc(act);
    rec.st1 = r.sva.alloc((size_t)rows * sizeof(float2));
    rec.Qh = r.sva.alloc((size_t)nb * tok * dp * sizeof(bf16));
    rec.Kh = r.sva.alloc((size_t)nb * tok * dp * sizeof(bf16));
    rec.Vh = r.sva.alloc((size_t)nb * tok * dp * sizeof(bf16));
    rec.rmax = r.sva.alloc((size_t)nb * tok * sizeof(float));
    rec.Oh = r.sva.alloc((size_t)nb * tok * dp * sizeof(bf16));
    rec.invl = r.sva.alloc((size_t)nb * tok * sizeof(float));
    rec.x1 = r.sva.alloc(act);
    rec.st2 = r.sva.alloc((size_t)rows * sizeof(float2));
    rec.Q2 = r.sva.alloc((size_t)nb * tok * dp2 * sizeof(bf16));
    rec.K2 = r.sva.alloc((size_t)nb * Tp * dp2 * sizeof(bf16));
    rec.V2 = r.sva.alloc((size_t)nb * Tp * dp2 * sizeof(bf16));
    const bool fused2 = attn_fused_supported(tok, Tp, dp2);   // cross attention on the fused kernels: P~ is never stored
    if (!fused2) rec.P2 = r.sva.alloc((size_t)nb * tok * Tp * sizeof(bf16));
    rec.rmax2 = r.sva.alloc((size_t)nb * tok * sizeof(float));
    rec.O2 = r.sva.alloc((size_t)nb * tok * dp2 * sizeof(bf16));
    rec.invl2 = r.sva.alloc((size_t)nb * tok * sizeof(float));
    rec.x2 = r.sva.alloc(act);
    rec.st3 = r.sva.alloc((size_t)rows * sizeof(float2));
    rec.hff = r.sva.alloc((size_t)rows * 8 * C * sizeof(bf16));
    rec.out = r.sva.alloc(act);
    const bf16* x = r.S<bf16>(rec.x);
    const size_t m = r.wsa.mark();
    bf16* n = r.Walloc<bf16>(act);          // normalised activations (GroupNorm / LayerNorm outputs), reused
    bf16* a = r.Walloc<bf16>(act);          // merged attention output, reused
    RC(ugn_forward(r, x, p.gn, rec.g, n, tok, 0, 1e-6f));
    bf16* t0 = r.S<bf16>(rec.t0);
    RC(gemm_launch(lin_op("unet.tf.proj_in", n, B, h, w, C, p.proj_in.fwd, C, p.proj_in.bias, nullptr, t0), ns, r.st));
    // ---- self attention
    launch_ln_fwd(t0, p.ln1.g, p.ln1.b, n, r.S<float2>(rec.st1), rows, C, 1e-5f, r.st);
    {
        const size_t m2 = r.wsa.mark();
        bf16* qkv = r.Walloc<bf16>(3 * act);
        RC(gemm_launch(lin_op("unet.attn1.qkv", n, B, h, w, C, p.qkv.fwd, 3 * C, nullptr, nullptr, qkv), ns, r.st));
        const long long bs = (long long)tok * 3 * C;
        launch_head_split(qkv, 3 * C, bs, 0, r.S<bf16>(rec.Qh), B, H, tok, tok, tok, d, dp, 0, r.st);
        launch_head_split(qkv, 3 * C, bs, C, r.S<bf16>(rec.Kh), B, H, tok, tok, tok, d, dp, 0, r.st);
        // (V's first padding channel is 1.0: the fused attention reads the softmax denominator off that column of O)
        launch_head_split(qkv, 3 * C, bs, 2 * C, r.S<bf16>(rec.Vh), B, H, tok, tok, tok, d, dp, 1, r.st);
        r.wsa.reset(m2);
    }
    RC(mh_attention_forward(r, r.S<bf16>(rec.Qh), r.S<bf16>(rec.Kh), r.S<bf16>(rec.Vh), nullptr, r.S<float>(rec.rmax),
                            r.S<float>(rec.invl), r.S<bf16>(rec.Oh), nb, tok, tok, dp, scale, d < dp ? d : -1));
    launch_head_merge(r.S<bf16>(rec.Oh), a, C, (long long)tok * C, 0, B, H, tok, d, dp, r.st);
    bf16* x1 = r.S<bf16>(rec.x1);
    RC(gemm_launch(lin_op("unet.attn1.out", a, B, h, w, C, p.out1.fwd, C, p.out1.bias, t0, x1), ns, r.st));
    // ---- cross attention (keys / values from the prompt embeddings, padded key tokens masked through the spare channel)
    launch_ln_fwd(x1, p.ln2.g, p.ln2.b, n, r.S<float2>(rec.st2), rows, C, 1e-5f, r.st);
    {
        const size_t m2 = r.wsa.mark();
        bf16* q = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.attn2.q", n, B, h, w, C, p.q2.fwd, C, nullptr, nullptr, q), ns, r.st));
        launch_head_split(q, C, (long long)tok * C, 0, r.S<bf16>(rec.Q2), B, H, tok, tok, tok, d, dp2, 1, r.st);
        bf16* kv = r.Walloc<bf16>((size_t)B * Tp * 2 * C * sizeof(bf16));
        RC(gemm_launch(lin_op("unet.attn2.kv", ctxp, B, 1, Tp, p.kv2.ci, p.kv2.fwd, 2 * C, nullptr, nullptr, kv), ns, r.st));
        launch_head_split(kv, 2 * C, (long long)Tp * 2 * C, 0, r.S<bf16>(rec.K2), B, H, Tp, T, Tp, d, dp2, 2, r.st);
        // (V's padding channel d = 1.0 on every row: the fused kernel's denominator column; padded keys have P~ = 0)
        launch_head_split(kv, 2 * C, (long long)Tp * 2 * C, C, r.S<bf16>(rec.V2), B, H, Tp, Tp, Tp, d, dp2, 1, r.st);
        r.wsa.reset(m2);
    }
    RC(mh_attention_forward(r, r.S<bf16>(rec.Q2), r.S<bf16>(rec.K2), r.S<bf16>(rec.V2), fused2 ? nullptr : r.S<bf16>(rec.P2),
                            r.S<float>(rec.rmax2), r.S<float>(rec.invl2), r.S<bf16>(rec.O2), nb, tok, Tp, dp2, scale,
                            fused2 ? d : -1));
    launch_head_merge(r.S<bf16>(rec.O2), a, C, (long long)tok * C, 0, B, H, tok, d, dp2, r.st);
    bf16* x2 = r.S<bf16>(rec.x2);
    RC(gemm_launch(lin_op("unet.attn2.out", a, B, h, w, C, p.out2.fwd, C, p.out2.bias, x1, x2), ns, r.st));
    // ---- GEGLU feed-forward
    launch_ln_fwd(x2, p.ln3.g, p.ln3.b, n, r.S<float2>(rec.st3), rows, C, 1e-5f, r.st);
    bf16* hff = r.S<bf16>(rec.hff);
    RC(gemm_launch(lin_op("unet.ff.proj", n, B, h, w, C, p.ff1.fwd, 8 * C, p.ff1.bias, nullptr, hff), ns, r.st));
    {
        const size_t m2 = r.wsa.mark();
        bf16* g = r.Walloc<bf16>(4 * act);
        launch_geglu_fwd(hff, g, rows, 4 * C, r.st);
        bf16* x3 = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.ff.out", g, B, h, w, 4 * C, p.ff2.fwd, C, p.ff2.bias, x2, x3), ns, r.st));
        RC(gemm_launch(lin_op("unet.tf.proj_out", x3, B, h, w, C, p.proj_out.fwd, C, p.proj_out.bias, x, r.S<bf16>(rec.out)),
                       ns, r.st));
        r.wsa.reset(m2);
    }
    r.wsa.reset(m);
    return 0;
}

int utf_backward(URun& r, const UTf& p, const TfRec& rec, const bf16* dout, bf16* dx, int Tp) {
    const int B = r.B, C = p.C, H = p.heads, d = p.d, dp = p.dp, dp2 = p.dp2, ns = r.u->base.num_sms;
    const int h = rec.h, w = rec.w, tok = h * w, nb = B * H;
    const long long rows = (long long)B * tok;
    const float scale = 1.0f / sqrtf((float)d);
    const size_t act = act_bytes(B, h, w, C);
    const size_t m = r.wsa.mark();
    bf16* g3 = r.Walloc<bf16>(act);   // d(x3)
    RC(gemm_launch(lin_op("unet.tf.proj_out.dgrad", dout, B, h, w, C, p.proj_out.bwd, C, nullptr, nullptr, g3), ns, r.st));
    bf16* g2 = r.Walloc<bf16>(act);   // d(x2)
    {   // feed-forward
        const size_t m2 = r.wsa.mark();
        bf16* dg = r.Walloc<bf16>(4 * act);
        RC(gemm_launch(lin_op("unet.ff.out.dgrad", g3, B, h, w, C, p.ff2.bwd, 4 * C, nullptr, nullptr, dg), ns, r.st));
        bf16* dh = r.Walloc<bf16>(8 * act);
        launch_geglu_bwd(r.S<bf16>(rec.hff), dg, dh, rows, 4 * C, r.st);
        bf16* dn = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.ff.proj.dgrad", dh, B, h, w, 8 * C, p.ff1.bwd, C, nullptr, nullptr, dn), ns, r.st));
        launch_ln_bwd(r.S<bf16>(rec.x2), dn, p.ln3.g, r.S<float2>(rec.st3), g3, g2, rows, C, r.st);
        r.wsa.reset(m2);
    }
    bf16* g1 = g3;                    // d(x1) (g3 is dead after the LayerNorm backward above)
    {   // cross attention: only the query path carries a gradient
        const size_t m2 = r.wsa.mark();
        bf16* da = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.attn2.out.dgrad", g2, B, h, w, C, p.out2.bwd, C, nullptr, nullptr, da), ns, r.st));
        bf16* dO = r.Walloc<bf16>((size_t)nb * tok * dp2 * sizeof(bf16));
        launch_head_split(da, C, (long long)tok * C, 0, dO, B, H, tok, tok, tok, d, dp2, 0, r.st);
        bf16* dQ = r.Walloc<bf16>((size_t)nb * tok * dp2 * sizeof(bf16));
        const bool fused2 = attn_fused_supported(tok, Tp, dp2);
        RC(mh_attention_backward(r, r.S<bf16>(rec.Q2), r.S<bf16>(rec.K2), r.S<bf16>(rec.V2),
                                 fused2 ? nullptr : r.S<bf16>(rec.P2), r.S<float>(rec.rmax2), r.S<float>(rec.invl2),
                                 r.S<bf16>(rec.O2), dO, dQ, nullptr, nullptr, nb, tok, Tp, dp2, scale));
        launch_head_merge(dQ, da, C, (long long)tok * C, 0, B, H, tok, d, dp2, r.st);
        bf16* dn = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.attn2.q.dgrad", da, B, h, w, C, p.q2.bwd, C, nullptr, nullptr, dn), ns, r.st));
        launch_ln_bwd(r.S<bf16>(rec.x1), dn, p.ln2.g, r.S<float2>(rec.st2), g2, g1, rows, C, r.st);
        r.wsa.reset(m2);
    }
    bf16* g0 = g2;                    // d(t0)
    {   // self attention
        const size_t m2 = r.wsa.mark();
        bf16* da = r.Walloc<bf16>(act);
        RC(gemm_launch(lin_op("unet.attn1.out.dgrad", g1, B, h, w, C, p.out1.bwd, C, nullptr, nullptr, da), ns, r.st));
        const size_t hb = (size_t)nb * tok * dp * sizeof(bf16);
        bf16* dO = r.Walloc<bf16>(hb);
        launch_head_split(da, C, (long long)tok * C, 0, dO, B, H, tok, tok, tok, d, dp, 0, r.st);
        bf16* dQ = r.Walloc<bf16>(hb);
        bf16* dK = r.Walloc<bf16>(hb);
        bf16* dV = r.Walloc<bf16>(hb);
        RC(mh_attention_backward(r, r.S<bf16>(rec.Qh), r.S<bf16>(rec.Kh), r.S<bf16>(rec.Vh), nullptr, r.S<float>(rec.rmax),
                                 r.S<float>(rec.invl), r.S<bf16>(rec.Oh), dO, dQ, dK, dV, nb, tok, tok, dp, scale));
        bf16* dqkv = r.Walloc<bf16>(3 * act);
        const long long bs = (long long)tok * 3 * C;
        launch_head_merge(dQ, dqkv, 3 * C, bs, 0, B, H, tok, d, dp, r.st);
        launch_head_merge(dK, dqkv, 3 * C, bs, C, B, H, tok, d, dp, r.st);
        launch_head_merge(dV, dqkv, 3 * C, bs, 2 * C, B, H, tok, d, dp, r.st);
        bf16* dn = da;
        RC(gemm_launch(lin_op("unet.attn1.qkv.dgrad", dqkv, B, h, w, 3 * C, p.qkv.bwd, C, nullptr, nullptr, dn), ns, r.st));
        launch_ln_bwd(r.S<bf16>(rec.t0), dn, p.ln1.g, r.S<float2>(rec.st1), g1, g0, rows, C, r.st);
        r.wsa.reset(m2);
    }
    bf16* dt = g1;
    RC(gemm_launch(lin_op("unet.tf.proj_in.dgrad", g0, B, h, w, C, p.proj_in.bwd, C, nullptr, nullptr, dt), ns, r.st));
    RC(ugn_backward(r, r.S<bf16>(rec.x), dt, p.gn, rec.g, dout, dx, tok, 0));
    r.wsa.reset(m);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the walks
// ------------------------------------------------------------------------------------------------
void ctx_pack(URun& r, const float* ctx, bf16* out, int B, int T, int Tp, int D);

int unet_forward_walk(TmlUnet* u, URun& r, Tape& tp, const float* sample, float t, const float* ctx, int h, int w, int T,
                      float* out) {
    const TmlUnetCfg& c = u->cfg;
    const int B = r.B, nb = c.num_blocks, ns = u->base.num_sms;
    const int C0 = c.block_out_channels[0];
    const int Tp = round_up(T, 64);
    tp = Tape();
    tp.B = B; tp.h = h; tp.w = w; tp.T = T;
    // per-call vectors: timestep embedding -> every resnet's conv1 bias
    float* e0 = r.Walloc<float>((size_t)C0 * sizeof(float));
    float* e1 = r.Walloc<float>((size_t)u->temb * sizeof(float));
    float* temb = r.Walloc<float>((size_t)u->temb * sizeof(float));
    float* tbias = r.Walloc<float>((size_t)u->tb_total * sizeof(float));
    launch_timestep_embed(t, e0, C0, r.st);
    launch_small_linear(e0, u->te_w1, u->te_b1, e1, u->temb, C0, 0, r.st);
    launch_small_linear(e1, u->te_w2, u->te_b2, temb, u->temb, u->temb, 1, r.st);
    launch_small_linear(temb, u->tp_w, u->tp_b, tbias, u->tb_total, u->temb, 1, r.st);
    bf16* ctxp = r.Walloc<bf16>((size_t)B * Tp * c.cross_attention_dim * sizeof(bf16));
    ctx_pack(r, ctx, ctxp, B, T, Tp, c.cross_attention_dim);
    {   // conv_in (4 -> C0, input channels padded to 64)
        const size_t m = r.wsa.mark();
        bf16* xp = r.Walloc<bf16>(act_bytes(B, h, w, 64));
        launch_nchw_pack64(sample, xp, B, c.in_channels, h * w, r.st);
        tp.x0 = r.sva.alloc(act_bytes(B, h, w, C0));
        RC(gemm_launch(dense_conv_op("unet.conv_in", xp, B, h, w, 64, u->conv_in.fwd, C0, 1, h, w, u->conv_in.bias, nullptr,
                                     r.S<bf16>(tp.x0)), ns, r.st));
        r.wsa.reset(m);
    }
    auto push_skip = [&](size_t off, int ch, int hh, int ww) {
        tp.skip_off.push_back(off); tp.skip_c.push_back(ch); tp.skip_h.push_back(hh); tp.skip_w.push_back(ww);
    };
    size_t cur = tp.x0;
    int hh = h, ww = w, cin = C0;
    push_skip(cur, C0, hh, ww);
    size_t ri = 0, ti = 0, di = 0, ui = 0;
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[i];
        for (int j = 0; j < c.layers_per_block; ++j) {
            tp.res.emplace_back();
            RC(ures_forward(r, u->res[ri], tp.res.back(), cur, hh, ww, tbias)); ++ri;
            cur = tp.res.back().out;
            if (c.down_has_attn[i]) {
                tp.tf.emplace_back();
                RC(utf_forward(r, u->tf[ti], tp.tf.back(), cur, hh, ww, ctxp, T, Tp)); ++ti;
                cur = tp.tf.back().out;
            }
            push_skip(cur, cout, hh, ww);
            cin = cout;
        }
        if (i != nb - 1) {
            if (hh % 2 || ww % 2) { set_error("unet: latent size must be a multiple of %d", 1 << (nb - 1)); return -30; }
            SampRec d;
            d.h = hh; d.w = ww; d.x = cur;
            hh /= 2; ww /= 2;
            d.out = r.sva.alloc(act_bytes(B, hh, ww, cout));
            const Conv3& cv = u->downs[di]; ++di;
            RC(gemm_launch(dense_conv_op("unet.downsample", r.S<bf16>(d.x), B, d.h, d.w, cv.ci, cv.fwd, cv.co, 2, hh, ww, cv.bias,
                                         nullptr, r.S<bf16>(d.out)), ns, r.st));
            tp.down.push_back(d);
            cur = d.out;
            push_skip(cur, cout, hh, ww);
        }
    }
    // mid block
    tp.res.emplace_back();
    RC(ures_forward(r, u->res[ri], tp.res.back(), cur, hh, ww, tbias)); ++ri;
    cur = tp.res.back().out;
    tp.tf.emplace_back();
    RC(utf_forward(r, u->tf[ti], tp.tf.back(), cur, hh, ww, ctxp, T, Tp)); ++ti;
    cur = tp.tf.back().out;
    tp.res.emplace_back();
    RC(ures_forward(r, u->res[ri], tp.res.back(), cur, hh, ww, tbias)); ++ri;
    cur = tp.res.back().out;
    // up blocks
    int sp = (int)tp.skip_off.size();
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[nb - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) {
            --sp;
            if (sp < 0 || tp.skip_h[sp] != hh || tp.skip_w[sp] != ww) { set_error("unet: skip connection mismatch"); return -31; }
            CatRec cr;
            cr.ca = cin; cr.cb = tp.skip_c[sp]; cr.skip = sp; cr.h = hh; cr.w = ww;
            cr.out = r.sva.alloc(act_bytes(B, hh, ww, cr.ca + cr.cb));
            const long long rows = (long long)B * hh * ww;
            launch_copy_cols(r.S<bf16>(cur), cr.ca, 0, r.S<bf16>(cr.out), cr.ca + cr.cb, 0, cr.ca, rows, r.st);
            launch_copy_cols(r.S<bf16>(tp.skip_off[sp]), cr.cb, 0, r.S<bf16>(cr.out), cr.ca + cr.cb, cr.ca, cr.cb, rows, r.st);
            tp.cat.push_back(cr);
            tp.res.emplace_back();
            if (u->res[ri].ci != cr.ca + cr.cb) { set_error("unet: up resnet %zu expects %d input channels, got %d", ri, u->res[ri].ci, cr.ca + cr.cb); return -31; }
            RC(ures_forward(r, u->res[ri], tp.res.back(), cr.out, hh, ww, tbias)); ++ri;
            cur = tp.res.back().out;
            cin = cout;
            if (c.up_has_attn[i]) {
                tp.tf.emplace_back();
                RC(utf_forward(r, u->tf[ti], tp.tf.back(), cur, hh, ww, ctxp, T, Tp)); ++ti;
                cur = tp.tf.back().out;
            }
        }
        if (i != nb - 1) {
            SampRec s;
            s.h = hh; s.w = ww; s.x = cur;
            const Conv3& cv = u->ups[ui]; ++ui;
            const size_t m = r.wsa.mark();
            bf16* up = r.Walloc<bf16>(act_bytes(B, 2 * hh, 2 * ww, cv.ci));
            launch_upsample2x(r.S<bf16>(cur), up, B, hh, ww, cv.ci, r.st);
            hh *= 2; ww *= 2;
            s.out = r.sva.alloc(act_bytes(B, hh, ww, cv.co));
            RC(gemm_launch(dense_conv_op("unet.upsample", up, B, hh, ww, cv.ci, cv.fwd, cv.co, 1, hh, ww, cv.bias, nullptr,
                                         r.S<bf16>(s.out)), ns, r.st));
            r.wsa.reset(m);
            tp.up.push_back(s);
            cur = s.out;
        }
    }
    tp.xlast = cur;
    tp.gout = alloc_gn(r.sva, B, C0);
    {   // conv_norm_out + SiLU + conv_out -> fp32 NCHW
        const size_t m = r.wsa.mark();
        bf16* a = r.Walloc<bf16>(act_bytes(B, hh, ww, C0));
        RC(ugn_forward(r, r.S<bf16>(tp.xlast), u->norm_out, tp.gout, a, hh * ww, 1, 1e-5f));
        GemmOp o = dense_conv_op("unet.conv_out", a, B, hh, ww, C0, u->conv_out.fwd, 16, 1, hh, ww, u->conv_out.bias, nullptr, nullptr);
        o.D = out; o.out_fp32 = 1; o.n_store = c.out_channels;
        o.D_sB = (int64_t)c.out_channels * hh * ww; o.D_sH = ww; o.D_sW = 1; o.D_sN = (int64_t)hh * ww;
        RC(gemm_launch(o, ns, r.st));
        r.wsa.reset(m);
    }
    tp.saved_bytes = r.sva.peak + 256;
    tp.valid = true;
    return 0;
}

int unet_backward_walk(TmlUnet* u, URun& r, const Tape& tp, const float* dout, float* dsample) {
    const TmlUnetCfg& c = u->cfg;
    const int B = r.B, nb = c.num_blocks, ns = u->base.num_sms;
    const int C0 = c.block_out_channels[0];
    const int h = tp.h, w = tp.w, Tp = round_up(tp.T, 64);
    // ping-pong gradient buffers sized for the widest activation (the concatenated resnet inputs of the up path)
    size_t gmax = act_bytes(B, h, w, 64);
    {
        size_t k = 0;
        for (const UResnet& p : u->res) { gmax = std::max(gmax, act_bytes(B, tp.res[k].h, tp.res[k].w, std::max(p.ci, p.co))); ++k; }
    }
    bf16* G[2] = {r.Walloc<bf16>(gmax), r.Walloc<bf16>(gmax)};
    int cur = 0;
    // gradients of the skip connections, written by the up path, consumed by the down path
    std::vector<bf16*> dskip(tp.skip_off.size(), nullptr);
    for (size_t k = 0; k < tp.skip_off.size(); ++k)
        dskip[k] = r.Walloc<bf16>(act_bytes(B, tp.skip_h[k], tp.skip_w[k], tp.skip_c[k]));
    g_dump_next = 0;
    int hh = h, ww = w;
    {   // d(out) -> d(conv_norm_out input)
        const size_t m = r.wsa.mark();
        bf16* d64 = r.Walloc<bf16>(act_bytes(B, hh, ww, 64));
        launch_nchw_pack64(dout, d64, B, c.out_channels, hh * ww, r.st);
        bf16* d_a = r.Walloc<bf16>(act_bytes(B, hh, ww, C0));
        RC(gemm_launch(dense_conv_op("unet.conv_out.dgrad", d64, B, hh, ww, 64, u->conv_out.bwd, C0, 1, hh, ww, nullptr, nullptr,
                                     d_a), ns, r.st));
        RC(ugn_backward(r, r.S<bf16>(tp.xlast), d_a, u->norm_out, tp.gout, nullptr, G[cur], hh * ww, 1));
        r.wsa.reset(m);
        dump_grad(G[cur], act_bytes(B, hh, ww, C0), r.st);
    }
    size_t ri = tp.res.size(), ti = tp.tf.size(), ci = tp.cat.size(), ui = tp.up.size(), di = tp.down.size();
    for (int i = nb - 1; i >= 0; --i) {   // up blocks, last first
        if (i != nb - 1) {
            --ui;
            const SampRec& s = tp.up[ui];
            const Conv3& cv = u->ups[ui];
            const size_t m = r.wsa.mark();
            bf16* d_up = r.Walloc<bf16>(act_bytes(B, 2 * s.h, 2 * s.w, cv.ci));
            RC(gemm_launch(dense_conv_op("unet.upsample.dgrad", G[cur], B, 2 * s.h, 2 * s.w, cv.co, cv.bwd, cv.ci, 1, 2 * s.h,
                                         2 * s.w, nullptr, nullptr, d_up), ns, r.st));
            launch_upsample2x_bwd(d_up, G[cur ^ 1], B, s.h, s.w, cv.ci, r.st);
            r.wsa.reset(m);
            cur ^= 1;
            hh = s.h; ww = s.w;
            dump_grad(G[cur], act_bytes(B, hh, ww, cv.ci), r.st);
        }
        for (int j = c.layers_per_block; j >= 0; --j) {
            if (c.up_has_attn[i]) {
                --ti;
                RC(utf_backward(r, u->tf[ti], tp.tf[ti], G[cur], G[cur ^ 1], Tp));
                cur ^= 1;
                dump_grad(G[cur], act_bytes(B, hh, ww, u->tf[ti].C), r.st);
            }
            --ri; --ci;
            const CatRec& cr = tp.cat[ci];
            RC(ures_backward(r, u->res[ri], tp.res[ri], G[cur], G[cur ^ 1]));
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, hh, ww, cr.ca + cr.cb), r.st);
            const long long rows = (long long)B * hh * ww;
            launch_copy_cols(G[cur], cr.ca + cr.cb, cr.ca, dskip[cr.skip], cr.cb, 0, cr.cb, rows, r.st);
            launch_copy_cols(G[cur], cr.ca + cr.cb, 0, G[cur ^ 1], cr.ca, 0, cr.ca, rows, r.st);
            cur ^= 1;
        }
    }
    // mid block
    --ri; RC(ures_backward(r, u->res[ri], tp.res[ri], G[cur], G[cur ^ 1])); cur ^= 1;
    dump_grad(G[cur], act_bytes(B, hh, ww, u->res[ri].ci), r.st);
    --ti; RC(utf_backward(r, u->tf[ti], tp.tf[ti], G[cur], G[cur ^ 1], Tp)); cur ^= 1;
    dump_grad(G[cur], act_bytes(B, hh, ww, u->tf[ti].C), r.st);
    --ri; RC(ures_backward(r, u->res[ri], tp.res[ri], G[cur], G[cur ^ 1])); cur ^= 1;
    dump_grad(G[cur], act_bytes(B, hh, ww, u->res[ri].ci), r.st);
    // down blocks, last first
    int sp = (int)tp.skip_off.size();
    for (int i = nb - 1; i >= 0; --i) {
        if (i != nb - 1) {
            --sp;   // the downsampler's output was a skip connection
            launch_add_bf16(G[cur], dskip[sp], G[cur], (long long)B * hh * ww * tp.skip_c[sp], r.st);
            --di;
            const SampRec& d = tp.down[di];
            const Conv3& cv = u->downs[di];
            const int oh = d.h / 2, ow = d.w / 2;
            for (int q = 0; q < 4; ++q) {
                const int ph = q >> 1, pw = q & 1;
                GemmOp o = dense_conv_op("unet.downsample.dgrad", G[cur], B, oh, ow, cv.co, cv.bwd_par[q], cv.ci, 1, oh, ow,
                                         nullptr, nullptr, G[cur ^ 1] + ((size_t)ph * d.w + pw) * cv.ci);
                o.D_sW = 2 * cv.ci; o.D_sH = (int64_t)2 * d.w * cv.ci; o.D_sB = (int64_t)d.h * d.w * cv.ci;
                RC(gemm_launch(o, ns, r.st));
            }
            cur ^= 1;
            hh = d.h; ww = d.w;
            dump_grad(G[cur], act_bytes(B, hh, ww, cv.ci), r.st);
        }
        for (int j = c.layers_per_block - 1; j >= 0; --j) {
            --sp;
            launch_add_bf16(G[cur], dskip[sp], G[cur], (long long)B * hh * ww * tp.skip_c[sp], r.st);
            if (c.down_has_attn[i]) {
                --ti;
                RC(utf_backward(r, u->tf[ti], tp.tf[ti], G[cur], G[cur ^ 1], Tp));
                cur ^= 1;
                dump_grad(G[cur], act_bytes(B, hh, ww, u->tf[ti].C), r.st);
            }
            --ri;
            RC(ures_backward(r, u->res[ri], tp.res[ri], G[cur], G[cur ^ 1]));
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, hh, ww, u->res[ri].ci), r.st);
        }
    }
    --sp;   // conv_in's output
    launch_add_bf16(G[cur], dskip[sp], G[cur], (long long)B * hh * ww * C0, r.st);
    {   // conv_in dgrad (N padded 4 -> 64) -> fp32 NCHW
        bf16* d_x = r.Walloc<bf16>(act_bytes(B, h, w, 64));
        RC(gemm_launch(dense_conv_op("unet.conv_in.dgrad", G[cur], B, h, w, C0, u->conv_in.bwd, 64, 1, h, w, nullptr, nullptr,
                                     d_x), ns, r.st));
        launch_nhwc64_unpack(d_x, dsample, B, c.in_channels, h * w, r.st);
    }
    if (ri != 0 || ti != 0 || ci != 0 || ui != 0 || di != 0 || sp != 0) { set_error("internal: unet backward walk out of step"); return -41; }
    return 0;
}

// tape + scratch size for a shape: a dry run (launchers disabled) of both walks, as in enc_layout
int unet_layout(TmlUnet* u, int B, int h, int w, int T) {
    if (B <= 0 || h <= 0 || w <= 0 || T <= 0) { set_error("bad shape B=%d h=%d w=%d T=%d", B, h, w, T); return -30; }
    if (w % 8) { set_error("unet: latent width must be a multiple of 8 (got %d)", w); return -30; }
    Tape& tp = u->tape;
    if (tp.valid && tp.B == B && tp.h == h && tp.w == w && tp.T == T && tp.ws_bytes) return 0;
    char* fake = reinterpret_cast<char*>(uintptr_t(1) << 20);   // never dereferenced
    size_t peak = 0;
    int rc = 0;
    g_dry_run = true;
    g_dry_validate = true;
    Tape t2;
    {
        URun r{u, fake, fake, Arena(), Arena(), nullptr, B};
        rc = unet_forward_walk(u, r, t2, reinterpret_cast<const float*>(fake), 0.f, reinterpret_cast<const float*>(fake), h, w, T,
                               reinterpret_cast<float*>(fake));
        peak = r.wsa.peak;
    }
    if (rc == 0) {
        URun r{u, fake, fake, Arena(), Arena(), nullptr, B};
        rc = unet_backward_walk(u, r, t2, reinterpret_cast<const float*>(fake), reinterpret_cast<float*>(fake));
        peak = std::max(peak, r.wsa.peak);
    }
    g_dry_run = false;
    g_dry_validate = false;
    if (rc) { tp = Tape(); return rc; }
    t2.ws_bytes = peak + 1024;
    tp = t2;
    return 0;
}

__global__ void ctx_pack_kernel(const float* __restrict__ ctx, bf16* __restrict__ out, int T, int Tp, int D, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int dd = (int)(i % D);
    const long long r = i / D;
    const int t = (int)(r % Tp);
    const long long b = r / Tp;
    out[i] = __float2bfloat16_rn(t < T ? ctx[((size_t)b * T + t) * D + dd] : 0.f);
}
void ctx_pack(URun& r, const float* ctx, bf16* out, int B, int T, int Tp, int D) {
    if (g_dry_run) return;
    const long long total = (long long)B * Tp * D;
    ctx_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, r.st>>>(ctx, out, T, Tp, D, total);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int tml_unet_create(const TmlUnetCfg* cfg, int device, TmlUnet** out) {
    if (!cfg || !out) { set_error("null argument"); return -1; }
    if (cfg->in_channels < 1 || cfg->in_channels > 64 || cfg->out_channels < 1 || cfg->out_channels > 16 ||
        cfg->norm_num_groups != 32 || cfg->num_blocks < 2 || cfg->num_blocks > 8 || cfg->layers_per_block < 1 ||
        cfg->num_heads < 1 || cfg->cross_attention_dim % 64) {
        set_error("unsupported unet config (need groups=32, cross_attention_dim %% 64 == 0, 2..8 blocks)");
        return -1;
    }
    for (int i = 0; i < cfg->num_blocks; ++i)
        if (cfg->block_out_channels[i] % 64) { set_error("unet block_out_channels must be multiples of 64"); return -1; }
    if (cfg->block_out_channels[0] * 4 > 2048 * 4) { set_error("unet too wide"); return -1; }
    int sms = 148;
    if (!g_host_only) {
        int ndev = 0;
        CUDA_OK(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return -1; }
        DeviceGuard guard(device);
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            set_error("tml_b200 needs an sm_100a device (Blackwell B200); device %d is sm_%d%d — there is no fallback path",
                      device, prop.major, prop.minor);
            return -2;
        }
        sms = prop.multiProcessorCount;
    }
    TmlUnet* u = new TmlUnet();
    u->cfg = *cfg;
    u->base.device = device;
    u->base.num_sms = sms;
    *out = u;
    return 0;
}

void tml_unet_destroy(TmlUnet* u) {
    if (!u) return;
    DeviceGuard guard(u->base.device);
    for (void* p : u->base.dev_allocs) cudaFree(p);
    delete u;
}

int tml_unet_set_weight(TmlUnet* u, const char* key, const void* ptr, int dtype, const int64_t* shape, int ndim) {
    if (!u) { set_error("null handle"); return -1; }
    u->finalized = false;
    return tml_encoder_set_weight(&u->base, key, ptr, dtype, shape, ndim);
}

int tml_unet_finalize(TmlUnet* u, void* stream) {
    (void)stream;
    if (!u) { set_error("null handle"); return -1; }
    TmlEncoder* e = &u->base;
    DeviceGuard guard(e->device);
    const TmlUnetCfg& c = u->cfg;
    const int nb = c.num_blocks, C0 = c.block_out_channels[0];
    u->temb = 4 * C0;
    {   // conv_in: input channels padded to 64; conv_out: N padded to 16, dgrad consumes d(out) padded to 64 channels
        const HostTensor* w = find(e, "conv_in.weight", (size_t)C0 * c.in_channels * 9);
        const HostTensor* b = find(e, "conv_in.bias", C0);
        if (!w || !b) return -20;
        std::vector<float> w64((size_t)C0 * 64 * 9, 0.f);
        for (int co = 0; co < C0; ++co)
            for (int ci = 0; ci < c.in_channels; ++ci)
                for (int k = 0; k < 9; ++k) w64[((size_t)co * 64 + ci) * 9 + k] = w->v[((size_t)co * c.in_channels + ci) * 9 + k];
        u->conv_in.ci = 64; u->conv_in.co = C0; u->conv_in.stride = 1;
        RC(make_packed(e, w64.data(), C0, 64, 0, &u->conv_in.fwd));
        RC(make_packed(e, w64.data(), C0, 64, 1, &u->conv_in.bwd));
        RC(upload<float>(e, b->v, &u->conv_in.bias));
        const HostTensor* wo = find(e, "conv_out.weight", (size_t)c.out_channels * C0 * 9);
        const HostTensor* bo = find(e, "conv_out.bias", c.out_channels);
        if (!wo || !bo) return -20;
        std::vector<float> w16((size_t)16 * C0 * 9, 0.f), b16(16, 0.f), wo64((size_t)64 * C0 * 9, 0.f);
        memcpy(w16.data(), wo->v.data(), (size_t)c.out_channels * C0 * 9 * 4);
        memcpy(wo64.data(), wo->v.data(), (size_t)c.out_channels * C0 * 9 * 4);
        for (int i = 0; i < c.out_channels; ++i) b16[i] = bo->v[i];
        u->conv_out.ci = C0; u->conv_out.co = 16; u->conv_out.stride = 1;
        RC(make_packed(e, w16.data(), 16, C0, 0, &u->conv_out.fwd));
        RC(make_packed(e, wo64.data(), 64, C0, 1, &u->conv_out.bwd));
        RC(upload<float>(e, b16, &u->conv_out.bias));
    }
    {   // time embedding MLP (fp32)
        const HostTensor* w1 = find(e, "time_embedding.linear_1.weight", (size_t)u->temb * C0);
        const HostTensor* b1 = find(e, "time_embedding.linear_1.bias", u->temb);
        const HostTensor* w2 = find(e, "time_embedding.linear_2.weight", (size_t)u->temb * u->temb);
        const HostTensor* b2 = find(e, "time_embedding.linear_2.bias", u->temb);
        if (!w1 || !b1 || !w2 || !b2) return -20;
        RC(upload<float>(e, w1->v, &u->te_w1)); RC(upload<float>(e, b1->v, &u->te_b1));
        RC(upload<float>(e, w2->v, &u->te_w2)); RC(upload<float>(e, b2->v, &u->te_b2));
    }
    u->res.clear(); u->tf.clear(); u->downs.clear(); u->ups.clear();
    std::vector<float> tp_w, tp_b;
    char key[256];
    std::vector<int> skip_c;
    int cin = C0;
    skip_c.push_back(C0);
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[i];
        for (int j = 0; j < c.layers_per_block; ++j) {
            snprintf(key, sizeof(key), "down_blocks.%d.resnets.%d", i, j);
            UResnet r;
            RC(make_uresnet(u, key, cin, cout, tp_w, tp_b, &r));
            u->res.push_back(r);
            cin = cout;
            if (c.down_has_attn[i]) {
                snprintf(key, sizeof(key), "down_blocks.%d.attentions.%d", i, j);
                UTf t;
                RC(make_utf(u, key, cout, &t));
                u->tf.push_back(t);
            }
            skip_c.push_back(cout);
        }
        if (i != nb - 1) {
            snprintf(key, sizeof(key), "down_blocks.%d.downsamplers.0.conv", i);
            Conv3 d;
            RC(make_down(u, key, cout, &d));
            u->downs.push_back(d);
            skip_c.push_back(cout);
        }
    }
    {
        UResnet r0, r1;
        UTf t;
        RC(make_uresnet(u, "mid_block.resnets.0", cin, cin, tp_w, tp_b, &r0));
        RC(make_utf(u, "mid_block.attentions.0", cin, &t));
        RC(make_uresnet(u, "mid_block.resnets.1", cin, cin, tp_w, tp_b, &r1));
        u->res.push_back(r0); u->tf.push_back(t); u->res.push_back(r1);
    }
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[nb - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) {
            const int cs = skip_c.back();
            skip_c.pop_back();
            snprintf(key, sizeof(key), "up_blocks.%d.resnets.%d", i, j);
            UResnet r;
            RC(make_uresnet(u, key, cin + cs, cout, tp_w, tp_b, &r));
            u->res.push_back(r);
            cin = cout;
            if (c.up_has_attn[i]) {
                snprintf(key, sizeof(key), "up_blocks.%d.attentions.%d", i, j);
                UTf t;
                RC(make_utf(u, key, cout, &t));
                u->tf.push_back(t);
            }
        }
        if (i != nb - 1) {
            snprintf(key, sizeof(key), "up_blocks.%d.upsamplers.0.conv", i);
            Conv3 up;
            RC(make_conv3(e, key, cout, cout, 1, &up));
            u->ups.push_back(up);
        }
    }
    RC(make_norm(e, "conv_norm_out", C0, &u->norm_out));
    u->tb_total = (int)tp_b.size();
    RC(upload<float>(e, tp_w, &u->tp_w));
    RC(upload<float>(e, tp_b, &u->tp_b));
    e->host.clear();
    u->finalized = true;
    u->tape = Tape();
    return 0;
}

int tml_unet_query(TmlUnet* u, int B, int h, int w, int ctx_tokens, size_t* workspace_bytes, size_t* saved_bytes) {
    if (!u || !u->finalized) { set_error("unet not finalized"); return -1; }
    RC(unet_layout(u, B, h, w, ctx_tokens));
    if (workspace_bytes) *workspace_bytes = u->tape.ws_bytes;
    if (saved_bytes) *saved_bytes = u->tape.saved_bytes;
    return 0;
}

int tml_unet_forward(TmlUnet* u, const float* sample, float timestep, const float* ctx, int B, int h, int w, int ctx_tokens,
                     float* out, void* saved, void* ws, void* stream) {
    if (!u || !u->finalized) { set_error("unet not finalized"); return -1; }
    if (!sample || !ctx || !out || !saved || !ws) { set_error("null buffer"); return -1; }
    DeviceGuard guard(u->base.device);
    RC(unet_layout(u, B, h, w, ctx_tokens));
    URun r{u, reinterpret_cast<char*>(saved), reinterpret_cast<char*>(ws), Arena(), Arena(), reinterpret_cast<cudaStream_t>(stream), B};
    Tape t2;
    RC(unet_forward_walk(u, r, t2, sample, timestep, ctx, h, w, ctx_tokens, out));
    if (r.wsa.peak > u->tape.ws_bytes || t2.saved_bytes != u->tape.saved_bytes) {
        set_error("internal: unet workspace overrun (%zu > %zu) or layout drift", r.wsa.peak, u->tape.ws_bytes);
        return -40;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_unet_backward(TmlUnet* u, const float* dout, int B, int h, int w, int ctx_tokens, const void* saved, float* dsample,
                      void* ws, void* stream) {
    if (!u || !u->finalized) { set_error("unet not finalized"); return -1; }
    if (!dout || !saved || !ws || !dsample) { set_error("null buffer"); return -1; }
    DeviceGuard guard(u->base.device);
    RC(unet_layout(u, B, h, w, ctx_tokens));
    URun r{u, const_cast<char*>(reinterpret_cast<const char*>(saved)), reinterpret_cast<char*>(ws), Arena(), Arena(),
           reinterpret_cast<cudaStream_t>(stream), B};
    RC(unet_backward_walk(u, r, u->tape, dout, dsample));
    if (r.wsa.peak > u->tape.ws_bytes) { set_error("internal: unet workspace overrun (%zu > %zu)", r.wsa.peak, u->tape.ws_bytes); return -40; }
    CUDA_OK(cudaGetLastError());
    return 0;
}

void tml_debug_set_host_only(int on) { g_host_only = on != 0; }

int tml_debug_attention(const void* Q, const void* K, const void* V, int nb, int tq, int tkv, int dp, int lcol, float scale,
                        void* O, float* rmax, float* inv_l, const void* dO, void* dQ, void* dK, void* dV, void* ws,
                        size_t ws_bytes, void* stream) {
    if (!Q || !K || !V || !O || !rmax || !inv_l) { set_error("null buffer"); return -1; }
    if (!attn_fused_supported(tq, tkv, dp)) { set_error("unsupported shape tq=%d tkv=%d dp=%d", tq, tkv, dp); return -1; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    RC(launch_attn_fused_fwd(reinterpret_cast<const bf16*>(Q), reinterpret_cast<const bf16*>(K), reinterpret_cast<const bf16*>(V),
                             nullptr, rmax, inv_l, reinterpret_cast<bf16*>(O), nb, tq, tkv, dp, lcol, scale, st));
    if (dO != nullptr) {
        const size_t rows = (size_t)nb * tq;
        const size_t need = rows * dp * sizeof(bf16) + 256 + (rows * 3 + 8) * sizeof(float);
        if (!ws || ws_bytes < need || !dQ) { set_error("attention backward needs %zu bytes of scratch", need); return -1; }
        bf16* dOs = reinterpret_cast<bf16*>(ws);
        float* Dp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((rows * dp * sizeof(bf16) + 255) & ~size_t(255)));
        RC(launch_attn_fused_bwd(reinterpret_cast<const bf16*>(Q), reinterpret_cast<const bf16*>(K), reinterpret_cast<const bf16*>(V),
                                 reinterpret_cast<const bf16*>(O), reinterpret_cast<const bf16*>(dO), rmax, inv_l, dOs, Dp,
                                 reinterpret_cast<bf16*>(dQ), reinterpret_cast<bf16*>(dK), reinterpret_cast<bf16*>(dV), nb, tq, tkv,
                                 dp, scale, st));
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_debug_unet_saved_tensor(TmlUnet* u, const char* name, int index, size_t* offset, int dims[4]) {
    if (!u || !name || !offset || !dims) { set_error("null argument"); return -1; }
    const Tape& tp = u->tape;
    if (!tp.valid) { set_error("no unet layout yet: call tml_unet_query/forward first"); return -1; }
    const std::string n(name);
    auto set = [&](size_t off, int h, int w, int c) { *offset = off; dims[0] = tp.B; dims[1] = h; dims[2] = w; dims[3] = c; return 0; };
    if (n == "conv_in") return set(tp.x0, tp.h, tp.w, u->cfg.block_out_channels[0]);
    if (n == "resnet_h1" || n == "resnet_out") {
        if (index < 0 || index >= (int)tp.res.size()) { set_error("index"); return -1; }
        const ResRec& r = tp.res[index];
        return set(n == "resnet_h1" ? r.h1 : r.out, r.h, r.w, u->res[index].co);
    }
    if (n == "tf_t0" || n == "tf_x1" || n == "tf_x2" || n == "tf_out") {
        if (index < 0 || index >= (int)tp.tf.size()) { set_error("index"); return -1; }
        const TfRec& t = tp.tf[index];
        const size_t off = n == "tf_t0" ? t.t0 : n == "tf_x1" ? t.x1 : n == "tf_x2" ? t.x2 : t.out;
        return set(off, t.h, t.w, u->tf[index].C);
    }
    if (n == "down_out") {
        if (index < 0 || index >= (int)tp.down.size()) { set_error("index"); return -1; }
        return set(tp.down[index].out, tp.down[index].h / 2, tp.down[index].w / 2, u->downs[index].co);
    }
    if (n == "up_out") {
        if (index < 0 || index >= (int)tp.up.size()) { set_error("index"); return -1; }
        return set(tp.up[index].out, 2 * tp.up[index].h, 2 * tp.up[index].w, u->ups[index].co);
    }
    if (n == "count_resnets") { *offset = tp.res.size(); return 0; }
    if (n == "count_tf") { *offset = tp.tf.size(); return 0; }
    set_error("unknown tensor '%s'", name);
    return -1;
}

}  // extern "C"
