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
#include "vae_impl.cuh"

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
    DeviceGuard guard(device);   // the caller's current device is restored on return
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
    DeviceGuard guard(e->device);
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
    DeviceGuard guard(e->device);
    const TmlEncoderCfg& c = e->cfg;
    const int C0 = c.block_out_channels[0];
    {   // conv_in forward as a 3x3 convolution over a 64-channel image [hi(3) | lo(3) | 0(58)]: both halves see the same weights
        if (c.in_channels != 3) { set_error("in_channels must be 3 (got %d)", c.in_channels); return -22; }
        const HostTensor* w = find(e, "encoder.conv_in.weight", (size_t)C0 * 27);
        const HostTensor* b = find(e, "encoder.conv_in.bias", C0);
        if (!w || !b) return -20;
        std::vector<float> w64((size_t)C0 * 64 * 9, 0.f);
        for (int co = 0; co < C0; ++co)
            for (int ci = 0; ci < 3; ++ci)
                for (int k = 0; k < 9; ++k)
                    w64[((size_t)co * 64 + ci) * 9 + k] = w64[((size_t)co * 64 + ci + 3) * 9 + k] = w->v[((size_t)co * 3 + ci) * 9 + k];
        Conv3& ci3 = e->conv_in_fwd;
        ci3.ci = 64; ci3.co = C0; ci3.stride = 1;
        RC(make_packed(e, w64.data(), C0, 64, 0, &ci3.fwd));
        RC(upload<float>(e, b->v, &ci3.bias));
        // Input gradient: the output is only 3 channels wide, so the 3x3 dgrad as a K = 9*C0 GEMM with N padded to 16
        // spends 144 tiny MMAs per 256 pixels.  Instead ONE K = C0 GEMM produces, per pixel, the 27 products
        // Y[p][(r,s,ci)] = sum_co dy[p][co] * W[co][ci][r][s]  (N = 27 padded to 32, 8 MMAs per 128 pixels) into fp32
        // planes, and a col2im kernel gathers dx[ci][h][w] = sum_{r,s} Y[(h-r+1, w-s+1)][(r,s,ci)].
        std::vector<float> wt((size_t)32 * C0, 0.f), zb(32, 0.f);
        for (int co = 0; co < C0; ++co)
            for (int ci = 0; ci < 3; ++ci)
                for (int k = 0; k < 9; ++k) wt[(size_t)(k * 3 + ci) * C0 + co] = w->v[((size_t)co * 3 + ci) * 9 + k];
        RC(make_lin_from(e, wt, zb, C0, 32, &e->conv_in_bwd));
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
    if (e->host.count("decoder.conv_in.weight")) RC(decoder_finalize(e));
    e->host.clear();
    e->finalized = true;
    e->lay = Layout();
    return 0;
}

static int enc_forward_walk(TmlEncoder* e, Run& r, const float* x, float* moments);
static int enc_backward_walk(TmlEncoder* e, Run& r, const float* dmoments, float* dx, float beta);

// build_layout + the scratch size: a dry run (launchers disabled) of both walks with the same arena gives the exact
// peak, so the size reported by tml_encoder_query is what the real walks use -- nothing is launched on an estimate.
static int enc_layout(TmlEncoder* e, int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) { set_error("bad shape B=%d H=%d W=%d", B, H, W); return -30; }
    RC(build_layout(e, B, H, W));
    Layout& L = e->lay;
    if (L.measured) return 0;
    char* fake = reinterpret_cast<char*>(uintptr_t(1) << 20);   // never dereferenced
    size_t peak = 0;
    int rc = 0;
    g_dry_run = true;
    {
        Run r{e, fake, fake, Arena(), nullptr, B};
        rc = enc_forward_walk(e, r, reinterpret_cast<const float*>(fake), reinterpret_cast<float*>(fake));
        peak = r.wsa.peak;
    }
    if (rc == 0) {
        Run r{e, fake, fake, Arena(), nullptr, B};
        rc = enc_backward_walk(e, r, reinterpret_cast<const float*>(fake), reinterpret_cast<float*>(fake), 0.f);
        peak = std::max(peak, r.wsa.peak);
    }
    g_dry_run = false;
    if (rc) { L = Layout(); return rc; }
    L.ws_bytes = peak + 1024;
    L.measured = true;
    return 0;
}

int tml_encoder_query(TmlEncoder* e, int B, int H, int W, size_t* workspace_bytes, size_t* saved_bytes) {
    if (!e) { set_error("null handle"); return -1; }
    RC(enc_layout(e, B, H, W));
    if (workspace_bytes) *workspace_bytes = e->lay.ws_bytes;
    if (saved_bytes) *saved_bytes = e->lay.saved_bytes;
    return 0;
}

int tml_encoder_forward(TmlEncoder* e, const float* x, int B, int H, int W, float* moments, void* saved, void* ws,
                        void* stream) {
    if (!e || !e->finalized) { set_error("encoder not finalized"); return -1; }
    if (!x || !moments || !saved || !ws) { set_error("null buffer"); return -1; }
    DeviceGuard guard(e->device);
    RC(enc_layout(e, B, H, W));
    Run r{e, reinterpret_cast<char*>(saved), reinterpret_cast<char*>(ws), Arena(), reinterpret_cast<cudaStream_t>(stream), B};
    RC(enc_forward_walk(e, r, x, moments));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int enc_forward_walk(TmlEncoder* e, Run& r, const float* x, float* moments) {
    const Layout& L = e->lay;
    const int B = L.B, H = L.H, W = L.W;
    const int C0 = e->cfg.block_out_channels[0];
    r.statbuf[0] = r.Walloc<float>(fused_partial_bytes(B, H, W));
    r.statbuf[1] = r.Walloc<float>(fused_partial_bytes(B, H, W));
    {   // conv_in: pack (hi/lo bf16 split of the fp32 image, 64 channels) + 3x3 convolution with the statistics fused
        const size_t m = r.wsa.mark();
        bf16* xp = r.Walloc<bf16>(act_bytes(B, H, W, 64));
        launch_conv_in_pack(x, xp, B, H, W, r.st);
        GemmOp ci = dense_conv_op("conv_in", xp, B, H, W, 64, e->conv_in_fwd.fwd, C0, 1, H, W, e->conv_in_fwd.bias, nullptr,
                                  r.S<bf16>(L.x0));
        r.pending = fuse_stats(ci, r.statbuf[1], H, W);
        RC(gemm_launch(ci, e->num_sms, r.st));
        r.wsa.reset(m);
    }
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
    if (L.measured && r.wsa.peak > L.ws_bytes) { set_error("internal: workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
    return 0;
}

int tml_encoder_backward(TmlEncoder* e, const float* dmoments, int B, int H, int W, const void* saved, float* dx,
                         float beta, void* ws, void* stream) {
    if (!e || !e->finalized) { set_error("encoder not finalized"); return -1; }
    if (!dmoments || !saved || !ws || !dx) { set_error("null buffer"); return -1; }
    DeviceGuard guard(e->device);
    RC(enc_layout(e, B, H, W));
    Run r{e, const_cast<char*>(reinterpret_cast<const char*>(saved)), reinterpret_cast<char*>(ws), Arena(),
          reinterpret_cast<cudaStream_t>(stream), B};
    RC(enc_backward_walk(e, r, dmoments, dx, beta));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int enc_backward_walk(TmlEncoder* e, Run& r, const float* dmoments, float* dx, float beta) {
    const Layout& L = e->lay;
    const int B = L.B, H = L.H, W = L.W;
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
        const size_t m = r.wsa.mark();
        float* Y = r.Walloc<float>((size_t)B * H * W * 32 * sizeof(float));   // [B][H][W][32] tap products (27 used)
        GemmOp o = dense_lin_op("conv_in.dgrad", G[cur], B, H, W, C0, e->conv_in_bwd.fwd, 32, nullptr, nullptr, nullptr);
        o.D = Y; o.out_fp32 = 1;   // dense rows of 32 floats = one 128-byte line per pixel
        RC(gemm_launch(o, e->num_sms, r.st));
        launch_conv_in_col2im(Y, dx, B, H, W, beta, r.st);
        r.wsa.reset(m);
    }
    if (L.measured && r.wsa.peak > L.ws_bytes) { set_error("internal: workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
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

int tml_universal_project(float* delta, const float* sources, int nsrc, float lo, float hi, int64_t n, void* stream) {
    if (!delta || (!sources && nsrc > 0)) { set_error("null buffer"); return -1; }
    if (nsrc <= 0 || n <= 0) return 0;
    launch_universal_project(delta, sources, nsrc, lo, hi, n, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

void tml_launch_counts(int64_t out[2]) {
    out[0] = gemm_launch_count();
    out[1] = kernel_launch_count();
}
void tml_debug_set_gemm_impl(int impl) { gemm_set_impl(impl); }

static int desc_to_op(const TmlGemmDesc* d, GemmOp& o) {
    if (!d) { set_error("null desc"); return -1; }
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
    o.in_gn_ss = reinterpret_cast<const float2*>(d->in_gn_ss);
    o.a_trans = d->a_trans; o.A_sK = d->A_sK;
    return 0;
}

int tml_debug_gn_chunks_per_image(const TmlGemmDesc* d) {
    GemmOp o;
    if (desc_to_op(d, o)) return -1;
    return gemm_gn_chunks_per_image(o);
}

int tml_debug_gemm(const TmlGemmDesc* d, void* stream) {
    GemmOp o;
    RC(desc_to_op(d, o));
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
int tml_debug_unet_gn_chunks_per_image(int HW, int C) { return (HW > 0 && C > 0 && C % 32 == 0 && C % 8 == 0) ? gng_num_chunks(HW, C) : -1; }

int tml_debug_pack_conv3x3(const float* w, int Co, int Ci, int mode, uint16_t* out, int* ntaps, int* dh, int* dw) {
    std::vector<uint16_t> h;
    if (pack_conv3x3(w, Co, Ci, mode, h, ntaps, dh, dw)) { set_error("bad pack mode %d", mode); return -1; }
    memcpy(out, h.data(), h.size() * 2);
    return 0;
}

}  // extern "C"
