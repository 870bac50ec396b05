// AutoencoderKL decoder forward + input-gradient backward (vae.decode at main.py:156 and the image-space
// losses of main.py:160,168) built from the same sm_100a kernels as the encoder: the tcgen05 implicit GEMM
// for every conv / linear / attention product, the GroupNorm / SiLU kernels, plus nearest-2x upsampling.
//
//   z [B,4,h,w] fp32 -> post_quant_conv (1x1) -> conv_in 4->C -> mid (res, attn, res)
//     -> up blocks: 3 resnets each (C: 512,512,256,128), nearest-2x + conv3x3 between blocks
//     -> GroupNorm + SiLU -> conv_out C0->3 -> image [B,3,8h,8w] fp32 NCHW
//
// Layout rules are the encoder's: bf16 NHWC activations, fp32 NCHW at the seams,
// `saved` keeps the input of every GroupNorm, qkv and the attention probabilities.
#include "vae_impl.cuh"

// ------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------
int decoder_finalize(TmlEncoder* e) {
    const TmlEncoderCfg& c = e->cfg;
    const int nb = c.num_blocks;
    const int Ctop = c.block_out_channels[nb - 1];
    const int C0 = c.block_out_channels[0];
    const int L = c.latent_channels;
    if (L != 4) { set_error("decoder: latent_channels must be 4"); return -1; }
    {   // post_quant_conv: [4][4][1][1]
        const HostTensor* w = find(e, "post_quant_conv.weight", (size_t)L * L);
        const HostTensor* b = find(e, "post_quant_conv.bias", L);
        if (!w || !b) return -20;
        RC(upload<float>(e, w->v, &e->pq_w));
        RC(upload<float>(e, b->v, &e->pq_b));
    }
    {   // conv_in: [Ctop][4][3][3] -> input channels padded to 64 (one 128-byte swizzle row)
        const HostTensor* w = find(e, "decoder.conv_in.weight", (size_t)Ctop * L * 9);
        const HostTensor* b = find(e, "decoder.conv_in.bias", Ctop);
        if (!w || !b) return -20;
        std::vector<float> w64((size_t)Ctop * 64 * 9, 0.f);
        for (int co = 0; co < Ctop; ++co)
            for (int ci = 0; ci < L; ++ci)
                for (int k = 0; k < 9; ++k) w64[((size_t)co * 64 + ci) * 9 + k] = w->v[((size_t)co * L + ci) * 9 + k];
        Conv3& ci = e->dec_conv_in;
        ci.ci = 64; ci.co = Ctop; ci.stride = 1;
        RC(make_packed(e, w64.data(), Ctop, 64, 0, &ci.fwd));
        RC(make_packed(e, w64.data(), Ctop, 64, 1, &ci.bwd));
        RC(upload<float>(e, b->v, &ci.bias));
    }
    e->dec_resnets.clear();
    e->ups.clear();
    char key[256];
    for (int m = 0; m < 2; ++m) {
        snprintf(key, sizeof(key), "decoder.mid_block.resnets.%d", m);
        Resnet r;
        RC(make_resnet(e, key, Ctop, Ctop, &r));
        e->dec_resnets.push_back(r);
    }
    e->dec_has_attn = c.mid_block_add_attention != 0;
    if (e->dec_has_attn) {
        const std::string a = "decoder.mid_block.attentions.0";
        Attn& at = e->dec_attn;
        at.C = Ctop;
        RC(make_norm(e, a + ".group_norm", Ctop, &at.gn));
        const char* names[3] = {".to_q", ".to_k", ".to_v"};
        std::vector<float> Wqkv((size_t)3 * Ctop * Ctop), bqkv((size_t)3 * Ctop);
        for (int q = 0; q < 3; ++q) {
            const HostTensor* w = find(e, a + names[q] + ".weight", (size_t)Ctop * Ctop);
            const HostTensor* b = find(e, a + names[q] + ".bias", Ctop);
            if (!w || !b) return -20;
            memcpy(&Wqkv[(size_t)q * Ctop * Ctop], w->v.data(), (size_t)Ctop * Ctop * 4);
            memcpy(&bqkv[(size_t)q * Ctop], b->v.data(), (size_t)Ctop * 4);
        }
        RC(make_lin_from(e, Wqkv, bqkv, Ctop, 3 * Ctop, &at.qkv));
        const HostTensor* wo = find(e, a + ".to_out.0.weight", (size_t)Ctop * Ctop);
        const HostTensor* bo = find(e, a + ".to_out.0.bias", Ctop);
        if (!wo || !bo) return -20;
        RC(make_lin_from(e, wo->v, bo->v, Ctop, Ctop, &at.out));
    }
    int cin = Ctop;
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[nb - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) {
            snprintf(key, sizeof(key), "decoder.up_blocks.%d.resnets.%d", i, j);
            Resnet r;
            RC(make_resnet(e, key, cin, cout, &r));
            e->dec_resnets.push_back(r);
            cin = cout;
        }
        if (i != nb - 1) {
            snprintf(key, sizeof(key), "decoder.up_blocks.%d.upsamplers.0.conv", i);
            Conv3 u;
            RC(make_conv3(e, key, cout, cout, 1, &u));
            e->ups.push_back(u);
        }
    }
    RC(make_norm(e, "decoder.conv_norm_out", C0, &e->dec_norm_out));
    {   // conv_out: [3][C0][3][3] -> N padded to 16; dgrad consumes d(image) padded to 64 channels
        const HostTensor* w = find(e, "decoder.conv_out.weight", (size_t)3 * C0 * 9);
        const HostTensor* b = find(e, "decoder.conv_out.bias", 3);
        if (!w || !b) return -20;
        std::vector<float> w16((size_t)16 * C0 * 9, 0.f), b16(16, 0.f), w64((size_t)64 * C0 * 9, 0.f);
        memcpy(w16.data(), w->v.data(), (size_t)3 * C0 * 9 * 4);
        memcpy(w64.data(), w->v.data(), (size_t)3 * C0 * 9 * 4);
        for (int i = 0; i < 3; ++i) b16[i] = b->v[i];
        Conv3& co = e->dec_conv_out;
        co.ci = C0; co.co = 16; co.stride = 1;
        RC(make_packed(e, w16.data(), 16, C0, 0, &co.fwd));
        RC(make_packed(e, w64.data(), 64, C0, 1, &co.bwd));   // B[ci][t*64 + m] = W[m][ci][r][s]
        RC(upload<float>(e, b16, &co.bias));
    }
    e->has_decoder = true;
    e->dlay = DecLayout();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
static int build_dec_layout(TmlEncoder* e, int B, int h, int w) {
    DecLayout& L = e->dlay;
    if (L.B == B && L.h == h && L.w == w && L.saved_bytes) return 0;
    const TmlEncoderCfg& c = e->cfg;
    const int nb = c.num_blocks;
    if (w % 8) { set_error("decoder: latent width must be a multiple of 8 (got %d)", w); return -30; }
    L = DecLayout();
    L.B = B; L.h = h; L.w = w;
    Arena S;
    const int Ctop = c.block_out_channels[nb - 1];
    // (the scratch size is measured by a dry run of the walks: dec_layout)
    int hh = h, ww = w;
    L.x0 = S.alloc(act_bytes(B, hh, ww, Ctop));
    size_t cur = L.x0;
    auto add_resnet = [&](int ci, int co) {
        ResnetRec r;
        r.h = hh; r.w = ww; r.x = cur;
        r.g1 = alloc_gn(S, B, ci);
        r.h1 = S.alloc(act_bytes(B, hh, ww, co));
        r.g2 = alloc_gn(S, B, co);
        r.out = S.alloc(act_bytes(B, hh, ww, co));
        L.res.push_back(r);
        cur = r.out;
    };
    add_resnet(Ctop, Ctop);
    if (c.mid_block_add_attention) {
        AttnRec a;
        a.h = hh; a.w = ww; a.x = cur;
        const size_t tok = (size_t)hh * ww;
        a.g = alloc_gn(S, B, Ctop);
        a.qkv = S.alloc((size_t)B * tok * 3 * Ctop * sizeof(bf16));
        a.P = S.alloc((size_t)B * tok * tok * sizeof(bf16));
        a.a = S.alloc(act_bytes(B, hh, ww, Ctop));
        a.inv_l = S.alloc((size_t)B * tok * sizeof(float));
        a.out = S.alloc(act_bytes(B, hh, ww, Ctop));
        // (scratch of the attention walks is measured by the dry run, see dec_layout)
        L.attn = a;
        cur = a.out;
    }
    add_resnet(Ctop, Ctop);
    int cin = Ctop;
    for (int i = 0; i < nb; ++i) {
        const int cout = c.block_out_channels[nb - 1 - i];
        for (int j = 0; j < c.layers_per_block + 1; ++j) { add_resnet(cin, cout); cin = cout; }
        if (i != nb - 1) {
            UpRec u;
            u.h = hh; u.w = ww;
            hh *= 2; ww *= 2;
            u.out = S.alloc(act_bytes(B, hh, ww, cout));
            L.ups.push_back(u);
            cur = u.out;
        }
    }
    L.gout = alloc_gn(S, B, cin);
    L.xlast = cur;
    L.Hl = hh; L.Wl = ww;
    L.saved_bytes = S.peak + 256;
    L.ws_bytes = 0;   // set by dec_layout from the dry run
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

static int dec_forward_walk(TmlEncoder* e, Run& r, const float* z, float* image);
static int dec_backward_walk(TmlEncoder* e, Run& r, const float* dimage, float* dz);

// layout + scratch size measured by a dry run of both walks (see enc_layout in encoder.cu)
static int dec_layout(TmlEncoder* e, int B, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0) { set_error("bad shape B=%d h=%d w=%d", B, h, w); return -30; }
    RC(build_dec_layout(e, B, h, w));
    DecLayout& L = e->dlay;
    if (L.measured) return 0;
    char* fake = reinterpret_cast<char*>(uintptr_t(1) << 20);   // never dereferenced
    size_t peak = 0;
    int rc = 0;
    g_dry_run = true;
    {
        Run r{e, fake, fake, Arena(), nullptr, B};
        rc = dec_forward_walk(e, r, reinterpret_cast<const float*>(fake), reinterpret_cast<float*>(fake));
        peak = r.wsa.peak;
    }
    if (rc == 0) {
        Run r{e, fake, fake, Arena(), nullptr, B};
        rc = dec_backward_walk(e, r, reinterpret_cast<const float*>(fake), reinterpret_cast<float*>(fake));
        peak = std::max(peak, r.wsa.peak);
    }
    g_dry_run = false;
    if (rc) { L = DecLayout(); return rc; }
    L.ws_bytes = peak + 1024;
    L.measured = true;
    return 0;
}

int tml_decoder_query(TmlEncoder* e, int B, int h, int w, size_t* workspace_bytes, size_t* saved_bytes) {
    if (!e || !e->finalized || !e->has_decoder) { set_error("decoder weights were not loaded"); return -1; }
    RC(dec_layout(e, B, h, w));
    if (workspace_bytes) *workspace_bytes = e->dlay.ws_bytes;
    if (saved_bytes) *saved_bytes = e->dlay.saved_bytes;
    return 0;
}

int tml_decoder_forward(TmlEncoder* e, const float* z, int B, int h, int w, float* image, void* saved, void* ws,
                        void* stream) {
    if (!e || !e->finalized || !e->has_decoder) { set_error("decoder weights were not loaded"); return -1; }
    if (!z || !image || !saved || !ws) { set_error("null buffer"); return -1; }
    DeviceGuard guard(e->device);
    RC(dec_layout(e, B, h, w));
    Run r{e, reinterpret_cast<char*>(saved), reinterpret_cast<char*>(ws), Arena(), reinterpret_cast<cudaStream_t>(stream), B};
    RC(dec_forward_walk(e, r, z, image));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int dec_forward_walk(TmlEncoder* e, Run& r, const float* z, float* image) {
    const DecLayout& L = e->dlay;
    const TmlEncoderCfg& c = e->cfg;
    const int nb = c.num_blocks, ns = e->num_sms;
    const int B = L.B, h = L.h, w = L.w;
    r.statbuf[0] = r.Walloc<float>(fused_partial_bytes(B, L.Hl, L.Wl));
    r.statbuf[1] = r.Walloc<float>(fused_partial_bytes(B, L.Hl, L.Wl));
    const int Ctop = c.block_out_channels[nb - 1];
    {   // post_quant_conv + conv_in
        const size_t m = r.wsa.mark();
        bf16* zp = r.Walloc<bf16>(act_bytes(B, h, w, 64));
        launch_latent_pack(z, e->pq_w, e->pq_b, zp, B, h * w, r.st);
        GemmOp o = dense_conv_op("dec.conv_in", zp, B, h, w, 64, e->dec_conv_in.fwd, Ctop, 1, h, w, e->dec_conv_in.bias,
                                 nullptr, r.S<bf16>(L.x0));
        r.pending = fuse_stats(o, r.statbuf[1], h, w);
        RC(gemm_launch(o, ns, r.st));
        r.wsa.reset(m);
    }
    size_t ri = 0;
    RC(resnet_forward(r, e->dec_resnets[ri], L.res[ri])); ++ri;
    if (e->dec_has_attn) RC(attn_forward(r, e->dec_attn, L.attn));
    RC(resnet_forward(r, e->dec_resnets[ri], L.res[ri])); ++ri;
    size_t ui = 0;
    for (int i = 0; i < nb; ++i) {
        for (int j = 0; j < c.layers_per_block + 1; ++j, ++ri) RC(resnet_forward(r, e->dec_resnets[ri], L.res[ri]));
        if (i != nb - 1) {
            const UpRec& u = L.ups[ui];
            const Conv3& cv = e->ups[ui];
            const size_t m = r.wsa.mark();
            bf16* up = r.Walloc<bf16>(act_bytes(B, 2 * u.h, 2 * u.w, cv.ci));
            launch_upsample2x(r.S<bf16>(L.res[ri - 1].out), up, B, u.h, u.w, cv.ci, r.st);
            GemmOp o = dense_conv_op("dec.upsample", up, B, 2 * u.h, 2 * u.w, cv.ci, cv.fwd, cv.co, 1, 2 * u.h, 2 * u.w,
                                     cv.bias, nullptr, r.S<bf16>(u.out));
            r.pending = fuse_stats(o, r.statbuf[1], 2 * u.h, 2 * u.w);
            RC(gemm_launch(o, ns, r.st));
            r.wsa.reset(m);
            ++ui;
        }
    }
    {   // conv_norm_out + SiLU + conv_out -> fp32 NCHW image
        const int H = L.Hl, W = L.Wl, C = e->dec_norm_out.C;
        bf16* a = r.Walloc<bf16>(act_bytes(B, H, W, C));
        RC(gn_forward(r, r.S<bf16>(L.xlast), e->dec_norm_out, L.gout, a, H * W, 1, r.pending));
        GemmOp o = dense_conv_op("dec.conv_out", a, B, H, W, C, e->dec_conv_out.fwd, 16, 1, H, W, e->dec_conv_out.bias,
                                 nullptr, nullptr);
        o.D = image; o.out_fp32 = 1; o.n_store = 3;
        o.D_sB = (int64_t)3 * H * W; o.D_sH = W; o.D_sW = 1; o.D_sN = (int64_t)H * W;
        RC(gemm_launch(o, ns, r.st));
    }
    if (L.measured && r.wsa.peak > L.ws_bytes) { set_error("internal: decoder workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
    return 0;
}

int tml_decoder_backward(TmlEncoder* e, const float* dimage, int B, int h, int w, const void* saved, float* dz,
                         void* ws, void* stream) {
    if (!e || !e->finalized || !e->has_decoder) { set_error("decoder weights were not loaded"); return -1; }
    if (!dimage || !saved || !ws || !dz) { set_error("null buffer"); return -1; }
    DeviceGuard guard(e->device);
    RC(dec_layout(e, B, h, w));
    Run r{e, const_cast<char*>(reinterpret_cast<const char*>(saved)), reinterpret_cast<char*>(ws), Arena(),
          reinterpret_cast<cudaStream_t>(stream), B};
    RC(dec_backward_walk(e, r, dimage, dz));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int dec_backward_walk(TmlEncoder* e, Run& r, const float* dimage, float* dz) {
    const DecLayout& L = e->dlay;
    const TmlEncoderCfg& c = e->cfg;
    const int nb = c.num_blocks, ns = e->num_sms;
    const int B = L.B, h = L.h, w = L.w;
    // ping-pong gradient buffers sized for the largest activation
    size_t gmax = act_bytes(B, h, w, c.block_out_channels[nb - 1]);
    {
        int hh = h, ww = w;
        for (int i = 0; i < nb; ++i) {
            const int cout = c.block_out_channels[nb - 1 - i];
            const int cprev = i ? c.block_out_channels[nb - i] : cout;
            gmax = std::max(gmax, std::max(act_bytes(B, hh, ww, cout), act_bytes(B, hh, ww, cprev)));
            if (i != nb - 1) { hh *= 2; ww *= 2; gmax = std::max(gmax, act_bytes(B, hh, ww, cout)); }
        }
    }
    bf16* G[2] = {r.Walloc<bf16>(gmax), r.Walloc<bf16>(gmax)};
    int cur = 0;
    g_dump_next = 0;
    {   // d(image) -> d(conv_norm_out input)
        const int H = L.Hl, W = L.Wl, C = e->dec_norm_out.C;
        const size_t m = r.wsa.mark();
        bf16* di64 = r.Walloc<bf16>(act_bytes(B, H, W, 64));
        launch_image_pack(dimage, di64, B, (long long)H * W, r.st);
        bf16* d_a = r.Walloc<bf16>(act_bytes(B, H, W, C));
        GemmOp go = dense_conv_op("dec.conv_out.dgrad", di64, B, H, W, 64, e->dec_conv_out.bwd, C, 1, H, W, nullptr, nullptr, d_a);
        float* pbuf = r.Walloc<float>(fused_partial_bytes(B, H, W));
        const Partials po = fuse_gn_bwd(r, go, r.S<bf16>(L.xlast), e->dec_norm_out, L.gout, 1, pbuf, H, W);
        RC(gemm_launch(go, ns, r.st));
        RC(gn_backward(r, r.S<bf16>(L.xlast), d_a, e->dec_norm_out, L.gout, nullptr, G[cur], H * W, 1, po));
        r.wsa.reset(m);
        dump_grad(G[cur], act_bytes(B, H, W, C), r.st);
    }
    size_t ri = e->dec_resnets.size();
    size_t ui = e->ups.size();
    for (int i = nb - 1; i >= 0; --i) {
        if (i != nb - 1) {
            // undo the Upsample2D that closes block i: conv dgrad, then sum the four nearest-neighbour copies
            --ui;
            const UpRec& u = L.ups[ui];
            const Conv3& cv = e->ups[ui];
            const size_t m = r.wsa.mark();
            bf16* d_up = r.Walloc<bf16>(act_bytes(B, 2 * u.h, 2 * u.w, cv.ci));
            RC(gemm_launch(dense_conv_op("dec.upsample.dgrad", G[cur], B, 2 * u.h, 2 * u.w, cv.co, cv.bwd, cv.ci, 1, 2 * u.h,
                                         2 * u.w, nullptr, nullptr, d_up), ns, r.st));
            launch_upsample2x_bwd(d_up, G[cur ^ 1], B, u.h, u.w, cv.ci, r.st);
            r.wsa.reset(m);
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, u.h, u.w, cv.ci), r.st);
        }
        for (int j = c.layers_per_block; j >= 0; --j) {
            --ri;
            RC(resnet_backward(r, e->dec_resnets[ri], L.res[ri], G[cur], G[cur ^ 1]));
            cur ^= 1;
            dump_grad(G[cur], act_bytes(B, L.res[ri].h, L.res[ri].w, e->dec_resnets[ri].ci), r.st);
        }
    }
    const size_t mid_bytes = act_bytes(B, h, w, c.block_out_channels[nb - 1]);
    --ri;
    RC(resnet_backward(r, e->dec_resnets[ri], L.res[ri], G[cur], G[cur ^ 1])); cur ^= 1;
    dump_grad(G[cur], mid_bytes, r.st);
    if (e->dec_has_attn) { RC(attn_backward(r, e->dec_attn, L.attn, G[cur], G[cur ^ 1])); cur ^= 1; dump_grad(G[cur], mid_bytes, r.st); }
    --ri;
    RC(resnet_backward(r, e->dec_resnets[ri], L.res[ri], G[cur], G[cur ^ 1])); cur ^= 1;
    dump_grad(G[cur], mid_bytes, r.st);
    {   // conv_in dgrad (N padded 4 -> 64) and post_quant_conv backward -> fp32 NCHW dz
        const int Ctop = c.block_out_channels[nb - 1];
        bf16* d_zp = r.Walloc<bf16>(act_bytes(B, h, w, 64));
        RC(gemm_launch(dense_conv_op("dec.conv_in.dgrad", G[cur], B, h, w, Ctop, e->dec_conv_in.bwd, 64, 1, h, w, nullptr,
                                     nullptr, d_zp), ns, r.st));
        launch_latent_unpack_bwd(d_zp, e->pq_w, dz, B, h * w, r.st);
    }
    if (L.measured && r.wsa.peak > L.ws_bytes) { set_error("internal: decoder workspace overrun (%zu > %zu)", r.wsa.peak, L.ws_bytes); return -40; }
    return 0;
}

size_t tml_image_loss_workspace(int B) { return image_loss_workspace_bytes(B); }

int tml_image_loss(const float* out, const float* target, const float* source, int B, int64_t per_image, float rec_lambda,
                   float pert_lambda, float* rec, float* pert, float* dout, void* ws, void* stream) {
    if (!out || !target || !ws) { set_error("null buffer"); return -1; }
    if (B <= 0) return 0;
    launch_image_loss(out, target, source, B, per_image, rec_lambda, pert_lambda, rec, pert, dout, ws,
                      reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_posterior_sample(const float* moments, const float* noise, float* z, int B, int h, int w, void* stream) {
    if (!moments || !z) { set_error("null buffer"); return -1; }
    launch_posterior_sample(moments, noise, z, B, h * w, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_posterior_sample_backward(const float* moments, const float* noise, const float* dz, float* dmoments, int B,
                                  int h, int w, void* stream) {
    if (!moments || !dz || !dmoments) { set_error("null buffer"); return -1; }
    launch_posterior_sample_bwd(moments, noise, dz, dmoments, B, h * w, reinterpret_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tml_debug_decoder_saved_tensor(TmlEncoder* e, const char* name, int index, size_t* offset, int dims[4]) {
    if (!e || !name || !offset || !dims) { set_error("null argument"); return -1; }
    const DecLayout& L = e->dlay;
    if (!L.saved_bytes) { set_error("no decoder layout yet"); return -1; }
    const std::string n(name);
    auto set = [&](size_t off, int h, int w, int c) { *offset = off; dims[0] = L.B; dims[1] = h; dims[2] = w; dims[3] = c; return 0; };
    if (n == "conv_in") return set(L.x0, L.h, L.w, e->dec_conv_in.co);
    if (n == "resnet_h1" || n == "resnet_out") {
        if (index < 0 || index >= (int)L.res.size()) { set_error("index"); return -1; }
        const ResnetRec& r = L.res[index];
        return set(n == "resnet_h1" ? r.h1 : r.out, r.h, r.w, e->dec_resnets[index].co);
    }
    if (n == "up_out") {
        if (index < 0 || index >= (int)L.ups.size()) { set_error("index"); return -1; }
        return set(L.ups[index].out, 2 * L.ups[index].h, 2 * L.ups[index].w, e->ups[index].co);
    }
    if (n == "attn_out") return set(L.attn.out, L.attn.h, L.attn.w, e->dec_attn.C);
    set_error("unknown tensor '%s'", name);
    return -1;
}

}  // extern "C"
