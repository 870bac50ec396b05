// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a).  Two kernels:
//
//   conv_gemm_tcgen05_kernel  -- pixel-major (M = 128 output pixels, N = output channels): every GEMM shape of the
//     path (3x3 s1/s2, 1x1, linears, attention products and their input-gradient forms).
//       * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared memory ring;
//         convolution taps are shifted TMA box origins on the NHWC tensor, the hardware zero-fills out-of-range
//         pixels (= the conv padding), so no im2col buffer exists; in halo mode one (mt+2) x 130-pixel tile per
//         64-channel chunk serves all nine taps through row-shifted UMMA descriptors;
//       * one elected thread issues tcgen05.mma (M=128 or, on CTA pairs, 256; N<=256; K=16; bf16 x bf16 -> fp32) with
//         the accumulator in TMEM, double buffered so that the epilogue of tile i overlaps the main loop of tile i+1;
//       * eight epilogue warps read the accumulator with tcgen05.ld (thread = output pixel), apply alpha / bias /
//         residual, reduce GroupNorm sums, and either store from registers (32 bytes per lane) or stage the tile in
//         shared memory for a TMA store;
//       * persistent grid (one CTA per SM), static round-robin tiles, warp-specialised roles synchronised only
//         through mbarriers; cta_group::2 pairs for the weight-heavy long-K shapes.
//   conv3x3_swapped_kernel  -- channel-major (M = 128/256 output channels, N = 256 pixels) for 3x3 stride-1
//     convolutions on rows of >= 128 pixels: where most of the FLOPs of the path run (see the comment above it).
//
// Replaces the cuDNN conv fwd/dgrad + cuBLAS linear calls PyTorch makes for the reference's
// vae.encode (main.py:75,191) and its autograd backward (main.py:176).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "gemm.h"
#include "ptx.cuh"

namespace tml {

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local message, C ABI returns negative codes)
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<int> g_impl{0};
static std::atomic<long> g_tc_launches{0};
void gemm_set_impl(int impl) { g_impl.store(impl); }
int gemm_get_impl() { return g_impl.load(); }
long gemm_launch_count() { return g_tc_launches.load(); }

// ------------------------------------------------------------------------------------------------
// optional per-launch CUDA-event timing of the tcgen05 kernel (bench.py's roofline numbers)
// ------------------------------------------------------------------------------------------------
struct TimedLaunch { cudaEvent_t a, b; double flops; char key[96]; };
static std::vector<TimedLaunch> g_timed;
static std::vector<cudaEvent_t> g_event_pool;
static bool g_timing = false;
static size_t g_timing_cap = 0;
static long g_timing_dropped = 0;

void gemm_timing_enable(int max_launches) {
    g_timing = max_launches > 0;
    g_timing_cap = max_launches > 0 ? (size_t)max_launches : 0;
    g_timing_dropped = 0;
    for (auto& t : g_timed) { g_event_pool.push_back(t.a); g_event_pool.push_back(t.b); }
    g_timed.clear();
    if (g_timing) g_timed.reserve(g_timing_cap);
}
static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
// Sums elapsed time / algorithmic flops over the recorded launches (caller has synchronised).
void gemm_timing_collect(double* total_ms, double* total_flops, long* launches, long* dropped) {
    double ms = 0.0, fl = 0.0;
    for (auto& t : g_timed) {
        float e = 0.f;
        if (cudaEventElapsedTime(&e, t.a, t.b) == cudaSuccess) ms += e;
        fl += t.flops;
    }
    *total_ms = ms; *total_flops = fl; *launches = (long)g_timed.size(); *dropped = g_timing_dropped;
}

// Per-shape table of the recorded launches: "name|M|N|K|gn|count|ms|flops" lines (caller has synchronised).
size_t gemm_timing_report(char* buf, size_t cap) {
    struct Agg { std::string key; int n = 0; double ms = 0, fl = 0; };
    std::vector<Agg> aggs;
    for (auto& t : g_timed) {
        float e = 0.f;
        if (cudaEventElapsedTime(&e, t.a, t.b) != cudaSuccess) continue;
        Agg* a = nullptr;
        for (auto& x : aggs) if (x.key == t.key) { a = &x; break; }
        if (!a) { aggs.push_back(Agg()); a = &aggs.back(); a->key = t.key; }
        a->n++; a->ms += e; a->fl += t.flops;
    }
    size_t off = 0;
    for (auto& a : aggs) {
        int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s|%d|%.4f|%.6g\n", a.key.c_str(), a.n, a.ms, a.fl);
        if (w < 0 || off + (size_t)w >= cap) break;
        off += (size_t)w;
    }
    return off;
}

// ------------------------------------------------------------------------------------------------
// tiling
// ------------------------------------------------------------------------------------------------
constexpr int kBlockK = 64;          // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaM = 128;
constexpr int kATileBytes = kUmmaM * kBlockK * 2;  // 16 KiB
constexpr int kMaxSmem = 227 * 1024;
constexpr int kBarrierBytes = 256;
constexpr int kThreads = 384;        // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
constexpr int kEpiThreads = 256;

static int largest_divisor_le(int n, int cap, int multiple_of) {
    for (int d = (cap < n ? cap : n); d >= 1; --d)
        if (n % d == 0 && d % multiple_of == 0) return d;
    return 0;
}

constexpr int kGnSmemBytes = 8192;  // epilogue scratch: per-tile GN constants + cross-warp reduction

static void tile_shape(int OH, int OW, int* TW, int* TH) {
    *TW = largest_divisor_le(OW, 128, 8);
    *TH = *TW ? largest_divisor_le(OH, 128 / *TW, 1) : 0;
}
int gemm_gn_tiles_per_image(int OH, int OW) {
    int TW, TH;
    tile_shape(OH, OW, &TW, &TH);
    return TW ? (OH / TH) * (OW / TW) : 0;
}

bool gemm_swapped_shape(const GemmOp& op);
int sw_px_per_warp(bool gnb);
int gemm_row_partials(const GemmOp& op) {
    GemmTiling t;
    const int rc = gemm_plan(op, &t);
    return rc ? rc : 2 * t.n_tiles;
}
// Partial-sum entries per image written by the fused GroupNorm reduction of this op.
int gemm_gn_chunks_per_image(const GemmOp& op) {
    if (op.gn_mode != 0 && gemm_swapped_shape(op))   // one entry per epilogue warp's pixel range (64 or 128 pixels)
        return op.OH * op.OW / (op.gn_mode == 2 ? sw_px_per_warp(true) : sw_px_per_warp(false));
    GemmOp shape = op;        // the geometry does not depend on the reduction's buffers, which a caller may set later
    shape.gn_mode = 0;
    GemmTiling t;
    if (gemm_plan(shape, &t) == 0 && t.h66) return t.tiles_h;   // pitch-66 tiles of 64-pixel rows: 33 per 64 rows
    return gemm_gn_tiles_per_image(op.OH, op.OW);
}

int gemm_plan(const GemmOp& op, GemmTiling* t) {
    if (op.A_C % kBlockK != 0) { set_error("%s: A_C=%d is not a multiple of 64", op.name, op.A_C); return -1; }
    if (op.N % 16 != 0) { set_error("%s: N=%d is not a multiple of 16", op.name, op.N); return -1; }
    if (op.stride != 1 && op.stride != 2) { set_error("%s: stride %d unsupported", op.name, op.stride); return -1; }
    if (op.ntaps < 1 || op.ntaps > kMaxTaps) { set_error("%s: ntaps=%d", op.name, op.ntaps); return -1; }
    int TW, TH;
    tile_shape(op.OH, op.OW, &TW, &TH);
    if (TW == 0) { set_error("%s: output width %d has no tile width (multiple of 8, <=128)", op.name, op.OW); return -1; }
    if (op.stride == 2 && (op.A_W % 2 != 0)) { set_error("%s: stride-2 input width must be even", op.name); return -1; }
    if (op.a_trans) {
        // the 128 (or 64) rows of a tile must be consecutive values of m = oh*OW + ow: whole rows, or one row segment
        if (op.ntaps != 1 || op.stride != 1 || op.dbg_shift != 0 || op.A_sK <= 0 || (op.A_sK * 2) % 16 != 0 ||
            (TW * TH) % 64 != 0 || !(TW == op.OW || TH == 1) || op.in_gn_ss != nullptr) {
            set_error("%s: transposed A needs one tap, stride 1 and tiles of 64/128 consecutive rows (TW=%d TH=%d OW=%d)",
                      op.name, TW, TH, op.OW);
            return -1;
        }
    }
    int BN = 0;
    if (op.N % 256 == 0) BN = 256;
    else if (op.N % 128 == 0) BN = 128;
    else if (op.N <= 256) BN = op.N;
    else {
        // any other width (e.g. 2240 attention tokens of a 320 x 448 image): the largest column tile that divides N
        for (int d = 256; d >= 32 && BN == 0; d -= 32) if (op.N % d == 0) BN = d;
        for (int d = 240; d >= 16 && BN == 0; d -= 16) if (op.N % d == 0) BN = d;
        if (BN == 0) { set_error("%s: N=%d has no column tile (a multiple of 16, at most 256)", op.name, op.N); return -1; }
    }
    // Small problems (a few tile waves): a narrower column tile can fill the last wave better than the widest one, e.g.
    // 4096 x 1280 (the UNet's 16 x 16 level) is 80 CTA-pair tiles of 256 columns = 2 waves of 74 pairs, but 256 tiles of
    // 160 columns = 2 waves of 148 CTAs at 0.62 of the work per tile.  Cost model: waves x (columns + fixed per-tile cost).
    // (measured: neutral in situ, 758 vs 805 TFLOP/s in isolation at 4096 x 1280 x 11520 because the narrower tiles lose the
    // CTA pairs' halved weight traffic -- off unless TML_BN_HEUR=1)
    static const bool no_bn_heur = !(getenv("TML_BN_HEUR") && getenv("TML_BN_HEUR")[0] == '1');   // experiment switch
    if (!no_bn_heur && op.N > 256 && op.epi_mode == 0 && op.gn_mode == 0 && !op.a_trans && TW * TH > 0 &&
        !(op.stride == 1 && op.ntaps == 9 && (op.OW % 128 == 0 || op.OW == 64))) {   // (halo modes keep their geometry)
        const long sub = (long)op.A_B * (op.OH / TH) * (op.OW / TW);
        const int rows_valid = TW * TH;
        auto cost = [&](int bn) -> double {
            const long nt = op.N / bn;
            const bool pair = bn == 256 && rows_valid == 128 && op.ntaps * op.A_C >= 2048 && sub % 2 == 0 && sub * nt >= 4;
            if (pair) return 0.9 * (double)((sub / 2 * nt + 73) / 74) * (bn + 48);
            const bool mt2 = bn <= 128 && op.B_sBatch == 0 && sub % 2 == 0 && sub * nt >= 2 * 148;
            if (mt2) return (double)((sub / 2 * nt + 147) / 148) * (2 * bn + 48);
            return (double)((sub * nt + 147) / 148) * (bn + 48);
        };
        double best = cost(BN);
        const double def = best;
        int best_bn = BN;
        const int cands[3] = {192, 160, 128};
        for (int c : cands)
            if (c < BN && op.N % c == 0 && cost(c) < best) { best = cost(c); best_bn = c; }
        if (best < 0.85 * def) BN = best_bn;
    }
    if (op.resid && BN < 32) { set_error("%s: residual needs N >= 32", op.name); return -1; }
    if (op.gn_mode != 0) {
        const int cpg = op.N / 32;
        if (op.out_fp32 || op.D_sN != 1 || op.N % 32 || (cpg != 4 && cpg != 8 && cpg != 16) || BN % 32 ||
            !op.gn_partial || (op.gn_mode == 2 && (!op.gn_x || !op.gn_ss || !op.gn_mr || !op.gn_gamma))) {
            set_error("%s: fused GroupNorm reduction needs a dense bf16 output with N in {128,256,512}", op.name);
            return -1;
        }
    }
    // Halo mode: stride-1 3x3 convolutions whose output rows split into 128-pixel segments.
    static const bool no_halo = getenv("TML_NO_HALO") && getenv("TML_NO_HALO")[0] == '1';   // tuning switch
    bool halo = !no_halo && op.stride == 1 && op.ntaps == 9 && op.OW % 128 == 0 && op.B_sBatch == 0 &&
                op.dbg_shift == 0 && op.OW == op.A_W && op.OH == op.A_H;
    if (halo) {
        unsigned seen = 0;
        for (int i = 0; i < 9; ++i) {
            if (op.dh[i] < -1 || op.dh[i] > 1 || op.dw[i] < -1 || op.dw[i] > 1) { halo = false; break; }
            seen |= 1u << ((op.dh[i] + 1) * 3 + op.dw[i] + 1);
        }
        if (seen != 0x1FFu) halo = false;
    }
    // Halo mode for rows of 64 pixels ("h66", the 512-channel layers at 64^2): the halo tile is {64 channels, 66 pixels
    // (x = -1 .. 64, the two ends zero-filled by TMA), 6 rows}, i.e. the image with a row pitch of 66 in SHARED memory only.
    // In that pitch-66 index space s = 66*y + x + 1 a tap is the affine shift 66*dh + dw, so an M tile is 128 consecutive
    // slots starting anywhere (two of every 66 are the zero columns: junk output rows that are never stored), read through
    // row-shifted descriptors like the 130-pixel tile.  64 rows x 66 = 33 tiles of 128 per image.  The A operand then costs
    // 50 KB per 64-channel chunk instead of nine 16 KB tap tiles: 64 -> 40 B/clk/SM from L2, which is what bounded these layers.
    static const bool no_h66 = getenv("TML_NO_H66") && getenv("TML_NO_H66")[0] == '1';   // tuning switch
    bool h66 = !no_halo && !no_h66 && !halo && op.stride == 1 && op.ntaps == 9 && op.OW == 64 && op.A_W == 64 &&
               op.OH == op.A_H && (op.OH * 66) % 128 == 0 && op.B_sBatch == 0 && op.dbg_shift == 0 && op.D_sN == 1 &&
               !op.out_fp32 && (op.n_store == 0 || op.n_store == op.N) && op.epi_mode == 0 && op.row_scale == nullptr;
    if (h66) {
        unsigned seen = 0;
        for (int i = 0; i < 9; ++i) {
            if (op.dh[i] < -1 || op.dh[i] > 1 || op.dw[i] < -1 || op.dw[i] > 1) { h66 = false; break; }
            seen |= 1u << ((op.dh[i] + 1) * 3 + op.dw[i] + 1);
        }
        if (seen != 0x1FFu) h66 = false;
    }
    t->h66 = h66 ? 1 : 0;
    if (h66) {
        t->halo = 1;
        t->out_bytes = 0;
        t->TW = 64; t->TH = 2; t->rows_valid = 128;
        t->tiles_w = 1; t->tiles_h = op.OH * 66 / 128;      // tiles per image (in the pitch-66 space)
        t->BN = BN; t->n_tiles = op.N / BN; t->kchunks = op.A_C / kBlockK;
        t->mt = 1;
        t->halo_bytes = ((6 * 66 * 128) + 1023) / 1024 * 1024;
        const long cta_m_tiles = (long)op.A_B * t->tiles_h;
        static const bool no_pair66 = getenv("TML_PAIR") && getenv("TML_PAIR")[0] == '0';   // tuning switch
        // (a pair = the same tile of two consecutive images, see decode_sub: needs an even number of images)
        // (narrower column tiles, e.g. 160 of the UNet's 320-channel layers, pair up as well: each CTA stages half of the
        // weight tile, which is what bounds these layers -- 80 B/clk/SM from L2 on single CTAs)
        static const int pair66_min_bn = getenv("TML_H66_PAIR_MIN_BN") ? atoi(getenv("TML_H66_PAIR_MIN_BN")) : 256;   // experiment switch
        t->pair = (!no_pair66 && (BN == 256 || (BN >= pair66_min_bn && BN % 32 == 0)) && op.A_B % 2 == 0 &&
                   cta_m_tiles * t->n_tiles >= 4) ? 1 : 0;
        t->stage_bytes = (t->pair ? BN / 2 : BN) * 128;
        int stages = (kMaxSmem - 1024 - kBarrierBytes - kGnSmemBytes - 2 * t->halo_bytes) / t->stage_bytes;
        if (stages > 8) stages = 8;
        if (stages < 2) { set_error("%s: halo tiles do not fit", op.name); return -1; }
        t->stages = stages;
        t->smem_bytes = size_t(2) * t->halo_bytes + size_t(stages) * t->stage_bytes + kBarrierBytes + kGnSmemBytes + 1024;
        return 0;
    }
    if (halo) { TW = 128; TH = 1; }
    t->halo = halo ? 1 : 0;
    t->out_bytes = 0;
    t->TW = TW;
    t->TH = TH;
    t->rows_valid = TW * TH;
    t->tiles_w = op.OW / TW;
    t->tiles_h = op.OH / TH;
    t->BN = BN;
    t->n_tiles = op.N / BN;
    t->kchunks = op.A_C / kBlockK;
    t->pair = 0;
    if (halo) {
        t->mt = (BN <= 128 && op.OH % 2 == 0) ? 2 : 1;
        t->halo_bytes = (((t->mt + 2) * 130 * 128) + 1023) / 1024 * 1024;
        // CTA pairs (cta_group::2, M = 256 over two SMs): each CTA stages its own halo tile and half of every
        // weight tile.  Correct (tests run it); off by default here because the layers that would profit run on the
        // operand-swapped kernel (TML_PAIR=1 switches it on for what is left in halo mode).
        static const bool use_pair = getenv("TML_PAIR") && getenv("TML_PAIR")[0] == '1';   // experiment switch
        const long cta_m_tiles = (long)op.A_B * (op.OH / t->mt) * t->tiles_w;
        t->pair = (use_pair && cta_m_tiles % 2 == 0 && cta_m_tiles * t->n_tiles >= 4) ? 1 : 0;
        t->stage_bytes = (t->pair ? BN / 2 : BN) * 128;   // the ring holds weight tiles only
        int stages = (kMaxSmem - 1024 - kBarrierBytes - kGnSmemBytes - 2 * t->halo_bytes) / t->stage_bytes;
        if (stages > 8) stages = 8;
        if (stages < 2) { set_error("%s: halo tiles do not fit", op.name); return -1; }
        t->stages = stages;
        t->smem_bytes = size_t(2) * t->halo_bytes + size_t(stages) * t->stage_bytes + kBarrierBytes + kGnSmemBytes + 1024;
        return 0;
    }
    t->halo_bytes = 0;
    // Two 128-row sub-tiles per CTA tile when the accumulators fit (2 x 2 x BN <= 512 TMEM columns):
    // every B (weight) tile is then fetched from L2 once per 256 output pixels instead of once per 128.
    const long sub_tiles = (long)op.A_B * t->tiles_h * t->tiles_w;
    static const bool no_mt2 = getenv("TML_NO_MT2") && getenv("TML_NO_MT2")[0] == '1';   // tuning switch
    t->mt = (!no_mt2 && !op.a_trans && BN <= 128 && op.B_sBatch == 0 && sub_tiles % 2 == 0 && sub_tiles * t->n_tiles >= 2 * 148) ? 2 : 1;
    // CTA pairs for the weight-heavy long-K tiles (BN = 256, K >= 2048: the 512-channel convolutions at 64^2): M = 256
    // pixels over two SMs, each CTA stages its own 128 pixels and HALF of every weight tile, which cuts the L2 -> SM
    // traffic that bounds these layers by a third (48 -> 32 KB per four MMAs).
    static const bool no_pair = getenv("TML_PAIR") && getenv("TML_PAIR")[0] == '0';   // tuning switch
    // (a per-image B operand is fine when both CTAs of a pair always work on the same image)
    const bool b_ok = op.B_sBatch == 0 || (t->tiles_h * t->tiles_w) % 2 == 0;
    // (the attention logits, K = 512 and a per-image B operand, are L2-bound on single CTAs: 48 KB per four MMAs)
    t->pair = (!no_pair && b_ok && op.dbg_shift == 0 && BN == 256 && t->mt == 1 &&
               t->rows_valid == 128 && (op.ntaps * op.A_C >= 2048 || op.epi_mode != 0) && sub_tiles % 2 == 0 &&
               sub_tiles * t->n_tiles >= 4) ? 1 : 0;
    // Plain dense outputs (no fused reduction) leave through shared memory and TMA stores: the per-lane
    // 32-byte global stores of the register epilogue cost one LSU request per sector (~0.2 ms per GB of output, measured),
    // which is what bounds the thin GEMMs (1x1 shortcuts, parity-class dgrads, attention logits).
    // Two column halves x two buffers of [128 rows][32 columns].
    static const bool no_tma_store = getenv("TML_NO_TMA_STORE") && getenv("TML_NO_TMA_STORE")[0] == '1';   // tuning switch
    const int es_out = op.out_fp32 ? 4 : 2;
    const bool dense_rows = op.D_sN == 1 && (op.n_store == 0 || op.n_store == op.N) && !(op.out_fp32 && op.beta != 0.f) &&
                            (op.D_sW * es_out) % 16 == 0 && (op.D_sH * es_out) % 16 == 0 && (op.D_sB * es_out) % 16 == 0;
    // (each of the 8 epilogue warps stages its own [32 rows][32 columns] x 2 buffers: rows_valid is a multiple of 32 and
    // a warp's 32 rows are 32 pixels of one row or whole rows of the tile)
    const bool warp_rows = t->rows_valid % 32 == 0 && (TW % 32 == 0 || 32 % TW == 0);
    if (op.epi_mode != 0) {
        if (op.gn_mode != 0 || op.out_fp32 || !dense_rows || !warp_rows || BN % 32 != 0 || op.dbg_shift != 0 || op.bias ||
            (op.epi_mode != 3 && op.epi_mode != 1 && op.epi_mode != 2)) {
            set_error("%s: row-wise epilogue %d needs a dense bf16 output of whole 32-row warps", op.name, op.epi_mode);
            return -1;
        }
        t->out_bytes = 8 * 2 * 32 * 32 * es_out;
    } else if (!no_tma_store && !t->pair && op.gn_mode == 0 && dense_rows && warp_rows && BN % 32 == 0 && op.dbg_shift == 0 &&
        getenv("TML_DBG_NO_EPI") == nullptr)
        t->out_bytes = 8 * 2 * 32 * 32 * es_out;
    int stage_bytes = t->mt * kATileBytes + (((t->pair ? BN / 2 : BN) * 128 + 1023) / 1024) * 1024;
    int stages = (kMaxSmem - 1024 - kBarrierBytes - kGnSmemBytes - t->out_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    // (the ring runs across tiles: with short K it holds several tiles' operands, which is what hides the load latency
    // of the memory-bound 1x1 / thin GEMMs -- capping it at one tile's k-blocks left them latency-bound)
    if (stages < 2) stages = 2;
    t->stages = stages;
    t->stage_bytes = stage_bytes;
    t->smem_bytes = size_t(t->out_bytes) + size_t(stages) * stage_bytes + kBarrierBytes + kGnSmemBytes + 1024;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int mode;  // 0: stride 1 (4-D map c,w,h,b)   1: stride 2 (5-D map c,wpar,w/2,h,b)
    int a_trans;   // A stored [batch][k][m] (3-D map m,k,b; boxes of 64 m x 64 k): MN-major UMMA operand
    int b_prefetch;   // > 0: L2 prefetch of the weight tile this many k-blocks ahead (weight-streaming layers: few output
                      // rows, tens of MB of weights coming from HBM once -- the ring alone keeps too few bytes in flight)
    int TW, TH, rows_valid;
    int tiles_w, tiles_h, nimg;
    int n_tiles, BN;
    int mt;    // M sub-tiles (128 rows each) per CTA tile: 2 shares every B tile between two accumulators
    int kchunks, ntaps;
    int dh[kMaxTaps], dw[kMaxTaps];
    int b_batched;
    int stages, stage_bytes;
    float alpha;
    const float* bias;
    const __nv_bfloat16* resid;
    long long R_sB, R_sH, R_sW;
    void* D;
    int out_fp32;
    long long D_sB, D_sH, D_sW, D_sN;
    int n_store;
    float beta;
    // fused GroupNorm reductions over the (bf16-rounded) output tile
    int gn_mode;   // 0 none, 1 (sum, sumsq), 2 (sum dxh, sum dxh*xh) of the GroupNorm backward
    int gn_cpg;    // channels per group of the output tensor (4, 8 or 16)
    float* gn_partial;               // [nimg][tiles_h*tiles_w][32][2]
    const __nv_bfloat16* gn_x;       // mode 2: the GroupNorm input, same layout as D
    const float2* gn_ss;             // mode 2: [nimg][N] (scale, shift)
    const float2* gn_mr;             // mode 2: [nimg][32] (mean, rstd)
    const float* gn_gamma;           // mode 2: [N]
    int gn_silu;
    int dbg_shift, dbg_bo;
    // halo mode (3x3 stride-1 convolutions, output rows of 128 pixels): one (mt+2) x 130-pixel halo tile per
    // 64-channel chunk serves all nine taps through row-shifted UMMA descriptors
    int halo, halo_bytes;
    int h66;           // halo mode on rows of 64 pixels: pitch-66 tile of 6 rows, M tile = 128 consecutive slots (see gemm_plan)
    int pair;          // halo mode on CTA pairs: cta_group::2 MMA (M = 256 over two SMs), each CTA stages half of B
    volatile int* hang_where;  // mapped host word that receives the id of a wait that timed out
    int out_bytes;     // > 0: dense outputs are staged in shared memory ([2 halves][2 buffers][128 rows][32 cols]) and TMA-stored
    int dbg_no_epi;    // experiment: 1 = the epilogue only hands the accumulator back, 2 = no global memory ops, 3 = no GN math
    int dbg_mma_only;  // experiment: operands are loaded for the first pass over the ring only
    // row-wise attention epilogues (see GemmOp::epi_mode) and the per-row output scale
    int epi_mode;
    const float* row_a;
    const float* row_b;
    float* row_part;
    float exp_scale;
    const float* row_scale;
    int rows_img, ow_full;   // output pixels per image, output row length (global row = img * rows_img + oh * ow_full + ow)
};

constexpr int kHaloW = 130;  // 128 output pixels + one halo pixel on each side

struct SubTile { int img, oh0, ow0, sub_in_img; };
__device__ __forceinline__ SubTile decode_sub(const TcParams& p, int mtile, int sub) {
    SubTile s;
    if (p.h66) {
        // mtile -> (image, tile k of the pitch-66 space): oh0 / ow0 carry the first slot's row and slot-in-row
        // (CTA pairs take the same tile of two consecutive images: one descriptor offset must serve both CTAs)
        const int mq = p.pair ? (mtile >> 1) : mtile;
        const int k = mq % p.tiles_h;
        s.img = p.pair ? 2 * (mq / p.tiles_h) + (mtile & 1) : mq / p.tiles_h;
        const int s0 = k * 128;
        s.oh0 = s0 / 66;
        s.ow0 = s0 - s.oh0 * 66;
        s.sub_in_img = k;
    } else if (p.halo) {
        const int tw_i = mtile % p.tiles_w;
        const int r = mtile / p.tiles_w;
        const int hp_n = p.tiles_h / p.mt;
        s.img = r / hp_n;
        s.oh0 = (r - s.img * hp_n) * p.mt + sub;   // TH == 1: one output row per sub-tile
        s.ow0 = tw_i * p.TW;
        s.sub_in_img = s.oh0 * p.tiles_w + tw_i;
    } else {
        int st = mtile * p.mt + sub;
        s.sub_in_img = st % (p.tiles_h * p.tiles_w);
        const int tw_i = st % p.tiles_w; st /= p.tiles_w;
        const int th_i = st % p.tiles_h;
        s.img = st / p.tiles_h;
        s.ow0 = tw_i * p.TW;
        s.oh0 = th_i * p.TH;
    }
    return s;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// 256-bit global accesses (sm_100: LDG/STG.256): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Epilogue for NC (<=32) accumulator columns held by one thread (= one output pixel): alpha, bias,
// residual (already in registers: rres, loaded one chunk ahead), store.  On return f[] holds the
// values as stored (bf16-rounded for bf16 outputs).
template <int NC>
__device__ __forceinline__ void epilogue_store(const TcParams& p, const uint32_t* v, float (&f)[NC],
                                               const uint4 (&rres)[4], bool valid, long long d_off, int n0, float alpha) {
    if (alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < NC; ++j) f[j] = __uint_as_float(v[j]) * alpha;
    } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) f[j] = __uint_as_float(v[j]);
    }
    if (!valid) return;
    if (p.bias != nullptr) {
        // n0 is a multiple of 16: 16-byte loads (the same addresses for every lane: one L1 line per instruction)
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
        }
    }
    if (p.resid != nullptr) {
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            const uint4 r = rres[j];
            f[8 * j + 0] += bf16_lo(r.x); f[8 * j + 1] += bf16_hi(r.x);
            f[8 * j + 2] += bf16_lo(r.y); f[8 * j + 3] += bf16_hi(r.y);
            f[8 * j + 4] += bf16_lo(r.z); f[8 * j + 5] += bf16_hi(r.z);
            f[8 * j + 6] += bf16_lo(r.w); f[8 * j + 7] += bf16_hi(r.w);
        }
    }
    if (p.dbg_no_epi == 2) return;
    if (p.D_sN == 1 && n0 + NC <= p.n_store && p.beta == 0.f) {
        if (p.out_fp32) {
            float* dp = reinterpret_cast<float*>(p.D) + d_off + n0;
#pragma unroll
            for (int j = 0; j < NC / 8; ++j)
                st_global_256(dp + 8 * j,
                              make_uint4(__float_as_uint(f[8 * j]), __float_as_uint(f[8 * j + 1]), __float_as_uint(f[8 * j + 2]), __float_as_uint(f[8 * j + 3])),
                              make_uint4(__float_as_uint(f[8 * j + 4]), __float_as_uint(f[8 * j + 5]), __float_as_uint(f[8 * j + 6]), __float_as_uint(f[8 * j + 7])));
        } else {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + d_off + n0;
            uint4 o[NC / 8];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
                o[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                  pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                if (p.gn_mode != 0) {   // the fused reductions see the values as stored
                    f[8 * j + 0] = bf16_lo(o[j].x); f[8 * j + 1] = bf16_hi(o[j].x); f[8 * j + 2] = bf16_lo(o[j].y); f[8 * j + 3] = bf16_hi(o[j].y);
                    f[8 * j + 4] = bf16_lo(o[j].z); f[8 * j + 5] = bf16_hi(o[j].z); f[8 * j + 6] = bf16_lo(o[j].w); f[8 * j + 7] = bf16_hi(o[j].w);
                }
            }
            // full 32-byte sectors per lane (a 16-byte store would leave every sector half written)
#pragma unroll
            for (int j = 0; j < NC / 16; ++j) st_global_256(dp + 16 * j, o[2 * j], o[2 * j + 1]);
        }
    } else {
        // strided / partially stored columns (e.g. fp32 NCHW moments, transposed operands)
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if (n0 + j < p.n_store) {
                long long o = d_off + (long long)(n0 + j) * p.D_sN;
                if (p.out_fp32) {
                    float* dp = reinterpret_cast<float*>(p.D) + o;
                    *dp = p.beta != 0.f ? fmaf(p.beta, *dp, f[j]) : f[j];
                } else {
                    reinterpret_cast<__nv_bfloat16*>(p.D)[o] = __float2bfloat16_rn(f[j]);
                }
            }
        }
    }
}

// 16 per-lane values -> their sums over the 32 lanes of the warp in 16 shuffles (halving butterfly).
// On return, lane L holds in v[0] the total of value index ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
// The order of additions is fixed, so the result is bitwise reproducible.
__device__ __forceinline__ void warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Per-thread partial sums over one 32-column chunk of one output row, per GroupNorm group:
//   mode 1: gv[2g] = sum f, gv[2g+1] = sum f^2            (statistics of the tensor just produced)
//   mode 2: gv[2g] = sum dxh, gv[2g+1] = sum dxh*xh        (GroupNorm backward reductions; f = dy)
// In mode 2 the inner loop accumulates sum(dxh) and sum(dxh*x); xh = (x - mean)*rstd is applied once
// per group afterwards: sum(dxh*xh) = rstd*(sum(dxh*x) - mean*sum(dxh)).
// CPG is a template parameter so gv[] stays in registers.
template <int CPG>
__device__ __forceinline__ void gn_chunk_sums(const TcParams& p, const float (&f)[32], float (&gv)[16], bool valid,
                                              const uint4 (&xreg)[4], int c, const float* gn_sc, const float* gn_sh,
                                              const float* gn_gm, const float2* gn_mrs) {
#pragma unroll
    for (int i = 0; i < 16; ++i) gv[i] = 0.f;
    if (!valid) return;
    if (p.gn_mode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            gv[2 * (j / CPG)] += f[j];
            gv[2 * (j / CPG) + 1] = fmaf(f[j], f[j], gv[2 * (j / CPG) + 1]);
        }
    } else {
        const bool silu = p.gn_silu != 0;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
            const uint4 xr = xreg[j8];
            const float xs[8] = {bf16_lo(xr.x), bf16_hi(xr.x), bf16_lo(xr.y), bf16_hi(xr.y),
                                 bf16_lo(xr.z), bf16_hi(xr.z), bf16_lo(xr.w), bf16_hi(xr.w)};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j8 * 8 + jj;
                float dxh = f[j] * gn_gm[c + j];
                if (silu) {
                    const float u = fmaf(xs[jj], gn_sc[c + j], gn_sh[c + j]);
                    const float sg = __fdividef(1.f, 1.f + __expf(-u));
                    dxh *= sg * fmaf(u, 1.f - sg, 1.f);
                }
                gv[2 * (j / CPG)] += dxh;
                gv[2 * (j / CPG) + 1] = fmaf(dxh, xs[jj], gv[2 * (j / CPG) + 1]);
            }
        }
#pragma unroll
        for (int g = 0; g < 32 / CPG; ++g) {
            const float2 m = gn_mrs[c / CPG + g];
            gv[2 * g + 1] = m.y * fmaf(-m.x, gv[2 * g], gv[2 * g + 1]);
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


// ------------------------------------------------------------------------------------------------
// Lean epilogues (EPI != 0): compile-time specialised for the short-K GEMMs whose run time is the epilogue
// (1x1 shortcuts, stride-2 parity dgrads, attention projections, conv_in tap products).  ncu of the generic
// epilogue on a K = 128 shortcut (profiles/r02_prof_thin_shortcut_base.txt): ~390 SASS instructions per warp and
// 32-column chunk -- runtime feature branches, operand-prefetch register shuffles, two 128-thread named barriers --
// and 8000 clk per 128 x 256 sub-tile where the MMAs need 1024.  Here a warp owns its 32 rows end to end:
// tcgen05.ld 32 columns -> alpha / bias (from shared memory) / residual -> pack -> its own [32 rows][32 cols]
// staging buffer (two of them) -> its own TMA store.  Only __syncwarp between the steps; ~70 instructions per chunk.
// ------------------------------------------------------------------------------------------------
// EPI_LEAN_ATTN = the bf16 lean epilogue with the row-wise attention modes (GemmOp::epi_mode 1..3) compiled in; the plain
// instantiations carry none of that code (sharing one instantiation cost the 1x1 shortcuts 27 %, measured).
constexpr int EPI_GENERIC = 0, EPI_LEAN_BF16 = 1, EPI_LEAN_F32 = 2, EPI_LEAN_ATTN = 3;

template <int EPI, bool PAIR>
__device__ __forceinline__ void lean_epilogue(const TcParams& p, const CUtensorMap* mapD, uint8_t* out_stage, float* bias_s,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base, int tile0,
                                              int tile_step, int total_tiles, uint32_t crank, volatile int* hw) {
    constexpr int ES = EPI == EPI_LEAN_F32 ? 4 : 2;
    constexpr int ROWB = 32 * ES;                       // bytes of one staged row (32 columns)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = warp - 4, q = e & 3, half = e >> 2;   // TMEM lane quadrant (rows 32q..32q+31), column half
    const int et = int(threadIdx.x) - 128;
    const int row = q * 32 + lane;
    const bool warp_valid = q * 32 < p.rows_valid;      // rows_valid is a multiple of 32 here (host checks)
    const int nch = p.BN >> 5;
    const int ch_lo = half == 0 ? 0 : (nch + 1) / 2, ch_hi = half == 0 ? (nch + 1) / 2 : nch;
    const int acc_cols = p.mt * p.BN;
    uint8_t* const sbase = out_stage + size_t(e) * (2 * 32 * ROWB);
    // the warp's 32 rows inside the tile's TW x TH pixel rectangle: 32 pixels of one row (TW >= 32) or 32/TW rows
    const int wr = (q * 32) / p.TW, wc = q * 32 - wr * p.TW;
    const int r_th = row / p.TW, r_tw = row - r_th * p.TW;
    const float alpha = p.alpha;
    const bool has_bias = p.bias != nullptr, has_res = p.resid != nullptr;
    const int mode = EPI == EPI_LEAN_ATTN ? p.epi_mode : 0;   // 0 plain, 1 row max, 2 exp, 3 softmax backward
    const int htag = int(crank) * 100;
    int acc = 0, buf = 0, bias_nt = -1;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int nt = tile % p.n_tiles;
        const int mtile = PAIR ? (tile / p.n_tiles) * 2 + int(crank) : tile / p.n_tiles;
        if (has_bias && nt != bias_nt) {   // this tile's bias columns -> shared memory (warp-uniform branch)
            named_bar_sync(1, kEpiThreads);
            for (int c = et; c < p.BN; c += kEpiThreads) bias_s[c] = __ldg(p.bias + nt * p.BN + c);
            named_bar_sync(1, kEpiThreads);
            bias_nt = nt;
        }
        mbar_wait(&tfull_bar[acc], acc_phase, hw, htag + 6);
        tc_fence_after();
        if (warp_valid) {
            for (int sub = 0; sub < p.mt; ++sub) {
                const SubTile stl = decode_sub(p, mtile, sub);
                const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * acc_cols + sub * p.BN);
                const __nv_bfloat16* rrow = nullptr;
                if (has_res)
                    rrow = p.resid + (long long)stl.img * p.R_sB + (long long)(stl.oh0 + r_th) * p.R_sH +
                           (long long)(stl.ow0 + r_tw) * p.R_sW + nt * p.BN;
                // row-wise modes: this thread's row constants / running reduction
                const long long grow = (long long)stl.img * p.rows_img + (long long)(stl.oh0 + r_th) * p.ow_full + (stl.ow0 + r_tw);
                float ra = 0.f, rb = 1.f, red = mode == 1 ? -INFINITY : 0.f;
                float al = alpha;                                         // plain mode: alpha x the per-row output scale
                if (mode == 0 && p.row_scale != nullptr) al *= __ldg(p.row_scale + grow);
                if (mode >= 2) ra = __ldg(p.row_a + grow);
                if (mode == 2) ra *= -p.exp_scale;                       // exp2(acc * c - m * c)
                if (mode == 3) rb = alpha * __ldg(p.row_b + grow);
                uint32_t v[32];
                if (ch_lo < ch_hi) tmem_ld32(t_addr + uint32_t(ch_lo * 32), v);
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    const int c = ch * 32;
                    uint4 rr[4];
                    if (has_res) {
                        ld_global_nc_256(rrow + c, rr[0], rr[1]);
                        ld_global_nc_256(rrow + c + 16, rr[2], rr[3]);
                    }
                    if (mode != 1) {
                        if (lane == 0) bulk_wait_group_read<1>();   // the store that last read this buffer is done with it
                        __syncwarp();
                    }
                    tmem_ld_wait();
                    float f[32];
                    if (mode == 0) {
                        if (has_bias) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 b4 = lds_f4(smem_u32(bias_s + c + 4 * j));
                                f[4 * j] = fmaf(__uint_as_float(v[4 * j]), al, b4.x);
                                f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), al, b4.y);
                                f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), al, b4.z);
                                f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), al, b4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * al;
                        }
                    } else if (mode == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) red = fmaxf(red, __uint_as_float(v[j]));
                    } else if (mode == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = ex2_approx(fmaf(__uint_as_float(v[j]), p.exp_scale, ra));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = (__uint_as_float(v[j]) - ra) * rb;
                    }
                    // the accumulator registers are consumed: the next chunk's TMEM load runs under the rest of this one
                    if (ch + 1 < ch_hi) tmem_ld32(t_addr + uint32_t(c + 32), v);
                    if (mode == 1) continue;
                    if (has_res) {
                        if (mode == 3) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[8 * j + 0] *= bf16_lo(rr[j].x); f[8 * j + 1] *= bf16_hi(rr[j].x);
                                f[8 * j + 2] *= bf16_lo(rr[j].y); f[8 * j + 3] *= bf16_hi(rr[j].y);
                                f[8 * j + 4] *= bf16_lo(rr[j].z); f[8 * j + 5] *= bf16_hi(rr[j].z);
                                f[8 * j + 6] *= bf16_lo(rr[j].w); f[8 * j + 7] *= bf16_hi(rr[j].w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[8 * j + 0] += bf16_lo(rr[j].x); f[8 * j + 1] += bf16_hi(rr[j].x);
                                f[8 * j + 2] += bf16_lo(rr[j].y); f[8 * j + 3] += bf16_hi(rr[j].y);
                                f[8 * j + 4] += bf16_lo(rr[j].z); f[8 * j + 5] += bf16_hi(rr[j].z);
                                f[8 * j + 6] += bf16_lo(rr[j].w); f[8 * j + 7] += bf16_hi(rr[j].w);
                            }
                        }
                    }
                    uint8_t* sb = sbase + size_t(buf) * (32 * ROWB);
                    uint8_t* rowp = sb + size_t(lane) * ROWB;
                    if constexpr (EPI == EPI_LEAN_F32) {
                        // 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            sts_u4(smem_u32(rowp) + ((j ^ (lane & 7)) << 4),
                                   make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                              __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3])));
                    } else {
                        // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 o = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                                       pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                            sts_u4(smem_u32(rowp) + ((j ^ ((lane >> 1) & 3)) << 4), o);
                            if (mode == 2)   // the softmax denominator sums the probabilities as stored (bf16), in a fixed order
                                red += ((bf16_lo(o.x) + bf16_hi(o.x)) + (bf16_lo(o.y) + bf16_hi(o.y))) +
                                       ((bf16_lo(o.z) + bf16_hi(o.z)) + (bf16_lo(o.w) + bf16_hi(o.w)));
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(mapD, sb, nt * p.BN + c, stl.ow0 + wc, stl.oh0 + wr, stl.img);
                        bulk_commit_group();
                    }
                    buf ^= 1;
                }
                if (mode == 1 || mode == 2)   // (a column half without chunks writes the identity)
                    p.row_part[(grow * p.n_tiles + nt) * 2 + half] = red;
            }
        }
        tc_fence_before();
        if constexpr (PAIR) {
            if (crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);   // the leader's MMA warp owns both accumulators
            else mbar_arrive(&tempty_bar[acc]);
        } else {
            mbar_arrive(&tempty_bar[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) bulk_wait_group<0>();   // staged tiles are in global memory before the CTA exits
}

// PAIR = true is the cta_group::2 instantiation (launched as clusters of two CTAs); the PAIR = false
// instantiation contains no cluster instruction and is launched as an ordinary grid.  EPI selects the epilogue
// (EPI_GENERIC: every feature behind runtime switches; EPI_LEAN_*: see above, single-CTA only).
template <bool PAIR, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                         const __grid_constant__ CUtensorMap mapD, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* out_stage = smem + 2 * size_t(p.halo_bytes);              // (halo mode: two halo tiles come first)
    uint8_t* ring = out_stage + size_t(p.out_bytes);                   // output staging of the TMA-store epilogue, then the ring
    uint8_t* bar_base = ring + size_t(p.stages) * p.stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);        // [stages]
    uint64_t* empty_bar = full_bar + 8;                                // [stages]
    uint64_t* tfull_bar = empty_bar + 8;                               // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
    uint64_t* hfull_bar = tempty_bar + 2;                              // [2]
    uint64_t* hempty_bar = hfull_bar + 2;                              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty_bar + 2);
    // GN scratch (epilogue warps only): red[4 warps][8 chunks][16] floats, then per-tile constants
    float* gn_red = reinterpret_cast<float*>(bar_base + kBarrierBytes);          // 2048 B
    float* gn_sc = gn_red + 4 * 8 * 16;                                          // [256]
    float* gn_sh = gn_sc + 256;                                                  // [256]
    float* gn_gm = gn_sh + 256;                                                  // [256]
    float2* gn_mrs = reinterpret_cast<float2*>(gn_gm + 256);                     // [64]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (p.out_bytes) tma_prefetch_desc(&mapD);
    }
    if (warp == 1 && lane == 0) {
        // pair mode: the "full" barriers live in the leader CTA: ONE arrival (the leader's) plus the transaction
        // bytes of both CTAs' TMA loads -- a per-stage remote arrive by the follower's producer throttles it to one
        // stage per cross-SM round trip; the accumulator-empty barrier collects both epilogues
        const uint32_t nprod = PAIR ? 2u : 1u;
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);      // pair mode: the leader's arrival announces both CTAs' bytes (see below)
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], kEpiThreads * nprod);
            mbar_init(&hfull_bar[a], 1);
            mbar_init(&hempty_bar[a], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before anyone signals across
    tc_fence_after();
    // broadcast through a shuffle so the compiler knows the value is warp-uniform (keeps it in a uniform register)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    uint32_t crank = 0u;
    if constexpr (PAIR) crank = cluster_ctarank();
    volatile int* hw = p.hang_where;
    const int htag = int(crank) * 100;

    const int sub_per_img = p.tiles_h * p.tiles_w;
    const int m_tiles = (p.nimg * sub_per_img) / p.mt;   // (halo: sub-tiles of a CTA tile are consecutive rows)
    const int total_tiles = PAIR ? (m_tiles / 2) * p.n_tiles : m_tiles * p.n_tiles;   // pair mode: tiles of the PAIR
    const int tile0 = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int tile_step = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
    const int kblocks = p.ntaps * p.kchunks;
    const int a_bytes = p.mt * kATileBytes;
    const uint32_t tx_bytes = uint32_t(p.mt) * uint32_t(p.rows_valid + (p.dbg_shift ? 8 : 0)) * 128u + uint32_t(p.BN) * 128u;
    const int acc_cols = p.mt * p.BN;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int nt = tile % p.n_tiles;
                const int mtile = PAIR ? (tile / p.n_tiles) * 2 + int(crank) : tile / p.n_tiles;
                if (p.halo) {
                    // B (weight) tiles only, in (chunk, tap) order; the halo tiles come from warp 3
                    for (int ch = 0; ch < p.kchunks; ++ch)
                        for (int tap = 0; tap < p.ntaps; ++tap) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u, hw, htag + 1);
                            if constexpr (PAIR) {
                                if (p.dbg_mma_only && (phase != 0 || tile != tile0)) {
                                    if (crank == 0) mbar_arrive(&full_bar[stage]);
                                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                                    continue;
                                }
                                // each CTA stages its half of the weight tile; bytes are counted on the leader's barrier
                                if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], uint32_t(p.BN) * 128u);
                                tma_load_3d_2sm(ring + size_t(stage) * p.stage_bytes, &mapB, &full_bar[stage],
                                                (tap * p.kchunks + ch) * kBlockK, nt * p.BN + int(crank) * (p.BN / 2), 0);
                                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                                continue;
                            }
                            if (p.dbg_mma_only && (phase != 0 || tile != tile0)) { mbar_arrive(&full_bar[stage]); if (++stage == p.stages) { stage = 0; phase ^= 1u; } continue; }
                            mbar_arrive_expect_tx(&full_bar[stage], uint32_t(p.BN) * 128u);
                            tma_load_3d(ring + size_t(stage) * p.stage_bytes, &mapB, &full_bar[stage],
                                        (tap * p.kchunks + ch) * kBlockK, nt * p.BN, 0);
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                    continue;
                }
                SubTile sb[2];
                for (int sub = 0; sub < p.mt; ++sub) sb[sub] = decode_sub(p, mtile, sub);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    const int tap = kb / p.kchunks;
                    const int c0 = (kb - tap * p.kchunks) * kBlockK;
                    uint8_t* sA = ring + size_t(stage) * p.stage_bytes;
                    uint8_t* sB = sA + a_bytes;
                    if constexpr (PAIR) {
                        // pairs (one sub-tile): each CTA stages its own 128 pixels (its half of M = 256) and
                        // its half of the weight tile; the leader alone arrives, announcing both CTAs' bytes
                        if (p.dbg_mma_only && (phase != 0 || tile != tile0)) {
                            if (crank == 0) mbar_arrive(&full_bar[stage]);
                        } else {
                            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * uint32_t(p.rows_valid) * 128u + uint32_t(p.BN) * 128u);
                            if (p.a_trans) {
                                // this CTA's 128 rows of the transposed operand: two 64-row x 64-k atoms
                                const int m0 = sb[0].oh0 * p.ow_full + sb[0].ow0;
                                tma_load_3d_2sm(sA, &mapA, &full_bar[stage], m0, c0, sb[0].img);
                                tma_load_3d_2sm(sA + 8192, &mapA, &full_bar[stage], m0 + 64, c0, sb[0].img);
                            } else if (p.mode == 0) {
                                tma_load_4d_2sm(sA, &mapA, &full_bar[stage], c0, sb[0].ow0 + p.dw[tap], sb[0].oh0 + p.dh[tap], sb[0].img);
                            } else {
                                const int dw = p.dw[tap], dh = p.dh[tap];
                                for (int i = 0; i < p.TH; ++i)
                                    tma_load_5d_2sm(sA + size_t(i) * p.TW * 128, &mapA, &full_bar[stage], c0, dw & 1,
                                                    sb[0].ow0 + (dw >> 1), 2 * (sb[0].oh0 + i) + dh, sb[0].img);
                            }
                            if (p.b_prefetch > 0 && kb + p.b_prefetch < kblocks)
                                tma_prefetch_l2_3d(&mapB, (kb + p.b_prefetch) * kBlockK, nt * p.BN + int(crank) * (p.BN / 2), 0);
                            tma_load_3d_2sm(sB, &mapB, &full_bar[stage], kb * kBlockK, nt * p.BN + int(crank) * (p.BN / 2),
                                            p.b_batched ? sb[0].img : 0);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (p.dbg_mma_only && (phase != 0 || tile != tile0)) { mbar_arrive(&full_bar[stage]); if (++stage == p.stages) { stage = 0; phase ^= 1u; } continue; }
                    mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                    for (int sub = 0; sub < p.mt; ++sub) {
                        uint8_t* dst = sA + sub * kATileBytes;
                        if (p.a_trans) {
                            // 64-row (m) x 64-k atoms column by column: rows m0 .. m0+63 at dst, m0+64 .. at dst + 8 KB
                            const int m0 = sb[sub].oh0 * p.ow_full + sb[sub].ow0;
                            for (int j = 0; j < (p.rows_valid >> 6); ++j)
                                tma_load_3d(dst + j * 8192, &mapA, &full_bar[stage], m0 + 64 * j, c0, sb[sub].img);
                        } else if (p.mode == 0) {
                            tma_load_4d(dst, &mapA, &full_bar[stage], c0, sb[sub].ow0 + p.dw[tap] - p.dbg_shift,
                                        sb[sub].oh0 + p.dh[tap], sb[sub].img);
                        } else {
                            const int dw = p.dw[tap], dh = p.dh[tap];
                            for (int i = 0; i < p.TH; ++i)
                                tma_load_5d(dst + size_t(i) * p.TW * 128, &mapA, &full_bar[stage], c0, dw & 1,
                                            sb[sub].ow0 + (dw >> 1), 2 * (sb[sub].oh0 + i) + dh, sb[sub].img);
                        }
                    }
                    if (p.b_prefetch > 0 && kb + p.b_prefetch < kblocks)
                        tma_prefetch_l2_3d(&mapB, (kb + p.b_prefetch) * kBlockK, nt * p.BN, 0);
                    tma_load_3d(sB, &mapB, &full_bar[stage], kb * kBlockK, nt * p.BN, p.b_batched ? sb[0].img : 0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================================== halo producer (halo mode)
        if (lane == 0 && p.halo) {
            int hs = 0;
            uint32_t hphase = 0;
            const uint32_t halo_tx = p.h66 ? 6u * 66u * 128u : uint32_t(p.mt + 2) * kHaloW * 128u;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const SubTile s0 = decode_sub(p, PAIR ? (tile / p.n_tiles) * 2 + int(crank) : tile / p.n_tiles, 0);
                for (int ch = 0; ch < p.kchunks; ++ch) {
                    mbar_wait(&hempty_bar[hs], hphase ^ 1u, hw, htag + 2);
                    if constexpr (PAIR) {
                        if (p.dbg_mma_only && (hphase != 0 || tile != tile0)) {
                            if (crank == 0) mbar_arrive(&hfull_bar[hs]);
                            hs ^= 1;
                            if (hs == 0) hphase ^= 1u;
                            continue;
                        }
                        if (crank == 0) mbar_arrive_expect_tx(&hfull_bar[hs], 2u * halo_tx);
                        tma_load_4d_2sm(smem + size_t(hs) * p.halo_bytes, &mapA, &hfull_bar[hs], ch * kBlockK,
                                        p.h66 ? -1 : s0.ow0 - 1, p.h66 ? s0.oh0 - 2 : s0.oh0 - 1, s0.img);
                        hs ^= 1;
                        if (hs == 0) hphase ^= 1u;
                        continue;
                    }
                    if (p.dbg_mma_only && (hphase != 0 || tile != tile0)) { mbar_arrive(&hfull_bar[hs]); hs ^= 1; if (hs == 0) hphase ^= 1u; continue; }
                    mbar_arrive_expect_tx(&hfull_bar[hs], halo_tx);
                    // rows oh0-1 .. oh0+mt, pixels ow0-1 .. ow0+128: out-of-range pixels arrive as zeros (= padding)
                    // (h66: pixels -1 .. 64 of rows r0-2 .. r0+3, r0 = row of the tile's first slot)
                    tma_load_4d(smem + size_t(hs) * p.halo_bytes, &mapA, &hfull_bar[hs], ch * kBlockK,
                                p.h66 ? -1 : s0.ow0 - 1, p.h66 ? s0.oh0 - 2 : s0.oh0 - 1, s0.img);
                    hs ^= 1;
                    if (hs == 0) hphase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The whole warp runs this loop convergently with warp-uniform values (descriptors, stage indices live in
        // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.  Issuing
        // from a divergent `if (lane == 0)` region costs ~16 instructions of R2UR/ELECT shuffling per MMA, which
        // made the N = 128 layers issue-bound (64 tensor cycles per MMA).
        if (crank == 0) {   // pair mode: only the leader CTA issues
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kUmmaM : kUmmaM, p.BN) | (p.a_trans ? kUmmaAMajorMN : 0u);
            // K = 16 step of the A descriptor (addr >> 4 units): 32 B inside the swizzle row, or two 1 KB K-groups
            const uint64_t a_step = p.a_trans ? 128u : 2u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0, hs = 0;
            uint32_t acc_phase = 0, hphase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, hw, htag + 3);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * acc_cols);
                if (p.halo) {
                    // h66: slot-in-row of the tile's first output slot (the same for both CTAs of a pair)
                    const int xi0 = p.h66 ? decode_sub(p, PAIR ? (tile / p.n_tiles) * 2 : tile / p.n_tiles, 0).ow0 : 0;
                    for (int ch = 0; ch < p.kchunks; ++ch) {
                        mbar_wait(&hfull_bar[hs], hphase, hw, htag + 4);
                        tc_fence_after();
                        const uint32_t h_addr = smem_u32(smem + size_t(hs) * p.halo_bytes);
                        for (int tap = 0; tap < p.ntaps; ++tap) {
                            mbar_wait(&full_bar[stage], phase, hw, htag + 5);
                            tc_fence_after();
                            const uint64_t b_desc = umma_desc_sw128(smem_u32(ring + size_t(stage) * p.stage_bytes));
                            // operand rows = 128 consecutive halo pixels starting at (row sub+dh+1, pixel dw+1)
                            // (h66: the tile holds rows r0-2 .. r0+3 at a pitch of 66 slots, slot = x + 1)
                            const uint32_t row0 = p.h66 ? uint32_t((p.dh[tap] + 2) * 66 + xi0 + p.dw[tap])
                                                        : uint32_t((p.dh[tap] + 1) * kHaloW + p.dw[tap] + 1);
                            const uint64_t a_desc0 = umma_desc_sw128(h_addr + row0 * 128u);
                            const uint32_t first = (ch | tap) != 0 ? 1u : 0u;
                            if (elect_one()) {
                                for (int sub = 0; sub < p.mt; ++sub) {
                                    // next output row = next halo row: +kHaloW rows of 128 B = +kHaloW*8 in the addr>>4 field
                                    const uint64_t a_desc = a_desc0 + uint64_t(sub * kHaloW * 8);
                                    const uint32_t d = d_tmem + uint32_t(sub * p.BN);
                                    if constexpr (PAIR) {
                                        umma_bf16_2sm(d, a_desc, b_desc, idesc, first);
                                        umma_bf16_2sm(d, a_desc + 2, b_desc + 2, idesc, 1u);
                                        umma_bf16_2sm(d, a_desc + 4, b_desc + 4, idesc, 1u);
                                        umma_bf16_2sm(d, a_desc + 6, b_desc + 6, idesc, 1u);
                                    } else {
                                        umma_bf16(d, a_desc, b_desc, idesc, first);
                                        umma_bf16(d, a_desc + 2, b_desc + 2, idesc, 1u);
                                        umma_bf16(d, a_desc + 4, b_desc + 4, idesc, 1u);
                                        umma_bf16(d, a_desc + 6, b_desc + 6, idesc, 1u);
                                    }
                                }
                                if constexpr (PAIR) umma_commit_2sm(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
                                // halo tile free once its nine taps have retired
                                if (tap == p.ntaps - 1) {
                                    if constexpr (PAIR) umma_commit_2sm(&hempty_bar[hs], 3); else umma_commit(&hempty_bar[hs]);
                                }
                            }
                            __syncwarp();
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                        hs ^= 1;
                        if (hs == 0) hphase ^= 1u;
                    }
                } else {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase, hw, htag + 5);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(ring + size_t(stage) * p.stage_bytes);
                        const uint64_t b_desc = umma_desc_sw128(a_addr + a_bytes);
                        uint64_t a_desc0 = p.a_trans ? umma_desc_sw128_mn(a_addr, 8192u, 1024u)
                                                     : umma_desc_sw128(a_addr + uint32_t(p.dbg_shift) * 128u);
                        if (p.dbg_bo) a_desc0 |= uint64_t(((a_addr + uint32_t(p.dbg_shift) * 128u) >> 7) & 7u) << 49;
                        const uint32_t first = kb != 0 ? 1u : 0u;
                        if (elect_one()) {
                            if constexpr (PAIR) {
                                umma_bf16_2sm(d_tmem, a_desc0, b_desc, idesc, first);
                                umma_bf16_2sm(d_tmem, a_desc0 + a_step, b_desc + 2, idesc, 1u);
                                umma_bf16_2sm(d_tmem, a_desc0 + 2 * a_step, b_desc + 4, idesc, 1u);
                                umma_bf16_2sm(d_tmem, a_desc0 + 3 * a_step, b_desc + 6, idesc, 1u);
                                umma_commit_2sm(&empty_bar[stage], 3);
                            } else {
                                for (int sub = 0; sub < p.mt; ++sub) {
                                    const uint64_t a_desc = a_desc0 + uint64_t(sub * (kATileBytes >> 4));
                                    const uint32_t d = d_tmem + uint32_t(sub * p.BN);
                                    // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the addr>>4 field
                                    umma_bf16(d, a_desc, b_desc, idesc, first);
                                    umma_bf16(d, a_desc + a_step, b_desc + 2, idesc, 1u);
                                    umma_bf16(d, a_desc + 2 * a_step, b_desc + 4, idesc, 1u);
                                    umma_bf16(d, a_desc + 3 * a_step, b_desc + 6, idesc, 1u);
                                }
                                umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
                            }
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
                if (elect_one()) {   // accumulator complete -> epilogue
                    if constexpr (PAIR) umma_commit_2sm(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4 && EPI != EPI_GENERIC) {
        if constexpr (EPI != EPI_GENERIC)
            lean_epilogue<EPI, PAIR>(p, &mapD, out_stage, gn_red, tfull_bar, tempty_bar, tmem_base, tile0, tile_step,
                                     total_tiles, crank, hw);
    } else if (warp >= 4) {
        // ===================================================================== epilogue
        // 8 warps: warp e handles TMEM lane quadrant e % 4 (rows 32q .. 32q+31 of the sub-tile) and
        // column half e / 4 of the tile, so every SM sub-partition runs two epilogue warps.
        const int e = warp - 4;
        const int q = e & 3;       // == warp % 4: this warp may touch TMEM lanes [32q, 32q+32)
        const int half = e >> 2;
        const int et = threadIdx.x - 128;  // 0..255 within the epilogue warps
        const int row = q * 32 + lane;
        const bool row_ok = row < p.rows_valid;
        const int r_th = row / p.TW, r_tw = row - r_th * p.TW;
        const int cpg = p.gn_cpg;
        const int nch = p.BN >> 5;                      // 32-column chunks in the tile (0 when BN == 16)
        const int ch_lo = half == 0 ? 0 : (nch + 1) / 2;  // this warp's chunk range
        const int ch_hi = half == 0 ? (nch + 1) / 2 : nch;
        const bool ld_res = p.resid != nullptr && p.dbg_no_epi != 2, ld_x = p.gn_mode == 2 && p.dbg_no_epi != 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const int nt = tile % p.n_tiles;
            const int mtile = PAIR ? (tile / p.n_tiles) * 2 + int(crank) : tile / p.n_tiles;
            bool waited = false;
            if (p.dbg_no_epi == 1) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
            }
            for (int sub = 0; sub < p.mt && p.dbg_no_epi != 1; ++sub) {
                const SubTile stl = decode_sub(p, mtile, sub);
                const int sub_in_img = stl.sub_in_img, img = stl.img;
                int oh = stl.oh0 + r_th, ow = stl.ow0 + r_tw;
                bool valid = row_ok;
                if (p.h66) {   // row = slot of the pitch-66 space: slot 0 / 65 of every row is a zero column (no output)
                    const int sl = stl.ow0 + row;
                    const int dr = sl / 66, xi = sl - dr * 66;
                    oh = stl.oh0 + dr; ow = xi - 1;
                    valid = xi >= 1 && xi <= 64;
                }
                const long long d_off = (long long)img * p.D_sB + (long long)oh * p.D_sH + (long long)ow * p.D_sW;
                const long long r_off = (long long)img * p.R_sB + (long long)oh * p.R_sH + (long long)ow * p.R_sW;
                const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * acc_cols + sub * p.BN);
                float alpha_row = p.alpha;   // alpha x the per-row scale (attention PV: 1 / softmax denominator)
                if (p.row_scale != nullptr && valid)
                    alpha_row *= __ldg(p.row_scale + (long long)img * p.rows_img + (long long)oh * p.ow_full + ow);

                // operands of the first chunk are requested before waiting for the accumulator
                uint4 rres[4], xreg[4];
                if (valid && ch_lo < ch_hi) {
                    const int n0 = nt * p.BN + ch_lo * 32;
                    if (ld_res) {
                        ld_global_nc_256(p.resid + r_off + n0, rres[0], rres[1]);
                        ld_global_nc_256(p.resid + r_off + n0 + 16, rres[2], rres[3]);
                    }
                    if (ld_x) {
                        ld_global_nc_256(p.gn_x + d_off + n0, xreg[0], xreg[1]);
                        ld_global_nc_256(p.gn_x + d_off + n0 + 16, xreg[2], xreg[3]);
                    }
                }
                if (p.gn_mode != 0) named_bar_sync(1, kEpiThreads);  // previous readers of the scratch are done
                if (p.gn_mode == 2) {
                    // per-tile constants of the GroupNorm being differentiated: scale/shift/gamma per channel,
                    // mean/rstd per group, for this image and this tile's channel range
                    for (int c = et; c < p.BN; c += kEpiThreads) {
                        const int n = nt * p.BN + c;
                        const float2 s2 = __ldg(&p.gn_ss[(size_t)img * (p.n_tiles * p.BN) + n]);
                        gn_sc[c] = s2.x; gn_sh[c] = s2.y;
                        gn_gm[c] = __ldg(&p.gn_gamma[n]);
                    }
                    for (int g = et; g < p.BN / cpg; g += kEpiThreads)
                        gn_mrs[g] = __ldg(&p.gn_mr[(size_t)img * 32 + (nt * p.BN) / cpg + g]);
                    named_bar_sync(1, kEpiThreads);
                }
                if (!waited) {
                    mbar_wait(&tfull_bar[acc], acc_phase, hw, htag + 6);
                    tc_fence_after();
                    waited = true;
                }

                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    const int c = ch * 32;
                    uint32_t v[32];
                    float f[32];
                    tmem_ld32(t_addr + uint32_t(c), v);
                    // operands of the next chunk go in flight while this one is processed
                    uint4 rnext[4], xnext[4];
                    if (valid && ch + 1 < ch_hi) {
                        const int n1 = nt * p.BN + c + 32;
                        if (ld_res) {
                            ld_global_nc_256(p.resid + r_off + n1, rnext[0], rnext[1]);
                            ld_global_nc_256(p.resid + r_off + n1 + 16, rnext[2], rnext[3]);
                        }
                        if (ld_x) {
                            ld_global_nc_256(p.gn_x + d_off + n1, xnext[0], xnext[1]);
                            ld_global_nc_256(p.gn_x + d_off + n1 + 16, xnext[2], xnext[3]);
                        }
                    }
                    tmem_ld_wait();
                    epilogue_store<32>(p, v, f, rres, valid, d_off, nt * p.BN + c, alpha_row);
                    if (p.gn_mode != 0 && p.dbg_no_epi != 3) {
                        float gv[16];
                        if (cpg == 4) gn_chunk_sums<4>(p, f, gv, valid, xreg, c, gn_sc, gn_sh, gn_gm, gn_mrs);
                        else if (cpg == 8) gn_chunk_sums<8>(p, f, gv, valid, xreg, c, gn_sc, gn_sh, gn_gm, gn_mrs);
                        else gn_chunk_sums<16>(p, f, gv, valid, xreg, c, gn_sc, gn_sh, gn_gm, gn_mrs);
                        warp_reduce16(gv, lane);
                        if ((lane & 1) == 0)
                            gn_red[(q * 8 + ch) * 16 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 +
                                   ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = gv[0];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) { rres[j] = rnext[j]; xreg[j] = xnext[j]; }
                }
                if (nch == 0 && half == 0) {  // BN == 16 (padded thin outputs)
                    uint32_t v[16];
                    float f[16];
                    tmem_ld16(t_addr, v);
                    tmem_ld_wait();
                    epilogue_store<16>(p, v, f, rres, valid, d_off, nt * p.BN, alpha_row);
                }
                if (p.gn_mode != 0) {
                    // cross-warp combine in fixed order, one (group, value) per thread: the chunk of a group
                    // was handled by the four row-quadrant warps of one column half
                    named_bar_sync(1, kEpiThreads);
                    const int gpc = 32 / cpg;                 // groups per 32-column chunk
                    const int ngroups = p.BN / cpg;           // groups in this tile
                    if (et < ngroups * 2) {
                        const int g = et >> 1, which = et & 1;
                        const int chunk = g / gpc, gi = g - chunk * gpc;
                        float tot = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) tot += gn_red[(w * 8 + chunk) * 16 + 2 * gi + which];
                        p.gn_partial[(((size_t)img * sub_per_img + sub_in_img) * 32 + (nt * p.BN) / cpg + g) * 2 + which] = tot;
                    }
                }
            }
            tc_fence_before();
            if constexpr (PAIR) {
                if (crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);   // the leader's MMA warp owns both accumulators
                else mbar_arrive(&tempty_bar[acc]);
            } else {
                mbar_arrive(&tempty_bar[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // nobody leaves while the peer may still signal its barriers / read its TMEM
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// "Swapped" 3x3 stride-1 convolution for exactly 128 output channels on wide rows (OW % 256 == 0).
//
// tcgen05.mma at N = 128 tops out ~27 % below N = 256 (measured MMA-only ceilings 1160 vs 1590 TFLOP/s), and the
// 128-channel layers at full resolution are a third of all GEMM time.  N is the weight dimension in the kernel
// above, so here the operands trade places:  D^T[channel][pixel] = W[channel][k] * X[pixel][k]^T  with
//   A (M = 128)  = one (tap, 64-channel chunk) slice of the packed weights          -- a ring of 16 KB stages
//   B (N = 256)  = 256 consecutive pixels of ONE input row, 64 channels              -- a ring of 258-pixel rows
// A tile is 256 pixels of one output row.  Each input row (y-1, y, y+1) x chunk is one ring slot of 258 pixels
// (the tile plus one halo pixel each side, zero-filled by TMA outside the image); the three horizontal taps read
// it through descriptors whose start is shifted by 0 / 128 / 256 bytes (same mechanism as halo mode above).
// The accumulator is [128 lanes = channels][256 columns = pixels], double buffered (2 x 256 TMEM columns).
// Epilogue: thread = channel, registers = 32 consecutive pixels; a warp's store of one pixel is 32 channels =
// 64 contiguous bytes (two full sectors).  GroupNorm statistics are per channel = per thread, so the fused
// reduction is a private accumulation plus two shuffles per group: no shared memory, no named barriers.
// ------------------------------------------------------------------------------------------------
// Geometry of the two instantiations.  Single CTA: a tile is 256 pixels of one row (a 258-pixel slot per input row).
// CTA pair (cta_group::2, for rows of only 128 pixels): the pair computes 256 channels x (2 rows x 128 pixels) with
// one M = 256, N = 256 MMA stream issued by the leader; CTA r stages channels [128r, 128r+128) of the weight slice
// (its half of A) and input row 2*yp + r (+ halo, a 130-pixel slot: its half of B).  The two 128-pixel halves of the
// N operand live in DIFFERENT shared memories, which is what makes a two-row tile expressible at all: inside one
// CTA, rows of 128+2 pixels cannot be 128 pixels apart.
#ifndef TML_SW_ROWSLOTS     // ring depths of the single-CTA geometry (A/B builds)
#define TML_SW_ROWSLOTS 3
#endif
#ifndef TML_SW_WSTAGES
#define TML_SW_WSTAGES 7
#endif
template <bool PAIR> struct SwGeo {
    static constexpr int kRowPx = PAIR ? 130 : 258;
    static constexpr int kRowBytes = PAIR ? 17 * 1024 : 33 * 1024;   // kRowPx * 128 B rounded up to the 1 KB swizzle atom
    static constexpr int kRowSlots = PAIR ? 4 : TML_SW_ROWSLOTS;
    static constexpr int kWStages = PAIR ? 8 : TML_SW_WSTAGES;
    static constexpr int kSmem = 1024 + kRowSlots * kRowBytes + kWStages * (128 * kBlockK * 2) + 512;
};
constexpr int kSwWBytes = 128 * kBlockK * 2;  // 16 KB: 128 output channels x 64 input channels
static_assert(SwGeo<false>::kSmem <= kMaxSmem && SwGeo<true>::kSmem <= kMaxSmem, "swapped-conv shared memory");

struct SwParams {
    int H, W, nimg, kchunks;
    int tap_of[3][3];            // packed-weight tap index of (row dh + 1, column dw + 1)
    const float* bias;           // [N] or null
    const __nv_bfloat16* resid;  // dense NHWC [nimg][H][W][N] or null
    __nv_bfloat16* D;            // dense NHWC [nimg][H][W][N]
    int gn_mode;                 // 0 none, 1 (sum, sumsq) of the output, 2 GroupNorm-backward sums (see GemmOp)
    float* gn_partial;           // [nimg][H * W/128 (mode 1) or H * W/64 (mode 2)][32][2]
    const __nv_bfloat16* gn_x;   // mode 2: the GroupNorm input, same layout as D
    const float2* gn_ss;         // mode 2: [nimg][N] (scale, shift)
    const float2* gn_mr;         // mode 2: [nimg][32] (mean, rstd)
    const float* gn_gamma;       // mode 2: [N]
    int gn_silu;
    volatile int* hang_where;
    int dbg_no_epi, dbg_mma_only;
    int l2_prefetch;             // row producer pulls the NEXT tile's input rows into L2 while this tile runs
    const float2* in_ss;         // XF instantiations: [nimg][64 * kchunks] (scale, shift) of the GroupNorm feeding this conv
};

// The GNB instantiations run 16 epilogue warps (64 pixels each, 8-pixel steps, <= 102 registers) because their
// epilogue is instruction-bound (exp + rcp + ~18 ALU ops per element): four warps per SM sub-partition instead of two
// is what lets it finish inside the main loop of the next tile.  Their partial sums are per 64-pixel segment.
#ifndef TML_SW_EPI16
#define TML_SW_EPI16 0
#endif
template <bool GNB> struct SwCfg {
    static constexpr int kEpiWarps = (GNB || TML_SW_EPI16) ? 16 : 8;
    static constexpr int kThreadsSw = 128 + 32 * kEpiWarps;
    static constexpr int kPx = 256 / (kEpiWarps / 4);   // pixels per epilogue warp
    static constexpr int kStep = kPx / 8;               // pixels per tcgen05.ld
};

int sw_px_per_warp(bool gnb) { return gnb ? SwCfg<true>::kPx : SwCfg<false>::kPx; }

// What one CTA does for tile number `tile`.
struct SwTile {
    int img;
    int chan0;        // first of this CTA's 128 output channels
    int brow, bx0;    // the input row / first pixel behind this CTA's half of the N operand
    int orow, ox0;    // output row / pixel of accumulator column 0; column c is pixel (orow + c / cols_row, ox0 + c % cols_row)
};
template <int NC, bool PAIR>
__device__ __forceinline__ SwTile sw_decode(const SwParams& p, int tile, int crank) {
    SwTile t;
    if constexpr (PAIR) {
        constexpr int NCP = NC / 2;
        const int cp = tile % NCP, pt = tile / NCP;
        const int segs = p.W >> 7;
        const int xs = pt % segs, rem = pt / segs;
        const int hp = p.H >> 1;
        const int yp = rem % hp;
        t.img = rem / hp;
        t.chan0 = cp * 256 + crank * 128;
        t.brow = 2 * yp + crank; t.bx0 = xs << 7;
        t.orow = 2 * yp; t.ox0 = xs << 7;
    } else {
        const int ct = tile % NC, pt = tile / NC;
        const int tiles_x = p.W >> 8;
        const int xh = pt % tiles_x, rem = pt / tiles_x;
        t.img = rem / p.H;
        t.chan0 = ct * 128;
        t.brow = rem % p.H; t.bx0 = xh << 8;
        t.orow = t.brow; t.ox0 = t.bx0;
    }
    return t;
}

// NC: output channels = 128 * NC; RES: residual add; GNB: fused GroupNorm-backward reductions (gn_mode 2);
// PAIR: the cta_group::2 instantiation (launched as clusters of two CTAs; the others contain no cluster instruction).
// XF (CTA pairs, forward convolutions): the input is the RAW GroupNorm input; four extra warps rewrite every staged row
// in shared memory as silu(x * scale + shift) between its TMA arrival and the MMAs that read it, so the normalised
// activation never exists in HBM.  Budget on pairs: 3 x 130 x 64 elements per 64-channel chunk and CTA against 36 MMAs of
// 128 clk = 5.4 elements/clk/SM (one tanh.approx + ~5.5 issue slots each).  Correct (kernel suite + the full parity suite
// pass with it switched on) but slower than the separate pass it replaces: see gemm_fuses_input_gn.  Row slots then fill in two steps: TMA -> the CTA's
// own `rland` barrier -> transform warps -> one arrival per CTA (the follower's with release.cluster) on the
// leader's `rfull`, which the MMA warp waits on with acquire.cluster.
constexpr int kXfThreads = 256;   // eight transform warps: two per SM sub-partition, so that one hides the other's latencies
template <int NC, bool RES, bool GNB, bool PAIR, bool XF = false>
__global__ void __launch_bounds__(SwCfg<GNB>::kThreadsSw + (XF ? kXfThreads : 0), 1)
conv3x3_swapped_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapRow,
                       const __grid_constant__ CUtensorMap mapTail, const SwParams p) {
    using G = SwGeo<PAIR>;
    constexpr int kRowSlots = G::kRowSlots, kWStages = G::kWStages, kRowBytes = G::kRowBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* rows = smem;
    uint8_t* wring = rows + size_t(kRowSlots) * kRowBytes;
    uint64_t* wfull = reinterpret_cast<uint64_t*>(wring + size_t(kWStages) * kSwWBytes);   // [8]
    uint64_t* wempty = wfull + 8;     // [8]
    uint64_t* rfull = wempty + 8;     // [4]
    uint64_t* rempty = rfull + 4;     // [4]
    uint64_t* tfull = rempty + 4;     // [2]
    uint64_t* tempty = tfull + 2;     // [2]
    uint64_t* rland = tempty + 2;     // [4]  XF: "the raw row has landed in this CTA's slot"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rland + 4);
    static_assert(!XF || (PAIR && !GNB), "the fused input normalisation exists for forward convolutions on CTA pairs");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapRow);
        if constexpr (!PAIR) tma_prefetch_desc(&mapTail);
    }
    if (warp == 1 && lane == 0) {
        // pair mode: the "full" barriers live in the leader CTA: ONE arrival (the leader's, which also announces the
        // transaction bytes of both CTAs' loads); the follower only issues its TMA, whose bytes are counted on the
        // leader's barrier (a complete_tx may precede the expect_tx of its phase).  A per-stage remote arrive by the
        // follower's producer throttled it to one stage per cross-SM round trip (measured: 855 vs 1800 TFLOP/s
        // without loads).  The accumulator-empty barrier collects both epilogues.
        const uint32_t nprod = PAIR ? 2u : 1u;
        for (int s = 0; s < kWStages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < kRowSlots; ++s) { mbar_init(&rfull[s], XF ? 2 : 1); mbar_init(&rempty[s], 1); mbar_init(&rland[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], nprod * 32 * SwCfg<GNB>::kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    volatile int* hw = p.hang_where;
    uint32_t crank = 0u;
    if constexpr (PAIR) crank = cluster_ctarank();
    const int htag = int(crank) * 100 + 200;

    // tile = (pixel tile, channel tile), channel tile fastest: CTAs that run side by side share their input rows in L2
    constexpr int N = 128 * NC;
    const int total_tiles = PAIR ? p.nimg * (p.H >> 1) * (p.W >> 7) * (NC / 2) : p.nimg * p.H * (p.W >> 8) * NC;
    const int tile0 = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int tstep = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);

    if (warp == 0) {
        // ===================================================================== weight producer
        if (lane == 0) {
            int ws = 0;
            uint32_t wphase = 0;
            bool first_pass = true;
            for (int tile = tile0; tile < total_tiles; tile += tstep) {
                const SwTile t = sw_decode<NC, PAIR>(p, tile, int(crank));
                for (int ch = 0; ch < p.kchunks; ++ch)
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) {
                            mbar_wait(&wempty[ws], wphase ^ 1u, hw, htag + 1);
                            uint8_t* dst = wring + size_t(ws) * kSwWBytes;
                            const int k0 = (p.tap_of[r][c] * p.kchunks + ch) * kBlockK;
                            if constexpr (PAIR) {
                                // each CTA stages its 128 channels; bytes are counted on the leader's barrier
                                if ((p.dbg_mma_only == 1 || p.dbg_mma_only == 2) && !first_pass) {
                                    if (crank == 0) mbar_arrive(&wfull[ws]);
                                } else {
                                    if (crank == 0) mbar_arrive_expect_tx(&wfull[ws], 2u * uint32_t(kSwWBytes));
                                    tma_load_3d_2sm(dst, &mapW, &wfull[ws], k0, t.chan0, 0);
                                }
                            } else {
                                if ((p.dbg_mma_only == 1 || p.dbg_mma_only == 2) && !first_pass) {
                                    mbar_arrive(&wfull[ws]);
                                } else {
                                    mbar_arrive_expect_tx(&wfull[ws], uint32_t(kSwWBytes));
                                    tma_load_3d(dst, &mapW, &wfull[ws], k0, t.chan0, 0);
                                }
                            }
                            if (++ws == kWStages) { ws = 0; wphase ^= 1u; first_pass = false; }
                        }
            }
        }
    } else if (warp == 3) {
        // ===================================================================== input-row producer
        if (lane == 0) {
            int rs = 0;
            uint32_t rphase = 0;
            bool first_pass = true;
            for (int tile = tile0; tile < total_tiles; tile += tstep) {
                const SwTile t = sw_decode<NC, PAIR>(p, tile, int(crank));
                if (p.l2_prefetch && tile + tstep < total_tiles) {
                    // The input rows are streamed from HBM (the first tile that touches a row misses L2); the ring only
                    // looks two slots ahead, so the next tile's rows are requested a whole tile early.
                    const SwTile n = sw_decode<NC, PAIR>(p, tile + tstep, int(crank));
                    for (int ch = 0; ch < p.kchunks; ++ch)
                        for (int r = 0; r < 3; ++r)
                            tma_prefetch_l2_4d(&mapRow, ch * kBlockK, n.bx0 - 1, n.brow + r - 1, n.img);
                }
                for (int ch = 0; ch < p.kchunks; ++ch)
                    for (int r = 0; r < 3; ++r) {
                        mbar_wait(&rempty[rs], rphase ^ 1u, hw, htag + 2);
                        uint8_t* dst = rows + size_t(rs) * kRowBytes;
                        // rows or pixels outside the image arrive as zeros (= the convolution padding)
                        if constexpr (XF) {
                            // the raw row lands on this CTA's own barrier; the transform warps pass it on to the leader's rfull
                            mbar_arrive_expect_tx(&rland[rs], uint32_t(G::kRowPx) * 128u);
                            tma_load_4d(dst, &mapRow, &rland[rs], ch * kBlockK, t.bx0 - 1, t.brow + r - 1, t.img);
                        } else if constexpr (PAIR) {
                            if ((p.dbg_mma_only == 1 || p.dbg_mma_only == 3) && !first_pass) {
                                if (crank == 0) mbar_arrive(&rfull[rs]);
                            } else {
                                if (crank == 0) mbar_arrive_expect_tx(&rfull[rs], 2u * uint32_t(G::kRowPx) * 128u);
                                tma_load_4d_2sm(dst, &mapRow, &rfull[rs], ch * kBlockK, t.bx0 - 1, t.brow + r - 1, t.img);
                            }
                        } else {
                            if ((p.dbg_mma_only == 1 || p.dbg_mma_only == 3) && !first_pass) {
                                mbar_arrive(&rfull[rs]);
                            } else {
                                mbar_arrive_expect_tx(&rfull[rs], uint32_t(G::kRowPx) * 128u);
                                // pixels x0-1 .. x0+254, then x0+255 .. x0+256 (a TMA box is at most 256 wide)
                                tma_load_4d(dst, &mapRow, &rfull[rs], ch * kBlockK, t.bx0 - 1, t.brow + r - 1, t.img);
                                tma_load_4d(dst + 256 * 128, &mapTail, &rfull[rs], ch * kBlockK, t.bx0 + 255, t.brow + r - 1, t.img);
                            }
                        }
                        if (++rs == kRowSlots) { rs = 0; rphase ^= 1u; first_pass = false; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (convergent warp)
        if (crank == 0) {   // pair mode: only the leader CTA issues
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kUmmaM : kUmmaM, 256);
            int ws = 0, rs = 0, acc = 0;
            uint32_t wphase = 0, rphase = 0, acc_phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstep) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u, hw, htag + 3);
                tc_fence_after();
                const uint32_t d = tmem_base + uint32_t(acc * 256);
                for (int ch = 0; ch < p.kchunks; ++ch)
                    for (int r = 0; r < 3; ++r) {
                        if constexpr (XF) mbar_wait_cluster(&rfull[rs], rphase, hw, htag + 4);   // rows rewritten by both CTAs' threads
                        else mbar_wait(&rfull[rs], rphase, hw, htag + 4);
                        tc_fence_after();
                        const uint32_t raddr = smem_u32(rows + size_t(rs) * kRowBytes);
                        for (int c = 0; c < 3; ++c) {
                            mbar_wait(&wfull[ws], wphase, hw, htag + 5);
                            tc_fence_after();
                            const uint64_t a_desc = umma_desc_sw128(smem_u32(wring + size_t(ws) * kSwWBytes));
                            // operand rows = consecutive pixels of the slot starting at pixel c (= dw + 1)
                            const uint64_t b_desc = umma_desc_sw128(raddr + uint32_t(c) * 128u);
                            const uint32_t first = (ch | r | c) != 0 ? 1u : 0u;
                            if (elect_one()) {
                                if constexpr (PAIR) {
                                    umma_bf16_2sm(d, a_desc, b_desc, idesc, first);
                                    umma_bf16_2sm(d, a_desc + 2, b_desc + 2, idesc, 1u);
                                    umma_bf16_2sm(d, a_desc + 4, b_desc + 4, idesc, 1u);
                                    umma_bf16_2sm(d, a_desc + 6, b_desc + 6, idesc, 1u);
                                    umma_commit_2sm(&wempty[ws], 3);
                                    if (c == 2) umma_commit_2sm(&rempty[rs], 3);
                                } else {
                                    umma_bf16(d, a_desc, b_desc, idesc, first);
                                    umma_bf16(d, a_desc + 2, b_desc + 2, idesc, 1u);
                                    umma_bf16(d, a_desc + 4, b_desc + 4, idesc, 1u);
                                    umma_bf16(d, a_desc + 6, b_desc + 6, idesc, 1u);
                                    umma_commit(&wempty[ws]);
                                    if (c == 2) umma_commit(&rempty[rs]);
                                }
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wphase ^= 1u; }
                        }
                        if (++rs == kRowSlots) { rs = 0; rphase ^= 1u; }
                    }
                if (elect_one()) {
                    if constexpr (PAIR) umma_commit_2sm(&tfull[acc], 3); else umma_commit(&tfull[acc]);
                }
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4 + SwCfg<GNB>::kEpiWarps) {
        // ===================================================================== input normalisation (XF: 8 warps)
        if constexpr (XF) {
            const int tt = int(threadIdx.x) - (128 + 32 * SwCfg<GNB>::kEpiWarps);   // 0 .. kXfThreads-1
            const int j = tt & 7;            // this thread's 16-byte chunk of a pixel row = channels 8j .. 8j+7 of the chunk
            const int r0 = tt >> 3;          // its pixels: r0, r0 + 32, ... (five at most of the slot's 130)
            constexpr int PXI = kXfThreads / 8, NIT = (G::kRowPx + PXI - 1) / PXI;
            const int C_in = p.kchunks * kBlockK;
            int rs = 0, nslot = 0;
            uint32_t rphase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstep) {
                const SwTile t = sw_decode<NC, PAIR>(p, tile, int(crank));
                for (int ch = 0; ch < p.kchunks; ++ch) {
                    // silu(u) = u * sigmoid(u) = h + h * tanh(h) with h = u / 2: scale and shift arrive pre-halved
                    float sc[8], sh[8];
                    const float2* ssp = p.in_ss + (size_t)t.img * C_in + ch * kBlockK + 8 * j;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 v = __ldg(ssp + e);
                        sc[e] = 0.5f * v.x; sh[e] = 0.5f * v.y;
                    }
                    for (int r = 0; r < 3; ++r) {
                        mbar_wait(&rland[rs], rphase, hw, htag + 7);
                        const int yin = t.brow + r - 1;
                        if (yin >= 0 && yin < p.H) {   // rows outside the image stay the zeros TMA wrote (= the padding)
                            uint8_t* slot = rows + size_t(rs) * kRowBytes;
                            uint4 u[NIT];
                            bool ok[NIT];
#pragma unroll
                            for (int i = 0; i < NIT; ++i) {       // all loads first: five independent 16-byte reads in flight
                                const int px = r0 + i * PXI, xin = t.bx0 - 1 + px;
                                ok[i] = px < G::kRowPx && xin >= 0 && xin < p.W;   // halo pixels outside the image: padding
                                if (ok[i]) u[i] = *reinterpret_cast<const uint4*>(slot + size_t(px) * 128 + ((j ^ (px & 7)) << 4));
                            }
#pragma unroll
                            for (int i = 0; i < NIT; ++i) {
                                if (!ok[i]) continue;
                                const int px = r0 + i * PXI;
                                const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
                                uint32_t o[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float h0 = fmaf(bf16_lo(w[e]), sc[2 * e], sh[2 * e]);
                                    const float h1 = fmaf(bf16_hi(w[e]), sc[2 * e + 1], sh[2 * e + 1]);
                                    o[e] = pack_bf16x2(fmaf(h0, tanh_approx(h0), h0), fmaf(h1, tanh_approx(h1), h1));
                                }
                                *reinterpret_cast<uint4*>(slot + size_t(px) * 128 + ((j ^ (px & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                            }
                        }
                        fence_proxy_async();       // generic-proxy writes -> visible to the tensor core's reads
                        // ONE arrival per CTA and slot, issued by a different warp each time: a release.cluster arrive on
                        // the peer's barrier stalls its thread for a cross-SM round trip (measured with four remote
                        // arrivals per slot: 660 instead of 1450 TFLOP/s), so the cost is rotated over the four warps
                        named_bar_sync(5, kXfThreads);
                        if (lane == 0 && (tt >> 5) == (nslot & 7)) {
                            if (crank != 0) mbar_arrive_remote(&rfull[rs], 0);   // release.cluster
                            else mbar_arrive(&rfull[rs]);
                        }
                        ++nslot;
                        if (++rs == kRowSlots) { rs = 0; rphase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (8 or 16 warps)
        // warp e: TMEM lane quadrant e % 4 (output channels 32q .. 32q+31 of this CTA's 128), pixel part e / 4
        constexpr int PX = SwCfg<GNB>::kPx, STEP = SwCfg<GNB>::kStep;
        constexpr int COLS_ROW = PAIR ? 128 : 256;   // accumulator columns per output row
        const int e = warp - 4;
        const int q = e & 3, part = e >> 2;
        constexpr int CPG = N / 32;      // channels per GroupNorm group = consecutive lanes per group (4, 8 or 16)
        const int segs_row = p.W / PX;   // partial-sum segments per image row
        const int et = int(threadIdx.x) - 128;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstep) {
            const SwTile t = sw_decode<NC, PAIR>(p, tile, int(crank));
            const int img = t.img;
            const int chn = t.chan0 + q * 32 + lane;
            const int wrow = t.orow + (part * PX) / COLS_ROW, wx = t.ox0 + (part * PX) % COLS_ROW;   // this warp's first pixel
            const size_t off = (((size_t)img * p.H + wrow) * p.W + wx) * N + chn;
            if constexpr (RES || GNB) {
                // Pull the NEXT tile's residual / GroupNorm-input slice (256 pixels x 256 bytes) into L2 a whole tile
                // ahead: epilogue thread t < 256 takes the pixel of accumulator column t (two 128-byte lines).
                const int nxt = tile + tstep;
                if (nxt < total_tiles && et < 256) {
                    const SwTile n = sw_decode<NC, PAIR>(p, nxt, int(crank));
                    const size_t noff = (((size_t)n.img * p.H + n.orow + et / COLS_ROW) * p.W + n.ox0 + et % COLS_ROW) * N + n.chan0;
                    if constexpr (RES) {
                        prefetch_l2(p.resid + noff);
                        prefetch_l2(p.resid + noff + 64);
                    }
                    if constexpr (GNB) {
                        prefetch_l2(p.gn_x + noff);
                        prefetch_l2(p.gn_x + noff + 64);
                    }
                }
            }
            // per-thread constants (thread = channel): requested before waiting for the accumulator
            const float bias_c = p.bias != nullptr ? __ldg(p.bias + chn) : 0.f;
            float sc = 0.f, sh = 0.f, gm = 0.f, mean = 0.f, rstd = 0.f;
            if constexpr (GNB) {
                const float2 s2 = __ldg(&p.gn_ss[(size_t)img * N + chn]);
                const float2 m2 = __ldg(&p.gn_mr[(size_t)img * 32 + chn / CPG]);
                sc = s2.x; sh = s2.y; mean = m2.x; rstd = m2.y;
                gm = __ldg(p.gn_gamma + chn);
                if (p.gn_silu) { sc *= 0.5f; sh *= 0.5f; gm *= 0.5f; }   // the tanh form of silu' below works on u/2
            }
            __nv_bfloat16* dp = p.D + off;
            const unsigned short* rp = reinterpret_cast<const unsigned short*>(p.resid) + off;
            const unsigned short* xp = reinterpret_cast<const unsigned short*>(p.gn_x) + off;
            // Residual / GroupNorm-input values are gathered per thread (thread = channel, one 2-byte load per
            // pixel; a warp's load of one pixel is 64 contiguous bytes).  They are fetched one step ahead of their
            // use, the first step before the accumulator is even complete.
            unsigned short ra[STEP], rb[STEP], xa[STEP], xb[STEP];
            auto fetch = [&](int k, unsigned short (&r)[STEP], unsigned short (&x)[STEP]) {
                if constexpr (RES) {
#pragma unroll
                    for (int j = 0; j < STEP; ++j) r[j] = __ldg(rp + (size_t)(k * STEP + j) * N);
                }
                if constexpr (GNB) {
#pragma unroll
                    for (int j = 0; j < STEP; ++j) x[j] = __ldg(xp + (size_t)(k * STEP + j) * N);
                }
            };
            float s = 0.f, ss = 0.f;
            auto process = [&](int k, const uint32_t (&v)[STEP], const unsigned short (&r)[STEP], const unsigned short (&x)[STEP]) {
#pragma unroll
                for (int j = 0; j < STEP; ++j) {
                    float f = __uint_as_float(v[j]) + bias_c;
                    if constexpr (RES) f += __uint_as_float(uint32_t(r[j]) << 16);
                    const __nv_bfloat16 b = __float2bfloat16_rn(f);
                    if (p.dbg_no_epi != 2) dp[(size_t)(k * STEP + j) * N] = b;
                    const float fr = __bfloat162float(b);   // reductions see the values as stored
                    if constexpr (GNB) {
                        // GroupNorm backward: dxh = dy * act'(x*sc+sh) * gamma; sums of dxh and dxh*x.
                        // silu'(u) = sg + u*sg*(1-sg) with sg = (1 + t)/2, t = tanh(u/2):  silu'(u) = (1 + t + h*(1 - t*t))/2,
                        // h = u/2 -- one MUFU and four FMAs instead of exp + rcp + six (sc, sh, gm arrive pre-halved).
                        const float xv = __uint_as_float(uint32_t(x[j]) << 16);
                        float dxh = fr * gm;
                        if (p.gn_silu) {
                            const float h = fmaf(xv, sc, sh);
                            const float t = tanh_approx(h);
                            const float z = fmaf(h, fmaf(-t, t, 1.f), t);
                            dxh = fmaf(dxh, z, dxh);
                        }
                        s += dxh;
                        ss = fmaf(dxh, xv, ss);
                    } else {
                        s += fr;
                        ss = fmaf(fr, fr, ss);
                    }
                }
            };
            if (p.dbg_no_epi != 1) fetch(0, ra, xa);
            mbar_wait(&tfull[acc], acc_phase, hw, htag + 6);
            tc_fence_after();
            if (p.dbg_no_epi != 1) {
                const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256 + part * PX);
#pragma unroll 1
                for (int k2 = 0; k2 < 4; ++k2) {
                    uint32_t v[STEP];
                    tmem_ld_n(t_addr + uint32_t(2 * k2 * STEP), v);
                    fetch(2 * k2 + 1, rb, xb);
                    tmem_ld_wait();
                    process(2 * k2, v, ra, xa);
                    tmem_ld_n(t_addr + uint32_t((2 * k2 + 1) * STEP), v);
                    if (k2 < 3) fetch(2 * k2 + 2, ra, xa);
                    tmem_ld_wait();
                    process(2 * k2 + 1, v, rb, xb);
                }
                if (p.gn_mode != 0) {
                    // sum(dxh*xh) = rstd * (sum(dxh*x) - mean * sum(dxh)); linear, so applied per thread
                    if constexpr (GNB) ss = rstd * fmaf(-mean, s, ss);
                    // group = CPG consecutive channels = CPG consecutive lanes; fixed order -> reproducible
#pragma unroll
                    for (int m = 1; m < CPG; m <<= 1) {
                        s += __shfl_xor_sync(0xffffffffu, s, m);
                        ss += __shfl_xor_sync(0xffffffffu, ss, m);
                    }
                    if ((lane & (CPG - 1)) == 0) {
                        const size_t seg = (size_t)wrow * segs_row + wx / PX;
                        float2* out = reinterpret_cast<float2*>(p.gn_partial) +
                                      (((size_t)img * p.H * segs_row + seg) * 32 + chn / CPG);
                        *out = make_float2(s, ss);
                    }
                }
            }
            tc_fence_before();
            if constexpr (PAIR) {
                if (crank != 0) mbar_arrive_remote(&tempty[acc], 0);   // the leader's MMA warp owns both accumulators
                else mbar_arrive(&tempty[acc]);
            } else {
                mbar_arrive(&tempty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // nobody leaves while the peer may still signal its barriers / read its TMEM
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || p == nullptr)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, const char* what, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -2; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed: %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", what, (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
                  box[1], box[2]);
        return -3;
    }
    return 0;
}

// (attn_fused.cu) bf16 / SWIZZLE_128B tensor map through the same driver entry point
int encode_map_bf16_sw128(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                          const cuuint32_t* box, const char* what) {
    return encode_map(m, base, rank, dims, strides_b, box, what);
}

// "dynamic shared memory limit raised" flags, one per (device, kernel family): the attribute is per device
static bool& attr_flag(int family) {
    static bool flags[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    return flags[family][dev & 63];
}

// One mapped host word per process: a kernel whose mbarrier wait times out stores the id of that wait here
// before trapping, so the failure can be attributed after the context is gone.
static int* g_hang_host = nullptr;
static int* g_hang_dev = nullptr;
static volatile int* hang_word_device() {
    if (!g_hang_host) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&g_hang_host), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            g_hang_host = nullptr;
            return nullptr;
        }
        *g_hang_host = 0;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_hang_dev), g_hang_host, 0) != cudaSuccess) {
            cudaGetLastError();
            g_hang_dev = nullptr;
        }
    }
    return g_hang_dev;
}
int gemm_last_hang() { return g_hang_host ? *g_hang_host : 0; }

// Is this op a dense 3x3 stride-1 convolution to 128 / 256 / 512 channels that the operand-swapped kernel tiles?
// Returns 0 (no), 1 (rows split into 256-pixel tiles: one CTA per tile) or 2 (rows of 128-pixel segments, an even
// number of rows, >= 256 channels: CTA pairs).  Shape only: the fused-reduction fields are checked at launch.
static int swapped_kind(const GemmOp& op) {
    static const bool off = getenv("TML_NO_SWAP") && getenv("TML_NO_SWAP")[0] == '1';            // tuning switch
    static const bool no_pair = getenv("TML_NO_SWAP_PAIR") && getenv("TML_NO_SWAP_PAIR")[0] == '1';   // tuning switch
    static const int max_n = getenv("TML_SWAP_MAX_N") ? atoi(getenv("TML_SWAP_MAX_N")) : 512;   // tuning switch
    if (off || op.stride != 1 || op.ntaps != 9 || (op.N != 128 && op.N != 256 && op.N != 512) || op.N > max_n ||
        op.OW % 128 != 0 || op.OW != op.A_W || op.OH != op.A_H || op.B_sBatch != 0 || op.dbg_shift != 0 ||
        op.alpha != 1.0f || op.out_fp32 || op.D_sN != 1 || op.A_C % kBlockK != 0 || (op.n_store != 0 && op.n_store != op.N))
        return 0;
    const int64_t sW = op.N, sH = (int64_t)op.OW * op.N, sB = (int64_t)op.OH * op.OW * op.N;
    if (op.D_sW != sW || op.D_sH != sH || op.D_sB != sB) return 0;
    if (op.resid && (op.R_sW != sW || op.R_sH != sH || op.R_sB != sB)) return 0;
    // measured in situ: pairs 1520-1610 TFLOP/s vs 1390-1500 for single CTAs on the 256-channel layers (a pair stages
    // each input row once for 256 channels), so pairs are used wherever they apply
    static const bool prefer_pair = !(getenv("TML_SWAP_PREFER_PAIR") && getenv("TML_SWAP_PREFER_PAIR")[0] == '0');   // tuning switch
    const bool pair_ok = !no_pair && op.OH % 2 == 0 && op.N >= 256 && (long)op.A_B * (op.OH / 2) * (op.OW / 128) * (op.N / 256) >= 2;
    if (op.OW % 256 == 0 && !(prefer_pair && pair_ok)) return 1;
    return pair_ok ? 2 : 0;
}
bool gemm_swapped_shape(const GemmOp& op) { return swapped_kind(op) != 0; }
bool gemm_fuses_input_gn(const GemmOp& op) {
    // OFF by default.  Measured on B200 (tools/gpu_r2_q.sh .. gpu_r2_s.sh): with the transform on the operand path the pair
    // convolutions run at 1200-1250 TFLOP/s isolated instead of 1630-1690 (eight transform warps: ~2100 warp-instructions
    // per 130 x 64 slot = a third of the SM's issue slots, one MUFU per element) and the whole step at 372-374 instead of
    // 386-388 image-PGD-iters/s: the separate apply pass it removes (3.5 ms per 64 images, at the HBM roofline) is cheaper.
    static const bool on = getenv("TML_FUSE_INGN") && getenv("TML_FUSE_INGN")[0] == '1';   // experiment switch
    return on && g_impl.load() == 0 && swapped_kind(op) == 2;
}
static bool swap_eligible(const GemmOp& op, int tap_of[3][3]) {
    if (!gemm_swapped_shape(op) || op.gn_mode < 0 || op.gn_mode > 2) return false;
    if (op.gn_mode != 0 && !op.gn_partial) return false;
    if (op.gn_mode == 2 && (!op.gn_x || !op.gn_ss || !op.gn_mr || !op.gn_gamma)) return false;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) tap_of[r][c] = -1;
    for (int i = 0; i < 9; ++i) {
        if (op.dh[i] < -1 || op.dh[i] > 1 || op.dw[i] < -1 || op.dw[i] > 1) return false;
        tap_of[op.dh[i] + 1][op.dw[i] + 1] = i;
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) if (tap_of[r][c] < 0) return false;
    return true;
}

static int gemm_launch_swapped(const GemmOp& op, const int tap_of[3][3], int num_sms, cudaStream_t stream) {
    int rc;
    const bool pair = swapped_kind(op) == 2;
    const bool xf = op.in_gn_ss != nullptr;
    if (xf && (!pair || op.gn_mode == 2)) {
        set_error("%s: the fused input normalisation needs the CTA-pair form of a forward convolution", op.name);
        return -1;
    }
    CUtensorMap mapW, mapRow, mapTail;
    {
        cuuint64_t dims[3] = {(cuuint64_t)op.ntaps * op.A_C, (cuuint64_t)op.N, 1};
        cuuint64_t str[2] = {(cuuint64_t)op.B_sN * 2, (cuuint64_t)op.B_sN * 2 * (cuuint64_t)op.N};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, 128, 1};
        if ((rc = encode_map(&mapW, op.Bm, 3, dims, str, box, op.name))) return rc;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)op.A_C, (cuuint64_t)op.A_W, (cuuint64_t)op.A_H, (cuuint64_t)op.A_B};
        cuuint64_t str[3] = {(cuuint64_t)op.A_sW * 2, (cuuint64_t)op.A_sH * 2, (cuuint64_t)op.A_sB * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, pair ? 130u : 256u, 1, 1};
        if ((rc = encode_map(&mapRow, op.A, 4, dims, str, box, op.name))) return rc;
        box[1] = 2;
        if ((rc = encode_map(&mapTail, op.A, 4, dims, str, box, op.name))) return rc;
    }
    SwParams p;
    memset(&p, 0, sizeof(p));
    p.H = op.OH; p.W = op.OW; p.nimg = op.A_B; p.kchunks = op.A_C / kBlockK;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) p.tap_of[r][c] = tap_of[r][c];
    p.bias = op.bias;
    p.resid = reinterpret_cast<const __nv_bfloat16*>(op.resid);
    p.D = reinterpret_cast<__nv_bfloat16*>(op.D);
    p.gn_mode = op.gn_mode;
    p.gn_partial = op.gn_partial;
    p.gn_x = reinterpret_cast<const __nv_bfloat16*>(op.gn_x);
    p.gn_ss = op.gn_ss; p.gn_mr = op.gn_mr; p.gn_gamma = op.gn_gamma; p.gn_silu = op.gn_silu;
    p.in_ss = op.in_gn_ss;
    p.hang_where = hang_word_device();
    { static const int pf = getenv("TML_SW_PREFETCH") ? atoi(getenv("TML_SW_PREFETCH")) : 0; p.l2_prefetch = pf; }   // tuning switch
    { static const int mo = getenv("TML_DBG_MMA_ONLY") ? atoi(getenv("TML_DBG_MMA_ONLY")) : 0; p.dbg_mma_only = mo;   // 1 none, 2 no weights, 3 no rows
      static const int ne = getenv("TML_DBG_NO_EPI") ? atoi(getenv("TML_DBG_NO_EPI")) : 0; p.dbg_no_epi = (ne == 1 || ne == 2) ? ne : 0; }
    const int nc = op.N / 128;
    const int total_tiles = pair ? op.A_B * (op.OH / 2) * (op.OW / 128) * (nc / 2) : op.A_B * op.OH * (op.OW / 256) * nc;
    int grid = pair ? (2 * total_tiles < num_sms ? 2 * total_tiles : (num_sms & ~1)) : (total_tiles < num_sms ? total_tiles : num_sms);
    const bool timed = g_timing && g_timed.size() < g_timing_cap;
    TimedLaunch tl;
    if (timed) {
        tl.a = take_event(); tl.b = take_event();
        tl.flops = 2.0 * (double)op.A_B * op.OH * op.OW * (double)op.N * (double)op.ntaps * op.A_C;
        snprintf(tl.key, sizeof(tl.key), "%s|%ld|%d|%d|%d", op.name, (long)op.A_B * op.OH * op.OW, op.N,
                 op.ntaps * op.A_C, (pair ? 3000 : 2000) + (xf ? 100 : 0) + op.gn_mode * 10);
        cudaEventRecord(tl.a, stream);
    } else if (g_timing) {
        ++g_timing_dropped;
    }
    {
        typedef void (*SwKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const SwParams);
#define SW_ROW(NC_, PAIR_)                                                                                   \
    {{conv3x3_swapped_kernel<NC_, false, false, PAIR_>, conv3x3_swapped_kernel<NC_, false, true, PAIR_>},    \
     {conv3x3_swapped_kernel<NC_, true, false, PAIR_>, conv3x3_swapped_kernel<NC_, true, true, PAIR_>}}
        static const SwKernel table[3][2][2] = {SW_ROW(1, false), SW_ROW(2, false), SW_ROW(4, false)};
        static const SwKernel ptable[2][2][2] = {SW_ROW(2, true), SW_ROW(4, true)};
#undef SW_ROW
        // [NC 2 / 4][residual]: forward convolutions on CTA pairs with the GroupNorm + SiLU of their input fused in
        static const SwKernel xtable[2][2] = {
            {conv3x3_swapped_kernel<2, false, false, true, true>, conv3x3_swapped_kernel<2, true, false, true, true>},
            {conv3x3_swapped_kernel<4, false, false, true, true>, conv3x3_swapped_kernel<4, true, false, true, true>}};
        bool& attr_set = attr_flag(0);   // cudaFuncSetAttribute is per device
        if (!attr_set) {
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    cudaError_t e = cudaSuccess;
                    for (int a = 0; a < 3 && e == cudaSuccess; ++a)
                        e = cudaFuncSetAttribute(table[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, SwGeo<false>::kSmem);
                    for (int a = 0; a < 2 && e == cudaSuccess; ++a)
                        e = cudaFuncSetAttribute(ptable[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, SwGeo<true>::kSmem);
                    for (int a = 0; a < 2 && e == cudaSuccess && c == 0; ++a)
                        e = cudaFuncSetAttribute(xtable[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, SwGeo<true>::kSmem);
                    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -4; }
                }
            attr_set = true;
        }
        const int threads = (op.gn_mode == 2 ? SwCfg<true>::kThreadsSw : SwCfg<false>::kThreadsSw) + (xf ? kXfThreads : 0);
        const int ri = op.resid ? 1 : 0, gi = op.gn_mode == 2 ? 1 : 0;
        if (pair) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = SwGeo<true>::kSmem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t le = cudaLaunchKernelEx(&cfg, xf ? xtable[nc == 2 ? 0 : 1][ri] : ptable[nc == 2 ? 0 : 1][ri][gi], mapW, mapRow, mapTail, p);
            if (le != cudaSuccess) { set_error("%s: cluster launch failed: %s", op.name, cudaGetErrorString(le)); return -5; }
        } else {
            table[nc == 1 ? 0 : nc == 2 ? 1 : 2][ri][gi]<<<grid, threads, SwGeo<false>::kSmem, stream>>>(mapW, mapRow, mapTail, p);
        }
    }
    if (timed) { cudaEventRecord(tl.b, stream); g_timed.push_back(tl); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s: launch failed: %s", op.name, cudaGetErrorString(e)); return -5; }
    g_tc_launches.fetch_add(1);
    return 0;
}

int gemm_launch_tc(const GemmOp& op, int num_sms, cudaStream_t stream) {
    {
        int tap_of[3][3];
        if (swap_eligible(op, tap_of)) return gemm_launch_swapped(op, tap_of, num_sms, stream);
    }
    if (op.in_gn_ss != nullptr) {
        set_error("%s: in_gn_ss is only implemented by the CTA-pair operand-swapped kernel (see gemm_fuses_input_gn)", op.name);
        return -1;
    }
    GemmTiling t;
    int rc = gemm_plan(op, &t);
    if (rc) return rc;
    if (op.epi_mode != 0 && ((op.epi_mode == 3) != (op.resid != nullptr) || (op.epi_mode != 3 && !op.row_part) ||
                             (op.epi_mode >= 2 && !op.row_a) || (op.epi_mode == 3 && !op.row_b) ||
                             (op.epi_mode != 1 && !op.D))) {
        set_error("%s: row-wise epilogue %d is missing one of its buffers", op.name, op.epi_mode);
        return -1;
    }

    CUtensorMap mapA, mapB;
    if (op.a_trans) {
        cuuint64_t dims[3] = {(cuuint64_t)op.OH * op.OW, (cuuint64_t)op.A_C, (cuuint64_t)op.A_B};
        cuuint64_t str[2] = {(cuuint64_t)op.A_sK * 2, (cuuint64_t)op.A_sB * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)kBlockK, 1};
        if ((rc = encode_map(&mapA, op.A, 3, dims, str, box, op.name))) return rc;
    } else if (op.stride == 1) {
        cuuint64_t dims[4] = {(cuuint64_t)op.A_C, (cuuint64_t)op.A_W, (cuuint64_t)op.A_H, (cuuint64_t)op.A_B};
        cuuint64_t str[3] = {(cuuint64_t)op.A_sW * 2, (cuuint64_t)op.A_sH * 2, (cuuint64_t)op.A_sB * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(t.TW + (op.dbg_shift ? 8 : 0)), (cuuint32_t)t.TH, 1};
        if (t.halo) { box[1] = 130; box[2] = (cuuint32_t)(t.mt + 2); }
        if (t.h66) { box[1] = 66; box[2] = 6; }
        if (op.dbg_shift && (t.TH != 1 || t.TW > 64 || t.mt != 1)) { set_error("dbg_shift needs TH=1, TW<=64, mt=1"); return -1; }
        if ((rc = encode_map(&mapA, op.A, 4, dims, str, box, op.name))) return rc;
    } else {
        cuuint64_t dims[5] = {(cuuint64_t)op.A_C, 2, (cuuint64_t)op.A_W / 2, (cuuint64_t)op.A_H, (cuuint64_t)op.A_B};
        cuuint64_t str[4] = {(cuuint64_t)op.A_sW * 2, (cuuint64_t)op.A_sW * 4, (cuuint64_t)op.A_sH * 2,
                             (cuuint64_t)op.A_sB * 2};
        cuuint32_t box[5] = {(cuuint32_t)kBlockK, 1, (cuuint32_t)t.TW, 1, 1};
        if ((rc = encode_map(&mapA, op.A, 5, dims, str, box, op.name))) return rc;
    }
    {
        const int ktot = op.ntaps * op.A_C;
        const int nb = op.B_sBatch ? op.A_B : 1;
        cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)op.N, (cuuint64_t)nb};
        cuuint64_t sb = op.B_sBatch ? (cuuint64_t)op.B_sBatch * 2 : (cuuint64_t)op.B_sN * 2 * (cuuint64_t)op.N;
        cuuint64_t str[2] = {(cuuint64_t)op.B_sN * 2, sb};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)(t.pair ? t.BN / 2 : t.BN), 1};
        if ((rc = encode_map(&mapB, op.Bm, 3, dims, str, box, op.name))) return rc;
    }

    CUtensorMap mapD = mapB;   // placeholder unless the staged epilogue is on (epi_mode 1 stores no tile)
    if (t.out_bytes && op.epi_mode != 1) {
        const cuuint64_t es = op.out_fp32 ? 4 : 2;
        cuuint64_t dims[4] = {(cuuint64_t)op.N, (cuuint64_t)op.OW, (cuuint64_t)op.OH, (cuuint64_t)op.A_B};
        cuuint64_t str[3] = {(cuuint64_t)op.D_sW * es, (cuuint64_t)op.D_sH * es, (cuuint64_t)op.D_sB * es};
        const cuuint32_t bw = t.TW < 32 ? (cuuint32_t)t.TW : 32u;      // one epilogue warp's 32 rows of the tile
        cuuint32_t box[4] = {32, bw, 32u / bw, 1};
        if ((rc = encode_map(&mapD, op.D, 4, dims, str, box, op.name,
                             op.out_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                             op.out_fp32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)))
            return rc;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.out_bytes = t.out_bytes;
    p.mode = op.stride == 2 ? 1 : 0;
    p.a_trans = op.a_trans;
    {   // weight-streaming layers: shared weights of >= 8 MB for at most 8192 output rows (the UNet's 16 x 16 and 8 x 8 levels)
        // (measured, tools/gpu_r2_pf.sh: no effect at 8 / 24 / 48 k-blocks -- these layers sit on the ~8 TB/s the L2 can hand to
        // the SMs, 33 B/clk/SM with every weight tile re-read by 16 row tiles, not on HBM latency: off by default)
        static const int pf = getenv("TML_B_PREFETCH") ? atoi(getenv("TML_B_PREFETCH")) : 0;   // experiment switch
        const double wbytes = 2.0 * op.N * op.ntaps * op.A_C;
        const long rows = (long)op.A_B * op.OH * op.OW;
        p.b_prefetch = (op.B_sBatch == 0 && !t.halo && wbytes >= 8e6 && rows <= 8192) ? pf : 0;
    }
    p.TW = t.TW; p.TH = t.TH; p.rows_valid = t.rows_valid;
    p.tiles_w = t.tiles_w; p.tiles_h = t.tiles_h; p.nimg = op.A_B;
    p.n_tiles = t.n_tiles; p.BN = t.BN; p.mt = t.mt;
    p.kchunks = t.kchunks; p.ntaps = op.ntaps;
    for (int i = 0; i < op.ntaps; ++i) { p.dh[i] = op.dh[i]; p.dw[i] = op.dw[i]; }
    p.b_batched = op.B_sBatch != 0;
    p.stages = t.stages;
    p.stage_bytes = t.stage_bytes;
    p.alpha = op.alpha;
    p.bias = op.bias;
    p.resid = reinterpret_cast<const __nv_bfloat16*>(op.resid);
    p.R_sB = op.R_sB; p.R_sH = op.R_sH; p.R_sW = op.R_sW;
    p.D = op.D; p.out_fp32 = op.out_fp32;
    p.D_sB = op.D_sB; p.D_sH = op.D_sH; p.D_sW = op.D_sW; p.D_sN = op.D_sN;
    p.n_store = op.n_store > 0 ? op.n_store : op.N;
    p.beta = op.out_fp32 ? op.beta : 0.f;
    p.gn_mode = op.gn_mode;
    p.gn_cpg = op.gn_mode ? op.N / 32 : 4;
    p.gn_partial = op.gn_partial;
    p.gn_x = reinterpret_cast<const __nv_bfloat16*>(op.gn_x);
    p.gn_ss = op.gn_ss; p.gn_mr = op.gn_mr; p.gn_gamma = op.gn_gamma; p.gn_silu = op.gn_silu;
    p.dbg_shift = op.dbg_shift; p.dbg_bo = op.dbg_bo;
    p.halo = t.halo; p.halo_bytes = t.halo_bytes; p.h66 = t.h66;
    p.pair = t.pair;
    p.epi_mode = op.epi_mode; p.row_a = op.row_a; p.row_b = op.row_b; p.row_part = op.row_part; p.exp_scale = op.exp_scale;
    p.row_scale = op.row_scale;
    p.rows_img = op.OH * op.OW; p.ow_full = op.OW;
    p.hang_where = hang_word_device();
    { static const bool mo = getenv("TML_DBG_MMA_ONLY") && getenv("TML_DBG_MMA_ONLY")[0] == '1'; p.dbg_mma_only = mo ? 1 : 0;
      static const int ne = getenv("TML_DBG_NO_EPI") ? atoi(getenv("TML_DBG_NO_EPI")) : 0; p.dbg_no_epi = ne; }

    bool& attr_set = attr_flag(1);   // cudaFuncSetAttribute is per device
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<true, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, EPI_LEAN_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, EPI_LEAN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, EPI_LEAN_ATTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<true, EPI_LEAN_ATTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -4; }
        attr_set = true;
    }
    const int total_tiles = (op.A_B * t.tiles_h * t.tiles_w / t.mt) * t.n_tiles;
    int grid = total_tiles < num_sms ? total_tiles : num_sms;
    if (t.pair) grid &= ~1;
    const bool timed = g_timing && g_timed.size() < g_timing_cap;
    TimedLaunch tl;
    if (timed) {
        tl.a = take_event(); tl.b = take_event();
        tl.flops = 2.0 * (double)op.A_B * op.OH * op.OW * (double)(op.n_store > 0 ? op.n_store : op.N) * (double)op.ntaps * op.A_C;
        snprintf(tl.key, sizeof(tl.key), "%s|%ld|%d|%d|%d", op.name, (long)op.A_B * op.OH * op.OW, op.N,
                 op.ntaps * op.A_C, (t.out_bytes ? 4000 : 0) + t.pair * 1000 + t.halo * 100 + op.gn_mode * 10 + t.mt);
        cudaEventRecord(tl.a, stream);
    } else if (g_timing) {
        ++g_timing_dropped;
    }
    if (t.pair) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = t.smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = op.epi_mode != 0
            ? cudaLaunchKernelEx(&cfg, conv_gemm_tcgen05_kernel<true, EPI_LEAN_ATTN>, mapA, mapB, mapD, p)
            : cudaLaunchKernelEx(&cfg, conv_gemm_tcgen05_kernel<true, EPI_GENERIC>, mapA, mapB, mapD, p);
        if (le != cudaSuccess) { set_error("%s: cluster launch failed: %s", op.name, cudaGetErrorString(le)); return -5; }
    } else {
        if (op.epi_mode != 0)
            conv_gemm_tcgen05_kernel<false, EPI_LEAN_ATTN><<<grid, kThreads, t.smem_bytes, stream>>>(mapA, mapB, mapD, p);
        else if (t.out_bytes && op.out_fp32)
            conv_gemm_tcgen05_kernel<false, EPI_LEAN_F32><<<grid, kThreads, t.smem_bytes, stream>>>(mapA, mapB, mapD, p);
        else if (t.out_bytes)
            conv_gemm_tcgen05_kernel<false, EPI_LEAN_BF16><<<grid, kThreads, t.smem_bytes, stream>>>(mapA, mapB, mapD, p);
        else
            conv_gemm_tcgen05_kernel<false, EPI_GENERIC><<<grid, kThreads, t.smem_bytes, stream>>>(mapA, mapB, mapD, p);
    }
    if (timed) { cudaEventRecord(tl.b, stream); g_timed.push_back(tl); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s: launch failed: %s", op.name, cudaGetErrorString(e)); return -5; }
    g_tc_launches.fetch_add(1);
    return 0;
}

thread_local bool g_dry_validate = false;   // dry runs also plan every GEMM (shape validation without a GPU)
int gemm_launch(const GemmOp& op, int num_sms, cudaStream_t stream) {
    if (g_dry_run) {
        if (g_dry_validate) { GemmTiling t; return gemm_plan(op, &t); }
        return 0;
    }
    if (g_impl.load() == 1) return gemm_launch_simt(op, stream);
    return gemm_launch_tc(op, num_sms, stream);
}

}  // namespace tml
