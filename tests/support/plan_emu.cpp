// plan_emu.cpp -- TEST-ONLY host mirror of k_stream's control flow.
//
// Links the product's host planner (imageprocessor_b200/csrc/plan.cpp) and walks
// its tables exactly as the CUDA kernel does (same tile/band ownership, same fp32
// fmaf order, same quantiser and ambiguity test), so the planner logic and the
// "certified fp32" error bound can be checked on a CPU-only box.  Never shipped,
// never linked into libipgpu.so; used by tests/test_plan_host.py only.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../imageprocessor_b200/csrc/plan.h"

using namespace ipg;

// optional capture (tests of the error bound): per target, dw * dh * 4 floats receiving fmaf(v, 256, 128) -- the
// 16.8 fixed-point value before floor() -- of every channel
static float *g_capture[2] = {nullptr, nullptr};
extern "C" void planemu_capture(float *t0, float *t1) { g_capture[0] = t0; g_capture[1] = t1; }
extern "C" int planemu_fix_d(int taps_x, int taps_y, int parts) { return certified_fix_d(taps_x, taps_y, parts); }
// 1: a geometry with the integer-moment vertical form (StreamGeom::vint_ok) is walked in that form, as the lean kernels do
static int g_vint = 0;
extern "C" void planemu_set_vint(int on) { g_vint = on; }

static inline int quant16(float v, int D, bool &amb)
{
    const uint32_t T = (uint32_t)std::min((int)std::floor(std::fmaf(v, 256.0f, 128.0f)), 0xffffff);
    amb |= ((T & 0xffffu) - (uint32_t)D) >= (65536u - 2u * (uint32_t)D);
    return (int)(T >> 16);
}

extern "C" int planemu_axis(int dn, int sn, int32_t *off, int32_t *first, double *w, double *inv, int *max_taps)
{
    auto p = get_axis_plan(dn, sn);
    if (off) memcpy(off, p->off.data(), p->off.size() * 4);
    if (first) memcpy(first, p->first.data(), p->first.size() * 4);
    if (w) memcpy(w, p->w.data(), p->w.size() * 8);
    if (inv) memcpy(inv, p->inv.data(), p->inv.size() * 8);
    if (max_taps) *max_taps = p->max_taps;
    return (int)p->w.size();
}

// specs: 6 ints per target {rect_x, rect_y, rect_w, rect_h, dw, dh}.
// flags[t]: dw*dh bytes, 1 where the kernel would queue an fp64 fix-up.
// info: {n_tiles, n_bands, tile_w, n_items, rows_read}
// SAMPLE = uint8_t: an *image.RGBA source (k_stream: samples are bytes, the vertical weights carry the 0x101);
// SAMPLE = uint16_t: the 16-bit premultiplied samples x/image's pass 1 reads from an NRGBA / YCbCr / Gray source
// (k_stream_planar: the lanes convert per tap, the vertical weights are unscaled).  Four samples per pixel.
template <typename SAMPLE>
static int run_impl(const SAMPLE *src, int stride, int W, int H, int n_targets, const int *specs,
                    const int *two_stage, uint8_t **dsts, uint8_t **flags, int bands_hint, int *info)
{
    constexpr bool WIDE = sizeof(SAMPLE) == 2;
    constexpr int OPAQUE = WIDE ? 0xffff : 0xff;
    StreamTargetSpec sp[2];
    for (int t = 0; t < n_targets; t++)
        sp[t] = StreamTargetSpec{specs[6 * t], specs[6 * t + 1], specs[6 * t + 2], specs[6 * t + 3], specs[6 * t + 4], specs[6 * t + 5]};
    auto g = get_stream_geom(W, H, sp, n_targets, false, bands_hint, WIDE ? 1.0 : 257.0);
    if (!g) return -1;
    long rows_read = 0;
    bool opaque_src = true; // then every column's alpha chain must equal the record's precomputed one
    for (int y = 0; y < H && opaque_src; y++)
        for (int x = 0; x < W; x++)
            if (src[(size_t)y * stride + (size_t)x * 4 + 3] != OPAQUE) { opaque_src = false; break; }
    for (const StreamItem &it : g->items) {
        const int tile = it.tile, band = it.band;
        const int cx0 = tile * g->tile_w;
        const int ys0 = g->band_y[band], yend = g->band_yend[band];
        // two accumulator sets per target, exactly as the kernel's VAcc; weights and emits come
        // from the GroupRec tables the kernel consumes (parity already resolved by the planner)
        std::vector<float> acc[2][2];
        const bool vint = g_vint && g->vint_ok && !WIDE && opaque_src; // (a non-opaque source is redone in the fp32 form)
        std::vector<uint32_t> iacc((size_t)STREAM_COLS * 4, 0u); // integer-moment form: M0 | M1 << 12 per byte column
        std::vector<float> irow((size_t)STREAM_COLS * 4, 0.f), carry((size_t)STREAM_COLS * 4, 0.f); // the kernel's cy / nx
        int piece_rows = 0;  // rows in the integer-moment piece being walked
        bool act[2] = {false, false};
        for (int t = 0; t < n_targets; t++) {
            acc[t][0].assign((size_t)STREAM_COLS * 4, 0.f);
            acc[t][1].assign((size_t)STREAM_COLS * 4, 0.f);
            act[t] = g->t[t].tile_ox[tile + 1] > g->t[t].tile_ox[tile] && g->t[t].band_tend[band] > ys0;
        }
        // the horizontal pass over a completed, vertically filtered row of target t
        auto emit_row = [&](int t, int oy, const std::vector<float> &row, int fix_d) -> int {
            const StreamTargetGeom &tg = g->t[t];
            for (int ox = tg.tile_ox[tile]; ox < tg.tile_ox[tile + 1]; ox++) {
                const int k0 = tg.ax->off[ox], n = tg.ax->off[ox + 1] - k0;
                const int e0 = tg.ax->first[ox] + sp[t].rect_x - cx0;
                if (e0 < 0 || e0 + n > g->slab_cols) return -2;
                // horizontal sum in the kernel's order: P threads per output, each over
                // its interleaved taps in order, then an xor-butterfly of the partial sums
                const int P = std::max(tg.tile_parts[tile] & 255, 1);
                if (tg.local) { // must lie inside the owning warp's 128 loaded columns
                    int w = 0;
                    while (w < 3 && ox >= tg.warp_ox[(size_t)tile * 4 + w + 1]) w++;
                    if (ox < tg.warp_ox[(size_t)tile * 4 + w] || e0 < w * g->warp_stride ||
                        e0 + n > w * g->warp_stride + STREAM_WARP_COLS) return -5;
                }
                float part[32][4];
                for (int pp = 0; pp < P; pp++)
                    for (int ch = 0; ch < 4; ch++) {
                        float a = 0.f;
                        for (int kk = pp; kk < n; kk += P) a = std::fmaf(row[(size_t)(e0 + kk) * 4 + ch], tg.xw[k0 + kk], a);
                        part[pp][ch] = a;
                    }
                for (int off = 1; off < P; off <<= 1) {
                    float nx[32][4];
                    for (int pp = 0; pp < P; pp++)
                        for (int ch = 0; ch < 4; ch++) nx[pp][ch] = part[pp][ch] + part[pp ^ off][ch];
                    memcpy(part, nx, sizeof nx);
                }
                float s[4] = {part[0][0], part[0][1], part[0][2], part[0][3]};
                for (int ch = 0; ch < 3; ch++) s[ch] = std::fmin(s[ch], s[3]);
                bool amb = false;
                uint8_t *d = dsts[t] + ((size_t)oy * sp[t].dw + ox) * 4;
                for (int ch = 0; ch < 4; ch++) d[ch] = (uint8_t)quant16(s[ch], fix_d, amb);
                if (g_capture[t])
                    for (int ch = 0; ch < 4; ch++) g_capture[t][((size_t)oy * sp[t].dw + ox) * 4 + ch] = std::fmaf(s[ch], 256.0f, 128.0f);
                if (flags && flags[t]) flags[t][(size_t)oy * sp[t].dw + ox] += amb ? 1 : 16; // 16: written once
            }
            return 0;
        };
        const int ngroups = (yend - ys0 + STREAM_GROUP - 1) / STREAM_GROUP;
        for (int gi = 0; gi < ngroups; gi++) {
            for (int k = 0; k < STREAM_GROUP; k++) {
                const int ys = ys0 + gi * STREAM_GROUP + k;
                if (ys < yend) rows_read++;
                for (int t = 0; t < n_targets; t++) {
                    if (!act[t]) continue;
                    const StreamTargetGeom &tg = g->t[t];
                    const GroupRec &G = g->grec[((size_t)g->band_grec_off[band] + (size_t)gi) * (size_t)g->rec_slots + (size_t)t];
                    if (vint) { // k_stream's v_rows_int: IDP.2A per channel and row, fp32 only when a segment ends
                        const GroupRecI &GI = *reinterpret_cast<const GroupRecI *>(&g->grec[((size_t)g->band_grec_off[band] + (size_t)gi) * (size_t)g->rec_slots + 1]);
                        if (ys >= yend && (GI.m[k] != 0 || GI.emit[k] != -1)) return -3;
                        // the record as the kernel takes it: rows up to end_k before the flush, the rest after it -- that is
                        // the row-by-row walk below iff end_k / end_e name the group's only end; and a piece never outgrows
                        // its moment word (<= 16 rows, r <= 15, multiplier 1 + (r << 12))
                        if (k == 0) {
                            int ends = 0, ke = STREAM_GROUP - 1, ee = -1;
                            for (int kk = 0; kk < STREAM_GROUP; kk++)
                                if (GI.emit[kk] != -1) { ends++; ke = kk; ee = GI.emit[kk]; }
                            if (ends > 1 || GI.end_k != ke || GI.end_e != ee) return -6;
                        }
                        if (GI.m[k] != 0) {
                            if ((GI.m[k] & 0xfffu) != 1u || (GI.m[k] >> 12) > 15u || ++piece_rows > 16) return -7;
                        } else if (GI.emit[k] != -1) return -7; // an end sits on a row that feeds the piece
                        if (GI.emit[k] != -1) piece_rows = 0;
                        for (int e = 0; e < STREAM_COLS; e++) {
                            const int c = cx0 + e;
                            uint8_t px[4] = {0, 0, 0, 255};
                            if (c < W && ys < yend) for (int q = 0; q < 4; q++) px[q] = (uint8_t)src[(size_t)ys * stride + (size_t)c * 4 + q];
                            for (int q = 0; q < 3; q++) iacc[(size_t)e * 4 + q] += (uint32_t)px[q] * GI.m[k];
                        }
                        if (GI.emit[k] == -1) continue;
                        for (int e = 0; e < STREAM_COLS; e++)
                            for (int q = 0; q < 3; q++) { // a piece ends: its moments become fp32
                                const uint32_t a = iacc[(size_t)e * 4 + q];
                                const float f0 = (float)(a & 0xfffu), f1 = (float)(a >> 12);
                                float &cy = irow[(size_t)e * 4 + q], &nx = carry[(size_t)e * 4 + q];
                                cy = std::fmaf(f1, GI.bR, std::fmaf(f0, GI.aR, cy));
                                nx = std::fmaf(f1, GI.bL, std::fmaf(f0, GI.aL, nx));
                                iacc[(size_t)e * 4 + q] = 0u;
                            }
                        if (GI.emit[k] == -3) continue;
                        if (GI.emit[k] >= 0) { // ... and its segment: the row is complete
                            std::vector<float> row = irow;
                            for (int e = 0; e < STREAM_COLS; e++) row[(size_t)e * 4 + 3] = 65535.0f;
                            const int rc = emit_row(t, GI.emit[k], row, tg.fix_d_vint);
                            if (rc) return rc;
                        }
                        irow = carry;
                        std::fill(carry.begin(), carry.end(), 0.f);
                        continue;
                    }
                    const GroupRow r = G.row[k];
                    if (ys >= yend && (r.w0 != 0.f || r.w1 != 0.f || G.emit[k] >= 0)) return -3; // padding rows must be inert
                    for (int e = 0; e < STREAM_COLS; e++) {
                        const int c = cx0 + e;
                        SAMPLE px[4] = {0, 0, 0, (SAMPLE)OPAQUE};
                        if (c < W && ys < yend) memcpy(px, src + (size_t)ys * stride + (size_t)c * 4, 4 * sizeof(SAMPLE));
                        if (two_stage[t]) {
                            for (int q = 0; q < 3; q++) px[q] = std::min(px[q], px[3]);
                            // cropAndResize's 1:1 first pass keeps uint8(c16 >> 8); scaleX_RGBA re-expands it
                            if (WIDE) for (int q = 0; q < 4; q++) px[q] = (SAMPLE)((px[q] >> 8) * 0x101);
                        }
                        for (int q = 0; q < 4; q++) {
                            float &a = acc[t][0][(size_t)e * 4 + q], &b = acc[t][1][(size_t)e * 4 + q];
                            a = std::fmaf((float)px[q], r.w0, a);
                            b = std::fmaf((float)px[q], r.w1, b);
                        }
                    }
                    // the opaque fast path takes alpha from the record instead of the accumulators
                    if (opaque_src) {
                        for (int e = 0; e < STREAM_COLS; e++) {
                            if (acc[t][0][(size_t)e * 4 + 3] != r.sa0 || acc[t][1][(size_t)e * 4 + 3] != r.sa1) return -4;
                        }
                    }
                    if (G.emit[k] >= 0) {
                        const int set = G.emit[k] & 1;
                        std::vector<float> row = acc[t][set];
                        std::fill(acc[t][set].begin(), acc[t][set].end(), 0.f);
                        const int rc = emit_row(t, G.emit[k] >> 1, row, tg.fix_d);
                        if (rc) return rc;
                    }
                }
            }
        }
        if (piece_rows != 0) return -8; // a band's last row closes its last piece
    }
    if (info) {
        info[0] = g->n_tiles; info[1] = g->n_bands; info[2] = g->tile_w;
        info[3] = (int)g->items.size(); info[4] = (int)std::min<long>(rows_read, 2147483647L);
        info[5] = (g_vint && g->vint_ok) ? g->t[0].fix_d_vint : g->t[0].fix_d; info[6] = n_targets > 1 ? g->t[1].fix_d : 0;
        info[7] = g->vint_ok ? 1 : 0;
    }
    return 0;
}

extern "C" int planemu_run(const uint8_t *src, int stride, int W, int H, int n_targets, const int *specs,
                           const int *two_stage, uint8_t **dsts, uint8_t **flags, int bands_hint, int *info)
{
    return run_impl<uint8_t>(src, stride, W, H, n_targets, specs, two_stage, dsts, flags, bands_hint, info);
}

// the same over 16-bit samples (stride in samples, 4 per pixel)
extern "C" int planemu_run16(const uint16_t *src, int stride, int W, int H, int n_targets, const int *specs,
                             const int *two_stage, uint8_t **dsts, uint8_t **flags, int bands_hint, int *info)
{
    return run_impl<uint16_t>(src, stride, W, H, n_targets, specs, two_stage, dsts, flags, bands_hint, info);
}

// ---- k_direct (small-support fp32 kernel, one thread per output pixel), mirrored: for every contributing row the
// horizontal fmaf chain, then its fmaf into the vertical sum; same quantiser, same window.  spec as above; SAMPLE as above.
template <typename SAMPLE>
static int direct_impl(const SAMPLE *src, int stride, int W, int H, const int *spec, int two_stage, uint8_t *dst, uint8_t *flags,
                       float *capture, int *info)
{
    constexpr bool WIDE = sizeof(SAMPLE) == 2;
    const StreamTargetSpec sp{spec[0], spec[1], spec[2], spec[3], spec[4], spec[5]};
    if (sp.rect_x < 0 || sp.rect_y < 0 || sp.rect_x + sp.rect_w > W || sp.rect_y + sp.rect_h > H) return -2;
    auto g = get_direct_geom(sp, WIDE ? 1.0 : 257.0, true); // (the kernel's arithmetic is certified for mild downscales as well)
    if (!g) return -1;
    const AxisPlan &ax = *g->ax, &ay = *g->ay;
    for (int oy = 0; oy < sp.dh; oy++)
        for (int ox = 0; ox < sp.dw; ox++) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const int kx0 = ax.off[ox], nx = ax.off[ox + 1] - kx0, ky0 = ay.off[oy], ny = ay.off[oy + 1] - ky0;
            const int x0 = ax.first[ox] + sp.rect_x, y0 = ay.first[oy] + sp.rect_y;
            for (int j = 0; j < ny; j++) {
                float s[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < nx; k++) {
                    SAMPLE px[4];
                    memcpy(px, src + (size_t)(y0 + j) * stride + (size_t)(x0 + k) * 4, 4 * sizeof(SAMPLE));
                    if (two_stage) {
                        for (int q = 0; q < 3; q++) px[q] = std::min(px[q], px[3]);
                        if (WIDE) for (int q = 0; q < 4; q++) px[q] = (SAMPLE)((px[q] >> 8) * 0x101);
                    }
                    for (int q = 0; q < 4; q++) s[q] = std::fmaf((float)px[q], g->xw[kx0 + k], s[q]);
                }
                for (int q = 0; q < 4; q++) acc[q] = std::fmaf(s[q], g->yw[ky0 + j], acc[q]);
            }
            for (int q = 0; q < 3; q++) acc[q] = std::fmin(acc[q], acc[3]);
            bool amb = false;
            uint8_t *d = dst + ((size_t)oy * sp.dw + ox) * 4;
            for (int q = 0; q < 4; q++) d[q] = (uint8_t)quant16(acc[q], g->fix_d, amb);
            if (flags) flags[(size_t)oy * sp.dw + ox] = amb ? 1 : 16;
            if (capture)
                for (int q = 0; q < 4; q++) capture[((size_t)oy * sp.dw + ox) * 4 + q] = std::fmaf(acc[q], 256.0f, 128.0f);
        }
    if (info) { info[0] = g->fix_d; info[1] = ax.max_taps; info[2] = ay.max_taps; }
    return 0;
}

extern "C" int planemu_direct(const uint8_t *src, int stride, int W, int H, const int *spec, int two_stage, uint8_t *dst,
                              uint8_t *flags, float *capture, int *info)
{
    return direct_impl<uint8_t>(src, stride, W, H, spec, two_stage, dst, flags, capture, info);
}
extern "C" int planemu_direct16(const uint16_t *src, int stride, int W, int H, const int *spec, int two_stage, uint8_t *dst,
                                uint8_t *flags, float *capture, int *info)
{
    return direct_impl<uint16_t>(src, stride, W, H, spec, two_stage, dst, flags, capture, info);
}
