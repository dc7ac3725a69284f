// kernels.cu -- hand-written sm_100a kernels for ImageProcessor's raster hot path.
//
//   k_stream      fp32, vertical-first streaming separable resample; one pass over
//                 the source produces up to two resampled outputs (resize + thumb)
//                 and, optionally, the watermark copy with its glyph blend.
//                 Replaces resizeImage (operations/resize.go:121-125),
//                 cropAndResize (operations/thumbnail.go:114-132) and the raster
//                 part of addTextWatermark (operations/watermark.go:91-92,151).
//   k_exact_*     fp64, the reference's exact operation order (x/image v0.33.0
//                 draw.Kernel.Scale: scaleX_<type> then scaleY_RGBA_Src), unfused
//                 multiply/add.  Whole outputs (REFERENCE mode, upscales, odd
//                 layouts) or only the pixels k_stream flagged (EXACT mode).
//   k_watermark   draw.Draw(dst,...,Src) conversion for every source layout +
//                 stdlib drawGlyphOver in string order (uint32, wrapping).
//
// No tensor cores: nothing here is a dense contraction; the bound is HBM.
#include "kernels.h"

namespace ipg {

// ---------------------------------------------------------------------------------
// source adaptors (the inner expressions of x/image draw/impl.go scaleX_<type>)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ size_t chroma_index(int layout, int s1, int x, int y)
{
    switch (layout) {
    case L_YCBCR444: return (size_t)y * s1 + x;
    case L_YCBCR422: return (size_t)y * s1 + (x >> 1);
    case L_YCBCR420: return (size_t)(y >> 1) * s1 + (x >> 1);
    default:         return (size_t)(y >> 1) * s1 + x; // 4:4:0
    }
}

// 16-bit premultiplied sample as the reference's pass 1 sees source pixel (x,y).
// Returns true when the alpha lane is the literal constant 1.0 (Gray, YCbCr).
__device__ __forceinline__ bool sample16(const SrcView &s, int x, int y, uint32_t p[4])
{
    switch (s.layout) {
    case L_RGBA8: {
        uint32_t q = __ldg((const uint32_t *)(s.p0 + (size_t)y * s.s0) + x);
        p[0] = (q & 0xff) * 0x101u; p[1] = ((q >> 8) & 0xff) * 0x101u;
        p[2] = ((q >> 16) & 0xff) * 0x101u; p[3] = (q >> 24) * 0x101u;
        return false;
    }
    case L_NRGBA8: {
        uint32_t q = __ldg((const uint32_t *)(s.p0 + (size_t)y * s.s0) + x);
        uint32_t pa = (q >> 24) * 0x101u;
        p[0] = (q & 0xff) * pa / 0xffu; p[1] = ((q >> 8) & 0xff) * pa / 0xffu;
        p[2] = ((q >> 16) & 0xff) * pa / 0xffu; p[3] = pa;
        return false;
    }
    case L_GRAY8: {
        uint32_t v = (uint32_t)__ldg(s.p0 + (size_t)y * s.s0 + x) * 0x101u;
        p[0] = p[1] = p[2] = v; p[3] = 0xffffu;
        return true;
    }
    default: {
        size_t ci = chroma_index(s.layout, s.s1, x, y);
        int yy1 = (int)__ldg(s.p0 + (size_t)y * s.s0 + x) * 0x10101;
        int cb1 = (int)__ldg(s.p1 + ci) - 128;
        int cr1 = (int)__ldg(s.p2 + ci) - 128;
        p[0] = (uint32_t)clampi((yy1 + 91881 * cr1) >> 8, 0, 0xffff);
        p[1] = (uint32_t)clampi((yy1 - 22554 * cb1 - 46802 * cr1) >> 8, 0, 0xffff);
        p[2] = (uint32_t)clampi((yy1 + 116130 * cb1) >> 8, 0, 0xffff);
        p[3] = 0xffffu;
        return true;
    }
    }
}

// cropAndResize's first Scale is 1:1 (one tap of weight 1): it stores
// uint8(min(c16,a16) >> 8) into an *image.RGBA; the second Scale then reads that
// through scaleX_RGBA.  Fold both into one sample.
__device__ __forceinline__ void to_cropped_rgba16(uint32_t p[4])
{
    uint32_t a = p[3];
    p[0] = (min(p[0], a) >> 8) * 0x101u;
    p[1] = (min(p[1], a) >> 8) * 0x101u;
    p[2] = (min(p[2], a) >> 8) * 0x101u;
    p[3] = (a >> 8) * 0x101u;
}

// x/image draw/scale.go ftou
__device__ __forceinline__ uint32_t ftou(double f)
{
    double v = __dadd_rn(__dmul_rn(65535.0, f), 0.5);
    int i = __double2int_rz(v); // NaN -> 0, like Go's clamp of the INT_MIN it yields
    return (uint32_t)clampi(i, 0, 0xffff);
}

// One output pixel in the reference's exact order: for every contributing source
// row, the horizontal sum (sequential, unfused), times invTotalWeightFFFF; then the
// vertical sum, premultiplied clamp, times invTotalWeight, ftou, >> 8.
__device__ uchar4 exact_pixel(const ExactJob &J, int ox, int oy)
{
    const int kx0 = __ldg(J.ax.off + ox), nx = __ldg(J.ax.off + ox + 1) - kx0;
    const int ky0 = __ldg(J.ay.off + oy), ny = __ldg(J.ay.off + oy + 1) - ky0;
    const int x0 = __ldg(J.ax.first + ox) + J.rect_x;
    const int y0 = __ldg(J.ay.first + oy) + J.rect_y;
    const double ifx = __ldg(J.ax.inv_ffff + ox);
    double pr = 0, pg = 0, pb = 0, pa = 0;
    for (int j = 0; j < ny; j++) {
        double xr = 0, xg = 0, xb = 0, xa = 0;
        bool const_alpha = false;
        for (int k = 0; k < nx; k++) {
            uint32_t p[4];
            const_alpha = sample16(J.src, x0 + k, y0 + j, p);
            if (J.two_stage) { to_cropped_rgba16(p); const_alpha = false; }
            const double w = __ldg(J.ax.w + kx0 + k);
            xr = __dadd_rn(xr, __dmul_rn((double)p[0], w));
            xg = __dadd_rn(xg, __dmul_rn((double)p[1], w));
            xb = __dadd_rn(xb, __dmul_rn((double)p[2], w));
            xa = __dadd_rn(xa, __dmul_rn((double)p[3], w));
        }
        const double wy = __ldg(J.ay.w + ky0 + j);
        const double ta = const_alpha ? 1.0 : __dmul_rn(xa, ifx);
        pr = __dadd_rn(pr, __dmul_rn(__dmul_rn(xr, ifx), wy));
        pg = __dadd_rn(pg, __dmul_rn(__dmul_rn(xg, ifx), wy));
        pb = __dadd_rn(pb, __dmul_rn(__dmul_rn(xb, ifx), wy));
        pa = __dadd_rn(pa, __dmul_rn(ta, wy));
    }
    if (pr > pa) pr = pa;
    if (pg > pa) pg = pa;
    if (pb > pa) pb = pa;
    const double iy = __ldg(J.ay.inv + oy);
    uchar4 o;
    o.x = (unsigned char)(ftou(__dmul_rn(pr, iy)) >> 8);
    o.y = (unsigned char)(ftou(__dmul_rn(pg, iy)) >> 8);
    o.z = (unsigned char)(ftou(__dmul_rn(pb, iy)) >> 8);
    o.w = (unsigned char)(ftou(__dmul_rn(pa, iy)) >> 8);
    return o;
}

__global__ void __launch_bounds__(256)
k_exact_tiles(const ExactJob *__restrict__ jobs, const ExactItem *__restrict__ items)
{
    const ExactItem it = items[blockIdx.x];
    const ExactJob &J = jobs[it.job];
    const int ox = it.tile_x * 32 + (threadIdx.x & 31);
    const int oy = it.tile_y * 8 + (threadIdx.x >> 5);
    if (ox >= J.dw || oy >= J.dh) return;
    const uchar4 o = exact_pixel(J, ox, oy);
    *(uchar4 *)(J.dst + (size_t)oy * J.dst_stride + (size_t)ox * 4) = o;
}

__global__ void __launch_bounds__(128)
k_exact_fix(const ExactJob *__restrict__ jobs, int n_jobs, FixList fix)
{
    const uint32_t cnt = *fix.count;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (cnt <= fix.capacity) {
        for (uint32_t i = tid; i < cnt; i += nthr) {
            const FixEntry e = fix.entries[i];
            const ExactJob &J = jobs[e.job];
            const uchar4 o = exact_pixel(J, e.x, e.y);
            *(uchar4 *)(J.dst + (size_t)e.y * J.dst_stride + (size_t)e.x * 4) = o;
        }
    } else {
        // the list overflowed: entries were dropped, so redo every stream target whole
        for (int j = 0; j < n_jobs; j++) {
            const ExactJob &J = jobs[j];
            const uint32_t npx = (uint32_t)J.dw * (uint32_t)J.dh;
            for (uint32_t i = tid; i < npx; i += nthr) {
                const int x = (int)(i % (uint32_t)J.dw), y = (int)(i / (uint32_t)J.dw);
                const uchar4 o = exact_pixel(J, x, y);
                *(uchar4 *)(J.dst + (size_t)y * J.dst_stride + (size_t)x * 4) = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// stdlib image/draw drawGlyphOver (Go 1.24): uint32 arithmetic, wrapping on purpose
// (the default colour 255,255,255,127 is not a valid premultiplied colour).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t glyph_over_px(uint32_t d, int x, int y, const WatermarkD &wm)
{
    const uint32_t m = 0xffffu;
    for (int k = 0; k < wm.n_glyphs; k++) {
        const GlyphD &g = wm.glyphs[k];
        if (x < g.x0 || x >= g.x1 || y < g.y0 || y >= g.y1) continue;
        uint32_t ma = __ldg(g.mask + (size_t)(y - g.y0 + g.mp_y) * g.mask_stride + (x - g.x0 + g.mp_x));
        if (ma == 0) continue;
        ma |= ma << 8;
        const uint32_t a = (m - (wm.sa * ma / m)) * 0x101u;
        const uint32_t r = (((d & 0xff) * a + wm.sr * ma) / m >> 8) & 0xff;
        const uint32_t g8 = ((((d >> 8) & 0xff) * a + wm.sg * ma) / m >> 8) & 0xff;
        const uint32_t b = ((((d >> 16) & 0xff) * a + wm.sb * ma) / m >> 8) & 0xff;
        const uint32_t al = (((d >> 24) * a + wm.sa * ma) / m >> 8) & 0xff;
        d = r | (g8 << 8) | (b << 16) | (al << 24);
    }
    return d;
}

__device__ __forceinline__ uint32_t blend_if_inside(uint32_t d, int x, int y, const WatermarkD &wm)
{
    if (x >= wm.bx0 && x < wm.bx1 && y >= wm.by0 && y < wm.by1) return glyph_over_px(d, x, y, wm);
    return d;
}

// draw.Draw(dst *image.RGBA, r, src, sp, draw.Src) per source type: copy,
// drawNRGBASrc, imageutil.DrawYCbCr (8-bit color.YCbCrToRGB), drawGray.
__device__ __forceinline__ uint32_t draw_src_px(const SrcView &s, int x, int y)
{
    switch (s.layout) {
    case L_RGBA8: return __ldg((const uint32_t *)(s.p0 + (size_t)y * s.s0) + x);
    case L_NRGBA8: {
        uint32_t q = __ldg((const uint32_t *)(s.p0 + (size_t)y * s.s0) + x);
        uint32_t sa = (q >> 24) * 0x101u;
        uint32_t r = ((q & 0xff) * sa / 0xffu) >> 8;
        uint32_t g = (((q >> 8) & 0xff) * sa / 0xffu) >> 8;
        uint32_t b = (((q >> 16) & 0xff) * sa / 0xffu) >> 8;
        return r | (g << 8) | (b << 16) | ((sa >> 8) << 24);
    }
    case L_GRAY8: {
        uint32_t v = __ldg(s.p0 + (size_t)y * s.s0 + x);
        return v * 0x010101u | 0xff000000u;
    }
    default: {
        size_t ci = chroma_index(s.layout, s.s1, x, y);
        int yy1 = (int)__ldg(s.p0 + (size_t)y * s.s0 + x) * 0x10101;
        int cb1 = (int)__ldg(s.p1 + ci) - 128;
        int cr1 = (int)__ldg(s.p2 + ci) - 128;
        uint32_t r = (uint32_t)clampi((yy1 + 91881 * cr1) >> 16, 0, 0xff);
        uint32_t g = (uint32_t)clampi((yy1 - 22554 * cb1 - 46802 * cr1) >> 16, 0, 0xff);
        uint32_t b = (uint32_t)clampi((yy1 + 116130 * cb1) >> 16, 0, 0xff);
        return r | (g << 8) | (b << 16) | 0xff000000u;
    }
    }
}

__global__ void __launch_bounds__(256)
k_watermark(const WmJob *__restrict__ jobs, const WmItem *__restrict__ items)
{
    const WmItem it = items[blockIdx.x];
    const WmJob &J = jobs[it.job];
    const int W = J.src.w;
    const int y1 = min(it.row0 + WM_ROWS, J.src.h);
    for (int y = it.row0; y < y1; y++) {
        uint8_t *drow = J.wm.dst + (size_t)y * J.wm.dst_stride;
        for (int x = threadIdx.x * 4; x < W; x += 256 * 4) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                o[j] = (x + j < W) ? blend_if_inside(draw_src_px(J.src, x + j, y), x + j, y, J.wm) : 0u;
            if (x + 4 <= W && ((J.wm.dst_stride | (int)(size_t)J.wm.dst) & 15) == 0) {
                *(uint4 *)(drow + (size_t)x * 4) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                for (int j = 0; j < 4 && x + j < W; j++) *(uint32_t *)(drow + (size_t)(x + j) * 4) = o[j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// k_stream: vertical-first fp32 streaming resample
//
// CTA = 128 threads x 4 source pixels = a 512-column slab; it owns `tile_w` of those
// columns, the rest is the right halo the widest horizontal support needs.  The CTA
// walks the source rows of its band once, top to bottom.  Per source row each thread
// converts its 4 RGBA pixels to fp32 once and multiply-adds them into the (at most)
// two output rows the row contributes to, per target: a tent of half-width `scale`
// centred every `scale` rows covers each source row exactly twice.  When a source
// row completes an output row (RowRec.emit), the CTA parks that vertically-filtered
// row in shared memory (XOR-swizzled float4 slots: conflict-free stores, <=2-way
// gathers), and threads 0..n_owned-1 run the horizontal gather, quantise with the
// reference's ftou()>>8, flag bytes too close to a quantiser step for the fp64
// fix-up, and store uchar4 (coalesced).
// Source bytes are read from HBM exactly once for resize + thumbnail + watermark.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> 3) & 7); }

__device__ __forceinline__ void unpack_px(uint32_t q, float *v)
{
    v[0] = (float)(q & 0xff);
    v[1] = (float)((q >> 8) & 0xff);
    v[2] = (float)((q >> 16) & 0xff);
    v[3] = (float)(q >> 24);
}

__device__ __forceinline__ uint32_t clamp_to_alpha(uint32_t q)
{
    uint32_t a = q >> 24;
    uint32_t r = min(q & 0xff, a), g = min((q >> 8) & 0xff, a), b = min((q >> 16) & 0xff, a);
    return r | (g << 8) | (b << 16) | (a << 24);
}

__device__ __forceinline__ uint4 load_px4(const uint8_t *row, int c, int W, bool vec)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c + 4 <= W && vec) {
        v = __ldcs((const uint4 *)(row + (size_t)c * 4)); // streaming: read once
    } else if (c < W) {
        const uint32_t *p = (const uint32_t *)row + c;
        v.x = __ldg(p);
        if (c + 1 < W) v.y = __ldg(p + 1);
        if (c + 2 < W) v.z = __ldg(p + 2);
        if (c + 3 < W) v.w = __ldg(p + 3);
    }
    return v;
}

__device__ __forceinline__ int quant16(float v, int D, bool &amb)
{
    // t = v + 0.5 in 16.8 fixed point; out byte = floor(t) >> 8 = T >> 16.
    const int T = __float2int_rd(fmaf(v, 256.0f, 128.0f));
    const int out = min(T >> 16, 255);
    const int lo = min(max(T - D, 0) >> 16, 255);
    const int hi = min((T + D) >> 16, 255);
    amb |= (lo != hi);
    return max(out, 0);
}

__device__ __forceinline__ void xpass(const StreamTarget &t, int tile, int oy, int cx0,
                                      const float4 *__restrict__ buf, const FixList &fix)
{
    const int ox0 = __ldg(t.tile_ox + tile);
    const int n_own = __ldg(t.tile_ox + tile + 1) - ox0;
    for (int j = threadIdx.x; j < n_own; j += STREAM_THREADS) {
        const int ox = ox0 + j;
        const int k0 = __ldg(t.xoff + ox);
        const int n = __ldg(t.xoff + ox + 1) - k0;
        const int e0 = __ldg(t.xfirst + ox) + t.rect_x - cx0;
        float r = 0.f, g = 0.f, b = 0.f, a = 0.f;
        for (int k = 0; k < n; k++) {
            const float w = __ldg(t.xw + k0 + k);
            const float4 q = buf[swz(e0 + k)];
            r = fmaf(q.x, w, r); g = fmaf(q.y, w, g);
            b = fmaf(q.z, w, b); a = fmaf(q.w, w, a);
        }
        r = fminf(r, a); g = fminf(g, a); b = fminf(b, a);
        bool amb = false;
        uchar4 o;
        o.x = (unsigned char)quant16(r, t.fix_d, amb);
        o.y = (unsigned char)quant16(g, t.fix_d, amb);
        o.z = (unsigned char)quant16(b, t.fix_d, amb);
        o.w = (unsigned char)quant16(a, t.fix_d, amb);
        *(uchar4 *)(t.dst + (size_t)oy * t.dst_stride + (size_t)ox * 4) = o;
        if (amb && fix.capacity) {
            const uint32_t idx = atomicAdd(fix.count, 1u);
            if (idx < fix.capacity) fix.entries[idx] = FixEntry{t.exact_job, ox, oy};
        }
    }
}

template <int NT, bool WM, bool CHECK>
__global__ void __launch_bounds__(STREAM_THREADS, (NT == 2 ? 3 : 4))
k_stream(const StreamJob *__restrict__ jobs, const StreamItem *__restrict__ items, FixList fix)
{
    __shared__ float4 rowbuf[2][STREAM_COLS];

    const StreamItem it = items[blockIdx.x];
    const StreamJob &J = jobs[it.job];
    const int tile = it.tile, band = it.band;
    const int W = J.src.w;
    const int cx0 = tile * J.tile_w;
    const int c = cx0 + (int)threadIdx.x * STREAM_PX;
    const int ys0 = __ldg(J.band_y + band), ys1 = __ldg(J.band_y + band + 1);
    const int yend = __ldg(J.band_yend + band);
    const int stride = J.src.s0;
    const bool vec = (((size_t)J.src.p0 | (size_t)stride) & 15) == 0;

    float acc_a[NT > 0 ? NT : 1][16], acc_b[NT > 0 ? NT : 1][16];
    const RowRec *rec[NT > 0 ? NT : 1];
    int tend[NT > 0 ? NT : 1]; // rows >= tend[T] contribute nothing to target T in this (tile, band)
#pragma unroll
    for (int T = 0; T < NT; T++) {
#pragma unroll
        for (int i = 0; i < 16; i++) { acc_a[T][i] = 0.f; acc_b[T][i] = 0.f; }
        rec[T] = nullptr;
        tend[T] = ys0;
        if (T < J.n_targets && __ldg(J.t[T].tile_ox + tile + 1) > __ldg(J.t[T].tile_ox + tile)) {
            rec[T] = J.t[T].rows + __ldg(J.t[T].band_rec_off + band);
            tend[T] = __ldg(J.t[T].band_tend + band);
        }
    }

    const uint8_t *row = J.src.p0 + (size_t)ys0 * stride;
    uint4 cur = load_px4(row, c, W, vec);
    int emits = 0;
    const bool own_col = (int)threadIdx.x * STREAM_PX < J.tile_w && c < W;

    for (int ys = ys0; ys < yend; ys++) {
        row += stride;
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (ys + 1 < yend) nxt = load_px4(row, c, W, vec);

        if (WM && J.has_wm && ys < ys1 && own_col) {
            uint4 o = cur;
            const WatermarkD &wm = J.wm;
            if (ys >= wm.by0 && ys < wm.by1 && c + 4 > wm.bx0 && c < wm.bx1) {
                o.x = blend_if_inside(o.x, c, ys, wm);
                o.y = blend_if_inside(o.y, c + 1, ys, wm);
                o.z = blend_if_inside(o.z, c + 2, ys, wm);
                o.w = blend_if_inside(o.w, c + 3, ys, wm);
            }
            uint8_t *d = wm.dst + (size_t)ys * wm.dst_stride + (size_t)c * 4;
            if (c + 4 <= W && (((size_t)wm.dst | (size_t)wm.dst_stride) & 15) == 0) {
                __stcs((uint4 *)d, o);
            } else {
                ((uint32_t *)d)[0] = o.x;
                if (c + 1 < W) ((uint32_t *)d)[1] = o.y;
                if (c + 2 < W) ((uint32_t *)d)[2] = o.z;
                if (c + 3 < W) ((uint32_t *)d)[3] = o.w;
            }
        }

        if (NT > 0) {
            float v[16];
            unpack_px(cur.x, v); unpack_px(cur.y, v + 4);
            unpack_px(cur.z, v + 8); unpack_px(cur.w, v + 12);
            bool slow = false;
            if (CHECK && J.check_premul) {
                const uint32_t m = min(min(cur.x, cur.y), min(cur.z, cur.w));
                slow = __any_sync(0xffffffffu, m < 0xff000000u);
            }
#pragma unroll
            for (int T = 0; T < NT; T++) {
                if (ys >= tend[T]) continue; // CTA-uniform
                const int4 rr = __ldg((const int4 *)(rec[T] + (ys - ys0)));
                const float wa = __int_as_float(rr.x), wb = __int_as_float(rr.y);
                if (CHECK && slow && J.t[T].two_stage) {
                    float vc[16];
                    unpack_px(clamp_to_alpha(cur.x), vc); unpack_px(clamp_to_alpha(cur.y), vc + 4);
                    unpack_px(clamp_to_alpha(cur.z), vc + 8); unpack_px(clamp_to_alpha(cur.w), vc + 12);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        acc_a[T][i] = fmaf(vc[i], wa, acc_a[T][i]);
                        acc_b[T][i] = fmaf(vc[i], wb, acc_b[T][i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        acc_a[T][i] = fmaf(v[i], wa, acc_a[T][i]);
                        acc_b[T][i] = fmaf(v[i], wb, acc_b[T][i]);
                    }
                }
                if (rr.z >= 0) { // CTA-uniform: this source row completes output row rr.z
                    float4 *buf = rowbuf[emits & 1];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        buf[swz((int)threadIdx.x * 4 + j)] =
                            make_float4(acc_a[T][4 * j], acc_a[T][4 * j + 1], acc_a[T][4 * j + 2], acc_a[T][4 * j + 3]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) { acc_a[T][i] = acc_b[T][i]; acc_b[T][i] = 0.f; }
                    __syncthreads();
                    xpass(J.t[T], tile, rr.z, cx0, buf, fix);
                    emits++;
                }
            }
        }
        cur = nxt;
    }
}

int stream_smem_bytes() { return (int)(2 * STREAM_COLS * sizeof(float4)); }

template <int NT, bool WM, bool CHECK>
static cudaError_t launch_stream_t(const StreamJob *jobs, const StreamItem *items, int n, FixList fix, cudaStream_t st)
{
    k_stream<NT, WM, CHECK><<<n, STREAM_THREADS, 0, st>>>(jobs, items, fix);
    return cudaGetLastError();
}

cudaError_t launch_stream(const StreamJob *jobs, const StreamItem *items, int n_items, int max_targets,
                          bool any_wm, bool any_check, FixList fix, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
#define IPG_DISPATCH(NT)                                                                          \
    if (any_wm) {                                                                                 \
        return any_check ? launch_stream_t<NT, true, true>(jobs, items, n_items, fix, st)         \
                         : launch_stream_t<NT, true, false>(jobs, items, n_items, fix, st);       \
    } else {                                                                                      \
        return any_check ? launch_stream_t<NT, false, true>(jobs, items, n_items, fix, st)        \
                         : launch_stream_t<NT, false, false>(jobs, items, n_items, fix, st);      \
    }
    switch (max_targets) {
    case 0: if (!any_wm) return cudaSuccess; return launch_stream_t<0, true, false>(jobs, items, n_items, fix, st);
    case 1: IPG_DISPATCH(1)
    default: IPG_DISPATCH(2)
    }
#undef IPG_DISPATCH
}

cudaError_t launch_exact_tiles(const ExactJob *jobs, const ExactItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_exact_tiles<<<n_items, 256, 0, st>>>(jobs, items);
    return cudaGetLastError();
}

cudaError_t launch_exact_fix(const ExactJob *jobs, int n_jobs, FixList fix, cudaStream_t st)
{
    if (!fix.capacity || n_jobs <= 0) return cudaSuccess;
    k_exact_fix<<<148 * 4, 128, 0, st>>>(jobs, n_jobs, fix);
    return cudaGetLastError();
}

cudaError_t launch_watermark(const WmJob *jobs, const WmItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_watermark<<<n_items, 256, 0, st>>>(jobs, items);
    return cudaGetLastError();
}

} // namespace ipg
