/*
 * ip_oracle.c -- CPU oracle (plain C, float64, scalar) for the ImageProcessor
 * worker hot path.  TEST INFRASTRUCTURE ONLY -- see ip_oracle.h.
 * PARITY UNPINNED (no reference tests / goldens / Go toolchain) -- see header.
 *
 * Build with -ffp-contract=off: Go on amd64 never fuses a*b+c, so neither may
 * this file (oracle/Makefile passes the flag).
 */

#include "ip_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------ */
/* geometry                                                                  */
/* ------------------------------------------------------------------------ */

/* operations/resize.go:63-72 */
void ipo_keep_aspect_dims(int ow, int oh, int w, int h, int *nw, int *nh)
{
    double width_ratio = (double)w / (double)ow;
    double height_ratio = (double)h / (double)oh;
    double ratio = width_ratio < height_ratio ? width_ratio : height_ratio; /* math.Min */
    *nw = (int)((double)ow * ratio);
    *nh = (int)((double)oh * ratio);
}

/* operations/thumbnail.go:52-63 */
void ipo_thumb_fit_dims(int ow, int oh, int size, int *nw, int *nh)
{
    if (ow > oh) {
        *nh = size;
        *nw = (int)((double)ow * (double)size / (double)oh);
    } else {
        *nw = size;
        *nh = (int)((double)oh * (double)size / (double)ow);
    }
}

/* operations/thumbnail.go:115-127 */
void ipo_crop_square(int ow, int oh, int *cx, int *cy, int *cs)
{
    if (ow > oh) {
        *cs = oh;
        *cx = (ow - oh) / 2;
        *cy = 0;
    } else {
        *cs = ow;
        *cx = 0;
        *cy = (oh - ow) / 2;
    }
}

/* operations/watermark.go:116,118: fixed.Int26_6(fontSize*64*1.2).Ceil() */
int ipo_watermark_height_px(double font_size)
{
    double v = font_size * 64;
    v = v * 1.2;
    int32_t fx = (int32_t)v;       /* float -> Int26_6 truncates */
    return (int)((fx + 0x3f) >> 6); /* Int26_6.Ceil */
}

/* operations/watermark.go:121-148 (margin 20; unknown -> bottom-right) */
void ipo_watermark_anchor(const char *position, int W, int H, int width_px,
                          int height_px, int *x, int *y)
{
    const int margin = 20;
    if (!strcmp(position, "top-left")) {
        *x = margin; *y = margin + height_px;
    } else if (!strcmp(position, "top-right")) {
        *x = W - width_px - margin; *y = margin + height_px;
    } else if (!strcmp(position, "top-center")) {
        *x = (W - width_px) / 2; *y = margin + height_px;
    } else if (!strcmp(position, "bottom-left")) {
        *x = margin; *y = H - margin;
    } else if (!strcmp(position, "bottom-center")) {
        *x = (W - width_px) / 2; *y = H - margin;
    } else if (!strcmp(position, "center")) {
        *x = (W - width_px) / 2; *y = (H + height_px) / 2;
    } else { /* "bottom-right" and default */
        *x = W - width_px - margin; *y = H - margin;
    }
}

/* strconv.Atoi: [+-]?[0-9]+ , must fit in int64. */
static int go_atoi(const char *s, size_t n, long long *out)
{
    size_t k = 0;
    int neg = 0;
    if (n == 0) return -1;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; k = 1; }
    if (k == n) return -1;
    unsigned long long v = 0;
    for (; k < n; k++) {
        if (s[k] < '0' || s[k] > '9') return -1;
        unsigned d = (unsigned)(s[k] - '0');
        if (v > (0x7fffffffffffffffULL + (neg ? 1ULL : 0ULL) - d) / 10ULL) return -1;
        v = v * 10 + d;
    }
    *out = neg ? -(long long)v : (long long)v;
    return 0;
}

/* watermark.go:188-190 goes through float64 min/max; same result on ints. */
static int clampi(long long v, int lo, int hi)
{
    double d = (double)v;
    if (d > (double)hi) d = (double)hi;
    if (d < (double)lo) d = (double)lo;
    return (int)d;
}

/* uint8(255 * opacity) as Go/amd64 evaluates it for a float64 variable. */
static uint8_t opacity_byte(double opacity)
{
    double v = 255 * opacity;
    return (uint8_t)(long long)v;
}

/* operations/watermark.go:159-186 (+ the caller's fallback, :94-97) */
int ipo_parse_color(const char *s, double opacity, uint8_t rgba[4])
{
    char buf[256];
    size_t n = 0;
    for (const char *p = s; *p && n + 1 < sizeof buf; p++)
        if (*p != ' ') buf[n++] = *p; /* strings.ReplaceAll(" ", "") */
    buf[n] = 0;
    const char *part[8];
    size_t plen[8];
    int np = 0;
    const char *start = buf;
    for (size_t k = 0;; k++) {
        if (buf[k] == ',' || buf[k] == 0) {
            if (np < 8) { part[np] = start; plen[np] = (size_t)(buf + k - start); }
            np++;
            start = buf + k + 1;
            if (buf[k] == 0) break;
        }
    }
    uint8_t a_op = opacity_byte(opacity);
    long long r, g, b;
    if ((np != 3 && np != 4) || go_atoi(part[0], plen[0], &r) ||
        go_atoi(part[1], plen[1], &g) || go_atoi(part[2], plen[2], &b)) {
        rgba[0] = rgba[1] = rgba[2] = 0; /* caller: "using black" */
        rgba[3] = a_op;
        return -1;
    }
    rgba[0] = (uint8_t)clampi(r, 0, 255);
    rgba[1] = (uint8_t)clampi(g, 0, 255);
    rgba[2] = (uint8_t)clampi(b, 0, 255);
    long long av;
    if (np == 4 && go_atoi(part[3], plen[3], &av) == 0)
        rgba[3] = (uint8_t)clampi(av, 0, 255);
    else
        rgba[3] = a_op;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Spec R: x/image draw.Kernel{Support:1, At: 1-t}.Scale                     */
/* ------------------------------------------------------------------------ */

typedef struct { int32_t i, j; double inv_total, inv_total_ffff; } span_t;
typedef struct { int32_t coord; double weight; } tap_t;
typedef struct { span_t *spans; tap_t *taps; int32_t n_spans, n_taps; } axis_t;

static void axis_free(axis_t *a) { free(a->spans); free(a->taps); }

/* x/image draw/scale.go newDistrib, BiLinear (Support 1, At(t)=1-t). */
static int axis_build(axis_t *a, int32_t dw, int32_t sw)
{
    double scale = (double)sw / (double)dw;
    double half_width = 1.0, arg_scale = 1.0;
    if (scale > 1) {
        half_width *= scale;
        arg_scale = 1 / scale;
    }
    a->n_spans = dw;
    a->spans = (span_t *)malloc(sizeof(span_t) * (size_t)(dw > 0 ? dw : 1));
    if (!a->spans) return -1;
    int64_t n = 0;
    for (int32_t x = 0; x < dw; x++) {
        double center = ((double)x + 0.5) * scale - 0.5;
        int32_t i = (int32_t)floor(center - half_width);
        if (i < 0) i = 0;
        int32_t j = (int32_t)ceil(center + half_width);
        if (j > sw) {
            j = sw;
            if (j < i) j = i;
        }
        a->spans[x].i = i;
        a->spans[x].j = j;
        a->spans[x].inv_total = center; /* parked, as upstream does */
        n += j - i;
    }
    a->taps = (tap_t *)malloc(sizeof(tap_t) * (size_t)(n > 0 ? n : 1));
    if (!a->taps) { free(a->spans); return -1; }
    int32_t len = 0;
    for (int32_t k = 0; k < dw; k++) {
        span_t b = a->spans[k];
        double total = 0.0;
        int32_t l = len;
        for (int32_t coord = b.i; coord < b.j; coord++) {
            double t = fabs((b.inv_total - (double)coord) * arg_scale);
            if (t >= 1.0) continue;
            double w = 1 - t;
            if (w == 0) continue;
            total += w;
            a->taps[len].coord = coord;
            a->taps[len].weight = w;
            len++;
        }
        total = 1 / total;
        a->spans[k].i = l;
        a->spans[k].j = len;
        a->spans[k].inv_total = total;
        a->spans[k].inv_total_ffff = total / 0xffff;
    }
    a->n_taps = len;
    return 0;
}

int ipo_distrib(int dw, int sw, int32_t *starts, int32_t *coords,
                double *weights, double *inv_total)
{
    axis_t a;
    if (axis_build(&a, dw, sw)) return -1;
    for (int k = 0; k < dw; k++) {
        if (starts) starts[k] = a.spans[k].i;
        if (inv_total) inv_total[k] = a.spans[k].inv_total;
    }
    if (starts) starts[dw] = a.n_taps;
    for (int k = 0; k < a.n_taps; k++) {
        if (coords) coords[k] = a.taps[k].coord;
        if (weights) weights[k] = a.taps[k].weight;
    }
    int n = a.n_taps;
    axis_free(&a);
    return n;
}

/* x/image draw/scale.go ftou */
static inline uint32_t ftou(double f)
{
    double v = 0xffff * f;
    v = v + 0.5;
    /* Go int32(NaN/out of range) on amd64 is 0x80000000 -> clamps to 0. */
    int32_t i = (v >= -2147483648.0 && v < 2147483648.0) ? (int32_t)v : INT32_MIN;
    if (i > 0xffff) return 0xffff;
    if (i > 0) return (uint32_t)i;
    return 0;
}

/* image.{RGBA,NRGBA}.Opaque(): full scan of the image's own bounds; YCbCr and
 * Gray are always opaque.  x/image draw/scale.go: op Over -> Src if opaque. */
static inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | (uint32_t)p[1]; }

static int image_opaque(const ipo_image *m)
{
    if (m->layout == IPO_RGBA64 || m->layout == IPO_NRGBA64) { /* image.RGBA64.Opaque(): every alpha == 0xffff */
        for (int y = 0; y < m->height; y++) {
            const uint8_t *row = m->plane[0] + (size_t)y * (size_t)m->stride[0];
            for (int x = 0; x < m->width; x++)
                if (be16(row + 8 * x + 6) != 0xffff) return 0;
        }
        return 1;
    }
    if (m->layout == IPO_PALETTED_RGBA || m->layout == IPO_PALETTED_NRGBA) {
        /* image.Paletted.Opaque(): the palette entries actually used by the pixels must all have alpha 0xffff */
        int present[256] = {0};
        for (int y = 0; y < m->height; y++) {
            const uint8_t *row = m->plane[0] + (size_t)y * (size_t)m->stride[0];
            for (int x = 0; x < m->width; x++) present[row[x]] = 1;
        }
        for (int i = 0; i < 256; i++)
            if (present[i] && m->plane[1][4 * i + 3] != 0xff) return 0;
        return 1;
    }
    if (m->layout == IPO_RGBA8 || m->layout == IPO_NRGBA8) {
        for (int y = 0; y < m->height; y++) {
            const uint8_t *row = m->plane[0] + (size_t)y * (size_t)m->stride[0];
            for (int x = 0; x < m->width; x++)
                if (row[4 * x + 3] != 0xff) return 0;
        }
        return 1;
    }
    return 1;
}

/* inline color.YCbCr.RGBA() as in x/image draw/impl.go scaleX_YCbCr* */
static inline void ycbcr_to_rgb16(int yy, int cb, int cr, uint32_t *r, uint32_t *g, uint32_t *b)
{
    int yy1 = yy * 0x10101;
    int cb1 = cb - 128;
    int cr1 = cr - 128;
    int pr = (yy1 + 91881 * cr1) >> 8;
    int pg = (yy1 - 22554 * cb1 - 46802 * cr1) >> 8;
    int pb = (yy1 + 116130 * cb1) >> 8;
    if (pr < 0) pr = 0; else if (pr > 0xffff) pr = 0xffff;
    if (pg < 0) pg = 0; else if (pg > 0xffff) pg = 0xffff;
    if (pb < 0) pb = 0; else if (pb > 0xffff) pb = 0xffff;
    *r = (uint32_t)pr; *g = (uint32_t)pg; *b = (uint32_t)pb;
}

/* 8-bit color.YCbCrToRGB as used by image/internal/imageutil.DrawYCbCr */
static inline void ycbcr_to_rgb8(int yy, int cb, int cr, uint8_t *r, uint8_t *g, uint8_t *b)
{
    int yy1 = yy * 0x10101;
    int cb1 = cb - 128;
    int cr1 = cr - 128;
    int pr = (yy1 + 91881 * cr1) >> 16;
    int pg = (yy1 - 22554 * cb1 - 46802 * cr1) >> 16;
    int pb = (yy1 + 116130 * cb1) >> 16;
    if (pr < 0) pr = 0; else if (pr > 0xff) pr = 0xff;
    if (pg < 0) pg = 0; else if (pg > 0xff) pg = 0xff;
    if (pb < 0) pb = 0; else if (pb > 0xff) pb = 0xff;
    *r = (uint8_t)pr; *g = (uint8_t)pg; *b = (uint8_t)pb;
}

static inline size_t chroma_index(const ipo_image *s, int x, int y)
{
    switch (s->layout) {
    case IPO_YCBCR444: return (size_t)y * (size_t)s->stride[1] + (size_t)x;
    case IPO_YCBCR422: return (size_t)y * (size_t)s->stride[1] + (size_t)(x / 2);
    case IPO_YCBCR420: return (size_t)(y / 2) * (size_t)s->stride[1] + (size_t)(x / 2);
    default:           return (size_t)(y / 2) * (size_t)s->stride[1] + (size_t)x; /* 440 */
    }
}

/* 16-bit premultiplied sample of one source pixel, per concrete type
 * (scaleX_RGBA / scaleX_NRGBA / scaleX_Gray / scaleX_YCbCr4xx). Returns 1 when
 * the alpha lane is the constant 1.0 (Gray, YCbCr). */
static inline int sample16(const ipo_image *s, int x, int y, uint32_t p[4])
{
    switch (s->layout) {
    case IPO_RGBA8: {
        const uint8_t *q = s->plane[0] + (size_t)y * (size_t)s->stride[0] + (size_t)x * 4;
        p[0] = (uint32_t)q[0] * 0x101; p[1] = (uint32_t)q[1] * 0x101;
        p[2] = (uint32_t)q[2] * 0x101; p[3] = (uint32_t)q[3] * 0x101;
        return 0;
    }
    case IPO_NRGBA8: {
        const uint8_t *q = s->plane[0] + (size_t)y * (size_t)s->stride[0] + (size_t)x * 4;
        uint32_t pa = (uint32_t)q[3] * 0x101;
        p[0] = (uint32_t)q[0] * pa / 0xff; p[1] = (uint32_t)q[1] * pa / 0xff;
        p[2] = (uint32_t)q[2] * pa / 0xff; p[3] = pa;
        return 0;
    }
    case IPO_GRAY8: {
        uint32_t v = (uint32_t)s->plane[0][(size_t)y * (size_t)s->stride[0] + (size_t)x] * 0x101;
        p[0] = p[1] = p[2] = v; p[3] = 0xffff;
        return 1;
    }
    /* ---- generic path: src.At(x, y).RGBA() / RGBA64At (x/image draw/impl.go scaleX_Image, scaleX_RGBA64Image);
     * alpha is data here, accumulated like the colour channels ---- */
    case IPO_RGBA64: { /* color.RGBA64.RGBA(): the stored values */
        const uint8_t *q = s->plane[0] + (size_t)y * (size_t)s->stride[0] + (size_t)x * 8;
        p[0] = be16(q); p[1] = be16(q + 2); p[2] = be16(q + 4); p[3] = be16(q + 6);
        return 0;
    }
    case IPO_NRGBA64: { /* color.NRGBA64.RGBA(): c = C * A / 0xffff (uint32) */
        const uint8_t *q = s->plane[0] + (size_t)y * (size_t)s->stride[0] + (size_t)x * 8;
        uint32_t a = be16(q + 6);
        p[0] = be16(q) * a / 0xffff; p[1] = be16(q + 2) * a / 0xffff; p[2] = be16(q + 4) * a / 0xffff; p[3] = a;
        return 0;
    }
    case IPO_GRAY16: { /* color.Gray16.RGBA(): (y, y, y, 0xffff) */
        uint32_t v = be16(s->plane[0] + (size_t)y * (size_t)s->stride[0] + (size_t)x * 2);
        p[0] = p[1] = p[2] = v; p[3] = 0xffff;
        return 0;
    }
    case IPO_PALETTED_RGBA: { /* Palette[i].(color.RGBA).RGBA(): c |= c << 8 */
        const uint8_t *e = s->plane[1] + 4 * (size_t)s->plane[0][(size_t)y * (size_t)s->stride[0] + (size_t)x];
        p[0] = (uint32_t)e[0] * 0x101; p[1] = (uint32_t)e[1] * 0x101; p[2] = (uint32_t)e[2] * 0x101; p[3] = (uint32_t)e[3] * 0x101;
        return 0;
    }
    case IPO_PALETTED_NRGBA: { /* Palette[i].(color.NRGBA).RGBA(): c |= c << 8; c *= A; c /= 0xff; a |= a << 8 */
        const uint8_t *e = s->plane[1] + 4 * (size_t)s->plane[0][(size_t)y * (size_t)s->stride[0] + (size_t)x];
        uint32_t a = e[3];
        p[0] = (uint32_t)e[0] * 0x101 * a / 0xff; p[1] = (uint32_t)e[1] * 0x101 * a / 0xff; p[2] = (uint32_t)e[2] * 0x101 * a / 0xff;
        p[3] = a * 0x101;
        return 0;
    }
    default: {
        size_t ci = chroma_index(s, x, y);
        int yy = s->plane[0][(size_t)y * (size_t)s->stride[0] + (size_t)x];
        ycbcr_to_rgb16(yy, s->plane[1][ci], s->plane[2][ci], &p[0], &p[1], &p[2]);
        p[3] = 0xffff;
        return 1;
    }
    }
}

static int layout_ok(const ipo_image *s)
{
    if (!s || s->layout < IPO_RGBA8 || s->layout > IPO_PALETTED_NRGBA || s->width <= 0 || s->height <= 0 || !s->plane[0]) return 0;
    if (s->layout >= IPO_YCBCR444 && s->layout <= IPO_YCBCR440) return s->plane[1] && s->plane[2];
    if (s->layout >= IPO_PALETTED_RGBA) return s->plane[1] != NULL;
    return 1;
}

int ipo_scale_bilinear(const ipo_image *src, int sx0, int sy0, int sw, int sh,
                       uint8_t *dst, int dst_stride, int dw, int dh, int op)
{
    if (!layout_ok(src) || !dst) return -1;
    if (dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0) return 0; /* adr.Empty() || sr.Empty() */
    if (sx0 < 0 || sy0 < 0 || sx0 + sw > src->width || sy0 + sh > src->height) return -1;

    axis_t hz, vt;
    if (axis_build(&hz, dw, sw)) return -1;
    if (axis_build(&vt, dh, sh)) { axis_free(&hz); return -1; }
    if (op == IPO_OP_OVER && image_opaque(src)) op = IPO_OP_SRC;

    /* fresh zeroed temp per call, as make([][4]float64, dw*sh) */
    double (*tmp)[4] = (double (*)[4])calloc((size_t)dw * (size_t)sh, sizeof(double[4]));
    if (!tmp) { axis_free(&hz); axis_free(&vt); return -1; }

    /* pass 1: scaleX_<type> */
    size_t t = 0;
    for (int32_t y = 0; y < sh; y++) {
        for (int32_t x = 0; x < dw; x++) {
            span_t s = hz.spans[x];
            double pr = 0, pg = 0, pb = 0, pa = 0;
            int const_alpha = 0;
            for (int32_t k = s.i; k < s.j; k++) {
                uint32_t p[4];
                double w = hz.taps[k].weight;
                const_alpha = sample16(src, sx0 + hz.taps[k].coord, sy0 + y, p);
                pr += (double)p[0] * w;
                pg += (double)p[1] * w;
                pb += (double)p[2] * w;
                if (!const_alpha) pa += (double)p[3] * w;
            }
            if (src->layout >= IPO_GRAY8 && src->layout <= IPO_YCBCR440) const_alpha = 1;
            tmp[t][0] = pr * s.inv_total_ffff;
            tmp[t][1] = pg * s.inv_total_ffff;
            tmp[t][2] = pb * s.inv_total_ffff;
            tmp[t][3] = const_alpha ? 1.0 : pa * s.inv_total_ffff;
            t++;
        }
    }

    /* pass 2: scaleY_RGBA_{Src,Over}, column-major like upstream */
    for (int32_t dx = 0; dx < dw; dx++) {
        uint8_t *d = dst + (size_t)dx * 4;
        for (int32_t dy = 0; dy < dh; dy++) {
            span_t s = vt.spans[dy];
            double pr = 0, pg = 0, pb = 0, pa = 0;
            for (int32_t k = s.i; k < s.j; k++) {
                const double *p = tmp[(size_t)vt.taps[k].coord * (size_t)dw + (size_t)dx];
                double w = vt.taps[k].weight;
                pr += p[0] * w;
                pg += p[1] * w;
                pb += p[2] * w;
                pa += p[3] * w;
            }
            if (pr > pa) pr = pa;
            if (pg > pa) pg = pa;
            if (pb > pa) pb = pa;
            if (op == IPO_OP_SRC) {
                d[0] = (uint8_t)(ftou(pr * s.inv_total) >> 8);
                d[1] = (uint8_t)(ftou(pg * s.inv_total) >> 8);
                d[2] = (uint8_t)(ftou(pb * s.inv_total) >> 8);
                d[3] = (uint8_t)(ftou(pa * s.inv_total) >> 8);
            } else {
                uint32_t pr0 = ftou(pr * s.inv_total);
                uint32_t pg0 = ftou(pg * s.inv_total);
                uint32_t pb0 = ftou(pb * s.inv_total);
                uint32_t pa0 = ftou(pa * s.inv_total);
                uint32_t pa1 = (0xffff - pa0) * 0x101;
                d[0] = (uint8_t)(((uint32_t)d[0] * pa1 / 0xffff + pr0) >> 8);
                d[1] = (uint8_t)(((uint32_t)d[1] * pa1 / 0xffff + pg0) >> 8);
                d[2] = (uint8_t)(((uint32_t)d[2] * pa1 / 0xffff + pb0) >> 8);
                d[3] = (uint8_t)(((uint32_t)d[3] * pa1 / 0xffff + pa0) >> 8);
            }
            d += dst_stride;
        }
    }
    free(tmp);
    axis_free(&hz);
    axis_free(&vt);
    return 0;
}

/* operations/resize.go:121-125 */
int ipo_resize_image(const ipo_image *src, int dw, int dh, uint8_t *dst)
{
    if (dw <= 0 || dh <= 0) return 0;
    memset(dst, 0, (size_t)dw * (size_t)dh * 4); /* image.NewRGBA */
    return ipo_scale_bilinear(src, 0, 0, src->width, src->height, dst, dw * 4, dw, dh,
                              IPO_OP_OVER);
}

/* operations/thumbnail.go:114-132 */
int ipo_crop_and_resize(const ipo_image *src, int size, uint8_t *dst)
{
    int cx, cy, cs;
    if (!layout_ok(src)) return -1;
    ipo_crop_square(src->width, src->height, &cx, &cy, &cs);
    uint8_t *cropped = (uint8_t *)calloc((size_t)cs * (size_t)cs, 4); /* image.NewRGBA */
    if (!cropped) return -1;
    int rc = ipo_scale_bilinear(src, cx, cy, cs, cs, cropped, cs * 4, cs, cs, IPO_OP_OVER);
    if (rc == 0) {
        ipo_image c;
        memset(&c, 0, sizeof c);
        c.layout = IPO_RGBA8;
        c.width = cs; c.height = cs;
        c.plane[0] = cropped; c.stride[0] = cs * 4;
        rc = ipo_resize_image(&c, size, size, dst);
    }
    free(cropped);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Spec W: stdlib image/draw                                                 */
/* ------------------------------------------------------------------------ */

/* draw.Draw(dst *image.RGBA, b, src, Point{}, draw.Src), by source type:
 * RGBA copy; drawNRGBASrc; imageutil.DrawYCbCr; drawGray. */
int ipo_draw_src(const ipo_image *src, uint8_t *dst, int dst_stride)
{
    if (!layout_ok(src) || !dst) return -1;
    for (int y = 0; y < src->height; y++) {
        uint8_t *d = dst + (size_t)y * (size_t)dst_stride;
        switch (src->layout) {
        case IPO_RGBA8:
            memcpy(d, src->plane[0] + (size_t)y * (size_t)src->stride[0], (size_t)src->width * 4);
            break;
        case IPO_NRGBA8: {
            const uint8_t *q = src->plane[0] + (size_t)y * (size_t)src->stride[0];
            for (int x = 0; x < src->width; x++, q += 4, d += 4) {
                uint32_t sa = (uint32_t)q[3] * 0x101;
                d[0] = (uint8_t)(((uint32_t)q[0] * sa / 0xff) >> 8);
                d[1] = (uint8_t)(((uint32_t)q[1] * sa / 0xff) >> 8);
                d[2] = (uint8_t)(((uint32_t)q[2] * sa / 0xff) >> 8);
                d[3] = (uint8_t)(sa >> 8);
            }
            break;
        }
        case IPO_GRAY8: {
            const uint8_t *q = src->plane[0] + (size_t)y * (size_t)src->stride[0];
            for (int x = 0; x < src->width; x++, d += 4) {
                d[0] = d[1] = d[2] = q[x];
                d[3] = 0xff;
            }
            break;
        }
        case IPO_RGBA64: case IPO_NRGBA64: case IPO_GRAY16: case IPO_PALETTED_RGBA: case IPO_PALETTED_NRGBA: {
            /* no fast path in image/draw: drawRGBA's generic loop, d = uint8(At(x, y).RGBA() >> 8) per channel */
            for (int x = 0; x < src->width; x++, d += 4) {
                uint32_t p[4];
                sample16(src, x, y, p);
                d[0] = (uint8_t)(p[0] >> 8); d[1] = (uint8_t)(p[1] >> 8); d[2] = (uint8_t)(p[2] >> 8); d[3] = (uint8_t)(p[3] >> 8);
            }
            break;
        }
        default: {
            const uint8_t *yp = src->plane[0] + (size_t)y * (size_t)src->stride[0];
            for (int x = 0; x < src->width; x++, d += 4) {
                size_t ci = chroma_index(src, x, y);
                ycbcr_to_rgb8(yp[x], src->plane[1][ci], src->plane[2][ci], &d[0], &d[1], &d[2]);
                d[3] = 0xff;
            }
            break;
        }
        }
    }
    return 0;
}

/* Go 1.24 image/draw drawGlyphOver; colour is color.RGBA (not re-premultiplied):
 * Uniform.RGBA() -> c*0x101 per channel.  All arithmetic uint32, wrapping. */
void ipo_glyph_over(uint8_t *dst, int dst_stride, const uint8_t rgba[4], const ipo_glyph *g)
{
    const uint32_t m = 0xffff;
    uint32_t sr = (uint32_t)rgba[0] * 0x101, sg = (uint32_t)rgba[1] * 0x101;
    uint32_t sb = (uint32_t)rgba[2] * 0x101, sa = (uint32_t)rgba[3] * 0x101;
    for (int y = g->y0, my = g->mp_y; y < g->y1; y++, my++) {
        uint8_t *d = dst + (size_t)y * (size_t)dst_stride + (size_t)g->x0 * 4;
        const uint8_t *mk = g->mask + (size_t)my * (size_t)g->mask_stride + (size_t)g->mp_x;
        for (int x = g->x0; x < g->x1; x++, d += 4, mk++) {
            uint32_t ma = *mk;
            if (ma == 0) continue;
            ma |= ma << 8;
            uint32_t a = (m - (sa * ma / m)) * 0x101;
            d[0] = (uint8_t)(((uint32_t)d[0] * a + sr * ma) / m >> 8);
            d[1] = (uint8_t)(((uint32_t)d[1] * a + sg * ma) / m >> 8);
            d[2] = (uint8_t)(((uint32_t)d[2] * a + sb * ma) / m >> 8);
            d[3] = (uint8_t)(((uint32_t)d[3] * a + sa * ma) / m >> 8);
        }
    }
}

/* operations/watermark.go:90-92 + :151 (per-glyph DrawMask in string order) */
int ipo_watermark(const ipo_image *src, uint8_t *dst, int dst_stride, const uint8_t rgba[4],
                  const ipo_glyph *glyphs, int n_glyphs)
{
    int rc = ipo_draw_src(src, dst, dst_stride);
    if (rc) return rc;
    for (int k = 0; k < n_glyphs; k++) {
        const ipo_glyph *g = &glyphs[k];
        if (g->x0 >= g->x1 || g->y0 >= g->y1) continue; /* dr.Empty() */
        if (g->x0 < 0 || g->y0 < 0 || g->x1 > src->width || g->y1 > src->height) return -1;
        ipo_glyph_over(dst, dst_stride, rgba, g);
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* What Go's image/jpeg writer derives from an *image.RGBA before its DCT     */
/* (Go 1.24 image/jpeg/writer.go rgbaToYCbCr + scale, image/color/ycbcr.go    */
/* RGBToYCbCr), written as the 4:2:0 image that yields the same blocks.       */
/* ------------------------------------------------------------------------ */
static void go_rgb_to_ycbcr(uint8_t r, uint8_t g, uint8_t b, uint8_t *yy, uint8_t *cb, uint8_t *cr)
{
    int32_t r1 = r, g1 = g, b1 = b;
    int32_t y = (19595 * r1 + 38470 * g1 + 7471 * b1 + (1 << 15)) >> 16;
    int32_t c = -11056 * r1 - 21712 * g1 + 32768 * b1 + (257 << 15);
    if (((uint32_t)c & 0xff000000u) == 0) c >>= 16; else c = ~(c >> 31);
    int32_t d = 32768 * r1 - 27440 * g1 - 5328 * b1 + (257 << 15);
    if (((uint32_t)d & 0xff000000u) == 0) d >>= 16; else d = ~(d >> 31);
    *yy = (uint8_t)y; *cb = (uint8_t)c; *cr = (uint8_t)d;
}

/* y: w x h; cb, cr: ((w+1)/2) x ((h+1)/2), tight strides.  Chroma of a 2x2 group = (sum + 2) >> 2 with coordinates
 * clamped to the image (the writer replicates the last column / row inside its 16x16 blocks). */
int ipo_rgba_to_ycbcr420(const uint8_t *rgba, int stride, int w, int h, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    if (!rgba || !y || !cb || !cr || w <= 0 || h <= 0) return -1;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const uint8_t *p = rgba + (size_t)j * (size_t)stride + (size_t)i * 4;
            uint8_t a, b, c;
            go_rgb_to_ycbcr(p[0], p[1], p[2], &a, &b, &c);
            y[(size_t)j * (size_t)w + (size_t)i] = a;
        }
    for (int cj = 0; cj < ch; cj++)
        for (int ci = 0; ci < cw; ci++) {
            int sb = 0, sr = 0;
            for (int dj = 0; dj < 2; dj++)
                for (int di = 0; di < 2; di++) {
                    int yy = 2 * cj + dj, xx = 2 * ci + di;
                    if (yy > h - 1) yy = h - 1;
                    if (xx > w - 1) xx = w - 1;
                    const uint8_t *p = rgba + (size_t)yy * (size_t)stride + (size_t)xx * 4;
                    uint8_t a, b, c;
                    go_rgb_to_ycbcr(p[0], p[1], p[2], &a, &b, &c);
                    sb += b; sr += c;
                }
            cb[(size_t)cj * (size_t)cw + (size_t)ci] = (uint8_t)((sb + 2) >> 2);
            cr[(size_t)cj * (size_t)cw + (size_t)ci] = (uint8_t)((sr + 2) >> 2);
        }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* CPU baseline driver (bench.py only)                                       */
/* ------------------------------------------------------------------------ */

typedef struct {
    const ipo_image *imgs;
    int n, ops, rw, rh, keep_aspect, thumb_size;
    const uint8_t *rgba;
    const ipo_glyph *glyphs;
    int n_glyphs;
    int next;
    uint64_t checksum;
    pthread_mutex_t mu;
} bench_job;

static uint64_t fold(const uint8_t *p, size_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (size_t k = 0; k < n; k += 97) h = (h ^ p[k]) * 1099511628211ULL;
    return h;
}

static void *bench_worker(void *arg)
{
    bench_job *job = (bench_job *)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int idx = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (idx >= job->n) break;
        const ipo_image *im = &job->imgs[idx];
        uint64_t h = 0;
        if (job->ops & 1) {
            int nw = job->rw, nh = job->rh;
            if (job->keep_aspect)
                ipo_keep_aspect_dims(im->width, im->height, job->rw, job->rh, &nw, &nh);
            uint8_t *out = (uint8_t *)malloc((size_t)nw * (size_t)nh * 4);
            ipo_resize_image(im, nw, nh, out);
            h ^= fold(out, (size_t)nw * (size_t)nh * 4);
            free(out);
        }
        if (job->ops & 2) {
            size_t sz = (size_t)job->thumb_size * (size_t)job->thumb_size * 4;
            uint8_t *out = (uint8_t *)malloc(sz);
            ipo_crop_and_resize(im, job->thumb_size, out);
            h ^= fold(out, sz);
            free(out);
        }
        if (job->ops & 4) {
            size_t sz = (size_t)im->width * (size_t)im->height * 4;
            uint8_t *out = (uint8_t *)calloc(sz, 1); /* image.NewRGBA zero-fills */
            ipo_watermark(im, out, im->width * 4, job->rgba, job->glyphs, job->n_glyphs);
            h ^= fold(out, sz);
            free(out);
        }
        pthread_mutex_lock(&job->mu);
        job->checksum ^= h + (uint64_t)idx;
        pthread_mutex_unlock(&job->mu);
    }
    return NULL;
}

double ipo_bench_batch(const ipo_image *imgs, int n, int n_threads, int ops, int rw, int rh,
                       int keep_aspect, int thumb_size, const uint8_t rgba[4],
                       const ipo_glyph *glyphs, int n_glyphs, uint64_t *checksum)
{
    bench_job job;
    memset(&job, 0, sizeof job);
    job.imgs = imgs; job.n = n; job.ops = ops; job.rw = rw; job.rh = rh;
    job.keep_aspect = keep_aspect; job.thumb_size = thumb_size;
    job.rgba = rgba; job.glyphs = glyphs; job.n_glyphs = n_glyphs;
    pthread_mutex_init(&job.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < n_threads; k++) pthread_create(&th[k], NULL, bench_worker, &job);
    for (int k = 0; k < n_threads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_mutex_destroy(&job.mu);
    if (checksum) *checksum = job.checksum;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
