/*
 * ip_oracle.h -- CPU oracle for the ImageProcessor worker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (imageprocessor_b200/,
 * include/, the C-ABI library) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (sj-shoff/ImageProcessor) ships no tests, no
 * golden images and no Go toolchain exists in the build container, so this is
 * a restatement of the published algorithms of its un-vendored dependencies:
 *   - golang.org/x/image v0.33.0  draw/scale.go, draw/impl.go   (go.mod:45)
 *   - Go 1.24.7 stdlib image/draw/draw.go, image/color/ycbcr.go,
 *     image/internal/imageutil                                   (go.mod:3)
 * anchored on the reference's own call sites:
 *   - operations/resize.go:61-75,121-125      (keep-aspect dims, resizeImage)
 *   - operations/thumbnail.go:48-64,114-132   (fit dims, cropAndResize)
 *   - operations/watermark.go:86-157,159-190  (draw.Draw Src, DrawString,
 *                                              parseColor, clamp)
 * It is pinned instead by fixtures computed with independent code (PyTorch's
 * antialiased bilinear in float64 + the x/image quantiser; big-int
 * drawGlyphOver; hand-derived geometry): tests/golden/make_golden.py ->
 * tests/golden/golden_v1.npz, checked by tests/test_golden.py.
 * oracle/verify_with_go/ regenerates them from the real reference
 * dependencies the first time a Go toolchain is available.
 */
#ifndef IP_ORACLE_H
#define IP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Concrete raster types image.Decode can hand the ops (image_processor.go:47).
 * Numeric values are shared with include/ipgpu.h on purpose. */
enum {
    IPO_RGBA8    = 0, /* *image.RGBA: alpha-premultiplied, 4 B/px            */
    IPO_NRGBA8   = 1, /* *image.NRGBA: straight alpha, 4 B/px                */
    IPO_GRAY8    = 2, /* *image.Gray                                         */
    IPO_YCBCR444 = 3, /* *image.YCbCr, planar                                */
    IPO_YCBCR422 = 4,
    IPO_YCBCR420 = 5,
    IPO_YCBCR440 = 6,
    /* 16-bit types (a 16-bit PNG): Pix holds big-endian uint16 per channel, as Go stores them.  x/image reaches them
     * through the generic / RGBA64Image path: src.RGBA64At(x, y) (== At(x, y).RGBA()) per tap, alpha accumulated. */
    IPO_RGBA64   = 7, /* *image.RGBA64: alpha-premultiplied, 8 B/px         */
    IPO_NRGBA64  = 8, /* *image.NRGBA64: straight alpha, 8 B/px             */
    IPO_GRAY16   = 9, /* *image.Gray16, 2 B/px                              */
    /* *image.Paletted (GIF, paletted PNG): plane[0] = 1 index byte per pixel, plane[1] = 256 x 4 palette bytes.
     * Oracle-only: the product takes the palette expanded on the host (RGBA8 / NRGBA8), which tests prove identical. */
    IPO_PALETTED_RGBA  = 10, /* palette entries are color.RGBA (GIF; PNG without tRNS) */
    IPO_PALETTED_NRGBA = 11, /* palette entries are color.NRGBA (PNG with a tRNS chunk) */
};

typedef struct {
    int32_t layout;
    int32_t width, height;
    const uint8_t *plane[3]; /* RGBA/NRGBA/Gray(16): plane[0]; YCbCr: Y, Cb, Cr; Paletted: indices, palette */
    int32_t stride[3];       /* bytes per row of each plane                   */
} ipo_image;

enum { IPO_OP_OVER = 0, IPO_OP_SRC = 1 };

/* ---- geometry (all double/int arithmetic exactly as the reference) ------- */
/* resize.go:63-72: ratio=min(W/ow,H/oh); nw=int(ow*ratio); nh=int(oh*ratio) */
void ipo_keep_aspect_dims(int ow, int oh, int w, int h, int *nw, int *nh);
/* thumbnail.go:52-63 */
void ipo_thumb_fit_dims(int ow, int oh, int size, int *nw, int *nh);
/* thumbnail.go:115-127 */
void ipo_crop_square(int ow, int oh, int *cx, int *cy, int *cs);
/* watermark.go:116-148: anchor (dot) for a position string; text metrics given */
void ipo_watermark_anchor(const char *position, int W, int H, int width_px,
                          int height_px, int *x, int *y);
/* watermark.go:116-118: ceil(Int26_6(fontSize*64*1.2)) */
int ipo_watermark_height_px(double font_size);
/* watermark.go:159-186; returns 0 ok, -1 parse error (caller falls back to
 * black {0,0,0,uint8(255*opacity)} as watermark.go:94-97 does). */
int ipo_parse_color(const char *s, double opacity, uint8_t rgba[4]);

/* ---- Spec R: x/image draw.BiLinear.Scale -------------------------------- */
/* Number of contribs and per-axis tables, for KATs. weights/coords may be NULL.
 * Returns total number of contribs. starts has dw+1 entries (CSR offsets). */
int ipo_distrib(int dw, int sw, int32_t *starts, int32_t *coords,
                double *weights, double *inv_total);

/* BiLinear.Scale(dst, dst.Bounds(), src, Rect(sx0,sy0,sx0+sw,sy0+sh), op, nil)
 * into an RGBA8 dst of dw x dh (dst_stride bytes per row).  For op=OVER the
 * existing dst bytes take part, as in scaleY_RGBA_Over.  Returns 0 / -1. */
int ipo_scale_bilinear(const ipo_image *src, int sx0, int sy0, int sw, int sh,
                       uint8_t *dst, int dst_stride, int dw, int dh, int op);

/* resize.go:121-125 resizeImage: zero-filled RGBA + Scale(..., Over).        */
int ipo_resize_image(const ipo_image *src, int dw, int dh, uint8_t *dst);
/* thumbnail.go:114-132 cropAndResize: 1:1 Scale into cropped RGBA, then
 * resizeImage(cropped,size,size).  dst is size*size*4.                       */
int ipo_crop_and_resize(const ipo_image *src, int size, uint8_t *dst);

/* ---- Spec W: stdlib image/draw ------------------------------------------ */
/* draw.Draw(dst *RGBA, bounds, src, Point{}, Src)  (watermark.go:91-92)      */
int ipo_draw_src(const ipo_image *src, uint8_t *dst, int dst_stride);

/* One draw.DrawMask(dst, dr, Uniform(col), ZP, mask *Alpha, mp, Over) as
 * freetype's DrawString issues per glyph -> stdlib drawGlyphOver.
 * dr = [x0,x1) x [y0,y1) already clipped to dst; mask origin at (0,0).       */
typedef struct {
    int32_t x0, y0, x1, y1;  /* destination rectangle                       */
    int32_t mp_x, mp_y;      /* mask point matching (x0,y0)                 */
    int32_t mask_stride;
    int32_t mask_w, mask_h;
    const uint8_t *mask;     /* *image.Alpha pixels                          */
} ipo_glyph;

void ipo_glyph_over(uint8_t *dst, int dst_stride, const uint8_t rgba[4],
                    const ipo_glyph *g);

/* addTextWatermark's raster part: draw_src then glyphs in string order.      */
int ipo_watermark(const ipo_image *src, uint8_t *dst, int dst_stride,
                  const uint8_t rgba[4], const ipo_glyph *glyphs, int n_glyphs);

/* ---- what image/jpeg's writer makes of an *image.RGBA (Go 1.24 writer.go rgbaToYCbCr + scale; color.RGBToYCbCr) ---- */
/* Planar 4:2:0 whose jpeg.Encode output equals that of the RGBA image: y is w x h, cb / cr ((w+1)/2) x ((h+1)/2). */
int ipo_rgba_to_ycbcr420(const uint8_t *rgba, int stride, int w, int h, uint8_t *y, uint8_t *cb, uint8_t *cr);

/* ---- Go 1.24 image/jpeg ENCODER (ip_jpeg_oracle.c): jpeg.Encode(w, m, &jpeg.Options{Quality: quality}) ------------- */
/* The checker of the device-side JPEG writer (ipg_op.dst_layout = IPG_LAYOUT_JPEG).  Each returns the number of bytes
 * of the complete file written to out, 0 for a bad argument, (size_t)-1 when cap is too small.
 * m = *image.RGBA (what resizeImage / cropAndResize / addTextWatermark return; resize.go:78-91, watermark.go:66-79). */
size_t ipo_jpeg_encode_rgba(const uint8_t *rgba, int stride, int w, int h, int quality, uint8_t *out, size_t cap);
/* m = *image.YCbCr (layout IPO_YCBCR444..440): the writer's yCbCrToYCbCr path. */
size_t ipo_jpeg_encode_ycbcr(const ipo_image *img, int quality, uint8_t *out, size_t cap);
/* m = *image.Gray: one component, 8 x 8 MCUs. */
size_t ipo_jpeg_encode_gray(const uint8_t *pix, int stride, int w, int h, int quality, uint8_t *out, size_t cap);

/* ---- CPU baseline driver (bench.py only) -------------------------------- */
/* Runs resize(+thumb crop)(+watermark) over n images of identical geometry on
 * n_threads pthreads, one image per thread at a time, fresh temp buffers per
 * call like the reference.  ops bit0=resize bit1=thumb bit2=watermark.
 * Outputs may be NULL (results discarded after a checksum).  Returns seconds. */
double ipo_bench_batch(const ipo_image *imgs, int n, int n_threads, int ops,
                       int rw, int rh, int keep_aspect, int thumb_size,
                       const uint8_t rgba[4], const ipo_glyph *glyphs,
                       int n_glyphs, uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
