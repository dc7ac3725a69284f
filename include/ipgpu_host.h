/*
 * ipgpu_host.h -- C ABI of the host layer of libipgpu.so: a C++ mirror of the
 * reference's processor package for the worker hot path, sitting on top of the
 * raster entry points of ipgpu.h.
 *
 * The reference is Go; no Go toolchain exists in the build image, so the host side
 * above the raster C ABI is written in C++ with the reference's own names, argument
 * meaning and error strings (paths relative to the reference root):
 *
 *   iph_process        processor.ImageProcessor.Process     internal/usecase/processor/image_processor.go:39-102
 *                      applyOperation / generatePath /      :104-182
 *                      getContentType
 *                      Resizer.Process / processStaticImage operations/resize.go:26-119
 *                      Thumbnailer.Process / cropAndResize  operations/thumbnail.go:25-132
 *                      Watermarker.Process / addTextWatermark / parseColor
 *                                                           operations/watermark.go:40-190
 *                      freetype.Context.DrawString / glyph  golang/freetype@e2365dfdc4a0 freetype.go
 *                      (pen movement, kerning, the (glyph, quarter-pixel) mask cache, the
 *                       per-rune DrawMask rectangle with mask point (0, dr.Min.Y - glyphRect.Min.Y))
 *   iph_process_batch  Worker.processWorker / processMessage internal/worker/worker.go:112-234,
 *                      reshaped to take a batch of decoded images (the one loop the GPU
 *                      path changes): all tickets are submitted, then awaited.
 *   task / result JSON domain.ProcessingTask / ProcessingResult  internal/domain/task.go:3-23
 *                      (no json tags: wire keys are the Go field names)
 *
 * What stays with the caller, as in the reference: image.Decode, the jpeg/png/gif
 * encoders, the truetype face + rasteriser, and fileRepository.SaveProcessed -- all
 * reached through the callbacks below (a Go host would not use this layer at all: it
 * keeps its own processor package and binds ipgpu.h directly).
 */
#ifndef IPGPU_HOST_H
#define IPGPU_HOST_H

#include "ipgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iph_processor iph_processor;

/* *image.Alpha mask of one rune plus its placement, as truetype/raster produce it for
 * freetype.Context.glyph(): mask bounds offset by (off_x, off_y) from the integer pen. */
typedef struct {
    int32_t advance_26_6;      /* Font.HMetric(scale, index).AdvanceWidth: what DrawString moves the pen by */
    int32_t off_x, off_y;      /* mask origin relative to (pen.X>>6, pen.Y>>6) */
    int32_t mask_w, mask_h, mask_stride;
    const uint8_t *mask;       /* valid until the next glyph_mask call on this user */
} iph_glyph;

typedef struct {
    void *user;
    /* jpeg.Encode(q) / png.Encode / gif.Encode of an *image.RGBA. format is "jpeg", "png" or
     * "gif". *out stays owned by the callee until release(). Return 0 on success. */
    int (*encode)(void *user, const uint8_t *rgba, int w, int h, int stride, const char *format,
                  int jpeg_quality, uint8_t **out, size_t *out_len);
    void (*release)(void *user, uint8_t *buf);
    /* fileRepository.SaveProcessed(ctx, path, reader, size, contentType). Return 0 on success. */
    int (*save_processed)(void *user, const char *path, const uint8_t *data, size_t size,
                          const char *content_type);
    /* truetype.Face.GlyphAdvance(rune): returns 0 and the advance, or non-zero if the face
     * lacks the rune (then it adds nothing to the width, watermark.go:110-115). */
    int (*glyph_advance)(void *user, uint32_t rune, double font_size, int32_t *advance_26_6);
    /* freetype.Context.rasterize(index, fx, fy): mask at sub-pixel (fx, fy) = (p.X & 63, p.Y & 63). */
    int (*glyph_mask)(void *user, uint32_t rune, double font_size, int fx, int fy, iph_glyph *out);
    /* Font.Kern(scale, prev, index), 26.6; may be NULL (no kerning). */
    int32_t (*kern)(void *user, uint32_t prev_rune, uint32_t rune, double font_size);
    /* Font.Index(rune): the glyph index freetype.Context.glyph() keys its 256 x 4 direct-mapped mask cache on
     * (slot = index % 256, quarter-pixel x).  May be NULL: the rune then stands in for the index. */
    uint32_t (*glyph_index)(void *user, uint32_t rune);
} iph_callbacks;

/* ctx may be NULL for host-only use (parameter handling, paths, JSON): raster work then
 * fails with "no raster engine" -- there is no CPU fallback. */
IPG_API iph_processor *iph_processor_new(ipg_ctx *ctx, const iph_callbacks *cb);
IPG_API void iph_processor_free(iph_processor *p);
/* Opt-in (encode half; off by default): results whose target format is JPEG -- format_switch of
 * resize.go:78-91 / thumbnail.go:68-81 / watermark.go:66-79 -- are encoded on the device as
 * jpeg.Encode(buf, img, &jpeg.Options{Quality: 85}) would encode them (ipg_op.dst_layout = IPG_LAYOUT_JPEG) and go
 * straight to save_processed; the encode callback is then called for PNG and GIF targets only.  Returns 0. */
IPG_API int iph_set_device_jpeg(iph_processor *p, int enabled);

/* ImageProcessor.Process on an already decoded image.  task_json is the broker message
 * value (ProcessingTask); decoded_format what image.Decode returned ("jpeg", "png", "gif").
 * img == NULL stands for a failed image.Decode (decode_error = its message, may be NULL).
 * *result_json receives the ProcessingResult marshalled as encoding/json would (free with
 * iph_free).  Returns 0 when Process returns a nil error, -1 otherwise; like the reference
 * a populated result comes back in both cases and the error text is in iph_last_error(). */
IPG_API int iph_process(iph_processor *p, const char *task_json, const ipg_image_desc *img,
                        const char *decoded_format, const char *decode_error, char **result_json);

/* The batching processWorker: n messages with their decoded images.  Every task's raster
 * work is submitted before any is awaited, so the engine coalesces them into batched
 * launches; rc[i] / result_json[i] / (optionally) errors[i] are per message. */
IPG_API int iph_process_batch(iph_processor *p, int n, const char *const *task_json,
                              const ipg_image_desc *imgs, const char *const *decoded_formats,
                              char **result_json, int *rc, char **errors);

IPG_API void iph_free(void *p);
IPG_API const char *iph_last_error(void);

/* ---- pure host helpers (reference arithmetic), exposed for tests ---------------- */
/* watermark.go:159-186 + caller fallback :94-97. Returns 0, or -1 when the reference
 * would log "Color parse error" and use black. */
IPG_API int iph_parse_color(const char *s, double opacity, uint8_t rgba[4]);
/* watermark.go:116,118 */
IPG_API int iph_watermark_height_px(double font_size);
/* watermark.go:121-148 */
IPG_API void iph_watermark_anchor(const char *position, int W, int H, int width_px, int height_px,
                                  int *x, int *y);
/* image_processor.go:129-162; params_json is the operation's Parameters object. */
IPG_API char *iph_generate_path(const char *image_id, const char *operation, const char *format,
                                const char *params_json);
/* image_processor.go:164-182 */
IPG_API const char *iph_content_type(const char *path);

#ifdef __cplusplus
}
#endif
#endif /* IPGPU_HOST_H */
