/*
 * ipgpu.h -- C ABI of libipgpu.so: the B200 (sm_100a) raster engine behind
 * ImageProcessor's worker hot path.
 *
 * What it replaces (reference = sj-shoff/ImageProcessor, paths relative to the
 * reference root): the three pure raster functions the worker reaches through
 * processor.ImageProcessor.Process (internal/usecase/processor/image_processor.go:39)
 *
 *   resizeImage(img, w, h)                  operations/resize.go:121-125
 *   (*Thumbnailer).cropAndResize(img, size) operations/thumbnail.go:114-132
 *   (*Watermarker).addTextWatermark(...)    operations/watermark.go:86-157
 *        (its raster part: draw.Draw(...,Src) at :91-92 and the per-glyph
 *         draw.DrawMask(...,Over) issued by freetype DrawString at :151)
 *
 * Everything else on that path stays on the host exactly as the reference has
 * it: parameter parsing, geometry (keep-aspect dims, crop square, watermark
 * anchor), glyph rasterisation, JPEG/PNG codecs, object paths, SaveProcessed.
 * Geometry is computed by the caller in double/int exactly as the reference
 * and passed in; the library never re-derives it.
 *
 * Rules of the boundary
 *   - plain C, pointers and sizes only; no C++/torch types.
 *   - every entry returns IPG_OK (0) or a negative ipg_status; the message for
 *     the calling thread is in ipg_last_error().  Nothing aborts, exits or
 *     throws across the ABI (a native crash would bypass the worker's
 *     recover(), internal/worker/worker.go:151-163).
 *   - the caller owns all host buffers.  ipg_submit() does not retain caller
 *     pointers past its return unless they lie inside memory obtained from
 *     ipg_alloc_pinned() (cgo pointer rule); outputs are valid after
 *     ipg_wait() returns IPG_OK.
 *   - there is no CPU fallback: without a usable CUDA device ipg_init() fails.
 *   - re-entrant: submit/wait may be called concurrently from any number of
 *     threads (WORKER_CONCURRENCY goroutines, worker.go:90-96).
 */
#ifndef IPGPU_H
#define IPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPG_ABI_VERSION 3

#if defined(__GNUC__)
#define IPG_API __attribute__((visibility("default")))
#else
#define IPG_API
#endif

typedef enum {
    IPG_OK = 0,
    IPG_ERR_INVALID = -1,   /* bad argument / unsupported combination        */
    IPG_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail)   */
    IPG_ERR_NOMEM = -3,     /* host or device allocation failed              */
    IPG_ERR_TIMEOUT = -4,   /* ipg_wait timed out; the ticket is still live  */
    IPG_ERR_NO_DEVICE = -5, /* no CUDA device: there is no CPU fallback      */
    IPG_ERR_SHUTDOWN = -6,  /* context is being destroyed                    */
    IPG_ERR_INTERNAL = -7,
} ipg_status;

/* Concrete raster types image.Decode hands the ops (image_processor.go:47). */
typedef enum {
    IPG_LAYOUT_RGBA8 = 0,    /* *image.RGBA  (alpha-premultiplied)            */
    IPG_LAYOUT_NRGBA8 = 1,   /* *image.NRGBA (straight alpha)                 */
    IPG_LAYOUT_GRAY8 = 2,    /* *image.Gray                                   */
    IPG_LAYOUT_YCBCR444 = 3, /* *image.YCbCr planar; chroma dims as image.NewYCbCr */
    IPG_LAYOUT_YCBCR422 = 4,
    IPG_LAYOUT_YCBCR420 = 5,
    IPG_LAYOUT_YCBCR440 = 6,
    /* 16-bit types (a 16-bit PNG): Pix as Go stores it, big-endian uint16 per channel.  x/image reaches them
     * through its generic / RGBA64Image path (RGBA64At per tap).  Low volume: vertical upscales run in the fp32 k_direct,
     * everything else whole-image in float64 (k_exact_tiles); the watermark frame is uint8(At().RGBA() >> 8). */
    IPG_LAYOUT_RGBA64 = 7,   /* *image.RGBA64 (alpha-premultiplied), 8 bytes per pixel */
    IPG_LAYOUT_NRGBA64 = 8,  /* *image.NRGBA64 (straight alpha), 8 bytes per pixel     */
    IPG_LAYOUT_GRAY16 = 9,   /* *image.Gray16, 2 bytes per pixel                        */
    /* Destination only (ipg_op.dst_layout, new in ABI 3): the result as the JPEG FILE the reference makes of it,
     * jpeg.Encode(buf, img, &jpeg.Options{Quality: q}) (operations/resize.go:78-91, watermark.go:66-79), encoded on the
     * device.  See ipg_op.dst_layout. */
    IPG_LAYOUT_JPEG = 16,
} ipg_layout;

typedef enum {
    IPG_MEM_HOST = 0,   /* pageable or pinned host memory                     */
    IPG_MEM_DEVICE = 1, /* device memory on the ticket's device (see ipg_submit_on) */
} ipg_memspace;

/* How the float part (resize / thumbnail) is evaluated.  The watermark blend is
 * integer and bit-exact in every mode. */
typedef enum {
    /* fp32 streaming kernel + fp64 reference-order re-evaluation of every output
     * byte whose fp32 value lies within a proven error bound of a quantiser
     * step: output is byte-identical to the fp64 reference algorithm. */
    IPG_PRECISION_EXACT = 0,
    /* fp32 streaming kernel only: max |diff| <= 1 per channel. */
    IPG_PRECISION_FAST = 1,
    /* whole image in fp64, reference operation order (slow; verification). */
    IPG_PRECISION_REFERENCE = 2,
} ipg_precision;

typedef struct {
    uint32_t struct_size;        /* = sizeof(ipg_config)                      */
    int32_t precision;           /* ipg_precision, default EXACT              */
    int32_t lanes_per_device;    /* streams + staging sets per device, default 3 */
    int32_t max_batch;           /* tickets coalesced into one launch sequence, default 16 */
    int32_t batch_window_us;     /* how long the batcher waits to fill a batch; <0: default 200, 0: none */
    int32_t fuse_targets;        /* resample targets per pass over the source: 0/1 = one per pass (default, measured fastest on
                                    B200), 2 = fuse two (general instantiation), 3 = fuse a narrow- and a wide-support target
                                    in the lean fused instantiation when eligible */
    uint64_t lane_device_bytes;  /* device arena per lane, default 1 GiB      */
    uint64_t lane_pinned_bytes;  /* pinned staging per lane (for non-pinned callers), default 256 MiB */
} ipg_config;

typedef struct ipg_ctx ipg_ctx;
typedef uint64_t ipg_ticket;

/* A decoded source image (what image.Decode returned). plane[0] only for
 * RGBA/NRGBA/Gray; Y, Cb, Cr for YCbCr. stride in bytes. */
typedef struct {
    int32_t layout;    /* ipg_layout */
    int32_t memspace;  /* ipg_memspace */
    int32_t width, height;
    const void *plane[3];
    int32_t stride[3];
    int32_t opaque_hint; /* 1: caller knows alpha==255 everywhere (JPEG); 0: unknown */
} ipg_image_desc;

/* One draw.DrawMask(dst, dr, Uniform(col), ZP, mask *image.Alpha, mp, Over) as
 * freetype's DrawString issues per rune (golang/freetype freetype.go DrawString;
 * call site watermark.go:151).  dr must already be clipped to the image. */
typedef struct {
    int32_t x0, y0, x1, y1; /* dr                                             */
    int32_t mp_x, mp_y;     /* mask point for (x0,y0) (freetype passes mp.X=0) */
    int32_t mask_w, mask_h, mask_stride;
    int32_t reserved0;
    const uint8_t *mask;    /* host memory; copied during ipg_submit          */
} ipg_glyph;

typedef enum {
    /* resizeImage(img, dst_w, dst_h): BiLinear.Scale of the whole image.
     * Also used for the non-crop thumbnail (thumbnail.go:52-64). */
    IPG_OP_RESIZE = 1,
    /* cropAndResize: 1:1 Scale of rect (quantises to RGBA8) then resizeImage to
     * dst_w x dst_h (= size x size). */
    IPG_OP_THUMB_CROP = 2,
    /* addTextWatermark raster part: full-frame draw.Draw(Src) convert/copy then
     * glyphs blended in order. dst is width x height RGBA8. */
    IPG_OP_WATERMARK = 3,
} ipg_op_kind;

typedef struct {
    int32_t kind;             /* ipg_op_kind                                  */
    int32_t dst_w, dst_h;     /* output size (host-computed, reference arithmetic) */
    int32_t rect_x, rect_y, rect_w, rect_h; /* THUMB_CROP source square       */
    uint8_t color[4];         /* WATERMARK: color.RGBA{R,G,B,A} from parseColor, not premultiplied */
    int32_t n_glyphs;
    const ipg_glyph *glyphs;
    /* destination, caller-owned, RGBA8 (image.RGBA), dst_h rows of dst_stride bytes */
    void *dst;
    int32_t dst_stride;
    int32_t dst_memspace;     /* ipg_memspace                                 */
    int32_t flags;            /* ipg_op_flags (new in ABI 2)                  */
    /* Destination layout (new in ABI 2).  0 / IPG_LAYOUT_RGBA8: the *image.RGBA the reference's raster functions return.
     * IPG_LAYOUT_YCBCR420 (opt-in, SURVEY 8f-3 first step, for results that will be JPEG-encoded): the result is handed
     * back as the planar 4:2:0 image Go's image/jpeg writer derives from that *image.RGBA before its DCT -- per pixel the
     * integer color.RGBToYCbCr of the (premultiplied) R, G, B bytes, chroma averaged 2 x 2 as writer.go's scale() does,
     * (sum + 2) >> 2, with the edge pixel replicated for odd sizes -- so jpeg.Encode of
     *   &image.YCbCr{Y, Cb, Cr, YStride, CStride, SubsampleRatio420, Rect(0, 0, w, h)}
     * emits the same bytes as jpeg.Encode of the RGBA result when dst_w and dst_h are multiples of 16 (otherwise the
     * writer pads its last MCUs from the edge chroma SAMPLE of a YCbCr image but from the edge PIXEL of an RGBA one, and
     * the files differ in those blocks: use IPG_LAYOUT_JPEG for byte identity at any size), and 1.5 instead of 4 bytes
     * per pixel cross PCIe.
     * dst / dst_stride are then the Y plane (dst_w bytes per row), dst_cb / dst_cr / dst_cstride the (dst_w+1)/2 x
     * (dst_h+1)/2 chroma planes.  Host planes must lie in ipg_alloc_pinned memory (no staging path for this layout).
     *
     * IPG_LAYOUT_JPEG (opt-in, new in ABI 3; SURVEY 8f-3, encode half): the result never leaves the device as pixels.
     * The engine runs the baseline JPEG writer of Go 1.24's image/jpeg on it -- rgbaToYCbCr (integer color.RGBToYCbCr,
     * edge pixels replicated), 4:2:0 chroma means, the jfdctint forward DCT of fdct.go, division by 8 * quant rounded
     * half away from zero, the Annex K Huffman tables, byte stuffing, and the writer's own header layout (SOI, one DQT,
     * SOF0, one DHT, SOS; no JFIF segment) -- all integer arithmetic, so the file is byte for byte what
     *   jpeg.Encode(buf, result, &jpeg.Options{Quality: jpeg_quality})
     * writes, and the host's encode step becomes SaveProcessed(bytes).  dst is then a byte buffer of dst_capacity
     * bytes in ipg_alloc_pinned memory (or device memory), dst_stride is ignored, and *dst_len -- also in
     * ipg_alloc_pinned memory, e.g. the last 8 bytes of the same allocation: both are written after ipg_submit returned,
     * and the library keeps no other caller pointer past the call -- receives the file length.  A file that does not fit dst_capacity fails
     * that ticket with IPG_ERR_NOMEM (w * h bytes is ample for photographs at quality 85; w * h * 3 + 4096 always fits
     * at the qualities the reference uses).  Images of 65536 pixels or more per side are refused as Go's writer refuses them. */
    int32_t dst_layout;
    void *dst_cb, *dst_cr;
    int32_t dst_cstride;
    int32_t jpeg_quality;     /* IPG_LAYOUT_JPEG: 1..100 as jpeg.Options.Quality (clamped like Go); 0 = 85, the reference's
                                 constant (domain/task.go:57) */
    uint64_t dst_capacity;    /* IPG_LAYOUT_JPEG: bytes available at dst */
    uint64_t *dst_len;        /* IPG_LAYOUT_JPEG: receives the file length (ipg_alloc_pinned memory, 8-byte aligned) */
} ipg_op;

typedef enum {
    /* IPG_OP_WATERMARK on an IPG_LAYOUT_RGBA8 source only.  draw.Draw(result, b, img, Point{}, Src) of an *image.RGBA is
     * a copy (watermark.go:91-92), so the watermarked image differs from its source only inside the union of the glyph
     * rectangles (~300 x 45 px for the default text).  With this flag the engine writes ONLY that box into dst and
     * leaves every other byte of dst untouched: the caller guarantees they already hold the source pixels -- e.g. dst IS
     * the source buffer (ops of one ticket all read the uploaded original, so aliasing is safe in any op order), or a
     * host copy of it.  Saves the full-frame device copy and (w*h*4 - box) bytes of D2H per image.  Ignored for other
     * source layouts (their draw.Draw is a conversion: the full frame is produced as usual). */
    IPG_OPF_WATERMARK_PATCH_ONLY = 1,
} ipg_op_flags;

typedef struct {
    uint64_t tickets_done;
    uint64_t batches;
    uint64_t kernels_launched;    /* kernel launches issued by this library   */
    uint64_t bytes_h2d, bytes_d2h;
    uint64_t exact_fixups;        /* output pixels re-evaluated in fp64       */
    uint64_t exact_fallbacks;     /* ops that ran whole-image in fp64         */
    uint64_t staged_copies;       /* host memcpy into/out of internal pinned staging */
    /* device time between CUDA events recorded on the launching stream, summed over batches */
    double kernel_ms;             /* all kernels                               */
    double stream_kernel_ms;      /* k_stream (fp32 resample + watermark copy) */
    double fix_kernel_ms;         /* k_exact_fix                               */
    double other_kernel_ms;       /* k_exact_tiles + k_watermark               */
    /* device-clock spans since ipg_reset_stats (max over devices): first kernel start ->
     * last kernel end, and first batch H2D start -> last batch D2H end */
    double kernel_span_ms;
    double batch_span_ms;
    /* of stream_kernel_ms: the lean single-target k_stream instantiation; and the stream jobs it ran */
    double stream_fast_kernel_ms;
    uint64_t fast_jobs;
    /* device time of the copies, summed over batches: first H2D start -> last H2D end on the upload stream, and first
     * D2H start -> last D2H end on the download stream (copies of different batches overlap each other and the kernels) */
    double h2d_ms, d2h_ms;
} ipg_stats;

/* ---- lifecycle ---------------------------------------------------------- */
/* device_ids==NULL or n<=0: use all visible devices. cfg may be NULL. */
IPG_API int ipg_init(const int *device_ids, int n, const ipg_config *cfg, ipg_ctx **out);
IPG_API void ipg_destroy(ipg_ctx *ctx);
IPG_API int ipg_device_count(const ipg_ctx *ctx);
IPG_API const char *ipg_last_error(void); /* thread-local, never NULL */
IPG_API int ipg_abi_version(void);

/* ---- memory the host may decode into / encode from ---------------------- */
IPG_API void *ipg_alloc_pinned(ipg_ctx *ctx, size_t bytes);
IPG_API void ipg_free_pinned(ipg_ctx *ctx, void *p);
/* device memory (for device-resident sources/destinations, IPG_MEM_DEVICE) */
IPG_API void *ipg_alloc_device(ipg_ctx *ctx, int device_index, size_t bytes);
IPG_API void ipg_free_device(ipg_ctx *ctx, int device_index, void *p);
IPG_API int ipg_copy_to_device(ipg_ctx *ctx, int device_index, void *dst, const void *src, size_t bytes);
IPG_API int ipg_copy_from_device(ipg_ctx *ctx, int device_index, void *dst, const void *src, size_t bytes);

/* ---- the hot path -------------------------------------------------------- */
/* Queue all ops of one decoded image (each op reads the ORIGINAL image,
 * image_processor.go:64-65).  The engine picks the device; tickets from
 * concurrent callers are coalesced into batched launches. */
IPG_API int ipg_submit(ipg_ctx *ctx, const ipg_image_desc *src, const ipg_op *ops, int n_ops,
               ipg_ticket *ticket);
/* Same, pinned to one device (required when any memspace is IPG_MEM_DEVICE). */
IPG_API int ipg_submit_on(ipg_ctx *ctx, int device_index, const ipg_image_desc *src, const ipg_op *ops,
                  int n_ops, ipg_ticket *ticket);
/* Block until the ticket finished (timeout_ms<0: forever). On IPG_OK the ticket
 * is consumed; on IPG_ERR_TIMEOUT it stays valid (ctx cancellation maps here);
 * any other code is that image's failure and consumes the ticket. */
IPG_API int ipg_wait(ipg_ctx *ctx, ipg_ticket ticket, int timeout_ms);
/* Wait for everything submitted so far on every device. */
IPG_API int ipg_flush(ipg_ctx *ctx);

/* ---- introspection ------------------------------------------------------- */
IPG_API int ipg_get_stats(ipg_ctx *ctx, ipg_stats *out);
/* ipg_flush, then zero the counters and restart the device-time spans. */
IPG_API int ipg_reset_stats(ipg_ctx *ctx);

/* ---- host-side helpers with the reference's arithmetic ------------------- */
/* resize.go:63-72 */
IPG_API void ipg_keep_aspect_dims(int ow, int oh, int w, int h, int *nw, int *nh);
/* thumbnail.go:52-63 */
IPG_API void ipg_thumb_fit_dims(int ow, int oh, int size, int *nw, int *nh);
/* thumbnail.go:115-127 */
IPG_API void ipg_crop_square(int ow, int oh, int *cx, int *cy, int *cs);

#ifdef __cplusplus
}
#endif
#endif /* IPGPU_H */
