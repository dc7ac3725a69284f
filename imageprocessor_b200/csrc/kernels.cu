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

// Launch-side state that CUDA keeps per device (function attributes, occupancy-derived grids) is cached per
// device: one process may drive every GPU of the box (ipg_init(NULL, 0, ...)).
enum { kMaxDevices = 64 };
static int current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

// ---------------------------------------------------------------------------------
// source adaptors (the inner expressions of x/image draw/impl.go scaleX_<type>)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ uint32_t div255(uint32_t x) { return __umulhi(x, 0x80808081u) >> 7; } // exact for every uint32

__device__ __forceinline__ size_t chroma_index(int layout, int s1, int x, int y)
{
    switch (layout) {
    case L_YCBCR444: return (size_t)y * s1 + x;
    case L_YCBCR422: return (size_t)y * s1 + (x >> 1);
    case L_YCBCR420: return (size_t)(y >> 1) * s1 + (x >> 1);
    default:         return (size_t)(y >> 1) * s1 + x; // 4:4:0
    }
}

__device__ __forceinline__ uint32_t be16(const uint8_t *p) { return ((uint32_t)__ldg(p) << 8) | (uint32_t)__ldg(p + 1); }
// Gray and YCbCr sources carry the literal constant 1.0 in the alpha lane of x/image's pass 1 (scaleX_Gray, scaleX_YCbCr*)
__device__ __forceinline__ bool layout_const_alpha(int layout) { return layout >= L_GRAY8 && layout <= L_YCBCR440; }

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
    // 16-bit types (x/image's generic / RGBA64Image path: RGBA64At per tap, alpha accumulated like a colour channel);
    // Pix is big-endian
    case L_RGBA64: {
        const uint8_t *q = s.p0 + (size_t)y * s.s0 + (size_t)x * 8;
        p[0] = be16(q); p[1] = be16(q + 2); p[2] = be16(q + 4); p[3] = be16(q + 6);
        return false;
    }
    case L_NRGBA64: { // color.NRGBA64.RGBA(): c = C * A / 0xffff in uint32
        const uint8_t *q = s.p0 + (size_t)y * s.s0 + (size_t)x * 8;
        const uint32_t a = be16(q + 6);
        p[0] = be16(q) * a / 0xffffu; p[1] = be16(q + 2) * a / 0xffffu; p[2] = be16(q + 4) * a / 0xffffu; p[3] = a;
        return false;
    }
    case L_GRAY16: {
        const uint32_t v = be16(s.p0 + (size_t)y * s.s0 + (size_t)x * 2);
        p[0] = p[1] = p[2] = v; p[3] = 0xffffu;
        return false;
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

// The fix-up list.  A flagged pixel costs ny x nx source samples (8 x 8 for the 4:1 resize, 29 x 29 for
// the 15:1 thumbnail) behind a chain of dependent global loads (entry -> job -> tap offsets -> weights
// and samples): walked naively it is pure latency.  So a warp takes 32 entries at a time:
//   header   everything the pixel needs from the entry, the job and the two axis tables;
//   narrow   supports up to 17 x 17: 32 entries per warp, one per lane -- their dependent header loads
//            overlap -- each finished by its own lane (exact_pixel_thread, k_exact_fix);
//   wide     queued by k_exact_fix, one pixel per warp in k_exact_fix_wide: the header is broadcast; ALL lanes stage the
//            two weight runs and the ny x nx block of 16-bit samples in shared memory (the loads of a
//            pixel are in flight together; rows in chunks if the block is large); then lanes take the
//            staged rows and run their horizontal sums (each sum sequential, unfused, in tap order)
//            and lanes 0..3 run the vertical sum of one channel each in row order.
// Same operations in the same order per value as exact_pixel.
enum { FIX_THREADS = 128, FIX_MAX_ROWS = 64, FIX_MAX_TAPS = 64, FIX_STAGE_PX = 1024, FIX_MLP = 16 };

struct FixWarpSmem {
    uint2 px[FIX_STAGE_PX];      // staged samples: (r16 | g16 << 16, b16 | a16 << 16)
    double tmp[FIX_MAX_ROWS][4]; // horizontally filtered rows
    double wx[FIX_MAX_TAPS];
    double wy[FIX_MAX_ROWS];
};

struct FixHdr {
    SrcView src;
    const double *wx, *wy;  // first weight of the pixel's taps on each axis
    uint8_t *dst;           // the pixel's 4 bytes
    double ifx, iy;
    int32_t x0, y0, nx, ny; // first tap (absolute source coordinates) and tap counts
    int32_t two_stage, job, ox, oy;
};

template <typename T> __device__ __forceinline__ T warp_bcast(const T &v, int src_lane)
{
    static_assert(sizeof(T) % 4 == 0, "word-sized POD");
    T r;
    const uint32_t *in = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *out = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) out[i] = __shfl_sync(0xffffffffu, in[i], src_lane);
    return r;
}

__device__ __forceinline__ FixHdr fix_header(const ExactJob *__restrict__ jobs, const FixEntry e)
{
    const ExactJob &J = jobs[e.job];
    FixHdr h;
    h.src = J.src;
    const int kx0 = __ldg(J.ax.off + e.x), ky0 = __ldg(J.ay.off + e.y);
    h.nx = __ldg(J.ax.off + e.x + 1) - kx0;
    h.ny = __ldg(J.ay.off + e.y + 1) - ky0;
    h.x0 = __ldg(J.ax.first + e.x) + J.rect_x;
    h.y0 = __ldg(J.ay.first + e.y) + J.rect_y;
    h.ifx = __ldg(J.ax.inv_ffff + e.x);
    h.iy = __ldg(J.ay.inv + e.y);
    h.wx = J.ax.w + kx0;
    h.wy = J.ay.w + ky0;
    h.dst = J.dst + (size_t)e.y * J.dst_stride + (size_t)e.x * 4;
    h.two_stage = J.two_stage;
    h.job = e.job; h.ox = e.x; h.oy = e.y;
    return h;
}

// One flagged pixel by ONE thread, from its header: for small supports (the 4:1 resize has 8 x 8 taps,
// 88 % of all fix-ups) 32 pixels per warp beat one pixel per warp by an order of magnitude in issued
// instructions.  Each row's samples are requested together (FIX_TMLP loads in flight per thread), then
// summed in tap order; the vertical sum runs row by row.  Operation order per value as exact_pixel.
enum { FIX_TMLP = 8, FIX_TROWS = 4, FIX_THREAD_TAPS = 17 };

// the x-sum of one staged row and its contribution to the vertical sums (reference order)
__device__ __forceinline__ void fix_row_rgba(const uint32_t *q, const double *w, int nx, bool two_stage, bool const_alpha, double ifx,
                                             double wy, double &pr, double &pg, double &pb, double &pa)
{
    double xr = 0, xg = 0, xb = 0, xa = 0;
#pragma unroll
    for (int u = 0; u < FIX_TMLP; u++) {
        if (u >= nx) break;
        uint32_t p[4] = {(q[u] & 0xff) * 0x101u, ((q[u] >> 8) & 0xff) * 0x101u, ((q[u] >> 16) & 0xff) * 0x101u, (q[u] >> 24) * 0x101u};
        if (two_stage) to_cropped_rgba16(p);
        xr = __dadd_rn(xr, __dmul_rn((double)p[0], w[u]));
        xg = __dadd_rn(xg, __dmul_rn((double)p[1], w[u]));
        xb = __dadd_rn(xb, __dmul_rn((double)p[2], w[u]));
        xa = __dadd_rn(xa, __dmul_rn((double)p[3], w[u]));
    }
    const double ta = const_alpha ? 1.0 : __dmul_rn(xa, ifx);
    pr = __dadd_rn(pr, __dmul_rn(__dmul_rn(xr, ifx), wy));
    pg = __dadd_rn(pg, __dmul_rn(__dmul_rn(xg, ifx), wy));
    pb = __dadd_rn(pb, __dmul_rn(__dmul_rn(xb, ifx), wy));
    pa = __dadd_rn(pa, __dmul_rn(ta, wy));
}

__device__ __forceinline__ void exact_pixel_thread(const FixHdr &h)
{
    const bool const_alpha = layout_const_alpha(h.src.layout) && !h.two_stage;
    double pr = 0, pg = 0, pb = 0, pa = 0;
    if (h.src.layout == L_RGBA8 && h.nx <= FIX_TMLP) {
        // the resize case: FIX_TROWS rows x nx taps requested together (every load of the pixel is 2-4 round
        // trips to DRAM instead of one per row), weights once
        double w[FIX_TMLP];
#pragma unroll
        for (int u = 0; u < FIX_TMLP; u++) w[u] = u < h.nx ? __ldg(h.wx + u) : 0.0;
        const uint8_t *base = h.src.p0 + (size_t)h.y0 * h.src.s0 + (size_t)h.x0 * 4;
        for (int j0 = 0; j0 < h.ny; j0 += FIX_TROWS) {
            uint32_t q[FIX_TROWS][FIX_TMLP];
            double wy[FIX_TROWS];
#pragma unroll
            for (int r = 0; r < FIX_TROWS; r++) {
                const bool rin = j0 + r < h.ny;
                const uint32_t *row = (const uint32_t *)(base + (size_t)(j0 + r) * h.src.s0);
                wy[r] = rin ? __ldg(h.wy + j0 + r) : 0.0;
#pragma unroll
                for (int u = 0; u < FIX_TMLP; u++) q[r][u] = (rin && u < h.nx) ? __ldg(row + u) : 0u;
            }
#pragma unroll
            for (int r = 0; r < FIX_TROWS; r++)
                if (j0 + r < h.ny) fix_row_rgba(q[r], w, h.nx, h.two_stage != 0, const_alpha, h.ifx, wy[r], pr, pg, pb, pa);
        }
    } else if (h.nx <= FIX_TMLP) {
        // the same for the other layouts (planar YCbCr, Gray, NRGBA): two rows of converted 16-bit samples at a time
        double w[FIX_TMLP];
#pragma unroll
        for (int u = 0; u < FIX_TMLP; u++) w[u] = u < h.nx ? __ldg(h.wx + u) : 0.0;
        for (int j0 = 0; j0 < h.ny; j0 += 2) {
            uint2 q[2][FIX_TMLP];
            double wy[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const bool rin = j0 + r < h.ny;
                wy[r] = rin ? __ldg(h.wy + j0 + r) : 0.0;
#pragma unroll
                for (int u = 0; u < FIX_TMLP; u++) {
                    uint32_t p[4] = {0u, 0u, 0u, 0u};
                    if (rin && u < h.nx) {
                        sample16(h.src, h.x0 + u, h.y0 + j0 + r, p);
                        if (h.two_stage) to_cropped_rgba16(p);
                    }
                    q[r][u] = make_uint2(p[0] | (p[1] << 16), p[2] | (p[3] << 16));
                }
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (j0 + r >= h.ny) break;
                double xr = 0, xg = 0, xb = 0, xa = 0;
#pragma unroll
                for (int u = 0; u < FIX_TMLP; u++) {
                    if (u >= h.nx) break;
                    xr = __dadd_rn(xr, __dmul_rn((double)(q[r][u].x & 0xffffu), w[u]));
                    xg = __dadd_rn(xg, __dmul_rn((double)(q[r][u].x >> 16), w[u]));
                    xb = __dadd_rn(xb, __dmul_rn((double)(q[r][u].y & 0xffffu), w[u]));
                    xa = __dadd_rn(xa, __dmul_rn((double)(q[r][u].y >> 16), w[u]));
                }
                const double ta = const_alpha ? 1.0 : __dmul_rn(xa, h.ifx);
                pr = __dadd_rn(pr, __dmul_rn(__dmul_rn(xr, h.ifx), wy[r]));
                pg = __dadd_rn(pg, __dmul_rn(__dmul_rn(xg, h.ifx), wy[r]));
                pb = __dadd_rn(pb, __dmul_rn(__dmul_rn(xb, h.ifx), wy[r]));
                pa = __dadd_rn(pa, __dmul_rn(ta, wy[r]));
            }
        }
    } else
    for (int j = 0; j < h.ny; j++) {
        double xr = 0, xg = 0, xb = 0, xa = 0;
        if (h.src.layout == L_RGBA8) {
            const uint32_t *row = (const uint32_t *)(h.src.p0 + (size_t)(h.y0 + j) * h.src.s0) + h.x0;
            for (int k0 = 0; k0 < h.nx; k0 += FIX_TMLP) {
                uint32_t q[FIX_TMLP];
                double w[FIX_TMLP];
#pragma unroll
                for (int u = 0; u < FIX_TMLP; u++) {
                    const bool in = k0 + u < h.nx;
                    q[u] = in ? __ldg(row + k0 + u) : 0u;
                    w[u] = in ? __ldg(h.wx + k0 + u) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < FIX_TMLP; u++) {
                    if (k0 + u >= h.nx) break;
                    uint32_t p[4] = {(q[u] & 0xff) * 0x101u, ((q[u] >> 8) & 0xff) * 0x101u, ((q[u] >> 16) & 0xff) * 0x101u,
                                     (q[u] >> 24) * 0x101u};
                    if (h.two_stage) to_cropped_rgba16(p);
                    xr = __dadd_rn(xr, __dmul_rn((double)p[0], w[u]));
                    xg = __dadd_rn(xg, __dmul_rn((double)p[1], w[u]));
                    xb = __dadd_rn(xb, __dmul_rn((double)p[2], w[u]));
                    xa = __dadd_rn(xa, __dmul_rn((double)p[3], w[u]));
                }
            }
        } else {
            for (int k = 0; k < h.nx; k++) {
                uint32_t p[4];
                sample16(h.src, h.x0 + k, h.y0 + j, p);
                if (h.two_stage) to_cropped_rgba16(p);
                const double w = __ldg(h.wx + k);
                xr = __dadd_rn(xr, __dmul_rn((double)p[0], w));
                xg = __dadd_rn(xg, __dmul_rn((double)p[1], w));
                xb = __dadd_rn(xb, __dmul_rn((double)p[2], w));
                xa = __dadd_rn(xa, __dmul_rn((double)p[3], w));
            }
        }
        const double wy = __ldg(h.wy + j);
        const double ta = const_alpha ? 1.0 : __dmul_rn(xa, h.ifx);
        pr = __dadd_rn(pr, __dmul_rn(__dmul_rn(xr, h.ifx), wy));
        pg = __dadd_rn(pg, __dmul_rn(__dmul_rn(xg, h.ifx), wy));
        pb = __dadd_rn(pb, __dmul_rn(__dmul_rn(xb, h.ifx), wy));
        pa = __dadd_rn(pa, __dmul_rn(ta, wy));
    }
    if (pr > pa) pr = pa;
    if (pg > pa) pg = pa;
    if (pb > pa) pb = pa;
    uchar4 o;
    o.x = (unsigned char)(ftou(__dmul_rn(pr, h.iy)) >> 8);
    o.y = (unsigned char)(ftou(__dmul_rn(pg, h.iy)) >> 8);
    o.z = (unsigned char)(ftou(__dmul_rn(pb, h.iy)) >> 8);
    o.w = (unsigned char)(ftou(__dmul_rn(pa, h.iy)) >> 8);
    *(uchar4 *)h.dst = o;
}

__device__ void exact_pixel_warp(const ExactJob *__restrict__ jobs, const FixHdr &h, FixWarpSmem &S)
{
    const int lane = threadIdx.x & 31;
    const int nx = h.nx, ny = h.ny;
    if (ny > FIX_MAX_ROWS || nx > FIX_MAX_TAPS || nx < 1) { // cannot stage: one lane does it alone
        if (lane == 0) *(uchar4 *)h.dst = exact_pixel(jobs[h.job], h.ox, h.oy);
        __syncwarp();
        return;
    }
    const bool const_alpha = layout_const_alpha(h.src.layout) && !h.two_stage; // Gray / YCbCr: alpha is the literal 1.0
    for (int k = lane; k < nx; k += 32) S.wx[k] = __ldg(h.wx + k);
    for (int j = lane; j < ny; j += 32) S.wy[j] = __ldg(h.wy + j);
    // staging index: lanes walk (row, tap) with the tap count rounded up to a power of two (no division)
    const int lg = 32 - __clz(nx - 1), nxp = 1 << lg; // nx = 1 -> lg 0
    const int rows_per_chunk = FIX_STAGE_PX / nx;     // >= 16
    const bool rgba = h.src.layout == L_RGBA8;
    for (int j0 = 0; j0 < ny; j0 += rows_per_chunk) {
        const int R = min(rows_per_chunk, ny - j0), n = R << lg;
        if (rgba) { // the common case without the per-sample layout switch: 8-bit premultiplied, 16 bits = byte * 0x101
            const uint8_t *base = h.src.p0 + (size_t)(h.y0 + j0) * h.src.s0 + (size_t)h.x0 * 4;
            for (int i0 = lane; i0 < n; i0 += 32 * FIX_MLP) { // FIX_MLP loads per lane in flight, then their unpacking
                uint32_t q[FIX_MLP];
#pragma unroll
                for (int u = 0; u < FIX_MLP; u++) {
                    const int i = i0 + 32 * u, jj = i >> lg, k = i & (nxp - 1);
                    q[u] = (i < n && k < nx) ? __ldg((const uint32_t *)(base + (size_t)jj * h.src.s0) + k) : 0u;
                }
#pragma unroll
                for (int u = 0; u < FIX_MLP; u++) {
                    const int i = i0 + 32 * u, jj = i >> lg, k = i & (nxp - 1);
                    if (i >= n || k >= nx) continue;
                    uint32_t p[4] = {(q[u] & 0xff) * 0x101u, ((q[u] >> 8) & 0xff) * 0x101u, ((q[u] >> 16) & 0xff) * 0x101u,
                                     (q[u] >> 24) * 0x101u};
                    if (h.two_stage) to_cropped_rgba16(p);
                    S.px[jj * nx + k] = make_uint2(p[0] | (p[1] << 16), p[2] | (p[3] << 16));
                }
            }
        } else {
#pragma unroll 4
            for (int i = lane; i < n; i += 32) {
                const int jj = i >> lg, k = i & (nxp - 1);
                if (k >= nx) continue;
                uint32_t p[4];
                sample16(h.src, h.x0 + k, h.y0 + j0 + jj, p);
                if (h.two_stage) to_cropped_rgba16(p);
                S.px[jj * nx + k] = make_uint2(p[0] | (p[1] << 16), p[2] | (p[3] << 16));
            }
        }
        __syncwarp();
        for (int jj = lane; jj < R; jj += 32) {
            const uint2 *row = S.px + jj * nx;
            double xr = 0, xg = 0, xb = 0, xa = 0;
            for (int k = 0; k < nx; k++) {
                const uint2 q = row[k];
                const double w = S.wx[k];
                xr = __dadd_rn(xr, __dmul_rn((double)(q.x & 0xffffu), w));
                xg = __dadd_rn(xg, __dmul_rn((double)(q.x >> 16), w));
                xb = __dadd_rn(xb, __dmul_rn((double)(q.y & 0xffffu), w));
                xa = __dadd_rn(xa, __dmul_rn((double)(q.y >> 16), w));
            }
            double *t = S.tmp[j0 + jj];
            t[0] = __dmul_rn(xr, h.ifx);
            t[1] = __dmul_rn(xg, h.ifx);
            t[2] = __dmul_rn(xb, h.ifx);
            t[3] = const_alpha ? 1.0 : __dmul_rn(xa, h.ifx);
        }
        __syncwarp();
    }
    double p = 0;
    if (lane < 4) {
        for (int j = 0; j < ny; j++) p = __dadd_rn(p, __dmul_rn(S.tmp[j][lane], S.wy[j]));
    }
    const double pa = __shfl_sync(0xffffffffu, p, 3);
    if (lane < 4) {
        if (p > pa) p = pa;
        h.dst[lane] = (unsigned char)(ftou(__dmul_rn(p, h.iy)) >> 8);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FIX_THREADS)
k_exact_fix(const ExactJob *__restrict__ jobs, int n_jobs, FixList fix)
{
    __shared__ FixWarpSmem ws[FIX_THREADS / 32];
    const uint32_t cnt = fix.count[0];
    if (cnt == 0) return;
    if (cnt <= fix.capacity) {
        // Warps claim batches of up to 32 entries (one per lane) from the cursor fix.count[1].  Batch b takes the
        // entries b, b + nbatch, b + 2 nbatch, ... (list neighbours -- the flagged pixels of one emitted row --
        // land in different batches).  A lane finishes its own pixel if the support is narrow; wide ones are
        // queued at the unused back end of the list (fix.count[2]) for k_exact_fix_wide, one warp each.
        const int lane = threadIdx.x & 31;
        const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
        const uint32_t width = min(32u, (cnt + nwarps - 1) / nwarps); // entries per batch: one batch per warp while they last
        const uint32_t nbatch = (cnt + width - 1) / width;
        for (;;) {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(fix.count + 1, 1u);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= nbatch) break;
            const uint32_t idx = b + (uint32_t)lane * nbatch;
            const bool valid = (uint32_t)lane < width && idx < cnt;
            FixHdr mine = {};
            if (valid) mine = fix_header(jobs, fix.entries[idx]);
            const bool small = valid && mine.nx <= FIX_THREAD_TAPS && mine.ny <= FIX_THREAD_TAPS;
            bool inline_wide = false;
            if (small) {
                exact_pixel_thread(mine);
            } else if (valid) {
                const uint32_t pos = atomicAdd(fix.count + 2, 1u);
                if (cnt + pos < fix.capacity) fix.entries[fix.capacity - 1u - pos] = FixEntry{mine.job, mine.ox, mine.oy};
                else inline_wide = true; // no room left at the back: the warp does it here
            }
            uint32_t big = __ballot_sync(0xffffffffu, inline_wide);
            while (big) {
                const int k = __ffs(big) - 1;
                big &= big - 1;
                exact_pixel_warp(jobs, warp_bcast(mine, k), ws[threadIdx.x >> 5]);
            }
        }
    } else {
        // the list overflowed: entries were dropped, so redo every stream target whole
        const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
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

// The wide supports queued by k_exact_fix at the back of the list: one pixel per warp and claim (cursor
// fix.count[3]), so a 29 x 29-tap thumbnail pixel never waits behind another one in the same warp.
__global__ void __launch_bounds__(FIX_THREADS)
k_exact_fix_wide(const ExactJob *__restrict__ jobs, FixList fix)
{
    __shared__ FixWarpSmem ws[FIX_THREADS / 32];
    const uint32_t cnt = fix.count[0];
    if (cnt > fix.capacity) return; // overflow: k_exact_fix redid every target whole
    const uint32_t n_wide = min(fix.count[2], fix.capacity - cnt);
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(fix.count + 3, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n_wide) break;
        FixHdr h = {};
        if (lane == 0) h = fix_header(jobs, fix.entries[fix.capacity - 1u - i]);
        exact_pixel_warp(jobs, warp_bcast(h, 0), ws[threadIdx.x >> 5]);
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
    case L_RGBA64: case L_NRGBA64: case L_GRAY16: { // no fast path in image/draw: drawRGBA's generic loop, uint8(At().RGBA() >> 8)
        uint32_t p[4];
        sample16(s, x, y, p);
        return (p[0] >> 8) | ((p[1] >> 8) << 8) | ((p[2] >> 8) << 16) | ((p[3] >> 8) << 24);
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

// 8-bit color.YCbCrToRGB of one pixel from the chroma terms of its chroma sample (imageutil.DrawYCbCr's inner expression)
__device__ __forceinline__ uint32_t ycc_px8(uint32_t y, int tr, int tg, int tb)
{
    const int yy1 = (int)y * 0x10101;
    const uint32_t r = (uint32_t)min(max((yy1 + tr) >> 16, 0), 0xff);
    const uint32_t g = (uint32_t)min(max((yy1 + tg) >> 16, 0), 0xff);
    const uint32_t b = (uint32_t)min(max((yy1 + tb) >> 16, 0), 0xff);
    return r | (g << 8) | (b << 16) | 0xff000000u;
}

// Full-frame draw.Draw(Src) conversion.  A thread converts 8 consecutive pixels of a row per step from vector loads
// (8 luma bytes + 4 or 8 bytes of each chroma plane, or two uint4 of NRGBA / one uint2 of Gray) into two 16-byte
// stores; rows whose planes are not aligned for that, and the ragged right edge, take the per-pixel expression.
__global__ void __launch_bounds__(256)
k_watermark(const WmJob *__restrict__ jobs, const WmItem *__restrict__ items)
{
    const WmItem it = items[blockIdx.x];
    const WmJob &J = jobs[it.job];
    const SrcView s = J.src;
    const int W = s.w;
    const int y1 = min(it.row0 + WM_ROWS, s.h);
    const int layout = s.layout;
    const bool ycc = layout >= L_YCBCR444 && layout <= L_YCBCR440;
    const bool sub_x = layout == L_YCBCR422 || layout == L_YCBCR420;
    const bool dst_vec = ((J.wm.dst_stride | (int)(size_t)J.wm.dst) & 15) == 0;
    bool vec = dst_vec && layout <= L_YCBCR440; // the 16-bit types take the per-pixel expression
    if (layout == L_RGBA8 || layout == L_NRGBA8) vec = vec && (((size_t)s.p0 | (size_t)s.s0) & 15) == 0;
    else vec = vec && (((size_t)s.p0 | (size_t)s.s0) & 7) == 0;
    if (ycc) vec = vec && (((size_t)s.p1 | (size_t)s.p2 | (size_t)s.s1 | (size_t)s.s2) & (sub_x ? 3 : 7)) == 0;
    const int Wv = vec ? (W & ~7) : 0; // columns converted 8 at a time
    for (int y = it.row0; y < y1; y++) {
        uint8_t *drow = J.wm.dst + (size_t)y * J.wm.dst_stride;
        for (int x = threadIdx.x * 8; x < Wv; x += 256 * 8) {
            uint32_t o[8];
            if (layout == L_RGBA8) {
                const uint4 *q = (const uint4 *)(s.p0 + (size_t)y * s.s0 + (size_t)x * 4);
                const uint4 a = __ldg(q), b = __ldg(q + 1);
                o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
            } else if (layout == L_NRGBA8) { // drawNRGBASrc: sa = A * 0x101; c = uint8((C * sa / 0xff) >> 8); a = uint8(sa >> 8)
                const uint4 *q = (const uint4 *)(s.p0 + (size_t)y * s.s0 + (size_t)x * 4);
                const uint4 a = __ldg(q), b = __ldg(q + 1);
                const uint32_t in[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t sa = (in[j] >> 24) * 0x101u;
                    const uint32_t r = div255((in[j] & 0xff) * sa) >> 8, g = div255(((in[j] >> 8) & 0xff) * sa) >> 8,
                                   bb = div255(((in[j] >> 16) & 0xff) * sa) >> 8;
                    o[j] = r | (g << 8) | (bb << 16) | (in[j] & 0xff000000u);
                }
            } else if (layout == L_GRAY8) {
                const uint2 g = __ldg((const uint2 *)(s.p0 + (size_t)y * s.s0 + x));
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = (((j < 4 ? g.x : g.y) >> (8 * (j & 3))) & 0xff) * 0x010101u | 0xff000000u;
            } else {
                const uint2 yy = __ldg((const uint2 *)(s.p0 + (size_t)y * s.s0 + x));
                const int cy = (layout == L_YCBCR420 || layout == L_YCBCR440) ? y >> 1 : y;
                uint2 cb, cr; // chroma samples under these 8 pixels: 4 (in .x) when subsampled horizontally, else 8
                if (sub_x) {
                    cb = make_uint2(__ldg((const uint32_t *)(s.p1 + (size_t)cy * s.s1 + (x >> 1))), 0u);
                    cr = make_uint2(__ldg((const uint32_t *)(s.p2 + (size_t)cy * s.s2 + (x >> 1))), 0u);
                } else {
                    cb = __ldg((const uint2 *)(s.p1 + (size_t)cy * s.s1 + x));
                    cr = __ldg((const uint2 *)(s.p2 + (size_t)cy * s.s2 + x));
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int c = sub_x ? j >> 1 : j;
                    const int cb1 = (int)(((c < 4 ? cb.x : cb.y) >> (8 * (c & 3))) & 0xff) - 128;
                    const int cr1 = (int)(((c < 4 ? cr.x : cr.y) >> (8 * (c & 3))) & 0xff) - 128;
                    o[j] = ycc_px8(((j < 4 ? yy.x : yy.y) >> (8 * (j & 3))) & 0xff, 91881 * cr1, -22554 * cb1 - 46802 * cr1, 116130 * cb1);
                }
            }
            uint4 *d = (uint4 *)(drow + (size_t)x * 4);
            d[0] = make_uint4(o[0], o[1], o[2], o[3]);
            d[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        for (int x = Wv + threadIdx.x; x < W; x += 256) // unaligned planes and the ragged right edge
            *(uint32_t *)(drow + (size_t)x * 4) = draw_src_px(s, x, y);
    }
}

// Ordered glyph blend, in place on the watermark destination, over the union box of
// the glyph rectangles only (~300x45 px): one thread per pixel walks the glyphs in string
// order, exactly like freetype's one-DrawMask-per-rune sequence.  Runs after the kernel
// that produced the full-frame copy/conversion, on the same stream.
__global__ void __launch_bounds__(256)
k_blend(const WatermarkD *__restrict__ wms, const BlendItem *__restrict__ items)
{
    const BlendItem it = items[blockIdx.x];
    const WatermarkD &wm = wms[it.wm];
    const int x = wm.bx0 + it.tile_x * 32 + (threadIdx.x & 31);
    const int y = wm.by0 + it.tile_y * 8 + (threadIdx.x >> 5);
    if (x >= wm.bx1 || y >= wm.by1) return;
    uint32_t *p = (uint32_t *)(wm.dst + (size_t)y * wm.dst_stride) + x;
    const uint32_t d = *p;
    const uint32_t o = glyph_over_px(d, x, y, wm);
    if (o != d) *p = o;
}

// ---------------------------------------------------------------------------------
// k_rgba_to_ycbcr420: the colour conversion and chroma down-sampling Go's image/jpeg writer applies to an *image.RGBA
// before its DCT (Go 1.24 image/jpeg/writer.go rgbaToYCbCr + scale; image/color/ycbcr.go RGBToYCbCr), so that the host
// can hand jpeg.Encode a 4:2:0 *image.YCbCr and get the bytes it would have produced from the RGBA result.
//   per pixel  yy = (19595 R + 38470 G + 7471 B + 1<<15) >> 16
//              cb = -11056 R - 21712 G + 32768 B + 257<<15;  cb = cb fits 24 bits ? cb >> 16 : (cb < 0 ? 0 : 255)   (same for cr)
//   per 2 x 2  c = (c00 + c01 + c10 + c11 + 2) >> 2, coordinates past the last column / row clamped (the writer replicates
//              the edge pixel inside its 16 x 16 blocks; with clamped reads a 4:2:0 image reproduces exactly that)
// A thread converts 2 rows x 8 pixels: 16 luma bytes and 4 + 4 chroma bytes.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void rgb_to_ycbcr(uint32_t q, int &yy, int &cb, int &cr)
{
    const int r = (int)(q & 0xff), g = (int)((q >> 8) & 0xff), b = (int)((q >> 16) & 0xff);
    yy = (19595 * r + 38470 * g + 7471 * b + (1 << 15)) >> 16;
    int c = -11056 * r - 21712 * g + 32768 * b + (257 << 15);
    cb = (((uint32_t)c & 0xff000000u) == 0 ? c >> 16 : ~(c >> 31)) & 0xff;
    c = 32768 * r - 27440 * g - 5328 * b + (257 << 15);
    cr = (((uint32_t)c & 0xff000000u) == 0 ? c >> 16 : ~(c >> 31)) & 0xff;
}

__global__ void __launch_bounds__(256)
k_rgba_to_ycbcr420(const YccJob *__restrict__ jobs, const YccItem *__restrict__ items)
{
    const YccItem it = items[blockIdx.x];
    const YccJob &J = jobs[it.job];
    const int bx = it.tile_x * 32 + (threadIdx.x & 31), by = it.tile_y * 8 + (threadIdx.x >> 5); // 8-pixel, 2-row block
    const int x0 = bx * 8, y0 = by * 2;
    if (x0 >= J.w || y0 >= J.h) return;
    uint32_t yv[2][2] = {{0u, 0u}, {0u, 0u}}; // 8 luma bytes per row, packed
    int sb[4] = {0, 0, 0, 0}, sr[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t *row = (const uint32_t *)(J.rgba + (size_t)min(y0 + j, J.h - 1) * J.rgba_pitch);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int yy, cb, cr;
            rgb_to_ycbcr(__ldg(row + min(x0 + i, J.w - 1)), yy, cb, cr);
            yv[j][i >> 2] |= (uint32_t)yy << (8 * (i & 3));
            sb[i >> 1] += cb;
            sr[i >> 1] += cr;
        }
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        if (y0 + j >= J.h) break;
        uint8_t *d = J.y + (size_t)(y0 + j) * J.y_pitch + x0;
        if (x0 + 8 <= J.w && (((size_t)d) & 7) == 0) *(uint2 *)d = make_uint2(yv[j][0], yv[j][1]);
        else
            for (int i = 0; i < 8 && x0 + i < J.w; i++) d[i] = (uint8_t)(yv[j][i >> 2] >> (8 * (i & 3)));
    }
    const int cw = (J.w + 1) >> 1, cx0 = bx * 4;
    uint32_t pb = 0, pr = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        pb |= (uint32_t)((sb[c] + 2) >> 2) << (8 * c);
        pr |= (uint32_t)((sr[c] + 2) >> 2) << (8 * c);
    }
    uint8_t *db = J.cb + (size_t)by * J.c_pitch + cx0, *dr = J.cr + (size_t)by * J.c_pitch + cx0;
    if (cx0 + 4 <= cw && ((((size_t)db) | ((size_t)dr)) & 3) == 0) {
        *(uint32_t *)db = pb;
        *(uint32_t *)dr = pr;
    } else {
        for (int c = 0; c < 4 && cx0 + c < cw; c++) { db[c] = (uint8_t)(pb >> (8 * c)); dr[c] = (uint8_t)(pr >> (8 * c)); }
    }
}

// The same blend for the patch-only watermark of an *image.RGBA source: the box pixels come from the SOURCE (draw.Draw
// Src of an RGBA is a copy), are blended in string order and land in the patch buffer / the caller's frame.  Every box
// pixel is written (blended or not): the patch is copied over the caller's buffer as a rectangle.
__global__ void __launch_bounds__(256)
k_blend_patch(const PatchJob *__restrict__ jobs, const BlendItem *__restrict__ items)
{
    const BlendItem it = items[blockIdx.x];
    const PatchJob &J = jobs[it.wm];
    const WatermarkD &wm = J.wm;
    const int x = wm.bx0 + it.tile_x * 32 + (threadIdx.x & 31);
    const int y = wm.by0 + it.tile_y * 8 + (threadIdx.x >> 5);
    if (x >= wm.bx1 || y >= wm.by1) return;
    const uint32_t d = __ldg((const uint32_t *)(J.src.p0 + (size_t)y * J.src.s0) + x);
    uint32_t *p = (uint32_t *)(wm.dst + (size_t)(y - J.oy) * wm.dst_stride) + (x - J.ox);
    *p = glyph_over_px(d, x, y, wm);
}

// ---------------------------------------------------------------------------------
// k_stream: vertical-first fp32 streaming resample
//
// CTA = 4 V warps + 1 producer warp.  Each V warp covers 128 source columns (32 lanes x
// 4 px); consecutive warps start `warp_stride` columns apart, so the CTA's slab is
// 3 * warp_stride + 128 columns wide and the CTA owns `tile_w` of them (the rest is the
// right halo the widest horizontal support needs).  The CTA walks the source rows of its
// band once, top to bottom, in groups of STREAM_GROUP rows:
//
//   producer   one elected lane drives a ring of STAGES stages with TMA bulk copies
//              (cp.async.bulk, SASS UBLKCP): per group, the rows and the group's records
//              land in one stage under one mbarrier phase.  For the watermark copy the
//              same lane bulk-STORES the owned columns of the landed rows straight from
//              the ring to the destination (draw.Draw(Src) of an *image.RGBA is a copy):
//              the copy costs no SM instructions and the source crosses HBM once for
//              resize + thumbnail + watermark.  A stage is refilled once the V warps
//              released it (`empty` mbarrier) and its store has read it (bulk-group wait).
//   V warps    wait for a stage, LDS.128 their 4 pixels of each of the 4 rows, convert
//              bytes to fp32 once (PRMT into 2^23+b, FADD2: the I2F.U8 the compiler would
//              pick runs on the quarter-rate XU pipe) and FFMA2 them into the two
//              accumulator sets of each target: a tent of half-width `scale` centred every
//              `scale` rows covers each source row exactly twice.  Weights arrive per SET
//              (the host resolved which set is which open output row).  While a warp has
//              only met opaque pixels its alpha sums are the host-computed chains in the
//              records, so the fast path carries no alpha arithmetic.
//   emit       when a source row completes an output row of a LOCAL target (narrow
//              horizontal support: the resize) each V warp parks its 128 filtered columns
//              in its private shared-memory strip and, after a __syncwarp, every lane
//              gathers one output pixel (taps in registers), quantises with the
//              reference's ftou()>>8, flags bytes too close to a quantiser step for the
//              fp64 fix-up and stores uchar4: no cross-warp hand-off at all.  For a SHARED
//              target (wide support: the 15:1 thumbnail, one emit per ~15 rows) the four
//              warps park into a CTA-wide double-buffered row, meet at a 128-thread named
//              barrier, and split the outputs among all 128 threads.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> 3) & 7); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends the thread in hardware up to the hint (ns) and wakes on completion
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
// ... on a precomputed 32-bit shared address.  smem_u32() of a ring barrier costs an S2R (SR_CgaCtaId: the address carries
// the CTA's rank in its cluster window) wherever the compiler rematerialises it -- once per group in the V loops -- so
// the V warps compute the ring's barrier addresses once and keep them.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void vwarps_bar() { asm volatile("bar.sync 1, %0;" ::"r"((int)STREAM_THREADS) : "memory"); }

__device__ __forceinline__ uint32_t clamp_to_alpha(uint32_t q)
{
    uint32_t a = q >> 24;
    uint32_t r = min(q & 0xff, a), g = min((q >> 8) & 0xff, a), b = min((q >> 16) & 0xff, a);
    return r | (g << 8) | (b << 16) | (a << 24);
}

__device__ __forceinline__ uint32_t quant16(float v, uint32_t D, uint32_t span, bool &amb)
{
    // t = v + 0.5 in 16.8 fixed point (T < 2^24 because v <= 65535): the output byte is
    // floor(t) >> 8 = T >> 16.  The byte is ambiguous when T lies within D of a multiple
    // of 65536: (low16 - D) mod 2^32 >= 65536 - 2D  (span = 65536 - 2D).
    const uint32_t T = (uint32_t)min(__float2int_rd(fmaf(v, 256.0f, 128.0f)), 0xffffff);
    amb |= ((T & 0xffffu) - D) >= span;
    return T >> 16;
}

struct __align__(16) StreamStage {
    uint4 rows[STREAM_GROUP][STREAM_THREADS]; // source rows of the slab (slab_cols * 4 bytes used)
    GroupRec rec[2];                          // this group's records, one per target
};

// Per-thread horizontal-pass table of a target in "cached" form (StreamTarget::tile_parts):
// every output is split over P adjacent V threads, each owning at most STREAM_XTAPS taps
// e0, e0 + P, e0 + 2P, ...  Lives in shared memory, not registers: the emit code exists
// once for all targets and the accumulators keep the register file.
template <int NT> struct XTab {
    int16_t ox[NT][STREAM_XROUNDS][STREAM_THREADS]; // output column of each round (< 32768), or -1 when the thread has none
    int16_t e0[NT][STREAM_XROUNDS][STREAM_THREADS]; // element of xbuf[T] holding its first tap
    float w[NT][STREAM_XTAPS_TAB][STREAM_THREADS];  // its weights (round r: entries r * ntap ...), 0 past the end
};
template <> struct XTab<0> {};

#ifndef IPG_ROW_UNROLL
#define IPG_ROW_UNROLL 4
#endif
constexpr int kRowUnroll = IPG_ROW_UNROLL; // rows of a group unrolled in the V loop
#ifndef IPG_FAST_UNROLL
#define IPG_FAST_UNROLL 4
#endif
constexpr int kFastUnroll = IPG_FAST_UNROLL; // ... in the lean instantiations (each unrolled row carries its own copy of the emit code)
enum { STREAM_XBUF = STREAM_COLS + 64 }; // padded: zero-weight taps past an output's support read finite data

// What the horizontal pass needs of a target, copied to shared memory once per CTA.
struct XInfo {
    uint8_t *dst;
    int32_t dst_stride, exact_job;
    uint32_t D;
    int32_t local, parts, ntap, rounds; // parts = threads per output (0: generic form), ntap = taps per thread and round
};

template <int NT, int STAGES> struct __align__(128) StreamSmem {
    StreamStage stage[STAGES];
    float4 xbuf[NT > 0 ? NT : 1][STREAM_XBUF];   // vertically filtered row of each target, XOR-swizzled:
                                                 // local target: four private 128-column warp strips;
                                                 // shared target: the CTA's slab
    XTab<NT> xt;
    uint64_t full[STAGES];                       // TMA completion of a stage
    uint64_t empty[STAGES];                      // the 4 V warps are done with a stage
    XInfo xi[2];                                 // per target; parts = tile_parts of this CTA's tile
    // what a refill / watermark store needs, once per CTA (the lanes that drive the ring read it from here)
    const uint8_t *ring_src, *ring_rec;          // first source row of the band at the tile's first column; the band's records
    uint8_t *ring_wm;                            // watermark destination of that row and column
    int32_t ring_wm_stride;
    uint32_t ring_rec_bytes;
    uint32_t done[STAGES];                       // no-producer-warp form: V warps done with the stage (the 4th one refills it)
};
#ifndef IPG_INLINE_PRODUCER
#define IPG_INLINE_PRODUCER 0
#endif
template <int NT, int LEAN = 0> struct StreamCfg {
    static constexpr bool FAST = LEAN != 0;
    // IPG_INLINE_PRODUCER=1 builds the lean single-target instantiations without a producer warp: lane 0 of V warp
    // (g mod 4) stores group g when it lands and the last V warp to finish a group refills its stage, so a CTA is 4
    // warps and 5 CTAs fit an SM at the same 96 registers (20 V warps per SM instead of 16, 3 ring stages).
    // Parity-green, but measured SLOWER (17.8 / 32.2 us per 12 MP image for resize / all three ops against
    // 15.1 / 24.6 with the producer warp): the refill then waits for the slowest warp's lane to get to it and
    // stalls that warp, and the ring is one stage shallower.  Kept as a tested option, off by default.
    static constexpr bool INLINE_PROD = IPG_INLINE_PRODUCER != 0 && (LEAN == 1 || LEAN == 2 || LEAN == 4);
    static constexpr int THREADS = INLINE_PROD ? (int)STREAM_THREADS : (int)STREAM_CTA;
    static constexpr int STAGES = INLINE_PROD ? STREAM_STAGES_INL : LEAN == 3 ? STREAM_STAGES_FAST2 : FAST ? STREAM_STAGES_FAST : NT == 2 ? STREAM_STAGES_2T : STREAM_STAGES_1T;
    static constexpr int CTAS_PER_SM = INLINE_PROD ? STREAM_CTAS_INL : LEAN == 3 ? STREAM_CTAS_FAST2 : FAST ? STREAM_CTAS_FAST : NT == 2 ? STREAM_CTAS_2T : STREAM_CTAS_1T;
    using Smem = StreamSmem<NT, STAGES>;
};

// ---- horizontal pass ---------------------------------------------------------------
#ifndef IPG_LEAN_OPAQUE_FINISH
#define IPG_LEAN_OPAQUE_FINISH 1
#endif
// OPAQUE (the lean instantiations: every pixel met so far had alpha 255, or the job is redone): the alpha sum is
// 65535 * (weights summing to 1) -- T = 0xffff80 + a few units, byte 255, 128 units from a quantiser step, never
// ambiguous while D < 128 -- so the byte is written as a constant, and the premultiplied clamp min(c, a) can only
// touch values within rounding of 65535 whose byte is 255 either way (a flag it might add or drop there is harmless:
// the fp64 re-evaluation returns 255 too).
template <bool OPAQUE = false, typename TI>
__device__ __forceinline__ void xfinish(const TI &t, uint32_t D, int ox, int oy, float2 rg, float2 ba, const FixList &fix)
{
    bool amb = false;
    const uint32_t span = 65536u - 2u * D;
    uint32_t o;
    if constexpr (OPAQUE && IPG_LEAN_OPAQUE_FINISH) {
        o = quant16(rg.x, D, span, amb) | (quant16(rg.y, D, span, amb) << 8) | (quant16(ba.x, D, span, amb) << 16) | 0xff000000u;
    } else {
        const float a = ba.y;
        const float r = fminf(rg.x, a), g = fminf(rg.y, a), b = fminf(ba.x, a);
        o = quant16(r, D, span, amb) | (quant16(g, D, span, amb) << 8) | (quant16(b, D, span, amb) << 16) | (quant16(a, D, span, amb) << 24);
    }
    *(uint32_t *)(t.dst + (size_t)oy * t.dst_stride + (size_t)ox * 4) = o;
    if (amb && fix.capacity) { // (keep this shape: the compiler aggregates the atomic per warp; other forms cost the V loop 5-30 %)
        const uint32_t idx = atomicAdd(fix.count, 1u);
        if (idx < fix.capacity) fix.entries[idx] = FixEntry{t.exact_job, ox, oy};
    }
}

// One output pixel: n taps of a parked row starting at element e0, weights from L1.
// FFMA2 carries the (r,g) and (b,a) pairs; taps in order, exactly as the emulator.
__device__ __forceinline__ void xgather(const float *__restrict__ wp, int e0, int n, const float4 *__restrict__ buf,
                                        float2 &rg, float2 &ba)
{
    rg = make_float2(0.f, 0.f);
    ba = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        const float w = __ldg(wp + k);
        const float4 q = buf[swz(e0 + k)];
        const float2 ww = make_float2(w, w);
        rg = __ffma2_rn(make_float2(q.x, q.y), ww, rg);
        ba = __ffma2_rn(make_float2(q.z, q.w), ww, ba);
    }
}

// This thread's entries of target T's cached-form table for the CTA's tile (all rounds).  pv is the tile's
// tile_parts word; a tile without outputs (P == 0) gets an empty one-round table.
template <typename XT>
__device__ __forceinline__ void xtab_fill(XT &xt, XInfo &xi, int T, const StreamTarget &t, int tile, int cx0, int ws, int tid,
                                          int pv, int fix_d)
{
    const int warp = tid >> 5, lane = tid & 31;
    const int P0 = pv & 255, P = max(P0, 1), ntap = P0 ? (pv >> 8) & 255 : 0, R = P0 ? pv >> 16 : 1;
    if (tid == 0) xi = XInfo{t.dst, t.dst_stride, t.exact_job, (uint32_t)fix_d, t.local, P, ntap, R};
    const int part = tid & (P - 1);
#pragma unroll
    for (int r = 0; r < STREAM_XROUNDS; r++) {
        int ox = -1, e0 = 0, k0 = 0, n = 0;
        if (P0 > 0 && r < R) {
            if (t.local) { // P adjacent lanes per output of the owning warp; strip w sits at elements [128 w, 128 w + 128)
                const int ox0 = __ldg(t.warp_ox + tile * 4 + warp);
                const int j = lane / P + r * (32 / P);
                if (j < __ldg(t.warp_ox + tile * 4 + warp + 1) - ox0) ox = ox0 + j;
                if (ox >= 0) e0 = __ldg(t.xfirst + ox) + t.rect_x - cx0 + warp * (STREAM_WARP_COLS - ws) + part;
            } else {       // P adjacent threads per output of the tile
                const int ox0 = __ldg(t.tile_ox + tile);
                const int j = tid / P + r * (STREAM_THREADS / P);
                if (j < __ldg(t.tile_ox + tile + 1) - ox0) ox = ox0 + j;
                if (ox >= 0) e0 = __ldg(t.xfirst + ox) + t.rect_x - cx0 + part;
            }
        }
        if (ox >= 0) {
            k0 = __ldg(t.xoff + ox);
            n = __ldg(t.xoff + ox + 1) - k0;
        }
        xt.ox[T][r][tid] = (int16_t)ox;
        xt.e0[T][r][tid] = (int16_t)e0;
        if (r < R) // rounds past R have no table entries (R * ntap <= STREAM_XTAPS_TAB is what the planner guarantees)
            for (int k = 0; k < ntap; k++)
                xt.w[T][r * ntap + k][tid] = (ox >= 0 && part + k * P < n) ? __ldg(t.xw + k0 + part + k * P) : 0.f;
    }
}

// The cached horizontal pass over the row parked in xbuf[T]: every round of this thread.  All V threads
// of the unit (warp if local, CTA if not) call it together: the butterfly shuffles are warp-wide.
template <bool OPAQUE = false, typename SM>
__device__ __forceinline__ void xcached(SM &sm, int T, int oy, int tid, const FixList &fix)
{
    const XInfo xi = sm.xi[T];
    const int P = xi.parts, ntap = xi.ntap;
    const float4 *buf = sm.xbuf[T];
    for (int r = 0; r < xi.rounds; r++) {
        const int ox = sm.xt.ox[T][r][tid];
        const int e0 = sm.xt.e0[T][r][tid];
        float2 rg = make_float2(0.f, 0.f), ba = make_float2(0.f, 0.f);
        if (ox >= 0) {
#pragma unroll 4
            for (int k = 0; k < ntap; k++) { // weight 0 past the end adds exactly nothing (the buffer is padded and finite)
                const float w = sm.xt.w[T][r * ntap + k][tid];
                const float4 q = buf[swz(e0 + k * P)];
                const float2 ww = make_float2(w, w);
                rg = __ffma2_rn(make_float2(q.x, q.y), ww, rg);
                ba = __ffma2_rn(make_float2(q.z, q.w), ww, ba);
            }
        }
        for (int off = 1; off < P; off <<= 1) { // butterfly over the P threads of an output (P divides 32)
            rg.x += __shfl_xor_sync(0xffffffffu, rg.x, off);
            rg.y += __shfl_xor_sync(0xffffffffu, rg.y, off);
            ba.x += __shfl_xor_sync(0xffffffffu, ba.x, off);
            ba.y += __shfl_xor_sync(0xffffffffu, ba.y, off);
        }
        if (ox >= 0 && (tid & (P - 1)) == 0) xfinish<OPAQUE>(xi, xi.D, ox, oy, rg, ba, fix);
    }
}

// Horizontal pass of the row the V warps just parked for target T (runtime value: this
// code exists once).  Local target: warp-private, one __syncwarp on each side.  Shared
// target: all 128 V threads, one named barrier on each side.
template <int NT, typename SM>
__device__ __noinline__ void xpass(const StreamJob &J, SM &sm, int T, int oy, int tile, int cx0, int vtid, const FixList &fix)
{
    const XInfo xi = sm.xi[T];
    const int warp = vtid >> 5, lane = vtid & 31;
    const bool local = xi.local != 0;
    const int P = xi.parts;
    const float4 *buf = sm.xbuf[T];
    if (local) __syncwarp(); else vwarps_bar(); // the row is parked
    if (P > 0) {
        xcached(sm, T, oy, vtid, fix); // cached form: taps and weights from the shared-memory table
    } else {
        const StreamTarget &t = J.t[T];
        int ox0, n_own, ebase = t.rect_x - cx0, j0 = vtid, step = STREAM_THREADS;
        if (local) { // strips sit at their slab position: warp * 128 + (column - warp * warp_stride)
            ox0 = __ldg(t.warp_ox + tile * 4 + warp);
            n_own = __ldg(t.warp_ox + tile * 4 + warp + 1) - ox0;
            ebase += warp * (STREAM_WARP_COLS - J.warp_stride);
            j0 = lane;
            step = 32;
        } else {
            ox0 = __ldg(t.tile_ox + tile);
            n_own = __ldg(t.tile_ox + tile + 1) - ox0;
        }
        for (int j = j0; j < n_own; j += step) {
            const int ox = ox0 + j;
            const int k0 = __ldg(t.xoff + ox);
            float2 rg, ba;
            xgather(t.xw + k0, __ldg(t.xfirst + ox) + ebase, __ldg(t.xoff + ox + 1) - k0, buf, rg, ba);
            xfinish(xi, xi.D, ox, oy, rg, ba, fix);
        }
    }
    if (local) __syncwarp(); else vwarps_bar(); // the buffer is reused by the next emit
}

// ---- vertical pass -------------------------------------------------------------------
// Per-target accumulators of one V thread (4 source pixels): two sets, one per open
// output row.  RGB of the 4 pixels is 12 floats = 6 packed pairs for FFMA2; the per-pixel
// alpha lanes exist only in the ALPHA instantiation.
template <bool ALPHA> struct VAcc;
template <> struct VAcc<false> { float2 rgb[2][6]; };
template <> struct VAcc<true> { float2 rgb[2][6]; float2 al[2][2]; };

#ifndef IPG_DP4A_MASK
#define IPG_DP4A_MASK 0
#endif
// byte k of q -> the bit pattern of 2^23 + b.  PRMT runs on the ALU pipe, the V loop's busiest; for the
// byte positions in IPG_DP4A_MASK the same word comes from IDP.4A (0x4B000000 + q . (1 << 8k)) on the
// FMA-side integer pipe instead.
template <int K> __device__ __forceinline__ uint32_t magic1(uint32_t q)
{
    if constexpr ((IPG_DP4A_MASK >> K) & 1) return __dp4a(q, 1u << (8 * K), 0x4B000000u);
    else return __byte_perm(q, 0x4B000000u, 0x7540 + K);
}
template <int KA, int KB> __device__ __forceinline__ float2 magic2(uint32_t qa, uint32_t qb)
{
    // two bytes -> (2^23 + b) bit patterns; the caller subtracts 2^23 with one FADD2
    return make_float2(__uint_as_float(magic1<KA>(qa)), __uint_as_float(magic1<KB>(qb)));
}
__device__ __forceinline__ void unpack_rgb(const uint4 &c, float2 *vp)
{
    const float2 m = make_float2(-8388608.0f, -8388608.0f);
    vp[0] = __fadd2_rn(magic2<0, 1>(c.x, c.x), m);
    vp[1] = __fadd2_rn(magic2<2, 0>(c.x, c.y), m);
    vp[2] = __fadd2_rn(magic2<1, 2>(c.y, c.y), m);
    vp[3] = __fadd2_rn(magic2<0, 1>(c.z, c.z), m);
    vp[4] = __fadd2_rn(magic2<2, 0>(c.z, c.w), m);
    vp[5] = __fadd2_rn(magic2<1, 2>(c.w, c.w), m);
}
__device__ __forceinline__ void unpack_alpha(const uint4 &c, float2 *va)
{
    const float2 m = make_float2(-8388608.0f, -8388608.0f);
    va[0] = __fadd2_rn(magic2<3, 3>(c.x, c.y), m);
    va[1] = __fadd2_rn(magic2<3, 3>(c.z, c.w), m);
}

// Park one completed, vertically-filtered row (accumulator set SET) at elements
// 4*slot .. 4*slot+3 of `buf` (XOR-swizzled) and clear the set.  `sa` is the opaque alpha
// chain value (ALPHA=false).
__device__ __forceinline__ void sts128(float4 *p, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
template <int SET, bool ALPHA>
__device__ __forceinline__ void park_row(VAcc<ALPHA> &S, float sa, float4 *buf, int slot)
{
    const int base = (slot * 4) & ~7, key = (slot >> 1) & 7, lo = (slot & 1) * 4; // swz(4*slot + j)
    const float2 *g = S.rgb[SET];
    float a0 = sa, a1 = sa, a2 = sa, a3 = sa;
    if constexpr (ALPHA) { a0 = S.al[SET][0].x; a1 = S.al[SET][0].y; a2 = S.al[SET][1].x; a3 = S.al[SET][1].y; }
    sts128(&buf[base | ((lo + 0) ^ key)], g[0].x, g[0].y, g[1].x, a0);
    sts128(&buf[base | ((lo + 1) ^ key)], g[1].y, g[2].x, g[2].y, a1);
    sts128(&buf[base | ((lo + 2) ^ key)], g[3].x, g[3].y, g[4].x, a2);
    sts128(&buf[base | ((lo + 3) ^ key)], g[4].y, g[5].x, g[5].y, a3);
#pragma unroll
    for (int i = 0; i < 6; i++) S.rgb[SET][i] = make_float2(0.f, 0.f);
    if constexpr (ALPHA) S.al[SET][0] = S.al[SET][1] = make_float2(0.f, 0.f);
}

// Park the set that source-row record `r` completes (emit word e: output row << 1 | set).  e is
// CTA-uniform.  A convergence barrier in each arm keeps the two set variants in real branches:
// if-converted, both variants' ~30 instructions would issue at every emit.
template <bool ALPHA>
__device__ __forceinline__ void park_emit(VAcc<ALPHA> &S, int e, const float4 &r, float4 *buf, int slot)
{
    if (e & 1) {
        park_row<1, ALPHA>(S, r.w, buf, slot);
        asm volatile("bar.warp.sync 0xffffffff; // set 1" ::: "memory");
    } else {
        park_row<0, ALPHA>(S, r.z, buf, slot);
        asm volatile("bar.warp.sync 0xffffffff; // set 0" ::: "memory");
    }
}

// What a V thread carries through the row loop besides its accumulators.
struct VCtx {
    int tile, cx0, vtid;
    int slot;        // this thread's 4-pixel slot within the slab (index into a ring row)
    int pslot[2];    // where it parks its 4 columns of target T (strip slot if local, slab slot if shared)
    bool act[2], clampt[2];
    // target 0 in its common form (local, one output per lane, <= STREAM_XTAPS taps): the horizontal
    // pass runs inline with the taps in registers; every other case goes through xpass()
    bool x0_inline;
    int x0_parts;            // lanes sharing one output (1, 2 or 4)
    int x0_ox;               // output column of this lane, or -1
    const float4 *x0_tap;    // address of its first tap in xbuf[0] when no swizzle step is crossed ... see x0_off
    int x0_e0;
    float x0_w[STREAM_XTAPS];
    // fallback watermark copy
    bool wm_v;
    uint8_t *wm_dst;
    int wm_stride, c, W, ys1;
};

// The common case, stripped of every test it does not need: one target, active in this tile,
// in the lane-per-output form, all four rows inside the band, opaque so far, watermark (if any)
// copied by the producer.  Everything else takes v_rows below.
template <int LEAN, typename SM>
__device__ __forceinline__ uint32_t v_rows_fast(VAcc<false> &S, const StreamJob &J, const StreamStage &stg, SM &sm, const VCtx &C,
                                                const FixList &fix)
{
    uint4 cur = stg.rows[0][C.slot];
    uint32_t opq = 0xffffffffu; // AND of every pixel word met: all four alpha bytes are 0xff iff opq >= 0xff000000
#pragma unroll kFastUnroll
    for (int k = 0; k < STREAM_GROUP; k++) {
        const uint4 nxt = stg.rows[(k + 1) & (STREAM_GROUP - 1)][C.slot];
        opq &= (cur.x & cur.y) & (cur.z & cur.w);
        float2 vp[6];
        unpack_rgb(cur, vp);
        const float4 r = *reinterpret_cast<const float4 *>(&stg.rec[0].row[k]); // LDS.128 broadcast
        const float2 w00 = make_float2(r.x, r.x), w11 = make_float2(r.y, r.y);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            S.rgb[0][i] = __ffma2_rn(vp[i], w00, S.rgb[0][i]);
            S.rgb[1][i] = __ffma2_rn(vp[i], w11, S.rgb[1][i]);
        }
        const int e = stg.rec[0].emit[k];
        if (e >= 0) { // CTA-uniform: this source row completes output row e>>1, held in set e&1
            park_emit<false>(S, e, r, sm.xbuf[0], C.pslot[0]);
            if constexpr (LEAN == 1) { // local target: this warp filters its own strip, one output per lane
                __syncwarp();
                const float4 *buf = sm.xbuf[0];
                float2 rg = make_float2(0.f, 0.f), ba = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < STREAM_XTAPS; q++) { // weight 0 past the end adds exactly nothing
                    const float4 v = buf[swz(C.x0_e0 + q * C.x0_parts)];
                    const float2 ww = make_float2(C.x0_w[q], C.x0_w[q]);
                    rg = __ffma2_rn(make_float2(v.x, v.y), ww, rg);
                    ba = __ffma2_rn(make_float2(v.z, v.w), ww, ba);
                }
                for (int off = 1; off < C.x0_parts; off <<= 1) { // lanes sharing an output (supports of 9..32 taps)
                    rg.x += __shfl_xor_sync(0xffffffffu, rg.x, off);
                    rg.y += __shfl_xor_sync(0xffffffffu, rg.y, off);
                    ba.x += __shfl_xor_sync(0xffffffffu, ba.x, off);
                    ba.y += __shfl_xor_sync(0xffffffffu, ba.y, off);
                }
                if (C.x0_ox >= 0 && (threadIdx.x & (C.x0_parts - 1)) == 0) xfinish<true>(sm.xi[0], sm.xi[0].D, C.x0_ox, e >> 1, rg, ba, fix);
                __syncwarp(); // the strip is reused by the next emit
            } else if constexpr (LEAN == 4 && !IPG_LEAN4_TABLES) {
                // (not this build's case: the engine gives such jobs to the wide-target instantiation)
            } else if constexpr (LEAN == 4) { // table forms, inline (no call: the accumulators stay in registers):
                const bool loc = sm.xi[0].local != 0; // a wide target, or a local one with several outputs per lane
                if (loc) __syncwarp(); else vwarps_bar();
                xcached<true>(sm, 0, e >> 1, (int)threadIdx.x, fix);
                if (loc) __syncwarp(); else vwarps_bar(); // the row buffer is reused by the next emit
            } else {           // wide support (the thumbnail): CTA-wide split pass, once per ~15 rows
                xpass<1>(J, sm, 0, e >> 1, C.tile, C.cx0, C.vtid, fix);
            }
        }
        cur = nxt;
    }
    return opq;
}

// The lean two-target case: target 0 local (inline lane-per-output pass), target 1 wide (split pass
// through xpass, once per ~15 rows) -- resize + thumbnail (+ watermark copy) with the source read once.
template <typename SM>
__device__ __forceinline__ void v_rows_fast2(VAcc<false> *S, const StreamJob &J, const StreamStage &stg, SM &sm, const VCtx &C,
                                             const FixList &fix)
{
    uint4 cur = stg.rows[0][C.slot];
#pragma unroll
    for (int k = 0; k < STREAM_GROUP; k++) {
        const uint4 nxt = stg.rows[(k + 1) & (STREAM_GROUP - 1)][C.slot];
        float2 vp[6];
        unpack_rgb(cur, vp);
        {
            const float4 r = *reinterpret_cast<const float4 *>(&stg.rec[0].row[k]); // LDS.128 broadcast
            const float2 w00 = make_float2(r.x, r.x), w11 = make_float2(r.y, r.y);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                S[0].rgb[0][i] = __ffma2_rn(vp[i], w00, S[0].rgb[0][i]);
                S[0].rgb[1][i] = __ffma2_rn(vp[i], w11, S[0].rgb[1][i]);
            }
            const int e = stg.rec[0].emit[k];
            if (e >= 0) { // CTA-uniform
                park_emit<false>(S[0], e, r, sm.xbuf[0], C.pslot[0]);
                __syncwarp();
                const float4 *buf = sm.xbuf[0];
                float2 rg = make_float2(0.f, 0.f), ba = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < STREAM_XTAPS; q++) { // weight 0 past the end adds exactly nothing
                    const float4 v = buf[swz(C.x0_e0 + q * C.x0_parts)];
                    const float2 ww = make_float2(C.x0_w[q], C.x0_w[q]);
                    rg = __ffma2_rn(make_float2(v.x, v.y), ww, rg);
                    ba = __ffma2_rn(make_float2(v.z, v.w), ww, ba);
                }
                for (int off = 1; off < C.x0_parts; off <<= 1) {
                    rg.x += __shfl_xor_sync(0xffffffffu, rg.x, off);
                    rg.y += __shfl_xor_sync(0xffffffffu, rg.y, off);
                    ba.x += __shfl_xor_sync(0xffffffffu, ba.x, off);
                    ba.y += __shfl_xor_sync(0xffffffffu, ba.y, off);
                }
                if (C.x0_ox >= 0 && (threadIdx.x & (C.x0_parts - 1)) == 0) xfinish(sm.xi[0], sm.xi[0].D, C.x0_ox, e >> 1, rg, ba, fix);
                __syncwarp(); // the strip is reused by the next emit
            }
        }
        if (C.act[1]) { // CTA-uniform: this tile overlaps the crop square
            const float4 r = *reinterpret_cast<const float4 *>(&stg.rec[1].row[k]);
            const float2 w00 = make_float2(r.x, r.x), w11 = make_float2(r.y, r.y);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                S[1].rgb[0][i] = __ffma2_rn(vp[i], w00, S[1].rgb[0][i]);
                S[1].rgb[1][i] = __ffma2_rn(vp[i], w11, S[1].rgb[1][i]);
            }
            const int e = stg.rec[1].emit[k];
            if (e >= 0) {
                park_emit<false>(S[1], e, r, sm.xbuf[1], C.pslot[1]);
                xpass<2>(J, sm, 1, e >> 1, C.tile, C.cx0, C.vtid, fix);
            }
        }
        cur = nxt;
    }
}

// The source rows of one ring stage into every active target; rows are taken one at a
// time (the next one is requested before this one's math) so the code stays small.
template <int NT, bool WM, bool ALPHA, typename SM>
__device__ __forceinline__ void v_rows(VAcc<ALPHA> *S, const StreamJob &J, const StreamStage &stg, SM &sm, const VCtx &C,
                                       int ys_first, int nr, const FixList &fix)
{
    uint4 cur = stg.rows[0][C.slot];
#pragma unroll kRowUnroll
    for (int k = 0; k < STREAM_GROUP; k++) {
        const uint4 nxt = stg.rows[(k + 1) & (STREAM_GROUP - 1)][C.slot];
        // band tail: rows past the end are stale ring contents (their weights are 0)
        if (k >= nr) cur = make_uint4(0xff000000u, 0xff000000u, 0xff000000u, 0xff000000u);
        if (WM && C.wm_v && ys_first + k < C.ys1) {
            uint32_t *d = (uint32_t *)(C.wm_dst + (size_t)(ys_first + k) * C.wm_stride + (size_t)C.c * 4);
            d[0] = cur.x;
            if (C.c + 1 < C.W) d[1] = cur.y;
            if (C.c + 2 < C.W) d[2] = cur.z;
            if (C.c + 3 < C.W) d[3] = cur.w;
        }
        if constexpr (NT > 0) {
            float2 vp[6], va[2];
            unpack_rgb(cur, vp);
            if (ALPHA) unpack_alpha(cur, va);
            int pending = 0, oyv[NT > 0 ? NT : 1];
#pragma unroll
            for (int T = 0; T < NT; T++) {
                if (!C.act[T]) continue; // CTA-uniform
                const float4 r = *reinterpret_cast<const float4 *>(&stg.rec[T].row[k]); // LDS.128 broadcast
                const float2 w00 = make_float2(r.x, r.x), w11 = make_float2(r.y, r.y);
                if constexpr (!ALPHA) {
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        S[T].rgb[0][i] = __ffma2_rn(vp[i], w00, S[T].rgb[0][i]);
                        S[T].rgb[1][i] = __ffma2_rn(vp[i], w11, S[T].rgb[1][i]);
                    }
                } else {
                    float2 up[6];
                    if (C.clampt[T]) {
                        // cropAndResize quantises to premultiplied RGBA8 first: channels clamp to alpha
                        const uint4 cc = make_uint4(clamp_to_alpha(cur.x), clamp_to_alpha(cur.y), clamp_to_alpha(cur.z),
                                                    clamp_to_alpha(cur.w));
                        unpack_rgb(cc, up);
                    } else {
#pragma unroll
                        for (int i = 0; i < 6; i++) up[i] = vp[i];
                    }
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        S[T].rgb[0][i] = __ffma2_rn(up[i], w00, S[T].rgb[0][i]);
                        S[T].rgb[1][i] = __ffma2_rn(up[i], w11, S[T].rgb[1][i]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        S[T].al[0][i] = __ffma2_rn(va[i], w00, S[T].al[0][i]);
                        S[T].al[1][i] = __ffma2_rn(va[i], w11, S[T].al[1][i]);
                    }
                }
                const int e = stg.rec[T].emit[k];
                oyv[T] = e >> 1;
                if (e >= 0) { // CTA-uniform: this source row completes output row e>>1, held in set e&1
                    park_emit<ALPHA>(S[T], e, r, sm.xbuf[T], C.pslot[T]);
                    if (T == 0 && C.x0_inline) {
                        __syncwarp();
                        float2 x0_rg = make_float2(0.f, 0.f), x0_ba = make_float2(0.f, 0.f);
                        if (C.x0_ox >= 0) {
                            const float4 *buf = sm.xbuf[0];
                            float2 &rg = x0_rg, &ba = x0_ba;
#pragma unroll
                            for (int q = 0; q < STREAM_XTAPS; q++) { // weight 0 past the end adds exactly nothing
                                const float4 v = buf[swz(C.x0_e0 + q * C.x0_parts)];
                                const float2 ww = make_float2(C.x0_w[q], C.x0_w[q]);
                                rg = __ffma2_rn(make_float2(v.x, v.y), ww, rg);
                                ba = __ffma2_rn(make_float2(v.z, v.w), ww, ba);
                            }
                        }
                        {
                            float2 &rg = x0_rg, &ba = x0_ba;
                            for (int off = 1; off < C.x0_parts; off <<= 1) {
                                rg.x += __shfl_xor_sync(0xffffffffu, rg.x, off);
                                rg.y += __shfl_xor_sync(0xffffffffu, rg.y, off);
                                ba.x += __shfl_xor_sync(0xffffffffu, ba.x, off);
                                ba.y += __shfl_xor_sync(0xffffffffu, ba.y, off);
                            }
                            if (C.x0_ox >= 0 && (threadIdx.x & (C.x0_parts - 1)) == 0)
                                xfinish(sm.xi[0], sm.xi[0].D, C.x0_ox, e >> 1, rg, ba, fix);
                        }
                        __syncwarp(); // the strip is reused by the next emit
                    } else {
                        pending |= 1 << T;
                    }
                }
            }
            if (pending & 1) xpass<NT>(J, sm, 0, oyv[0], C.tile, C.cx0, C.vtid, fix);
            if constexpr (NT > 1) {
                if (pending & 2) xpass<NT>(J, sm, 1, oyv[1], C.tile, C.cx0, C.vtid, fix);
            }
        }
        cur = nxt;
    }
}

// FAST (NT == 1 only): the lean instantiation for the common case -- one target whose horizontal
// pass is in a cached form (local: one output per lane; wide: split over adjacent threads), watermark
// copied by the producer, every pixel opaque.  It carries none of the
// general paths, so it fits 4 CTAs per SM in ~96 registers and runs the fused resize + watermark
// copy near the HBM roofline.  Opacity is checked, not assumed: a warp that meets a non-opaque
// pixel raises the job's redo flag and the general instantiation, launched right after over the same
// items, redoes exactly the flagged jobs (it exits at once for the others).
template <int NT, bool WM, int LEAN>
__global__ void __launch_bounds__((StreamCfg<NT, LEAN>::THREADS), (StreamCfg<NT, LEAN>::CTAS_PER_SM))
k_stream(const StreamJob *__restrict__ jobs, const StreamItem *__restrict__ items, FixList fix)
{
    // LEAN 1: one local target (lane-per-output pass inline); 2: one wide target (split pass);
    // 3: a local and a wide target fused (the source is read once for resize + thumbnail);
    // 4: one target, local or wide decided per CTA (a row loop per horizontal-pass form, all inline, plus the
    //    integer-moment loop of wide 8-bit targets): CTAs of every kind share the SMs
    constexpr bool FAST = LEAN != 0;
    constexpr bool VINT = LEAN == 2 || LEAN == 4; // these run a wide target's vertical pass in the integer-moment form when the job has one
    static_assert(!FAST || (LEAN == 3 ? NT == 2 : NT == 1), "lean instantiations: one target, or local + wide");
    constexpr int STAGES = StreamCfg<NT, LEAN>::STAGES;
    constexpr bool INLINE = StreamCfg<NT, LEAN>::INLINE_PROD;
    using Smem = typename StreamCfg<NT, LEAN>::Smem;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // The shared-window address of the CTA's memory carries its rank in the cluster (SR_CgaCtaId << 24) and the compiler
    // rematerialises it -- an S2R in front of every group's loads and barrier waits -- rather than keep it in a register.
    // Computed once and made opaque, it stays; the assumption gives the pointer its address space back (LDS / STS).
    uint32_t sm32 = smem_u32(smem_raw);
    asm volatile("" : "+r"(sm32));
    Smem *smp = reinterpret_cast<Smem *>(__cvta_shared_to_generic(sm32));
    __builtin_assume(__isShared(smp));
    Smem &sm = *smp;

    const StreamItem it = items[blockIdx.x];
    const StreamJob &J = jobs[it.job];
    // a job the lean kernel already ran is redone here only if it met a non-opaque pixel
    if (!FAST && J.fast_path != 0 && (J.redo_flag == nullptr || *(volatile const int32_t *)J.redo_flag == 0)) return;
    const int tile = it.tile, band = it.band;
    const int W = J.src.w;
    const int cx0 = tile * J.tile_w;
    const int ys0 = __ldg(J.band_y + band), ys1 = __ldg(J.band_y + band + 1);
    const int yend = __ldg(J.band_yend + band);
    const int stride = J.src.s0;
    const int warp = threadIdx.x >> 5;
    const int ngroups = (yend - ys0 + STREAM_GROUP - 1) / STREAM_GROUP;
    const int tid = (int)threadIdx.x;
    const int ws = J.warp_stride;
    const int slot = (warp & 3) * (ws >> 2) + (tid & 31); // this thread's 4-pixel slot within the slab
    const int c = cx0 + slot * STREAM_PX;
    const uint32_t row_bytes = (uint32_t)(((min(J.slab_cols, W - cx0) * 4) + 15) & ~15);
    // watermark copy: owned columns of the band's own rows; by TMA when 16-byte granular
    const bool has_wm = WM && J.has_wm;
    const uint32_t wm_bytes = (uint32_t)(min(J.tile_w, W - cx0) * 4);
    const bool wm_tma = has_wm && (wm_bytes & 15) == 0 && ((((size_t)J.wm.dst) | (size_t)J.wm.dst_stride) & 15) == 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], STREAM_THREADS / 32);
            sm.done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.ring_src = J.src.p0 + (size_t)ys0 * stride + (size_t)cx0 * 4;
        sm.ring_rec = (const uint8_t *)(J.grec + (size_t)__ldg(J.band_grec_off + band) * (size_t)J.rec_slots);
        sm.ring_rec_bytes = (uint32_t)(NT > 0 ? J.rec_slots * (int)sizeof(GroupRec) : 0);
        sm.ring_wm = has_wm ? J.wm.dst + (size_t)ys0 * J.wm.dst_stride + (size_t)cx0 * 4 : nullptr;
        sm.ring_wm_stride = has_wm ? J.wm.dst_stride : 0;
    }
    // slab columns past the image edge are never written by TMA: park opaque black there
    // so they neither look transparent nor feed anything (no output taps them)
    if (tid < STREAM_THREADS && c >= W) {
        for (int s = 0; s < STAGES; s++)
            for (int k = 0; k < STREAM_GROUP; k++)
                sm.stage[s].rows[k][slot] = make_uint4(0xff000000u, 0xff000000u, 0xff000000u, 0xff000000u);
    }
    // horizontal-pass tables of the targets whose tile is in cached form, one entry per V thread
    if constexpr (NT > 0) {
        if (tid < STREAM_THREADS) {
#pragma unroll
            for (int T = 0; T < NT; T++) {
                if (T >= J.n_targets) continue;
                const StreamTarget &t = J.t[T];
                for (int e = tid; e < STREAM_XBUF; e += STREAM_THREADS) sm.xbuf[T][e] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int pv = __ldg(t.tile_parts + tile);
                // the lean kernels always run the table form (a tile without outputs gets an empty table);
                // the general one keeps its generic loops when the tile has no cached form
                if (FAST || (pv & 255)) xtab_fill(sm.xt, sm.xi[T], T, t, tile, cx0, ws, tid, pv, (VINT && J.vint) ? t.fix_d_vint : t.fix_d);
                else if (tid == 0) sm.xi[T] = XInfo{t.dst, t.dst_stride, t.exact_job, (uint32_t)t.fix_d, t.local, 0, 0, 1};
            }
        }
    }
    __syncthreads();

    // one ring stage <- one group of source rows + its records, under one mbarrier phase
    auto refill = [&](int group, int stage) {
        const uint32_t rec_bytes = sm.ring_rec_bytes;
        const int nr = min(STREAM_GROUP, yend - ys0 - group * STREAM_GROUP);
        StreamStage &st = sm.stage[stage];
        mbar_arrive_expect_tx(&sm.full[stage], row_bytes * (uint32_t)nr + rec_bytes);
        const uint8_t *g = sm.ring_src + (size_t)group * STREAM_GROUP * stride;
        for (int k = 0; k < nr; k++, g += stride) tma_load_1d(&st.rows[k][0], g, row_bytes, &sm.full[stage]);
        if (rec_bytes) tma_load_1d(&st.rec[0], sm.ring_rec + (size_t)group * rec_bytes, rec_bytes, &sm.full[stage]);
    };
    // the watermark copy of a landed group: the owned columns of the rows the band owns, ring -> destination
    auto store_group = [&](int group, int stage) {
        const int nw = min(STREAM_GROUP, ys1 - ys0 - group * STREAM_GROUP);
        if (nw > 0) {
            const int wm_stride = sm.ring_wm_stride;
            uint8_t *d = sm.ring_wm + (size_t)group * STREAM_GROUP * wm_stride;
            for (int k = 0; k < nw; k++, d += wm_stride) tma_store_1d(d, &sm.stage[stage].rows[k][0], wm_bytes);
        }
        tma_store_commit(); // one (possibly empty) bulk group per ring group keeps the wait counts uniform
    };
    if constexpr (!INLINE) {
        if (warp == STREAM_THREADS / 32) {
            // ===== producer: one lane drives the ring =====
            if ((tid & 31) != 0) return;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int g = 0; g < min(STAGES, ngroups); g++) refill(g, g); // prologue: fill the ring
            int s = 0, sp = 0;        // stage of group g / of group g-1
            uint32_t ph = 0, php = 0; // their phases
            for (int g = 0; g < ngroups; g++) {
                if (wm_tma) {
                    if (ys1 - ys0 - g * STREAM_GROUP > 0) mbar_wait(&sm.full[s], ph);
                    store_group(g, s);
                }
                if (g >= 1) {
                    if (g - 1 + STAGES < ngroups) {
                        mbar_wait(&sm.empty[sp], php);            // V warps are done with group g-1
                        if (wm_tma) tma_store_wait_read<1>();     // ... and so is its store (all but group g's)
                        refill(g - 1 + STAGES, sp);
                    }
                    if (++sp == STAGES) { sp = 0; php ^= 1; }
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (wm_tma) tma_store_wait_all(); // shared memory must outlive the reads
            return;
        }
    } else {
        // no producer warp: lane 0 of V warp (g mod 4) stores group g when it lands; the last V warp to finish a
        // group refills its stage (nobody ever waits for a slower warp); here the ring is filled for the first time
        if ((tid & 31) == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0)
            for (int g = 0; g < min(STAGES, ngroups); g++) refill(g, g);
    }

    // ===== V warps: vertical pass, and the horizontal pass of every row they complete =====
    uint32_t bar32 = smem_u32(&sm.full[0]); // full[s] at bar32 + 8 s, empty[s] kEmptyOff further on (see mbar_wait_a)
    asm volatile("" : "+r"(bar32));
    constexpr uint32_t kEmptyOff = (uint32_t)(offsetof(Smem, empty) - offsetof(Smem, full));
    VCtx C;
    C.tile = tile; C.cx0 = cx0; C.vtid = tid; C.slot = slot;
    C.c = c; C.W = W; C.ys1 = ys1;
#pragma unroll
    for (int T = 0; T < 2; T++) {
        C.act[T] = T < NT && T < J.n_targets && __ldg(J.t[T].tile_ox + tile + 1) > __ldg(J.t[T].tile_ox + tile) &&
                   __ldg(J.t[T].band_tend + band) > ys0;
        C.clampt[T] = T < NT && T < J.n_targets && J.t[T].two_stage != 0;
        C.pslot[T] = (T < NT && T < J.n_targets && J.t[T].local) ? tid : slot;
    }
    C.x0_inline = false;
    C.x0_parts = 1;
    C.x0_ox = -1;
    C.x0_e0 = 0;
    C.x0_tap = nullptr;
#pragma unroll
    for (int k = 0; k < STREAM_XTAPS; k++) C.x0_w[k] = 0.f;
    if constexpr (NT > 0) {
        if (LEAN == 1 || LEAN == 3 || (LEAN == 4 && sm.xi[0].local && sm.xi[0].rounds == 1) ||
            (!FAST && C.act[0] && sm.xi[0].local && sm.xi[0].parts >= 1 && sm.xi[0].rounds == 1)) {
            C.x0_inline = true;
            C.x0_parts = sm.xi[0].parts;
            C.x0_ox = sm.xt.ox[0][0][tid];
            C.x0_e0 = sm.xt.e0[0][0][tid];
#pragma unroll
            for (int k = 0; k < STREAM_XTAPS; k++) C.x0_w[k] = k < sm.xi[0].ntap ? sm.xt.w[0][k][tid] : 0.f; // the table holds ntap entries
        }
    }
    // fallback watermark copy by the V warps when the rows are not 16-byte granular
    C.wm_v = has_wm && !wm_tma && c < min(W, cx0 + J.tile_w) &&
             (warp == 3 || (tid & 31) * STREAM_PX < ws); // overlapped columns: the right-hand warp writes them
    C.wm_dst = has_wm ? J.wm.dst : nullptr;
    C.wm_stride = has_wm ? J.wm.dst_stride : 0;

    VAcc<false> S[NT > 0 ? NT : 1];
#pragma unroll
    for (int T = 0; T < NT; T++)
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int i = 0; i < 6; i++) S[T].rgb[k][i] = make_float2(0.f, 0.f);
    int rs = 0, g = 0;
    uint32_t rph = 0;
    auto advance = [&]() {
        __syncwarp();
        if ((tid & 31) == 0) {
            if constexpr (INLINE) {
                // the lane that stored this group lets the copy finish reading the stage before it counts as done (the
                // copy was issued a whole group ago)
                if (wm_tma && (g & 3) == warp) tma_store_wait_read<0>();
            }
            mbar_arrive_a(bar32 + kEmptyOff + 8u * (uint32_t)rs);
            if constexpr (INLINE) {
                if (atomicAdd(&sm.done[rs], 1u) == STREAM_THREADS / 32 - 1) { // the last of the four: refill the stage
                    sm.done[rs] = 0;
                    if (g + STAGES < ngroups) {
                        mbar_wait(&sm.empty[rs], rph); // complete by now: taken for its acquire, as the producer warp does
                        refill(g + STAGES, rs);
                    }
                }
            }
        }
        if (++rs == STAGES) { rs = 0; rph ^= 1; }
    };

    // Integer-moment vertical pass (GroupRecI, ipg_device.h): a wide target whose rows between two output centres
    // enter two exact integer moments per byte column -- one IDP.2A per channel and row, no byte -> fp32 unpack, no
    // per-row weights -- and become fp32 once per piece (a segment, or half of one longer than 16 rows): the open row
    // gains aR M0 + bR M1, the carry into the next one aL M0 + bL M1.  The parked row then takes the same CTA-wide
    // horizontal pass as the fp32 form.
    if constexpr (VINT) {
        if (J.vint) { // CTA-uniform
            uint32_t A[12];
            float2 cy[6], nx[6]; // the open output row so far, and the carry into the next one
#pragma unroll
            for (int i = 0; i < 12; i++) A[i] = 0u;
#pragma unroll
            for (int i = 0; i < 6; i++) cy[i] = nx[i] = make_float2(0.f, 0.f);
            const bool check = J.redo_flag != nullptr;
            float4 *xb = sm.xbuf[0];
            const int pbase = (slot * 4) & ~7, pkey = (slot >> 1) & 7, plo = (slot & 1) * 4; // swz(4 * slot + j), as park_row
            // the 12 byte columns of one source row into their moment words: [m, 0] takes bytes 0 / 2, [0, m] bytes 1 / 3
            auto row_in = [&](const uint4 &q, uint32_t m) {
                const uint32_t mh = m << 16;
                A[0] = __dp2a_lo(m, q.x, A[0]); A[1] = __dp2a_lo(mh, q.x, A[1]); A[2] = __dp2a_hi(m, q.x, A[2]);
                A[3] = __dp2a_lo(m, q.y, A[3]); A[4] = __dp2a_lo(mh, q.y, A[4]); A[5] = __dp2a_hi(m, q.y, A[5]);
                A[6] = __dp2a_lo(m, q.z, A[6]); A[7] = __dp2a_lo(mh, q.z, A[7]); A[8] = __dp2a_hi(m, q.z, A[8]);
                A[9] = __dp2a_lo(m, q.w, A[9]); A[10] = __dp2a_lo(mh, q.w, A[10]); A[11] = __dp2a_hi(m, q.w, A[11]);
            };
            for (; g < ngroups; g++) {
                mbar_wait_a(bar32 + 8u * (uint32_t)rs, rph);
                if constexpr (INLINE) {
                    if (wm_tma && (g & 3) == warp && (tid & 31) == 0) store_group(g, rs);
                }
                const StreamStage &stg = sm.stage[rs];
                const GroupRecI &G = *reinterpret_cast<const GroupRecI *>(&stg.rec[1]);
                const int nr = yend - ys0 - g * STREAM_GROUP;
                // At most one segment ends in a group (after row end_k; STREAM_GROUP - 1 when none): the rows up to it
                // go in first, then the end -- the one copy of that code: the row loop stays small enough for the
                // instruction cache it shares with the resize CTAs -- then the rows after it.
                const int ke = G.end_k, e = G.end_e;
                uint32_t opq = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < STREAM_GROUP; k++) {
                    const uint4 q = stg.rows[k][slot];
                    if (k < nr) opq &= (q.x & q.y) & (q.z & q.w); // (rows past the band's end are stale ring contents, m = 0)
                    row_in(q, k <= ke ? G.m[k] : 0u);
                }
                if (e != -1) { // CTA-uniform: a piece ends in this group: its moments become fp32
                    const float4 cf = *reinterpret_cast<const float4 *>(&G.aR);
                    const float2 aR = make_float2(cf.x, cf.x), bR = make_float2(cf.y, cf.y);
                    const float2 aL = make_float2(cf.z, cf.z), bL = make_float2(cf.w, cf.w);
                    float2 f0[6], f1[6];
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        const uint32_t a0 = A[2 * i], a1 = A[2 * i + 1];
                        f0[i] = make_float2((float)(a0 & 0xfffu), (float)(a1 & 0xfffu));
                        f1[i] = make_float2((float)(a0 >> 12), (float)(a1 >> 12));
                        cy[i] = __ffma2_rn(f1[i], bR, __ffma2_rn(f0[i], aR, cy[i]));
                        A[2 * i] = A[2 * i + 1] = 0u;
                    }
                    if (e == -3) { // half a segment: the carry goes on accumulating
#pragma unroll
                        for (int i = 0; i < 6; i++) nx[i] = __ffma2_rn(f1[i], bL, __ffma2_rn(f0[i], aL, nx[i]));
                    } else {       // ... and its segment with it
                        if (e >= 0) { // output row e is complete
                            sts128(&xb[pbase | ((plo + 0) ^ pkey)], cy[0].x, cy[0].y, cy[1].x, 65535.0f);
                            sts128(&xb[pbase | ((plo + 1) ^ pkey)], cy[1].y, cy[2].x, cy[2].y, 65535.0f);
                            sts128(&xb[pbase | ((plo + 2) ^ pkey)], cy[3].x, cy[3].y, cy[4].x, 65535.0f);
                            sts128(&xb[pbase | ((plo + 3) ^ pkey)], cy[4].y, cy[5].x, cy[5].y, 65535.0f);
                        }
                        // the carry, with this piece's share, becomes the open row (written where the old row sat: no moves)
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            cy[i] = __ffma2_rn(f1[i], bL, __ffma2_rn(f0[i], aL, nx[i]));
                            nx[i] = make_float2(0.f, 0.f);
                        }
                        if (e >= 0) {
                            vwarps_bar();
                            xcached<true>(sm, 0, e, tid, fix);
                            vwarps_bar(); // the row buffer is reused by the next emit
                        }
                    }
                    if (ke < STREAM_GROUP - 1) {
#pragma unroll
                        for (int k = 1; k < STREAM_GROUP; k++) row_in(stg.rows[k][slot], k > ke ? G.m[k] : 0u);
                    }
                }
                if (check && __any_sync(0xffffffffu, opq < 0xff000000u) && (tid & 31) == 0) atomicExch(J.redo_flag, 1);
                advance();
            }
            if constexpr (INLINE) {
                if (wm_tma && (tid & 31) == 0) tma_store_wait_all();
            }
            return;
        }
    }

    // phase 1: every pixel this warp has met so far is opaque -- alpha comes from the records.
    // The general kernel scans a group before it runs it (it switches to per-pixel alpha from that
    // group on); the lean single-target kernels only raise the redo flag, so they fold the scan into
    // the row loop (one AND per row on registers they hold anyway) and vote after the group.
    bool switch_alpha = false;
    constexpr bool FOLD = FAST && LEAN != 3;
    const bool check = NT > 0 && (!FAST || J.redo_flag != nullptr); // read once: the asm memory clobbers would reload it per group
    for (; g < ngroups; g++) {
        mbar_wait_a(bar32 + 8u * (uint32_t)rs, rph);
        if constexpr (INLINE) {
            if (wm_tma && (g & 3) == warp && (tid & 31) == 0) store_group(g, rs);
        }
        const StreamStage &stg = sm.stage[rs];
        const int nr = yend - ys0 - g * STREAM_GROUP;
        const bool folded = FOLD && nr >= STREAM_GROUP; // a partial last group holds stale rows: scan it the guarded way
        if (check && !folded) {
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < STREAM_GROUP; k++) {
                const uint4 q = stg.rows[k][slot];
                if (k < nr) m = min(m, min(min(q.x, q.y), min(q.z, q.w)));
            }
            if (__any_sync(0xffffffffu, m < 0xff000000u)) {
                if constexpr (FAST) { // not this kernel's case: have the general one redo the job
                    if ((tid & 31) == 0) atomicExch(J.redo_flag, 1);
                } else {
                    switch_alpha = true;
                    break;
                }
            }
        }
        if constexpr (LEAN == 3) v_rows_fast2(S, J, stg, sm, C, fix);
        else if constexpr (FAST) {
            // The merged instantiation keeps one row loop per horizontal-pass form and picks per CTA: the lane-per-output
            // local pass (the 12 MP resize) never shares its loop body with the table forms' code, so what a batch of
            // one kind executes fits the instruction cache (measured: 24.1 -> 23.7 us per 12 MP image, r+t+w).
            uint32_t opq;
            if constexpr (LEAN == 4) opq = C.x0_inline ? v_rows_fast<1>(S[0], J, stg, sm, C, fix) : v_rows_fast<4>(S[0], J, stg, sm, C, fix);
            else opq = v_rows_fast<LEAN>(S[0], J, stg, sm, C, fix);
            if (check && folded && __any_sync(0xffffffffu, opq < 0xff000000u) && (tid & 31) == 0) atomicExch(J.redo_flag, 1);
        } else              v_rows<NT, WM, false>(S, J, stg, sm, C, ys0 + g * STREAM_GROUP, nr, fix);
        advance();
    }
    if constexpr (INLINE) {
        if (wm_tma && (tid & 31) == 0) tma_store_wait_all(); // shared memory must outlive this lane's stores
    }
    if constexpr (FAST) return;
    if constexpr (NT > 0) {
        if (!switch_alpha) return;
        // phase 2: first non-opaque pixel this warp meets: materialise the per-pixel alpha lanes from the
        // chains entering this group (bit-identical to having carried them all along)
        VAcc<true> A[NT];
#pragma unroll
        for (int T = 0; T < NT; T++) {
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int i = 0; i < 6; i++) A[T].rgb[k][i] = S[T].rgb[k][i];
            const float s0 = sm.stage[rs].rec[T].seed0, s1 = sm.stage[rs].rec[T].seed1;
            A[T].al[0][0] = A[T].al[0][1] = make_float2(s0, s0);
            A[T].al[1][0] = A[T].al[1][1] = make_float2(s1, s1);
        }
        for (; g < ngroups; g++) {
            mbar_wait_a(bar32 + 8u * (uint32_t)rs, rph); // (already complete for the group that triggered the switch)
            v_rows<NT, WM, true>(A, J, sm.stage[rs], sm, C, ys0 + g * STREAM_GROUP, yend - ys0 - g * STREAM_GROUP, fix);
            advance();
        }
    }
}

// ---------------------------------------------------------------------------------
// k_stream_planar: the lean streaming resample for planar YCbCr sources (*image.YCbCr,
// 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0) and *image.Gray, one target per pass, local or wide per CTA.
//
// Same structure as k_stream<1,false,4>; what differs is the source: the producer lands the
// Y row and the chroma rows that belong to each of the group's 4 luma rows (nearest chroma
// sample, no interpolation), and the V lanes convert their 4 pixels to the 16-bit RGB that
// x/image's scaleX_YCbCr<ratio> feeds its filter -- the inlined color.YCbCr.RGBA() integer
// formula, per tap, before any filtering -- so the result is the reference's, not that of an
// 8-bit RGBA intermediate.  cropAndResize's 1:1 first pass stores uint8(c16 >> 8), which the
// second pass re-expands (two_stage).  Alpha is the constant 0xffff.  H2D is 1.5 bytes per
// pixel instead of 4 for a 4:2:0 JPEG.
// ---------------------------------------------------------------------------------
// NRGBA = true is the same kernel for *image.NRGBA (PNG with straight alpha): rows of 4 bytes per
// pixel land as in k_stream; every V lane premultiplies its 4 pixels at 16 bits exactly as
// scaleX_NRGBA does (a16 = A * 0x101; c16 = C * a16 / 0xff in uint32) and carries per-pixel alpha
// accumulators (the alpha lane is data here, never a constant).
template <bool NRGBA> struct PlanarStage;
template <> struct __align__(16) PlanarStage<false> {
    uint8_t y[STREAM_GROUP][STREAM_COLS];
    uint8_t cb[STREAM_GROUP][STREAM_COLS];
    uint8_t cr[STREAM_GROUP][STREAM_COLS];
    GroupRec rec[1];
};
template <> struct __align__(16) PlanarStage<true> {
    uint4 rows[STREAM_GROUP][STREAM_THREADS];
    GroupRec rec[1];
};
template <bool NRGBA, int STAGES> struct __align__(128) PlanarSmem {
    PlanarStage<NRGBA> stage[STAGES];
    float4 xbuf[1][STREAM_XBUF];
    XTab<1> xt;
    uint64_t full[STAGES], empty[STAGES];
    XInfo xi[2];
    // fused watermark conversion (draw.Draw(Src) of the planar / NRGBA source into the RGBA8 frame): destination of the
    // band's first own row at the tile's first column, and how many rows the band owns (0: no watermark on this job)
    uint8_t *wm_row0;
    int32_t wm_stride, wm_rows;
};
enum { PLANAR_STAGES = 16 / STREAM_GROUP, PLANAR_CTAS = 4, NRGBA_CTAS = 3 };

// two 16-bit samples -> fp32 pair via the 2^23 mantissa trick (exact for values < 2^23)
__device__ __forceinline__ float2 u16x2_f32(uint32_t a, uint32_t b)
{
    return __fadd2_rn(make_float2(__uint_as_float(0x4B000000u | a), __uint_as_float(0x4B000000u | b)),
                      make_float2(-8388608.0f, -8388608.0f));
}

// 16-bit sample >> SH clamped to [0, 0xffff >> (SH - 8)]: color.YCbCr.RGBA()'s "(x >> 8) clamped to 16 bits" for SH = 8;
// for the crop stage, which keeps uint8(c16 >> 8), SH = 16 gives that byte directly (the two clamps commute with the shift).
template <int SH> __device__ __forceinline__ uint32_t ycc_chan(int v)
{
    return (uint32_t)min(max(v >> SH, 0), SH == 8 ? 0xffff : 0xff);
}

// The V warps' row loop of k_stream_planar.  KIND 0: one chroma sample per pixel in the row (4:4:4, 4:4:0) -- or,
// with NRGBA, no chroma at all; 1: one per pixel pair (4:2:2, 4:2:0); 2: *image.Gray.  TWO: cropAndResize's 1:1 first
// pass stored uint8(c16 >> 8) into an *image.RGBA and scaleX_RGBA re-expands it (byte * 0x101).
// WM: this job also carries the watermark's full-frame conversion: a lane that owns its 4 columns stores, for every row
// the band owns, the 8-bit pixels draw.Draw(Src) would produce -- the high bytes of the 16-bit samples it has just
// computed (clamp and shift commute: uint8(YCbCrToRGB) == RGBA() >> 8; drawNRGBASrc's uint8((C * sa / 0xff) >> 8) is the
// premultiplied sample >> 8), so the source is read once for the resize and the watermark frame.
template <bool NRGBA, int KIND, bool TWO, bool WM, typename SM>
__device__ __forceinline__ void planar_vloop(SM &sm, int ngroups, int slot, int tid, const FixList &fix, int wm_cols)
{
    constexpr int STAGES = PLANAR_STAGES;
    VAcc<NRGBA> S;
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
        for (int i = 0; i < 6; i++) S.rgb[k][i] = make_float2(0.f, 0.f);
        if constexpr (NRGBA) S.al[k][0] = S.al[k][1] = make_float2(0.f, 0.f);
    }
    const bool local = sm.xi[0].local != 0;
    const int pslot = local ? tid : slot;
    int rs = 0;
    uint32_t rph = 0;
    for (int g = 0; g < ngroups; g++) {
        mbar_wait(&sm.full[rs], rph);
        const auto &stg = sm.stage[rs];
#pragma unroll 2
        for (int k = 0; k < STREAM_GROUP; k++) {
            uint32_t c16[12];
            [[maybe_unused]] uint32_t a16[4];
            if constexpr (NRGBA) {
                // ---- 4 straight-alpha pixels -> 16-bit premultiplied exactly as x/image scaleX_NRGBA
                const uint4 q4 = stg.rows[k][slot]; // (rows past the band's end: stale ring contents, weight 0)
                const uint32_t q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t pa = (q[j] >> 24) * 0x101u;
                    uint32_t r = div255((q[j] & 0xff) * pa), gg = div255(((q[j] >> 8) & 0xff) * pa), b = div255(((q[j] >> 16) & 0xff) * pa);
                    if constexpr (TWO) { r = (r >> 8) * 0x101u; gg = (gg >> 8) * 0x101u; b = (b >> 8) * 0x101u; }
                    c16[3 * j] = r; c16[3 * j + 1] = gg; c16[3 * j + 2] = b;
                    a16[j] = pa; // (pa >> 8) * 0x101 == pa: the crop stage keeps alpha as it is
                }
            } else if constexpr (KIND == 2) {
                // ---- *image.Gray: scaleX_Gray feeds y16 = Y * 0x101 on all three channels; the crop stage changes nothing
                // ((Y * 0x101 >> 8) * 0x101 == Y * 0x101)
                const uint32_t y4 = *reinterpret_cast<const uint32_t *>(&stg.y[k][slot * 4]);
#pragma unroll
                for (int j = 0; j < 4; j++) c16[3 * j] = c16[3 * j + 1] = c16[3 * j + 2] = ((y4 >> (8 * j)) & 0xff) * 0x101u;
            } else {
                // ---- 4 pixels -> 16-bit RGB exactly as color.YCbCr.RGBA() (x/image scaleX_YCbCr*): the chroma terms
                // once per chroma sample, then per pixel yy1 + term, >> 8, clamp
                constexpr int NC = KIND == 1 ? 2 : 4; // chroma samples under this thread's 4 pixels
                constexpr int SH = TWO ? 16 : 8;
                const uint32_t y4 = *reinterpret_cast<const uint32_t *>(&stg.y[k][slot * 4]);
                uint32_t cb4, cr4;
                if constexpr (KIND == 1) {
                    cb4 = *reinterpret_cast<const uint16_t *>(&stg.cb[k][slot * 2]);
                    cr4 = *reinterpret_cast<const uint16_t *>(&stg.cr[k][slot * 2]);
                } else {
                    cb4 = *reinterpret_cast<const uint32_t *>(&stg.cb[k][slot * 4]);
                    cr4 = *reinterpret_cast<const uint32_t *>(&stg.cr[k][slot * 4]);
                }
                int tr[NC], tg[NC], tb[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const int cb1 = (int)((cb4 >> (8 * c)) & 0xff) - 128, cr1 = (int)((cr4 >> (8 * c)) & 0xff) - 128;
                    tr[c] = 91881 * cr1;
                    tg[c] = -22554 * cb1 - 46802 * cr1;
                    tb[c] = 116130 * cb1;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int yy1 = (int)((y4 >> (8 * j)) & 0xff) * 0x10101;
                    const int c = KIND == 1 ? j >> 1 : j;
                    uint32_t r = ycc_chan<SH>(yy1 + tr[c]), gg = ycc_chan<SH>(yy1 + tg[c]), b = ycc_chan<SH>(yy1 + tb[c]);
                    if constexpr (TWO) { r *= 0x101u; gg *= 0x101u; b *= 0x101u; }
                    c16[3 * j] = r; c16[3 * j + 1] = gg; c16[3 * j + 2] = b;
                }
            }
            if constexpr (WM) {
                const int row = g * STREAM_GROUP + k;
                if (wm_cols > 0 && row < sm.wm_rows) { // wm_cols: how many of this lane's 4 columns it owns and the image has
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t al = 0xffu;
                        if constexpr (NRGBA) al = a16[j] >> 8;
                        o[j] = (c16[3 * j] >> 8) | ((c16[3 * j + 1] >> 8) << 8) | ((c16[3 * j + 2] >> 8) << 16) | (al << 24);
                    }
                    uint8_t *d = sm.wm_row0 + (size_t)row * sm.wm_stride + (size_t)slot * 16;
                    if (wm_cols == 4) *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
                    else
                        for (int j = 0; j < wm_cols; j++) reinterpret_cast<uint32_t *>(d)[j] = o[j];
                }
            }
            float2 vp[6];
#pragma unroll
            for (int i = 0; i < 6; i++) vp[i] = u16x2_f32(c16[2 * i], c16[2 * i + 1]);
            const float4 r = *reinterpret_cast<const float4 *>(&stg.rec[0].row[k]);
            const float2 w00 = make_float2(r.x, r.x), w11 = make_float2(r.y, r.y);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                S.rgb[0][i] = __ffma2_rn(vp[i], w00, S.rgb[0][i]);
                S.rgb[1][i] = __ffma2_rn(vp[i], w11, S.rgb[1][i]);
            }
            if constexpr (NRGBA) {
                const float2 va[2] = {u16x2_f32(a16[0], a16[1]), u16x2_f32(a16[2], a16[3])};
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    S.al[0][i] = __ffma2_rn(va[i], w00, S.al[0][i]);
                    S.al[1][i] = __ffma2_rn(va[i], w11, S.al[1][i]);
                }
            }
            const int e = stg.rec[0].emit[k];
            if (e >= 0) { // CTA-uniform
                park_emit<NRGBA>(S, e, r, sm.xbuf[0], pslot);
                if (local) __syncwarp(); else vwarps_bar();
                xcached(sm, 0, e >> 1, tid, fix);
                if (local) __syncwarp(); else vwarps_bar();
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&sm.empty[rs]);
        if (++rs == STAGES) { rs = 0; rph ^= 1; }
    }
}

template <bool NRGBA>
__global__ void __launch_bounds__(STREAM_CTA, (NRGBA ? NRGBA_CTAS : PLANAR_CTAS))
k_stream_planar(const StreamJob *__restrict__ jobs, const StreamItem *__restrict__ items, FixList fix)
{
    using Smem = PlanarSmem<NRGBA, PLANAR_STAGES>;
    constexpr int STAGES = PLANAR_STAGES;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t sm32 = smem_u32(smem_raw); // kept opaque: see k_stream (no S2R SR_CgaCtaId per group)
    asm volatile("" : "+r"(sm32));
    Smem *smp = reinterpret_cast<Smem *>(__cvta_shared_to_generic(sm32));
    __builtin_assume(__isShared(smp));
    Smem &sm = *smp;

    const StreamItem it = items[blockIdx.x];
    const StreamJob &J = jobs[it.job];
    const int tile = it.tile, band = it.band;
    const int W = J.src.w;
    const int cx0 = tile * J.tile_w;
    const int ys0 = __ldg(J.band_y + band);
    const int yend = __ldg(J.band_yend + band);
    const int warp = threadIdx.x >> 5;
    const int ngroups = (yend - ys0 + STREAM_GROUP - 1) / STREAM_GROUP;
    const int tid = (int)threadIdx.x;
    const int ws = J.warp_stride;
    const int slot = (warp & 3) * (ws >> 2) + (tid & 31);
    const int layout = J.src.layout;
    const bool sub_x = layout == L_YCBCR422 || layout == L_YCBCR420; // chroma at half horizontal resolution
    const bool sub_y = layout == L_YCBCR420 || layout == L_YCBCR440; // ... vertical
    const bool gray = layout == L_GRAY8;                              // *image.Gray: one plane, r = g = b = Y * 0x101
    const int ncols = min(J.slab_cols, W - cx0);
    const uint32_t y_bytes = (uint32_t)(((NRGBA ? ncols * 4 : ncols) + 15) & ~15);
    const uint32_t c_bytes = (uint32_t)(((sub_x ? (ncols + 1) >> 1 : ncols) + 15) & ~15);

    const int ys1 = __ldg(J.band_y + band + 1);
    const bool has_wm = J.has_wm != 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], STREAM_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.wm_row0 = has_wm ? J.wm.dst + (size_t)ys0 * J.wm.dst_stride + (size_t)cx0 * 4 : nullptr;
        sm.wm_stride = has_wm ? J.wm.dst_stride : 0;
        sm.wm_rows = has_wm ? ys1 - ys0 : 0;
    }
    if constexpr (NRGBA) {
        // slab columns past the image edge are never written by TMA: park transparent black there (finite, feeds nothing)
        if (tid < STREAM_THREADS && cx0 + slot * STREAM_PX >= W) {
            for (int s = 0; s < STAGES; s++)
                for (int k = 0; k < STREAM_GROUP; k++) sm.stage[s].rows[k][slot] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // horizontal-pass table (cached form; the engine only sends jobs whose tiles have one)
    if (tid < STREAM_THREADS) {
        for (int e = tid; e < STREAM_XBUF; e += STREAM_THREADS) sm.xbuf[0][e] = make_float4(0.f, 0.f, 0.f, 0.f);
        xtab_fill(sm.xt, sm.xi[0], 0, J.t[0], tile, cx0, ws, tid, __ldg(J.t[0].tile_parts + tile), J.t[0].fix_d);
    }
    __syncthreads();

    if (warp == STREAM_THREADS / 32) {
        // ===== producer =====
        if ((tid & 31) != 0) return;
        if constexpr (NRGBA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint8_t *grec = (const uint8_t *)(J.grec + (size_t)__ldg(J.band_grec_off + band));
        const int ccx0 = sub_x ? cx0 >> 1 : cx0;
        auto refill = [&](int group, int stage) {
            const int y0 = ys0 + group * STREAM_GROUP;
            const int nr = min(STREAM_GROUP, yend - y0);
            auto &st = sm.stage[stage];
            if constexpr (NRGBA) {
                mbar_arrive_expect_tx(&sm.full[stage], y_bytes * (uint32_t)nr + (uint32_t)sizeof(GroupRec));
                for (int k = 0; k < nr; k++)
                    tma_load_1d(&st.rows[k][0], J.src.p0 + (size_t)(y0 + k) * J.src.s0 + (size_t)cx0 * 4, y_bytes, &sm.full[stage]);
            } else {
                mbar_arrive_expect_tx(&sm.full[stage], (y_bytes + (gray ? 0u : 2 * c_bytes)) * (uint32_t)nr + (uint32_t)sizeof(GroupRec));
                for (int k = 0; k < nr; k++) {
                    const int y = y0 + k, cy = sub_y ? y >> 1 : y;
                    tma_load_1d(&st.y[k][0], J.src.p0 + (size_t)y * J.src.s0 + cx0, y_bytes, &sm.full[stage]);
                    if (gray) continue;
                    tma_load_1d(&st.cb[k][0], J.src.p1 + (size_t)cy * J.src.s1 + ccx0, c_bytes, &sm.full[stage]);
                    tma_load_1d(&st.cr[k][0], J.src.p2 + (size_t)cy * J.src.s2 + ccx0, c_bytes, &sm.full[stage]);
                }
            }
            tma_load_1d(&st.rec[0], grec + (size_t)group * sizeof(GroupRec), (uint32_t)sizeof(GroupRec), &sm.full[stage]);
        };
        for (int g = 0; g < min(STAGES, ngroups); g++) refill(g, g);
        int sp = 0;
        uint32_t php = 0;
        for (int g = 1; g < ngroups; g++) {
            if (g - 1 + STAGES < ngroups) {
                mbar_wait(&sm.empty[sp], php);
                refill(g - 1 + STAGES, sp);
            }
            if (++sp == STAGES) { sp = 0; php ^= 1; }
        }
        return;
    }

    // ===== V warps: one specialisation of the row loop per (source kind, crop stage), chosen per CTA =====
    const bool two_stage = J.t[0].two_stage != 0;
    // the watermark frame: columns this lane owns (the overlap of consecutive warps belongs to the right-hand one, as in
    // k_stream) and the image has; the engine attaches a watermark to single-stage (resize) jobs only
    int wm_cols = 0;
    if (has_wm) {
        const int c = cx0 + slot * STREAM_PX;
        const bool own = c < min(W, cx0 + J.tile_w) && (warp == 3 || (tid & 31) * STREAM_PX < ws);
        if (own) wm_cols = min(STREAM_PX, min(W, cx0 + J.tile_w) - c);
    }
    if constexpr (NRGBA) {
        if (two_stage)   planar_vloop<true, 0, true, false>(sm, ngroups, slot, tid, fix, 0);
        else if (has_wm) planar_vloop<true, 0, false, true>(sm, ngroups, slot, tid, fix, wm_cols);
        else             planar_vloop<true, 0, false, false>(sm, ngroups, slot, tid, fix, 0);
    } else if (gray) {
        if (two_stage)   planar_vloop<false, 2, true, false>(sm, ngroups, slot, tid, fix, 0);
        else if (has_wm) planar_vloop<false, 2, false, true>(sm, ngroups, slot, tid, fix, wm_cols);
        else             planar_vloop<false, 2, false, false>(sm, ngroups, slot, tid, fix, 0);
    } else if (sub_x) {
        if (two_stage)   planar_vloop<false, 1, true, false>(sm, ngroups, slot, tid, fix, 0);
        else if (has_wm) planar_vloop<false, 1, false, true>(sm, ngroups, slot, tid, fix, wm_cols);
        else             planar_vloop<false, 1, false, false>(sm, ngroups, slot, tid, fix, 0);
    } else {
        if (two_stage)   planar_vloop<false, 0, true, false>(sm, ngroups, slot, tid, fix, 0);
        else if (has_wm) planar_vloop<false, 0, false, true>(sm, ngroups, slot, tid, fix, wm_cols);
        else             planar_vloop<false, 0, false, false>(sm, ngroups, slot, tid, fix, 0);
    }
}

// ---------------------------------------------------------------------------------
// k_direct: small-support fp32 resample, one thread per output pixel.
//
// For every contributing source row the horizontal sum (fmaf, taps in order), then its fmaf into the vertical sum:
// ny * nx taps read through L1 (neighbouring threads share most of them).  RGBA8 bytes go to fp32 with the 2^23 trick
// and the vertical weights carry the 0x101; every other layout goes through sample16() (16-bit samples, unscaled
// weights).  |T_fp32 - T_exact| obeys the same bound as the streaming kernels (plan.cpp certified_fix_d with one thread
// per output: every rounding is a round-to-nearest of a partial sum below 2^16 in final units), so the same ambiguity
// window flags the bytes k_exact_fix re-evaluates in float64.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_direct(const DirectJob *__restrict__ jobs, const DirectItem *__restrict__ items, FixList fix)
{
    const DirectItem it = items[blockIdx.x];
    const DirectJob &J = jobs[it.job];
    const int ox = it.tile_x * 32 + (threadIdx.x & 31);
    const int oy = it.tile_y * 8 + (threadIdx.x >> 5);
    if (ox >= J.dw || oy >= J.dh) return;
    const int kx0 = __ldg(J.xoff + ox), nx = __ldg(J.xoff + ox + 1) - kx0;
    const int ky0 = __ldg(J.yoff + oy), ny = __ldg(J.yoff + oy + 1) - ky0;
    const int x0 = __ldg(J.xfirst + ox) + J.rect_x, y0 = __ldg(J.yfirst + oy) + J.rect_y;
    const bool rgba8 = J.src.layout == L_RGBA8;
    const bool two = J.two_stage != 0;
    float2 rg = make_float2(0.f, 0.f), ba = make_float2(0.f, 0.f);
    for (int j = 0; j < ny; j++) {
        float2 srg = make_float2(0.f, 0.f), sba = make_float2(0.f, 0.f);
        if (rgba8) {
            const uint32_t *row = (const uint32_t *)(J.src.p0 + (size_t)(y0 + j) * J.src.s0) + x0;
            for (int k = 0; k < nx; k++) {
                uint32_t q = __ldg(row + k);
                if (two) q = clamp_to_alpha(q); // the crop stage stores min(c, a) >> 8 of byte * 0x101: the clamped byte
                const float w = __ldg(J.xw + kx0 + k);
                const float2 m = make_float2(-8388608.0f, -8388608.0f), ww = make_float2(w, w);
                srg = __ffma2_rn(__fadd2_rn(magic2<0, 1>(q, q), m), ww, srg);
                sba = __ffma2_rn(__fadd2_rn(magic2<2, 3>(q, q), m), ww, sba);
            }
        } else {
            for (int k = 0; k < nx; k++) {
                uint32_t p[4];
                sample16(J.src, x0 + k, y0 + j, p);
                if (two) to_cropped_rgba16(p);
                const float w = __ldg(J.xw + kx0 + k);
                const float2 ww = make_float2(w, w);
                srg = __ffma2_rn(u16x2_f32(p[0], p[1]), ww, srg);
                sba = __ffma2_rn(u16x2_f32(p[2], p[3]), ww, sba);
            }
        }
        const float wy = __ldg(J.yw + ky0 + j);
        const float2 wwy = make_float2(wy, wy);
        rg = __ffma2_rn(srg, wwy, rg);
        ba = __ffma2_rn(sba, wwy, ba);
    }
    xfinish(J, (uint32_t)J.fix_d, ox, oy, rg, ba, fix);
}

template <bool NRGBA>
static cudaError_t launch_stream_planar_t(const StreamJob *jobs, const StreamItem *items, int n_items, FixList fix, cudaStream_t st)
{
    using Smem = PlanarSmem<NRGBA, PLANAR_STAGES>;
    // the attribute belongs to the function handle of ONE device's primary context: keep the flag per device
    static bool configured[kMaxDevices] = {}; // per instantiation; benign race (idempotent attribute)
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_stream_planar<NRGBA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    k_stream_planar<NRGBA><<<n_items, STREAM_CTA, sizeof(Smem), st>>>(jobs, items, fix);
    return cudaGetLastError();
}

cudaError_t launch_stream_planar(const StreamJob *jobs, const StreamItem *items, int n_items, bool nrgba, FixList fix, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    return nrgba ? launch_stream_planar_t<true>(jobs, items, n_items, fix, st) : launch_stream_planar_t<false>(jobs, items, n_items, fix, st);
}

int stream_smem_bytes() { return (int)sizeof(StreamCfg<2>::Smem); }
// shared memory must not be what limits the residency the launch bounds ask for (227 KB per SM, 1 KB per CTA reserved)
static_assert(sizeof(StreamCfg<1, 4>::Smem) + 1024 <= 227 * 1024 / StreamCfg<1, 4>::CTAS_PER_SM, "lean k_stream: shared memory limits occupancy");
static_assert(sizeof(StreamCfg<1, 0>::Smem) + 1024 <= 227 * 1024 / STREAM_CTAS_1T, "k_stream<1>: shared memory limits occupancy");
static_assert(sizeof(StreamCfg<2, 0>::Smem) + 1024 <= 227 * 1024 / STREAM_CTAS_2T, "k_stream<2>: shared memory limits occupancy");
static_assert(sizeof(StreamCfg<2, 3>::Smem) + 1024 <= 227 * 1024 / STREAM_CTAS_FAST2, "lean fused k_stream: shared memory limits occupancy");
static_assert(sizeof(PlanarSmem<false, PLANAR_STAGES>) + 1024 <= 227 * 1024 / PLANAR_CTAS, "k_stream_planar: shared memory limits occupancy");
static_assert(sizeof(PlanarSmem<true, PLANAR_STAGES>) + 1024 <= 227 * 1024 / NRGBA_CTAS, "k_stream_planar<NRGBA>: shared memory limits occupancy");

template <int NT, bool WM, int LEAN>
static cudaError_t launch_stream_t(const StreamJob *jobs, const StreamItem *items, int n, FixList fix, cudaStream_t st)
{
    using Smem = typename StreamCfg<NT, LEAN>::Smem;
    static bool configured[kMaxDevices] = {}; // per instantiation AND device (the attribute is per primary context); benign race
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_stream<NT, WM, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    k_stream<NT, WM, LEAN><<<n, StreamCfg<NT, LEAN>::THREADS, sizeof(Smem), st>>>(jobs, items, fix);
    return cudaGetLastError();
}

// Rows must be bulk-copyable (16-byte aligned base and stride): the engine re-pitches
// anything else into its arena first.
cudaError_t launch_stream(const StreamJob *jobs, const StreamItem *items, int n_items, int max_targets,
                          bool any_wm, FixList fix, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    switch (max_targets) {
    case 0: return any_wm ? launch_stream_t<0, true, 0>(jobs, items, n_items, fix, st) : cudaSuccess;
    case 1: return any_wm ? launch_stream_t<1, true, 0>(jobs, items, n_items, fix, st)
                          : launch_stream_t<1, false, 0>(jobs, items, n_items, fix, st);
    default: return any_wm ? launch_stream_t<2, true, 0>(jobs, items, n_items, fix, st)
                           : launch_stream_t<2, false, 0>(jobs, items, n_items, fix, st);
    }
}

// A lean instantiation over items of jobs with StreamJob::fast_path == kind (1: local target, 2: wide target).
cudaError_t launch_stream_fast(const StreamJob *jobs, const StreamItem *items, int n_items, int kind, bool any_wm,
                               FixList fix, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    if (kind == 1)
        return any_wm ? launch_stream_t<1, true, 1>(jobs, items, n_items, fix, st)
                      : launch_stream_t<1, false, 1>(jobs, items, n_items, fix, st);
    if (kind == 4)
        return any_wm ? launch_stream_t<1, true, 4>(jobs, items, n_items, fix, st)
                      : launch_stream_t<1, false, 4>(jobs, items, n_items, fix, st);
    if (kind == 3)
        return any_wm ? launch_stream_t<2, true, 3>(jobs, items, n_items, fix, st)
                      : launch_stream_t<2, false, 3>(jobs, items, n_items, fix, st);
    return any_wm ? launch_stream_t<1, true, 2>(jobs, items, n_items, fix, st)
                  : launch_stream_t<1, false, 2>(jobs, items, n_items, fix, st);
}

cudaError_t launch_direct(const DirectJob *jobs, const DirectItem *items, int n_items, FixList fix, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_direct<<<n_items, 256, 0, st>>>(jobs, items, fix);
    return cudaGetLastError();
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
    // one resident wave each: the warps claim work from cursors.  Sized per device (benign race: idempotent)
    static int grid[kMaxDevices] = {}, grid_wide[kMaxDevices] = {};
    const int dev = current_device();
    if (grid[dev] == 0) {
        int sms = 148, per_sm = 4, per_sm_wide = 4;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_exact_fix, FIX_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_wide, k_exact_fix_wide, FIX_THREADS, 0) != cudaSuccess || per_sm_wide < 1) per_sm_wide = 4;
        grid_wide[dev] = sms * per_sm_wide;
        grid[dev] = sms * per_sm;
    }
    k_exact_fix<<<grid[dev], FIX_THREADS, 0, st>>>(jobs, n_jobs, fix);
    k_exact_fix_wide<<<grid_wide[dev], FIX_THREADS, 0, st>>>(jobs, fix);
    return cudaGetLastError();
}

cudaError_t launch_blend(const WatermarkD *wms, const BlendItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_blend<<<n_items, 256, 0, st>>>(wms, items);
    return cudaGetLastError();
}

cudaError_t launch_rgba_to_ycbcr420(const YccJob *jobs, const YccItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_rgba_to_ycbcr420<<<n_items, 256, 0, st>>>(jobs, items);
    return cudaGetLastError();
}

cudaError_t launch_blend_patch(const PatchJob *jobs, const BlendItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_blend_patch<<<n_items, 256, 0, st>>>(jobs, items);
    return cudaGetLastError();
}

cudaError_t launch_watermark(const WmJob *jobs, const WmItem *items, int n_items, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    k_watermark<<<n_items, 256, 0, st>>>(jobs, items);
    return cudaGetLastError();
}

} // namespace ipg
