// ipg_device.h -- structs shared by the host engine and the sm_100a kernels.
// Everything here is plain-old-data living in the per-batch parameter blob that
// the engine uploads once per launch sequence; pointers are device pointers.
#pragma once
#include <stdint.h>

namespace ipg {

enum Layout : int32_t {
    L_RGBA8 = 0, L_NRGBA8 = 1, L_GRAY8 = 2,
    L_YCBCR444 = 3, L_YCBCR422 = 4, L_YCBCR420 = 5, L_YCBCR440 = 6,
    L_RGBA64 = 7, L_NRGBA64 = 8, L_GRAY16 = 9,
};

struct SrcView {
    const uint8_t *p0, *p1, *p2; // device planes
    int32_t s0, s1, s2;          // strides (bytes)
    int32_t w, h;
    int32_t layout;
};

// One axis of x/image's distrib in reference (float64) form, CSR by output index.
struct AxisExact {
    const int32_t *off;     // [n+1] tap offsets
    const int32_t *first;   // [n]   first source coord (relative to the source rect)
    const double *inv;      // [n]   invTotalWeight
    const double *inv_ffff; // [n]   invTotalWeight / 0xffff
    const double *w;        // [off[n]] unnormalised tent weights
};

// A resample evaluated in fp64, reference operation order (whole image or the
// pixels a stream kernel flagged as ambiguous).
struct ExactJob {
    SrcView src;
    int32_t rect_x, rect_y;   // source rectangle origin
    int32_t two_stage;        // cropAndResize: samples are the 8-bit cropped RGBA
    int32_t dw, dh;
    int32_t dst_stride;
    uint8_t *dst;
    AxisExact ax, ay;
};

struct ExactItem { int32_t job; int32_t tile_x, tile_y; int32_t pad; };

// Output pixel flagged by a stream kernel for fp64 re-evaluation.  k_exact_fix gives a narrow-support
// pixel to one lane and re-queues the wide-support ones at the unused back end of the list for
// k_exact_fix_wide (a whole warp each).
struct FixEntry { int32_t job; int32_t x, y; };
struct FixList {
    FixEntry *entries;
    uint32_t *count;     // device counters, 4 words, zeroed per batch: [0] entries appended by the stream kernels
                         // (may exceed capacity: then every target is redone whole); [2] wide entries re-queued
                         // from the back; [1], [3] the fix kernels' work cursors
    uint32_t capacity;
};

// Streaming (vertical-first, fp32) resample plan ---------------------------------
// One record per source row a band walks: row contributes w_a to the lowest
// still-open output row and w_b to the next one; emit>=0: the lowest open row is
// complete after this source row and is output row `emit`.
struct RowRec { float wa, wb; int32_t emit; int32_t pad; };

// What the kernel actually consumes: the RowRecs of STREAM_GROUP consecutive source rows
// of one target, re-expressed per accumulator SET (the two open output rows live in two
// register sets that swap roles at every emitted row; the host resolves that parity), plus
// the alpha sums an all-opaque image would have accumulated (so the opaque fast path does
// no alpha arithmetic at all).  One block of n_targets GroupRecs per (band, group) rides
// into shared memory with the group's source rows in the same TMA transaction.
// Source rows per ring stage / TMA barrier phase.  The per-group fixed cost of a V warp (mbarrier wait, opacity vote,
// stage release, ring bookkeeping: ~50-60 issued instructions, ncu r2: 15-23 % of the kernel at 4 rows) is paid once per
// group.  IPG_GROUP=8 halves it with half as many, twice as large stages (same shared memory): measured on the B200
// (r2, tools/jobs/r2_variants.sh) each pass ALONE gains (resize + watermark 0.914 -> 0.939 of the HBM roofline, thumbnail
// 0.447 -> 0.503) but the merged launch the engine actually runs loses (25.2 -> 25.8 us per 12 MP image, bench frac
// 0.667 -> 0.641): two 8-row stages leave one stage of slack where four 4-row stages leave three.  Default stays 4.
#ifndef IPG_GROUP
#define IPG_GROUP 4
#endif
struct GroupRow { float w0, w1;    // weight of this source row for accumulator set 0 / 1
                  float sa0, sa1; };// fmaf(255, w, .) chains of set 0 / 1 after this row (before an emit clears it)
struct GroupRec {
    GroupRow row[IPG_GROUP];       // STREAM_GROUP rows
    int32_t emit[IPG_GROUP];       // -1, or (output row << 1 | set): the row completes that set
    float seed0, seed1;            // the two alpha chains entering the group
    int32_t pad[2];
};
static_assert(sizeof(GroupRec) % 16 == 0, "group records ride in TMA bulk copies (16-byte granular)");

// Integer-moment form of a wide target's vertical pass (plan.cpp build_vint; the 15:1 thumbnail).  Between the
// centres of two consecutive output rows ("segment") the tent weights of the two open output rows are LINEAR in the
// row index r, so a run of rows contributes to both through the two exact integer moments M0 = sum x and
// M1 = sum r * x of every byte column.  A V thread keeps both in ONE 32-bit word per channel, M0 in bits 0..11
// (<= 16 rows * 255) and M1 from bit 12 (<= 255 * 120), fed by one IDP.2A per channel and row with the 16-bit multiplier
// m = 1 + (r << 12): no byte -> fp32 unpack, no per-row weights, 12 accumulators instead of 24.  A "piece" is a segment,
// or half of one that is longer than 16 rows (scales up to 32:1).  When a piece ends the moments are flushed to fp32,
//     row = fl(bR * M1 + fl(aR * M0 + row)),   carry = fl(bL * M1 + fl(aL * M0 + carry)),   moments = 0
// and when it also ends its segment, `row` is the vertically filtered output row (parked for the horizontal pass), then
// row = carry, carry = 0.  At most one piece ends per group.  Same size as a GroupRec: it rides in the slot after the
// target's fp32 record (StreamJob::rec_slots), which the on-demand redo still uses.
struct GroupRecI {
    uint32_t m[IPG_GROUP];     // per source row: 1 + (r << 12), or 0 when the row feeds no output this band owns
    int32_t emit[IPG_GROUP];   // -1: nothing; -3: a piece ends after this row; -2: ... and its segment, whose row this band does not own;
                               // >= 0: ... and its segment: output row `emit` is complete
    float aR, bR, aL, bL;      // of the piece that ends in this group
    int32_t end_k, end_e;      // that end again, as the kernel takes it: row index within the group (IPG_GROUP - 1 when none: then
                               // every row is "before the end") and its emit word (-1 when none)
    uint8_t pad[sizeof(GroupRec) - 8 * IPG_GROUP - 24];
};
static_assert(sizeof(GroupRecI) == sizeof(GroupRec), "the integer-form record takes a GroupRec slot");
static_assert(IPG_GROUP == 4 || IPG_GROUP == 8, "the row loop indexes the next row with a power-of-two mask");

struct StreamTarget {
    uint8_t *dst;
    int32_t dst_stride;
    int32_t dw, dh;
    int32_t rect_x, rect_y;      // source rect origin (crop)
    int32_t two_stage;           // samples clamped to alpha (8-bit crop stage) first
    int32_t exact_job;           // ExactJob index used for fix-ups
    int32_t fix_d;               // ambiguity half-width, 1/256 of a 16-bit step
    int32_t fix_d_vint;          // ... of the integer-moment vertical form (StreamJob::vint)
    const int32_t *xoff;         // [dw+1]
    const int32_t *xfirst;       // [dw]   relative to rect_x
    const float *xw;             // normalised fp32 weights (sum 1)
    const int32_t *tile_ox;      // [n_tiles+1] output columns owned by each column tile
    int32_t local;               // 1: narrow support, each V warp runs its own horizontal pass
    const int32_t *warp_ox;      // local: [n_tiles*4+1] output columns owned by each (tile, warp)
    const int32_t *tile_parts;   // [n_tiles] horizontal-pass form of the tile, P | taps_per_thread << 8 | rounds << 16:
                                 // P == 0 = generic loops; P >= 1 = "cached": every output is split over P adjacent
                                 // V threads holding taps_per_thread interleaved taps each, and a thread takes
                                 // `rounds` outputs (taps_per_thread * rounds <= STREAM_XTAPS_TAB)
    const RowRec *rows;          // per-band records, concatenated
    const int32_t *band_rec_off; // [n_bands] first record of each band
    const int32_t *band_tend;    // [n_bands] one past the last source row that contributes
    const int32_t *band_oy;      // [n_bands+1] first output row owned by each band
};

struct GlyphD {
    int32_t x0, y0, x1, y1;
    int32_t mp_x, mp_y, mask_stride, pad;
    const uint8_t *mask; // device
};

struct WatermarkD {
    uint8_t *dst;
    int32_t dst_stride;
    int32_t n_glyphs;
    const GlyphD *glyphs;
    int32_t bx0, by0, bx1, by1;  // union of glyph rects
    uint32_t sr, sg, sb, sa;     // Uniform.RGBA(): c * 0x101
};

// 1: the merged lean instantiation (k_stream<1,WM,4>) carries the table forms of the horizontal pass inline (wide
// targets in the fp32 form, local targets with several outputs per lane); 0: it carries only the lane-per-output local
// pass and the integer-moment wide pass, and the engine routes every other lean job to the wide-target instantiation
// (k_stream<1,WM,2>, table forms behind a call) on the side stream: a smaller row loop for the instruction cache.
#ifndef IPG_LEAN4_TABLES
#define IPG_LEAN4_TABLES 1
#endif

// occupancy knobs (overridable for experiments: make EXTRA=-DIPG_CTAS_2T=2 ...)
#ifndef IPG_CTAS_2T
#define IPG_CTAS_2T 2
#endif
#ifndef IPG_STAGES_2T
#define IPG_STAGES_2T (32 / IPG_GROUP)
#endif
#ifndef IPG_CTAS_1T
#define IPG_CTAS_1T 3
#endif
#ifndef IPG_STAGES_1T
#define IPG_STAGES_1T (IPG_GROUP == 4 ? 5 : 3)
#endif
#ifndef IPG_CTAS_FAST
#define IPG_CTAS_FAST 4
#endif
#ifndef IPG_STAGES_FAST
#define IPG_STAGES_FAST (16 / IPG_GROUP)
#endif
#ifndef IPG_CTAS_FAST2
#define IPG_CTAS_FAST2 3
#endif
#ifndef IPG_CTAS_INL
#define IPG_CTAS_INL 5
#endif
#ifndef IPG_STAGES_INL
#define IPG_STAGES_INL (IPG_GROUP == 4 ? 3 : 2)
#endif
#ifndef IPG_STAGES_FAST2
#define IPG_STAGES_FAST2 (IPG_GROUP == 4 ? 3 : 2)
#endif

// k_stream CTA: 4 V warps + one producer warp (one elected lane drives the TMA ring: bulk
// loads of source rows + group records, bulk stores of the watermark copy straight out of
// the ring).  Each V warp covers 128 source columns (32 lanes x 4 px); consecutive warps
// start `warp_stride` <= 128 columns apart, so a warp sees the right halo of its own
// outputs and runs the horizontal pass of narrow-support ("local") targets on its own,
// one output per lane, with no cross-warp hand-off.  Wide-support targets (the 15:1
// thumbnail) park their rows CTA-wide and meet at a 128-thread named barrier instead.
enum {
    STREAM_THREADS = 128, STREAM_PX = 4, STREAM_WARP_COLS = 32 * STREAM_PX,
    STREAM_COLS = STREAM_THREADS * STREAM_PX, // widest slab (warp_stride == 128)
    STREAM_GROUP = IPG_GROUP,   // source rows per ring stage / TMA barrier phase (<= 16 KB)
    STREAM_PTHREADS = 32,       // producer warp
    STREAM_CTA = STREAM_THREADS + STREAM_PTHREADS,
    STREAM_XTAPS = 8,           // local targets: taps per thread kept in registers
    STREAM_XTAPS_TAB = 16,      // shared-memory table: taps per thread, all rounds together
    STREAM_XROUNDS = 4,         // ... and outputs per thread (rounds) in the table forms (mild downscales)
    STREAM_LOCAL_MAX_HALO = 16, // a target is local when its widest support - 1 fits this overlap
    STREAM_STAGES_2T = IPG_STAGES_2T, STREAM_CTAS_2T = IPG_CTAS_2T,   // ring depth / CTAs per SM, two-target instantiation
    STREAM_STAGES_1T = IPG_STAGES_1T, STREAM_CTAS_1T = IPG_CTAS_1T,   // ... otherwise
    STREAM_STAGES_FAST = IPG_STAGES_FAST, STREAM_CTAS_FAST = IPG_CTAS_FAST, // the lean single-target instantiations
    STREAM_STAGES_FAST2 = IPG_STAGES_FAST2, STREAM_CTAS_FAST2 = IPG_CTAS_FAST2, // the lean local + wide instantiation
    STREAM_STAGES_INL = IPG_STAGES_INL, STREAM_CTAS_INL = IPG_CTAS_INL,         // the lean single-target instantiations without a producer warp
};

struct StreamJob {
    SrcView src;
    int32_t n_targets;
    int32_t has_wm;
    int32_t tile_w;            // source columns owned per tile (multiple of 4)
    int32_t warp_stride;       // columns between the first columns of consecutive V warps (multiple of 4)
    int32_t slab_cols;         // 3 * warp_stride + 128: columns a CTA loads per row
    int32_t n_tiles, n_bands;
    int32_t check_premul;      // RGBA8 source, alpha unknown, a two_stage target exists
    int32_t fast_path;         // 1 / 2 / 3: a lean instantiation (local / wide / local + wide targets) runs this job;
                               // the general one only redoes it on demand
    int32_t rec_slots;         // GroupRec slots per group in grec (n_targets, + 1 when target 0's integer-moment record rides along)
    int32_t vint;              // 1: the lean kernels run target 0's vertical pass in the integer-moment form (GroupRecI in slot n_targets)
    int32_t *redo_flag;        // fast_path: raised by the lean kernel on a non-opaque pixel (nullptr: caller vouches for opacity)
    const int32_t *band_y;     // [n_bands+1] owned source rows of each band
    const int32_t *band_yend;  // [n_bands]   one past the last row the band must read
    const GroupRec *grec;      // [groups of all bands][rec_slots]
    const int32_t *band_grec_off; // [n_bands] first group of each band
    StreamTarget t[2];
    WatermarkD wm;
};

struct StreamItem { int32_t job; int16_t tile, band; };

// Small-support fp32 resample, one thread per output pixel (k_direct): vertical upscales, which cannot stream (many output
// rows open at once).  Same certificate and fix list as the streaming kernels.
enum { DIRECT_MAX_TAPS = 4 };
struct DirectJob {
    SrcView src;
    int32_t rect_x, rect_y;      // source rect origin (crop)
    int32_t two_stage;           // samples pass cropAndResize's 8-bit crop stage first
    int32_t dw, dh, dst_stride;
    uint8_t *dst;
    const int32_t *xoff, *xfirst; // [dw+1], [dw]
    const float *xw;              // normalised fp32 horizontal weights (sum 1)
    const int32_t *yoff, *yfirst; // [dh+1], [dh]
    const float *yw;              // normalised fp32 vertical weights times the sample scale (257 for 8-bit samples, 1 for 16-bit)
    int32_t exact_job, fix_d;
};
struct DirectItem { int32_t job; int32_t tile_x, tile_y; int32_t pad; }; // 32 x 8 output pixels

// Standalone convert/copy + watermark blend (any layout) -------------------------
struct WmJob {
    SrcView src;
    WatermarkD wm;
};
struct WmItem { int32_t job; int32_t row0; };  // WM_ROWS rows per item
// Patch-only watermark of an RGBA8 source (IPG_OPF_WATERMARK_PATCH_ONLY): the glyph union box alone is produced, read
// from the source, blended, written at (x - ox, y - oy) of wm.dst (a box-sized patch buffer, or the caller's device
// frame with ox = oy = 0).
struct PatchJob {
    SrcView src;
    WatermarkD wm;
    int32_t ox, oy;
    int32_t pad[2];
};
struct BlendItem { int32_t wm; int32_t tile_x, tile_y; }; // 32x8 px of a watermark's glyph box
// RGBA8 result -> planar YCbCr 4:2:0 exactly as Go's image/jpeg writer derives it (k_rgba_to_ycbcr420)
struct YccJob {
    const uint8_t *rgba;
    int32_t rgba_pitch;
    int32_t w, h;
    uint8_t *y, *cb, *cr;
    int32_t y_pitch, c_pitch;
    int32_t pad[2];
};
struct YccItem { int32_t job; int32_t tile_x, tile_y; int32_t pad; }; // 256 x 16 pixels
enum { WM_ROWS = 8 };

} // namespace ipg
