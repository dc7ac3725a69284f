// plan.h -- host-side planning: per-axis weight tables exactly as x/image's
// newDistrib builds them (float64, same operation order), and the derived
// streaming plan (column tiles, row bands, per-row records) for k_stream.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "ipg_device.h"

namespace ipg {

// x/image v0.33.0 draw/scale.go newDistrib for BiLinear (Support 1, At(t)=1-t),
// CSR by output index.  Taps of one output are a contiguous run of source coords.
struct AxisPlan {
    int32_t dn = 0, sn = 0;
    int32_t max_taps = 0;
    std::vector<int32_t> off;    // [dn+1]
    std::vector<int32_t> first;  // [dn]
    std::vector<double> inv;     // [dn] 1/sum(w)
    std::vector<double> inv_ffff;// [dn] inv/0xffff
    std::vector<double> w;       // unnormalised weights
    bool contiguous = true;      // taps of each output are consecutive coords
};

std::shared_ptr<const AxisPlan> get_axis_plan(int dn, int sn); // cached, thread-safe

// Half-width of the ambiguity window of the fp32 streaming kernels, in 1/256 of a 16-bit step, for a target whose
// widest supports are taps_x / taps_y and whose horizontal pass splits an output over `parts` threads (plan.cpp).
int certified_fix_d(int taps_x, int taps_y, int parts);
// ... when the vertical pass runs in the integer-moment form: `vert_units` bounds its fp32 error (plan.cpp build_vint)
int certified_fix_d_vint(int taps_x, int parts, double vert_units);

struct StreamTargetSpec {
    int32_t rect_x, rect_y, rect_w, rect_h;
    int32_t dw, dh;
    bool operator==(const StreamTargetSpec &o) const
    {
        return rect_x == o.rect_x && rect_y == o.rect_y && rect_w == o.rect_w && rect_h == o.rect_h &&
               dw == o.dw && dh == o.dh;
    }
};

struct StreamTargetGeom {
    std::shared_ptr<const AxisPlan> ax, ay;
    std::vector<float> xw;            // normalised fp32 horizontal weights
    std::vector<int32_t> tile_ox;     // [n_tiles+1]
    bool local = false;               // horizontal pass per V warp (narrow support)
    std::vector<int32_t> warp_ox;     // local: [n_tiles*4+1]
    std::vector<int32_t> tile_parts;  // [n_tiles] 0 = generic, P >= 1 = cached with P threads per output
    std::vector<RowRec> rows;         // per-band records
    std::vector<int32_t> band_rec_off;// [n_bands]
    std::vector<int32_t> band_tend;   // [n_bands] one past last source row with a contribution
    std::vector<int32_t> band_oy;     // [n_bands+1] first output row owned by each band
    int32_t fix_d = 0;
    int32_t fix_d_vint = 0;           // certificate of the integer-moment vertical form (StreamGeom::vint_ok)
};

// Everything k_stream needs that depends only on geometry (not on pointers).
struct StreamGeom {
    int32_t W = 0, H = 0;
    int32_t n_targets = 0;
    bool has_wm = false;
    int32_t tile_w = 0, n_tiles = 0, n_bands = 0;
    int32_t warp_stride = STREAM_WARP_COLS, slab_cols = STREAM_COLS;
    std::vector<int32_t> band_y;    // [n_bands+1]
    std::vector<int32_t> band_yend; // [n_bands]
    StreamTargetGeom t[2];
    std::vector<GroupRec> grec;         // [groups of all bands][rec_slots]: what k_stream consumes
    int32_t rec_slots = 0;              // n_targets, + 1 when vint_ok: target 0's GroupRecI follows its fp32 record
    bool vint_ok = false;               // one wide 8-bit target whose vertical pass has the integer-moment form (GroupRecI)
    std::vector<int32_t> band_grec_off; // [n_bands] first group of each band
    std::vector<StreamItem> items;  // (tile, band) pairs that have work; job index left 0
    bool lean2_ok = false;          // two targets, [0] local and [1] wide, both in cached forms: the lean fused instantiation
    bool lean_regs_ok = false;      // ... and local with one output per lane group everywhere (taps in registers)
    bool lean_ok = false;           // one target, every tile with outputs in a cached horizontal-pass form:
                                    // the lean k_stream instantiation can run it (given opaque pixels)
};

// fp32 tables of the small-support direct kernel (k_direct) for one target: per axis the newDistrib taps with weights
// normalised to sum 1 (the vertical ones times the sample scale).  Cached like the stream geometry.
struct DirectGeom {
    std::shared_ptr<const AxisPlan> ax, ay;
    std::vector<float> xw, yw;
    int32_t fix_d = 0;
};
// nullptr unless the target is a vertical upscale (no streaming form exists: many output rows are open at once) with
// at most DIRECT_MAX_TAPS taps per output row... vertically, any horizontal support up to 200 taps.  With
// mild_downscales (IPG_DIRECT=2, experiments) downscales of <= DIRECT_MAX_TAPS taps per axis are taken too; measured
// on the B200 (r2, tools/size_sweep.py) the streaming kernel is faster there -- 12.9 vs 16.3 us at 1152x864, 15.8 vs
// 25.9 at 1632x1224, 17.6 vs 28.1 at 2048x1536 (resize + thumbnail per image) -- so the engine does not ask for them.
std::shared_ptr<const DirectGeom> get_direct_geom(const StreamTargetSpec &t, double sample_scale, bool mild_downscales = false);

// Returns nullptr when the geometry cannot stream (upscale in y, more than two
// output rows open at once, support wider than a slab, ...): caller uses k_exact.
// `sample_scale` maps source samples to the 16-bit scale (257 for 8-bit RGBA).
std::shared_ptr<const StreamGeom> get_stream_geom(int W, int H, const StreamTargetSpec *targets,
                                                  int n_targets, bool has_wm, int n_bands_hint,
                                                  double sample_scale);

} // namespace ipg
