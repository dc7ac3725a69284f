// plan.cpp -- see plan.h.  Compiled with -ffp-contract=off: the tables must equal
// what Go/amd64 computes for golang.org/x/image v0.33.0 draw/scale.go newDistrib.
#include "plan.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace ipg {

// Plan cache with least-recently-used eviction: when full, the older half (by last use) goes; a long-running worker
// on a mixed-size stream keeps its working set instead of rebuilding everything after a clear.  Entries are shared_ptrs:
// whoever still uses an evicted plan keeps it alive (engine.cpp pins the plans of a batch until its blob is built).
template <typename K, typename V> class AgedCache {
public:
    explicit AgedCache(size_t cap) : cap_(cap) {}
    bool find(const K &k, V &out)
    {
        auto it = map_.find(k);
        if (it == map_.end()) return false;
        it->second.second = ++tick_;
        out = it->second.first;
        return true;
    }
    void put(const K &k, const V &v)
    {
        if (map_.size() >= cap_) {
            std::vector<uint64_t> ages;
            ages.reserve(map_.size());
            for (auto &kv : map_) ages.push_back(kv.second.second);
            std::nth_element(ages.begin(), ages.begin() + ages.size() / 2, ages.end());
            const uint64_t cut = ages[ages.size() / 2];
            for (auto it = map_.begin(); it != map_.end();) it = it->second.second < cut ? map_.erase(it) : std::next(it);
        }
        map_[k] = {v, ++tick_};
    }
    size_t size() const { return map_.size(); }

private:
    size_t cap_;
    uint64_t tick_ = 0;
    std::map<K, std::pair<V, uint64_t>> map_;
};

static std::shared_ptr<const AxisPlan> build_axis(int32_t dn, int32_t sn)
{
    auto p = std::make_shared<AxisPlan>();
    p->dn = dn;
    p->sn = sn;
    p->off.assign((size_t)dn + 1, 0);
    p->first.assign((size_t)dn, 0);
    p->inv.assign((size_t)dn, 0.0);
    p->inv_ffff.assign((size_t)dn, 0.0);

    // BiLinear: Support = 1, At(t) = 1 - t.  When shrinking, the support is
    // widened by the scale and the kernel argument shrunk by it.
    const double scale = (double)sn / (double)dn;
    double half_width = 1.0, arg_scale = 1.0;
    if (scale > 1) {
        half_width *= scale;
        arg_scale = 1 / scale;
    }
    for (int32_t x = 0; x < dn; x++) {
        const double center = ((double)x + 0.5) * scale - 0.5;
        int32_t i = (int32_t)std::floor(center - half_width);
        if (i < 0) i = 0;
        int32_t j = (int32_t)std::ceil(center + half_width);
        if (j > sn) {
            j = sn;
            if (j < i) j = i;
        }
        double total = 0.0;
        int32_t kept_first = -1, kept_last = -1, kept = 0;
        for (int32_t coord = i; coord < j; coord++) {
            const double t = std::fabs((center - (double)coord) * arg_scale);
            if (t >= 1.0) continue;
            const double wt = 1 - t;
            if (wt == 0) continue;
            total += wt;
            p->w.push_back(wt);
            if (kept_first < 0) kept_first = coord;
            kept_last = coord;
            kept++;
        }
        if (kept > 0 && kept_last - kept_first + 1 != kept) p->contiguous = false;
        p->first[x] = kept_first < 0 ? 0 : kept_first;
        p->off[(size_t)x + 1] = (int32_t)p->w.size();
        total = 1 / total;
        p->inv[x] = total;
        p->inv_ffff[x] = total / 0xffff;
        p->max_taps = std::max(p->max_taps, kept);
    }
    return p;
}

std::shared_ptr<const AxisPlan> get_axis_plan(int dn, int sn)
{
    static std::mutex mu;
    static AgedCache<std::pair<int, int>, std::shared_ptr<const AxisPlan>> cache(4096);
    {
        std::lock_guard<std::mutex> lk(mu);
        std::shared_ptr<const AxisPlan> hit;
        if (cache.find({dn, sn}, hit)) return hit;
    }
    auto p = build_axis(dn, sn);
    std::lock_guard<std::mutex> lk(mu);
    cache.put({dn, sn}, p);
    return p;
}

// ---------------------------------------------------------------------------------

// The fp32 certificate (DESIGN.md 2.1 derives it term by term).  Units: 1/256 of a 16-bit step, i.e. the LSB of
// T = floor(v * 256 + 128), the 16.8 fixed-point value whose bits 16.. are the output byte.  For one output channel
//   |T_fp32 - T_exact| <= 2                     the two fp32-rounded weights on every term (relative 2^-24 each, value < 2^16)
//                        + 0.5 * taps_y          one round-to-nearest per vertical fmaf, partial sums < 2^16 (half ulp = 2^-9)
//                        + 0.5 * taps_x          one per horizontal fmaf
//                        + log2(P)               the xor-butterfly over the P threads sharing an output (<= 1 per level)
//                        + 0.5                   the rounding of fmaf(v, 256, 128) itself (values in [2^23, 2^24): ulp 1)
//                        + 1                     floor() against the exact, unfloored value
//                        (+ < 0.001              the float64 reference's own rounding)
// so D = ceil((taps_x + taps_y) / 2) + log2(P) + 4 covers it, and one more unit of margin is added on top.
// A byte can differ from the float64 result only when a multiple of 65536 lies within D of T: those are flagged.
int certified_fix_d(int taps_x, int taps_y, int parts)
{
    int lg = 0;
    while ((1 << lg) < parts) lg++;
    return (taps_x + taps_y + 1) / 2 + lg + 5;
}

// The integer-moment vertical form (build_vint) replaces the 0.5 * taps_y of the vertical fmaf chain and the vertical
// weight's unit by `vert_units`, the bound build_vint derives for "moments -> fp32 row"; the horizontal terms stay:
//   1 (fp32 horizontal weight) + 0.5 * taps_x + log2(P) + 0.5 (fmaf(v, 256, 128)) + 1 (floor) and one unit of margin.
int certified_fix_d_vint(int taps_x, int parts, double vert_units)
{
    int lg = 0;
    while ((1 << lg) < parts) lg++;
    return (int)std::ceil(vert_units + 0.5 * taps_x + 2.5) + lg + 1;
}

// Integer-moment records of a wide 8-bit target (GroupRecI, ipg_device.h).  Segment a = source rows [B[a], B[a+1]) with
// B[a] = ceil(centre of output row a): there the tent of row a falls (weight (1 - (o - c_a) / s) - r / s at r rows past
// an origin o) and the tent of row a + 1 rises ((1 - (c_{a+1} - o) / s) + r / s); rows above B[0] only raise row 0's.
// A segment of more than 16 rows (scales 16:1 ... 32:1) is cut in two pieces of <= 16 rows, each with its own origin,
// so that M0 keeps its 12 bits: every piece is flushed to fp32 on its own.  The model is CHECKED against the
// newDistrib table (every owned output row, every source row of the band) and the form is refused when it deviates by
// more than 1e-10 of a weight or when two pieces could end in one group.  `vert_units` returns the bound of
// |fp32 row - exact row| in 1/256 of a 16-bit step.  Per flush, with M0 <= 255 n and M1 <= 255 sum(r) over the piece's
// rows, the running values bounded by 65535 (1 + 1e-6) (true partial sums of a convex combination) and every rounding
// at most 2^-24 of the magnitude it rounds:
//   row   = fl(bR * M1 + fl(aR * M0 + row))   (|aR| M0 + |bR| M1) coefficients + (|aR| M0 + |row|) + |row'|
//   carry = fl(bL * M1 + fl(aL * M0 + carry)) (|aL| M0 + |bL| M1) coefficients + (|aL| M0 + |carry|) + |carry'|
// summed over the flushes of segment a - 1 (carry) and of segment a (row) for output row a.
static bool build_vint(const StreamGeom &g, const StreamTargetSpec &s, double sample_scale, std::vector<GroupRecI> &out,
                       double &vert_units)
{
    const StreamTargetGeom &t = g.t[0];
    const AxisPlan &ay = *t.ay;
    const double scale = (double)s.rect_h / (double)s.dh;
    if (sample_scale != 257.0 || !(scale >= (double)STREAM_GROUP + 1.0) || !(scale <= 32.0)) return false;
    const double arg = 1 / scale;
    std::vector<double> cen((size_t)s.dh);
    std::vector<int32_t> B((size_t)s.dh + 1);
    for (int32_t a = 0; a < s.dh; a++) {
        cen[a] = ((double)a + 0.5) * scale - 0.5;
        B[a] = std::min(std::max((int32_t)std::ceil(cen[a]), 0), s.rect_h);
    }
    B[s.dh] = s.rect_h;
    auto last = [&](int32_t oy) { return ay.first[oy] + (ay.off[oy + 1] - ay.off[oy]) - 1; };
    const double vmax = 65535.0 * (1 + 1e-6);
    double worst = 0.0;
    out.clear();
    for (int b = 0; b < g.n_bands; b++) {
        const int32_t Y0 = g.band_y[b];
        const int32_t ng = (g.band_yend[b] - Y0 + STREAM_GROUP - 1) / STREAM_GROUP;
        const int32_t oyA = t.band_oy[b], oyB = t.band_oy[b + 1], tend = t.band_tend[b];
        const size_t base = out.size();
        out.resize(base + (size_t)ng);
        // what the kernel will compute, in double, per piece: checked against the table below
        struct Piece { int32_t sg; double aR, bR, aL, bL; };
        std::vector<Piece> pieces;
        std::vector<int32_t> row_piece((size_t)ng * STREAM_GROUP, -1), row_r((size_t)ng * STREAM_GROUP, 0);
        int32_t cur_first = -1, cur_n = 0;   // rows (relative to Y0) of the piece being walked: first, count
        int32_t cur_o = INT32_MIN;           // ... and its origin
        int32_t seg_hint = std::max(oyA - 1, -1); // the segments are met in order: no search per row
        int64_t cur_sr = 0;
        double L_prev = 0.0;                 // error weight of the carry that entered the current segment
        double L_cur = 0.0, R_cur = 0.0;     // ... accumulated by this segment's flushes: into the next row's carry / into its own row
        double cy_mag = 0.0, nx_mag = 0.0;   // bounds of the running row / carry values
        for (int32_t gi = 0; gi < ng; gi++) {
            GroupRecI &G = out[base + (size_t)gi];
            memset(&G, 0, sizeof G);
            G.end_k = STREAM_GROUP - 1;
            G.end_e = -1;
            int ends = 0;
            for (int k = 0; k < STREAM_GROUP; k++) {
                G.emit[k] = -1;
                const int32_t ys = Y0 + gi * STREAM_GROUP + k, yr = ys - s.rect_y;
                if (oyB <= oyA || ys >= tend || yr < ay.first[oyA]) continue;
                // the segment of this row: the largest a with B[a] <= yr, -1 above B[0]
                while (seg_hint + 1 < s.dh && B[seg_hint + 1] <= yr) seg_hint++;
                const int32_t sg = seg_hint;
                if (sg < oyA - 1 || sg > oyB - 1 || (sg >= 0 && B[sg] > yr)) return false;
                const int32_t lo = sg >= 0 ? B[sg] : 0, hi = B[sg + 1], n = hi - lo;
                if (n < 1 || n > 32) return false;
                const int32_t half = n > 16 ? (n + 1) / 2 : n;       // rows of the first piece
                const int32_t o = yr < lo + half ? lo : lo + half, r = yr - o;
                const bool seg_end = yr == hi - 1, piece_end = seg_end || yr == lo + half - 1;
                if (r < 0 || r > 15) return false;
                if (o != cur_o) {
                    if (cur_n != 0) return false; // the previous piece was never flushed
                    cur_o = o;
                }
                if (cur_n == 0) cur_first = ys - Y0;
                G.m[k] = 1u + ((uint32_t)r << 12);
                row_r[(size_t)(ys - Y0)] = r;
                cur_n++;
                cur_sr += r;
                if (cur_n > 16) return false;
                if (ys == tend - 1 && !seg_end) return false; // the band's last row closes its last segment
                if (!piece_end) continue;
                if (++ends > 1) return false;
                Piece e{sg, 0, 0, 0, 0};
                if (sg >= oyA) { // falling half of output row sg
                    e.aR = (1 - ((double)o - cen[sg]) * arg) * ay.inv[sg] * sample_scale;
                    e.bR = -arg * ay.inv[sg] * sample_scale;
                }
                if (sg + 1 < oyB) { // rising half of output row sg + 1
                    e.aL = (1 - (cen[sg + 1] - (double)o) * arg) * ay.inv[sg + 1] * sample_scale;
                    e.bL = arg * ay.inv[sg + 1] * sample_scale;
                }
                G.aR = (float)e.aR; G.bR = (float)e.bR; G.aL = (float)e.aL; G.bL = (float)e.bL;
                G.emit[k] = !seg_end ? -3 : sg >= oyA ? sg : -2;
                G.end_k = k;
                G.end_e = G.emit[k];
                for (int32_t rr = cur_first; rr < cur_first + cur_n; rr++) row_piece[(size_t)rr] = (int32_t)pieces.size();
                pieces.push_back(e);
                const double M0 = 255.0 * (double)cur_n, M1 = 255.0 * (double)cur_sr;
                const double mR = std::fabs(e.aR) * M0 + std::fabs(e.bR) * M1, mL = std::fabs(e.aL) * M0 + std::fabs(e.bL) * M1;
                R_cur += mR + (std::fabs(e.aR) * M0 + std::min(cy_mag, vmax)) + vmax;
                L_cur += mL + (std::fabs(e.aL) * M0 + std::min(nx_mag, vmax)) + std::min(nx_mag + mL, vmax);
                cy_mag += mR;
                nx_mag += mL;
                cur_n = 0;
                cur_sr = 0;
                if (seg_end) {
                    if (sg >= oyA) worst = std::max(worst, (L_prev + R_cur) * 256.0 / 16777216.0);
                    L_prev = L_cur;
                    cy_mag = nx_mag;
                    L_cur = R_cur = nx_mag = 0.0;
                }
            }
        }
        if (cur_n != 0) return false;
        // the model against the table: every owned output row over every source row the band walks
        for (int32_t a = oyA; a < oyB; a++) {
            // (rows outside output a's support and outside segments a - 1 and a have neither a tap nor a model weight)
            const int32_t r0 = std::min(ay.first[a], a > 0 ? B[a - 1] : 0), r1 = std::max(last(a), B[a + 1] - 1);
            for (int32_t ys = std::max(Y0, r0 + s.rect_y); ys < std::min(Y0 + ng * STREAM_GROUP, r1 + s.rect_y + 1); ys++) {
                const int32_t yr = ys - s.rect_y;
                double table = 0.0, model = 0.0;
                if (yr >= ay.first[a] && yr <= last(a)) table = ay.w[(size_t)ay.off[a] + (size_t)(yr - ay.first[a])] * ay.inv[a] * sample_scale;
                const int32_t pi = row_piece[(size_t)(ys - Y0)], r = row_r[(size_t)(ys - Y0)];
                if (pi >= 0) {
                    const Piece &e = pieces[(size_t)pi];
                    if (e.sg == a) model = e.aR + e.bR * r;
                    else if (e.sg + 1 == a) model = e.aL + e.bL * r;
                }
                if (std::fabs(table - model) > 1e-10 * sample_scale) return false;
            }
        }
    }
    vert_units = worst + 0.01; // + the model tolerance (<= 64 rows * 255 * 1e-10 * 257 of a 16-bit step)
    return true;
}

static bool build_target(StreamGeom &g, int ti, const StreamTargetSpec &s, double sample_scale)
{
    StreamTargetGeom &t = g.t[ti];
    t.ax = get_axis_plan(s.dw, s.rect_w);
    t.ay = get_axis_plan(s.dh, s.rect_h);
    const AxisPlan &ax = *t.ax, &ay = *t.ay;
    if (!ax.contiguous || !ay.contiguous) return false;

    // horizontal: normalised fp32 weights
    t.xw.resize(ax.w.size());
    for (int32_t ox = 0; ox < s.dw; ox++)
        for (int32_t k = ax.off[ox]; k < ax.off[ox + 1]; k++) t.xw[k] = (float)(ax.w[k] * ax.inv[ox]);

    // the streaming vertical pass keeps two output rows open: first/last source
    // rows of consecutive outputs must both be strictly increasing and row a+2
    // must start after row a ended.
    auto last = [&](int32_t oy) { return ay.first[oy] + (ay.off[oy + 1] - ay.off[oy]) - 1; };
    for (int32_t oy = 0; oy < s.dh; oy++) {
        if (ay.off[oy + 1] == ay.off[oy]) return false;
        if (oy + 1 < s.dh && (ay.first[oy + 1] < ay.first[oy] || last(oy + 1) <= last(oy))) return false;
        if (oy + 2 < s.dh && ay.first[oy + 2] <= last(oy)) return false;
    }

    // column tiles: an output column belongs to the tile holding its first tap
    t.tile_ox.assign((size_t)g.n_tiles + 1, s.dw);
    {
        int32_t ox = 0;
        for (int32_t tile = 0; tile < g.n_tiles; tile++) {
            t.tile_ox[tile] = ox;
            const int32_t cx1 = (tile + 1) * g.tile_w;
            while (ox < s.dw && ax.first[ox] + s.rect_x < cx1) {
                const int32_t lastcol = ax.first[ox] + s.rect_x + (ax.off[ox + 1] - ax.off[ox]) - 1;
                if (lastcol >= tile * g.tile_w + g.slab_cols) return false; // halo too small
                ox++;
            }
        }
        t.tile_ox[g.n_tiles] = ox;
        if (ox != s.dw) return false;
    }
    // local targets: within a tile, an output column belongs to the V warp whose owned range
    // [w * warp_stride, (w + 1) * warp_stride) (the last warp: up to tile_w) holds its first tap;
    // all its taps must lie inside that warp's 128 loaded columns
    t.warp_ox.clear();
    if (t.local) {
        t.warp_ox.assign((size_t)g.n_tiles * 4 + 1, s.dw);
        int32_t ox = 0;
        for (int32_t tile = 0; tile < g.n_tiles; tile++) {
            for (int32_t w = 0; w < 4; w++) {
                t.warp_ox[(size_t)tile * 4 + w] = ox;
                const int32_t c0 = tile * g.tile_w + w * g.warp_stride;
                const int32_t c1 = w < 3 ? std::min(c0 + g.warp_stride, (tile + 1) * g.tile_w) : (tile + 1) * g.tile_w;
                while (ox < t.tile_ox[tile + 1] && ax.first[ox] + s.rect_x < c1) {
                    const int32_t firstcol = ax.first[ox] + s.rect_x;
                    const int32_t lastcol = firstcol + (ax.off[ox + 1] - ax.off[ox]) - 1;
                    if (firstcol < c0 || lastcol >= c0 + STREAM_WARP_COLS) return false;
                    ox++;
                }
            }
            if (ox != t.tile_ox[tile + 1]) return false;
        }
        t.warp_ox[(size_t)g.n_tiles * 4] = ox;
    }

    // horizontal-pass form per tile (see StreamTarget::tile_parts)
    t.tile_parts.assign((size_t)g.n_tiles, 0);
    for (int32_t tile = 0; tile < g.n_tiles; tile++) {
        const int32_t o0 = t.tile_ox[tile], o1 = t.tile_ox[tile + 1];
        int32_t maxn = 0;
        for (int32_t ox = o0; ox < o1; ox++) maxn = std::max(maxn, ax.off[ox + 1] - ax.off[ox]);
        if (o1 == o0) continue;
        // cheapest (rounds, P) whose table fits: P threads per output, `rounds` outputs per thread
        const int32_t units = t.local ? 32 : STREAM_THREADS;
        int32_t maxo = o1 - o0;
        if (t.local) {
            maxo = 0;
            for (int32_t w = 0; w < 4; w++)
                maxo = std::max(maxo, t.warp_ox[(size_t)tile * 4 + w + 1] - t.warp_ox[(size_t)tile * 4 + w]);
        }
        for (int32_t R = 1; R <= STREAM_XROUNDS && !t.tile_parts[tile]; R++) {
            for (int32_t P = 1; P <= (t.local ? 4 : 32); P <<= 1) {
                if (maxo * P > units * R) break;
                const int32_t ntap = (maxn + P - 1) / P;
                if (ntap * R > STREAM_XTAPS_TAB || (t.local && R == 1 && ntap > STREAM_XTAPS)) continue;
                t.tile_parts[tile] = P | (ntap << 8) | (R << 16);
                break;
            }
        }
    }

    // row bands: an output row belongs to the band holding its first source row
    t.band_rec_off.assign((size_t)g.n_bands, 0);
    t.band_tend.assign((size_t)g.n_bands, 0);
    t.band_oy.assign((size_t)g.n_bands + 1, s.dh);
    int32_t oy = 0;
    for (int32_t b = 0; b < g.n_bands; b++) {
        const int32_t Y0 = g.band_y[b], Y1 = g.band_y[b + 1];
        const int32_t oyA = oy;
        while (oy < s.dh && ay.first[oy] + s.rect_y < Y1) oy++;
        const int32_t oyB = oy;
        t.band_oy[b] = oyA;
        t.band_rec_off[b] = (int32_t)t.rows.size();
        int32_t tend = Y0;
        if (oyB > oyA) tend = last(oyB - 1) + s.rect_y + 1;
        t.band_tend[b] = tend;
        int32_t a = oyA;
        for (int32_t ys = Y0; ys < tend; ys++) {
            const int32_t yr = ys - s.rect_y;
            RowRec r{0.f, 0.f, -1, 0};
            auto weight = [&](int32_t o) -> float {
                if (o >= oyB || yr < ay.first[o] || yr > last(o)) return 0.f;
                const int32_t k = ay.off[o] + (yr - ay.first[o]);
                return (float)(ay.w[k] * ay.inv[o] * sample_scale);
            };
            r.wa = weight(a);
            r.wb = weight(a + 1);
            if (a + 2 < oyB && yr >= ay.first[a + 2]) return false;
            if (a < oyB && last(a) == yr) {
                r.emit = a;
                a++;
                if (a < oyB && last(a) == yr) return false;
            }
            t.rows.push_back(r);
        }
        if (a != oyB) return false;
    }
    if (oy != s.dh) return false;

    t.fix_d = certified_fix_d(ax.max_taps, ay.max_taps, [&] {
        int p = 1;
        for (int32_t v : t.tile_parts) p = std::max(p, (int)(v & 255));
        return p;
    }());
    return true;
}

static std::shared_ptr<const StreamGeom> build_stream(int W, int H, const StreamTargetSpec *targets, int n_targets,
                                                      bool has_wm, int n_bands, double sample_scale)
{
    auto g = std::make_shared<StreamGeom>();
    g->W = W;
    g->H = H;
    g->n_targets = n_targets;
    g->has_wm = has_wm;

    int max_taps_x = 1, local_halo = 0;
    double max_scale_y = 1.0;
    for (int i = 0; i < n_targets; i++) {
        const StreamTargetSpec &s = targets[i];
        if (s.dw <= 0 || s.dh <= 0 || s.rect_w <= 0 || s.rect_h <= 0) return nullptr;
        if (s.rect_h < s.dh) return nullptr; // vertical upscale: many open rows -> k_exact
        if (s.dw > 32767) return nullptr;    // the kernel's per-thread tables hold output columns in 16 bits
        auto ax = get_axis_plan(s.dw, s.rect_w);
        max_taps_x = std::max(max_taps_x, ax->max_taps);
        max_scale_y = std::max(max_scale_y, (double)s.rect_h / (double)s.dh);
        // narrow horizontal support: every V warp filters its own columns (no hand-off)
        g->t[i].local = ax->max_taps - 1 <= STREAM_LOCAL_MAX_HALO;
        if (g->t[i].local) local_halo = std::max(local_halo, ax->max_taps - 1);
    }
    // consecutive V warps overlap by the local halo, rounded up to the 4-pixel thread granule
    g->warp_stride = STREAM_WARP_COLS - (local_halo + STREAM_PX - 1) / STREAM_PX * STREAM_PX;
    g->slab_cols = 3 * g->warp_stride + STREAM_WARP_COLS;
    // planar (16-bit sample) sources: tile origins must keep the Y and the half-resolution chroma rows
    // 16-byte aligned for the bulk copies
    const int col_align = sample_scale == 1.0 ? 32 : STREAM_PX;
    const int tile_w_max = ((g->slab_cols - (max_taps_x - 1)) / col_align) * col_align;
    if (tile_w_max < 64) return nullptr;
    g->n_tiles = (W + tile_w_max - 1) / tile_w_max;
    g->tile_w = (((W + g->n_tiles - 1) / g->n_tiles) + col_align - 1) / col_align * col_align;
    if (g->tile_w > tile_w_max) g->tile_w = tile_w_max;
    g->n_tiles = (W + g->tile_w - 1) / g->tile_w;

    // bands: at least ~8 vertical supports tall so the re-read tail stays small
    const int min_rows = std::max(32, (int)std::ceil(16.0 * max_scale_y));
    n_bands = std::max(1, std::min(n_bands, H / min_rows));
    if (n_bands < 1) n_bands = 1;
    g->n_bands = n_bands;
    g->band_y.resize((size_t)n_bands + 1);
    for (int b = 0; b <= n_bands; b++) g->band_y[b] = (int32_t)((int64_t)H * b / n_bands);

    for (int i = 0; i < n_targets; i++)
        if (!build_target(*g, i, targets[i], sample_scale)) return nullptr;

    g->band_yend.resize((size_t)n_bands);
    for (int b = 0; b < n_bands; b++) {
        int32_t e = has_wm ? g->band_y[b + 1] : g->band_y[b];
        for (int i = 0; i < n_targets; i++) e = std::max(e, g->t[i].band_tend[b]);
        if (e > H) return nullptr;
        g->band_yend[b] = e;
    }
    // one wide 8-bit target: its vertical pass may also have the integer-moment form (one more record slot per group)
    std::vector<GroupRecI> irec;
    double vert_units = 0.0;
    g->vint_ok = n_targets == 1 && !g->t[0].local && build_vint(*g, targets[0], sample_scale, irec, vert_units);
    if (g->vint_ok) {
        int parts = 1;
        for (int32_t v : g->t[0].tile_parts) parts = std::max(parts, (int)(v & 255));
        g->t[0].fix_d_vint = certified_fix_d_vint(g->t[0].ax->max_taps, parts, vert_units);
        if (g->t[0].fix_d_vint >= 120) g->vint_ok = false; // (an opaque alpha sits 128 units from a quantiser step)
    }
    const int slots = n_targets + (g->vint_ok ? 1 : 0);
    g->rec_slots = slots;
    // Group records: the RowRecs re-expressed per accumulator set (parity resolved here), with
    // the opaque-alpha chains the kernel's fast path reads instead of computing.
    g->band_grec_off.resize((size_t)n_bands);
    for (int b = 0; b < n_bands; b++) {
        const int32_t Y0 = g->band_y[b];
        const int32_t ng = (g->band_yend[b] - Y0 + STREAM_GROUP - 1) / STREAM_GROUP;
        const size_t base = g->grec.size() / (size_t)std::max(slots, 1);
        g->band_grec_off[b] = (int32_t)base;
        if (n_targets == 0) continue;
        g->grec.resize((base + (size_t)ng) * (size_t)slots);
        if (g->vint_ok) // the integer-form records of this band's groups, in the slot after the fp32 one
            for (int32_t gi = 0; gi < ng; gi++)
                memcpy(&g->grec[(base + (size_t)gi) * (size_t)slots + 1], &irec[base + (size_t)gi], sizeof(GroupRec));
        for (int i = 0; i < n_targets; i++) {
            const StreamTargetGeom &t = g->t[i];
            int par = 0;
            float sa[2] = {0.f, 0.f};
            const float asamp = (float)(65535.0 / sample_scale); // opaque alpha in source-sample units (255 for 8-bit RGBA)
            for (int32_t gi = 0; gi < ng; gi++) {
                GroupRec &G = g->grec[(base + (size_t)gi) * (size_t)slots + (size_t)i];
                G.seed0 = sa[0];
                G.seed1 = sa[1];
                G.pad[0] = G.pad[1] = 0;
                for (int k = 0; k < STREAM_GROUP; k++) {
                    const int32_t ys = Y0 + gi * STREAM_GROUP + k;
                    RowRec r{0.f, 0.f, -1, 0};
                    if (ys < t.band_tend[b]) r = t.rows[(size_t)t.band_rec_off[b] + (size_t)(ys - Y0)];
                    float w[2];
                    w[par] = r.wa;
                    w[par ^ 1] = r.wb;
                    sa[0] = std::fmaf(asamp, w[0], sa[0]);
                    sa[1] = std::fmaf(asamp, w[1], sa[1]);
                    G.row[k] = GroupRow{w[0], w[1], sa[0], sa[1]};
                    G.emit[k] = -1;
                    if (r.emit >= 0) {
                        G.emit[k] = (r.emit << 1) | par;
                        sa[par] = 0.f;
                        par ^= 1;
                    }
                }
            }
        }
    }
    g->lean2_ok = n_targets == 2 && g->t[0].local && !g->t[1].local;
    for (int i = 0; i < 2 && g->lean2_ok; i++)
        for (int tile = 0; tile < g->n_tiles && g->lean2_ok; tile++)
            if (g->t[i].tile_ox[tile + 1] > g->t[i].tile_ox[tile] &&
                ((g->t[i].tile_parts[tile] & 255) < 1 || (i == 0 && (g->t[i].tile_parts[tile] >> 16) != 1)))
                g->lean2_ok = false; // the fused kernel keeps target 0's taps in registers: one output per lane group
    g->lean_regs_ok = n_targets == 1 && g->t[0].local;
    for (int tile = 0; tile < g->n_tiles && g->lean_regs_ok; tile++)
        if (g->t[0].tile_ox[tile + 1] > g->t[0].tile_ox[tile] && (g->t[0].tile_parts[tile] >> 16) != 1) g->lean_regs_ok = false;
    g->lean_ok = n_targets == 1;
    for (int tile = 0; tile < g->n_tiles && g->lean_ok; tile++)
        if (g->t[0].tile_ox[tile + 1] > g->t[0].tile_ox[tile] && (g->t[0].tile_parts[tile] & 255) < 1) g->lean_ok = false;
    for (int b = 0; b < n_bands; b++) {
        for (int tile = 0; tile < g->n_tiles; tile++) {
            bool work = has_wm;
            for (int i = 0; i < n_targets && !work; i++)
                work = g->t[i].band_tend[b] > g->band_y[b] && g->t[i].tile_ox[tile + 1] > g->t[i].tile_ox[tile];
            if (work && g->band_yend[b] > g->band_y[b]) g->items.push_back(StreamItem{0, (int16_t)tile, (int16_t)b});
        }
    }
    return g;
}

std::shared_ptr<const DirectGeom> get_direct_geom(const StreamTargetSpec &s, double sample_scale, bool mild_downscales)
{
    if (s.dw <= 0 || s.dh <= 0 || s.rect_w <= 0 || s.rect_h <= 0) return nullptr;
    if (s.rect_h >= s.dh && !mild_downscales) return nullptr; // streams: measured faster there (see plan.h)
    using Key = std::tuple<int, int, int, int, long long>;
    static std::mutex mu;
    static AgedCache<Key, std::shared_ptr<const DirectGeom>> cache(2048);
    const Key key{s.dw, s.rect_w, s.dh, s.rect_h, (long long)(sample_scale * 65536.0)};
    {
        std::lock_guard<std::mutex> lk(mu);
        std::shared_ptr<const DirectGeom> hit;
        if (cache.find(key, hit)) return hit;
    }
    std::shared_ptr<DirectGeom> g;
    auto ax = get_axis_plan(s.dw, s.rect_w), ay = get_axis_plan(s.dh, s.rect_h);
    const bool v_upscale = s.rect_h < s.dh;
    if (ax->contiguous && ay->contiguous && ay->max_taps <= DIRECT_MAX_TAPS && (ax->max_taps <= DIRECT_MAX_TAPS || (v_upscale && ax->max_taps <= 256))) {
        g = std::make_shared<DirectGeom>();
        g->ax = ax;
        g->ay = ay;
        g->xw.resize(ax->w.size());
        for (int32_t o = 0; o < s.dw; o++)
            for (int32_t k = ax->off[o]; k < ax->off[o + 1]; k++) g->xw[k] = (float)(ax->w[k] * ax->inv[o]);
        g->yw.resize(ay->w.size());
        for (int32_t o = 0; o < s.dh; o++)
            for (int32_t k = ay->off[o]; k < ay->off[o + 1]; k++) g->yw[k] = (float)(ay->w[k] * ay->inv[o] * sample_scale);
        g->fix_d = certified_fix_d(ax->max_taps, ay->max_taps, 1);
        if (g->fix_d >= 120) g.reset(); // an opaque alpha sits 128 units from a quantiser step: keep the window inside it
    }
    std::lock_guard<std::mutex> lk(mu);
    cache.put(key, g);
    return g;
}

namespace {
struct GeomKey {
    int W, H, n_targets, has_wm, n_bands;
    long long scale_bits;
    StreamTargetSpec t[2];
    bool operator<(const GeomKey &o) const
    {
        auto tup = [](const GeomKey &k) {
            return std::make_tuple(k.W, k.H, k.n_targets, k.has_wm, k.n_bands, k.scale_bits, k.t[0].rect_x, k.t[0].rect_y,
                                   k.t[0].rect_w, k.t[0].rect_h, k.t[0].dw, k.t[0].dh, k.t[1].rect_x, k.t[1].rect_y,
                                   k.t[1].rect_w, k.t[1].rect_h, k.t[1].dw, k.t[1].dh);
        };
        return tup(*this) < tup(o);
    }
};
} // namespace

std::shared_ptr<const StreamGeom> get_stream_geom(int W, int H, const StreamTargetSpec *targets, int n_targets,
                                                  bool has_wm, int n_bands_hint, double sample_scale)
{
    static std::mutex mu;
    static AgedCache<GeomKey, std::shared_ptr<const StreamGeom>> cache(1024); // value may be nullptr (infeasible)
    GeomKey key{};
    key.W = W; key.H = H; key.n_targets = n_targets; key.has_wm = has_wm; key.n_bands = n_bands_hint;
    key.scale_bits = (long long)(sample_scale * 65536.0);
    for (int i = 0; i < n_targets && i < 2; i++) key.t[i] = targets[i];
    {
        std::lock_guard<std::mutex> lk(mu);
        std::shared_ptr<const StreamGeom> hit;
        if (cache.find(key, hit)) return hit;
    }
    auto g = (n_targets <= 2) ? build_stream(W, H, targets, n_targets, has_wm, n_bands_hint, sample_scale) : nullptr;
    std::lock_guard<std::mutex> lk(mu);
    cache.put(key, g);
    return g;
}

} // namespace ipg
