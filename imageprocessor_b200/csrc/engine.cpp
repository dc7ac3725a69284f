// engine.cpp -- libipgpu.so host engine + C ABI (include/ipgpu.h).
//
// One submission queue and two service threads per device (a batcher that turns
// queued tickets into one launch sequence on a free lane, and a completer that
// retires lanes in order); `lanes_per_device` lanes, each a CUDA stream with its
// own device arena and pinned parameter blob, so the H2D of batch k+1, the kernels
// of batch k and the D2H of batch k-1 overlap.  Images shard by ticket across
// devices; there is no cross-device exchange, hence no collective.
//
// The drop-in boundary is processor.ImageProcessor.Process
// (internal/usecase/processor/image_processor.go:39): the Go side keeps decode,
// parameter parsing, geometry, glyph rasterisation, encode and SaveProcessed and
// calls ipg_submit/ipg_wait where it called resizeImage / cropAndResize /
// addTextWatermark.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ipgpu.h"
#include "ipg_device.h"
#include "kernels.h"
#define IPG_WITH_CUDA_RUNTIME 1
#include "jpeg.h"
#include "plan.h"

namespace ipg {

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
static std::string cuda_msg(const char *what, cudaError_t e)
{
    return std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}
#define IPG_CU(call)                                                      \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) throw std::runtime_error(cuda_msg(#call, e__)); \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int plane_count(int layout) { return (layout >= IPG_LAYOUT_YCBCR444 && layout <= IPG_LAYOUT_YCBCR440) ? 3 : 1; }
static void plane_dims(int layout, int plane, int w, int h, int *pw_bytes, int *ph)
{
    if (plane == 0) {
        *pw_bytes = (layout == IPG_LAYOUT_RGBA8 || layout == IPG_LAYOUT_NRGBA8) ? w * 4
                    : (layout == IPG_LAYOUT_RGBA64 || layout == IPG_LAYOUT_NRGBA64) ? w * 8
                    : layout == IPG_LAYOUT_GRAY16 ? w * 2 : w;
        *ph = h;
        return;
    }
    int cw = w, ch = h; // image.NewYCbCr chroma sizes
    if (layout == IPG_LAYOUT_YCBCR422 || layout == IPG_LAYOUT_YCBCR420) cw = (w + 1) / 2;
    if (layout == IPG_LAYOUT_YCBCR420 || layout == IPG_LAYOUT_YCBCR440) ch = (h + 1) / 2;
    *pw_bytes = cw;
    *ph = ch;
}

// ---------------------------------------------------------------------------------
// pinned memory: caller-visible allocations (registry) and an internal staging pool
// ---------------------------------------------------------------------------------
class PinnedRegistry {
public:
    void add(void *p, size_t n)
    {
        std::lock_guard<std::mutex> lk(mu_);
        map_[(uintptr_t)p] = n;
    }
    bool remove(void *p)
    {
        std::lock_guard<std::mutex> lk(mu_);
        return map_.erase((uintptr_t)p) > 0;
    }
    bool contains(const void *p, size_t n)
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = map_.upper_bound((uintptr_t)p);
        if (it == map_.begin()) return false;
        --it;
        return (uintptr_t)p + n <= it->first + it->second;
    }
    std::vector<void *> drain()
    {
        std::lock_guard<std::mutex> lk(mu_);
        std::vector<void *> v;
        for (auto &kv : map_) v.push_back((void *)kv.first);
        map_.clear();
        return v;
    }

private:
    std::mutex mu_;
    std::map<uintptr_t, size_t> map_;
};

// First-fit pool over one pinned slab; alloc blocks until space is free.
class StagingPool {
public:
    bool init(size_t bytes)
    {
        if (cudaHostAlloc((void **)&base_, bytes, cudaHostAllocPortable) != cudaSuccess) return false;
        size_ = bytes;
        free_[0] = bytes;
        return true;
    }
    void destroy()
    {
        if (base_) cudaFreeHost(base_);
        base_ = nullptr;
    }
    size_t size() const { return size_; }
    // Waits at most timeout_ms for space (the ABI returns a status, it does not block for ever: staging of a
    // pageable destination is only released by ipg_wait, so a caller that submits more than the pool holds before it
    // waits would otherwise hang).  *timed_out tells a full pool from a request that can never fit.
    uint8_t *alloc(size_t n, const std::atomic<bool> &stop, int timeout_ms, bool *timed_out)
    {
        if (timed_out) *timed_out = false;
        n = align_up(std::max<size_t>(n, 1), 256);
        if (n > size_) return nullptr;
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(std::max(timeout_ms, 0));
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            for (auto it = free_.begin(); it != free_.end(); ++it) {
                if (it->second >= n) {
                    size_t off = it->first, rem = it->second - n;
                    free_.erase(it);
                    if (rem) free_[off + n] = rem;
                    used_[off] = n;
                    return base_ + off;
                }
            }
            if (stop.load()) return nullptr;
            if (std::chrono::steady_clock::now() >= deadline) {
                if (timed_out) *timed_out = true;
                return nullptr;
            }
            cv_.wait_for(lk, std::chrono::milliseconds(20));
        }
    }
    void release(uint8_t *p)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu_);
        size_t off = (size_t)(p - base_);
        auto u = used_.find(off);
        if (u == used_.end()) return;
        size_t n = u->second;
        used_.erase(u);
        auto nx = free_.lower_bound(off);
        if (nx != free_.end() && off + n == nx->first) {
            n += nx->second;
            nx = free_.erase(nx);
        }
        if (nx != free_.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) {
                pv->second += n;
                cv_.notify_all();
                return;
            }
        }
        free_[off] = n;
        cv_.notify_all();
    }

private:
    uint8_t *base_ = nullptr;
    size_t size_ = 0;
    std::mutex mu_;
    std::condition_variable cv_;
    std::map<size_t, size_t> free_, used_;
};

// ---------------------------------------------------------------------------------
// tickets
// ---------------------------------------------------------------------------------
struct GlyphRec {
    ipg_glyph g;
    std::vector<uint8_t> mask; // owned copy, tight stride = mask_w
};

struct OpRec {
    int kind = 0;
    int dw = 0, dh = 0;
    int rx = 0, ry = 0, rw = 0, rh = 0;
    uint8_t color[4] = {0, 0, 0, 0};
    std::vector<GlyphRec> glyphs;
    void *dst = nullptr;
    int dst_stride = 0;
    int dst_mem = 0;
    int flags = 0;
    bool ycc_out = false;     // dst_layout == IPG_LAYOUT_YCBCR420: dst / dst_cb / dst_cr are the planes, dev_out an arena RGBA temp
    void *dst_cb = nullptr, *dst_cr = nullptr;
    int dst_cstride = 0;
    bool jpeg_out = false;    // dst_layout == IPG_LAYOUT_JPEG: dst is a byte buffer for the file, dev_out an arena RGBA temp
    int jpeg_quality = 85;
    size_t dst_capacity = 0;
    uint64_t *dst_len = nullptr;
    bool patch_only = false;  // IPG_OPF_WATERMARK_PATCH_ONLY on an RGBA8 source: only the glyph union box is produced
    int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0; // ... that box (empty: nothing to do)
    uint8_t *stage = nullptr; // staging for a non-pinned host dst (tight rows)
    uint8_t *dev_out = nullptr;
    size_t dev_pitch = 0;
};

struct Ticket {
    uint64_t id = 0;
    int dev = 0;
    ipg_image_desc src{};
    uint8_t *src_stage[3] = {nullptr, nullptr, nullptr};
    std::vector<OpRec> ops;
    size_t dev_bytes = 0; // arena bytes this ticket needs
    size_t cost = 0;      // for device selection
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    int status = IPG_OK;
    std::string err;
};
using TicketP = std::shared_ptr<Ticket>;

// ---------------------------------------------------------------------------------
// lanes, devices, context
// ---------------------------------------------------------------------------------
struct Batch {
    std::vector<TicketP> tickets;
    int lane = -1;
    bool failed = false;
    std::string err;
    int n_kernels = 0;
    uint64_t h2d = 0, d2h = 0;
    uint64_t exact_fallbacks = 0;
    uint64_t fast_jobs = 0;
    bool has_fix = false;
    // results encoded on the device (IPG_LAYOUT_JPEG): the file length is known only after the kernels ran, so the
    // completer reads it (result words, copied with the batch's read-backs) and issues the exact-size copy itself
    struct JpegOut { const uint8_t *dev; void *host; size_t cap; uint64_t *len_out; int ticket; int result; bool to_host; };
    std::vector<JpegOut> jpegs;
};

struct Lane {
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;     // side stream: the general k_stream launch that does not depend on the lean one
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evf = nullptr;      // after the lean k_stream launch
    cudaEvent_t begin = nullptr;    // before the batch's first H2D (upload stream)
    cudaEvent_t uploaded = nullptr; // after its last H2D (upload stream)
    cudaEvent_t dl_begin = nullptr; // before its first D2H (download stream, after the wait for the kernels)
    cudaEvent_t done = nullptr;     // after its last D2H (download stream)
    uint8_t *arena = nullptr;
    size_t arena_bytes = 0;
    uint8_t *param_host = nullptr; // pinned
    size_t param_cap = 0;
    uint32_t *fix_count_host = nullptr; // pinned copy of FixList::count (4 words)
    uint32_t *jpeg_result_host = nullptr; // pinned copy of the batch's JpegJob::result words (4 per job)
    cudaEvent_t done2 = nullptr;    // after the completer's exact-size JPEG copies (download stream 2)
    bool busy = false;
};

struct Ctx;
struct Device {
    Ctx *ctx = nullptr;
    int index = 0;
    int cuda_id = 0;
    std::vector<Lane> lanes;
    std::mutex mu;
    std::condition_variable cv_q, cv_lane, cv_idle;
    std::deque<TicketP> queue;
    std::deque<std::unique_ptr<Batch>> inflight;
    std::atomic<uint64_t> outstanding{0};
    std::thread batcher, completer;
    StagingPool staging;
    // device-time spans since the last ipg_reset_stats, in ms after `epoch`
    // One upload and one download stream per device, shared by the lanes in batch order:
    // with a copy stream per lane the driver maps several lanes onto one copy engine and an
    // H2D queues behind another lane's D2H; two dedicated streams always run both directions.
    cudaStream_t up = nullptr, down = nullptr;
    cudaStream_t down2 = nullptr;   // the completer's exact-size copies of device-encoded JPEG files
    cudaEvent_t epoch = nullptr;
    // Compute sections of consecutive batches (different lanes) are ordered by events owned by the earlier batch's lane:
    // the next batch's stream kernels wait for `last_stream` (end of the previous batch's k_stream section) and its
    // fix / blend section waits for `last_compute` (end of the previous batch's last kernel).  So the small latency-bound
    // tail of batch k (k_exact_fix, k_exact_fix_wide, k_blend: ~8 % of a step) runs beside the bandwidth-bound k_stream
    // of batch k+1 instead of in front of it, while two k_stream sections still never evict each other.
    cudaEvent_t last_stream = nullptr;
    cudaEvent_t last_compute = nullptr;
    cudaEvent_t prev_compute = nullptr; // ... and of the batch before that: a stream section never starts before the tail
                                        // two batches back has finished (callers that resubmit the same destination
                                        // buffers back to back keep the old "later batch wins" behaviour)
    double k_first = 1e300, k_last = -1, b_first = 1e300, b_last = -1;
};

struct Ctx {
    ipg_config cfg{};
    std::vector<std::unique_ptr<Device>> devs;
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> next_id{1};
    std::mutex tmu;
    std::unordered_map<uint64_t, TicketP> tickets;
    PinnedRegistry pinned;
    bool trace = false; // IPG_TRACE=1: per-batch device timeline on stderr
    // Resample targets fused into one pass over the source.  One (default): on B200 two single-target
    // passes (3 CTAs/SM each) beat one two-target pass (its 48 accumulators per thread leave 2 CTAs/SM)
    // even though the source is read twice -- measured 32.0 vs 39.0 us per 12 MP image.  ipg_config.fuse_targets = 2
    // (or IPG_FUSE_TARGETS=2) restores the single pass.
    // 1: one target per pass (default: measured fastest on B200 -- 28.4 us per 12 MP image for the full pipeline);
    // 2: two targets in the general instantiation (39 us); 3: a local and a wide target in the lean fused
    // instantiation when eligible (38 us: 48 accumulators per thread cost more occupancy than the second read saves)
    int fuse_targets = 1;
    // local- and wide-target lean jobs in ONE launch (k_stream<1,WM,4>): CTAs of the HBM-bound resize + watermark
    // pass and of the issue-bound thumbnail pass share the SMs (26.2 vs 28.4 us per 12 MP image).  IPG_MERGE_LEAN=0
    // launches them separately (k_stream<1,WM,1> and <1,WM,2> on two streams), which is how bench.py times each pass alone.
    bool merge_lean = true;
    uint32_t band_cta_target = 1400;
    uint32_t fix_capacity = 0;     // IPG_FIX_CAPACITY: fix-list entries per batch (0: sized from the batch); tests force the overflow paths with it
    int use_direct = 1;          // IPG_DIRECT=0: vertical upscales take the whole-image fp64 kernel as in round 1; 2: mild downscales too go to k_direct
    bool overlap_streams = true; // IPG_NO_OVERLAP=1: lean and general k_stream launches back to back (per-kernel timing)
    // IPG_OVERLAP_TAIL=1: a batch's stream kernels start as soon as the previous batch's STREAM section ends, beside its
    // fix / blend tail.  Measured (r2): +1.4 % images/s device-resident, but the tail kernels then share SMs with
    // k_stream and its own duration (what the roofline is quoted on) grows 3 %; off by default.
    bool overlap_tail = false;
    // IPG_VINT=0: wide 8-bit targets (the 15:1 thumbnail) keep the fp32 vertical pass instead of the integer-moment form
    bool use_vint = true;
    int staging_timeout_ms = 2000; // IPG_STAGING_TIMEOUT_MS: how long ipg_submit waits for pinned staging before IPG_ERR_NOMEM
    // stats
    std::atomic<uint64_t> s_done{0}, s_batches{0}, s_kernels{0}, s_h2d{0}, s_d2h{0}, s_fix{0}, s_fallback{0}, s_staged{0};
    std::mutex smu;
    double s_stream_ms = 0, s_fix_ms = 0, s_other_ms = 0, s_fast_ms = 0, s_h2d_ms = 0, s_d2h_ms = 0;
    std::atomic<uint64_t> s_fast_jobs{0};
};

} // namespace ipg

struct ipg_ctx : ipg::Ctx {};

namespace ipg {

// bump allocator over a lane's device arena
struct Arena {
    uint8_t *base;
    size_t cap, off = 0;
    uint8_t *take(size_t n, size_t a = 256)
    {
        size_t o = align_up(off, a);
        if (o + n > cap) return nullptr;
        off = o + n;
        return base + o;
    }
};

// host-side image of the parameter blob with the matching device base address
struct Blob {
    uint8_t *host;
    uint8_t *dev;
    size_t cap, off = 0;
    bool overflow = false;
    // Tables already in the blob, keyed by (address, bytes) of the host vector.  The vectors belong to cached plans
    // (StreamGeom / AxisPlan); the plan caches may evict at any time, on this thread or another device's batcher,
    // so every plan a batch reads from is pinned in `keep` until the blob is complete -- a freed vector's address
    // can then never be reused by a different table while `seen` still maps it.
    std::map<std::pair<const void *, size_t>, size_t> seen;
    std::vector<std::shared_ptr<const void>> keep;
    template <typename P> void hold(const std::shared_ptr<P> &p) { if (p) keep.push_back(std::static_pointer_cast<const void>(p)); }
    template <typename T> T *dptr(size_t o) const { return (T *)(dev + o); }
    size_t put(const void *data, size_t bytes, size_t a = 16)
    {
        size_t o = align_up(off, a);
        if (o + bytes > cap) {
            overflow = true;
            return 0;
        }
        if (bytes) memcpy(host + o, data, bytes);
        off = o + bytes;
        return o;
    }
    template <typename T> const T *put_vec(const std::vector<T> &v)
    {
        const std::pair<const void *, size_t> key{(const void *)v.data(), v.size() * sizeof(T)};
        auto it = seen.find(key);
        if (it != seen.end()) return dptr<const T>(it->second);
        size_t o = put(v.data(), v.size() * sizeof(T), 16);
        if (!overflow) seen[key] = o;
        return dptr<const T>(o);
    }
    size_t reserve(size_t bytes, size_t a = 16)
    {
        size_t o = align_up(off, a);
        if (o + bytes > cap) {
            overflow = true;
            return 0;
        }
        off = o + bytes;
        return o;
    }
};

static AxisExact pack_axis(Blob &b, const AxisPlan &p)
{
    AxisExact a;
    a.off = b.put_vec(p.off);
    a.first = b.put_vec(p.first);
    a.inv = b.put_vec(p.inv);
    a.inv_ffff = b.put_vec(p.inv_ffff);
    a.w = b.put_vec(p.w);
    return a;
}

// Device-encoded JPEG result (IPG_LAYOUT_JPEG): buffer sizes of one job.  The scan can never exceed 1248 bytes per MCU
// (6 blocks x [20 DC bits + 63 x (16-bit code + 10 value bits)]); the caller's capacity bounds it further.
enum { kJpegMaxJobs = 4096 };
struct JpegSizes { int mcu_w, n_mcu; size_t scan_cap, out_cap, total; };
static JpegSizes jpeg_sizes(int w, int h, size_t dst_capacity, bool out_in_arena)
{
    JpegSizes z{};
    z.mcu_w = (w + 15) / 16;
    z.n_mcu = z.mcu_w * ((h + 15) / 16);
    const size_t cap = std::min<size_t>(dst_capacity, 0xffff0000u);
    z.scan_cap = align_up(std::max<size_t>(std::min((size_t)z.n_mcu * 1248 + 64, cap), 512), JPEG_CHUNK);
    z.out_cap = std::min(cap, (size_t)JPEG_HDR_MAX + 2 * z.scan_cap + 2);
    z.total = (size_t)(z.n_mcu + 31) / 32 * 6 * JPEG_SLOT_WORDS * JPEG_SLOT_STRIDE * 4 + (size_t)z.n_mcu * (6 * 4 + 4) + z.scan_cap + (z.scan_cap / JPEG_CHUNK + 1) * 4 +
              (out_in_arena ? z.out_cap : 0) + 8 * 256;
    return z;
}

static size_t ticket_device_bytes(const Ticket &t)
{
    size_t n = 0;
    if (t.src.memspace == IPG_MEM_HOST ||
        (t.src.layout == IPG_LAYOUT_RGBA8 &&
         ((((uintptr_t)t.src.plane[0]) | (uintptr_t)t.src.stride[0] | ((uintptr_t)t.src.width * 4)) & 15))) {
        for (int p = 0; p < plane_count(t.src.layout); p++) {
            int wb, ph;
            plane_dims(t.src.layout, p, t.src.width, t.src.height, &wb, &ph);
            n += (align_up((size_t)wb, 256) + 256) * (size_t)ph + 256;
        }
    }
    for (auto &op : t.ops) {
        if (op.patch_only) n += (align_up((size_t)std::max(op.bx1 - op.bx0, 0) * 4, 256) + 256) * (size_t)std::max(op.by1 - op.by0, 0) + 256;
        else if (op.ycc_out) // the RGBA result in the arena, plus the three planes when they are read back to the host
            n += (align_up((size_t)std::max(op.dw, 0) * 4, 256) + 256) * (size_t)std::max(op.dh, 0) +
                 2 * (align_up((size_t)std::max(op.dw, 0), 256) + 256) * (size_t)std::max(op.dh, 0) + 1024;
        else if (op.jpeg_out) // the RGBA result in the arena, coefficients, the scan in both forms
            n += (align_up((size_t)std::max(op.dw, 0) * 4, 256) + 256) * (size_t)std::max(op.dh, 0) +
                 jpeg_sizes(std::max(op.dw, 1), std::max(op.dh, 1), op.dst_capacity, op.dst_mem == IPG_MEM_HOST).total + sizeof(JpegTables) + JPEG_HDR_MAX + 1024;
        else if (op.dst_mem == IPG_MEM_HOST) n += (align_up((size_t)std::max(op.dw, 0) * 4, 256) + 256) * (size_t)std::max(op.dh, 0) + 256;
        for (auto &g : op.glyphs) n += align_up(g.mask.size(), 256) + 256;
        n += 4096;
    }
    return n + 4096;
}

// The device-side JPEG writer's share of a batch: one JpegJob per result that leaves as a file, its buffers carved from
// the lane's arena, its tables (one set per quality) and header in the parameter blob, and the CTA lists of its passes.
struct JpegPlan {
    std::vector<JpegJob> jobs;
    std::vector<JpegDctItem> dct;     // k_jpeg_dct / k_jpeg_emit: 32 MCUs each
    std::vector<JpegStuffItem> stuff; // k_jpeg_zero / k_jpeg_ffcount / k_jpeg_write
    uint32_t *d_results = nullptr;    // 4 words per job
};
static void plan_jpeg_jobs(const std::vector<std::pair<OpRec *, int>> &jpeg_ops, Arena &arena, Blob &blob, Batch &B, JpegPlan &P)
{
    if (jpeg_ops.empty()) return;
    if (jpeg_ops.size() > kJpegMaxJobs) throw std::runtime_error("too many JPEG results in one batch; lower max_batch");
    P.d_results = (uint32_t *)arena.take(16 * jpeg_ops.size(), 256);
    if (!P.d_results) throw std::runtime_error("device arena exhausted (JPEG results)");
    std::map<int, const JpegTables *> tabs; // per quality
    for (auto &jo : jpeg_ops) {
        OpRec &op = *jo.first;
        const bool to_host = op.dst_mem == IPG_MEM_HOST;
        const JpegSizes z = jpeg_sizes(op.dw, op.dh, op.dst_capacity, to_host);
        JpegJob j{};
        j.rgba = op.dev_out; j.rgba_pitch = (int)op.dev_pitch; j.w = op.dw; j.h = op.dh;
        j.mcu_w = z.mcu_w; j.n_mcu = z.n_mcu;
        auto tb = tabs.find(op.jpeg_quality);
        if (tb == tabs.end()) {
            JpegTables T;
            jpeg_build_tables(op.jpeg_quality, &T);
            tb = tabs.emplace(op.jpeg_quality, blob.dptr<const JpegTables>(blob.put(&T, sizeof T, 16))).first;
        }
        j.tab = tb->second;
        uint8_t hdr[JPEG_HDR_MAX];
        j.hdr_len = (uint32_t)jpeg_build_header(op.jpeg_quality, op.dw, op.dh, hdr);
        j.hdr = blob.dptr<const uint8_t>(blob.put(hdr, j.hdr_len, 16));
        j.acs = (uint32_t *)arena.take((size_t)(z.n_mcu + 31) / 32 * 6 * JPEG_SLOT_WORDS * JPEG_SLOT_STRIDE * 4);
        j.side = (uint32_t *)arena.take((size_t)z.n_mcu * 6 * 4);
        j.mcu_off = (uint32_t *)arena.take((size_t)z.n_mcu * 4);
        j.words = (uint32_t *)arena.take(z.scan_cap);
        j.cap_bytes = (uint32_t)z.scan_cap;
        j.chunk_off = (uint32_t *)arena.take((z.scan_cap / JPEG_CHUNK + 1) * 4);
        j.out = to_host ? arena.take(z.out_cap) : (uint8_t *)op.dst;
        j.out_cap = (uint32_t)z.out_cap;
        if (!j.acs || !j.side || !j.mcu_off || !j.words || !j.chunk_off || !j.out) throw std::runtime_error("device arena exhausted (JPEG writer)");
        const int ji = (int)P.jobs.size();
        j.result = P.d_results + 4 * ji;
        P.jobs.push_back(j);
        for (int m = 0; m < z.n_mcu; m += JPEG_DCT_MCUS) P.dct.push_back(JpegDctItem{ji, m});
        // CTAs of the stuffing passes: the scan's real size is known on the device only, so size them for the capacity, 32
        // chunks (4 per warp) each; CTAs past the real end exit at once
        const int parts = (int)std::max<size_t>(1, std::min<size_t>(JPEG_STUFF_PARTS, (z.scan_cap / JPEG_CHUNK + 31) / 32));
        for (int q = 0; q < parts; q++) P.stuff.push_back(JpegStuffItem{ji, (int16_t)q, (int16_t)parts});
        B.jpegs.push_back(Batch::JpegOut{j.out, op.dst, (size_t)op.dst_capacity, op.dst_len, jo.second, ji, to_host});
    }
}

// Build and enqueue one batch on a lane.  Throws std::runtime_error on failure.
static void launch_batch(Ctx &c, Device &d, Lane &L, Batch &B)
{
    IPG_CU(cudaSetDevice(d.cuda_id));
    cudaStream_t st = L.st;       // kernels
    cudaStream_t up = d.up;       // host -> device
    cudaStream_t down = d.down;   // device -> host
    IPG_CU(cudaEventRecord(L.begin, up));
    Arena arena{L.arena, L.arena_bytes};
    uint8_t *blob_dev = arena.take(L.param_cap, 256);
    if (!blob_dev) throw std::runtime_error("device arena smaller than the parameter blob");
    Blob blob{L.param_host, blob_dev, L.param_cap};

    std::vector<StreamJob> sjobs;
    std::vector<StreamItem> sitems;   // general k_stream instantiations, independent of the lean launch
    std::vector<StreamItem> ritems;   // general k_stream instantiation: on-demand redo of lean jobs
    std::vector<StreamItem> fitems;   // lean instantiation, local target
    std::vector<StreamItem> f2items;  // lean instantiation, wide target
    std::vector<StreamItem> f3items;  // lean instantiation, local + wide targets fused
    std::vector<StreamItem> pitems;   // k_stream_planar<false> (YCbCr and Gray sources)
    std::vector<StreamItem> nitems;   // k_stream_planar<true> (NRGBA sources)
    bool any_wm_fast = false, any_wm_fast2 = false, any_wm_fast3 = false;
    size_t max_jobs = 0;
    for (auto &tp : B.tickets) max_jobs += tp->ops.size() + 1;
    int32_t *redo_flags = (int32_t *)arena.take(4 * max_jobs + 256); // one per stream job, raised on the device
    std::vector<ExactJob> fixjobs;    // one per stream target (EXACT mode)
    std::vector<ExactJob> xjobs;      // whole-output fp64 jobs
    std::vector<ExactItem> xitems;
    std::vector<WmJob> wjobs;
    std::vector<WmItem> witems;
    std::vector<WatermarkD> blends;   // every watermark of the batch that has glyphs
    std::vector<BlendItem> bitems;
    std::vector<YccJob> yjobs;        // results handed back as planar YCbCr 4:2:0 (dst_layout)
    std::vector<YccItem> yitems;
    std::vector<std::pair<OpRec *, int>> jpeg_ops; // results encoded on the device (dst_layout JPEG), with their ticket index
    std::vector<DirectJob> djobs;     // small-support targets (vertical upscales, mild downscales): k_direct
    std::vector<DirectItem> ditems;
    std::vector<PatchJob> pjobs;      // patch-only watermarks (RGBA8 sources): glyph box alone
    std::vector<BlendItem> pbitems;
    struct Readback { uint8_t *dev; size_t pitch; void *host; size_t hstride; size_t row_bytes; int rows; };
    std::vector<Readback> readbacks;
    int max_nt = 0;
    bool any_wm = false;
    uint64_t fix_px = 0;
    const int precision = c.cfg.precision;

    // band count: aim for >= ~2 waves of CTAs over the whole batch
    size_t est_ctas = 0;
    for (auto &tp : B.tickets) est_ctas += (size_t)(tp->src.width + 479) / 480;
    const size_t cta_target = c.band_cta_target; // CTAs a batch should at least give (IPG_BAND_CTAS; default 1400 = ~2.4 waves of 4 x 148)
    int bands_hint = (int)std::min<size_t>(32, std::max<size_t>(1, (cta_target + est_ctas - 1) / std::max<size_t>(est_ctas, 1)));

    int ticket_index = -1;
    for (auto &tp : B.tickets) {
        Ticket &t = *tp;
        ticket_index++;
        // ---- source view on the device
        SrcView sv{};
        sv.w = t.src.width;
        sv.h = t.src.height;
        sv.layout = t.src.layout;
        const uint8_t *dp[3] = {nullptr, nullptr, nullptr};
        int ds[3] = {0, 0, 0};
        for (int p = 0; p < plane_count(t.src.layout); p++) {
            int wb, ph;
            plane_dims(t.src.layout, p, t.src.width, t.src.height, &wb, &ph);
            if (t.src.memspace == IPG_MEM_DEVICE) {
                dp[p] = (const uint8_t *)t.src.plane[p];
                ds[p] = t.src.stride[p];
                if (t.src.layout == IPG_LAYOUT_RGBA8 && precision != IPG_PRECISION_REFERENCE &&
                    ((((uintptr_t)dp[p]) | (uintptr_t)ds[p] | (uintptr_t)wb) & 15)) {
                    // k_stream moves rows with TMA bulk copies (16-byte granular): re-pitch once
                    size_t pitch = align_up((size_t)wb, 256);
                    uint8_t *dv = arena.take(pitch * (size_t)ph);
                    if (!dv) throw std::runtime_error("device arena exhausted (re-pitched source)");
                    IPG_CU(cudaMemcpy2DAsync(dv, pitch, dp[p], (size_t)ds[p], (size_t)wb, (size_t)ph, cudaMemcpyDeviceToDevice, up));
                    dp[p] = dv;
                    ds[p] = (int)pitch;
                }
            } else {
                const void *hp = t.src_stage[p] ? (const void *)t.src_stage[p] : t.src.plane[p];
                size_t hs = t.src_stage[p] ? (size_t)wb : (size_t)t.src.stride[p];
                // keep the host stride as device pitch when rows stay 16-byte aligned: one linear DMA
                const bool linear = (hs % 16) == 0 && hs < (size_t)wb + 256;
                size_t pitch = linear ? hs : align_up((size_t)wb, 256);
                uint8_t *dv = arena.take(pitch * (size_t)ph);
                if (!dv) throw std::runtime_error("device arena exhausted (source)");
                if (linear)
                    IPG_CU(cudaMemcpyAsync(dv, hp, hs * (size_t)(ph - 1) + (size_t)wb, cudaMemcpyHostToDevice, up));
                else
                    IPG_CU(cudaMemcpy2DAsync(dv, pitch, hp, hs, (size_t)wb, (size_t)ph, cudaMemcpyHostToDevice, up));
                B.h2d += (uint64_t)wb * (uint64_t)ph;
                dp[p] = dv;
                ds[p] = (int)pitch;
            }
        }
        sv.p0 = dp[0]; sv.p1 = dp[1]; sv.p2 = dp[2];
        sv.s0 = ds[0]; sv.s1 = ds[1]; sv.s2 = ds[2];

        // ---- destinations on the device
        for (auto &op : t.ops) {
            op.dev_out = nullptr;
            if (op.dw <= 0 || op.dh <= 0) continue;
            if (op.patch_only) { // the box alone: a patch buffer read back as a rectangle, or the caller's device frame
                const int bw = op.bx1 - op.bx0, bh = op.by1 - op.by0;
                if (bw <= 0 || bh <= 0) continue;
                if (op.dst_mem == IPG_MEM_DEVICE) {
                    op.dev_out = (uint8_t *)op.dst;
                    op.dev_pitch = (size_t)op.dst_stride;
                } else {
                    const size_t pitch = align_up((size_t)bw * 4, 256);
                    op.dev_out = arena.take(pitch * (size_t)bh);
                    if (!op.dev_out) throw std::runtime_error("device arena exhausted (watermark patch)");
                    op.dev_pitch = pitch;
                    if (op.stage) readbacks.push_back({op.dev_out, pitch, op.stage, (size_t)bw * 4, (size_t)bw * 4, bh});
                    else readbacks.push_back({op.dev_out, pitch, (uint8_t *)op.dst + (size_t)op.by0 * (size_t)op.dst_stride + (size_t)op.bx0 * 4,
                                              (size_t)op.dst_stride, (size_t)bw * 4, bh});
                }
                continue;
            }
            if (op.ycc_out) { // RGBA result in the arena; converted after the last kernel that writes it; planes go back
                const size_t pitch = align_up((size_t)op.dw * 4, 256);
                op.dev_out = arena.take(pitch * (size_t)op.dh);
                if (!op.dev_out) throw std::runtime_error("device arena exhausted (destination)");
                op.dev_pitch = pitch;
                const int cw = (op.dw + 1) / 2, ch = (op.dh + 1) / 2;
                YccJob yj{};
                yj.rgba = op.dev_out; yj.rgba_pitch = (int)pitch; yj.w = op.dw; yj.h = op.dh;
                if (op.dst_mem == IPG_MEM_DEVICE) {
                    yj.y = (uint8_t *)op.dst; yj.cb = (uint8_t *)op.dst_cb; yj.cr = (uint8_t *)op.dst_cr;
                    yj.y_pitch = op.dst_stride; yj.c_pitch = op.dst_cstride;
                } else {
                    const size_t yp = align_up((size_t)op.dw, 256), cp = align_up((size_t)cw, 256);
                    yj.y = arena.take(yp * (size_t)op.dh);
                    yj.cb = arena.take(cp * (size_t)ch);
                    yj.cr = arena.take(cp * (size_t)ch);
                    if (!yj.y || !yj.cb || !yj.cr) throw std::runtime_error("device arena exhausted (YCbCr destination)");
                    yj.y_pitch = (int)yp; yj.c_pitch = (int)cp;
                    readbacks.push_back({yj.y, yp, op.dst, (size_t)op.dst_stride, (size_t)op.dw, op.dh});
                    readbacks.push_back({yj.cb, cp, op.dst_cb, (size_t)op.dst_cstride, (size_t)cw, ch});
                    readbacks.push_back({yj.cr, cp, op.dst_cr, (size_t)op.dst_cstride, (size_t)cw, ch});
                }
                const int ji = (int)yjobs.size();
                yjobs.push_back(yj);
                for (int ty = 0; ty < (op.dh + 15) / 16; ty++)
                    for (int tx = 0; tx < (op.dw + 255) / 256; tx++) yitems.push_back(YccItem{ji, tx, ty, 0});
            } else if (op.jpeg_out) { // RGBA result in the arena; the writer's buffers are carved once every ticket is placed
                const size_t pitch = align_up((size_t)op.dw * 4, 256);
                op.dev_out = arena.take(pitch * (size_t)op.dh);
                if (!op.dev_out) throw std::runtime_error("device arena exhausted (destination)");
                op.dev_pitch = pitch;
                jpeg_ops.push_back({&op, ticket_index});
            } else if (op.dst_mem == IPG_MEM_DEVICE) {
                op.dev_out = (uint8_t *)op.dst;
                op.dev_pitch = (size_t)op.dst_stride;
            } else {
                void *hp = op.stage ? (void *)op.stage : op.dst;
                size_t hs = op.stage ? (size_t)op.dw * 4 : (size_t)op.dst_stride;
                const bool linear = (hs % 16) == 0 && hs < (size_t)op.dw * 4 + 256;
                size_t pitch = linear ? hs : align_up((size_t)op.dw * 4, 256);
                op.dev_out = arena.take(pitch * (size_t)op.dh);
                if (!op.dev_out) throw std::runtime_error("device arena exhausted (destination)");
                op.dev_pitch = pitch;
                readbacks.push_back({op.dev_out, pitch, hp, hs, (size_t)op.dw * 4, op.dh});
            }
        }

        // ---- watermark descriptors
        auto make_wm = [&](OpRec &op) -> WatermarkD {
            WatermarkD w{};
            w.dst = op.dev_out;
            w.dst_stride = (int)op.dev_pitch;
            w.sr = op.color[0] * 0x101u; w.sg = op.color[1] * 0x101u;
            w.sb = op.color[2] * 0x101u; w.sa = op.color[3] * 0x101u;
            std::vector<GlyphD> gd;
            int bx0 = INT32_MAX, by0 = INT32_MAX, bx1 = INT32_MIN, by1 = INT32_MIN;
            for (auto &g : op.glyphs) {
                if (g.g.x0 >= g.g.x1 || g.g.y0 >= g.g.y1) continue;
                const size_t o = blob.put(g.mask.data(), g.mask.size(), 16); // masks ride in the blob
                GlyphD x{};
                x.x0 = g.g.x0; x.y0 = g.g.y0; x.x1 = g.g.x1; x.y1 = g.g.y1;
                x.mp_x = g.g.mp_x; x.mp_y = g.g.mp_y;
                x.mask_stride = g.g.mask_w;
                x.mask = blob.dptr<const uint8_t>(o);
                gd.push_back(x);
                bx0 = std::min(bx0, x.x0); by0 = std::min(by0, x.y0);
                bx1 = std::max(bx1, x.x1); by1 = std::max(by1, x.y1);
            }
            w.n_glyphs = (int)gd.size();
            if (gd.empty()) { bx0 = by0 = bx1 = by1 = 0; }
            size_t o = blob.put(gd.data(), gd.size() * sizeof(GlyphD), 16);
            w.glyphs = blob.dptr<const GlyphD>(o);
            w.bx0 = bx0; w.by0 = by0; w.bx1 = bx1; w.by1 = by1;
            if (w.n_glyphs > 0) {
                const int bi = (int)blends.size();
                blends.push_back(w);
                for (int ty = 0; ty < (by1 - by0 + 7) / 8; ty++)
                    for (int tx = 0; tx < (bx1 - bx0 + 31) / 32; tx++) bitems.push_back(BlendItem{bi, tx, ty});
            }
            return w;
        };

        auto add_exact = [&](OpRec &op, std::vector<ExactJob> &vec) -> int {
            ExactJob j{};
            j.src = sv;
            j.two_stage = op.kind == IPG_OP_THUMB_CROP;
            if (op.kind == IPG_OP_THUMB_CROP) { j.rect_x = op.rx; j.rect_y = op.ry; }
            const int sw = op.kind == IPG_OP_THUMB_CROP ? op.rw : sv.w;
            const int sh = op.kind == IPG_OP_THUMB_CROP ? op.rh : sv.h;
            j.dw = op.dw; j.dh = op.dh;
            j.dst = op.dev_out; j.dst_stride = (int)op.dev_pitch;
            auto pax = get_axis_plan(op.dw, sw), pay = get_axis_plan(op.dh, sh);
            blob.hold(pax);
            blob.hold(pay);
            j.ax = pack_axis(blob, *pax);
            j.ay = pack_axis(blob, *pay);
            vec.push_back(j);
            return (int)vec.size() - 1;
        };
        auto add_exact_whole = [&](OpRec &op) {
            int ji = add_exact(op, xjobs);
            for (int ty = 0; ty < (op.dh + 7) / 8; ty++)
                for (int tx = 0; tx < (op.dw + 31) / 32; tx++) xitems.push_back(ExactItem{ji, tx, ty, 0});
        };

        // what a stream job / target takes from its cached geometry (tables ride in the blob), for both streaming kernels
        auto fill_job = [&](StreamJob &j, const StreamGeom &geom, int nt) {
            j.src = sv;
            j.n_targets = nt;
            j.tile_w = geom.tile_w;
            j.warp_stride = geom.warp_stride;
            j.slab_cols = geom.slab_cols;
            j.n_tiles = geom.n_tiles;
            j.n_bands = geom.n_bands;
            j.band_y = blob.put_vec(geom.band_y);
            j.band_yend = blob.put_vec(geom.band_yend);
            j.grec = blob.put_vec(geom.grec);
            j.rec_slots = geom.rec_slots;
            j.band_grec_off = blob.put_vec(geom.band_grec_off);
        };
        auto fill_target = [&](StreamTarget &o, const StreamTargetGeom &tgm, OpRec &op, const StreamTargetSpec &spec) {
            o.dst = op.dev_out;
            o.dst_stride = (int)op.dev_pitch;
            o.dw = op.dw; o.dh = op.dh;
            o.rect_x = spec.rect_x; o.rect_y = spec.rect_y;
            o.two_stage = op.kind == IPG_OP_THUMB_CROP;
            o.fix_d = tgm.fix_d;
            o.fix_d_vint = tgm.fix_d_vint;
            o.xoff = blob.put_vec(tgm.ax->off);
            o.xfirst = blob.put_vec(tgm.ax->first);
            o.xw = blob.put_vec(tgm.xw);
            o.tile_ox = blob.put_vec(tgm.tile_ox);
            o.local = tgm.local ? 1 : 0;
            o.warp_ox = tgm.local ? blob.put_vec(tgm.warp_ox) : nullptr;
            o.tile_parts = blob.put_vec(tgm.tile_parts);
            o.rows = blob.put_vec(tgm.rows);
            o.band_rec_off = blob.put_vec(tgm.band_rec_off);
            o.band_tend = blob.put_vec(tgm.band_tend);
            o.band_oy = blob.put_vec(tgm.band_oy);
            o.exact_job = -1;
            if (precision == IPG_PRECISION_EXACT) {
                o.exact_job = add_exact(op, fixjobs);
                fix_px += (uint64_t)o.dw * (uint64_t)o.dh;
            }
        };

        // ---- classify ops
        std::vector<OpRec *> res, wms;
        for (auto &op : t.ops) {
            if (op.dw <= 0 || op.dh <= 0) continue;
            if (op.patch_only) {
                if (!op.dev_out) continue; // no glyph touches the image
                PatchJob pj{};
                pj.src = sv;
                pj.wm = make_wm(op);   // (its entry in `blends` is replaced by the patch job's own items below)
                if (pj.wm.n_glyphs > 0) { // make_wm queued in-place blend tiles for this watermark: take them back
                    const int bi = (int)blends.size() - 1;
                    while (!bitems.empty() && bitems.back().wm == bi) bitems.pop_back();
                    blends.pop_back();
                }
                pj.ox = op.dst_mem == IPG_MEM_DEVICE ? 0 : op.bx0;
                pj.oy = op.dst_mem == IPG_MEM_DEVICE ? 0 : op.by0;
                const int ji = (int)pjobs.size();
                pjobs.push_back(pj);
                for (int ty = 0; ty < (op.by1 - op.by0 + 7) / 8; ty++)
                    for (int tx = 0; tx < (op.bx1 - op.bx0 + 31) / 32; tx++) pbitems.push_back(BlendItem{ji, tx, ty});
                continue;
            }
            (op.kind == IPG_OP_WATERMARK ? wms : res).push_back(&op);
        }
        const bool streamable = precision != IPG_PRECISION_REFERENCE && sv.layout == L_RGBA8 &&
                                (sv.s0 % 4) == 0 && ((size_t)sv.p0 % 4) == 0;
        // vertical upscales have no streaming form (many output rows open at once) and the reference upscales small
        // images too (resize.go:63-72): one thread per output pixel (k_direct, fp32 + the same certificate) takes them,
        // where round 1 ran the whole image in fp64
        if (precision != IPG_PRECISION_REFERENCE && (sv.layout != L_RGBA8 || streamable)) {
            std::vector<OpRec *> keep;
            for (auto *op : res) {
                StreamTargetSpec sp{0, 0, sv.w, sv.h, op->dw, op->dh};
                if (op->kind == IPG_OP_THUMB_CROP) sp = StreamTargetSpec{op->rx, op->ry, op->rw, op->rh, op->dw, op->dh};
                auto dg = c.use_direct ? get_direct_geom(sp, sv.layout == L_RGBA8 ? 257.0 : 1.0, c.use_direct == 2) : nullptr;
                if (!dg) { keep.push_back(op); continue; }
                blob.hold(dg);
                DirectJob j{};
                j.src = sv;
                j.rect_x = sp.rect_x; j.rect_y = sp.rect_y;
                j.two_stage = op->kind == IPG_OP_THUMB_CROP;
                j.dw = op->dw; j.dh = op->dh;
                j.dst = op->dev_out; j.dst_stride = (int)op->dev_pitch;
                j.xoff = blob.put_vec(dg->ax->off); j.xfirst = blob.put_vec(dg->ax->first); j.xw = blob.put_vec(dg->xw);
                j.yoff = blob.put_vec(dg->ay->off); j.yfirst = blob.put_vec(dg->ay->first); j.yw = blob.put_vec(dg->yw);
                j.fix_d = dg->fix_d;
                j.exact_job = -1;
                if (precision == IPG_PRECISION_EXACT) {
                    j.exact_job = add_exact(*op, fixjobs);
                    fix_px += (uint64_t)op->dw * (uint64_t)op->dh;
                }
                const int ji = (int)djobs.size();
                djobs.push_back(j);
                for (int ty = 0; ty < (op->dh + 7) / 8; ty++)
                    for (int tx = 0; tx < (op->dw + 31) / 32; tx++) ditems.push_back(DirectItem{ji, tx, ty, 0});
                B.fast_jobs++;
            }
            res.swap(keep);
        }
        size_t wi = 0;
        if (streamable) {
            size_t ri = 0;
            while (ri < res.size() || wi < wms.size()) {
                // take up to fuse_targets resample targets + one watermark per pass over the source
                OpRec *tg[2] = {nullptr, nullptr};
                int nt = 0;
                StreamTargetSpec spec[2];
                auto spec_of = [&](OpRec *op) {
                    StreamTargetSpec s{0, 0, sv.w, sv.h, op->dw, op->dh};
                    if (op->kind == IPG_OP_THUMB_CROP) s = StreamTargetSpec{op->rx, op->ry, op->rw, op->rh, op->dw, op->dh};
                    return s;
                };
                OpRec *wm = wi < wms.size() ? wms[wi] : nullptr;
                std::shared_ptr<const StreamGeom> geom;
                bool lean2 = false;
                if (c.fuse_targets == 3 && ri + 1 < res.size()) {
                    // fuse the next two targets when the lean local + wide instantiation can take them
                    // (either order: the local one goes first), otherwise one target per pass
                    for (int order = 0; order < 2 && !lean2; order++) {
                        OpRec *a = res[ri + order], *b = res[ri + 1 - order];
                        StreamTargetSpec sp2[2] = {spec_of(a), spec_of(b)};
                        auto g2 = get_stream_geom(sv.w, sv.h, sp2, 2, wm != nullptr, bands_hint, 257.0);
                        if (g2 && g2->lean2_ok) {
                            geom = g2;
                            tg[0] = a; tg[1] = b;
                            spec[0] = sp2[0]; spec[1] = sp2[1];
                            nt = 2;
                            ri += 2;
                            lean2 = true;
                        }
                    }
                }
                const int per_pass = c.fuse_targets == 2 ? 2 : 1;
                while (!lean2 && ri < res.size() && nt < per_pass) {
                    OpRec *op = res[ri];
                    spec[nt] = spec_of(op);
                    tg[nt++] = op;
                    ri++;
                }
                if (!geom && (nt > 0 || wm)) geom = get_stream_geom(sv.w, sv.h, spec, nt, wm != nullptr, bands_hint, 257.0);
                if (!geom && nt == 2) { // retry the targets one at a time
                    ri -= 1;
                    nt = 1;
                    tg[1] = nullptr;
                    geom = get_stream_geom(sv.w, sv.h, spec, 1, wm != nullptr, bands_hint, 257.0);
                }
                if (!geom) {
                    if (nt == 1) { add_exact_whole(*tg[0]); B.exact_fallbacks++; }
                    if (nt == 0 && wm) break; // cannot happen (wm-only always streams); fall to k_watermark
                    continue;
                }
                if (wm) wi++;
                blob.hold(geom);
                StreamJob j{};
                fill_job(j, *geom, nt);
                j.has_wm = wm != nullptr;
                for (int k = 0; k < nt; k++) {
                    StreamTarget &o = j.t[k];
                    fill_target(o, geom->t[k], *tg[k], spec[k]);
                    if (o.two_stage && !t.src.opaque_hint) j.check_premul = 1;
                }
                if (wm) j.wm = make_wm(*wm);
                const int ji = (int)sjobs.size();
                // the lean instantiation takes the common case; it needs TMA-storable watermark rows
                const bool wm_tma_ok = !wm || ((sv.w % 4) == 0 && ((((uintptr_t)wm->dev_out) | (uintptr_t)wm->dev_pitch) & 15) == 0);
                j.fast_path = 0;
                // a wide target's vertical pass in the integer-moment form (the lean single-target kernels only)
                const bool vint = c.use_vint && geom->vint_ok;
                if (wm_tma_ok && redo_flags && (size_t)ji < max_jobs) {
                    // 1: the main lean launch (merged: local lane-per-output pass, integer-moment wide pass and -- when built
                    // with them -- the table forms); 2: the wide-target instantiation on the side stream (every table form)
                    if (geom->lean_ok) j.fast_path = (geom->lean_regs_ok || (c.merge_lean && (IPG_LEAN4_TABLES || vint))) ? 1 : 2;
                    else if (lean2 && geom->lean2_ok) j.fast_path = 3;
                }
                j.redo_flag = (j.fast_path && !t.src.opaque_hint) ? redo_flags + ji : nullptr;
                j.vint = (vint && (j.fast_path == 1 || j.fast_path == 2)) ? 1 : 0;
                sjobs.push_back(j);
                if (!j.fast_path) {
                    max_nt = std::max(max_nt, nt);
                    any_wm |= wm != nullptr;
                } else {
                    (j.fast_path == 1 ? any_wm_fast : j.fast_path == 2 ? any_wm_fast2 : any_wm_fast3) |= wm != nullptr;
                    B.fast_jobs++;
                }
                for (auto it : geom->items) {
                    it.job = ji;
                    if (!j.fast_path) sitems.push_back(it);
                    else {
                        (j.fast_path == 1 ? fitems : j.fast_path == 2 ? f2items : f3items).push_back(it);
                        if (j.redo_flag) ritems.push_back(it);
                    }
                }
            }
        } else {
            // planar YCbCr, Gray, NRGBA: the lean 16-bit-sample kernel takes every resample whose geometry streams in a
            // cached form
            const bool planar_ok = precision != IPG_PRECISION_REFERENCE &&
                                   ((sv.layout >= L_YCBCR444 && sv.layout <= L_YCBCR440) || sv.layout == L_GRAY8 ||
                                    sv.layout == L_NRGBA8) &&
                                   ((((uintptr_t)sv.p0) | ((uintptr_t)sv.p1) | ((uintptr_t)sv.p2) | (uintptr_t)sv.s0 |
                                     (uintptr_t)sv.s1 | (uintptr_t)sv.s2) & 15) == 0;
            for (auto *op : res) {
                std::shared_ptr<const StreamGeom> geom;
                StreamTargetSpec sp1{0, 0, sv.w, sv.h, op->dw, op->dh};
                if (op->kind == IPG_OP_THUMB_CROP) sp1 = StreamTargetSpec{op->rx, op->ry, op->rw, op->rh, op->dw, op->dh};
                // the watermark's full-frame conversion rides on a single-stage (resize) pass when one exists: the V lanes
                // hold the 16-bit samples whose high bytes draw.Draw(Src) would store, so the planar source is read once
                // for both (it needs 16-byte aligned destination rows: the arena's always are)
                OpRec *wm = nullptr;
                if (planar_ok && op->kind == IPG_OP_RESIZE && wi < wms.size() &&
                    ((((uintptr_t)wms[wi]->dev_out) | (uintptr_t)wms[wi]->dev_pitch) & 15) == 0) {
                    auto gw = get_stream_geom(sv.w, sv.h, &sp1, 1, true, bands_hint, 1.0);
                    if (gw && gw->lean_ok) { geom = gw; wm = wms[wi]; }
                }
                if (!geom && planar_ok) geom = get_stream_geom(sv.w, sv.h, &sp1, 1, false, bands_hint, 1.0);
                if (!geom || !geom->lean_ok) {
                    add_exact_whole(*op);
                    if (precision != IPG_PRECISION_REFERENCE) B.exact_fallbacks++;
                    continue;
                }
                if (wm) wi++;
                blob.hold(geom);
                StreamJob j{};
                fill_job(j, *geom, 1);
                const StreamTargetGeom &tgm = geom->t[0];
                fill_target(j.t[0], tgm, *op, sp1);
                j.fast_path = 5;
                if (wm) {
                    j.has_wm = 1;
                    j.wm = make_wm(*wm);
                }
                const int ji = (int)sjobs.size();
                sjobs.push_back(j);
                B.fast_jobs++;
                for (auto it : geom->items) {
                    // tiles outside the crop square have no outputs: skip them (unless they carry watermark rows)
                    if (!wm && tgm.tile_ox[it.tile + 1] == tgm.tile_ox[it.tile]) continue;
                    it.job = ji;
                    (sv.layout == L_NRGBA8 ? nitems : pitems).push_back(it);
                }
            }
        }
        for (; wi < wms.size(); wi++) {
            WmJob j{};
            j.src = sv;
            j.wm = make_wm(*wms[wi]);
            const int ji = (int)wjobs.size();
            wjobs.push_back(j);
            for (int y = 0; y < sv.h; y += WM_ROWS) witems.push_back(WmItem{ji, y});
        }
    }

    // ---- device-side JPEG writer (dst_layout JPEG): one job per result
    JpegPlan jp;
    plan_jpeg_jobs(jpeg_ops, arena, blob, B, jp);
    std::vector<JpegJob> &jjobs = jp.jobs;
    std::vector<JpegDctItem> &jdct = jp.dct;
    std::vector<JpegStuffItem> &jstuff = jp.stuff;
    uint32_t *d_jresults = jp.d_results;

    // ---- fix list (EXACT mode)
    FixList fix{nullptr, nullptr, 0};
    if (!fixjobs.empty()) {
        uint64_t cap = std::min<uint64_t>(fix_px / 8 + 4096, 8u << 20);
        if (c.fix_capacity) cap = c.fix_capacity;
        uint8_t *cnt = arena.take(256);
        uint8_t *ent = arena.take((size_t)cap * sizeof(FixEntry));
        if (!cnt || !ent) throw std::runtime_error("device arena exhausted (fix list)");
        fix.count = (uint32_t *)cnt;
        fix.entries = (FixEntry *)ent;
        fix.capacity = (uint32_t)cap;
        IPG_CU(cudaMemsetAsync(cnt, 0, 16, st)); // FixList::count: append counter, re-queue counter, two work cursors
        B.has_fix = true;
    }

    // ---- parameter blob upload
    const StreamJob *d_sjobs = blob.dptr<const StreamJob>(blob.put(sjobs.data(), sjobs.size() * sizeof(StreamJob), 16));
    const StreamItem *d_sitems = blob.dptr<const StreamItem>(blob.put(sitems.data(), sitems.size() * sizeof(StreamItem), 16));
    const StreamItem *d_fitems = blob.dptr<const StreamItem>(blob.put(fitems.data(), fitems.size() * sizeof(StreamItem), 16));
    const StreamItem *d_f2items = blob.dptr<const StreamItem>(blob.put(f2items.data(), f2items.size() * sizeof(StreamItem), 16));
    const StreamItem *d_f3items = blob.dptr<const StreamItem>(blob.put(f3items.data(), f3items.size() * sizeof(StreamItem), 16));
    const StreamItem *d_pitems = blob.dptr<const StreamItem>(blob.put(pitems.data(), pitems.size() * sizeof(StreamItem), 16));
    const StreamItem *d_nitems = blob.dptr<const StreamItem>(blob.put(nitems.data(), nitems.size() * sizeof(StreamItem), 16));
    const StreamItem *d_ritems = blob.dptr<const StreamItem>(blob.put(ritems.data(), ritems.size() * sizeof(StreamItem), 16));
    const ExactJob *d_fixjobs = blob.dptr<const ExactJob>(blob.put(fixjobs.data(), fixjobs.size() * sizeof(ExactJob), 16));
    const ExactJob *d_xjobs = blob.dptr<const ExactJob>(blob.put(xjobs.data(), xjobs.size() * sizeof(ExactJob), 16));
    const ExactItem *d_xitems = blob.dptr<const ExactItem>(blob.put(xitems.data(), xitems.size() * sizeof(ExactItem), 16));
    const WmJob *d_wjobs = blob.dptr<const WmJob>(blob.put(wjobs.data(), wjobs.size() * sizeof(WmJob), 16));
    const WmItem *d_witems = blob.dptr<const WmItem>(blob.put(witems.data(), witems.size() * sizeof(WmItem), 16));
    const WatermarkD *d_blends = blob.dptr<const WatermarkD>(blob.put(blends.data(), blends.size() * sizeof(WatermarkD), 16));
    const BlendItem *d_bitems = blob.dptr<const BlendItem>(blob.put(bitems.data(), bitems.size() * sizeof(BlendItem), 16));
    const YccJob *d_yjobs = blob.dptr<const YccJob>(blob.put(yjobs.data(), yjobs.size() * sizeof(YccJob), 16));
    const YccItem *d_yitems = blob.dptr<const YccItem>(blob.put(yitems.data(), yitems.size() * sizeof(YccItem), 16));
    const DirectJob *d_djobs = blob.dptr<const DirectJob>(blob.put(djobs.data(), djobs.size() * sizeof(DirectJob), 16));
    const DirectItem *d_ditems = blob.dptr<const DirectItem>(blob.put(ditems.data(), ditems.size() * sizeof(DirectItem), 16));
    const PatchJob *d_pjobs = blob.dptr<const PatchJob>(blob.put(pjobs.data(), pjobs.size() * sizeof(PatchJob), 16));
    const BlendItem *d_pbitems = blob.dptr<const BlendItem>(blob.put(pbitems.data(), pbitems.size() * sizeof(BlendItem), 16));
    const JpegJob *d_jjobs = blob.dptr<const JpegJob>(blob.put(jjobs.data(), jjobs.size() * sizeof(JpegJob), 16));
    const JpegDctItem *d_jdct = blob.dptr<const JpegDctItem>(blob.put(jdct.data(), jdct.size() * sizeof(JpegDctItem), 16));
    const JpegStuffItem *d_jstuff = blob.dptr<const JpegStuffItem>(blob.put(jstuff.data(), jstuff.size() * sizeof(JpegStuffItem), 16));
    if (blob.overflow) throw std::runtime_error("parameter blob overflow (batch too heterogeneous); lower max_batch");
    IPG_CU(cudaMemcpyAsync(blob_dev, L.param_host, blob.off, cudaMemcpyHostToDevice, up));
    B.h2d += blob.off;
    IPG_CU(cudaEventRecord(L.uploaded, up));
    IPG_CU(cudaStreamWaitEvent(st, L.uploaded, 0));

    // ---- kernels.  Compute sections of different lanes run one after another (each kernel
    // already fills the GPU; overlapping them only makes them evict each other), while the
    // copies of the other lanes overlap with them.
    if (cudaEvent_t prev = c.overlap_tail ? d.last_stream : d.last_compute) IPG_CU(cudaStreamWaitEvent(st, prev, 0));
    if (c.overlap_tail && d.prev_compute) IPG_CU(cudaStreamWaitEvent(st, d.prev_compute, 0));
    if ((!fitems.empty() || !f2items.empty() || !f3items.empty()) && redo_flags) IPG_CU(cudaMemsetAsync(redo_flags, 0, 4 * max_jobs, st));
    IPG_CU(cudaEventRecord(L.ev[0], st));
    if (!ditems.empty()) {
        IPG_CU(launch_direct(d_djobs, d_ditems, (int)ditems.size(), fix, st));
        B.n_kernels++;
    }
    // Two streams: the lean local-target launch (resize + watermark copy) on the lane's stream; beside it, on a
    // side stream, the lean wide-target launch (thumbnail) and the general launch over whatever neither lean kernel
    // takes -- each fills the other's ramp and tail.  The on-demand redo of lean jobs follows both.
    const bool side_work = !f2items.empty() || !sitems.empty();
    const bool main_work = !fitems.empty() || !f3items.empty() || !pitems.empty() || !nitems.empty();
    const bool side = side_work && main_work && c.overlap_streams;
    cudaStream_t s2 = side ? L.st2 : st;
    if (side) {
        IPG_CU(cudaEventRecord(L.fork, st));
        IPG_CU(cudaStreamWaitEvent(L.st2, L.fork, 0));
    }
    auto launch_main = [&]() { // the lean launches that carry most of the bytes
        if (!pitems.empty()) {
            IPG_CU(launch_stream_planar(d_sjobs, d_pitems, (int)pitems.size(), false, fix, st));
            B.n_kernels++;
        }
        if (!nitems.empty()) {
            IPG_CU(launch_stream_planar(d_sjobs, d_nitems, (int)nitems.size(), true, fix, st));
            B.n_kernels++;
        }
        if (!f3items.empty()) {
            IPG_CU(launch_stream_fast(d_sjobs, d_f3items, (int)f3items.size(), 3, any_wm_fast3, fix, st));
            B.n_kernels++;
        }
        if (!fitems.empty()) {
            IPG_CU(launch_stream_fast(d_sjobs, d_fitems, (int)fitems.size(), c.merge_lean ? 4 : 1, any_wm_fast, fix, st));
            B.n_kernels++;
        }
        IPG_CU(cudaEventRecord(L.evf, st));
    };
    if (!side && main_work) launch_main(); // no overlap: timed alone
    if (!f2items.empty()) {
        IPG_CU(launch_stream_fast(d_sjobs, d_f2items, (int)f2items.size(), 2, any_wm_fast2, fix, s2));
        B.n_kernels++;
    }
    if (!sitems.empty()) {
        IPG_CU(launch_stream(d_sjobs, d_sitems, (int)sitems.size(), max_nt, any_wm, fix, s2));
        B.n_kernels++;
    }
    if (side) {
        IPG_CU(cudaEventRecord(L.join, L.st2));
        launch_main();
        IPG_CU(cudaStreamWaitEvent(st, L.join, 0));
    } else if (!main_work) {
        IPG_CU(cudaEventRecord(L.evf, st));
    }
    if (!ritems.empty()) {
        IPG_CU(launch_stream(d_sjobs, d_ritems, (int)ritems.size(), f3items.empty() ? 1 : 2,
                             any_wm_fast || any_wm_fast2 || any_wm_fast3, fix, st));
        B.n_kernels++;
    }
    IPG_CU(cudaEventRecord(L.ev[1], st));
    d.last_stream = L.ev[1];
    if (c.overlap_tail && d.last_compute) IPG_CU(cudaStreamWaitEvent(st, d.last_compute, 0)); // tails stay in batch order
    if (!fixjobs.empty()) {
        IPG_CU(launch_exact_fix(d_fixjobs, (int)fixjobs.size(), fix, st));
        B.n_kernels += 2; // k_exact_fix + k_exact_fix_wide
    }
    IPG_CU(cudaEventRecord(L.ev[2], st));
    if (!xitems.empty()) {
        IPG_CU(launch_exact_tiles(d_xjobs, d_xitems, (int)xitems.size(), st));
        B.n_kernels++;
    }
    if (!witems.empty()) {
        IPG_CU(launch_watermark(d_wjobs, d_witems, (int)witems.size(), st));
        B.n_kernels++;
    }
    if (!bitems.empty()) {
        IPG_CU(launch_blend(d_blends, d_bitems, (int)bitems.size(), st));
        B.n_kernels++;
    }
    if (!pbitems.empty()) {
        IPG_CU(launch_blend_patch(d_pjobs, d_pbitems, (int)pbitems.size(), st));
        B.n_kernels++;
    }
    if (!yitems.empty()) { // after every kernel that writes an RGBA result (stream, fix-ups, blends)
        IPG_CU(launch_rgba_to_ycbcr420(d_yjobs, d_yitems, (int)yitems.size(), st));
        B.n_kernels++;
    }
    if (!jjobs.empty()) { // ... and the JPEG writer over the results that leave as files
        IPG_CU(cudaMemsetAsync(d_jresults, 0, 16 * jjobs.size(), st));
        IPG_CU(launch_jpeg(d_jjobs, (int)jjobs.size(), d_jdct, (int)jdct.size(), d_jstuff, (int)jstuff.size(), st));
        B.n_kernels += JPEG_LAUNCHES;
    }
    IPG_CU(cudaEventRecord(L.ev[3], st));
    d.prev_compute = d.last_compute;
    d.last_compute = L.ev[3];

    // ---- read back
    IPG_CU(cudaStreamWaitEvent(down, L.ev[3], 0));
    IPG_CU(cudaEventRecord(L.dl_begin, down));
    if (B.has_fix) IPG_CU(cudaMemcpyAsync(L.fix_count_host, fix.count, 16, cudaMemcpyDeviceToHost, down));
    if (!jjobs.empty()) IPG_CU(cudaMemcpyAsync(L.jpeg_result_host, d_jresults, 16 * jjobs.size(), cudaMemcpyDeviceToHost, down));
    for (auto &r : readbacks) {
        if (r.pitch == r.hstride)
            IPG_CU(cudaMemcpyAsync(r.host, r.dev, r.pitch * (size_t)(r.rows - 1) + r.row_bytes, cudaMemcpyDeviceToHost, down));
        else
            IPG_CU(cudaMemcpy2DAsync(r.host, r.hstride, r.dev, r.pitch, r.row_bytes, (size_t)r.rows, cudaMemcpyDeviceToHost, down));
        B.d2h += (uint64_t)r.row_bytes * (uint64_t)r.rows;
    }
    IPG_CU(cudaEventRecord(L.done, down));
}

static void finish_ticket(Ctx &c, Device &d, const TicketP &t, int status, const std::string &err)
{
    for (int p = 0; p < 3; p++) {
        if (t->src_stage[p]) d.staging.release(t->src_stage[p]);
        t->src_stage[p] = nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(t->mu);
        t->status = status;
        t->err = err;
        t->done = true;
    }
    t->cv.notify_all();
    d.outstanding.fetch_sub(t->cost);
    c.s_done++;
}

static void batcher_main(Ctx *c, Device *d)
{
    cudaSetDevice(d->cuda_id);
    for (;;) {
        std::unique_ptr<Batch> B(new Batch);
        int lane = -1;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_q.wait(lk, [&] { return c->stop.load() || !d->queue.empty(); });
            if (c->stop.load() && d->queue.empty()) return;
            // a free lane first: tickets keep arriving while we wait for one
            d->cv_lane.wait(lk, [&] {
                for (auto &L : d->lanes)
                    if (!L.busy) return true;
                return c->stop.load();
            });
            for (size_t i = 0; i < d->lanes.size(); i++)
                if (!d->lanes[i].busy) { lane = (int)i; break; }
            if (lane < 0) return;
            // short window to let concurrent submitters fill the batch
            if ((int)d->queue.size() < c->cfg.max_batch && c->cfg.batch_window_us > 0) {
                d->cv_q.wait_for(lk, std::chrono::microseconds(c->cfg.batch_window_us),
                                 [&] { return (int)d->queue.size() >= c->cfg.max_batch || c->stop.load(); });
            }
            Lane &L = d->lanes[lane];
            size_t budget = L.arena_bytes - L.param_cap - (64u << 20);
            size_t used = 0;
            while (!d->queue.empty() && (int)B->tickets.size() < c->cfg.max_batch) {
                TicketP &t = d->queue.front();
                if (!B->tickets.empty() && used + t->dev_bytes > budget) break;
                used += t->dev_bytes;
                B->tickets.push_back(t);
                d->queue.pop_front();
            }
            L.busy = true;
            B->lane = lane;
        }
        try {
            launch_batch(*c, *d, d->lanes[lane], *B);
        } catch (const std::exception &e) {
            B->failed = true;
            B->err = e.what();
            cudaGetLastError();
        }
        {
            std::lock_guard<std::mutex> lk(d->mu);
            d->inflight.push_back(std::move(B));
        }
        d->cv_idle.notify_all();
    }
}

static void completer_main(Ctx *c, Device *d)
{
    cudaSetDevice(d->cuda_id);
    for (;;) {
        std::unique_ptr<Batch> B;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_idle.wait(lk, [&] { return !d->inflight.empty() || c->stop.load(); });
            if (d->inflight.empty()) {
                if (c->stop.load()) return;
                continue;
            }
            B = std::move(d->inflight.front());
            d->inflight.pop_front();
        }
        Lane &L = d->lanes[B->lane];
        int status = IPG_OK;
        std::string err;
        std::vector<int> tstatus; // per ticket, when only some of a batch's tickets fail (a JPEG file that does not fit)
        if (B->failed) {
            cudaDeviceSynchronize();
            status = IPG_ERR_CUDA;
            err = B->err;
            if (err.find("arena") != std::string::npos || err.find("blob") != std::string::npos) status = IPG_ERR_NOMEM;
        } else {
            cudaError_t e = cudaEventSynchronize(L.done);
            if (e != cudaSuccess) {
                status = IPG_ERR_CUDA;
                err = cuda_msg("batch execution", e);
            } else {
                float a = 0, b = 0, o = 0, f = 0;
                cudaEventElapsedTime(&f, L.ev[0], L.evf);
                cudaEventElapsedTime(&a, L.ev[0], L.ev[1]);
                cudaEventElapsedTime(&b, L.ev[1], L.ev[2]);
                cudaEventElapsedTime(&o, L.ev[2], L.ev[3]);
                float up_ms = 0, down_ms = 0;
                cudaEventElapsedTime(&up_ms, L.begin, L.uploaded);
                cudaEventElapsedTime(&down_ms, L.dl_begin, L.done);
                float t0 = 0, t1 = 0, t2 = 0, t3 = 0;
                cudaEventElapsedTime(&t0, d->epoch, L.begin);
                cudaEventElapsedTime(&t1, d->epoch, L.ev[0]);
                cudaEventElapsedTime(&t2, d->epoch, L.ev[3]);
                cudaEventElapsedTime(&t3, d->epoch, L.done);
                if (c->trace)
                    fprintf(stderr, "[ipg trace] dev %d lane %d tickets %zu: begin %.3f ms, kernels %.3f..%.3f ms, done %.3f ms (h2d %.1f MB, d2h %.1f MB)\n",
                            d->index, B->lane, B->tickets.size(), t0, t1, t2, t3, B->h2d / 1e6, B->d2h / 1e6);
                std::lock_guard<std::mutex> lk(c->smu);
                c->s_stream_ms += a;
                c->s_fast_ms += f;
                c->s_fix_ms += b;
                c->s_other_ms += o;
                c->s_h2d_ms += up_ms;
                c->s_d2h_ms += down_ms;
                if (B->n_kernels > 0) {
                    d->k_first = std::min(d->k_first, (double)t1);
                    d->k_last = std::max(d->k_last, (double)t2);
                }
                d->b_first = std::min(d->b_first, (double)t0);
                d->b_last = std::max(d->b_last, (double)t3);
            }
            // Files encoded on the device: their lengths arrived with the read-backs; fetch exactly those bytes now.
            // Only this thread uses the second download stream, so waiting on it waits for nothing else.
            if (status == IPG_OK && !B->jpegs.empty()) {
                tstatus.assign(B->tickets.size(), IPG_OK);
                bool any_copy = false;
                for (auto &j : B->jpegs) {
                    const uint32_t *r = L.jpeg_result_host + 4 * j.result;
                    if (r[1] != 0 || r[0] == 0 || r[0] > j.cap) {
                        tstatus[(size_t)j.ticket] = IPG_ERR_NOMEM;
                        if (j.len_out) *j.len_out = 0;
                        continue;
                    }
                    if (j.len_out) *j.len_out = r[0];
                    if (!j.to_host) continue;
                    cudaError_t ce = cudaMemcpyAsync(j.host, j.dev, r[0], cudaMemcpyDeviceToHost, d->down2);
                    if (ce != cudaSuccess) { status = IPG_ERR_CUDA; err = cuda_msg("JPEG read-back", ce); break; }
                    B->d2h += r[0];
                    any_copy = true;
                }
                if (status == IPG_OK && any_copy) {
                    cudaEventRecord(L.done2, d->down2);
                    cudaError_t ce = cudaEventSynchronize(L.done2);
                    if (ce != cudaSuccess) { status = IPG_ERR_CUDA; err = cuda_msg("JPEG read-back", ce); }
                    else {
                        float t4 = 0, dl = 0;
                        cudaEventElapsedTime(&t4, d->epoch, L.done2);
                        cudaEventElapsedTime(&dl, L.done, L.done2);
                        std::lock_guard<std::mutex> lk(c->smu);
                        d->b_last = std::max(d->b_last, (double)t4);
                        c->s_d2h_ms += dl;
                    }
                }
            }
            if (B->has_fix && status == IPG_OK) c->s_fix += L.fix_count_host[0];
            c->s_batches++;
            c->s_kernels += (uint64_t)B->n_kernels;
            c->s_h2d += B->h2d;
            c->s_d2h += B->d2h;
            c->s_fallback += B->exact_fallbacks;
            c->s_fast_jobs += B->fast_jobs;
        }
        for (size_t i = 0; i < B->tickets.size(); i++) {
            if (status == IPG_OK && i < tstatus.size() && tstatus[i] != IPG_OK)
                finish_ticket(*c, *d, B->tickets[i], tstatus[i], "the JPEG file does not fit dst_capacity (or the scan its device buffer)");
            else finish_ticket(*c, *d, B->tickets[i], status, err);
        }
        {
            std::lock_guard<std::mutex> lk(d->mu);
            L.busy = false;
        }
        d->cv_lane.notify_all();
        d->cv_idle.notify_all();
    }
}

// ---------------------------------------------------------------------------------
static int validate(const ipg_image_desc *src, const ipg_op *ops, int n_ops)
{
    if (!src || !ops || n_ops <= 0) return fail(IPG_ERR_INVALID, "null source/ops or n_ops <= 0");
    if (src->layout < IPG_LAYOUT_RGBA8 || src->layout > IPG_LAYOUT_GRAY16) return fail(IPG_ERR_INVALID, "unknown source layout");
    if (src->layout >= IPG_LAYOUT_RGBA64 && src->width > (1 << 27)) return fail(IPG_ERR_INVALID, "source image too large");
    if (src->width <= 0 || src->height <= 0) return fail(IPG_ERR_INVALID, "source image is empty");
    if ((int64_t)src->width * src->height > (int64_t)1 << 30) return fail(IPG_ERR_INVALID, "source image too large");
    for (int p = 0; p < plane_count(src->layout); p++) {
        int wb, ph;
        plane_dims(src->layout, p, src->width, src->height, &wb, &ph);
        if (!src->plane[p]) return fail(IPG_ERR_INVALID, "source plane pointer is null");
        if (src->stride[p] < wb) return fail(IPG_ERR_INVALID, "source stride smaller than a row");
    }
    if (src->memspace != IPG_MEM_HOST && src->memspace != IPG_MEM_DEVICE) return fail(IPG_ERR_INVALID, "bad source memspace");
    if (src->memspace == IPG_MEM_DEVICE && src->layout <= IPG_LAYOUT_NRGBA8 && ((src->stride[0] & 3) || ((uintptr_t)src->plane[0] & 3)))
        return fail(IPG_ERR_INVALID, "device RGBA source must be 4-byte aligned");
    for (int i = 0; i < n_ops; i++) {
        const ipg_op &o = ops[i];
        if (o.kind != IPG_OP_RESIZE && o.kind != IPG_OP_THUMB_CROP && o.kind != IPG_OP_WATERMARK)
            return fail(IPG_ERR_INVALID, "unsupported operation kind");
        if (o.dst_w > 65536 || o.dst_h > 65536) return fail(IPG_ERR_INVALID, "destination too large");
        if (o.kind == IPG_OP_WATERMARK && (o.dst_w != src->width || o.dst_h != src->height))
            return fail(IPG_ERR_INVALID, "watermark destination must have the source size");
        if (o.dst_layout != IPG_LAYOUT_RGBA8 && o.dst_layout != IPG_LAYOUT_YCBCR420 && o.dst_layout != IPG_LAYOUT_JPEG)
            return fail(IPG_ERR_INVALID, "destination layout must be RGBA8, YCBCR420 or JPEG");
        if (o.dst_layout == IPG_LAYOUT_JPEG) {
            if (o.dst_w <= 0 || o.dst_h <= 0) return fail(IPG_ERR_INVALID, "a JPEG result needs a non-empty image");
            if (o.dst_w >= 65536 || o.dst_h >= 65536) return fail(IPG_ERR_INVALID, "jpeg: image is too large to encode");
            if (!o.dst || !o.dst_len) return fail(IPG_ERR_INVALID, "JPEG destination needs dst and dst_len");
            if (o.dst_capacity < 1024) return fail(IPG_ERR_INVALID, "JPEG destination capacity below 1024 bytes");
            if (o.flags & IPG_OPF_WATERMARK_PATCH_ONLY) return fail(IPG_ERR_INVALID, "a patch-only watermark has no JPEG form");
        } else if (o.dst_layout == IPG_LAYOUT_YCBCR420 && o.dst_w > 0 && o.dst_h > 0) {
            if (!o.dst || !o.dst_cb || !o.dst_cr) return fail(IPG_ERR_INVALID, "destination plane pointer is null");
            if (o.dst_stride < o.dst_w || o.dst_cstride < (o.dst_w + 1) / 2) return fail(IPG_ERR_INVALID, "destination plane stride smaller than a row");
            if (o.flags & IPG_OPF_WATERMARK_PATCH_ONLY) return fail(IPG_ERR_INVALID, "a patch-only watermark has no YCbCr form");
        } else if (o.dst_w > 0 && o.dst_h > 0) {
            if (!o.dst) return fail(IPG_ERR_INVALID, "destination pointer is null");
            if (o.dst_stride < o.dst_w * 4) return fail(IPG_ERR_INVALID, "destination stride smaller than a row");
            if (o.dst_memspace == IPG_MEM_DEVICE && ((o.dst_stride & 3) || ((uintptr_t)o.dst & 3)))
                return fail(IPG_ERR_INVALID, "device destination must be 4-byte aligned");
        }
        if (o.kind == IPG_OP_THUMB_CROP) {
            if (o.rect_w <= 0 || o.rect_h <= 0 || o.rect_x < 0 || o.rect_y < 0 || o.rect_x + o.rect_w > src->width ||
                o.rect_y + o.rect_h > src->height)
                return fail(IPG_ERR_INVALID, "crop rectangle outside the source");
        }
        if (o.kind == IPG_OP_WATERMARK) {
            if (o.n_glyphs < 0 || (o.n_glyphs > 0 && !o.glyphs)) return fail(IPG_ERR_INVALID, "bad glyph list");
            for (int g = 0; g < o.n_glyphs; g++) {
                const ipg_glyph &G = o.glyphs[g];
                if (G.x0 >= G.x1 || G.y0 >= G.y1) continue; // dr.Empty(): nothing drawn
                if (G.x0 < 0 || G.y0 < 0 || G.x1 > src->width || G.y1 > src->height)
                    return fail(IPG_ERR_INVALID, "glyph rectangle must be clipped to the image");
                if (!G.mask || G.mask_w <= 0 || G.mask_h <= 0 || G.mask_stride < G.mask_w)
                    return fail(IPG_ERR_INVALID, "bad glyph mask");
                if (G.mp_x < 0 || G.mp_y < 0 || G.mp_x + (G.x1 - G.x0) > G.mask_w || G.mp_y + (G.y1 - G.y0) > G.mask_h)
                    return fail(IPG_ERR_INVALID, "glyph rectangle exceeds its mask");
            }
        }
    }
    return IPG_OK;
}

static int submit_impl(Ctx *c, int dev_index, const ipg_image_desc *src, const ipg_op *ops, int n_ops, ipg_ticket *out)
{
    if (!c || !out) return fail(IPG_ERR_INVALID, "null context or ticket pointer");
    if (c->stop.load()) return fail(IPG_ERR_SHUTDOWN, "context is shutting down");
    int rc = validate(src, ops, n_ops);
    if (rc) return rc;
    bool any_device_mem = src->memspace == IPG_MEM_DEVICE;
    for (int i = 0; i < n_ops; i++) any_device_mem |= ops[i].dst_memspace == IPG_MEM_DEVICE;
    if (dev_index < 0 && any_device_mem) return fail(IPG_ERR_INVALID, "device-resident buffers need ipg_submit_on");
    if (dev_index >= (int)c->devs.size()) return fail(IPG_ERR_INVALID, "device index out of range");
    if (dev_index < 0) { // least outstanding work
        uint64_t best = UINT64_MAX;
        for (size_t i = 0; i < c->devs.size(); i++) {
            uint64_t o = c->devs[i]->outstanding.load();
            if (o < best) { best = o; dev_index = (int)i; }
        }
    }
    Device &d = *c->devs[dev_index];

    auto t = std::make_shared<Ticket>();
    t->dev = dev_index;
    t->src = *src;
    t->ops.resize((size_t)n_ops);
    for (int i = 0; i < n_ops; i++) {
        const ipg_op &o = ops[i];
        OpRec &r = t->ops[i];
        r.kind = o.kind; r.dw = o.dst_w; r.dh = o.dst_h;
        r.rx = o.rect_x; r.ry = o.rect_y; r.rw = o.rect_w; r.rh = o.rect_h;
        memcpy(r.color, o.color, 4);
        r.dst = o.dst; r.dst_stride = o.dst_stride; r.dst_mem = o.dst_memspace;
        r.flags = o.flags;
        r.ycc_out = o.dst_layout == IPG_LAYOUT_YCBCR420;
        r.dst_cb = o.dst_cb; r.dst_cr = o.dst_cr; r.dst_cstride = o.dst_cstride;
        r.jpeg_out = o.dst_layout == IPG_LAYOUT_JPEG;
        if (r.jpeg_out) {
            r.jpeg_quality = o.jpeg_quality == 0 ? 85 : std::min(100, std::max(1, (int)o.jpeg_quality));
            r.dst_capacity = (size_t)o.dst_capacity;
            r.dst_len = o.dst_len;
            // both are written after ipg_submit returned (by the completer thread): the library keeps no caller pointer
            // past the call unless it lies in its own pinned memory (cgo pointer rule)
            if (!c->pinned.contains(o.dst_len, sizeof(uint64_t)) ||
                (o.dst_memspace == IPG_MEM_HOST && !c->pinned.contains(o.dst, (size_t)o.dst_capacity)))
                return fail(IPG_ERR_INVALID, "a JPEG destination and its dst_len must lie in ipg_alloc_pinned memory (dst may be device memory)");
        }
        if (r.ycc_out && o.dst_memspace == IPG_MEM_HOST && o.dst_w > 0 && o.dst_h > 0) {
            const int cw = (o.dst_w + 1) / 2, ch = (o.dst_h + 1) / 2;
            if (!c->pinned.contains(o.dst, (size_t)o.dst_stride * (size_t)(o.dst_h - 1) + (size_t)o.dst_w) ||
                !c->pinned.contains(o.dst_cb, (size_t)o.dst_cstride * (size_t)(ch - 1) + (size_t)cw) ||
                !c->pinned.contains(o.dst_cr, (size_t)o.dst_cstride * (size_t)(ch - 1) + (size_t)cw))
                return fail(IPG_ERR_INVALID, "YCbCr destination planes must lie in ipg_alloc_pinned (or device) memory");
        }
        if (o.kind == IPG_OP_WATERMARK && (o.flags & IPG_OPF_WATERMARK_PATCH_ONLY) && src->layout == IPG_LAYOUT_RGBA8 && o.dst_w > 0 && o.dst_h > 0) {
            r.patch_only = true; // the union of the non-empty glyph rectangles (what freetype's DrawMask calls can touch)
            int bx0 = INT32_MAX, by0 = INT32_MAX, bx1 = INT32_MIN, by1 = INT32_MIN;
            for (int g = 0; g < o.n_glyphs; g++) {
                const ipg_glyph &G = o.glyphs[g];
                if (G.x0 >= G.x1 || G.y0 >= G.y1) continue;
                bx0 = std::min(bx0, G.x0); by0 = std::min(by0, G.y0);
                bx1 = std::max(bx1, G.x1); by1 = std::max(by1, G.y1);
            }
            if (bx0 < bx1 && by0 < by1) { r.bx0 = bx0; r.by0 = by0; r.bx1 = bx1; r.by1 = by1; }
        }
        if (o.kind == IPG_OP_WATERMARK) {
            for (int g = 0; g < o.n_glyphs; g++) {
                const ipg_glyph &G = o.glyphs[g];
                GlyphRec gr;
                gr.g = G;
                if (G.x0 < G.x1 && G.y0 < G.y1) {
                    gr.mask.resize((size_t)G.mask_w * (size_t)G.mask_h);
                    for (int y = 0; y < G.mask_h; y++)
                        memcpy(gr.mask.data() + (size_t)y * G.mask_w, G.mask + (size_t)y * G.mask_stride, (size_t)G.mask_w);
                }
                gr.g.mask = nullptr;
                r.glyphs.push_back(std::move(gr));
            }
        }
    }
    // caller memory that is not ours is copied now (cgo: no retained pointers)
    if (src->memspace == IPG_MEM_HOST) {
        for (int p = 0; p < plane_count(src->layout); p++) {
            int wb, ph;
            plane_dims(src->layout, p, src->width, src->height, &wb, &ph);
            size_t span = (size_t)src->stride[p] * (size_t)(ph - 1) + (size_t)wb;
            if (c->pinned.contains(src->plane[p], span)) continue;
            bool timed_out = false;
            uint8_t *s = d.staging.alloc((size_t)wb * (size_t)ph, c->stop, c->staging_timeout_ms, &timed_out);
            if (!s) {
                for (int q = 0; q < p; q++) d.staging.release(t->src_stage[q]);
                if (timed_out)
                    return fail(IPG_ERR_NOMEM, "pinned staging pool exhausted by tickets that were not waited for yet (ipg_wait earlier tickets, "
                                               "raise lane_pinned_bytes, or use ipg_alloc_pinned buffers)");
                return fail(IPG_ERR_NOMEM, "image larger than the pinned staging pool (raise lane_pinned_bytes or decode into ipg_alloc_pinned memory)");
            }
            for (int y = 0; y < ph; y++)
                memcpy(s + (size_t)y * wb, (const uint8_t *)src->plane[p] + (size_t)y * src->stride[p], (size_t)wb);
            t->src_stage[p] = s;
            c->s_staged++;
        }
    }
    for (auto &r : t->ops) {
        if (r.dst_mem != IPG_MEM_HOST || r.dw <= 0 || r.dh <= 0 || r.ycc_out || r.jpeg_out) continue;
        size_t span = (size_t)r.dst_stride * (size_t)(r.dh - 1) + (size_t)r.dw * 4;
        if (c->pinned.contains(r.dst, span)) continue;
        if (r.patch_only && (r.bx1 <= r.bx0 || r.by1 <= r.by0)) continue; // nothing will be written
        bool timed_out = false;
        const size_t stage_bytes = r.patch_only ? (size_t)(r.bx1 - r.bx0) * 4 * (size_t)(r.by1 - r.by0) : (size_t)r.dw * 4 * (size_t)r.dh;
        r.stage = d.staging.alloc(stage_bytes, c->stop, c->staging_timeout_ms, &timed_out);
        if (!r.stage) {
            for (auto &q : t->ops) { d.staging.release(q.stage); q.stage = nullptr; }
            for (int p = 0; p < 3; p++) d.staging.release(t->src_stage[p]);
            if (timed_out)
                return fail(IPG_ERR_NOMEM, "pinned staging pool exhausted by outputs of tickets that were not waited for yet (staging of a pageable "
                                           "destination is held until ipg_wait: wait for earlier tickets, raise lane_pinned_bytes, or use "
                                           "ipg_alloc_pinned destinations)");
            return fail(IPG_ERR_NOMEM, "output larger than the pinned staging pool");
        }
    }
    t->dev_bytes = ticket_device_bytes(*t);
    t->cost = (size_t)src->width * (size_t)src->height * 4 + 65536;
    if (t->dev_bytes + d.lanes[0].param_cap + (64u << 20) > d.lanes[0].arena_bytes) {
        for (auto &q : t->ops) { d.staging.release(q.stage); q.stage = nullptr; }
        for (int p = 0; p < 3; p++) d.staging.release(t->src_stage[p]);
        return fail(IPG_ERR_NOMEM, "image does not fit a lane's device arena (raise lane_device_bytes)");
    }
    t->id = c->next_id.fetch_add(1);
    {
        std::lock_guard<std::mutex> lk(c->tmu);
        c->tickets[t->id] = t;
    }
    d.outstanding.fetch_add(t->cost);
    {
        std::lock_guard<std::mutex> lk(d.mu);
        d.queue.push_back(t);
    }
    d.cv_q.notify_all();
    *out = t->id;
    return IPG_OK;
}

static int wait_impl(Ctx *c, ipg_ticket id, int timeout_ms)
{
    if (!c) return fail(IPG_ERR_INVALID, "null context");
    TicketP t;
    {
        std::lock_guard<std::mutex> lk(c->tmu);
        auto it = c->tickets.find(id);
        if (it == c->tickets.end()) return fail(IPG_ERR_INVALID, "unknown or already consumed ticket");
        t = it->second;
    }
    {
        std::unique_lock<std::mutex> lk(t->mu);
        if (timeout_ms < 0) {
            t->cv.wait(lk, [&] { return t->done; });
        } else if (!t->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return t->done; })) {
            return fail(IPG_ERR_TIMEOUT, "timed out waiting for ticket");
        }
    }
    {
        std::lock_guard<std::mutex> lk(c->tmu);
        if (c->tickets.erase(id) == 0) return fail(IPG_ERR_INVALID, "ticket consumed by another waiter");
    }
    Device &d = *c->devs[t->dev];
    for (auto &r : t->ops) {
        if (!r.stage) continue;
        if (t->status == IPG_OK) {
            if (r.patch_only) {
                const size_t bw4 = (size_t)(r.bx1 - r.bx0) * 4;
                for (int y = r.by0; y < r.by1; y++)
                    memcpy((uint8_t *)r.dst + (size_t)y * r.dst_stride + (size_t)r.bx0 * 4, r.stage + (size_t)(y - r.by0) * bw4, bw4);
            } else {
                for (int y = 0; y < r.dh; y++)
                    memcpy((uint8_t *)r.dst + (size_t)y * r.dst_stride, r.stage + (size_t)y * r.dw * 4, (size_t)r.dw * 4);
            }
            c->s_staged++;
        }
        d.staging.release(r.stage);
        r.stage = nullptr;
    }
    if (t->status != IPG_OK) return fail(t->status, t->err);
    return IPG_OK;
}

static void destroy_impl(Ctx *c)
{
    if (!c) return;
    c->stop.store(true);
    for (auto &dp : c->devs) {
        Device &d = *dp;
        d.cv_q.notify_all();
        d.cv_lane.notify_all();
        d.cv_idle.notify_all();
        if (d.batcher.joinable()) d.batcher.join();
        // let the completer retire what is in flight
        for (;;) {
            {
                std::lock_guard<std::mutex> lk(d.mu);
                if (d.inflight.empty()) break;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        d.cv_idle.notify_all();
        if (d.completer.joinable()) d.completer.join();
        // anything still queued never ran
        for (auto &t : d.queue) finish_ticket(*c, d, t, IPG_ERR_SHUTDOWN, "context destroyed");
        d.queue.clear();
        cudaSetDevice(d.cuda_id);
        for (auto &L : d.lanes) {
            if (L.st) cudaStreamSynchronize(L.st);
            for (auto &e : L.ev) if (e) cudaEventDestroy(e);
            if (L.evf) cudaEventDestroy(L.evf);
            if (L.begin) cudaEventDestroy(L.begin);
            if (L.uploaded) cudaEventDestroy(L.uploaded);
            if (L.done) cudaEventDestroy(L.done);
            if (L.dl_begin) cudaEventDestroy(L.dl_begin);
            if (L.arena) cudaFree(L.arena);
            if (L.param_host) cudaFreeHost(L.param_host);
            if (L.fix_count_host) cudaFreeHost(L.fix_count_host);
            if (L.jpeg_result_host) cudaFreeHost(L.jpeg_result_host);
            if (L.done2) cudaEventDestroy(L.done2);
            if (L.st) cudaStreamDestroy(L.st);
            if (L.st2) cudaStreamDestroy(L.st2);
            if (L.fork) cudaEventDestroy(L.fork);
            if (L.join) cudaEventDestroy(L.join);
        }
        if (d.up) { cudaStreamSynchronize(d.up); cudaStreamDestroy(d.up); }
        if (d.down) { cudaStreamSynchronize(d.down); cudaStreamDestroy(d.down); }
        if (d.down2) { cudaStreamSynchronize(d.down2); cudaStreamDestroy(d.down2); }
        if (d.epoch) cudaEventDestroy(d.epoch);
        d.staging.destroy();
    }
    for (void *p : c->pinned.drain()) cudaFreeHost(p);
    delete static_cast<ipg_ctx *>(c);
}

} // namespace ipg

// =================================================================================
// C ABI
// =================================================================================
using namespace ipg;

extern "C" {

int ipg_abi_version(void) { return IPG_ABI_VERSION; }
const char *ipg_last_error(void) { return g_err.c_str(); }

int ipg_init(const int *device_ids, int n, const ipg_config *cfg, ipg_ctx **out)
{
    if (!out) return fail(IPG_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(IPG_ERR_NO_DEVICE, std::string("no CUDA device available (there is no CPU fallback): ") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    ipg_ctx *c = nullptr;
    try {
        c = new ipg_ctx;
        ipg_config k{};
        if (cfg) memcpy(&k, cfg, std::min<size_t>(sizeof k, cfg->struct_size ? cfg->struct_size : sizeof k));
        if (k.precision < IPG_PRECISION_EXACT || k.precision > IPG_PRECISION_REFERENCE) k.precision = IPG_PRECISION_EXACT;
        if (k.lanes_per_device <= 0) k.lanes_per_device = 3;
        if (k.lanes_per_device > 16) k.lanes_per_device = 16;
        if (k.max_batch <= 0) k.max_batch = 16;
        if (!cfg || k.batch_window_us < 0) k.batch_window_us = 200;
        if (k.lane_device_bytes == 0) k.lane_device_bytes = 1ull << 30;
        if (k.lane_pinned_bytes == 0) k.lane_pinned_bytes = 256ull << 20;
        c->cfg = k;
        c->trace = getenv("IPG_TRACE") && atoi(getenv("IPG_TRACE")) != 0;
        c->fuse_targets = (k.fuse_targets >= 1 && k.fuse_targets <= 3) ? k.fuse_targets : 1;
        if (getenv("IPG_MERGE_LEAN")) c->merge_lean = atoi(getenv("IPG_MERGE_LEAN")) != 0;
        if (getenv("IPG_VINT")) c->use_vint = atoi(getenv("IPG_VINT")) != 0;
        if (getenv("IPG_BAND_CTAS")) c->band_cta_target = (uint32_t)std::max(1, atoi(getenv("IPG_BAND_CTAS")));
        if (getenv("IPG_FIX_CAPACITY")) c->fix_capacity = (uint32_t)std::max(1, atoi(getenv("IPG_FIX_CAPACITY")));
        c->overlap_streams = !(getenv("IPG_NO_OVERLAP") && atoi(getenv("IPG_NO_OVERLAP")) != 0);
        if (getenv("IPG_DIRECT")) c->use_direct = std::min(2, std::max(0, atoi(getenv("IPG_DIRECT"))));
        if (getenv("IPG_OVERLAP_TAIL")) c->overlap_tail = atoi(getenv("IPG_OVERLAP_TAIL")) != 0;
        if (getenv("IPG_STAGING_TIMEOUT_MS")) c->staging_timeout_ms = std::max(0, atoi(getenv("IPG_STAGING_TIMEOUT_MS")));
        if (getenv("IPG_FUSE_TARGETS")) c->fuse_targets = std::min(3, std::max(1, atoi(getenv("IPG_FUSE_TARGETS"))));
        std::vector<int> ids;
        if (device_ids && n > 0) ids.assign(device_ids, device_ids + n);
        else for (int i = 0; i < count; i++) ids.push_back(i);
        for (size_t i = 0; i < ids.size(); i++) {
            if (ids[i] < 0 || ids[i] >= count) throw std::runtime_error("device id out of range");
            cudaDeviceProp prop;
            IPG_CU(cudaGetDeviceProperties(&prop, ids[i]));
            if (prop.major < 10)
                throw std::runtime_error(std::string("device ") + prop.name + " is not sm_100-class; this library is built for sm_100a only");
            std::unique_ptr<Device> d(new Device);
            d->ctx = c;
            d->index = (int)i;
            d->cuda_id = ids[i];
            IPG_CU(cudaSetDevice(ids[i]));
            d->lanes.resize((size_t)k.lanes_per_device);
            const size_t param_cap = 32u << 20;
            if (k.lane_device_bytes < param_cap + (128u << 20)) throw std::runtime_error("lane_device_bytes too small (< 160 MiB)");
            for (auto &L : d->lanes) {
                IPG_CU(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
                IPG_CU(cudaStreamCreateWithFlags(&L.st2, cudaStreamNonBlocking));
                IPG_CU(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
                IPG_CU(cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming));
                for (auto &ev : L.ev) IPG_CU(cudaEventCreate(&ev));
                IPG_CU(cudaEventCreate(&L.evf));
                IPG_CU(cudaEventCreate(&L.begin));
                IPG_CU(cudaEventCreate(&L.uploaded));
                IPG_CU(cudaEventCreate(&L.done));
                IPG_CU(cudaEventCreate(&L.dl_begin));
                L.arena_bytes = (size_t)k.lane_device_bytes;
                IPG_CU(cudaMalloc((void **)&L.arena, L.arena_bytes));
                L.param_cap = param_cap;
                IPG_CU(cudaHostAlloc((void **)&L.param_host, L.param_cap, cudaHostAllocPortable));
                IPG_CU(cudaHostAlloc((void **)&L.fix_count_host, 64, cudaHostAllocPortable));
                IPG_CU(cudaHostAlloc((void **)&L.jpeg_result_host, 16 * kJpegMaxJobs, cudaHostAllocPortable));
                IPG_CU(cudaEventCreate(&L.done2));
            }
            if (!d->staging.init((size_t)k.lane_pinned_bytes * (size_t)k.lanes_per_device))
                throw std::runtime_error("pinned staging allocation failed");
            IPG_CU(cudaStreamCreateWithFlags(&d->up, cudaStreamNonBlocking));
            IPG_CU(cudaStreamCreateWithFlags(&d->down, cudaStreamNonBlocking));
            IPG_CU(cudaStreamCreateWithFlags(&d->down2, cudaStreamNonBlocking));
            IPG_CU(cudaEventCreate(&d->epoch));
            IPG_CU(cudaEventRecord(d->epoch, d->lanes[0].st));
            IPG_CU(cudaEventSynchronize(d->epoch));
            c->devs.push_back(std::move(d));
        }
        for (auto &d : c->devs) {
            d->batcher = std::thread(batcher_main, c, d.get());
            d->completer = std::thread(completer_main, c, d.get());
        }
    } catch (const std::exception &ex) {
        std::string m = ex.what();
        if (c) destroy_impl(c);
        cudaGetLastError();
        return fail(IPG_ERR_CUDA, m);
    }
    *out = c;
    return IPG_OK;
}

void ipg_destroy(ipg_ctx *ctx) { destroy_impl(ctx); }

int ipg_device_count(const ipg_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

void *ipg_alloc_pinned(ipg_ctx *ctx, size_t bytes)
{
    if (!ctx || bytes == 0) { fail(IPG_ERR_INVALID, "null context or zero size"); return nullptr; }
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); fail(IPG_ERR_NOMEM, cuda_msg("cudaHostAlloc", e)); return nullptr; }
    ctx->pinned.add(p, bytes);
    return p;
}

void ipg_free_pinned(ipg_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    if (ctx->pinned.remove(p)) cudaFreeHost(p);
}

void *ipg_alloc_device(ipg_ctx *ctx, int device_index, size_t bytes)
{
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size() || bytes == 0) {
        fail(IPG_ERR_INVALID, "bad context/device/size");
        return nullptr;
    }
    cudaSetDevice(ctx->devs[device_index]->cuda_id);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); fail(IPG_ERR_NOMEM, cuda_msg("cudaMalloc", e)); return nullptr; }
    return p;
}

void ipg_free_device(ipg_ctx *ctx, int device_index, void *p)
{
    if (!ctx || !p || device_index < 0 || device_index >= (int)ctx->devs.size()) return;
    cudaSetDevice(ctx->devs[device_index]->cuda_id);
    cudaFree(p);
}

int ipg_copy_to_device(ipg_ctx *ctx, int device_index, void *dst, const void *src, size_t bytes)
{
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(IPG_ERR_INVALID, "bad context/device");
    cudaSetDevice(ctx->devs[device_index]->cuda_id);
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(IPG_ERR_CUDA, cuda_msg("cudaMemcpy H2D", e)); }
    return IPG_OK;
}

int ipg_copy_from_device(ipg_ctx *ctx, int device_index, void *dst, const void *src, size_t bytes)
{
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(IPG_ERR_INVALID, "bad context/device");
    cudaSetDevice(ctx->devs[device_index]->cuda_id);
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(IPG_ERR_CUDA, cuda_msg("cudaMemcpy D2H", e)); }
    return IPG_OK;
}

int ipg_submit(ipg_ctx *ctx, const ipg_image_desc *src, const ipg_op *ops, int n_ops, ipg_ticket *ticket)
{
    try {
        return submit_impl(ctx, -1, src, ops, n_ops, ticket);
    } catch (const std::exception &e) {
        return fail(IPG_ERR_INTERNAL, e.what());
    }
}

int ipg_submit_on(ipg_ctx *ctx, int device_index, const ipg_image_desc *src, const ipg_op *ops, int n_ops,
                  ipg_ticket *ticket)
{
    if (device_index < 0) return fail(IPG_ERR_INVALID, "device index out of range");
    try {
        return submit_impl(ctx, device_index, src, ops, n_ops, ticket);
    } catch (const std::exception &e) {
        return fail(IPG_ERR_INTERNAL, e.what());
    }
}

int ipg_wait(ipg_ctx *ctx, ipg_ticket ticket, int timeout_ms)
{
    try {
        return wait_impl(ctx, ticket, timeout_ms);
    } catch (const std::exception &e) {
        return fail(IPG_ERR_INTERNAL, e.what());
    }
}

int ipg_flush(ipg_ctx *ctx)
{
    if (!ctx) return fail(IPG_ERR_INVALID, "null context");
    for (auto &dp : ctx->devs) {
        Device &d = *dp;
        std::unique_lock<std::mutex> lk(d.mu);
        d.cv_lane.wait(lk, [&] {
            if (!d.queue.empty() || !d.inflight.empty()) return false;
            for (auto &L : d.lanes)
                if (L.busy) return false;
            return true;
        });
    }
    return IPG_OK;
}

int ipg_get_stats(ipg_ctx *ctx, ipg_stats *out)
{
    if (!ctx || !out) return fail(IPG_ERR_INVALID, "null argument");
    memset(out, 0, sizeof *out);
    out->tickets_done = ctx->s_done.load();
    out->batches = ctx->s_batches.load();
    out->kernels_launched = ctx->s_kernels.load();
    out->bytes_h2d = ctx->s_h2d.load();
    out->bytes_d2h = ctx->s_d2h.load();
    out->exact_fixups = ctx->s_fix.load();
    out->exact_fallbacks = ctx->s_fallback.load();
    out->staged_copies = ctx->s_staged.load();
    std::lock_guard<std::mutex> lk(ctx->smu);
    out->stream_kernel_ms = ctx->s_stream_ms;
    out->fix_kernel_ms = ctx->s_fix_ms;
    out->other_kernel_ms = ctx->s_other_ms;
    out->stream_fast_kernel_ms = ctx->s_fast_ms;
    out->fast_jobs = ctx->s_fast_jobs.load();
    out->h2d_ms = ctx->s_h2d_ms;
    out->d2h_ms = ctx->s_d2h_ms;
    out->kernel_ms = ctx->s_stream_ms + ctx->s_fix_ms + ctx->s_other_ms;
    for (auto &d : ctx->devs) {
        if (d->k_last >= 0) out->kernel_span_ms = std::max(out->kernel_span_ms, d->k_last - d->k_first);
        if (d->b_last >= 0) out->batch_span_ms = std::max(out->batch_span_ms, d->b_last - d->b_first);
    }
    return IPG_OK;
}

int ipg_reset_stats(ipg_ctx *ctx)
{
    if (!ctx) return fail(IPG_ERR_INVALID, "null context");
    int rc = ipg_flush(ctx);
    if (rc) return rc;
    ctx->s_done = 0; ctx->s_batches = 0; ctx->s_kernels = 0; ctx->s_h2d = 0; ctx->s_d2h = 0;
    ctx->s_fix = 0; ctx->s_fallback = 0; ctx->s_staged = 0;
    std::lock_guard<std::mutex> lk(ctx->smu);
    ctx->s_stream_ms = ctx->s_fix_ms = ctx->s_other_ms = ctx->s_fast_ms = ctx->s_h2d_ms = ctx->s_d2h_ms = 0;
    ctx->s_fast_jobs = 0;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d->cuda_id);
        cudaEventRecord(d->epoch, d->lanes[0].st);
        cudaEventSynchronize(d->epoch);
        d->k_first = d->b_first = 1e300;
        d->k_last = d->b_last = -1;
    }
    return IPG_OK;
}

/* internal/usecase/processor/operations/resize.go:63-72 */
void ipg_keep_aspect_dims(int ow, int oh, int w, int h, int *nw, int *nh)
{
    const double wr = (double)w / (double)ow;
    const double hr = (double)h / (double)oh;
    const double ratio = wr < hr ? wr : hr; // math.Min (no NaNs here)
    *nw = (int)((double)ow * ratio);
    *nh = (int)((double)oh * ratio);
}

/* operations/thumbnail.go:52-63 */
void ipg_thumb_fit_dims(int ow, int oh, int size, int *nw, int *nh)
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
void ipg_crop_square(int ow, int oh, int *cx, int *cy, int *cs)
{
    if (ow > oh) {
        *cs = oh; *cx = (ow - oh) / 2; *cy = 0;
    } else {
        *cs = ow; *cx = 0; *cy = (oh - ow) / 2;
    }
}

} // extern "C"
