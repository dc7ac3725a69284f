// host.cpp -- host layer of libipgpu.so (include/ipgpu_host.h): a C++ mirror of the
// reference's processor package around the raster C ABI.  No pixel arithmetic happens
// here; it is parameter handling, geometry, glyph layout, object paths and the
// task/result JSON, each following the reference file:line quoted at the function.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ipgpu_host.h"

namespace {

thread_local std::string g_herr;

// ---------------------------------------------------------------------------------
// a small JSON value (the broker message is encoding/json output of domain.ProcessingTask)
// ---------------------------------------------------------------------------------
struct JV {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JV> arr;
    std::vector<std::pair<std::string, JV>> obj;
    const JV *get(const char *key) const
    {
        if (kind != Obj) return nullptr;
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

static void utf8_append(std::string &s, uint32_t cp)
{
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
}

struct JParser {
    const char *p, *end;
    std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool fail(const char *m) { if (err.empty()) err = m; return false; }
    bool hex4(uint32_t &v)
    {
        if (end - p < 4) return fail("unexpected end of JSON input");
        v = 0;
        for (int i = 0; i < 4; i++, p++) {
            int c = *p, d;
            if (c >= '0' && c <= '9') d = c - '0';
            else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
            else return fail("invalid character in \\u escape");
            v = v * 16 + (uint32_t)d;
        }
        return true;
    }
    bool string(std::string &out)
    {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("unexpected end of JSON input");
                char c = *p++;
                switch (c) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    uint32_t cp;
                    if (!hex4(cp)) return false;
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        const char *save = p;
                        p += 2;
                        uint32_t lo;
                        if (!hex4(lo)) return false;
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else { p = save; cp = 0xFFFD; }
                    } else if (cp >= 0xD800 && cp < 0xE000) cp = 0xFFFD;
                    utf8_append(out, cp);
                    break;
                }
                default: return fail("invalid escape in string");
                }
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unexpected end of JSON input");
        p++;
        return true;
    }
    bool value(JV &v, int depth = 0)
    {
        if (depth > 64) return fail("JSON nested too deeply");
        ws();
        if (p >= end) return fail("unexpected end of JSON input");
        if (*p == '{') {
            v.kind = JV::Obj;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':' after object key");
                p++;
                JV c;
                if (!value(c, depth + 1)) return false;
                v.obj.emplace_back(std::move(k), std::move(c));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; return true; }
                return fail("expected ',' or '}' in object");
            }
        }
        if (*p == '[') {
            v.kind = JV::Arr;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return true; }
            for (;;) {
                JV c;
                if (!value(c, depth + 1)) return false;
                v.arr.push_back(std::move(c));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; return true; }
                return fail("expected ',' or ']' in array");
            }
        }
        if (*p == '"') { v.kind = JV::Str; return string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JV::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JV::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JV::Null; p += 4; return true; }
        if (*p == '-' || (*p >= '0' && *p <= '9')) {
            char *e = nullptr;
            std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
            v.num = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) return fail("invalid number");
            p += e - tmp.c_str();
            v.kind = JV::Num;
            return true;
        }
        return fail("invalid character looking for beginning of value");
    }
};

static bool json_parse(const char *text, JV &out, std::string &err)
{
    JParser P{text, text + strlen(text), ""};
    if (!P.value(out)) { err = P.err; return false; }
    P.ws();
    if (P.p != P.end) { err = "invalid character after top-level value"; return false; }
    return true;
}

// encoding/json string encoding (HTML-safe escaping on, as json.Marshal does)
static void json_quote(std::string &o, const std::string &s)
{
    static const char *hex = "0123456789abcdef";
    o += '"';
    for (size_t i = 0; i < s.size();) {
        unsigned char c = (unsigned char)s[i];
        if (c < 0x80) {
            if (c == '"' || c == '\\') { o += '\\'; o += (char)c; }
            else if (c == '\n') o += "\\n";
            else if (c == '\r') o += "\\r";
            else if (c == '\t') o += "\\t";
            else if (c < 0x20 || c == '<' || c == '>' || c == '&') { o += "\\u00"; o += hex[c >> 4]; o += hex[c & 15]; }
            else o += (char)c;
            i++;
            continue;
        }
        // validate one UTF-8 sequence; invalid bytes become U+FFFD like encoding/json
        int n = (c >= 0xF0 && c < 0xF8) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2) ? 2 : 0;
        bool ok = n > 0 && i + (size_t)n <= s.size();
        uint32_t cp = 0;
        if (ok) {
            cp = c & (0xFF >> (n + 1));
            for (int k = 1; k < n; k++) {
                unsigned char d = (unsigned char)s[i + k];
                if ((d & 0xC0) != 0x80) { ok = false; break; }
                cp = (cp << 6) | (d & 0x3F);
            }
            if (ok && ((n == 3 && cp < 0x800) || (n == 4 && (cp < 0x10000 || cp > 0x10FFFF)) || (cp >= 0xD800 && cp < 0xE000))) ok = false;
        }
        if (!ok) { o += "\\ufffd"; i++; continue; }
        if (cp == 0x2028 || cp == 0x2029) { o += "\\u202"; o += hex[cp & 15]; }
        else o.append(s, i, (size_t)n);
        i += (size_t)n;
    }
    o += '"';
}

// ---------------------------------------------------------------------------------
// domain (internal/domain/task.go:3-23, image.go:43-49)
// ---------------------------------------------------------------------------------
struct Operation {
    std::string type;
    JV params; // map[string]interface{}: numbers are float64 after the JSON round trip
};
struct Task {
    std::string id, image_id, original_path, bucket, format;
    std::vector<Operation> ops;
};
struct Result {
    std::string id, image_id, status = "completed", error;
    std::map<std::string, std::string> paths; // encoding/json sorts map keys
    std::string marshal() const
    {
        std::string o = "{\"ID\":";
        json_quote(o, id);
        o += ",\"ImageID\":";
        json_quote(o, image_id);
        o += ",\"Status\":";
        json_quote(o, status);
        o += ",\"ProcessedPaths\":{";
        bool first = true;
        for (auto &kv : paths) {
            if (!first) o += ',';
            first = false;
            json_quote(o, kv.first);
            o += ':';
            json_quote(o, kv.second);
        }
        o += "},\"Error\":";
        json_quote(o, error);
        o += '}';
        return o;
    }
};

static std::string jstr(const JV *v) { return v && v->kind == JV::Str ? v->str : std::string(); }

static bool parse_task(const char *text, Task &t, std::string &err)
{
    JV root;
    if (!text) { err = "unexpected end of JSON input"; return false; }
    if (!json_parse(text, root, err)) return false;
    if (root.kind != JV::Obj) { err = "json: cannot unmarshal non-object into Go value of type domain.ProcessingTask"; return false; }
    t.id = jstr(root.get("ID"));
    t.image_id = jstr(root.get("ImageID"));
    t.original_path = jstr(root.get("OriginalPath"));
    t.bucket = jstr(root.get("Bucket"));
    t.format = jstr(root.get("Format"));
    if (const JV *ops = root.get("Operations")) {
        if (ops->kind == JV::Arr) {
            for (auto &o : ops->arr) {
                Operation op;
                op.type = jstr(o.get("Type"));
                if (const JV *p = o.get("Parameters")) op.params = *p;
                t.ops.push_back(std::move(op));
            }
        }
    }
    return true;
}

// Go's int(float64): truncation toward zero (out-of-range is implementation-defined; clamp)
static int go_int(double v)
{
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int)-2147483647 - 1;
    return (int)v;
}
static bool param_num(const JV &params, const char *key, double &out)
{
    const JV *v = params.get(key);
    if (!v || v->kind != JV::Num) return false;
    out = v->num;
    return true;
}
static bool param_bool(const JV &params, const char *key)
{
    const JV *v = params.get(key);
    return v && v->kind == JV::Bool && v->b;
}
static bool param_str(const JV &params, const char *key, std::string &out)
{
    const JV *v = params.get(key);
    if (!v || v->kind != JV::Str) return false;
    out = v->str;
    return true;
}
static std::string lower(std::string s)
{
    for (auto &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

// image_processor.go:129-162
static std::string generate_path(const std::string &image_id, const std::string &op, const std::string &format, const JV &params)
{
    const std::string base = "processed/";
    char buf[64];
    if (op == "resize") {
        double w = 0, h = 0;
        int wi = param_num(params, "width", w) ? go_int(w) : 0;
        int hi = param_num(params, "height", h) ? go_int(h) : 0;
        snprintf(buf, sizeof buf, "%dx%d", wi, hi);
        return base + "resize/" + image_id + "/" + buf + "." + format;
    }
    if (op == "thumbnail") {
        double s = 0;
        int size = param_num(params, "size", s) ? go_int(s) : 0;
        if (size == 0) size = 200; // domain.DefaultThumbnailSize
        snprintf(buf, sizeof buf, "%d", size);
        return base + "thumbnails/" + image_id + "/" + buf + "." + format;
    }
    if (op == "watermark") return base + "watermarked/" + image_id + "/watermarked." + format;
    return base + lower(op) + "/" + image_id + "/processed." + format;
}

// image_processor.go:164-182 (filepath.Ext: from the last '.' of the last path element)
static const char *content_type(const std::string &path)
{
    size_t slash = path.find_last_of('/');
    size_t dot = path.find_last_of('.');
    std::string ext;
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) ext = lower(path.substr(dot));
    if (ext == ".jpg" || ext == ".jpeg") return "image/jpeg";
    if (ext == ".png") return "image/png";
    if (ext == ".gif") return "image/gif";
    if (ext == ".webp") return "image/webp";
    if (ext == ".bmp") return "image/bmp";
    if (ext == ".tiff" || ext == ".tif") return "image/tiff";
    return "image/jpeg";
}

// strconv.Atoi
static bool go_atoi(const std::string &s, long long &out)
{
    size_t i = 0;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) i++;
    if (i >= s.size()) return false;
    long long v = 0;
    for (size_t k = i; k < s.size(); k++) {
        if (s[k] < '0' || s[k] > '9') return false;
        if (v > (9223372036854775807LL - (s[k] - '0')) / 10) return false; // out of range
        v = v * 10 + (s[k] - '0');
    }
    out = s[0] == '-' ? -v : v;
    return true;
}
static int clampi(long long v, int lo, int hi) { return (int)std::max<long long>(lo, std::min<long long>(hi, v)); }

// watermark.go:159-186, with the caller's fallback to black (:94-97) folded in
static int parse_color(const std::string &in, double opacity, uint8_t rgba[4])
{
    const uint8_t a_op = (uint8_t)(long long)(255 * opacity); // uint8(255 * opacity): truncation
    std::string s;
    for (char c : in)
        if (c != ' ') s += c;
    std::vector<std::string> parts;
    size_t pos = 0;
    for (;;) {
        size_t c = s.find(',', pos);
        parts.push_back(s.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
        if (c == std::string::npos) break;
        pos = c + 1;
    }
    long long r, g, b;
    if ((parts.size() != 3 && parts.size() != 4) || !go_atoi(parts[0], r) || !go_atoi(parts[1], g) || !go_atoi(parts[2], b)) {
        rgba[0] = rgba[1] = rgba[2] = 0;
        rgba[3] = a_op;
        return -1;
    }
    rgba[0] = (uint8_t)clampi(r, 0, 255);
    rgba[1] = (uint8_t)clampi(g, 0, 255);
    rgba[2] = (uint8_t)clampi(b, 0, 255);
    rgba[3] = a_op;
    long long av;
    if (parts.size() == 4 && go_atoi(parts[3], av)) rgba[3] = (uint8_t)clampi(av, 0, 255);
    return 0;
}

// watermark.go:116,118: fixed.Int26_6(fontSize*64*1.2).Ceil()
static int watermark_height_px(double font_size)
{
    const int32_t fx = (int32_t)(font_size * 64 * 1.2);
    return (fx + 63) >> 6;
}

// watermark.go:121-148 (Go integer division truncates toward zero)
static void watermark_anchor(const std::string &position, int W, int H, int wpx, int hpx, int *x, int *y)
{
    const int m = 20;
    if (position == "top-left") { *x = m; *y = m + hpx; }
    else if (position == "top-right") { *x = W - wpx - m; *y = m + hpx; }
    else if (position == "top-center") { *x = (W - wpx) / 2; *y = m + hpx; }
    else if (position == "bottom-left") { *x = m; *y = H - m; }
    else if (position == "bottom-center") { *x = (W - wpx) / 2; *y = H - m; }
    else if (position == "center") { *x = (W - wpx) / 2; *y = (H + hpx) / 2; }
    else { *x = W - wpx - m; *y = H - m; } // bottom-right and anything unknown
}

// `for _, x := range text`: UTF-8 decode, invalid bytes yield U+FFFD one byte at a time
static std::vector<uint32_t> runes_of(const std::string &s)
{
    std::vector<uint32_t> out;
    for (size_t i = 0; i < s.size();) {
        unsigned char c = (unsigned char)s[i];
        if (c < 0x80) { out.push_back(c); i++; continue; }
        int n = (c >= 0xF0 && c < 0xF8) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2) ? 2 : 0;
        bool ok = n > 0 && i + (size_t)n <= s.size();
        uint32_t cp = 0;
        if (ok) {
            cp = c & (0xFF >> (n + 1));
            for (int k = 1; k < n; k++) {
                unsigned char d = (unsigned char)s[i + k];
                if ((d & 0xC0) != 0x80) { ok = false; break; }
                cp = (cp << 6) | (d & 0x3F);
            }
            if (ok && ((n == 3 && cp < 0x800) || (n == 4 && (cp < 0x10000 || cp > 0x10FFFF)) || (cp >= 0xD800 && cp < 0xE000))) ok = false;
        }
        if (!ok) { out.push_back(0xFFFD); i++; continue; }
        out.push_back(cp);
        i += (size_t)n;
    }
    return out;
}

// ---------------------------------------------------------------------------------
// one planned operation of one image
// ---------------------------------------------------------------------------------
struct GlyphOwned {
    ipg_glyph g;
    std::shared_ptr<std::vector<uint8_t>> mask;
};
struct PlannedOp {
    std::string type, out_format, path, encode_err_prefix;
    int dw = 0, dh = 0;
    ipg_op op{};
    std::vector<GlyphOwned> glyphs;
    std::vector<ipg_glyph> glyph_arr;
    uint8_t *dst = nullptr; // pinned
    size_t dst_bytes = 0;
    bool jpeg_dev = false;  // the engine returns this result as the JPEG file (iph_set_device_jpeg): dst holds the file
    uint64_t *jpeg_len_ptr = nullptr; // ... of *this many bytes once the ticket is done (the tail of dst)
};

} // namespace

// Pinned output buffers come from a few large slabs (ipg_alloc_pinned once per slab) that are sub-allocated first-fit:
// a mixed-size stream asks for a differently sized watermark output per image, and one cudaHostAlloc / cudaFreeHost per
// image (each a device-synchronising driver call of ~0.3 ms per MB) would put the allocator, not the copy engines, on
// the critical path.  The pool is bounded: slabs that are entirely free are returned to the driver once the pool holds
// more than kPoolCapBytes, so one odd task cannot pin host memory for the life of the processor (the Go reference would
// have freed such a buffer at the next GC).
class PinnedSlabs {
public:
    static constexpr size_t kSlabBytes = (size_t)256 << 20;
    static constexpr size_t kPoolCapBytes = (size_t)2 << 30;
    uint8_t *take(ipg_ctx *ctx, size_t n)
    {
        n = (std::max<size_t>(n, 1) + 4095) & ~(size_t)4095;
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &s : slabs_)
            if (uint8_t *p = carve(s, n)) return p;
        if (!ctx) return nullptr;
        Slab s;
        s.bytes = std::max(n, kSlabBytes);
        s.base = (uint8_t *)ipg_alloc_pinned(ctx, s.bytes);
        if (!s.base && s.bytes > n) { // a whole slab did not fit: try the bare request
            s.bytes = n;
            s.base = (uint8_t *)ipg_alloc_pinned(ctx, s.bytes);
        }
        if (!s.base) return nullptr;
        s.free_[0] = s.bytes;
        total_ += s.bytes;
        slabs_.push_back(std::move(s));
        return carve(slabs_.back(), n);
    }
    void give(ipg_ctx *ctx, uint8_t *p)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < slabs_.size(); i++) {
            Slab &s = slabs_[i];
            if (p < s.base || p >= s.base + s.bytes) continue;
            const size_t off = (size_t)(p - s.base);
            auto u = s.used.find(off);
            if (u == s.used.end()) return;
            size_t n = u->second;
            s.used.erase(u);
            auto nx = s.free_.lower_bound(off);
            if (nx != s.free_.end() && off + n == nx->first) { n += nx->second; nx = s.free_.erase(nx); }
            if (nx != s.free_.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second == off) { pv->second += n; n = 0; }
            }
            if (n) s.free_[off] = n;
            if (s.used.empty() && total_ > kPoolCapBytes) { // shrink: this slab is idle and the pool is over its cap
                total_ -= s.bytes;
                if (ctx) ipg_free_pinned(ctx, s.base);
                slabs_.erase(slabs_.begin() + (long)i);
            }
            return;
        }
    }
    void release_all(ipg_ctx *ctx)
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &s : slabs_)
            if (ctx) ipg_free_pinned(ctx, s.base);
        slabs_.clear();
        total_ = 0;
    }
    size_t total_bytes()
    {
        std::lock_guard<std::mutex> lk(mu_);
        return total_;
    }

private:
    struct Slab {
        uint8_t *base = nullptr;
        size_t bytes = 0;
        std::map<size_t, size_t> free_, used; // offset -> bytes
    };
    static uint8_t *carve(Slab &s, size_t n)
    {
        for (auto it = s.free_.begin(); it != s.free_.end(); ++it) {
            if (it->second < n) continue;
            const size_t off = it->first, rem = it->second - n;
            s.free_.erase(it);
            if (rem) s.free_[off + n] = rem;
            s.used[off] = n;
            return s.base + off;
        }
        return nullptr;
    }
    std::mutex mu_;
    std::vector<Slab> slabs_;
    size_t total_ = 0;
};

struct iph_processor {
    ipg_ctx *ctx = nullptr;
    iph_callbacks cb{};
    PinnedSlabs pinned; // output buffers, reused across calls
    bool device_jpeg = false; // iph_set_device_jpeg
    int jpeg_first_div = 1;   // first-attempt file buffer = pixels / this + 4096 bytes (IPH_JPEG_FIRST_DIV: tests force the retry with it)
    uint8_t *take_pinned(size_t n) { return pinned.take(ctx, n); }
    void give_pinned(uint8_t *p) { pinned.give(ctx, p); }
};

namespace {

struct Job { // one ProcessingTask in flight
    Task task;
    Result res;
    std::string target_format;
    std::vector<PlannedOp> ops; // ops that passed parameter handling, in task order
    std::string plan_err;       // the error of the first op that did not (Process stops there)
    std::string plan_err_type;
    bool submitted = false;
    ipg_ticket ticket = 0;
    std::string fatal;          // error returned by Process
    const ipg_image_desc *img = nullptr; // the decoded image (the caller's, valid for the duration of the call)
};

static void format_switch(const std::string &format, bool watermark, std::string &out_format)
{
    // resize.go:78-91 / thumbnail.go:68-81 / watermark.go:66-79
    const std::string f = lower(format);
    if (f == "jpg" || f == "jpeg") out_format = "jpeg";
    else if (f == "png") out_format = "png";
    else if (f == "gif") out_format = watermark ? "jpeg" : "gif";
    else out_format = "jpeg";
}

// freetype.Context.DrawString (freetype.go) for one text: per-rune DrawMask rectangles.
static bool layout_text(iph_processor *P, const std::string &text, double font_size, int W, int H, int px, int py,
                        PlannedOp &po, std::string &err)
{
    int32_t pen_x = px * 64, pen_y = py * 64; // freetype.Pt
    uint32_t prev = 0;
    bool has_prev = false;
    // Context.glyph() (golang/freetype freetype.go): a direct-mapped cache of nGlyphs(256) x nXFractions(4) x
    // nYFractions(1) slots, slot = ((index % 256) * 4 + fx / 16) * 1 + fy / 64; a hit needs the SAME glyph index in
    // the slot (the sub-pixel offset is implied by the slot: the first occurrence's mask is reused for the whole
    // quarter-pixel bucket); a miss rasterises at this occurrence's (fx, fy) and OVERWRITES the slot.  The Context is
    // made per addTextWatermark call (watermark.go:98), so the cache starts empty for every text.  Runes that map to
    // one glyph index (e.g. all runes the face lacks -> .notdef) share masks, exactly as there.
    struct Cached { bool valid = false; uint32_t index = 0; iph_glyph g{}; std::shared_ptr<std::vector<uint8_t>> mask; };
    std::map<int, Cached> cache; // slot -> entry (sparse stand-in for the 1024-entry array)
    for (uint32_t r : runes_of(text)) {
        if (has_prev && P->cb.kern) pen_x += P->cb.kern(P->cb.user, prev, r, font_size);
        const int ix = pen_x >> 6, fx = pen_x & 63, iy = pen_y >> 6, fy = pen_y & 63;
        // Font.Index(rune); hosts without the callback get the rune itself as a stand-in (distinct runes then never
        // share a mask and collide only when equal mod 256)
        const uint32_t index = P->cb.glyph_index ? P->cb.glyph_index(P->cb.user, r) : r;
        const int slot = ((int)(index % 256u) * 4 + fx / 16) * 1 + fy / 64;
        Cached &e = cache[slot];
        if (!e.valid || e.index != index) {
            iph_glyph g{};
            if (!P->cb.glyph_mask || P->cb.glyph_mask(P->cb.user, r, font_size, fx, fy, &g) != 0) {
                err = "failed to draw watermark text: glyph rasterisation failed";
                return false;
            }
            Cached c;
            c.valid = true;
            c.index = index;
            c.g = g;
            c.mask = std::make_shared<std::vector<uint8_t>>();
            if (g.mask && g.mask_w > 0 && g.mask_h > 0) {
                c.mask->resize((size_t)g.mask_w * (size_t)g.mask_h);
                for (int y = 0; y < g.mask_h; y++)
                    memcpy(c.mask->data() + (size_t)y * g.mask_w, g.mask + (size_t)y * g.mask_stride, (size_t)g.mask_w);
            }
            c.g.mask = nullptr;
            c.g.mask_stride = g.mask_w;
            e = std::move(c);
        }
        const Cached &c = e;
        if (c.g.mask_w > 0 && c.g.mask_h > 0) {
            // glyphRect = mask.Bounds().Add(offset + (ix, iy)); dr = clip.Intersect(glyphRect)
            const int gx0 = ix + c.g.off_x, gy0 = iy + c.g.off_y;
            const int x0 = std::max(gx0, 0), y0 = std::max(gy0, 0);
            const int x1 = std::min(gx0 + c.g.mask_w, W), y1 = std::min(gy0 + c.g.mask_h, H);
            if (x0 < x1 && y0 < y1) {
                GlyphOwned go;
                go.mask = c.mask;
                go.g.x0 = x0; go.g.y0 = y0;
                // DrawMask(dst, dr, src, ZP, mask, mp = (0, dr.Min.Y - glyphRect.Min.Y), Over): mp.X is 0 even
                // when dr was clipped on the left; DrawMask then clips dr to the mask bounds from mp
                go.g.x1 = std::min(x1, x0 + c.g.mask_w);
                go.g.y1 = y1;
                go.g.mp_x = 0;
                go.g.mp_y = y0 - gy0;
                go.g.mask_w = c.g.mask_w; go.g.mask_h = c.g.mask_h; go.g.mask_stride = c.g.mask_w;
                go.g.mask = go.mask->data();
                po.glyphs.push_back(std::move(go));
            }
        }
        pen_x += c.g.advance_26_6;
        prev = r;
        has_prev = true;
    }
    return true;
}

enum { kMaxDstSide = 65536 };

// Resizer.Process / Thumbnailer.Process / Watermarker.Process up to the raster call:
// parameters, geometry.  Returns false with the operation's error text.
static bool plan_op(iph_processor *P, const Operation &o, const ipg_image_desc &img, const std::string &format,
                    PlannedOp &po, std::string &err)
{
    const int ow = img.width, oh = img.height;
    po.type = o.type;
    const bool gif = lower(format) == "gif";
    if (o.type == "resize") {
        double w, h;
        if (!param_num(o.params, "width", w)) { err = "width parameter is required and must be a number"; return false; }
        if (!param_num(o.params, "height", h)) { err = "height parameter is required and must be a number"; return false; }
        const int width = go_int(w), height = go_int(h);
        if (width <= 0 || height <= 0) { err = "width and height must be positive numbers"; return false; }
        po.dw = width;
        po.dh = height;
        if (param_bool(o.params, "keep_aspect")) ipg_keep_aspect_dims(ow, oh, width, height, &po.dw, &po.dh);
        po.op.kind = IPG_OP_RESIZE;
        format_switch(format, false, po.out_format);
        po.encode_err_prefix = gif ? "failed to encode gif: " : "failed to encode resized image: ";
    } else if (o.type == "thumbnail") {
        double s;
        int size = param_num(o.params, "size", s) ? go_int(s) : 200;
        if (size <= 0) { err = "size must be a positive number"; return false; }
        if (param_bool(o.params, "crop_to_fit")) {
            int cx, cy, cs;
            ipg_crop_square(ow, oh, &cx, &cy, &cs);
            po.op.kind = IPG_OP_THUMB_CROP;
            po.op.rect_x = cx; po.op.rect_y = cy; po.op.rect_w = cs; po.op.rect_h = cs;
            po.dw = po.dh = size;
        } else {
            po.op.kind = IPG_OP_RESIZE;
            ipg_thumb_fit_dims(ow, oh, size, &po.dw, &po.dh);
        }
        format_switch(format, false, po.out_format);
        po.encode_err_prefix = gif ? "failed to encode gif thumbnail: " : "failed to encode thumbnail: ";
    } else if (o.type == "watermark") {
        std::string text, position, font_color;
        double opacity, font_size;
        if (!param_str(o.params, "text", text) || text.empty()) text = "\xC2\xA9 ImageProcessor"; // "© ImageProcessor"
        if (!param_num(o.params, "opacity", opacity) || opacity <= 0) opacity = 0.5;
        if (!param_str(o.params, "position", position)) position = "bottom-right";
        if (!param_num(o.params, "font_size", font_size) || font_size <= 0) font_size = 36;
        if (!param_str(o.params, "font_color", font_color)) font_color = "255,255,255";
        po.op.kind = IPG_OP_WATERMARK;
        po.dw = ow;
        po.dh = oh;
        parse_color(font_color, opacity, po.op.color); // on a parse error: black at the same alpha
        int32_t text_width = 0;
        for (uint32_t r : runes_of(text)) {
            int32_t adv = 0;
            if (P->cb.glyph_advance && P->cb.glyph_advance(P->cb.user, r, font_size, &adv) == 0) text_width += adv;
        }
        const int width_px = (text_width + 63) >> 6, height_px = watermark_height_px(font_size);
        int px, py;
        watermark_anchor(position, ow, oh, width_px, height_px, &px, &py);
        if (!layout_text(P, text, font_size, ow, oh, px, py, po, err)) {
            err = "failed to add watermark: " + err;
            return false;
        }
        format_switch(format, true, po.out_format);
        po.encode_err_prefix = "failed to encode watermarked image: ";
    } else {
        err = "unsupported operation type: " + o.type;
        return false;
    }
    // The raster engine takes destinations up to 65536 px a side (ipg_submit rejects more): refuse here, before a
    // Kafka-supplied {"width": 60000, "height": 60000} reaches the pinned allocator.
    if (po.dw > kMaxDstSide || po.dh > kMaxDstSide) {
        err = "output dimensions " + std::to_string(po.dw) + "x" + std::to_string(po.dh) + " exceed the raster engine's limit of " +
              std::to_string(kMaxDstSide) + " px per side";
        return false;
    }
    po.path = generate_path("", o.type, po.out_format, o.params); // image id filled by the caller
    return true;
}

static void job_fail(Job &j, const std::string &result_error, const std::string &err)
{
    j.res.status = "failed";
    j.res.error = result_error;
    j.fatal = err;
}

// Destination buffers + ipg_submit of a planned task.  With iph_set_device_jpeg, results whose target format is JPEG
// are requested as files (IPG_LAYOUT_JPEG, quality 85 = domain/task.go:57): 1 byte per pixel of room is ample for
// photographs (noise needs 0.8-0.95); `roomy` (the retry after a file did not fit) gives 8, more than any scan can take.
static void submit_job(iph_processor *P, Job &j, bool roomy)
{
    std::vector<ipg_op> arr;
    for (auto &po : j.ops) {
        po.jpeg_dev = P->device_jpeg && po.out_format == "jpeg" && po.dw > 0 && po.dh > 0 && po.dw < 65536 && po.dh < 65536;
        const size_t px = (size_t)std::max(po.dw, 0) * (size_t)std::max(po.dh, 0);
        // a device-encoded file: room for the file, then 8 bytes for its length (the engine writes both after ipg_submit
        // returned, so both live in pinned memory)
        const size_t file_cap = ((roomy ? px * 8 : px / (size_t)P->jpeg_first_div) + 4096 + 7) & ~(size_t)7;
        po.dst_bytes = po.jpeg_dev ? file_cap + 8 : px * 4;
        if (po.dst_bytes) {
            po.dst = P->take_pinned(po.dst_bytes);
            if (!po.dst) {
                j.plan_err_type = po.type;
                j.plan_err = "failed to process operation " + po.type + ": " + ipg_last_error();
                break;
            }
        }
        po.op.dst_w = po.dw;
        po.op.dst_h = po.dh;
        po.op.dst = po.dst;
        po.op.dst_stride = (int32_t)((size_t)po.dw * 4); // dw <= 65536: fits
        po.op.dst_memspace = IPG_MEM_HOST;
        if (po.jpeg_dev) {
            po.op.dst_layout = IPG_LAYOUT_JPEG;
            po.op.jpeg_quality = 85;
            po.op.dst_capacity = file_cap;
            po.jpeg_len_ptr = (uint64_t *)(po.dst + file_cap);
            *po.jpeg_len_ptr = 0;
            po.op.dst_len = po.jpeg_len_ptr;
        }
        po.glyph_arr.clear();
        for (auto &g : po.glyphs) po.glyph_arr.push_back(g.g);
        po.op.n_glyphs = (int)po.glyph_arr.size();
        po.op.glyphs = po.glyph_arr.data();
        arr.push_back(po.op);
    }
    if (arr.size() != j.ops.size()) { // allocation failed part-way: nothing of this task runs
        for (auto &po : j.ops) { P->give_pinned(po.dst); po.dst = nullptr; }
        j.ops.clear();
        return;
    }
    if (ipg_submit(P->ctx, j.img, arr.data(), (int)arr.size(), &j.ticket) != IPG_OK) {
        j.plan_err_type = j.ops[0].type;
        j.plan_err = "failed to process operation " + j.ops[0].type + ": " + ipg_last_error();
        for (auto &po : j.ops) { P->give_pinned(po.dst); po.dst = nullptr; }
        j.ops.clear();
        return;
    }
    j.submitted = true;
}

// Process up to and including the submission of the raster work
static void job_begin(iph_processor *P, Job &j, const char *task_json, const ipg_image_desc *img, const char *decoded_format,
                      const char *decode_error)
{
    std::string perr;
    if (!parse_task(task_json, j.task, perr)) { // worker.go:167-171: the worker rejects the message before Process
        j.res.status = "failed";
        j.res.error = "failed to unmarshal task: " + perr;
        j.fatal = "failed to unmarshal task: " + perr;
        return;
    }
    j.res.id = j.task.id;
    j.res.image_id = j.task.image_id;
    if (!img) { // image_processor.go:47-53
        const std::string why = decode_error ? decode_error : "image: unknown format";
        job_fail(j, "Failed to decode image: " + why, "failed to decode image: " + why);
        return;
    }
    j.target_format = j.task.format.empty() ? std::string(decoded_format ? decoded_format : "") : j.task.format;
    for (auto &o : j.task.ops) {
        PlannedOp po;
        std::string err;
        if (!plan_op(P, o, *img, j.target_format, po, err)) {
            j.plan_err_type = o.type;
            // applyOperation wraps operation errors, but not the unsupported-type one (image_processor.go:104-127)
            j.plan_err = (err.rfind("unsupported operation type", 0) == 0) ? err : "failed to process operation " + o.type + ": " + err;
            break;
        }
        po.path = generate_path(j.task.image_id, o.type, po.out_format, o.params);
        j.ops.push_back(std::move(po));
    }
    if (j.ops.empty()) return;
    if (!P->ctx) {
        j.plan_err_type = j.ops[0].type;
        j.plan_err = "failed to process operation " + j.ops[0].type + ": no raster engine (there is no CPU fallback)";
        j.ops.clear();
        return;
    }
    j.img = img;
    submit_job(P, j, false);
}

// wait for the raster work, then encode + SaveProcessed per operation in task order
static void job_finish(iph_processor *P, Job &j)
{
    if (!j.fatal.empty()) return;
    std::string raster_err;
    if (j.submitted) {
        int rc = ipg_wait(P->ctx, j.ticket, -1);
        bool any_jpeg = false;
        for (auto &po : j.ops) any_jpeg |= po.jpeg_dev;
        if (rc == IPG_ERR_NOMEM && any_jpeg) { // a file did not fit its buffer (not a photograph): once more with room for any scan
            for (auto &po : j.ops) { P->give_pinned(po.dst); po.dst = nullptr; }
            j.submitted = false;
            submit_job(P, j, true);
            rc = j.submitted ? ipg_wait(P->ctx, j.ticket, -1) : IPG_OK;
        }
        if (rc != IPG_OK) raster_err = ipg_last_error();
    }
    for (auto &po : j.ops) {
        if (!j.fatal.empty()) break;
        if (!raster_err.empty()) {
            const std::string e = "failed to process operation " + po.type + ": " + raster_err;
            job_fail(j, "Operation " + po.type + " failed: " + e, "operation " + po.type + " failed: " + e);
            break;
        }
        if (po.jpeg_dev) { // jpeg.Encode already happened on the device: dst holds the file
            const int rc = P->cb.save_processed ? P->cb.save_processed(P->cb.user, po.path.c_str(), po.dst, (size_t)*po.jpeg_len_ptr, content_type(po.path)) : -1;
            if (rc != 0) {
                job_fail(j, "Failed to save processed image: save failed", "failed to save processed image: save failed");
                break;
            }
            j.res.paths[po.type] = po.path;
            continue;
        }
        uint8_t *enc = nullptr;
        size_t enc_len = 0;
        if (!P->cb.encode || P->cb.encode(P->cb.user, po.dst, po.dw, po.dh, po.dw * 4, po.out_format.c_str(), 85, &enc, &enc_len) != 0) {
            const std::string e = "failed to process operation " + po.type + ": " + po.encode_err_prefix + "encoder failed";
            job_fail(j, "Operation " + po.type + " failed: " + e, "operation " + po.type + " failed: " + e);
            break;
        }
        const int rc = P->cb.save_processed ? P->cb.save_processed(P->cb.user, po.path.c_str(), enc, enc_len, content_type(po.path)) : -1;
        if (P->cb.release) P->cb.release(P->cb.user, enc);
        if (rc != 0) {
            job_fail(j, "Failed to save processed image: save failed", "failed to save processed image: save failed");
            break;
        }
        j.res.paths[po.type] = po.path;
    }
    for (auto &po : j.ops) P->give_pinned(po.dst);
    if (j.fatal.empty() && !j.plan_err.empty())
        job_fail(j, "Operation " + j.plan_err_type + " failed: " + j.plan_err, "operation " + j.plan_err_type + " failed: " + j.plan_err);
}

static char *dup_cstr(const std::string &s)
{
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

} // namespace

extern "C" {

iph_processor *iph_processor_new(ipg_ctx *ctx, const iph_callbacks *cb)
{
    iph_processor *p = new (std::nothrow) iph_processor;
    if (!p) return nullptr;
    p->ctx = ctx;
    if (cb) p->cb = *cb;
    return p;
}

int iph_set_device_jpeg(iph_processor *p, int enabled)
{
    if (!p) return -1;
    p->device_jpeg = enabled != 0;
    if (const char *e = getenv("IPH_JPEG_FIRST_DIV")) p->jpeg_first_div = std::max(1, atoi(e));
    return 0;
}

void iph_processor_free(iph_processor *p)
{
    if (!p) return;
    p->pinned.release_all(p->ctx);
    delete p;
}

int iph_process(iph_processor *p, const char *task_json, const ipg_image_desc *img, const char *decoded_format,
                const char *decode_error, char **result_json)
{
    if (!p) { g_herr = "null processor"; return -1; }
    try {
        Job j;
        job_begin(p, j, task_json, img, decoded_format, decode_error);
        job_finish(p, j);
        if (result_json) *result_json = dup_cstr(j.res.marshal());
        g_herr = j.fatal;
        return j.fatal.empty() ? 0 : -1;
    } catch (const std::exception &e) {
        g_herr = std::string("internal error: ") + e.what();
        if (result_json) *result_json = nullptr;
        return -1;
    }
}

int iph_process_batch(iph_processor *p, int n, const char *const *task_json, const ipg_image_desc *imgs,
                      const char *const *decoded_formats, char **result_json, int *rc, char **errors)
{
    if (!p || n < 0 || !task_json || !imgs) { g_herr = "bad arguments"; return -1; }
    try {
        std::vector<Job> jobs((size_t)n);
        for (int i = 0; i < n; i++) job_begin(p, jobs[(size_t)i], task_json[i], &imgs[i], decoded_formats ? decoded_formats[i] : nullptr, nullptr);
        int failed = 0;
        for (int i = 0; i < n; i++) {
            Job &j = jobs[(size_t)i];
            job_finish(p, j);
            if (result_json) result_json[i] = dup_cstr(j.res.marshal());
            if (rc) rc[i] = j.fatal.empty() ? 0 : -1;
            if (errors) errors[i] = j.fatal.empty() ? nullptr : dup_cstr(j.fatal);
            if (!j.fatal.empty()) { failed++; g_herr = j.fatal; }
        }
        return failed;
    } catch (const std::exception &e) {
        g_herr = std::string("internal error: ") + e.what();
        return -1;
    }
}

void iph_free(void *p) { free(p); }
const char *iph_last_error(void) { return g_herr.c_str(); }

int iph_parse_color(const char *s, double opacity, uint8_t rgba[4]) { return parse_color(s ? s : "", opacity, rgba); }
int iph_watermark_height_px(double font_size) { return watermark_height_px(font_size); }
void iph_watermark_anchor(const char *position, int W, int H, int width_px, int height_px, int *x, int *y)
{
    watermark_anchor(position ? position : "", W, H, width_px, height_px, x, y);
}
char *iph_generate_path(const char *image_id, const char *operation, const char *format, const char *params_json)
{
    JV params;
    std::string err;
    if (params_json && *params_json) json_parse(params_json, params, err);
    return dup_cstr(generate_path(image_id ? image_id : "", operation ? operation : "", format ? format : "", params));
}
const char *iph_content_type(const char *path) { return content_type(path ? path : ""); }

} // extern "C"
