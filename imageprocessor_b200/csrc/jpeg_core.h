// jpeg_core.h -- the per-thread arithmetic of the device-side baseline JPEG writer (jpeg.cu), written so that the
// same functions compile for the host: tests/support/jpeg_emu.cpp runs them serially on the CPU and compares the file
// with the oracle before the kernels ever meet a GPU.
//
// What it produces is, byte for byte, what Go 1.24's image/jpeg writer emits for an *image.RGBA -- the type
// resizeImage / cropAndResize / addTextWatermark return and the reference hands to
//   jpeg.Encode(buf, img, &jpeg.Options{Quality: 85})      operations/resize.go:78-91, watermark.go:66-79
// i.e. writer.go's rgbaToYCbCr (color.RGBToYCbCr per pixel, coordinates clamped to the image), scale (2 x 2 chroma means,
// (sum + 2) >> 2), fdct.go's jfdctint, writeBlock's round-half-away division by 8 * quant, the Annex K Huffman tables
// in scan order Y0 Y1 Y2 Y3 Cb Cr per 16 x 16 MCU, byte stuffing, and the writer's headers (no JFIF segment).
// Integer arithmetic throughout: there is nothing to certify, the bytes are equal or the test fails.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IPG_HD __host__ __device__ __forceinline__
#else
#define IPG_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define IPG_UNROLL _Pragma("unroll")
#else
#define IPG_UNROLL
#endif

namespace ipg {

// Per quality, built on the host (jpeg_host.cpp) and shipped in the parameter blob.
struct JpegTables {
    int32_t half[2][64];   // (8 * quant[q][zig]) >> 1: rounding term of writer.go's div()
    uint32_t recip[2][64]; // floor(2^32 / (8 * quant)) + 1: n / d == umulhi(n, recip) for every n this path can produce
    uint32_t lut[4][256];  // huffmanLUT: nBits << 24 | code; [0] luminance DC, [1] luminance AC, [2] chrominance DC, [3] chrominance AC
};

// One result of one ticket: the RGBA image a raster kernel left in the arena -> a complete JPEG file.
struct JpegJob {
    const uint8_t *rgba;      // RGBA8 result (premultiplied bytes; the writer ignores alpha)
    int32_t rgba_pitch;
    int32_t w, h;
    int32_t mcu_w, n_mcu;     // 16 x 16 MCUs per row, and in all
    const JpegTables *tab;
    int16_t *coef;            // [n_mcu][6][64] quantised coefficients in zig-zag order (Y0 Y1 Y2 Y3 Cb Cr)
    uint32_t *side;           // [n_mcu][6] per block: (uint16) dc | AC bits << 16
    uint32_t *mcu_off;        // [n_mcu] bit offset of each MCU in the unstuffed scan
    uint32_t *words;          // the unstuffed scan, MSB-first bits in 32-bit words (k_jpeg_zero clears what k_jpeg_emit will fill)
    uint32_t cap_bytes;       // capacity of `words` in bytes (multiple of 16) and of the scan part of `out`
    uint32_t *chunk_off;      // [cap_bytes / 512 + 1] stuffed-byte offset of each 512-byte chunk of the unstuffed scan
    uint8_t *out;             // header | stuffed scan | EOI
    uint32_t out_cap;         // bytes
    const uint8_t *hdr;       // SOI .. SOS header (jpeg_host.cpp), hdr_len bytes
    uint32_t hdr_len;
    uint32_t *result;         // [0] file length in bytes, [1] status (0 ok, 1 the scan does not fit), [2] unstuffed scan bytes, [3] scan bits
};

// ---- colour -----------------------------------------------------------------------------------------------------
// image/color/ycbcr.go RGBToYCbCr on the low three bytes of an RGBA8 pixel word (R | G << 8 | B << 16).
IPG_HD int32_t jpeg_luma(uint32_t px)
{
    const int32_t r = px & 0xff, g = (px >> 8) & 0xff, b = (px >> 16) & 0xff;
    return (19595 * r + 38470 * g + 7471 * b + (1 << 15)) >> 16;
}
// Cb: (kr, kg, kb) = (-11056, -21712, 32768); Cr: (32768, -27440, -5328)
IPG_HD int32_t jpeg_chroma(uint32_t px, int32_t kr, int32_t kg, int32_t kb)
{
    const int32_t r = px & 0xff, g = (px >> 8) & 0xff, b = (px >> 16) & 0xff;
    int32_t c = kr * r + kg * g + kb * b + (257 << 15);
    c = ((uint32_t)c & 0xff000000u) == 0 ? c >> 16 : ~(c >> 31);
    return c & 0xff;
}

// ---- forward DCT (fdct.go: jfdctint, 13-bit constants, pass1Bits = 2, output scaled up by 8) ------------------------
IPG_HD void jpeg_fdct_1d(int32_t &s0, int32_t &s1, int32_t &s2, int32_t &s3, int32_t &s4, int32_t &s5, int32_t &s6, int32_t &s7,
                         const int first_pass)
{
    constexpr int32_t fix_0_298631336 = 2446, fix_0_390180644 = 3196, fix_0_541196100 = 4433, fix_0_765366865 = 6270,
                      fix_0_899976223 = 7373, fix_1_175875602 = 9633, fix_1_501321110 = 12299, fix_1_847759065 = 15137,
                      fix_1_961570560 = 16069, fix_2_053119869 = 16819, fix_2_562915447 = 20995, fix_3_072711026 = 25172;
    constexpr int constBits = 13, pass1Bits = 2;
    int32_t tmp0 = s0 + s7, tmp1 = s1 + s6, tmp2 = s2 + s5, tmp3 = s3 + s4;
    int32_t tmp10 = tmp0 + tmp3, tmp12 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp13 = tmp1 - tmp2;
    tmp0 = s0 - s7; tmp1 = s1 - s6; tmp2 = s2 - s5; tmp3 = s3 - s4;
    const int sh = first_pass ? constBits - pass1Bits : constBits + pass1Bits;
    if (first_pass) { // rows: level shift folded in (8 * centerJSample), results scaled by 1 << pass1Bits
        s0 = (tmp10 + tmp11 - 8 * 128) * (1 << pass1Bits);
        s4 = (tmp10 - tmp11) * (1 << pass1Bits);
    } else {          // columns: the pass-1 scaling is removed with rounding
        tmp10 += 1 << (pass1Bits - 1);
        s0 = (tmp10 + tmp11) >> pass1Bits;
        s4 = (tmp10 - tmp11) >> pass1Bits;
    }
    int32_t z1 = (tmp12 + tmp13) * fix_0_541196100;
    z1 += 1 << (sh - 1);
    s2 = (z1 + tmp12 * fix_0_765366865) >> sh;
    s6 = (z1 - tmp13 * fix_1_847759065) >> sh;
    tmp10 = tmp0 + tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp0 + tmp2; tmp13 = tmp1 + tmp3;
    z1 = (tmp12 + tmp13) * fix_1_175875602;
    z1 += 1 << (sh - 1);
    tmp0 *= fix_1_501321110; tmp1 *= fix_3_072711026; tmp2 *= fix_2_053119869; tmp3 *= fix_0_298631336;
    tmp10 *= -fix_0_899976223; tmp11 *= -fix_2_562915447; tmp12 *= -fix_0_390180644; tmp13 *= -fix_1_961570560;
    tmp12 += z1; tmp13 += z1;
    s1 = (tmp0 + tmp10 + tmp12) >> sh;
    s3 = (tmp1 + tmp11 + tmp13) >> sh;
    s5 = (tmp2 + tmp11 + tmp12) >> sh;
    s7 = (tmp3 + tmp10 + tmp13) >> sh;
}
// b: 64 samples (0..255), natural order, all indices compile-time so the block lives in registers on the device
IPG_HD void jpeg_fdct(int32_t *b)
{
IPG_UNROLL
    for (int y = 0; y < 8; y++)
        jpeg_fdct_1d(b[y * 8 + 0], b[y * 8 + 1], b[y * 8 + 2], b[y * 8 + 3], b[y * 8 + 4], b[y * 8 + 5], b[y * 8 + 6], b[y * 8 + 7], 1);
IPG_UNROLL
    for (int x = 0; x < 8; x++)
        jpeg_fdct_1d(b[0 * 8 + x], b[1 * 8 + x], b[2 * 8 + x], b[3 * 8 + x], b[4 * 8 + x], b[5 * 8 + x], b[6 * 8 + x], b[7 * 8 + x], 0);
}

// (zig-zag index, natural index) pairs: writer.go's unzig
#define IPG_JPEG_ZIGZAG(X)                                                                                                         \
    X(0, 0) X(1, 1) X(2, 8) X(3, 16) X(4, 9) X(5, 2) X(6, 3) X(7, 10) X(8, 17) X(9, 24) X(10, 32) X(11, 25) X(12, 18) X(13, 11)   \
    X(14, 4) X(15, 5) X(16, 12) X(17, 19) X(18, 26) X(19, 33) X(20, 40) X(21, 48) X(22, 41) X(23, 34) X(24, 27) X(25, 20)          \
    X(26, 13) X(27, 6) X(28, 7) X(29, 14) X(30, 21) X(31, 28) X(32, 35) X(33, 42) X(34, 49) X(35, 56) X(36, 57) X(37, 50)          \
    X(38, 43) X(39, 36) X(40, 29) X(41, 22) X(42, 15) X(43, 23) X(44, 30) X(45, 37) X(46, 44) X(47, 51) X(48, 58) X(49, 59)        \
    X(50, 52) X(51, 45) X(52, 38) X(53, 31) X(54, 39) X(55, 46) X(56, 53) X(57, 60) X(58, 61) X(59, 54) X(60, 47) X(61, 55)        \
    X(62, 62) X(63, 63)

IPG_HD uint32_t jpeg_umulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
IPG_HD uint32_t jpeg_nbits(uint32_t a) // bitCount: bits needed for a (0 for 0)
{
#if defined(__CUDA_ARCH__)
    return 32u - (uint32_t)__clz((int)a);
#else
    return a ? 32u - (uint32_t)__builtin_clz(a) : 0u;
#endif
}

// writer.go div(a, 8 * quant): nearest, halves away from zero.  |a| + half < 2^18 and the divisor < 2^11, so the
// multiply-high by floor(2^32 / d) + 1 is exact (error term n * (recip * d - 2^32) < 2^32).
IPG_HD int32_t jpeg_quant(int32_t a, int32_t half, uint32_t recip)
{
    const uint32_t n = (uint32_t)(a < 0 ? -a : a) + (uint32_t)half;
    const int32_t q = (int32_t)jpeg_umulhi(n, recip);
    return a < 0 ? -q : q;
}

IPG_HD uint32_t jpeg_dc_bits(int diff, const uint32_t *lut_dc)
{
    const uint32_t nb = jpeg_nbits((uint32_t)(diff < 0 ? -diff : diff));
    return (lut_dc[nb] >> 24) + nb;
}

// Colour transform + FDCT + quantisation of block `blk` (0..3 luma quadrants, 4 Cb, 5 Cr) of one MCU: 64 zig-zag int16
// to `coef`, (dc | AC bits << 16) to *side.  px(lx, ly) returns the RGBA8 word of the MCU's pixel (lx, ly), 0 <= lx, ly < 16,
// with the image coordinate clamped to the last column / row as writer.go's rgbaToYCbCr clamps it.
template <typename LoadPx>
IPG_HD void jpeg_block(const JpegTables &T, int blk, int16_t *coef, uint32_t *side, LoadPx px)
{
    int32_t b[64];
    if (blk < 4) {
        const int bx = (blk & 1) * 8, by = (blk & 2) * 4;
IPG_UNROLL
        for (int j = 0; j < 8; j++)
IPG_UNROLL
            for (int i = 0; i < 8; i++) b[8 * j + i] = jpeg_luma(px(bx + i, by + j));
    } else {
        const int32_t kr = blk == 4 ? -11056 : 32768, kg = blk == 4 ? -21712 : -27440, kb = blk == 4 ? 32768 : -5328;
IPG_UNROLL
        for (int j = 0; j < 8; j++)
IPG_UNROLL
            for (int i = 0; i < 8; i++) // writer.go scale(): mean of the 2 x 2 group of per-pixel 8-bit chroma values
                b[8 * j + i] = (jpeg_chroma(px(2 * i, 2 * j), kr, kg, kb) + jpeg_chroma(px(2 * i + 1, 2 * j), kr, kg, kb) +
                                jpeg_chroma(px(2 * i, 2 * j + 1), kr, kg, kb) + jpeg_chroma(px(2 * i + 1, 2 * j + 1), kr, kg, kb) + 2) >> 2;
    }
    jpeg_fdct(b);
    const int q = blk < 4 ? 0 : 1;
    int16_t c[64];
#define IPG_JPEG_Q(zig, nat) c[zig] = (int16_t)jpeg_quant(b[nat], T.half[q][zig], T.recip[q][zig]);
    IPG_JPEG_ZIGZAG(IPG_JPEG_Q)
#undef IPG_JPEG_Q
    uint32_t bits = 0; // AC bits, counted here while the coefficients are in registers
    {
        const uint32_t *lut = T.lut[2 * q + 1];
        int run = 0;
IPG_UNROLL
        for (int zig = 1; zig < 64; zig++) {
            const int v = c[zig];
            if (v == 0) {
                run++;
            } else {
                bits += (uint32_t)(run >> 4) * (lut[0xf0] >> 24);
                const uint32_t nb = jpeg_nbits((uint32_t)(v < 0 ? -v : v));
                bits += (lut[(run & 15) << 4 | nb] >> 24) + nb;
                run = 0;
            }
        }
        if (run > 0) bits += lut[0x00] >> 24;
    }
    *side = (uint32_t)(uint16_t)c[0] | bits << 16;
IPG_UNROLL
    for (int k = 0; k < 64; k += 8) { // eight 16-byte stores
        uint32_t w0 = (uint16_t)c[k] | (uint32_t)(uint16_t)c[k + 1] << 16, w1 = (uint16_t)c[k + 2] | (uint32_t)(uint16_t)c[k + 3] << 16;
        uint32_t w2 = (uint16_t)c[k + 4] | (uint32_t)(uint16_t)c[k + 5] << 16, w3 = (uint16_t)c[k + 6] | (uint32_t)(uint16_t)c[k + 7] << 16;
        uint32_t *o = reinterpret_cast<uint32_t *>(coef + k);
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4 *>(o) = make_uint4(w0, w1, w2, w3);
#else
        o[0] = w0; o[1] = w1; o[2] = w2; o[3] = w3;
#endif
    }
}

// DC predictor of block `blk` of MCU m: the previous block of the same component in scan order (0 at the start).
IPG_HD int jpeg_prev_dc(const uint32_t *side, int m, int blk)
{
    if (blk >= 1 && blk <= 3) return (int16_t)(side[m * 6 + blk - 1] & 0xffff);
    if (m == 0) return 0;
    return (int16_t)(side[(m - 1) * 6 + (blk == 0 ? 3 : blk)] & 0xffff);
}
// Bits MCU m occupies in the scan.
IPG_HD uint32_t jpeg_mcu_bits(const JpegTables &T, const uint32_t *side, int m)
{
    uint32_t bits = 0;
    for (int blk = 0; blk < 6; blk++) {
        const uint32_t s = side[m * 6 + blk];
        bits += (s >> 16) + jpeg_dc_bits((int16_t)(s & 0xffff) - jpeg_prev_dc(side, m, blk), T.lut[blk < 4 ? 0 : 2]);
    }
    return bits;
}

// MSB-first bit writer into 32-bit words shared with the neighbouring MCUs (whole words are OR-ed in).
struct JpegBitWriter {
    uint32_t *words;
    uint32_t wi;     // word being filled
    uint32_t nacc;   // bits of it taken (by the predecessor MCU and by us)
    uint64_t acc;    // our bits, right-aligned
    IPG_HD void flush_word(uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicOr(words + wi, v);
#else
        words[wi] |= v;
#endif
        wi++;
    }
    IPG_HD void put(uint32_t bits, uint32_t n) // n <= 16
    {
        acc = (acc << n) | bits;
        nacc += n;
        if (nacc >= 32) {
            nacc -= 32;
            flush_word((uint32_t)(acc >> nacc));
            acc &= ((uint64_t)1 << nacc) - 1;
        }
    }
    IPG_HD void finish()
    {
        if (nacc) flush_word((uint32_t)(acc << (32 - nacc)));
    }
};

// Entropy-code MCU m at its bit offset.  `last`: also writes the writer's final emit(0x7f, 7) padding.
IPG_HD void jpeg_mcu_emit(const JpegJob &J, const JpegTables &T, int m, bool last)
{
    const uint32_t off = J.mcu_off[m];
    JpegBitWriter bw{J.words, off >> 5, off & 31, 0};
    for (int blk = 0; blk < 6; blk++) {
        const int16_t *c = J.coef + ((size_t)m * 6 + blk) * 64;
        const uint32_t *ldc = T.lut[blk < 4 ? 0 : 2], *lac = T.lut[blk < 4 ? 1 : 3];
        { // emitHuffRLE(dc table, 0, dc - prevDC)
            const int diff = (int)c[0] - jpeg_prev_dc(J.side, m, blk);
            const uint32_t nb = jpeg_nbits((uint32_t)(diff < 0 ? -diff : diff));
            const uint32_t x = ldc[nb];
            bw.put(x & 0xffffff, x >> 24);
            if (nb) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1), nb);
        }
        int run = 0;
        for (int zig = 1; zig < 64; zig++) {
            const int v = c[zig];
            if (v == 0) { run++; continue; }
            while (run > 15) {
                const uint32_t z = lac[0xf0];
                bw.put(z & 0xffffff, z >> 24);
                run -= 16;
            }
            const uint32_t nb = jpeg_nbits((uint32_t)(v < 0 ? -v : v));
            const uint32_t x = lac[run << 4 | nb];
            bw.put(x & 0xffffff, x >> 24);
            bw.put((uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1), nb);
            run = 0;
        }
        if (run > 0) {
            const uint32_t z = lac[0x00];
            bw.put(z & 0xffffff, z >> 24);
        }
    }
    if (last && bw.nacc % 8) bw.put((1u << (8 - bw.nacc % 8)) - 1, 8 - bw.nacc % 8); // pad the last byte with ones
    bw.finish();
}

// Byte k of the unstuffed scan (MSB-first words).
IPG_HD uint32_t jpeg_scan_byte(const uint32_t *words, uint32_t k) { return (words[k >> 2] >> (24 - 8 * (k & 3))) & 0xff; }

} // namespace ipg
