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
    uint32_t *acs;            // per block: its AC bitstring (jpeg_slot_index: JPEG_SLOT_WORDS words, 32 slots interleaved)
    uint32_t *side;           // [n_mcu][6] per block (Y0 Y1 Y2 Y3 Cb Cr): (uint16) quantised dc | AC bits << 16
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
IPG_HD int jpeg_ctz64(uint64_t a) // a != 0
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)a) - 1;
#else
    return __builtin_ctzll(a);
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

// MSB-first bit writer into a block's private slot: word k of the slot lives at acs[k * stride] (the slots of the 32
// lanes of a warp are interleaved word by word, so both k_jpeg_dct's stores and k_jpeg_emit's loads coalesce).
struct JpegSlotWriter {
    uint32_t *acs;
    int stride;
    uint32_t total;  // bits written
    uint32_t nacc;
    uint64_t acc;
    IPG_HD void put(uint32_t bits, uint32_t n) // n <= 32
    {
        acc = (acc << n) | bits;
        nacc += n;
        total += n;
        if (nacc >= 32) {
            nacc -= 32;
            *acs = (uint32_t)(acc >> nacc);
            acs += stride;
            acc &= ((uint64_t)1 << nacc) - 1;
        }
    }
    IPG_HD void finish()
    {
        if (nacc) *acs = (uint32_t)(acc << (32 - nacc));
    }
};

// Slot of block `blk` of MCU m: the blocks are grouped by (32 consecutive MCUs, blk); within a group the 32 slots
// interleave.  Returns the first word; the slot's word k is JPEG_SLOT_STRIDE words further each.
enum { JPEG_SLOT_WORDS = 52, JPEG_SLOT_STRIDE = 32 }; // 63 coefficients x (16-bit code + 10 value bits) = 1638 bits <= 52 words
IPG_HD size_t jpeg_slot_index(int m, int blk) { return ((size_t)(m >> 5) * 6 + blk) * (JPEG_SLOT_WORDS * JPEG_SLOT_STRIDE) + (m & 31); }

// Colour transform of luma quadrant `blk` (0..3) of an MCU, as writer.go's rgbaToYCbCr + scale see it: the quadrant's
// 64 luma samples to b (natural order), and the 4 x 4 chroma means of its 2 x 2 pixel groups, (sum of the four per-pixel
// 8-bit values + 2) >> 2, packed four to a word: cbw[r] / crw[r] = chroma row r of the quadrant, byte i = column i.
// px(lx, ly) returns the RGBA8 word of the MCU's pixel (lx, ly), 0 <= lx, ly < 16, with the image coordinate clamped to
// the last column / row as rgbaToYCbCr clamps it.
template <typename LoadPx>
IPG_HD void jpeg_quadrant(int blk, int32_t *b, uint32_t *cbw, uint32_t *crw, LoadPx px)
{
    const int bx = (blk & 1) * 8, by = (blk & 2) * 4;
IPG_UNROLL
    for (int r = 0; r < 4; r++) {
        uint32_t wcb = 0, wcr = 0;
IPG_UNROLL
        for (int i = 0; i < 4; i++) {
            const uint32_t p00 = px(bx + 2 * i, by + 2 * r), p01 = px(bx + 2 * i + 1, by + 2 * r);
            const uint32_t p10 = px(bx + 2 * i, by + 2 * r + 1), p11 = px(bx + 2 * i + 1, by + 2 * r + 1);
            b[16 * r + 2 * i] = jpeg_luma(p00);
            b[16 * r + 2 * i + 1] = jpeg_luma(p01);
            b[16 * r + 8 + 2 * i] = jpeg_luma(p10);
            b[16 * r + 8 + 2 * i + 1] = jpeg_luma(p11);
            const int32_t cb = (jpeg_chroma(p00, -11056, -21712, 32768) + jpeg_chroma(p01, -11056, -21712, 32768) +
                                jpeg_chroma(p10, -11056, -21712, 32768) + jpeg_chroma(p11, -11056, -21712, 32768) + 2) >> 2;
            const int32_t cr = (jpeg_chroma(p00, 32768, -27440, -5328) + jpeg_chroma(p01, 32768, -27440, -5328) +
                                jpeg_chroma(p10, 32768, -27440, -5328) + jpeg_chroma(p11, 32768, -27440, -5328) + 2) >> 2;
            wcb |= (uint32_t)cb << (8 * i);
            wcr |= (uint32_t)cr << (8 * i);
        }
        cbw[r] = wcb;
        crw[r] = wcr;
    }
}
// Where quadrant q's chroma row r goes among the 16 words of a chroma block (row-major, two words per row).
IPG_HD int jpeg_chroma_word(int q, int r) { return ((q >> 1) * 4 + r) * 2 + (q & 1); }
// The 8 x 8 chroma block from those 16 words; word(k) returns word k.
template <typename LoadWord>
IPG_HD void jpeg_chroma_block(int32_t *b, LoadWord word)
{
IPG_UNROLL
    for (int k = 0; k < 16; k++) {
        const uint32_t w = word(k);
IPG_UNROLL
        for (int i = 0; i < 4; i++) b[4 * k + i] = (int32_t)((w >> (8 * i)) & 0xff);
    }
}

// FDCT + quantisation + AC entropy coding of one block whose 64 samples are in b: the AC bitstring to the block's slot,
// (dc | AC bits << 16) to *side.  half / recip: the block's quantiser rows (luminance or chrominance); lut_ac: its AC
// Huffman table.
template <typename CoefStore>
IPG_HD void jpeg_block_code(int32_t *b, const int32_t *half, const uint32_t *recip, const uint32_t *lut_ac, uint32_t *acs, int acs_stride,
                            uint32_t *side, CoefStore coef)
{
    jpeg_fdct(b);
    // Quantise in zig-zag order.  The non-zero AC coefficients go to `coef` (the kernel: shared memory, transposed) and
    // into a 63-bit map, so that the entropy coder below is a loop over the set bits -- as many rounds as the block has
    // non-zero coefficients, not 63 -- and its code stays small (the fully unrolled form thrashed the instruction cache).
    uint64_t nz = 0;
    int dc = 0;
#define IPG_JPEG_Q(zig, nat)                                                                                                    \
    {                                                                                                                             \
        const int v = jpeg_quant(b[nat], half[zig], recip[zig]);                                                                  \
        if (zig == 0) dc = v;                                                                                                     \
        else if (v != 0) { coef.set(zig, v); nz |= (uint64_t)1 << zig; }                                                          \
    }
    IPG_JPEG_ZIGZAG(IPG_JPEG_Q)
#undef IPG_JPEG_Q
    // The AC coefficients are entropy-coded right here into this block's private slot: only the DC code depends on
    // another block (the predictor), so k_jpeg_emit later writes that code and shift-copies the slot's bits to the
    // block's place in the scan.
    JpegSlotWriter sw{acs, acs_stride, 0, 0, 0};
    int prev = 0;
    while (nz) {
        const int zig = jpeg_ctz64(nz);
        nz &= nz - 1;
        int run = zig - prev - 1;
        prev = zig;
        const int v = coef.get(zig);
        while (run > 15) { // ZRL
            sw.put(lut_ac[0xf0] & 0xffffff, lut_ac[0xf0] >> 24);
            run -= 16;
        }
        const uint32_t nb = jpeg_nbits((uint32_t)(v < 0 ? -v : v));
        const uint32_t x = lut_ac[run << 4 | nb];
        sw.put((x & 0xffffff) << nb | ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1)), (x >> 24) + nb); // <= 16 + 10 bits
    }
    if (prev != 63) sw.put(lut_ac[0x00] & 0xffffff, lut_ac[0x00] >> 24); // EOB
    sw.finish();
    *side = (uint32_t)(uint16_t)dc | sw.total << 16; // total <= 63 * 26 bits
}

// DC predictor of block `blk` of the MCU whose six side words start at side[mi * 6]: the previous block of the same
// component in scan order; `first`: the MCU is the scan's first (predictors start at 0).
IPG_HD int jpeg_prev_dc(const uint32_t *side, int mi, int blk, bool first)
{
    if (blk >= 1 && blk <= 3) return (int16_t)(side[mi * 6 + blk - 1] & 0xffff);
    if (first) return 0;
    return (int16_t)(side[(mi - 1) * 6 + (blk == 0 ? 3 : blk)] & 0xffff);
}
// Bits block `blk` of that MCU occupies in the scan: its DC code and value, and its AC bitstring.
IPG_HD uint32_t jpeg_block_bits(const uint32_t *lut_dc_lum, const uint32_t *lut_dc_chr, const uint32_t *side, int mi, int blk, bool first)
{
    const uint32_t s = side[mi * 6 + blk];
    return (s >> 16) + jpeg_dc_bits((int16_t)(s & 0xffff) - jpeg_prev_dc(side, mi, blk, first), blk < 4 ? lut_dc_lum : lut_dc_chr);
}
IPG_HD uint32_t jpeg_mcu_bits(const JpegTables &T, const uint32_t *side, int m)
{
    uint32_t bits = 0;
    for (int blk = 0; blk < 6; blk++) bits += jpeg_block_bits(T.lut[0], T.lut[2], side, m, blk, m == 0);
    return bits;
}

// MSB-first bit writer into 32-bit words shared with the neighbouring blocks (whole words are OR-ed in).
struct JpegBitWriter {
    uint32_t *words;
    uint32_t wi;     // word being filled
    uint32_t nacc;   // bits of it taken (by the predecessor and by us)
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
    IPG_HD void put(uint32_t bits, uint32_t n) // n <= 32
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

// Block `blk` of an MCU into the scan at bit offset `off`: emitHuffRLE(dc table, 0, dc - prevDC), then the AC bitstring
// k_jpeg_dct left in the block's slot.  `last` (the scan's last block): also the writer's final emit(0x7f, 7) padding.
IPG_HD void jpeg_block_emit(uint32_t *words, const uint32_t *lut_dc, uint32_t side_word, int prev_dc, uint32_t off,
                            const uint32_t *acs, int acs_stride, bool last)
{
    JpegBitWriter bw{words, off >> 5, off & 31, 0};
    const int diff = (int)(int16_t)(side_word & 0xffff) - prev_dc;
    const uint32_t nb = jpeg_nbits((uint32_t)(diff < 0 ? -diff : diff));
    const uint32_t x = lut_dc[nb];
    bw.put((x & 0xffffff) << nb | ((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1)), (x >> 24) + nb); // <= 9 + 11 bits
    const uint32_t acbits = side_word >> 16;
    for (uint32_t k = 0; k * 32 < acbits; k++) {
        const uint32_t w = acs[(size_t)k * acs_stride];
        const uint32_t n = acbits - k * 32 < 32 ? acbits - k * 32 : 32;
        bw.put(w >> (32 - n), n);
    }
    if (last && bw.nacc % 8) bw.put((1u << (8 - bw.nacc % 8)) - 1, 8 - bw.nacc % 8); // pad the last byte with ones
    bw.finish();
}

// Byte k of the unstuffed scan (MSB-first words).
IPG_HD uint32_t jpeg_scan_byte(const uint32_t *words, uint32_t k) { return (words[k >> 2] >> (24 - 8 * (k & 3))) & 0xff; }

} // namespace ipg
