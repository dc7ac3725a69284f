/*
 * ip_jpeg_oracle.c -- CPU restatement of Go's image/jpeg ENCODER for the outputs of the worker hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ip_oracle.h): the checker for the device-side baseline JPEG writer
 * (imageprocessor_b200/csrc/jpeg.cu, ipg_op.dst_layout = IPG_LAYOUT_JPEG).
 *
 * What it restates, by name (Go 1.24.7 stdlib, go.mod:3; not under /root/reference):
 *   image/jpeg/writer.go   Encode, writeDQT, writeSOF0, writeDHT, writeSOS, writeBlock, emit, emitHuff,
 *                          emitHuffRLE, div, rgbaToYCbCr, yCbCrToYCbCr, grayToY, scale, unscaledQuant,
 *                          theHuffmanSpec, huffmanLUT.init, bitCount
 *   image/jpeg/fdct.go     fdct (jfdctint: 13-bit constants, pass1Bits = 2, results scaled up by 8)
 *   image/color/ycbcr.go   RGBToYCbCr
 * anchored on the reference's call sites jpeg.Encode(buf, img, &jpeg.Options{Quality: 85}):
 *   operations/resize.go:78-91, operations/thumbnail.go (same switch), operations/watermark.go:66-79;
 *   quality constant domain/task.go:57.
 *
 * PARITY with the Go binary is UNPINNED like the rest of the oracle (no Go toolchain here).  What pins this file
 * instead (tests/test_jpeg_oracle.py): libjpeg-turbo, through PIL, is an independent implementation of the same
 * baseline process -- Annex K tables scaled by the same quality rule, the same jfdctint forward DCT, round-half-away
 * quantisation, the Annex K Huffman tables -- so for an image whose chroma is constant over each 2 x 2 group the
 * entropy-coded segment, the DQT payload and the DHT payload of this writer must equal libjpeg-turbo's byte for byte.
 * They do.  The colour conversion (RGBToYCbCr) and the 2 x 2 chroma mean are pinned by ip_oracle.c's
 * ipo_rgba_to_ycbcr420 (tests/test_golden.py) and cross-checked here on sizes that need no MCU padding.
 */
#include "ip_oracle.h"
#include <string.h>

/* ---- tables (natural order as ITU T.81 Annex K prints them; writer.go stores the same numbers in zig-zag order) ---- */
static const uint8_t k_quant_natural[2][64] = {
    {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
     18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99},
};
/* unzig[zig] = natural index of the zig-th coefficient */
static const uint8_t k_unzig[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
};

typedef struct { uint8_t count[16]; const uint8_t *value; int n; } huff_spec;
static const uint8_t k_dc_values[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t k_ac_lum_values[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa,
};
static const uint8_t k_ac_chr_values[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa,
};
/* theHuffmanSpec: luminance DC, luminance AC, chrominance DC, chrominance AC (huffIndex order) */
static const huff_spec k_huff[4] = {
    {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, k_dc_values, 12},
    {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125}, k_ac_lum_values, 162},
    {{0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, k_dc_values, 12},
    {{0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}, k_ac_chr_values, 162},
};

typedef struct {
    uint8_t *out;
    size_t cap, n;
    int overflow;
    uint32_t bits, nbits;
    uint8_t quant[2][64];     /* zig-zag order, as writer.go keeps them */
    uint32_t lut[4][256];     /* huffmanLUT: nBits << 24 | code */
} enc;

static void put(enc *e, uint8_t b)
{
    if (e->n < e->cap) e->out[e->n] = b; else e->overflow = 1;
    e->n++;
}
static void put_n(enc *e, const uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) put(e, p[i]); }

/* writer.go emit: bits are packed MSB first; a 0xff data byte is followed by a stuffed 0x00 */
static void emit(enc *e, uint32_t bits, uint32_t nbits)
{
    nbits += e->nbits;
    bits <<= 32 - nbits;
    bits |= e->bits;
    while (nbits >= 8) {
        uint8_t b = (uint8_t)(bits >> 24);
        put(e, b);
        if (b == 0xff) put(e, 0x00);
        bits <<= 8;
        nbits -= 8;
    }
    e->bits = bits;
    e->nbits = nbits;
}
static void emit_huff(enc *e, int h, int32_t value)
{
    uint32_t x = e->lut[h][value];
    emit(e, x & ((1u << 24) - 1), x >> 24);
}
static uint32_t bit_count(int32_t a) /* bitCount[a] for a < 256, 8 + bitCount[a >> 8] above */
{
    uint32_t n = 0;
    while (a) { n++; a >>= 1; }
    return n;
}
static void emit_huff_rle(enc *e, int h, int32_t run, int32_t value)
{
    int32_t a = value, b = value;
    if (a < 0) { a = -value; b = value - 1; }
    uint32_t nbits = bit_count(a);
    emit_huff(e, h, run << 4 | (int32_t)nbits);
    if (nbits > 0) emit(e, (uint32_t)b & ((1u << nbits) - 1), nbits);
}
/* writer.go div: a / b rounded to nearest, halves away from zero */
static int32_t div_round(int32_t a, int32_t b)
{
    if (a >= 0) return (a + (b >> 1)) / b;
    return -((-a + (b >> 1)) / b);
}

/* fdct.go */
enum {
    fix_0_298631336 = 2446, fix_0_390180644 = 3196, fix_0_541196100 = 4433, fix_0_765366865 = 6270,
    fix_0_899976223 = 7373, fix_1_175875602 = 9633, fix_1_501321110 = 12299, fix_1_847759065 = 15137,
    fix_1_961570560 = 16069, fix_2_053119869 = 16819, fix_2_562915447 = 20995, fix_3_072711026 = 25172,
    constBits = 13, pass1Bits = 2, centerJSample = 128,
};
static void fdct(int32_t *b)
{
    for (int y = 0; y < 8; y++) { /* pass 1: rows */
        int32_t *s = b + y * 8;
        int32_t x0 = s[0], x1 = s[1], x2 = s[2], x3 = s[3], x4 = s[4], x5 = s[5], x6 = s[6], x7 = s[7];
        int32_t tmp0 = x0 + x7, tmp1 = x1 + x6, tmp2 = x2 + x5, tmp3 = x3 + x4;
        int32_t tmp10 = tmp0 + tmp3, tmp12 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp13 = tmp1 - tmp2;
        tmp0 = x0 - x7; tmp1 = x1 - x6; tmp2 = x2 - x5; tmp3 = x3 - x4;
        s[0] = (tmp10 + tmp11 - 8 * centerJSample) * (1 << pass1Bits);
        s[4] = (tmp10 - tmp11) * (1 << pass1Bits);
        int32_t z1 = (tmp12 + tmp13) * fix_0_541196100;
        z1 += 1 << (constBits - pass1Bits - 1);
        s[2] = (z1 + tmp12 * fix_0_765366865) >> (constBits - pass1Bits);
        s[6] = (z1 - tmp13 * fix_1_847759065) >> (constBits - pass1Bits);
        tmp10 = tmp0 + tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp0 + tmp2; tmp13 = tmp1 + tmp3;
        z1 = (tmp12 + tmp13) * fix_1_175875602;
        z1 += 1 << (constBits - pass1Bits - 1);
        tmp0 *= fix_1_501321110; tmp1 *= fix_3_072711026; tmp2 *= fix_2_053119869; tmp3 *= fix_0_298631336;
        tmp10 *= -fix_0_899976223; tmp11 *= -fix_2_562915447; tmp12 *= -fix_0_390180644; tmp13 *= -fix_1_961570560;
        tmp12 += z1; tmp13 += z1;
        s[1] = (tmp0 + tmp10 + tmp12) >> (constBits - pass1Bits);
        s[3] = (tmp1 + tmp11 + tmp13) >> (constBits - pass1Bits);
        s[5] = (tmp2 + tmp11 + tmp12) >> (constBits - pass1Bits);
        s[7] = (tmp3 + tmp10 + tmp13) >> (constBits - pass1Bits);
    }
    for (int x = 0; x < 8; x++) { /* pass 2: columns; removes pass1Bits, leaves the overall factor 8 */
        int32_t tmp0 = b[0 * 8 + x] + b[7 * 8 + x], tmp1 = b[1 * 8 + x] + b[6 * 8 + x];
        int32_t tmp2 = b[2 * 8 + x] + b[5 * 8 + x], tmp3 = b[3 * 8 + x] + b[4 * 8 + x];
        int32_t tmp10 = tmp0 + tmp3 + (1 << (pass1Bits - 1)), tmp12 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp13 = tmp1 - tmp2;
        tmp0 = b[0 * 8 + x] - b[7 * 8 + x]; tmp1 = b[1 * 8 + x] - b[6 * 8 + x];
        tmp2 = b[2 * 8 + x] - b[5 * 8 + x]; tmp3 = b[3 * 8 + x] - b[4 * 8 + x];
        b[0 * 8 + x] = (tmp10 + tmp11) >> pass1Bits;
        b[4 * 8 + x] = (tmp10 - tmp11) >> pass1Bits;
        int32_t z1 = (tmp12 + tmp13) * fix_0_541196100;
        z1 += 1 << (constBits + pass1Bits - 1);
        b[2 * 8 + x] = (z1 + tmp12 * fix_0_765366865) >> (constBits + pass1Bits);
        b[6 * 8 + x] = (z1 - tmp13 * fix_1_847759065) >> (constBits + pass1Bits);
        tmp10 = tmp0 + tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp0 + tmp2; tmp13 = tmp1 + tmp3;
        z1 = (tmp12 + tmp13) * fix_1_175875602;
        z1 += 1 << (constBits + pass1Bits - 1);
        tmp0 *= fix_1_501321110; tmp1 *= fix_3_072711026; tmp2 *= fix_2_053119869; tmp3 *= fix_0_298631336;
        tmp10 *= -fix_0_899976223; tmp11 *= -fix_2_562915447; tmp12 *= -fix_0_390180644; tmp13 *= -fix_1_961570560;
        tmp12 += z1; tmp13 += z1;
        b[1 * 8 + x] = (tmp0 + tmp10 + tmp12) >> (constBits + pass1Bits);
        b[3 * 8 + x] = (tmp1 + tmp11 + tmp13) >> (constBits + pass1Bits);
        b[5 * 8 + x] = (tmp2 + tmp11 + tmp12) >> (constBits + pass1Bits);
        b[7 * 8 + x] = (tmp3 + tmp10 + tmp13) >> (constBits + pass1Bits);
    }
}

/* writeBlock: fdct, quantise, Huffman-code one 8 x 8 block of component class q (0 luminance, 1 chrominance) */
static int32_t write_block(enc *e, int32_t *b, int q, int32_t prev_dc)
{
    fdct(b);
    int32_t dc = div_round(b[0], 8 * (int32_t)e->quant[q][0]);
    emit_huff_rle(e, 2 * q + 0, 0, dc - prev_dc);
    int h = 2 * q + 1;
    int32_t run = 0;
    for (int zig = 1; zig < 64; zig++) {
        int32_t ac = div_round(b[k_unzig[zig]], 8 * (int32_t)e->quant[q][zig]);
        if (ac == 0) {
            run++;
        } else {
            while (run > 15) { emit_huff(e, h, 0xf0); run -= 16; }
            emit_huff_rle(e, h, run, ac);
            run = 0;
        }
    }
    if (run > 0) emit_huff(e, h, 0x00);
    return dc;
}

static void go_rgb_to_ycbcr(uint8_t r, uint8_t g, uint8_t b, int32_t *yy, int32_t *cb, int32_t *cr)
{
    int32_t r1 = r, g1 = g, b1 = b;
    int32_t y = (19595 * r1 + 38470 * g1 + 7471 * b1 + (1 << 15)) >> 16;
    int32_t c = -11056 * r1 - 21712 * g1 + 32768 * b1 + (257 << 15);
    if (((uint32_t)c & 0xff000000u) == 0) c >>= 16; else c = ~(c >> 31);
    int32_t d = 32768 * r1 - 27440 * g1 - 5328 * b1 + (257 << 15);
    if (((uint32_t)d & 0xff000000u) == 0) d >>= 16; else d = ~(d >> 31);
    *yy = (uint8_t)y; *cb = (uint8_t)c; *cr = (uint8_t)d;
}

/* scale: the 16 x 16 region held in 4 blocks -> one 8 x 8 block, 2 x 2 means (sum + 2) >> 2 */
static void scale_blocks(int32_t *dst, int32_t src[4][64])
{
    for (int i = 0; i < 4; i++) {
        int dst_off = (i & 2) << 4 | (i & 1) << 2;
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) {
                int j = 16 * y + 2 * x;
                int32_t sum = src[i][j] + src[i][j + 1] + src[i][j + 8] + src[i][j + 9];
                dst[8 * y + x + dst_off] = (sum + 2) >> 2;
            }
    }
}

static void enc_init(enc *e, uint8_t *out, size_t cap, int quality)
{
    memset(e, 0, sizeof *e);
    e->out = out;
    e->cap = cap;
    if (quality < 1) quality = 1; else if (quality > 100) quality = 100;
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 64; j++) {
            int x = k_quant_natural[i][k_unzig[j]];
            x = (x * scale + 50) / 100;
            if (x < 1) x = 1; else if (x > 255) x = 255;
            e->quant[i][j] = (uint8_t)x;
        }
    for (int t = 0; t < 4; t++) { /* huffmanLUT.init */
        uint32_t code = 0;
        int k = 0;
        for (int i = 0; i < 16; i++) {
            uint32_t nbits = (uint32_t)(i + 1) << 24;
            for (int j = 0; j < k_huff[t].count[i]; j++) {
                e->lut[t][k_huff[t].value[k]] = nbits | code;
                code++;
                k++;
            }
            code <<= 1;
        }
    }
}

static void marker_header(enc *e, uint8_t marker, int len)
{
    put(e, 0xff); put(e, marker); put(e, (uint8_t)(len >> 8)); put(e, (uint8_t)(len & 0xff));
}

static void write_headers(enc *e, int w, int h, int ncomp)
{
    put(e, 0xff); put(e, 0xd8); /* SOI */
    marker_header(e, 0xdb, 2 + 2 * (1 + 64)); /* writeDQT: both tables in one segment */
    for (int i = 0; i < 2; i++) { put(e, (uint8_t)i); put_n(e, e->quant[i], 64); }
    marker_header(e, 0xc0, 8 + 3 * ncomp); /* writeSOF0 */
    put(e, 8);
    put(e, (uint8_t)(h >> 8)); put(e, (uint8_t)(h & 0xff));
    put(e, (uint8_t)(w >> 8)); put(e, (uint8_t)(w & 0xff));
    put(e, (uint8_t)ncomp);
    if (ncomp == 1) {
        put(e, 1); put(e, 0x11); put(e, 0x00);
    } else {
        static const uint8_t samp[3] = {0x22, 0x11, 0x11}, tq[3] = {0, 1, 1};
        for (int i = 0; i < 3; i++) { put(e, (uint8_t)(i + 1)); put(e, samp[i]); put(e, tq[i]); }
    }
    int nspec = ncomp == 1 ? 2 : 4, len = 2; /* writeDHT: all tables in one segment */
    for (int i = 0; i < nspec; i++) len += 1 + 16 + k_huff[i].n;
    marker_header(e, 0xc4, len);
    static const uint8_t tc_th[4] = {0x00, 0x10, 0x01, 0x11};
    for (int i = 0; i < nspec; i++) {
        put(e, tc_th[i]);
        put_n(e, k_huff[i].count, 16);
        put_n(e, k_huff[i].value, (size_t)k_huff[i].n);
    }
    if (ncomp == 1) {
        static const uint8_t sos_y[10] = {0xff, 0xda, 0x00, 0x08, 0x01, 0x01, 0x00, 0x00, 0x3f, 0x00};
        put_n(e, sos_y, sizeof sos_y);
    } else {
        static const uint8_t sos_ycbcr[14] = {0xff, 0xda, 0x00, 0x0c, 0x03, 0x01, 0x00, 0x02, 0x11, 0x03, 0x11, 0x00, 0x3f, 0x00};
        put_n(e, sos_ycbcr, sizeof sos_ycbcr);
    }
}

/* A per-pixel (Y, Cb, Cr) source for the colour path of writeSOS */
typedef void (*px_fn)(const void *ctx, int x, int y, int32_t *yy, int32_t *cb, int32_t *cr);

typedef struct { const uint8_t *pix; int stride; } rgba_src;
static void px_rgba(const void *ctx, int x, int y, int32_t *yy, int32_t *cb, int32_t *cr)
{
    const rgba_src *s = (const rgba_src *)ctx;
    const uint8_t *p = s->pix + (size_t)y * (size_t)s->stride + (size_t)x * 4;
    go_rgb_to_ycbcr(p[0], p[1], p[2], yy, cb, cr);
}
static void px_ycbcr(const void *ctx, int x, int y, int32_t *yy, int32_t *cb, int32_t *cr)
{
    const ipo_image *m = (const ipo_image *)ctx;
    int cx = x, cy = y; /* COffset per subsample ratio */
    if (m->layout == IPO_YCBCR422 || m->layout == IPO_YCBCR420) cx = x / 2;
    if (m->layout == IPO_YCBCR420 || m->layout == IPO_YCBCR440) cy = y / 2;
    *yy = m->plane[0][(size_t)y * (size_t)m->stride[0] + (size_t)x];
    *cb = m->plane[1][(size_t)cy * (size_t)m->stride[1] + (size_t)cx];
    *cr = m->plane[2][(size_t)cy * (size_t)m->stride[2] + (size_t)cx];
}

static size_t encode_color(px_fn px, const void *ctx, int w, int h, int quality, uint8_t *out, size_t cap)
{
    if (w <= 0 || h <= 0 || w >= 1 << 16 || h >= 1 << 16) return 0; /* "jpeg: image is too large to encode" */
    enc e;
    enc_init(&e, out, cap, quality);
    write_headers(&e, w, h, 3);
    int32_t b[64], cb[4][64], cr[4][64];
    int32_t prev_y = 0, prev_cb = 0, prev_cr = 0;
    const int xmax = w - 1, ymax = h - 1;
    for (int y = 0; y < h; y += 16)
        for (int x = 0; x < w; x += 16) {
            for (int i = 0; i < 4; i++) {
                const int px0 = x + (i & 1) * 8, py0 = y + (i & 2) * 4;
                for (int j = 0; j < 8; j++) {
                    int sy = py0 + j;
                    if (sy > ymax) sy = ymax;
                    for (int k = 0; k < 8; k++) {
                        int sx = px0 + k;
                        if (sx > xmax) sx = xmax;
                        px(ctx, sx, sy, &b[8 * j + k], &cb[i][8 * j + k], &cr[i][8 * j + k]);
                    }
                }
                prev_y = write_block(&e, b, 0, prev_y);
            }
            scale_blocks(b, cb);
            prev_cb = write_block(&e, b, 1, prev_cb);
            scale_blocks(b, cr);
            prev_cr = write_block(&e, b, 1, prev_cr);
        }
    emit(&e, 0x7f, 7); /* pad the last byte with ones */
    put(&e, 0xff); put(&e, 0xd9);
    return e.overflow ? (size_t)-1 : e.n;
}

size_t ipo_jpeg_encode_rgba(const uint8_t *rgba, int stride, int w, int h, int quality, uint8_t *out, size_t cap)
{
    if (!rgba || !out) return 0;
    rgba_src s = {rgba, stride};
    return encode_color(px_rgba, &s, w, h, quality, out, cap);
}

size_t ipo_jpeg_encode_ycbcr(const ipo_image *img, int quality, uint8_t *out, size_t cap)
{
    if (!img || !out || img->layout < IPO_YCBCR444 || img->layout > IPO_YCBCR440) return 0;
    return encode_color(px_ycbcr, img, img->width, img->height, quality, out, cap);
}

size_t ipo_jpeg_encode_gray(const uint8_t *pix, int stride, int w, int h, int quality, uint8_t *out, size_t cap)
{
    if (!pix || !out || w <= 0 || h <= 0 || w >= 1 << 16 || h >= 1 << 16) return 0;
    enc e;
    enc_init(&e, out, cap, quality);
    write_headers(&e, w, h, 1);
    int32_t b[64], prev = 0;
    for (int y = 0; y < h; y += 8)
        for (int x = 0; x < w; x += 8) {
            for (int j = 0; j < 8; j++)
                for (int k = 0; k < 8; k++) { /* grayToY */
                    int sx = x + k, sy = y + j;
                    if (sx > w - 1) sx = w - 1;
                    if (sy > h - 1) sy = h - 1;
                    b[8 * j + k] = pix[(size_t)sy * (size_t)stride + (size_t)sx];
                }
            prev = write_block(&e, b, 0, prev);
        }
    emit(&e, 0x7f, 7);
    put(&e, 0xff); put(&e, 0xd9);
    return e.overflow ? (size_t)-1 : e.n;
}
