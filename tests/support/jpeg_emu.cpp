// jpeg_emu.cpp -- test-only: the device JPEG writer's per-thread functions (imageprocessor_b200/csrc/jpeg_core.h, the same
// source the kernels compile) run serially on the CPU, in the kernels' own decomposition (block per thread, MCU per
// thread, 16 scan bytes per lane, 512 per warp), so that its logic is compared with the oracle without a GPU
// (tests/test_jpeg_emu.py).  The scans and warp collectives of jpeg.cu are plain loops here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../imageprocessor_b200/csrc/jpeg.h"

using namespace ipg;

extern "C" long jpeg_emu_encode(const uint8_t *rgba, int pitch, int w, int h, int quality, uint8_t *out, size_t out_cap, size_t scan_cap)
{
    JpegTables T;
    jpeg_build_tables(quality, &T);
    uint8_t hdr[JPEG_HDR_MAX];
    const size_t hdr_len = jpeg_build_header(quality, w, h, hdr);
    JpegJob J{};
    J.rgba = rgba; J.rgba_pitch = pitch; J.w = w; J.h = h;
    J.mcu_w = (w + 15) / 16;
    J.n_mcu = J.mcu_w * ((h + 15) / 16);
    J.tab = &T;
    std::vector<uint32_t> acs((size_t)(J.n_mcu + 31) / 32 * 6 * JPEG_SLOT_WORDS * JPEG_SLOT_STRIDE, 0xdeadbeefu);
    std::vector<uint32_t> side((size_t)J.n_mcu * 6), mcu_off(J.n_mcu);
    scan_cap = (scan_cap + JPEG_CHUNK - 1) / JPEG_CHUNK * JPEG_CHUNK;
    std::vector<uint32_t> words(scan_cap / 4, 0u), chunk_off(scan_cap / JPEG_CHUNK + 1, 0u);
    uint32_t result[4] = {0, 0, 0, 0};
    J.acs = acs.data(); J.side = side.data(); J.mcu_off = mcu_off.data(); J.words = words.data();
    J.cap_bytes = (uint32_t)scan_cap; J.chunk_off = chunk_off.data(); J.out = out; J.out_cap = (uint32_t)out_cap;
    J.hdr = hdr; J.hdr_len = (uint32_t)hdr_len; J.result = result;
    // k_jpeg_dct
    for (int m = 0; m < J.n_mcu; m++) {
        const int x0 = (m % J.mcu_w) * 16, y0 = (m / J.mcu_w) * 16;
        uint32_t cmean[2][16];
        auto px = [&](int lx, int ly) {
            const int sx = x0 + lx < w - 1 ? x0 + lx : w - 1, sy = y0 + ly < h - 1 ? y0 + ly : h - 1;
            uint32_t v;
            memcpy(&v, rgba + (size_t)sy * pitch + (size_t)sx * 4, 4);
            return v;
        };
        for (int blk = 0; blk < 6; blk++) {
            int32_t b[64];
            const int q = blk < 4 ? 0 : 1;
            if (blk < 4) {
                uint32_t cbw[4], crw[4];
                jpeg_quadrant(blk, b, cbw, crw, px);
                for (int r = 0; r < 4; r++) { cmean[0][jpeg_chroma_word(blk, r)] = cbw[r]; cmean[1][jpeg_chroma_word(blk, r)] = crw[r]; }
            } else {
                const uint32_t *cw = cmean[blk - 4];
                jpeg_chroma_block(b, [cw](int k) { return cw[k]; });
            }
            struct CoefLocal {
                int16_t *c;
                void set(int zig, int v) { c[zig] = (int16_t)v; }
                int get(int zig) const { return c[zig]; }
            };
            int16_t cbuf[64];
            jpeg_block_code(b, T.half[q], T.recip[q], T.lut[2 * q + 1], J.acs + jpeg_slot_index(m, blk), JPEG_SLOT_STRIDE, J.side + (size_t)m * 6 + blk,
                            CoefLocal{cbuf});
        }
    }
    // k_jpeg_offsets
    uint64_t bits = 0;
    for (int m = 0; m < J.n_mcu; m++) {
        J.mcu_off[m] = (uint32_t)bits;
        bits += jpeg_mcu_bits(T, J.side, m);
    }
    result[3] = (uint32_t)bits;
    result[2] = (uint32_t)((bits + 7) >> 3);
    if (bits >= 0xffffffffull || ((bits + 7) >> 3) > scan_cap) return -1;
    // k_jpeg_emit
    for (int m = 0; m < J.n_mcu; m++) {
        uint32_t off = J.mcu_off[m];
        for (int blk = 0; blk < 6; blk++) {
            jpeg_block_emit(J.words, T.lut[blk < 4 ? 0 : 2], J.side[m * 6 + blk], jpeg_prev_dc(J.side, m, blk, m == 0), off,
                            J.acs + jpeg_slot_index(m, blk), JPEG_SLOT_STRIDE, m == J.n_mcu - 1 && blk == 5);
            off += jpeg_block_bits(T.lut[0], T.lut[2], J.side, m, blk, m == 0);
        }
    }
    // k_jpeg_ffcount + k_jpeg_chunks
    const uint32_t U = result[2], n_chunks = (U + JPEG_CHUNK - 1) / JPEG_CHUNK;
    uint32_t ff = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        J.chunk_off[c] = ff;
        for (uint32_t k = c * JPEG_CHUNK; k < (c + 1) * JPEG_CHUNK; k++) ff += jpeg_scan_byte(J.words, k) == 0xff;
    }
    const uint64_t len = hdr_len + U + ff + 2;
    if (len > out_cap) return -1;
    // k_jpeg_write
    memcpy(out, hdr, hdr_len);
    for (uint32_t c = 0; c < n_chunks; c++) {
        uint32_t before = 0;
        for (int lane = 0; lane < 32; lane++) {
            const uint32_t k0 = c * JPEG_CHUNK + lane * 16;
            uint8_t *o = out + hdr_len + k0 + J.chunk_off[c] + before;
            for (int k = 0; k < 16; k++)
                if (k0 + k < U) {
                    const uint32_t b = jpeg_scan_byte(J.words, k0 + k);
                    *o++ = (uint8_t)b;
                    if (b == 0xff) { *o++ = 0; before++; }
                }
        }
    }
    out[len - 2] = 0xff;
    out[len - 1] = 0xd9;
    return (long)len;
}
