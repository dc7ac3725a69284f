"""Integer identities the sm_100a kernels rely on where they do NOT spell the reference's expression literally
(imageprocessor_b200/csrc/kernels.cu).  Each is checked exhaustively over the inputs that can occur, against the
reference's own form (golang.org/x/image draw/impl.go scaleX_NRGBA / scaleX_YCbCr*, image/color ycbcr.go), in
numpy: no GPU, no oracle."""
import numpy as np


def test_div255_by_multiply_high_is_exact():
    # k_stream_planar<NRGBA>: c16 = C * a16 / 0xff with a16 = A * 0x101 (scaleX_NRGBA), the division as
    # umulhi(x, 0x80808081) >> 7
    c = np.arange(256, dtype=np.uint64)[:, None]
    a = np.arange(256, dtype=np.uint64)[None, :]
    x = c * (a * 0x101)
    assert x.max() < 2 ** 32
    got = ((x * 0x80808081) >> 32) >> 7
    assert np.array_equal(got, x // 0xFF)
    # the crop stage keeps alpha: (a16 >> 8) * 0x101 == a16
    a16 = np.arange(256, dtype=np.uint32) * 0x101
    assert np.array_equal((a16 >> 8) * 0x101, a16)


def _clamp(v, lo, hi):
    return np.minimum(np.maximum(v, lo), hi)


def test_ycbcr_terms_and_folded_crop_stage():
    # planar_vloop: the chroma terms are computed once per chroma sample and added to yy1 per pixel; for the crop
    # stage (cropAndResize's 1:1 first pass keeps uint8(c16 >> 8)) the two shifts fold into one: clamp(v >> 16, 0, 255)
    y = np.arange(256, dtype=np.int64)
    cb = np.arange(256, dtype=np.int64) - 128
    cr = np.arange(256, dtype=np.int64) - 128
    yy1 = y * 0x10101
    # r depends on (y, cr), b on (y, cb): 65,536 cases each; g on all three: 16.7 M
    vr = yy1[:, None] + 91881 * cr[None, :]
    vb = yy1[:, None] + 116130 * cb[None, :]
    vg = yy1[:, None, None] + (-22554 * cb[None, :, None] - 46802 * cr[None, None, :])
    for v in (vr, vb, vg):
        assert np.abs(v).max() < 2 ** 31                       # int32 arithmetic in the kernel does not wrap
        ref16 = _clamp(v >> 8, 0, 0xFFFF)                       # color.YCbCr.RGBA(): (x >> 8) clamped to 16 bits
        assert np.array_equal(_clamp(v >> 16, 0, 0xFF), ref16 >> 8)   # ycc_chan<16> == uint8(ref16 >> 8)
    # Gray: (Y * 0x101 >> 8) * 0x101 == Y * 0x101, so the crop stage changes nothing
    g16 = y * 0x101
    assert np.array_equal((g16 >> 8) * 0x101, g16)


def test_magic_number_byte_to_float():
    # unpack_rgb / u16x2_f32: bits 0x4B000000 | v is the float 2^23 + v for v < 2^23, so one FADD of -2^23 yields v
    v = np.concatenate([np.arange(0, 65536, dtype=np.uint32), np.array([2 ** 23 - 1], np.uint32)])
    f = (np.uint32(0x4B000000) | v).view(np.float32) - np.float32(8388608.0)
    assert np.array_equal(f.astype(np.uint32), v)
    # IDP.4A form (IPG_DP4A_MASK): 0x4B000000 + q . (1 << 8k) is the same word as the PRMT form for any byte
    q = np.arange(256, dtype=np.uint32)
    assert np.array_equal(np.uint32(0x4B000000) + q, np.uint32(0x4B000000) | q)


def test_quantiser_ambiguity_window():
    # quant16: T = floor(v * 256 + 128) in 16.8 fixed point; byte = T >> 16; "ambiguous" iff T lies within D of a
    # multiple of 65536, written as one unsigned compare ((T & 0xffff) - D) mod 2^32 >= 65536 - 2 D
    for D in (6, 22, 64, 200):
        low = np.arange(65536, dtype=np.uint32)
        got = ((low - np.uint32(D)) & np.uint32(0xFFFFFFFF)) >= np.uint32(65536 - 2 * D)
        want = (low < D) | (low >= 65536 - D)
        assert np.array_equal(got, want)


def _dp2a(a, b, c, hi):
    """PTX dp2a.{lo,hi}.u32.u32 d, a, b, c: a = two 16-bit lanes, b = four bytes, of which .lo takes bytes 0, 1 and .hi
    bytes 2, 3: d = c + a.lo16 * b[sel] + a.hi16 * b[sel + 1]  (mod 2^32)."""
    a, b, c = a.astype(np.uint64), b.astype(np.uint64), c.astype(np.uint64)
    sel = 16 if hi else 0
    b0, b1 = (b >> sel) & 0xFF, (b >> (sel + 8)) & 0xFF
    return ((c + (a & 0xFFFF) * b0 + (a >> 16) * b1) & 0xFFFFFFFF).astype(np.uint32)


def test_integer_moment_word_by_dp2a():
    # k_stream's integer-moment loop: one IDP.2A per channel and source row with m = 1 + (r << 12) adds x to bits 0..11
    # (M0 = sum x) and r * x from bit 12 on (M1 = sum r * x) of the channel's word; [m, 0] picks byte 0 (.lo) or 2 (.hi)
    # of the pixel word, [0, m] = m << 16 picks byte 1 (.lo) -- the alpha byte is never touched
    rng = np.random.default_rng(0)
    q = rng.integers(0, 2 ** 32, 4096, dtype=np.uint64).astype(np.uint32)
    for r in range(16):
        m = np.full(q.shape, 1 + (r << 12), np.uint32)
        zero = np.zeros_like(q)
        byte = lambda k: (q >> np.uint32(8 * k)) & np.uint32(0xFF)   # noqa: E731
        assert np.array_equal(_dp2a(m, q, zero, False), byte(0) * m)
        assert np.array_equal(_dp2a(m << np.uint32(16), q, zero, False), byte(1) * m)
        assert np.array_equal(_dp2a(m, q, zero, True), byte(2) * m)
        assert m.max() < 2 ** 16                                      # the multiplier fits its 16-bit lane
    # a piece of 16 rows of the largest byte: M0 = 4080 stays below bit 12, the word below 2^32, and both moments come
    # back out; the fp32 conversions of the flush are exact (both < 2^24)
    acc = np.zeros(1, np.uint32)
    px = np.full(1, 0xFFFFFFFF, np.uint32)
    for r in range(16):
        acc = _dp2a(np.full(1, 1 + (r << 12), np.uint32), px, acc, False)
    assert int(acc[0] & 0xFFF) == 16 * 255 and int(acc[0] >> 12) == 255 * sum(range(16))
    assert 16 * 255 < 2 ** 12 and (255 * 120 << 12) + 4080 < 2 ** 32
    assert np.float32(acc[0] >> 12) == 255 * 120 and np.float32(acc[0] & 0xFFF) == 4080
    # random pieces: the packed word equals the two sums computed apart
    rows = rng.integers(0, 256, (200, 16), dtype=np.uint32)
    n = rng.integers(1, 17, 200)
    for x, k in zip(rows, n):
        w = np.uint32(0)
        for r in range(int(k)):
            w = _dp2a(np.array([1 + (r << 12)], np.uint32), np.array([x[r]], np.uint32), np.array([w], np.uint32), False)[0]
        assert int(w & 0xFFF) == int(x[:k].sum()) and int(w >> 12) == int((np.arange(k) * x[:k]).sum())
