"""Case list and seeded input generators shared by make_golden.py (which stores the
expected outputs) and tests/test_golden.py (which regenerates the inputs)."""
import numpy as np

# name, source kind, width, height, seed, ops
RESAMPLE_CASES = [
    ("rgba_opaque_4x3", "rgba", 400, 300, 11, [("resize", 102, 76), ("thumb", 20)]),
    ("rgba_premul", "rgba_premul", 333, 251, 12, [("resize", 85, 64), ("thumb", 33)]),
    ("rgba_portrait", "rgba", 240, 320, 13, [("resize", 57, 76), ("thumb", 24)]),
    ("rgba_upscale", "rgba", 64, 48, 14, [("resize", 101, 76)]),
    ("rgba_identity", "rgba_premul", 50, 40, 15, [("resize", 50, 40)]),
    ("nrgba", "nrgba", 210, 160, 16, [("resize", 70, 53), ("thumb", 16)]),
    ("gray", "gray", 301, 203, 17, [("resize", 64, 43), ("thumb", 25)]),
    ("ycbcr444", "ycbcr444", 203, 151, 18, [("resize", 51, 38), ("thumb", 19)]),
    ("ycbcr422", "ycbcr422", 203, 151, 19, [("resize", 51, 38), ("thumb", 19)]),
    ("ycbcr420", "ycbcr420", 203, 151, 20, [("resize", 51, 38), ("thumb", 19)]),
    ("ycbcr440", "ycbcr440", 203, 151, 21, [("resize", 51, 38), ("thumb", 19)]),
    ("rgba_12mp_ratio", "rgba", 1000, 750, 22, [("resize", 256, 192), ("thumb", 50)]),
]

# name, width, height, seed, color.RGBA bytes (not premultiplied), glyph count
BLEND_CASES = [
    ("white127", 160, 64, 31, (255, 255, 255, 127), 7),
    ("green200", 160, 64, 32, (10, 200, 30, 200), 7),
    ("black127", 97, 50, 33, (0, 0, 0, 127), 5),
    ("opaque_red", 97, 50, 34, (255, 0, 0, 255), 5),
]


def chroma_shape(kind, w, h):
    if kind == "ycbcr444":
        return h, w
    if kind == "ycbcr422":
        return h, (w + 1) // 2
    if kind == "ycbcr420":
        return (h + 1) // 2, (w + 1) // 2
    return (h + 1) // 2, w


def make_source(kind, w, h, seed):
    """Planes of the decoded source (uint8), as image.Decode would hand them over."""
    rng = np.random.default_rng(seed)
    if kind in ("rgba", "rgba_premul", "nrgba"):
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        if kind == "rgba":
            a[..., 3] = 255
        elif kind == "rgba_premul":
            al = a[..., 3:4].astype(np.uint16)
            a[..., :3] = (a[..., :3].astype(np.uint16) * al // 255).astype(np.uint8)
        return (a,)
    if kind == "gray":
        return (rng.integers(0, 256, (h, w), dtype=np.uint8),)
    ch, cw = chroma_shape(kind, w, h)
    return (rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8),
            rng.integers(0, 256, (ch, cw), dtype=np.uint8))


def source_kind_base(kind):
    return "rgba" if kind == "rgba_premul" else kind


def make_blend_case(w, h, seed, n):
    """(dst RGBA8, [(x0,y0,x1,y1,mask,mp_x,mp_y)]): overlapping glyph boxes, clipped to the image,
    mask point as freetype passes it (mp.X = 0, mp.Y = rows clipped off the top)."""
    rng = np.random.default_rng(seed)
    dst = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    glyphs = []
    for k in range(n):
        gw, gh = int(rng.integers(10, 22)), int(rng.integers(14, 30))
        m = rng.integers(0, 256, (gh, gw), dtype=np.uint8)
        m[rng.random((gh, gw)) < 0.3] = 0
        gx, gy = 6 + 13 * k, int(rng.integers(-6, h - 20))      # 13 px pitch < width: neighbours overlap
        x0, y0, x1, y1 = max(gx, 0), max(gy, 0), min(gx + gw, w), min(gy + gh, h)
        if x0 < x1 and y0 < y1:
            glyphs.append((x0, y0, x1, y1, m, 0, y0 - gy))
    return dst, glyphs
