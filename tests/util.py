"""Helpers shared by the tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bits(a):
    a = np.ascontiguousarray(np.squeeze(a))
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same_bits(a, b) -> bool:
    a, b = bits(a), bits(b)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def assert_same_bits(a, b, what=""):
    a2, b2 = bits(a), bits(b)
    assert a2.shape == b2.shape, f"{what}: shape {a2.shape} vs {b2.shape}"
    bad = int((a2 != b2).sum())
    assert bad == 0, f"{what}: {bad} of {a2.size} elements differ bitwise"


def load_golden(case):
    from mygpuraytracer_b200.podscene import PodScene

    scene = PodScene.load(os.path.join(GOLDEN, case + ".b2s"))
    z = np.load(os.path.join(GOLDEN, case + "_stages.npz"))
    depths = []
    k = 0
    while f"d{k}_hit_t" in z:
        depths.append({n[len(f"d{k}_"):]: z[n] for n in z.files if n.startswith(f"d{k}_")})
        k += 1
    return scene, depths, z["image"], z["albedo"], z["nlive"]


# oracle stage name -> golden (reference dump) stage name
STAGE_MAP = [
    ("ray_origin", "in_origin"), ("ray_dir", "in_dir"), ("ray_pixel", "in_pixel"), ("hit_t", "hit_t"),
    ("hit_normal", "hit_normal"), ("hit_material", "hit_material"), ("sorted_pixel", "sorted_pixel"),
    ("shaded_color", "shaded_color"), ("shaded_bounces", "shaded_bounces"), ("shaded_origin", "shaded_origin"),
    ("shaded_dir", "shaded_dir"), ("partition_pixel", "part_pixel"),
]


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def decode_png_rgb8(path):
    """Minimal PNG decoder (8-bit RGB, non-interlaced, all five filter types) -> (H, W, 3) uint8."""
    import struct
    import zlib

    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        assert zlib.crc32(typ + body) == crc, f"bad CRC in chunk {typ}"
        if typ == b"IHDR":
            w, h, depth, ctype, _c, _f, interlace = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, interlace) == (8, 2, 0), "expected 8-bit RGB, non-interlaced"
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8)
    stride = w * 3
    raw = raw.reshape(h, stride + 1)
    out = np.zeros((h, stride), np.int32)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        cur = np.zeros(stride, np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            for i in range(stride):
                a = cur[i - 3] if i >= 3 else 0
                b = prev[i]
                c = prev[i - 3] if i >= 3 else 0
                if ft == 1:
                    p = a
                elif ft == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + p) & 255
        out[y] = cur
        prev = cur
    return out.astype(np.uint8).reshape(h, w, 3)


def decode_hdr_rgbe(path):
    """Minimal Radiance .hdr reader -> (H, W, 4) uint8 RGBE (-Y h +X w; new-style RLE or flat scanlines)."""
    data = open(path, "rb").read()
    assert data.startswith(b"#?RADIANCE") or data.startswith(b"#?RGBE"), "not a Radiance file"
    end = data.index(b"\n\n")
    assert b"FORMAT=32-bit_rle_rgbe" in data[:end], "expected 32-bit_rle_rgbe"
    nl = data.index(b"\n", end + 2)
    dims = data[end + 2:nl].split()
    assert dims[0] == b"-Y" and dims[2] == b"+X", dims
    h, w = int(dims[1]), int(dims[3])
    pos = nl + 1
    out = np.zeros((h, w, 4), np.uint8)
    for y in range(h):
        rle = 8 <= w < 32768 and data[pos] == 2 and data[pos + 1] == 2 and ((data[pos + 2] << 8) | data[pos + 3]) == w
        if not rle:
            out[y] = np.frombuffer(data, np.uint8, w * 4, pos).reshape(w, 4)
            pos += w * 4
            continue
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                n = data[pos]
                pos += 1
                if n > 128:
                    n -= 128
                    out[y, x:x + n, c] = data[pos]
                    pos += 1
                else:
                    out[y, x:x + n, c] = np.frombuffer(data, np.uint8, n, pos)
                    pos += n
                x += n
            assert x == w, "RLE run crosses the scanline"
    assert pos == len(data), f"{len(data) - pos} trailing bytes"
    return out
