"""saveImage / savePNG and the denoiser inputs.

CPU: the numpy restatement (oracle.save_image_rgb8) against pixels decoded
from PNGs the REFERENCE'S OWN image.cpp + stb_image_write produced
(tests/golden/png_golden.npz, made by tests/golden/make_png_golden.py).
GPU: b2pt_resolve_rgb8 / b2pt_save_png / b2pt_resolve_color through the C ABI
against that restatement, bit for bit, on the golden buffer and on a render.
"""
import os
import subprocess

import numpy as np
import pytest

from mygpuraytracer_b200 import abi, api, scenes
from oracle import oracle
from util import GOLDEN, ROOT, assert_same_bits, decode_png_rgb8

REF_PNG = os.path.join(ROOT, "oracle", "_ref", "ref_png")


def _golden():
    z = np.load(os.path.join(GOLDEN, "png_golden.npz"))
    return z["image"], int(z["width"]), int(z["height"]), int(z["samples"]), z["png_divided"], z["png_plain"]


def test_oracle_quantiser_matches_the_reference_png_pixels():
    img, w, h, samples, divided, plain = _golden()
    assert np.array_equal(oracle.save_image_rgb8(img, w, h, samples, divide=True), divided)
    assert np.array_equal(oracle.save_image_rgb8(img, w, h, samples, divide=False), plain)
    # the mirror is part of saveImage (main.cpp:126), not of the quantiser
    assert np.array_equal(oracle.save_image_rgb8(img, w, h, samples, mirror_x=False)[:, ::-1], divided)


def test_png_decoder_handles_every_filter_type(tmp_path):
    """The test-side decoder is itself checked against zlib-built PNGs with filters 0-4."""
    import struct
    import zlib

    rng = np.random.default_rng(3)
    w, h = 13, 9
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    rows = b""
    prev = np.zeros(w * 3, np.int32)
    for y in range(h):
        cur = img[y].reshape(-1).astype(np.int32)
        ft = y % 5
        line = np.zeros_like(cur)
        for i in range(cur.size):
            a = cur[i - 3] if i >= 3 else 0
            b = prev[i]
            c = prev[i - 3] if i >= 3 else 0
            if ft == 0:
                p = 0
            elif ft == 1:
                p = a
            elif ft == 2:
                p = b
            elif ft == 3:
                p = (a + b) >> 1
            else:
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            line[i] = (cur[i] - p) & 255
        rows += bytes([ft]) + line.astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b"")
    path = tmp_path / "t.png"
    path.write_bytes(png)
    assert np.array_equal(decode_png_rgb8(str(path)), img)


def test_png_writer_container_roundtrip(tmp_path):
    """csrc/host/png_writer.h on its own (compiled with g++, no GPU): chunk CRCs, the stored-deflate stream
    and the Adler-32 are accepted by zlib, and the pixels come back unchanged -- including an image whose raw
    size crosses several 65535-byte stored blocks."""
    src = tmp_path / "t.cpp"
    src.write_text(
        '#include <stdio.h>\n#include <stdlib.h>\n#include <vector>\n'
        f'#include "{ROOT}/mygpuraytracer_b200/csrc/host/png_writer.h"\n'
        'int main(int argc, char** argv) { int w = atoi(argv[2]), h = atoi(argv[3]); std::vector<uint8_t> px((size_t)w * h * 3);\n'
        '  FILE* f = fopen(argv[1], "rb"); if (!f || fread(px.data(), 1, px.size(), f) != px.size()) return 2; fclose(f);\n'
        '  std::string e = b2pt_host::write_png_rgb8(argv[4], w, h, px.data()); if (!e.empty()) { puts(e.c_str()); return 1; }\n'
        '  return b2pt_host::write_png_rgb8("/nonexistent_dir/x.png", w, h, px.data()).empty() ? 3 : 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O1", str(src), "-o", str(exe)])
    rng = np.random.default_rng(11)
    for w, h in ((1, 1), (7, 3), (301, 217)):  # 301*217*3 + 217 = 196 168 raw bytes: four stored blocks
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        raw, out = tmp_path / "in.raw", tmp_path / f"o_{w}.png"
        img.tofile(raw)
        subprocess.check_call([str(exe), str(raw), str(w), str(h), str(out)])
        assert np.array_equal(decode_png_rgb8(str(out)), img)


@pytest.mark.gpu
def test_resolve_and_png_on_the_golden_buffer(tmp_path):
    """Crafted accumulation buffer (values on, just below and just above every
    k/255 step) injected with b2pt_set_device_image: the CUDA quantiser equals
    the reference's PNG pixels."""
    import torch

    img, w, h, samples, divided, plain = _golden()
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=w, height=h)).pod
    dev = torch.from_numpy(img.copy()).cuda()
    with api.Renderer(pod, abi.default_options()) as r:
        api._check(r.lib.b2pt_set_device_image(r._h, dev.data_ptr()))
        assert np.array_equal(r.resolve_rgb8(abi.AOV_IMAGE, samples, True), divided)
        assert np.array_equal(r.resolve_rgb8(abi.AOV_IMAGE, 1, True), plain)
        assert np.array_equal(r.resolve_rgb8(abi.AOV_IMAGE, samples, False)[:, ::-1], divided)
        out = str(tmp_path / "g.png")
        r.save_png(out, abi.AOV_IMAGE, samples)
        assert np.array_equal(decode_png_rgb8(out), divided)
        assert_same_bits(r.resolve_color(samples), oracle.denoise_color(img, samples), "denoiser colour input")
        with pytest.raises(api.B2ptError):
            r.resolve_rgb8(abi.AOV_IMAGE, 0, True)
        with pytest.raises(api.B2ptError):
            r.save_png(str(tmp_path / "no_such_dir" / "x.png"), abi.AOV_IMAGE, 1)
        api._check(r.lib.b2pt_set_device_image(r._h, None))


@pytest.mark.gpu
def test_saved_render_matches_oracle_and_reference_writer(tmp_path):
    pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=96, height=64)).pod
    n = 12
    with api.Renderer(pod, abi.default_options()) as r:
        r.render(1, n, 1)
        img, alb = r.read()
        png_img, png_alb = str(tmp_path / "img.png"), str(tmp_path / "alb.png")
        r.save_png(png_img, abi.AOV_IMAGE, n)
        r.save_png(png_alb, abi.AOV_ALBEDO, n)
        color = r.resolve_color(n)
    want_img = oracle.save_image_rgb8(img, 96, 64, n, divide=True)
    want_alb = oracle.save_image_rgb8(alb, 96, 64, n, divide=False)
    assert np.array_equal(decode_png_rgb8(png_img), want_img)
    assert np.array_equal(decode_png_rgb8(png_alb), want_alb)
    assert want_img.max() > 200 and want_img.std() > 10, "the render is not blank"
    assert_same_bits(color, oracle.denoise_color(img, n), "denoiser colour input")
    if os.path.exists(REF_PNG):  # the reference's own writer on the same accumulation buffer
        raw = str(tmp_path / "in.raw")
        img.astype(np.float32).tofile(raw)
        subprocess.check_call([REF_PNG, raw, "96", "64", str(n), "1", str(tmp_path / "ref")], stdout=subprocess.DEVNULL)
        assert np.array_equal(decode_png_rgb8(str(tmp_path / "ref.png")), want_img)


# ---- saveImage + image::saveHDR (apps/src/image.cpp:41-45; the call is commented out at main.cpp:163) ----------
REF_HDR = os.path.join(ROOT, "oracle", "_ref", "ref_hdr")


def _hdr_golden():
    z = np.load(os.path.join(GOLDEN, "hdr_golden.npz"))
    for tag in "ab":
        w, h, samples = (int(x) for x in z[f"{tag}_dims"])
        yield z[f"{tag}_image"], w, h, samples, z[f"{tag}_divided"], z[f"{tag}_plain"]


def test_oracle_rgbe_matches_the_reference_hdr_pixels():
    """The numpy restatement of stbi_write_hdr's pixel encoding against RGBE bytes decoded from files the
    REFERENCE'S OWN image.cpp + stb_image_write wrote (RLE scanlines at width 37, flat ones at width 5)."""
    from util import decode_hdr_rgbe  # noqa: F401  (the decoder made the golden arrays)

    for img, w, h, samples, divided, plain in _hdr_golden():
        assert np.array_equal(oracle.save_image_rgbe(img, w, h, samples, divide=True), divided)
        assert np.array_equal(oracle.save_image_rgbe(img, w, h, samples, divide=False), plain)
        assert np.array_equal(oracle.save_image_rgbe(img, w, h, samples, mirror_x=False)[:, ::-1], divided)
        assert divided[..., 3].max() > 128 and (divided == 0).all(axis=-1).any(), "large radiances and the 1e-32 cut-off are covered"


def test_hdr_writer_on_its_own(tmp_path):
    """csrc/host/hdr_writer.h compiled with g++ (no GPU): the files decode to the reference's RGBE pixels."""
    from util import decode_hdr_rgbe

    src = tmp_path / "t.cpp"
    src.write_text(
        '#include <stdio.h>\n#include <stdlib.h>\n#include <vector>\n'
        f'#include "{ROOT}/mygpuraytracer_b200/csrc/host/hdr_writer.h"\n'
        'int main(int argc, char** argv) { int w = atoi(argv[2]), h = atoi(argv[3]); std::vector<float> px((size_t)w * h * 3);\n'
        '  FILE* f = fopen(argv[1], "rb"); if (!f || fread(px.data(), 4, px.size(), f) != px.size()) return 2; fclose(f);\n'
        '  std::string e = b2pt_host::write_hdr_rgb(argv[4], w, h, px.data()); if (!e.empty()) { puts(e.c_str()); return 1; }\n'
        '  return b2pt_host::write_hdr_rgb("/nonexistent_dir/x.hdr", w, h, px.data()).empty() ? 3 : 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", str(src), "-o", str(exe)])
    for k, (img, w, h, samples, divided, plain) in enumerate(_hdr_golden()):
        # the writer takes the pixels saveImage hands to image::saveHDR: mirrored, not divided here
        px = img.reshape(h, w, 3)[:, ::-1, :].astype(np.float32)
        raw, out = tmp_path / f"in{k}.raw", tmp_path / f"o{k}.hdr"
        np.ascontiguousarray(px).tofile(raw)
        subprocess.check_call([str(exe), str(raw), str(w), str(h), str(out)])
        assert np.array_equal(decode_hdr_rgbe(str(out)), plain)


@pytest.mark.gpu
def test_save_hdr_matches_oracle_and_reference_writer(tmp_path):
    from util import decode_hdr_rgbe

    pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=96, height=64)).pod
    n = 12
    with api.Renderer(pod, abi.default_options()) as r:
        r.render(1, n, 1)
        img, alb = r.read()
        hdr_img, hdr_alb = str(tmp_path / "img.hdr"), str(tmp_path / "alb.hdr")
        r.save_hdr(hdr_img, abi.AOV_IMAGE, n)
        r.save_hdr(hdr_alb, abi.AOV_ALBEDO, n)
        with pytest.raises(api.B2ptError):
            r.save_hdr(str(tmp_path / "no_such_dir" / "x.hdr"), abi.AOV_IMAGE, 1)
        with pytest.raises(api.B2ptError):
            r.save_hdr(hdr_img, abi.AOV_IMAGE, 0)
    want_img = oracle.save_image_rgbe(img, 96, 64, n, divide=True)
    want_alb = oracle.save_image_rgbe(alb, 96, 64, n, divide=False)
    assert np.array_equal(decode_hdr_rgbe(hdr_img), want_img)
    assert np.array_equal(decode_hdr_rgbe(hdr_alb), want_alb)
    assert want_img[..., 3].max() >= 128, "the render is not blank"
    if os.path.exists(REF_HDR):  # the reference's own writer on the same accumulation buffer
        raw = str(tmp_path / "in.raw")
        img.astype(np.float32).tofile(raw)
        subprocess.check_call([REF_HDR, raw, "96", "64", str(n), "1", str(tmp_path / "ref")], stdout=subprocess.DEVNULL)
        assert np.array_equal(decode_hdr_rgbe(str(tmp_path / "ref.hdr")), want_img)
