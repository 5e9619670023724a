"""Golden vector for the saveImage / saveHDR hand-off, made by the REFERENCE'S
OWN image.cpp + stb_image_write (oracle/_ref/ref_hdr, see oracle/Makefile).

Two crafted accumulation buffers -- 37x11 (stb writes RLE scanlines from width 8)
and 5x3 (flat scanlines) -- with values around the power-of-two exponent steps,
zeros, tiny values below stb's 1e-32 cut-off and large radiances are pushed through
the reference's saveImage loop (x mirror, / samples) and image::saveHDR; the decoded
RGBE bytes are stored next to the inputs in hdr_golden.npz.  Radiance is never
negative on the path (products of colours in [0,1] and emittances), so the buffers
hold no negative values (the reference's float -> unsigned char cast of a negative
value is undefined behaviour).
Needs /root/reference (through oracle/_ref); the output is committed.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import decode_hdr_rgbe  # noqa: E402

REF_HDR = os.path.join(ROOT, "oracle", "_ref", "ref_hdr")


def craft(w, h, samples, seed):
    rng = np.random.default_rng(seed)
    buf = (rng.uniform(0.0, 1.0, size=(h * w, 3)) ** 4 * 40.0 * samples).astype(np.float32)
    flat = buf.reshape(-1)
    pows = (np.float32(2.0) ** np.arange(-20, 12, dtype=np.float32)) * np.float32(samples)
    k = min(pows.size, flat.size // 4)
    flat[:k] = pows[:k]
    flat[k:2 * k] = np.nextafter(pows[:k], np.float32(0))
    flat[2 * k:3 * k] = np.nextafter(pows[:k], np.float32(1e30))
    flat[-9:] = [0.0, 0.0, 0.0, 1e-33, 1e-35, 0.0, 3e-39, 0.5, 1e-30]
    return buf


def main():
    assert os.path.exists(REF_HDR), "build oracle/_ref first: make -C oracle ref"
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, (w, h, samples) in {"a": (37, 11, 7), "b": (5, 3, 1)}.items():
            buf = craft(w, h, samples, 0x5EED + w)
            raw = os.path.join(tmp, f"{tag}.raw")
            buf.tofile(raw)
            out[f"{tag}_image"] = buf
            out[f"{tag}_dims"] = np.array([w, h, samples], np.int32)
            for name, divide in (("divided", 1), ("plain", 0)):
                base = os.path.join(tmp, f"{tag}_{name}")
                subprocess.check_call([REF_HDR, raw, str(w), str(h), str(samples), str(divide), base],
                                      stdout=subprocess.DEVNULL)
                out[f"{tag}_{name}"] = decode_hdr_rgbe(base + ".hdr")
    np.savez_compressed(os.path.join(HERE, "hdr_golden.npz"), **out)
    print("wrote hdr_golden.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
