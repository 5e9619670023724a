"""Golden vector for the saveImage / savePNG hand-off, made by the REFERENCE'S
OWN image.cpp + stb_image_write (oracle/_ref/ref_png, see oracle/Makefile).

A crafted 37x11 accumulation buffer (values below 0, above 1, exactly on the
k/255 quantisation steps, denormals) is pushed through the reference's
saveImage loop with samples=7 and through the albedo branch (no division);
the decoded PNG pixels are stored next to the input in png_golden.npz.
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
from util import decode_png_rgb8  # noqa: E402

REF_PNG = os.path.join(ROOT, "oracle", "_ref", "ref_png")


def main():
    assert os.path.exists(REF_PNG), "build oracle/_ref first: make -C oracle ref"
    w, h, samples = 37, 11, 7
    rng = np.random.default_rng(0x5EED)
    buf = rng.uniform(-0.5, 1.5 * samples, size=(h * w, 3)).astype(np.float32)
    steps = (np.arange(0, 256, dtype=np.float32) / np.float32(255.0) * np.float32(samples))
    buf.reshape(-1)[: steps.size] = steps                      # exactly on the steps
    buf.reshape(-1)[steps.size: 2 * steps.size] = np.nextafter(steps, np.float32(-1e9))
    buf.reshape(-1)[2 * steps.size: 3 * steps.size] = np.nextafter(steps, np.float32(1e9))
    buf.reshape(-1)[-6:] = [0.0, -0.0, 1e-42, float(samples), np.float32(samples) * 2, -3.0]
    out = {"image": buf, "width": w, "height": h, "samples": samples}
    with tempfile.TemporaryDirectory() as tmp:
        raw = os.path.join(tmp, "in.raw")
        buf.tofile(raw)
        for name, divide in (("divided", 1), ("plain", 0)):
            base = os.path.join(tmp, name)
            subprocess.check_call([REF_PNG, raw, str(w), str(h), str(samples), str(divide), base],
                                  stdout=subprocess.DEVNULL)
            out[f"png_{name}"] = decode_png_rgb8(base + ".png")
    np.savez_compressed(os.path.join(HERE, "png_golden.npz"), **out)
    print("wrote png_golden.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
