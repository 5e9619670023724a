"""Run the reference binaries under ``oracle/_ref``.  TEST INFRASTRUCTURE, NOT PRODUCT.

``oracle/_ref`` holds the reference's own sources compiled by ``oracle/Makefile``
(in the build container, where ``/root/reference`` exists) plus the run tree
the reference loader expects.  The directory is git-ignored but travels to the
GPU box with the snapshot, so nothing here reads ``/root/reference`` at run
time.  Only ``tests/`` and ``bench.py``'s reference legs import this module.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import tempfile
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
RUN_BIN = os.path.join(REF_DIR, "run", "bin")
RUN_SCENES = os.path.join(REF_DIR, "run", "scenes")
RUN_MODELS = os.path.join(REF_DIR, "run", "models")

SPACESHIP_OBJ = "Intergalactic_Spaceship-(Wavefront).obj"


def have(binary: str) -> bool:
    return os.access(os.path.join(REF_DIR, binary), os.X_OK) and os.path.isdir(RUN_BIN)


def scene_variant(name: str, out_path: str, width: Optional[int] = None, height: Optional[int] = None,
                  iterations: Optional[int] = None, depth: Optional[int] = None) -> str:
    """Copy run/scenes/<name>.txt with RES / ITERATIONS / DEPTH overridden
    (every shipped scene says 800x800, 5000, 8 -- SURVEY.md Q29)."""
    with open(os.path.join(RUN_SCENES, name + ".txt")) as f:
        txt = f.read()
    if width and height:
        txt = re.sub(r"(?m)^RES\s+\d+\s+\d+", f"RES         {width} {height}", txt)
    if iterations:
        txt = re.sub(r"(?m)^ITERATIONS\s+\d+", f"ITERATIONS  {iterations}", txt)
    if depth is not None:
        txt = re.sub(r"(?m)^DEPTH\s+\d+", f"DEPTH       {depth}", txt)
    with open(out_path, "w") as f:
        f.write(txt)
    return out_path


def install_spaceship_obj(obj_path: str) -> None:
    """Place a stand-in mesh where cornellObj/cornellSpaceship look for the
    missing 'Intergalactic_Spaceship-(Wavefront).obj' (run/models)."""
    dst = os.path.join(RUN_MODELS, SPACESHIP_OBJ)
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(os.path.abspath(obj_path), dst)


def run(binary: str, scene_txt: str, out_dir: Optional[str] = None, b2s: Optional[str] = None, iters: int = 1,
        iter_first: int = 1, dump_iter: Optional[int] = None, extra=(), timeout: float = 3600.0,
        env: Optional[Dict[str, str]] = None) -> Dict:
    """Run a reference binary from run/bin (the loader's cwd-relative paths)
    and return the JSON it prints on its REF_*_RESULT line."""
    cmd = [os.path.join(REF_DIR, binary), "--scene", os.path.abspath(scene_txt), "--iters", str(iters),
           "--iter-first", str(iter_first)]
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        cmd += ["--out", os.path.abspath(out_dir)]
    if b2s:
        cmd += ["--b2s", os.path.abspath(b2s)]
    if dump_iter is not None:
        cmd += ["--dump-iter", str(dump_iter)]
    cmd += list(extra)
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run(cmd, cwd=RUN_BIN, capture_output=True, text=True, timeout=timeout, env=e)
    if p.returncode != 0:
        raise RuntimeError(f"{binary} failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    for line in p.stdout.splitlines():
        m = re.match(r"REF_(CPU|GPU|ADAPTER)_RESULT (\{.*\})", line)
        if m:
            return json.loads(m.group(2))
    return {}


def load_dump(out_dir: str) -> Dict:
    """Read a stage dump directory: {'nlive', 'image', 'albedo', 'depths': [ {name: array} ]}."""
    res: Dict = {"depths": []}
    for f in sorted(os.listdir(out_dir)):
        if not f.endswith(".npy"):
            continue
        a = np.load(os.path.join(out_dir, f))
        m = re.match(r"d(\d+)_(.*)\.npy", f)
        if m:
            d = int(m.group(1))
            while len(res["depths"]) <= d:
                res["depths"].append({})
            res["depths"][d][m.group(2)] = a
        else:
            res[f[:-4]] = a
    return res


def tmpdir(prefix: str = "b2pt_ref_") -> str:
    return tempfile.mkdtemp(prefix=prefix)
