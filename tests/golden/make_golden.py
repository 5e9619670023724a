"""Generate the golden fixtures from the REFERENCE'S OWN CODE.

Runs ``oracle/_ref/ref_cpu`` -- a host build of the reference's unmodified
intersections.h / interactions.h / scene.cpp (see oracle/Makefile) -- on small
renders of the reference's scenes and stores

  <case>.b2s          the scene as the reference's loader produced it
  <case>_stages.npz   every stage array of iteration 1 and the accumulated
                      image / albedo after two iterations

Needs /root/reference (through oracle/_ref), so it only runs in the build
container; the outputs are committed.  Usage: python tests/golden/make_golden.py
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import harness  # noqa: E402
from mygpuraytracer_b200 import scenes  # noqa: E402

# case -> (scene, W, H, extra ref_cpu flags)
CASES = {
    "cornell_32x32": ("cornell", 32, 32, []),
    "cornellGlass_32x32": ("cornellGlass", 32, 32, []),
    "cornellGlass_dof_32x24": ("cornellGlass", 32, 24, ["--dof"]),
    "cornellGlass_noaa_24x32": ("cornellGlass", 24, 32, ["--no-aa"]),
    "sphere_16x16": ("sphere", 16, 16, []),
    "quadbox_32x32": ("quadbox", 32, 32, []),
    # OBJ syntax corner cases and concave polygons (tinyobjloader's ear clipping), see hardobj.obj
    "hardobj_16x16": ("hardobj", 16, 16, []),
    # scene-format variants (make_scene_variants.py) whose geometry / materials differ from the shipped scenes:
    # rotated geoms with a non-uniform scale (an ellipsoid), a material that is reflective AND refractive
    "rot_scale_48x20": ("variant:rot_scale", 48, 20, []),
    "both_refl_refr_48x20": ("variant:both_refl_refr", 48, 20, []),
    # quadbox with all four texture maps (texquad.mtl, PNG fixtures in texquad/): bump mapping, emission texels, kd / ks
    "texquad_32x32": ("texquad", 32, 32, []),
    # the same box with a 1-channel kd and a 2-channel ks map: three bytes are read per texel whatever the channel count
    "greyquad_32x32": ("greyquad", 32, 32, []),
}


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    for f in ("quadbox", "hardobj", "texquad", "greyquad"):
        shutil.copyfile(os.path.join(HERE, f + ".obj"), os.path.join(harness.RUN_MODELS, f + ".obj"))
        shutil.copyfile(os.path.join(HERE, f + ".mtl"), os.path.join(harness.RUN_MODELS, "materials", f + ".mtl"))
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    for f in os.listdir(os.path.join(HERE, "texquad")):
        shutil.copyfile(os.path.join(HERE, "texquad", f), os.path.join(tex_dir, f))
    for case, (scene, w, h, extra) in CASES.items():
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        if scene.startswith("variant:"):
            shutil.copyfile(os.path.join(HERE, "variants", scene.split(":")[1] + ".txt"), txt)
        elif scene in ("quadbox", "hardobj", "texquad", "greyquad"):
            with open(txt, "w") as f:
                f.write(scenes.scene_text("cornellObj", width=w, height=h, obj_path=f"../models/{scene}.obj"))
        else:
            harness.scene_variant(scene, txt, w, h)
        b2s = os.path.join(HERE, case + ".b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=2, dump_iter=1, extra=extra)
        dump = harness.load_dump(os.path.join(d, "out"))
        arrays = {"image": dump["image"], "albedo": dump["albedo"], "nlive": dump["nlive"]}
        for k, st in enumerate(dump["depths"]):
            for name, a in st.items():
                arrays[f"d{k}_{name}"] = a
        np.savez_compressed(os.path.join(HERE, case + "_stages.npz"), **arrays)
        print(case, os.path.getsize(b2s), os.path.getsize(os.path.join(HERE, case + "_stages.npz")))
        shutil.rmtree(d)


def texture_crc():
    """texture_crc.txt: CRC-32 of the four 4096x4096 spaceship textures as the reference's loader (stb_image)
    holds them, for tests/test_loader.py::test_jpeg_decoder_matches_reference_texels."""
    import zlib

    from mygpuraytracer_b200 import assets, standin_mesh
    from mygpuraytracer_b200.podscene import PodScene

    root = assets.prepare()
    harness.install_spaceship_obj(standin_mesh.ensure_obj(os.path.join(root, "models"), 1000))
    d = harness.tmpdir()
    txt = os.path.join(d, "s.txt")
    harness.scene_variant("cornellObj", txt, 16, 16)
    b2s = os.path.join(d, "s.b2s")
    harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
    ref = PodScene.load(b2s)
    with open(os.path.join(HERE, "texture_crc.txt"), "w") as f:
        for name, tex in zip(["kd", "ks", "bump", "ke"], ref.textures):
            f.write(f"{name} {zlib.crc32(tex.tobytes()):08x}\n")
    shutil.rmtree(d)
    print(open(os.path.join(HERE, "texture_crc.txt")).read())


if __name__ == "__main__":
    main()
    texture_crc()
