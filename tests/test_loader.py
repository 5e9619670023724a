"""The scene loader (scene text format, OBJ/MTL, JPEG/PPM textures, camera
orbit) against PODs written by the reference's own loader (tests/golden/*.b2s)."""
import os
import shutil

import numpy as np
import pytest

from mygpuraytracer_b200 import abi, api, assets, scenes, standin_mesh
from mygpuraytracer_b200.podscene import PodScene
from util import GOLDEN, assert_same_bits


def assert_same_scene(ref: PodScene, mine: PodScene, textures=True):
    for f in ref.geoms.dtype.names:
        assert_same_bits(ref.geoms[f], mine.geoms[f], f"geoms.{f}")
    assert ref.materials.tobytes() == mine.materials.tobytes()
    for f in ref.camera.dtype.names:
        assert_same_bits(ref.camera[f], mine.camera[f], f"camera.{f}")
    assert_same_bits(ref.face_pos, mine.face_pos, "face_pos")
    assert_same_bits(ref.face_uv, mine.face_uv, "face_uv")
    assert (ref.trace_depth, ref.iterations) == (mine.trace_depth, mine.iterations)
    if textures:
        assert len(ref.textures) == len(mine.textures)
        for a, b in zip(ref.textures, mine.textures):
            assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("case,name,w,h", [("cornell_32x32", "cornell", 32, 32), ("cornellGlass_32x32", "cornellGlass", 32, 32),
                                           ("sphere_16x16", "sphere", 16, 16), ("cornellGlass_dof_32x24", "cornellGlass", 32, 24)])
def test_scene_files_load_bit_identical_to_reference(tmp_path, case, name, w, h):
    ref = PodScene.load(os.path.join(GOLDEN, case + ".b2s"))
    path = scenes.write_scene(name, str(tmp_path / "scenes" / "s.txt"), width=w, height=h)
    assert_same_scene(ref, api.Scene(path).pod)
    # the same through the RES override of the loader
    path2 = scenes.write_scene(name, str(tmp_path / "scenes" / "s800.txt"))
    assert_same_scene(ref, api.Scene(path2, width=w, height=h).pod)


def test_obj_with_quads_and_no_texture_maps(tmp_path):
    """Quads split along the shorter diagonal like tinyobjloader 2.0.0, MTL
    without maps (the case the reference indexes out of bounds, Q19)."""
    ref = PodScene.load(os.path.join(GOLDEN, "quadbox_32x32.b2s"))
    (tmp_path / "models" / "materials").mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "quadbox.obj"), tmp_path / "models")
    shutil.copy(os.path.join(GOLDEN, "quadbox.mtl"), tmp_path / "models" / "materials")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=32, height=32,
                              obj_path="../models/quadbox.obj")
    mine = api.Scene(path).pod
    assert_same_scene(ref, mine)
    assert len(mine.face_pos) == 14 and mine.geoms["tex_kd"][6] == -1
    assert mine.geoms["material_id"][6] == 6  # appended material (scene.cpp:221-231)


def test_obj_syntax_corner_cases_and_concave_polygons(tmp_path):
    """tests/golden/hardobj.obj: CRLF lines, tabs, exponents and signs, three-component vt, negative
    (relative) indices, several o / g / usemtl / s statements, two MTL materials (the reference uses the
    first of the file for every face) and polygons with 5 - 8 corners, concave, non-planar and with
    collinear corners, which tinyobjloader 2.0.0 ear-clips: same triangles, same order, same bits as the
    reference's loader."""
    ref = PodScene.load(os.path.join(GOLDEN, "hardobj_16x16.b2s"))
    (tmp_path / "models" / "materials").mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "hardobj.obj"), tmp_path / "models")
    shutil.copy(os.path.join(GOLDEN, "hardobj.mtl"), tmp_path / "models" / "materials")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path="../models/hardobj.obj")
    mine = api.Scene(path).pod
    assert_same_scene(ref, mine)
    assert len(mine.face_pos) == 34
    assert abs(float(mine.materials["index_of_refraction"][-1]) - 1.33) < 1e-6  # `newmtl first`, not `second`


@pytest.mark.parametrize("name", ["camera", "tabs_numbers", "permuted_transform", "no_trailing_newline", "triangle_geom",
                                  "rot_scale", "both_refl_refr", "iter_depth"])
def test_scene_format_variants_match_the_reference_loader(name):
    """tests/golden/variants: one property of the scene text format each (see make_scene_variants.py), the
    .b2s written by the reference's own scene.cpp + the camera recompute of main.cpp."""
    ref = PodScene.load(os.path.join(GOLDEN, "variants", name + ".b2s"))
    mine = api.Scene(os.path.join(GOLDEN, "variants", name + ".txt")).pod
    assert_same_scene(ref, mine)
    if name == "triangle_geom":
        assert int(mine.geoms["type"][-1]) == 2 and len(mine.geoms) == 8
    if name == "iter_depth":
        assert (mine.trace_depth, mine.iterations) == (3, 17)


@pytest.mark.parametrize("name", ["minimal", "no_ni", "ke", "tabs_crlf", "comments_unknown", "second_first"])
def test_mtl_variants_match_the_reference_loader(tmp_path, name):
    """tests/golden/mtl_variants (see make_mtl_variants.py): the material the reference appends for an OBJ
    geom, byte for byte -- tinyobjloader's zero defaults, Ke[0] as emittance, only the first `newmtl` counts."""
    want = np.load(os.path.join(GOLDEN, "mtl_variants", name + "_material.npy")).tobytes()
    (tmp_path / "models" / "materials").mkdir(parents=True)
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    obj = obj.replace("mtllib quadbox.mtl", f"mtllib mv_{name}.mtl").replace("usemtl plain", "usemtl a")
    (tmp_path / "models" / f"mv_{name}.obj").write_text(obj)
    shutil.copy(os.path.join(GOLDEN, "mtl_variants", name + ".mtl"), tmp_path / "models" / "materials" / f"mv_{name}.mtl")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/mv_{name}.obj")
    mine = api.Scene(path).pod
    assert mine.materials[-1:].tobytes() == want, mine.materials[-1]


def test_crlf_and_comments_are_tolerated(tmp_path):
    txt = scenes.scene_text("cornell", width=8, height=8)
    p = tmp_path / "crlf.txt"
    p.write_bytes(("// a comment line\r\n" + txt.replace("\n", "\r\n")).encode())
    a = api.Scene(str(p)).pod
    b = api.Scene(scenes.write_scene("cornell", str(tmp_path / "lf.txt"), width=8, height=8)).pod
    assert_same_scene(b, a)


def test_loader_errors(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text("MATERIAL 3\nRGB 1 1 1\n")
    with pytest.raises(api.B2ptError) as e:
        api.Scene(str(p))
    assert e.value.code == -1
    p.write_text(scenes.scene_text("cornellObj", width=8, height=8, obj_path="../models/nope.obj"))
    with pytest.raises(api.B2ptError) as e:
        api.Scene(str(p))
    assert e.value.code == -4


def test_standin_mesh_properties(tmp_path):
    pos, uv, vn, faces = standin_mesh.build(2000)
    assert 1800 <= len(faces) <= 2300
    assert uv.min() > 0.0 and uv.max() < 1.0
    p0, p1, p2 = (pos[faces[:, k]].astype(np.float64) for k in range(3))
    n = np.cross(p1 - p0, p2 - p0)
    assert (np.linalg.norm(n, axis=1) > 0).all()  # no degenerate faces
    # closed and consistently wound: signed volume > 0 and every edge shared by exactly two faces
    assert (p0 * n).sum() / 6.0 > 1.0
    key = lambda a: tuple(np.round(pos[a], 6))
    edges = {}
    for f in faces:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            e = (key(a), key(b))
            edges[e] = edges.get(e, 0) + 1
    assert all(edges.get((b, a), 0) == c == 1 for (a, b), c in edges.items())
    # written and re-read through the loader: same float bits
    root = assets.prepare(str(tmp_path / "run"), triangles=2000, procedural_size=64)
    sc = api.Scene(assets.scene_file("cornellObj", 16, 16, root=root)).pod
    assert len(sc.face_pos) == len(faces)
    assert_same_bits(sc.face_pos.reshape(-1, 3, 3), pos[faces], "OBJ round trip")
    assert_same_bits(sc.face_uv.reshape(-1, 3, 2), uv[faces], "uv round trip")
    assert len(sc.textures) == 4 and sc.textures[0].shape == (64, 64, 3)
    assert sc.materials["index_of_refraction"][-1] == 2.0  # Ni of the spaceship MTL


@pytest.mark.parametrize("name", ["s420_64x48", "s420_37x29", "s422_50x20", "s444_33x17", "s420_2x2", "s420_1x5",
                                  "s420_restart_40x40", "grey_21x13", "prog420_90x70", "prog444_57x43", "prog422_70x50",
                                  "prog_grey_95x77", "prog420_restart_88x72"])
def test_jpeg_sampling_layouts_match_stb_image(tmp_path, name):
    """tests/golden/jpeg (see make_jpeg_golden.py): chroma-subsampled, odd-sized, restart-interval and grey
    baseline files and progressive files of each kind come out byte-identical to what the reference's loader
    (stb_image) holds."""
    want = np.load(os.path.join(GOLDEN, "jpeg", name + ".npy"))
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "jpeg", name + ".jpg"), tmp_path / "textures" / f"jg_{name}.jpg")
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / f"jg_{name}.obj").write_text(obj.replace("mtllib quadbox.mtl", f"mtllib jg_{name}.mtl"))
    (tmp_path / "models" / "materials" / f"jg_{name}.mtl").write_text(
        f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/jg_{name}.jpg\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/jg_{name}.obj")
    mine = api.Scene(path).pod
    assert len(mine.textures) == 1
    assert mine.textures[0].shape == want.shape and np.array_equal(mine.textures[0], want)


@pytest.mark.parametrize("name", ["options_s", "bump_bm", "bump_kw", "map_bump_lower", "all_four", "clamp_blend",
                                  "swallowed_name", "mm_type", "missing_file", "name_with_blanks"])
def test_mtl_texture_statements_match_the_reference_loader(tmp_path, name):
    """tests/golden/mtl_textures (see make_mtl_texture_golden.py): which texture slots of the geom are filled
    and with which texels, for MTL texture options, the bump spellings, names with blanks and missing files."""
    import zlib

    exp = np.load(os.path.join(GOLDEN, "mtl_textures", "expected.npz"))
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    for dst, src in {"opt tex a.png": "rgb_93x71.png", "opt_b.png": "rgba_85x77.png", "opt_c.png": "palette_90x80.png",
                     "opt_d.png": "stored_70x60.png"}.items():
        shutil.copy(os.path.join(GOLDEN, "png", src), tmp_path / "textures" / dst)
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / f"mo_{name}.obj").write_text(obj.replace("mtllib quadbox.mtl", f"mtllib mo_{name}.mtl"))
    shutil.copy(os.path.join(GOLDEN, "mtl_textures", name + ".mtl"), tmp_path / "models" / "materials" / f"mo_{name}.mtl")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/mo_{name}.obj")
    mine = api.Scene(path).pod
    slots = [int(mine.geoms[k][6]) for k in ("tex_kd", "tex_ks", "tex_bump", "tex_ke")]
    crcs = [zlib.crc32(np.ascontiguousarray(t).tobytes()) for t in mine.textures]
    assert slots == exp[name + "_slots"].tolist()
    assert crcs == exp[name + "_crcs"].tolist()


@pytest.mark.parametrize("name", ["vertex_colors", "v_w", "vt_one", "lines_points", "trailing_ws", "no_newline_end",
                                  "vp_and_unknown", "usemtl_unknown", "two_mtllibs", "face_v_vt", "two_corner_face"])
def test_obj_statement_variants_match_the_reference_loader(tmp_path, name):
    """tests/golden/obj_variants (see make_obj_variants.py): faces and the appended material, bit for bit."""
    want = np.load(os.path.join(GOLDEN, "obj_variants", name + ".npz"))
    (tmp_path / "models" / "materials").mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "obj_variants", name + ".obj"), tmp_path / "models" / f"ov_{name}.obj")
    shutil.copy(os.path.join(GOLDEN, "quadbox.mtl"), tmp_path / "models" / "materials")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/ov_{name}.obj")
    mine = api.Scene(path).pod
    assert_same_bits(want["face_pos"], mine.face_pos, "face_pos")
    assert_same_bits(want["face_uv"], mine.face_uv, "face_uv")
    assert mine.materials[-1:].tobytes() == want["material"].tobytes()


def _png_names():
    z = os.path.join(GOLDEN, "png", "texels.npz")
    return sorted(np.load(z).files) if os.path.exists(z) else []


@pytest.mark.parametrize("name", _png_names())
def test_png_layouts_match_stb_image(tmp_path, name):
    """tests/golden/png (see make_png_texture_golden.py): grey / grey+alpha / RGB / RGBA, palettes of 2, 4 and
    8 bits with and without tRNS, 1-bit and 16-bit grey, a transparent colour on RGB and grey, stored and
    dynamic-Huffman deflate streams: channel count and texels byte-identical to the reference's loader."""
    want = np.load(os.path.join(GOLDEN, "png", "texels.npz"))[name]
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "png", name + ".png"), tmp_path / "textures" / f"pg_{name}.png")
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / f"pg_{name}.obj").write_text(obj.replace("mtllib quadbox.mtl", f"mtllib pg_{name}.mtl"))
    (tmp_path / "models" / "materials" / f"pg_{name}.mtl").write_text(
        f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/pg_{name}.png\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/pg_{name}.obj")
    mine = api.Scene(path).pod
    assert len(mine.textures) == 1
    assert mine.textures[0].shape == want.shape and np.array_equal(mine.textures[0], want)


def _tga_names():
    z = os.path.join(GOLDEN, "tga", "texels.npz")
    return sorted(np.load(z).files) if os.path.exists(z) else []


@pytest.mark.parametrize("name", _tga_names())
def test_tga_layouts_match_stb_image(tmp_path, name):
    """tests/golden/tga (see make_tga_golden.py): grey / RGB / RGBA / colour-mapped / RGB555 files, raw and
    run-length encoded, bottom-up and top-down: channel count and texels byte-identical to the reference's loader."""
    want = np.load(os.path.join(GOLDEN, "tga", "texels.npz"))[name]
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "tga", name + ".tga"), tmp_path / "textures" / f"tg_{name}.tga")
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / f"tg_{name}.obj").write_text(obj.replace("mtllib quadbox.mtl", f"mtllib tg_{name}.mtl"))
    (tmp_path / "models" / "materials" / f"tg_{name}.mtl").write_text(
        f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/tg_{name}.tga\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/tg_{name}.obj")
    mine = api.Scene(path).pod
    assert len(mine.textures) == 1
    assert mine.textures[0].shape == want.shape and np.array_equal(mine.textures[0], want)


def _bmp_names():
    z = os.path.join(GOLDEN, "bmp", "texels.npz")
    return sorted(np.load(z).files) if os.path.exists(z) else []


@pytest.mark.parametrize("name", _bmp_names())
def test_bmp_layouts_match_stb_image(tmp_path, name):
    """tests/golden/bmp (see make_bmp_golden.py): 1 / 4 / 8-bit palettes, RGB555, RGB565 bit fields, 24- and 32-bit,
    top-down rows, the OS/2 header, an all-zero alpha channel: channel count and texels byte-identical to the
    reference's loader."""
    want = np.load(os.path.join(GOLDEN, "bmp", "texels.npz"))[name]
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    shutil.copy(os.path.join(GOLDEN, "bmp", name + ".bmp"), tmp_path / "textures" / f"bm_{name}.bmp")
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / f"bm_{name}.obj").write_text(obj.replace("mtllib quadbox.mtl", f"mtllib bm_{name}.mtl"))
    (tmp_path / "models" / "materials" / f"bm_{name}.mtl").write_text(
        f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/bm_{name}.bmp\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16,
                              obj_path=f"../models/bm_{name}.bmp".replace(".bmp", ".obj"))
    mine = api.Scene(path).pod
    assert len(mine.textures) == 1
    assert mine.textures[0].shape == want.shape and np.array_equal(mine.textures[0], want)


def test_png_decoder_rejects_damaged_files(tmp_path):
    """Truncated and corrupted PNG data never crashes the loader and never yields a partial texture."""
    data = open(os.path.join(GOLDEN, "png", "rgb_93x71.png"), "rb").read()
    for d in ("models/materials", "textures"):
        (tmp_path / d).mkdir(parents=True)
    obj = open(os.path.join(GOLDEN, "quadbox.obj")).read()
    (tmp_path / "models" / "bad.obj").write_text(obj.replace("mtllib quadbox.mtl", "mtllib bad.mtl"))
    (tmp_path / "models" / "materials" / "bad.mtl").write_text("newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/bad.png\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=16, height=16, obj_path="../models/bad.obj")
    rng = np.random.default_rng(5)
    cases = [data[:100], data[: len(data) // 2], data[:-20]]
    huge = bytearray(data)
    huge[16:24] = bytes([0x7f, 0xff, 0xff, 0xff, 0x7f, 0xff, 0xff, 0xff])  # IHDR width and height 2^31 - 1
    cases.append(bytes(huge))
    for _ in range(12):
        b = bytearray(data)
        for k in rng.integers(60, len(b) - 16, 8):
            b[k] ^= int(rng.integers(1, 256))
        cases.append(bytes(b))
    for blob in cases:
        (tmp_path / "textures" / "bad.png").write_bytes(blob)
        sc = api.Scene(path)
        pod = sc.pod
        # a texture that does not load leaves the geom without one, like the reference ("Failed to load Kd
        # texture file", scene.cpp:150-154); a flipped bit in the pixel data may still decode: then the shape holds
        assert len(pod.textures) == 0 or pod.textures[0].shape == (71, 93, 3)
        if len(pod.textures) == 0:
            assert int(pod.geoms["tex_kd"][6]) == -1
            # ... and the loader says so (b2pt_scene_warnings), as the reference prints the name
            assert len(sc.warnings) == 1 and "bad.png" in sc.warnings[0] and "could not be decoded" in sc.warnings[0]
        else:
            assert sc.warnings == []
    os.remove(tmp_path / "textures" / "bad.png")
    sc = api.Scene(path)
    assert len(sc.pod.textures) == 0 and len(sc.warnings) == 1 and "not found" in sc.warnings[0]


def test_jpeg_decoder_matches_reference_texels():
    """4:4:4 baseline JPEG -> the same bytes stb_image produces.  The golden
    is a CRC of texels dumped by the reference loader (see make_golden.py);
    it needs the reference JPEGs, which only exist where the build copied them."""
    import zlib
    tex = os.path.join(assets.DEFAULT_ROOT, "textures", "Intergalactic Spaceship_emi.jpg")
    crc_file = os.path.join(GOLDEN, "texture_crc.txt")
    if not (os.path.exists(tex) and os.path.exists(crc_file)):
        pytest.skip("reference textures not present")
    want = dict(l.split() for l in open(crc_file).read().splitlines())
    sc = api.Scene(assets.scene_file("cornellObj", 8, 8)).pod
    names = ["kd", "ks", "bump", "ke"]
    for n, t in zip(names, sc.textures):
        assert t.shape == (4096, 4096, 3)
        assert f"{zlib.crc32(t.tobytes()):08x}" == want[n], n


def test_camera_block_is_validated(tmp_path):
    """RES and DEPTH the renderer could not accept are refused by the loader (the reference allocates blindly)."""
    base = scenes.scene_text("cornell", width=8, height=8)
    for bad, code in (("RES         -8 8", -1), ("RES         0 8", -1), ("RES         999999999 16", -1),
                      ("DEPTH       -2", -6), ("DEPTH       63", -6)):
        key = bad.split()[0]
        txt = "\n".join(bad if line.startswith(key + " ") else line for line in base.split("\n"))
        p = tmp_path / "bad.txt"
        p.write_text(txt)
        with pytest.raises(api.B2ptError) as e:
            api.Scene(str(p))
        assert e.value.code == code, bad


@pytest.mark.parametrize("what", ["scene", "obj", "mtl"])
def test_text_parsers_survive_mutations(tmp_path, what):
    """Byte noise, deleted spans, injected tokens (huge numbers, NaN, NUL, stray keywords) and replaced lines in
    the scene file, the OBJ and the MTL: the loader answers with a scene or an error code, never a crash
    (6000 + 3000 mutations were clean when this was written; this keeps a short run in the suite)."""
    for d in ("models/materials", "scenes"):
        (tmp_path / d).mkdir(parents=True)
    obj = open(os.path.join(GOLDEN, "hardobj.obj"), "rb").read().replace(b"hardobj.mtl", b"f.mtl")
    mtl = open(os.path.join(GOLDEN, "hardobj.mtl"), "rb").read()
    scene = scenes.scene_text("cornellObj", width=16, height=16, obj_path="../models/f.obj").encode()
    files = {"scene": (tmp_path / "scenes" / "s.txt", scene), "obj": (tmp_path / "models" / "f.obj", obj),
             "mtl": (tmp_path / "models" / "materials" / "f.mtl", mtl)}
    for path, data in files.values():
        path.write_bytes(data)
    tokens = [b"-1", b"0", b"999999999", b"-999999999", b"1e309", b"nan", b"", b" ", b"\n", b"/", b"//", b"f", b"v", b"vt",
              b"OBJECT 3", b"MATERIAL 9", b"CAMERA", b"\x00", b"\xff", b"1/2/3/4", b"-0", b"2147483648", b"obj", b"usemtl",
              b"mtllib x y z"]
    rng = np.random.default_rng(20261018)
    target, original = files[what]
    loaded = 0
    for _ in range(250):
        b = bytearray(original)
        for _k in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, len(b)))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                b[k] = int(rng.integers(0, 256))
            elif mode == 1:
                del b[k:k + int(rng.integers(1, 20))]
            elif mode == 2:
                b[k:k] = tokens[int(rng.integers(0, len(tokens)))]
            else:
                j = b.find(b"\n", k)
                j = len(b) if j < 0 else j
                i0 = b.rfind(b"\n", 0, k) + 1
                b[i0:j] = tokens[int(rng.integers(0, len(tokens)))] + b" " + tokens[int(rng.integers(0, len(tokens)))]
        target.write_bytes(bytes(b))
        try:
            pod = api.Scene(str(files["scene"][0])).pod
            loaded += 1
            assert 0 < pod.n_pixels < (1 << 30) and 0 <= pod.trace_depth <= 62
        except api.B2ptError as e:
            assert e.code < 0
    assert loaded > 0


def _multimat_tree(tmp_path):
    (tmp_path / "models" / "materials").mkdir(parents=True)
    (tmp_path / "textures").mkdir()
    shutil.copy(os.path.join(GOLDEN, "multimat.obj"), tmp_path / "models")
    shutil.copy(os.path.join(GOLDEN, "multimat.mtl"), tmp_path / "models" / "materials")
    for f in os.listdir(os.path.join(GOLDEN, "texquad")):
        shutil.copy(os.path.join(GOLDEN, "texquad", f), tmp_path / "textures")
    return scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=32, height=32, obj_path="../models/multimat.obj")


def test_per_face_materials_match_tinyobj(tmp_path):
    """per_face_materials=1 keeps what the reference reads and discards (apps/src/scene.cpp:121-122): the ids are
    tinyobjloader's own (tests/golden/multimat_tinyobj.json, from the reference's vendored copy), every MTL material
    becomes a scene material converted like material 0 (scene.cpp:221-231) with its own four maps."""
    import json

    gold = json.load(open(os.path.join(GOLDEN, "multimat_tinyobj.json")))
    path = _multimat_tree(tmp_path)
    plain = api.Scene(path).pod
    pod = api.Scene(path, per_face_materials=True).pod
    assert plain.face_material is None and plain.material_textures is None
    n_mtl = len(gold["materials"])
    base = len(plain.materials) - 1                      # the reference appends ONE material per OBJ
    assert len(pod.materials) == base + n_mtl
    assert np.array_equal(pod.materials[: base + 1], plain.materials)   # material 0 is the reference's
    g = int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0])
    assert pod.geoms["material_id"][g] == base
    fb, fc = int(pod.geoms["face_begin"][g]), int(pod.geoms["face_count"][g])
    assert fc == len(gold["material_ids"])
    want = np.array([base + max(i, 0) for i in gold["material_ids"]], np.int32)   # no / unknown usemtl -> material 0
    assert np.array_equal(pod.face_material[fb: fb + fc], want)
    for k, m in enumerate(gold["materials"]):
        sm = pod.materials[base + k]
        assert np.array_equal(sm["color"], np.array(m["kd"], np.float32))
        assert np.array_equal(sm["specular_color"], np.array(m["ks"], np.float32))
        assert sm["index_of_refraction"] == np.float32(m["ior"]) and sm["emittance"] == np.float32(m["ke"][0])
        assert sm["has_reflective"] == 0 and sm["has_refractive"] == 0 and sm["specular_exponent"] == 0
        has_maps = bool(m["map_kd"])
        assert all((t >= 0) == has_maps for t in pod.material_textures[base + k])
    assert (pod.material_textures[:base] == -1).all()
    # the geom keeps material 0's maps (none here), as in reference mode
    assert pod.geoms["tex_kd"][g] == plain.geoms["tex_kd"][g] == -1
    # one decoded image per file, however many materials name it
    assert len(pod.textures) == 4


def test_per_face_materials_change_the_render_and_roundtrip_b2s(tmp_path):
    """The oracle defines the behaviour: a face is shaded with ITS material and maps; with the ids dropped the same
    scene renders as the reference does.  The .b2s container keeps the two arrays."""
    from mygpuraytracer_b200.podscene import PodScene
    from oracle import oracle

    path = _multimat_tree(tmp_path)
    pod = api.Scene(path, per_face_materials=True, width=24, height=24).pod
    plain = api.Scene(path, width=24, height=24).pod
    a = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 2, 1)
    b = oracle.render(plain, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 2, 1)
    assert not np.array_equal(a[0], b[0]) and not np.array_equal(a[1], b[1])
    # dropping the ids (every face -> material 0, no per-material maps) is the reference's scene
    same = pod.copy()
    same.face_material = None
    same.material_textures = None
    c = oracle.render(same, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 2, 1)
    assert np.array_equal(c[0], b[0]) and np.array_equal(c[1], b[1])
    f = str(tmp_path / "mm.b2s")
    pod.save(f)
    back = PodScene.load(f)
    assert np.array_equal(back.face_material, pod.face_material) and np.array_equal(back.material_textures, pod.material_textures)
    assert_same_scene(pod, back)
