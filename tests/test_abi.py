"""The C-ABI library loads and exports every symbol include/b2pt.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re

from mygpuraytracer_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b2pt.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2pt_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()  # raises AttributeError naming missing exports
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libb2pt.so lacks {n}"
    # and the Python mirror declares exactly the header's functions
    assert sorted(abi.SYMBOLS) == names


def test_abi_version_and_defaults():
    lib = api.load_library()
    assert lib.b2pt_abi_version() == 2
    o = abi.Options()
    lib.b2pt_default_options(C.byref(o))
    d = abi.default_options()
    assert o.struct_size == C.sizeof(abi.Options) == d.struct_size
    for f, _t in abi.Options._fields_:
        if f != "reserved":
            assert getattr(o, f) == getattr(d, f), f
    # the reference's compile-time switches, apps/src/pathtrace.cu:36-42,279-280
    assert (o.antialiasing, o.depth_of_field, o.sort_by_material) == (1, 0, 1)
    assert abs(o.lens_radius - 0.8) < 1e-7 and o.focal_distance == 11.0


def test_struct_layouts_match_reference_sizes():
    # Material 44 B, Camera 84 B
    assert C.sizeof(abi.Material) == 44
    assert C.sizeof(abi.Camera) == 84
    assert C.sizeof(abi.Geom) == 8 + 3 * 64 + 6 * 4


def test_errors_are_returned_not_fatal():
    lib = api.load_library()
    h = C.c_void_p()
    rc = lib.b2pt_scene_load(b"/nonexistent/scene.txt", None, C.byref(h))
    assert rc == -4 and b"cannot open" in lib.b2pt_last_error()
    assert lib.b2pt_render(None, 1, 1, 1) == -1
    lib.b2pt_destroy(None)  # pathtraceFree before the first Init is legal (main.cpp:245-248)
