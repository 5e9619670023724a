"""The reference's five scenes, written out in its own text format.

``apps/scenes/{cornell,cornellGlass,cornellObj,cornellSpaceship,sphere}.txt``
are tiny data files: a list of materials, one camera and a list of objects.
They are described here as Python data and serialised by :func:`write_scene`
in the line format ``apps/src/scene.cpp`` parses (MATERIAL blocks of seven
property lines, a CAMERA block of five lines plus EYE/LOOKAT/UP, OBJECT blocks
ending at a blank line).  ``tests/test_loader.py`` checks that loading a file
written here yields the same POD scene, bit for bit, as the reference's loader
produces from the reference's own file (golden ``.b2s`` fixtures).

The resolution, iteration count and depth of every shipped scene are
800x800 / 5000 / 8; BASELINE.json's configurations override them.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

from . import standin_mesh

# (RGB, SPECEX, SPECRGB, REFL, REFR, REFRIOR, EMITTANCE)
_LIGHT = ((1, 1, 1), 0, (0, 0, 0), 0, 0, 0, 5)
_WHITE = ((.98, .98, .98), 0, (0, 0, 0), 0, 0, 0, 0)
_RED = ((.85, .35, .35), 0, (0, 0, 0), 0, 0, 0, 0)
_GREEN = ((.35, .85, .35), 0, (0, 0, 0), 0, 0, 0, 0)
_MIRROR = ((.98, .98, .98), 0, (.98, .98, .98), 1, 0, 0, 0)
_GLASS = ((.98, .98, .98), 0, (.85, .85, .98), 0, 1, 1.65, 0)

# (type, material, TRANS, ROTAT, SCALE)
_BOX = [
    ("cube", 0, (0, 10, 0), (0, 0, 0), (3, .3, 3)),       # ceiling light
    ("cube", 1, (0, 0, 0), (0, 0, 0), (10, .01, 10)),     # floor
    ("cube", 1, (0, 10, 0), (0, 0, 90), (.01, 10, 10)),   # ceiling
    ("cube", 1, (0, 5, -5), (0, 90, 0), (.01, 10, 10)),   # back wall
    ("cube", 2, (-5, 5, 0), (0, 0, 0), (.01, 10, 10)),    # left wall
    ("cube", 3, (5, 5, 0), (0, 0, 0), (.01, 10, 10)),     # right wall
]
_SHIP = ("obj", None, (1, 3, 3), (0, 20, 180), (1, 1, 1))
_CAMERA = dict(res=(800, 800), fovy=45, iterations=5000, depth=8, eye=(0.0, 5, 10.5), lookat=(0, 5, 0), up=(0, 1, 0))

SCENES: Dict[str, dict] = {
    "cornell": dict(file="cornell", materials=[_LIGHT, _WHITE, _RED, _GREEN, _MIRROR],
                    objects=_BOX + [("sphere", 1, (-1, 4, -1), (0, 0, 0), (3, 3, 3))]),
    "cornellGlass": dict(file="cornell", materials=[_LIGHT, _WHITE, _RED, _GREEN, _MIRROR, _GLASS],
                         objects=_BOX + [("sphere", 5, (-1, 4, -1), (0, 0, 0), (3, 3, 3))]),
    "cornellObj": dict(file="cornell", materials=[_LIGHT, _WHITE, _RED, _GREEN, _MIRROR, _GLASS], objects=_BOX + [_SHIP]),
    "cornellSpaceship": dict(file="cornell", materials=[_LIGHT, _WHITE, _RED, _GREEN, _MIRROR, _GLASS],
                             objects=_BOX + [("sphere", 1, (-2, 7, -1), (0, 0, 0), (2, 2, 2)),
                                             ("sphere", 5, (1, 6, 0), (0, 0, 0), (2, 2, 2)), _SHIP]),
    "sphere": dict(file="sphere", materials=[_LIGHT], objects=[("sphere", 0, (0, 0, 0), (0, 0, 0), (3, 3, 3))]),
    # NOT one of the reference's scenes: two OBJ geoms in one scene (the loader appends one material per OBJ,
    # scene.cpp:221-231), the second one scaled by 1.5, so that its object-space distances (world / 1.5) compete with
    # world-space ones exactly as the reference's mixed `min` does.  Used by the parity tests of the several-meshes path.
    "twoShips": dict(file="cornell", materials=[_LIGHT, _WHITE, _RED, _GREEN, _MIRROR, _GLASS],
                     objects=_BOX + [("sphere", 5, (1, 6, 0), (0, 0, 0), (2, 2, 2)), _SHIP,
                                     ("obj", None, (-2, 6.5, -2), (30, 0, 10), (1.5, 1.5, 1.5))]),
}


def _num(x) -> str:
    return repr(float(x)) if isinstance(x, float) else str(x)


def scene_text(name: str, width: Optional[int] = None, height: Optional[int] = None, iterations: Optional[int] = None,
               depth: Optional[int] = None, obj_path: str = "../models/" + standin_mesh.OBJ_NAME) -> str:
    sc = SCENES[name]
    cam = dict(_CAMERA)
    out: List[str] = []
    for i, (rgb, specex, specrgb, refl, refr, ior, emit) in enumerate(sc["materials"]):
        out += [f"MATERIAL {i}", "RGB         " + " ".join(map(_num, rgb)), f"SPECEX      {_num(specex)}",
                "SPECRGB     " + " ".join(map(_num, specrgb)), f"REFL        {_num(refl)}", f"REFR        {_num(refr)}",
                f"REFRIOR     {_num(ior)}", f"EMITTANCE   {_num(emit)}", ""]
    w, h = (width, height) if width and height else cam["res"]
    out += ["CAMERA", f"RES         {w} {h}", f"FOVY        {_num(cam['fovy'])}",
            f"ITERATIONS  {iterations or cam['iterations']}", f"DEPTH       {depth if depth is not None else cam['depth']}",
            f"FILE        {sc['file']}", "EYE         " + " ".join(map(_num, cam["eye"])),
            "LOOKAT      " + " ".join(map(_num, cam["lookat"])), "UP          " + " ".join(map(_num, cam["up"])), "", ""]
    for i, (kind, mat, tr, ro, scl) in enumerate(sc["objects"]):
        out += [f"OBJECT {i}", kind]
        out += [obj_path] if kind == "obj" else [f"material {mat}"]
        out += ["TRANS       " + " ".join(map(_num, tr)), "ROTAT       " + " ".join(map(_num, ro)),
                "SCALE       " + " ".join(map(_num, scl)), ""]
    return "\n".join(out)


def write_scene(name: str, path: str, **kw) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(scene_text(name, **kw))
    return path


def uses_mesh(name: str) -> bool:
    return any(o[0] == "obj" for o in SCENES[name]["objects"])
