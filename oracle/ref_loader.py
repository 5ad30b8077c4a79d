"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference under stub Blender modules.

Runs the files under ``/root/reference`` (this container) or their staged byte-for-byte copies under
``oracle/_ref`` (the GPU box; see ``oracle/stage_ref.py``).  Used by
``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by the CPU tests that pin
``oracle/icp_oracle.py`` against the live reference when it is present.

Nothing from the reference enters the repo's history: its modules are executed from where they lie.
The stubs exist because ``functions/general.py:26`` star-imports ``functions/common``,
whose modules import ``bpy``/``bmesh``/``mathutils`` and read ``bpy.app.version`` at import
time (functions/common/blender.py:24-30, wrappers.py:45), and ``operators/icp_align.py``
imports ``bpy``, ``mathutils.Matrix`` and ``mathutils.bvhtree.BVHTree`` (:24-27).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

from . import mathutils_shim as ms


def _find_root():
    """ICP_REFERENCE_ROOT, else the read-only checkout, else the staged copy (oracle/stage_ref.py)."""
    env = os.environ.get("ICP_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")):
        if os.path.isfile(os.path.join(cand, "functions", "general.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_PKG = "refpkg_object_alignment"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "functions", "general.py"))


class _Anything:
    """Permissive stand-in for bpy.types classes / bpy.props factories."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, name):
        return _Anything()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _install_stubs():
    if "bpy" in sys.modules and getattr(sys.modules["bpy"], "_icp_stub", False):
        return
    bpy = _module("bpy", _icp_stub=True)
    bpy.app = types.SimpleNamespace(version=(3, 2, 2), version_string="3.2.2")
    bpy.context = types.SimpleNamespace(scene=types.SimpleNamespace(objects=[]))
    bpy.data = types.SimpleNamespace()
    bpy.ops = _Anything()
    bpy.utils = _Anything()
    bpy_types = _module("bpy.types")
    for cls in ("Object", "Scene", "ViewLayer", "Operator", "Panel", "AddonPreferences",
                "CollectionProperty", "Mesh", "Material", "Image", "Collection"):
        setattr(bpy_types, cls, type(cls, (), {}))
    bpy.types = bpy_types
    bpy_props = _module("bpy.props")
    for p in ("IntProperty", "BoolProperty", "FloatProperty", "EnumProperty", "StringProperty",
              "FloatVectorProperty", "PointerProperty", "CollectionProperty",
              "IntVectorProperty", "BoolVectorProperty"):
        setattr(bpy_props, p, lambda *a, **k: None)
    bpy.props = bpy_props
    mathutils = _module("mathutils", Vector=ms.Vector, Matrix=ms.Matrix,
                        Euler=type("Euler", (), {}), Color=type("Color", (), {}),
                        Quaternion=ms.Quaternion)
    bvhtree = _module("mathutils.bvhtree")
    from .icp_oracle import NearestVertexBVH

    bvhtree.BVHTree = NearestVertexBVH
    mathutils.bvhtree = bvhtree
    bpy_extras = _module("bpy_extras")
    view3d_utils = _module("bpy_extras.view3d_utils")
    for fn in ("location_3d_to_region_2d", "region_2d_to_vector_3d", "region_2d_to_origin_3d"):
        setattr(view3d_utils, fn, lambda *a, **k: None)
    bpy_extras.view3d_utils = view3d_utils
    gpu_extras = _module("gpu_extras")
    gpu_batch = _module("gpu_extras.batch", batch_for_shader=lambda *a, **k: None)   # align_pick_points.py:29
    gpu_extras.batch = gpu_batch
    mods = {
        "bpy": bpy, "bpy.types": bpy_types, "bpy.props": bpy_props,
        "bmesh": _module("bmesh"), "bgl": _module("bgl"), "blf": _module("blf"),
        "gpu": _module("gpu"), "gpu_extras": gpu_extras, "gpu_extras.batch": gpu_batch,
        "mathutils": mathutils, "mathutils.bvhtree": bvhtree,
        "bpy_extras": bpy_extras, "bpy_extras.view3d_utils": view3d_utils,
    }
    for k, v in mods.items():
        sys.modules.setdefault(k, v)


def _synthetic_package(name, path):
    """A package object whose __init__.py is NOT executed (so no registration side effects)."""
    if name in sys.modules:
        return sys.modules[name]
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    pkg.__package__ = name
    sys.modules[name] = pkg
    return pkg


def load_pick_points():
    """The reference's picked-points operator module (operators/align_pick_points.py), for its
    ``align_obj`` (:415-465) -- the interactive parts are never called."""
    load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(_PKG + ".operators.align_pick_points")


def load_reference():
    """Returns (general_module, icp_align_module) executed from /root/reference, unmodified."""
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    _install_stubs()
    _synthetic_package(_PKG, REFERENCE_ROOT)
    _synthetic_package(_PKG + ".operators", os.path.join(REFERENCE_ROOT, "operators"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # general.py:106 has an invalid escape in a docstring
        general = importlib.import_module(_PKG + ".functions.general")
        icp_align = importlib.import_module(_PKG + ".operators.icp_align")
    return general, icp_align
