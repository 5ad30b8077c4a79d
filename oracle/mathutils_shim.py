"""TEST INFRASTRUCTURE ONLY -- fp32 stand-ins for Blender's C ``mathutils`` types.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything under ``oracle/``.  The product
package (``object_alignment_b200``) never does.

Blender's ``mathutils`` (pinned by the reference only through
``bl_info["blender"] = (3, 2, 2)``, /root/reference/__init__.py:22) is a C module
that is absent from /root/reference and from this image.  The reference's hot
path touches exactly these pieces of it:

* ``Matrix.inverted()``, ``Matrix @ Vector``, ``Matrix @ Matrix``
  (/root/reference/functions/general.py:262-266,287,299,304;
  /root/reference/operators/icp_align.py:116-121)
* ``Vector - Vector``, ``Vector.length``, ``Vector[i]`` (general.py:299,320-321)
* ``Matrix.Identity(4)``, ``m[y][z] = x``, ``to_translation()``,
  ``to_quaternion().angle`` (icp_align.py:116-130)

The arithmetic restated here is the published behaviour of Blender 3.2's
``source/blender/python/mathutils`` and ``blenlib/intern/math_matrix.c`` as far
as it is known without the source at hand (it CANNOT be verified in this image;
SURVEY.md section 8c lists these as assumptions):

S1  ``Matrix @ Vector``: the 3-vector is extended with w = 1.0f; every output
    row is ``(float) sum_c (double)((float)(m[r][c] * v[c]))`` -- float products,
    double accumulation starting from 0.0, one rounding to float on store.
S2  ``Matrix.inverted()``: float adjugate / float determinant (Graphics-Gems
    cofactor expansion over Blender's column-major ``float[4][4]``), every
    element divided by the determinant, all in fp32; ValueError when det == 0.
S3  ``Matrix @ Matrix``: same float-product / double-accumulate / float-store rule.
S5  ``Vector.length``: ``sqrt`` of a double sum of float squares taken z, y, x.

The CUDA engine implements the same S1..S5 so that q, the world distance and the
targets agree bit for bit with this shim; differences between this shim and real
Blender can only move results by fp32 ulps (index parity is then exempt for
degenerate queries, rotation/translation parity is 1e-5 -- BASELINE.json).
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------- helpers
def _f(x):
    return f32(x)


def matvec43(m: np.ndarray, v) -> np.ndarray:
    """S1 on raw arrays: m is (4,4) float32, v is 3 floats; returns (3,) float32.

    Vectorised over the three rows, but the order of operations per row is exactly
    ((p0 + p1) + p2) + p3 with float products and double accumulation."""
    v4 = np.empty(4, dtype=f32)
    v4[0], v4[1], v4[2], v4[3] = v[0], v[1], v[2], 1.0
    p = (m[:3] * v4).astype(f64)          # float32 products, then widened (exact)
    return (((p[:, 0] + p[:, 1]) + p[:, 2]) + p[:, 3]).astype(f32)


def matmul44(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """S3 on raw arrays: (4,4) float32 @ (4,4) float32 -> (4,4) float32."""
    out = np.empty((4, 4), dtype=f32)
    for r in range(4):
        for c in range(4):
            acc = f64(0.0)
            for k in range(4):
                acc = acc + f64(f32(a[r, k] * b[k, c]))
            out[r, c] = f32(acc)
    return out


def _det2(a, b, c, d):
    return f32(f32(a * d) - f32(b * c))


def _det3(a1, a2, a3, b1, b2, b3, c1, c2, c3):
    t0 = f32(a1 * _det2(b2, b3, c2, c3))
    t1 = f32(b1 * _det2(a2, a3, c2, c3))
    t2 = f32(c1 * _det2(a2, a3, b2, b3))
    return f32(f32(t0 - t1) + t2)


def invert44(m_rowmajor: np.ndarray) -> np.ndarray:
    """S2 on raw arrays: fp32 adjugate/determinant inverse; raises ValueError if singular."""
    # Blender keeps matrices column-major: m[i][j] is (row j, column i).
    m = np.ascontiguousarray(m_rowmajor, dtype=f32).T
    a1, b1, c1, d1 = m[0]
    a2, b2, c2, d2 = m[1]
    a3, b3, c3, d3 = m[2]
    a4, b4, c4, d4 = m[3]
    det = f32(
        f32(
            f32(
                f32(a1 * _det3(b2, b3, b4, c2, c3, c4, d2, d3, d4))
                - f32(b1 * _det3(a2, a3, a4, c2, c3, c4, d2, d3, d4))
            )
            + f32(c1 * _det3(a2, a3, a4, b2, b3, b4, d2, d3, d4))
        )
        - f32(d1 * _det3(a2, a3, a4, b2, b3, b4, c2, c3, c4))
    )
    if det == f32(0.0):
        raise ValueError("Matrix.invert(ed): matrix does not have an inverse")
    r = np.empty((4, 4), dtype=f32)
    r[0, 0] = _det3(b2, b3, b4, c2, c3, c4, d2, d3, d4)
    r[1, 0] = -_det3(a2, a3, a4, c2, c3, c4, d2, d3, d4)
    r[2, 0] = _det3(a2, a3, a4, b2, b3, b4, d2, d3, d4)
    r[3, 0] = -_det3(a2, a3, a4, b2, b3, b4, c2, c3, c4)
    r[0, 1] = -_det3(b1, b3, b4, c1, c3, c4, d1, d3, d4)
    r[1, 1] = _det3(a1, a3, a4, c1, c3, c4, d1, d3, d4)
    r[2, 1] = -_det3(a1, a3, a4, b1, b3, b4, d1, d3, d4)
    r[3, 1] = _det3(a1, a3, a4, b1, b3, b4, c1, c3, c4)
    r[0, 2] = _det3(b1, b2, b4, c1, c2, c4, d1, d2, d4)
    r[1, 2] = -_det3(a1, a2, a4, c1, c2, c4, d1, d2, d4)
    r[2, 2] = _det3(a1, a2, a4, b1, b2, b4, d1, d2, d4)
    r[3, 2] = -_det3(a1, a2, a4, b1, b2, b4, c1, c2, c4)
    r[0, 3] = -_det3(b1, b2, b3, c1, c2, c3, d1, d2, d3)
    r[1, 3] = _det3(a1, a2, a3, c1, c2, c3, d1, d2, d3)
    r[2, 3] = -_det3(a1, a2, a3, b1, b2, b3, d1, d2, d3)
    r[3, 3] = _det3(a1, a2, a3, b1, b2, b3, c1, c2, c3)
    with np.errstate(all="ignore"):
        r = (r / det).astype(f32)
    return np.ascontiguousarray(r.T)


def length3(v) -> float:
    """S5: sqrt of the double sum of float squares, accumulated z, y, x."""
    v = np.asarray(v, dtype=f32)
    p = (v * v).astype(f64)
    return math.sqrt(float((p[2] + p[1]) + p[0]))


# --------------------------------------------------------------------------- types
class Vector:
    """3-component fp32 vector with the few operations the reference uses."""

    __slots__ = ("v",)

    def __init__(self, seq=(0.0, 0.0, 0.0)):
        self.v = np.array([seq[0], seq[1], seq[2]], dtype=f32)

    def __getitem__(self, i):
        return float(self.v[i])

    def __setitem__(self, i, x):
        self.v[i] = x

    def __len__(self):
        return 3

    def __iter__(self):
        return (float(x) for x in self.v)

    def __sub__(self, other):
        return Vector(self.v - other.v)

    def __add__(self, other):
        return Vector(self.v + other.v)

    @property
    def length(self) -> float:
        return length3(self.v)

    def copy(self):
        return Vector(self.v)

    def __repr__(self):
        return "Vector((%r, %r, %r))" % tuple(float(x) for x in self.v)


class _Row:
    """Row proxy so that ``m[y][z] = x`` (icp_align.py:119) writes through as float."""

    __slots__ = ("a", "r")

    def __init__(self, a, r):
        self.a = a
        self.r = r

    def __getitem__(self, c):
        return float(self.a[self.r, c])

    def __setitem__(self, c, x):
        self.a[self.r, c] = f32(x)

    def __len__(self):
        return self.a.shape[1]

    def __iter__(self):
        return (float(x) for x in self.a[self.r])


class Quaternion:
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, w, x, y, z):
        self.w, self.x, self.y, self.z = w, x, y, z

    @property
    def angle(self) -> float:
        return 2.0 * math.acos(max(-1.0, min(1.0, self.w)))


class Matrix:
    """4x4 fp32 matrix, row access like mathutils (``m[row][col]``)."""

    __slots__ = ("a",)

    def __init__(self, rows=None):
        if rows is None:
            self.a = np.identity(4, dtype=f32)
        else:
            self.a = np.array(rows, dtype=f32).reshape(4, 4).copy()

    @classmethod
    def Identity(cls, n):
        assert n == 4
        return cls()

    def __getitem__(self, r):
        return _Row(self.a, r)

    def copy(self):
        return Matrix(self.a)

    def inverted(self):
        return Matrix(invert44(self.a))

    def __matmul__(self, other):
        if isinstance(other, Matrix):
            return Matrix(matmul44(self.a, other.a))
        if isinstance(other, Vector):
            return Vector(matvec43(self.a, other.v))
        return NotImplemented

    def to_translation(self):
        return Vector(self.a[:3, 3])

    def to_quaternion(self):
        # Only ``abs(quat.angle)`` is consumed, for reporting (icp_align.py:139).
        r = self.a[:3, :3].astype(f64)
        # normalise columns (scale may be != 1 in ROT_LOC_SCALE mode)
        n = np.linalg.norm(r, axis=0)
        n[n == 0] = 1.0
        r = r / n
        tr = float(r[0, 0] + r[1, 1] + r[2, 2])
        w = math.sqrt(max(0.0, 1.0 + tr)) / 2.0
        return Quaternion(w, 0.0, 0.0, 0.0)

    def __repr__(self):
        return "Matrix(%r)" % (self.a.tolist(),)


# --------------------------------------------------------------------------- duck-typed scene
class VertexGroupElement:
    __slots__ = ("group", "weight")

    def __init__(self, group, weight):
        self.group = group
        self.weight = weight


class MeshVertex:
    __slots__ = ("co", "index", "groups")

    def __init__(self, co, index, groups=()):
        self.co = co
        self.index = index
        self.groups = groups


class _VertexSeq:
    """``mesh.vertices``: indexable/iterable of MeshVertex + Blender's bulk ``foreach_get``."""

    def __init__(self, coords_f32, group_members=None):
        self._co = np.ascontiguousarray(coords_f32, dtype=f32)
        self._members = group_members or {}
        self._cache = {}

    def __len__(self):
        return self._co.shape[0]

    def __getitem__(self, i):
        i = int(i)
        v = self._cache.get(i)
        if v is None:
            groups = tuple(
                VertexGroupElement(gi, float(w[i]))
                for gi, (mask, w) in self._members.items()
                if mask[i]
            )
            v = MeshVertex(Vector(self._co[i]), i, groups)
            if len(self._cache) < 4_000_000:
                self._cache[i] = v
        return v

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def foreach_get(self, attr, buf):
        assert attr == "co"
        np.asarray(buf).reshape(-1)[:] = self._co.reshape(-1)


class _LoopTriangles:
    """``mesh.loop_triangles`` with Blender's bulk ``foreach_get("vertices", buf)``."""

    def __init__(self, tris):
        self._t = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)

    def __len__(self):
        return self._t.shape[0]

    def foreach_get(self, attr, buf):
        assert attr == "vertices"
        np.asarray(buf).reshape(-1)[:] = self._t.reshape(-1)


class Mesh:
    def __init__(self, coords_f32, group_members=None, faces=None):
        self.vertices = _VertexSeq(coords_f32, group_members)
        self.loop_triangles = _LoopTriangles(faces if faces is not None else np.zeros((0, 3), np.int32))

    def calc_loop_triangles(self):
        pass


class VertexGroup:
    def __init__(self, name, index):
        self.name = name
        self.index = index


class _VertexGroups:
    def __init__(self, groups):
        self._g = list(groups)

    def __iter__(self):
        return iter(self._g)

    def __contains__(self, name):
        return any(g.name == name for g in self._g)

    def __len__(self):
        return len(self._g)

    def __getitem__(self, name):
        for g in self._g:
            if g.name == name:
                return g
        raise KeyError(name)


class Object:
    """Duck-typed ``bpy.types.Object`` exposing what icp_align.py:51-80,121-132 touches.

    ``groups`` maps group name -> (member_mask bool[M], weight float32[M]).
    """

    def __init__(self, name, coords, matrix_world=None, groups=None, faces=None):
        self.name = name
        self.type = "MESH"
        self.rotation_mode = "XYZ"
        self.matrix_world = Matrix(matrix_world) if matrix_world is not None else Matrix()
        groups = groups or {}
        vgs, members = [], {}
        for gi, (gname, (mask, w)) in enumerate(groups.items()):
            vgs.append(VertexGroup(gname, gi))
            members[gi] = (np.asarray(mask, dtype=bool), np.asarray(w, dtype=f32))
        self.vertex_groups = _VertexGroups(vgs)
        self.data = Mesh(coords, members, faces)
        self.tag_count = 0

    def update_tag(self):
        self.tag_count += 1

    def __eq__(self, other):
        return self is other

    def __ne__(self, other):
        return self is not other

    def __hash__(self):
        return id(self)


class _ViewLayer:
    def __init__(self):
        self.updates = 0

    def update(self):
        self.updates += 1


class _WindowManager:
    def __init__(self):
        self.timers, self.handlers = [], []

    def event_timer_add(self, time_step=0.01, window=None):
        self.timers.append(time_step)
        return len(self.timers)

    def event_timer_remove(self, timer):
        self.timers.remove(self.timers[timer - 1]) if self.timers else None

    def modal_handler_add(self, op):
        self.handlers.append(op)


class _Area:
    def __init__(self):
        self.redraws = 0

    def tag_redraw(self):
        self.redraws += 1


class Event:
    def __init__(self, type_):
        self.type = type_


class Context:
    """Duck-typed ``bpy.context`` (icp_align.py:43-53,133; icp_align_feedback.py:47-49,110)."""

    def __init__(self, active, selected):
        self.object = active
        self.selected_objects = list(selected)
        self.view_layer = _ViewLayer()
        self.window_manager = _WindowManager()
        self.window = None
        self.area = _Area()

    def evaluated_depsgraph_get(self):
        return None


class Settings:
    """Stand-in for ObjectAlignmentPreferences (/root/reference/lib/preferences.py:31-72)."""

    def __init__(self, **kw):
        self.icp_iterations = 50
        self.redraw_frequency = 10
        self.use_sample = False
        self.sample_fraction = 0.5
        self.min_start = 0.5
        self.target_d = 0.01
        self.use_target = True
        self.take_m_with = False
        self.align_meth = "0"
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)
