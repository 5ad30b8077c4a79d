"""Synthetic vertex sets for the BASELINE.json configs.

No Blender and no mesh files exist in this image, so "identical meshes" means these
seeded, procedurally generated fp32 vertex arrays.  Pure numpy; not part of the hot path.

C1  geodesic icosphere, frequency 10 -> 1002 vertices (10 f^2 + 2)
C2+ "bunny-like" lobed surface r(u) = 1 + 0.25 sin(3 theta) cos(2 phi)
        + 0.15 exp(-8 |u - e1|^2) + 0.15 exp(-8 |u - e2|^2),  u uniform on the sphere
"""
from __future__ import annotations

import math

import numpy as np

_AXIS = np.array([1.0, 2.0, 3.0]) / math.sqrt(14.0)
_E1 = np.array([0.45, 0.35, 0.82]) / np.linalg.norm([0.45, 0.35, 0.82])
_E2 = np.array([-0.45, 0.35, 0.82]) / np.linalg.norm([-0.45, 0.35, 0.82])


def rotation_about_axis(angle_deg: float, axis=_AXIS) -> np.ndarray:
    """3x3 rotation (float64) by angle_deg about a unit axis (Rodrigues)."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    t = math.radians(angle_deg)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.identity(3) + math.sin(t) * k + (1 - math.cos(t)) * (k @ k)


def rigid_matrix(angle_deg: float, translation, axis=_AXIS) -> np.ndarray:
    """4x4 float32 world matrix: rotate about `axis`, then translate."""
    m = np.identity(4)
    m[:3, :3] = rotation_about_axis(angle_deg, axis)
    m[:3, 3] = translation
    return m.astype(np.float32)


def icosphere(frequency: int = 10, return_faces: bool = False):
    """Class-I geodesic icosahedron, unit radius: (10 f^2 + 2, 3) float32, deterministic order.
    With return_faces also the (20 f^2, 3) int32 triangle list."""
    f = int(frequency)
    phi = (1 + math.sqrt(5)) / 2
    v = np.array([[-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0],
                  [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
                  [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v[0])
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
             (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
             (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
             (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    seen = {}
    out = []
    tris = []
    for (a, b, c) in faces:
        local = {}
        for i in range(f + 1):
            for j in range(f + 1 - i):
                k = f - i - j
                # integer barycentric key on the (sorted) corner ids makes shared points unique
                key = tuple(sorted(((a, i), (b, j), (c, k))))
                key = tuple((cid, w) for cid, w in key if w != 0)
                if key not in seen:
                    seen[key] = len(out)
                    p = (i * v[a] + j * v[b] + k * v[c]) / f
                    out.append(p / np.linalg.norm(p))
                local[(i, j)] = seen[key]
        for i in range(f):
            for j in range(f - i):
                tris.append((local[(i, j)], local[(i + 1, j)], local[(i, j + 1)]))
                if i + j < f - 1:
                    tris.append((local[(i + 1, j)], local[(i + 1, j + 1)], local[(i, j + 1)]))
    pts = np.array(out, dtype=np.float64)
    assert pts.shape[0] == 10 * f * f + 2, pts.shape
    if return_faces:
        t = np.array(tris, dtype=np.int32)
        assert t.shape[0] == 20 * f * f, t.shape
        return pts.astype(np.float32), t
    return pts.astype(np.float32)


def bunny_like(n: int, seed: int, noise: float = 0.0) -> np.ndarray:
    """n samples of the lobed test surface (float32), optional isotropic Gaussian noise."""
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    theta = np.arccos(np.clip(u[:, 2], -1.0, 1.0))
    ph = np.arctan2(u[:, 1], u[:, 0])
    r = (1.0 + 0.25 * np.sin(3 * theta) * np.cos(2 * ph)
         + 0.15 * np.exp(-8 * ((u - _E1) ** 2).sum(axis=1))
         + 0.15 * np.exp(-8 * ((u - _E2) ** 2).sum(axis=1)))
    p = u * r[:, None]
    if noise > 0:
        p = p + rng.normal(0.0, noise, size=p.shape)
    return p.astype(np.float32)


def bunny_like_mesh(n: int, seed: int):
    """The lobed test surface as a watertight triangle mesh: n vertices (float32) and the
    (2n - 4, 3) int32 triangles of the convex hull of their unit directions (a Delaunay
    triangulation of the sphere, carried to the star-shaped surface)."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    verts = bunny_like(n, seed)
    tris = ConvexHull(u).simplices.astype(np.int32)
    return verts, tris


def triangle_corners(verts: np.ndarray, tris: np.ndarray) -> np.ndarray:
    """(F, 9) float32 corner coordinates a.xyz, b.xyz, c.xyz (the icp_set_base_triangles layout)."""
    return np.ascontiguousarray(np.asarray(verts, np.float32)[np.asarray(tris)].reshape(-1, 9))


def apply_rigid(points: np.ndarray, angle_deg: float, translation, axis=_AXIS) -> np.ndarray:
    """Rotate then translate a point set (float64 arithmetic, float32 result)."""
    r = rotation_about_axis(angle_deg, axis)
    return (points.astype(np.float64) @ r.T + np.asarray(translation, dtype=np.float64)).astype(np.float32)


def config_c1():
    """C1: two 1002-vertex icospheres; moving = permuted (seed 1), rotated 10 deg, shifted."""
    base = icosphere(10)
    rng = np.random.default_rng(1)
    perm = rng.permutation(base.shape[0])
    moving = apply_rigid(base[perm], 10.0, (0.05, -0.03, 0.02))
    return dict(base=base, moving=moving, thresh=0.5, iters=20)


def config_cloud(n_moving: int, n_base: int, seed_base: int, seed_moving: int,
                 angle_deg: float = 5.0, translation=(0.02, -0.01, 0.03), noise: float = 1e-3):
    """C2/C3/C4 family: base and moving are independent samples of the same surface."""
    base = bunny_like(n_base, seed_base)
    moving = apply_rigid(bunny_like(n_moving, seed_moving, noise=noise), angle_deg, translation)
    return dict(base=base, moving=moving, thresh=0.5, iters=50)


def config_c5(n: int = 5_000_000, mask: str = "include"):
    """C5: 25 % outliers in the 2x bounding box + an include or exclude vertex-group mask."""
    cfg = config_cloud(n, n, 500, 501)
    mv = cfg["moving"]
    rng = np.random.default_rng(502)
    n_out = n // 4
    sel = rng.choice(n, size=n_out, replace=False)
    lo, hi = mv.min(axis=0), mv.max(axis=0)
    c, h = (lo + hi) / 2, (hi - lo)
    mv[sel] = (c + (rng.random((n_out, 3)) - 0.5) * 2 * h).astype(np.float32)
    rng = np.random.default_rng(503)
    if mask == "include":
        member = rng.random(n) < 0.70
        w = np.where(rng.random(n) < 0.85, 1.0, rng.random(n) * 0.9).astype(np.float32)
        groups = {"icp_include": (member, w)}
    else:
        member = rng.random(n) < 0.40
        w = np.where(rng.random(n) < 0.25, rng.random(n) * 0.1, 1.0).astype(np.float32)
        groups = {"icp_exclude": (member, w)}
    cfg.update(groups=groups, reject_fraction=0.25, thresh=float("inf"))
    return cfg
