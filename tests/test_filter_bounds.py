"""CPU property test of the brute-force kernels' lower-bound filters (no GPU, no oracle needed).

The slab / cylinder / 3-axis tests of `csrc/nn_bruteforce_slab.cuh` and `nn_bruteforce_cyl.cuh` may
skip a base vertex only if it can neither win nor tie.  Restated in numpy with the kernels' exact
fp32 operation order (fma emulated through float64, whose product of two float32 is exact), the
property checked here is the one the exactness argument rests on ("Error budget"):

    for every query q and EVERY vertex j of a tile, with  best = S4(q, b_j)  (j ties the running best):
        F1_a(q, j) <= thr1_a,   F1_b(q, j) <= thr1_b      (slab test along either axis)
        F(q, j)    <= thr_f                               (cylinder test)
        E3(q, j)   <= thr_3                               (3-axis expansion test)

so a vertex whose exact-metric distance equals (or beats) the current best always reaches the exact
S4 evaluation.  Data: tiles at unit scale, 800 units from the origin, at 1e-4 and 1e+3 scale, flat and
needle-shaped tiles, queries inside, near and far from the tile.
"""
import zlib

import numpy as np
import pytest

F = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def s4_d2(q, b):
    dx, dy, dz = (q[..., 0] - b[..., 0]).astype(F), (q[..., 1] - b[..., 1]).astype(F), (q[..., 2] - b[..., 2]).astype(F)
    return fma(dz, dz, fma(dy, dy, (dx * dx).astype(F)))


def tile_planes(pts):
    """k_build_slab_tiles: centre, centred planes, w2, per-axis squares, b2 (axes a, b, c = x, y, z)."""
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    cen = (F(0.5) * lo + F(0.5) * hi).astype(F)
    p = (pts - cen).astype(F)
    wa, wb = (p[:, 0] * p[:, 0]).astype(F), (p[:, 1] * p[:, 1]).astype(F)
    w2 = fma(p[:, 1], p[:, 1], wa)
    b2 = F(fma(p[:, 2], p[:, 2], w2).max() * F(1.00001))
    return cen, p, wa, wb, w2, b2


def filter_margins(q, pts):
    """Smallest slack (threshold - bound) over all (query, vertex) pairs, per test, with best = S4(q, b_j)."""
    cen, p, wa, wb, w2, b2 = tile_planes(pts)
    Q = q[:, None, :]
    best = s4_d2(np.broadcast_to(Q, (q.shape[0], pts.shape[0], 3)), np.broadcast_to(pts[None], (q.shape[0], pts.shape[0], 3)))
    pq = (q - cen).astype(F)                                    # centred query
    pa, pb, pc = pq[:, 0:1], pq[:, 1:2], pq[:, 2:3]
    na, nb = (pa * pa).astype(F), (pb * pb).astype(F)
    # slab_setup
    nq3_s = fma(pc, pc, (na + nb).astype(F))
    delta_s = (F(4.0e-6) * (nq3_s + b2).astype(F)).astype(F)
    thr1_a = ((best + delta_s).astype(F) - na).astype(F)
    thr1_b = ((best + delta_s).astype(F) - nb).astype(F)
    f1_a = fma((F(-2.0) * pa).astype(F), p[None, :, 0], wa[None])
    f1_b = fma((F(-2.0) * pb).astype(F), p[None, :, 1], wb[None])
    # cyl_setup / cyl_recheck
    nq2 = fma(pb, pb, (pa * pa).astype(F))
    nq3 = fma(pc, pc, nq2)
    delta = (F(4.0e-6) * (nq3 + b2).astype(F)).astype(F)
    ax, ay, az = (F(-2.0) * pa).astype(F), (F(-2.0) * pb).astype(F), (F(-2.0) * pc).astype(F)
    thr_f = ((best + delta).astype(F) - nq2).astype(F)
    thr_3 = (thr_f - (pc * pc).astype(F)).astype(F)
    f = fma(ay, p[None, :, 1], fma(ax, p[None, :, 0], w2[None]))
    zc = p[None, :, 2]
    e3 = fma(zc, (zc + az).astype(F), f)
    scale = (nq3 + b2).astype(np.float64)                       # s^2-like normaliser of the slack
    rel = lambda thr, val: float(((thr.astype(np.float64) - val.astype(np.float64)) / scale).min())
    return dict(slab_a=rel(thr1_a, f1_a), slab_b=rel(thr1_b, f1_b), cyl=rel(thr_f, f), e3=rel(thr_3, e3))


def _tile(rng, kind, scale, offset):
    n = 1024
    if kind == "blob":
        pts = rng.standard_normal((n, 3))
    elif kind == "flat":                      # a surface patch: thin along z
        pts = rng.standard_normal((n, 3)) * np.array([1.0, 1.0, 0.01])
    elif kind == "needle":                    # long along x
        pts = rng.standard_normal((n, 3)) * np.array([5.0, 0.02, 0.02])
    else:                                     # lattice: many exactly tied distances
        g = np.arange(10.0)
        pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
        pts = np.concatenate([pts, pts[:24]])
    return (pts * 0.07 * scale + offset).astype(F)


@pytest.mark.parametrize("kind", ["blob", "flat", "needle", "lattice"])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1.0, 800.0), (1e-4, 0.0), (1e-4, 0.3), (1e3, -5e4)])
def test_no_vertex_that_ties_the_best_is_filtered(kind, scale, offset):
    rng = np.random.default_rng(zlib.crc32(repr((kind, scale, offset)).encode()))
    pts = _tile(rng, kind, scale, np.array([offset, -0.6 * offset, 0.3 * offset]))
    cen = pts.mean(axis=0)
    ext = float(np.abs(pts - cen).max())
    queries = [
        pts[rng.integers(0, len(pts), 64)],                                                       # on vertices
        (pts[rng.integers(0, len(pts), 64)] + rng.standard_normal((64, 3)) * 0.01 * ext),         # near the surface
        (cen + rng.standard_normal((64, 3)) * ext),                                               # inside the tile
        (cen + rng.standard_normal((64, 3)) * 30 * ext),                                          # far away
        (cen + np.array([1.0, 0.0, 0.0]) * ext * rng.uniform(-40, 40, (64, 1))),                  # along one axis
    ]
    q = np.concatenate(queries).astype(F)
    m = filter_margins(q, pts)
    # every slack is positive; the kernels' delta is 4e-6 (|q'|^2 + B^2) and the analysed need is
    # 24.5 u s^2 ~ 1.5e-6 s^2 <= 2.9e-6 (|q'|^2 + B^2), so at least ~1e-6 of it must be left over
    for name, slack in m.items():
        assert slack > 5e-7, (name, slack)


def test_margin_is_not_vacuous():
    """With delta = 0 the same data does violate the inequalities: the test can fail."""
    rng = np.random.default_rng(5)
    pts = _tile(rng, "flat", 1.0, np.array([800.0, -500.0, 300.0]))
    q = (pts[rng.integers(0, len(pts), 256)] + rng.standard_normal((256, 3)) * 1e-3).astype(F)
    m = filter_margins(q, pts)
    # slack is reported relative to (|q'|^2 + B^2); subtracting the 4e-6 the kernels add shows what
    # the raw rounding error alone would do
    assert min(m.values()) - 4.0e-6 < 0.0


# ------------------------------------------------------------------------------------------------
# q16 kernel (csrc/nn_bruteforce_q16.cuh): 16-bit quantised interval test.  Property: for every query q
# and EVERY vertex j of a tile with best = S4(q, b_j) (j ties the running best), j's quantised
# coordinate lies inside the query's integer interval on all three axes -- for the alu-only loop
# (exact per-half wrap-around) and for the split loop (plain 32-bit add: a carry out of the low half
# may raise the high half by one, the width is one step larger).
# ------------------------------------------------------------------------------------------------
Q16_LEVELS = F(65534.0)


def q16_build(pts):
    """k_build_q16_tiles: per-axis lower corner, steps per unit (fp32), u = round((p - lo) * inv) in double."""
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    ext = (hi - lo).astype(F)
    ext = np.maximum(ext, np.maximum(np.maximum(np.abs(lo), np.abs(hi)) * F(1.0e-6), F(1.0e-30))).astype(F)
    inv = (Q16_LEVELS / ext).astype(F)
    t = (pts.astype(np.float64) - lo.astype(np.float64)) * inv.astype(np.float64)
    u = np.clip(np.floor(t + 0.5), 0.0, 65534.0).astype(np.int64)
    return lo, inv, u


def q16_setup(q, T, lo, inv, carry):
    """q16_setup<CARRY>: (li, W) per query with the kernel's fp32 operation order; W = 0 means "never"."""
    c = ((q - lo).astype(F) * inv).astype(F)
    r = (T * inv).astype(F)
    e = fma(np.full_like(c, F(4.8e-7)), (np.abs(c) + np.abs(r)).astype(F), np.full_like(c, F(2.0)))
    lf = ((c - r).astype(F) - e).astype(F)
    hf = ((c + r).astype(F) + e).astype(F)
    never = (hf < 0) | (lf > Q16_LEVELS) | (T < 0)
    li = np.floor(np.clip(lf, 0.0, 65534.0)).astype(np.int64)
    hi = np.ceil(np.clip(hf, 0.0, 65534.0)).astype(np.int64)
    W = hi - li + 1 + ((li > 0) & carry)
    return li, np.where(never, 0, W)


def q16_hit(u_lo_half, u_hi_half, li, W, carry):
    """The hot loop's test on one word (two vertices): per half (x + (65536 - li)) mod 2^16 < W; with
    `carry` the two halves are added as ONE 32-bit integer (IMAD), so the low half's carry enters the high."""
    nlo = (65536 - li) & 0xFFFF
    if not carry:
        return ((u_lo_half + nlo) & 0xFFFF) < W, ((u_hi_half + nlo) & 0xFFFF) < W
    word = (u_lo_half | (u_hi_half << 16)) + (nlo | (nlo << 16))
    return (word & 0xFFFF) < W, ((word >> 16) & 0xFFFF) < W


@pytest.mark.parametrize("carry", [False, True])
@pytest.mark.parametrize("kind", ["blob", "flat", "needle", "lattice"])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1.0, 800.0), (1e-4, 0.0), (1e-4, 0.3), (1e3, -5e4)])
def test_q16_interval_never_misses_a_vertex_that_ties_the_best(kind, scale, offset, carry):
    rng = np.random.default_rng(zlib.crc32(repr(("q16", kind, scale, offset)).encode()))
    pts = _tile(rng, kind, scale, np.array([offset, -0.6 * offset, 0.3 * offset]))
    cen = pts.mean(axis=0)
    ext = float(np.abs(pts - cen).max())
    q = np.concatenate([
        pts[rng.integers(0, len(pts), 48)],
        pts[rng.integers(0, len(pts), 48)] + rng.standard_normal((48, 3)) * 0.01 * ext,
        cen + rng.standard_normal((48, 3)) * ext,
        cen + rng.standard_normal((48, 3)) * 30 * ext,
    ]).astype(F)
    lo, inv, u = q16_build(pts)
    nq, nv = q.shape[0], pts.shape[0]
    best = s4_d2(np.broadcast_to(q[:, None, :], (nq, nv, 3)), np.broadcast_to(pts[None], (nq, nv, 3)))
    T = (np.sqrt(best).astype(F) * F(1.000001)).astype(F)              # q16_radius
    neighbour = np.roll(u, 1, axis=0)                                   # the other vertex of the same word
    for axis in range(3):
        li, W = q16_setup(q[:, None, axis], T, lo[axis], inv[axis], carry)
        assert (W > 0).all(), "a vertex at distance `best` made its own tile unreachable"
        # j in the low half of a word (any neighbour above it), and j in the high half (the low half's
        # carry may or may not arrive: try the neighbour that carries and the one that does not)
        hit_low, _ = q16_hit(u[None, :, axis], neighbour[None, :, axis], li, W, carry)
        assert hit_low.all()
        for other in (neighbour[None, :, axis], np.zeros_like(u[None, :, axis]), np.full_like(u[None, :, axis], 65534)):
            _, hit_high = q16_hit(np.broadcast_to(other, li.shape), u[None, :, axis], li, W, carry)
            assert hit_high.all()
        # padding (u = 65535, the high half of a ragged tile's last words) must not be needed for
        # correctness; it may only cost a false positive, which the exact metric rejects (coordinates +inf)


@pytest.mark.parametrize("kind", ["blob", "flat", "needle", "lattice"])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1.0, 800.0), (1e-4, 0.0), (1e-4, 0.3), (1e3, -5e4)])
def test_q16_tile_cull_never_drops_a_tile_that_holds_a_tie_of_the_best(kind, scale, offset):
    """The tile-level cull of k_nn_bruteforce_q16 (fp32, the kernel's operation order): a tile is skipped for a
    query when the tile's exact bounding box misses [q - T, q + T] on ANY axis, T = sqrt(best)(1 + 1e-6)(1 + 1e-6).
    Property: a tile that holds a vertex j with S4(q, b_j) == best is never skipped -- and the cull is not
    vacuous: with half the radius it does drop such tiles."""
    rng = np.random.default_rng(zlib.crc32(repr(("q16cull", kind, scale, offset)).encode()))
    pts = _tile(rng, kind, scale, np.array([offset, -0.6 * offset, 0.3 * offset]))
    cen = pts.mean(axis=0)
    ext = float(np.abs(pts - cen).max())
    q = np.concatenate([
        pts[rng.integers(0, len(pts), 48)] + rng.standard_normal((48, 3)) * 0.01 * ext,
        cen + rng.standard_normal((48, 3)) * ext,
        cen + rng.standard_normal((48, 3)) * 30 * ext,
    ]).astype(F)
    nq, nv = q.shape[0], pts.shape[0]
    best = s4_d2(np.broadcast_to(q[:, None, :], (nq, nv, 3)), np.broadcast_to(pts[None], (nq, nv, 3)))

    def missed(tm, lo, hi):
        m = np.zeros(tm.shape, bool)
        for a in range(3):
            qa = q[:, None, a]
            m |= ((qa - hi[a]).astype(F) > tm) | ((lo[a] - qa).astype(F) > tm)
        return m

    # sub-tiles of 8 vertices each: every vertex is the "best" once, its own sub-tile must survive
    order = np.argsort(pts[:, 0], kind="stable")
    for g in range(0, nv, 64):
        sel = order[g:g + 8]
        lo, hi = pts[sel].min(axis=0), pts[sel].max(axis=0)
        T = (np.sqrt(best[:, sel]).astype(F) * F(1.000001)).astype(F)      # q16_radius
        assert not missed((T * F(1.000001)).astype(F), lo, hi).any()
    far = cen + np.array([40 * ext, 0, 0])
    lo, hi = (pts + far).astype(F).min(axis=0), (pts + far).astype(F).max(axis=0)
    T = (np.sqrt(best.min(axis=1, keepdims=True)).astype(F) * F(1.000001)).astype(F)
    assert missed((T * F(1.000001)).astype(F), lo, hi)[:96].all()         # a tile far beyond the radius is dropped


def test_q16_filter_actually_prunes():
    """Not vacuous: far vertices are outside the interval, and a radius that is too small loses vertices."""
    rng = np.random.default_rng(9)
    pts = _tile(rng, "flat", 1.0, np.zeros(3))
    lo, inv, u = q16_build(pts)
    q = pts[:64] + F(1e-3)
    best = s4_d2(q, pts[:64])
    T = (np.sqrt(best).astype(F) * F(1.000001)).astype(F)
    li, W = q16_setup(q[:, None, 0], T[:, None], lo[0], inv[0], False)
    inside = ((u[None, :, 0] - li) % 65536) < W
    assert inside.mean() < 0.2                                          # most of the tile is pruned along x alone
    li, W = q16_setup(q[:, None, 0], (T * F(0.5))[:, None], lo[0], inv[0], False)
    own = ((u[None, :64, 0] - li) % 65536) < W
    assert not np.diagonal(own).all()                                   # half the radius drops true neighbours


def test_global_pruning_bound_is_valid():
    """k_transform_queries' bound for base-sharded runs: the previous pass's GLOBAL winner v of a slot lies
    within sqrt(d_old) + |q_new - q_old| of the new query, so the new winner's S4 distance never exceeds
    (sqrt(d_old) + |q_new - q_old|)^2 (1 + 4e-6) -- whatever moved the query."""
    rng = np.random.default_rng(17)
    base = (rng.standard_normal((4000, 3)) * 0.5).astype(F)
    q_old = (rng.standard_normal((3000, 3)) * 0.5).astype(F)
    d_all = s4_d2(np.broadcast_to(q_old[:, None, :], (3000, 4000, 3)), np.broadcast_to(base[None], (3000, 4000, 3)))
    win = d_all.argmin(axis=1)
    d_old = d_all[np.arange(3000), win]
    for jump in (1e-4, 1e-2, 1.0):
        q_new = (q_old + rng.standard_normal((3000, 3)) * jump).astype(F)
        e = (q_new - q_old).astype(F)
        dq = np.sqrt(fma(e[:, 2], e[:, 2], fma(e[:, 1], e[:, 1], (e[:, 0] * e[:, 0]).astype(F)))).astype(F)
        sr = (np.sqrt(d_old).astype(F) + dq).astype(F)
        bound = ((sr * sr).astype(F) * F(1.000004)).astype(F)
        d_new_same_vertex = s4_d2(q_new, base[win])
        assert (d_new_same_vertex <= bound).all()
        d_new_all = s4_d2(np.broadcast_to(q_new[:, None, :], (3000, 4000, 3)), np.broadcast_to(base[None], (3000, 4000, 3)))
        assert (d_new_all.min(axis=1) <= bound).all()


# ------------------------------------------------------------------------------------------------
# q16 query classes (csrc/nn_bruteforce_q16.cuh: k_q16_class_keys, q16_first_hard), restated: the class of a
# slot from the exponent of its squared pruning radius relative to the bin of the median, and the rule that
# decides which classes the gated expansion-filter launch serves.  (Performance logic only: every setting
# returns the same keys -- that is what test_q16_class_order_and_gated_filter_on_a_small_cloud_with_outliers
# checks on the GPU.)
# ------------------------------------------------------------------------------------------------
def q16_classes(d2):
    e = ((np.asarray(d2, F).view(np.uint32) >> 23) & 0xFF).astype(np.int64)
    hist = np.bincount(e, minlength=256)
    csum = np.cumsum(hist)
    m = int(np.argmax(2 * csum >= e.size))                     # the first bin whose prefix reaches n / 2
    cls = np.where(e <= m + 2, 0, np.minimum(15, (e - m - 1) // 2))
    return cls, m


def q16_first_hard(cls, hard_class=3):
    n, cnt = cls.size, np.bincount(cls, minlength=16)
    if (n - cnt[0]) * 10 > n:
        return int(cnt[0])
    return int(cnt[:hard_class].sum())


def test_q16_classes_are_monotone_and_the_gate_follows_the_populations():
    rng = np.random.default_rng(41)
    inl = (rng.random(100_000) * 4e-6 + 1e-7).astype(F)         # squared radii of an aligned cloud
    cls, m = q16_classes(inl)
    order = np.argsort(inl, kind="stable")
    assert (np.diff(cls[order]) >= 0).all()                     # a larger radius never gets a lower class
    assert (cls == 0).mean() > 0.95 and cls.max() <= 2
    assert q16_first_hard(cls) == inl.size                      # nothing for the gated launch: classes < 3 stay in q16
    out = np.concatenate([inl[:75_000], (rng.random(25_000) * 0.5 + 0.01).astype(F)])   # 25 % far outliers
    cls, _ = q16_classes(out)
    assert (cls[:75_000] <= 1).all() and (cls[75_000:] >= 4).all()
    assert q16_first_hard(cls) == int((cls == 0).sum())         # a cloud with outliers: everything above class 0 is gated
    tail = np.concatenate([inl[:99_000], (rng.random(1_000) * 0.5 + 0.01).astype(F)])   # a thin far tail (1 %)
    cls, _ = q16_classes(tail)
    assert q16_first_hard(cls) == int((cls < 3).sum()) == 99_000
    assert q16_first_hard(cls, hard_class=16) == tail.size      # ICP_Q16_HARD_CLASS=16: the gate is off
