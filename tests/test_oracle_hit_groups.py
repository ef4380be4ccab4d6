"""CPU: the oracle's procedural-primitive build and its any-hit / intersection machinery.

Pinned by the reference: the AABB load (UT:2616-2664), the structural rules of FL/BVHValidator.cpp:58-176 applied to a
procedural BLAS, the leaf-flag encoding (FL/BottomLevelComputeAABBs.hlsl:32-40), IsOpaque / Cull
(FL/TraverseFunction.hlsli:119-134, :193-198) and the literal control flow of Fallback_ReportHit (:136-158) and the
leaf handler (:651-722).  The intersection and any-hit PROGRAMS are ours (the application has only the no-op
ShadowAnyHit and no intersection shader): they are checked against closed-form answers, not against the reference.
"""
import numpy as np

import oracle
from dxrexperiments_b200 import scenes, types as T
from helpers import random_rays, ut_quad, ut_rays


def _aabbs(n, seed, lo=-10.0, hi=10.0, size=(0.2, 1.5)):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(lo, hi, size=(n, 3))
    h = rng.uniform(size[0], size[1], size=(n, 3)) * 0.5
    return np.concatenate([c - h, c + h], axis=1).astype(np.float32)


def test_load_procedural_geometry_kat():
    # LoadProceduralGeometry test, UT:2616-2664: three AABBs incl. a degenerate one come back verbatim, typed 2
    aabbs = np.array([[-1, -1, -1, 1, 1, 1], [-1, -500, -1, 1, 2000, 1], [1, 1, 1, 1, 1, 1]], np.float32)
    b = oracle.Blas([dict(aabbs=aabbs, flags=T.GEOMETRY_FLAG_NONE)])
    p = b.unsorted_prims()
    assert (p["type"] == T.PRIMITIVE_TYPE_PROCEDURAL).all()
    np.testing.assert_array_equal(p["v"][:, :6], aabbs)
    np.testing.assert_array_equal(p["v"][:, 6:], 0)
    np.testing.assert_array_equal(b.scene_aabb(), np.array([-1, -500, -1, 1, 2000, 1], np.float32))


def test_aabb_stride_is_honoured():
    aabbs = _aabbs(50, 3)
    padded = np.zeros((50, 10), np.float32)
    padded[:, :6] = aabbs
    padded[:, 6:] = 777.0
    a = oracle.Blas([dict(aabbs=aabbs)])
    b = oracle.Blas([dict(aabbs=padded, stride=40)])
    np.testing.assert_array_equal(a.blob(), b.blob())


def test_procedural_blas_structure():
    aabbs = _aabbs(300, 5)
    b = oracle.Blas([dict(aabbs=aabbs, flags=T.GEOMETRY_FLAG_NONE)])
    d = T.parse_blas_blob(b.blob())
    n = d["n"]
    assert n == 300
    nodes, prims, meta = d["nodes"], d["prims"], d["meta"]
    leaf = (nodes["flags"] & T.LEAF_FLAG) != 0
    assert leaf.sum() == n and (~leaf).sum() == n - 1
    # every leaf carries IsLeafFlag | IsProceduralGeometryFlag and its box is the AABB itself (no padding)
    assert ((nodes["flags"][leaf] & T.PROCEDURAL_FLAG) != 0).all()
    slots = nodes["flags"][leaf] & 0x00FFFFFF
    assert sorted(slots.tolist()) == list(range(n))
    mn, mx = prims["v"][slots][:, 0:3], prims["v"][slots][:, 3:6]
    c = (mn + mx) * np.float32(0.5)
    np.testing.assert_array_equal(nodes["center"][leaf], c)
    np.testing.assert_array_equal(nodes["halfDim"][leaf], mx - c)
    # metadata follows the sort; Morton codes come from the AABB centre
    perm = b.perm()
    np.testing.assert_array_equal(meta["prim"], perm)
    codes = b.morton()
    sa = b.scene_aabb()
    for i in (0, 17, 299):
        cc = ((aabbs[i, :3] + aabbs[i, 3:]) / np.float32(2.0)).astype(np.float32)
        assert codes[i] == oracle.morton_code_from_centroid(cc, sa)
    # BVHValidator: children boxes inside the parent's (FL/BVHValidator.cpp:120-150)
    for i in np.nonzero(~leaf)[0]:
        l, r = nodes["flags"][i] & 0x00FFFFFF, nodes["right"][i]
        for ch in (l, r):
            assert (nodes["center"][ch] - nodes["halfDim"][ch] >= nodes["center"][i] - nodes["halfDim"][i] - 1e-4).all()
            assert (nodes["center"][ch] + nodes["halfDim"][ch] <= nodes["center"][i] + nodes["halfDim"][i] + 1e-4).all()


def _sphere_reference(rays, aabbs):
    """Closest inscribed-sphere hit in float64 (brute force)."""
    c = (aabbs[:, :3].astype(np.float64) + aabbs[:, 3:]) * 0.5
    r = (aabbs[:, 3:].astype(np.float64) - c).min(axis=1)
    o, d = rays["origin"].astype(np.float64), rays["direction"].astype(np.float64)
    best = np.full(len(rays), np.inf)
    idx = np.full(len(rays), -1)
    for k in range(len(aabbs)):
        oc = o - c[k]
        a = (d * d).sum(1)
        b = (oc * d).sum(1)
        cc = (oc * oc).sum(1) - r[k] ** 2
        disc = b * b - a * cc
        ok = disc >= 0
        s = np.sqrt(np.where(ok, disc, 0))
        t0, t1 = (-b - s) / a, (-b + s) / a
        t = np.where(t0 >= rays["tmin"], t0, t1)
        ok &= (t >= rays["tmin"]) & (t < best) & (t < rays["tmax"])
        best = np.where(ok, t, best)
        idx = np.where(ok, k, idx)
    return best, idx


def test_sphere_program_against_closed_form():
    aabbs = _aabbs(120, 9, size=(0.8, 2.5))
    b = oracle.Blas([dict(aabbs=aabbs)])
    t = oracle.Tlas([b], [scenes.IDENTITY_3X4])
    rays = random_rays(4000, seed=2, lo=(-12, -12, -12), hi=(12, 12, 12), tmin=1e-3)
    h = t.trace_hit_groups(rays, [[T.ANYHIT_NONE, T.INTERSECTION_SPHERE]], threads=4)
    best, idx = _sphere_reference(rays, aabbs)
    hit = h["primitive_index"] != T.NO_HIT
    assert hit.sum() > 400
    # robust comparison: skip rays whose discriminant is ill-conditioned in fp32 (grazing hits)
    agree = (hit == (idx >= 0)) & ((~hit) | (h["primitive_index"] == idx))
    assert agree.mean() > 0.995
    both = hit & (h["primitive_index"] == idx)
    np.testing.assert_allclose(h["t"][both], best[both], rtol=2e-3, atol=2e-3)
    # attributes are the unit normal's x, y; HitKind() says whether the ray entered or left the sphere
    kind = h["leaf_slot"][both] >> 24
    assert set(np.unique(kind)) <= {0, 1}
    n2 = (h["bary"][both] ** 2).sum(1)
    assert (n2 <= 1.0 + 1e-3).all()
    # without an intersection program a procedural primitive can never be hit
    h0 = t.trace_hit_groups(rays, [[T.ANYHIT_NONE, T.INTERSECTION_NONE]])
    assert (h0["primitive_index"] == T.NO_HIT).all()
    assert (t.trace(rays)["primitive_index"] == T.NO_HIT).all()


def test_box_program_enter_and_exit():
    aabbs = np.array([[-1, -1, 2, 1, 1, 4]], np.float32)
    t = oracle.Tlas([oracle.Blas([dict(aabbs=aabbs)])], [scenes.IDENTITY_3X4])
    rays = np.zeros(3, T.RAY_DTYPE)
    rays["origin"] = [[0, 0, 0], [0, 0, 3], [5, 0, 0]]
    rays["direction"] = [[0, 0, 1], [0, 0, 1], [0, 0, 1]]
    rays["tmax"] = 100.0
    h = t.trace_hit_groups(rays, [[0, T.INTERSECTION_BOX]])
    np.testing.assert_array_equal(h["t"], np.array([2.0, 1.0, 100.0], np.float32))
    np.testing.assert_array_equal(h["leaf_slot"] >> 24, [0, 1, 0])  # enter, exit (origin inside), miss
    np.testing.assert_array_equal(h["primitive_index"], [0, 0, T.NO_HIT])


def test_report_hit_runs_any_hit_only_under_force_non_opaque():
    # Fallback_ReportHit: geomOpaque is the literal `true` (FL/TraverseFunction.hlsli:147-149)
    aabbs = np.array([[-1, -1, 2, 1, 1, 4]], np.float32)
    t = oracle.Tlas([oracle.Blas([dict(aabbs=aabbs, flags=T.GEOMETRY_FLAG_NONE)])], [scenes.IDENTITY_3X4])
    rays = np.zeros(1, T.RAY_DTYPE)
    rays["direction"] = [[0, 0, 1]]
    rays["tmax"] = 100.0
    prog = [[T.ANYHIT_IGNORE, T.INTERSECTION_BOX]]
    assert t.trace_hit_groups(rays, prog)["t"][0] == 2.0  # any-hit not consulted although the geometry is non-opaque
    assert t.trace_hit_groups(rays, prog, ray_flags=T.RAY_FLAG_FORCE_NON_OPAQUE)["primitive_index"][0] == T.NO_HIT
    # the leaf-level cull uses the real geometry flag (:646-648)
    assert t.trace_hit_groups(rays, prog, ray_flags=T.RAY_FLAG_CULL_NON_OPAQUE)["primitive_index"][0] == T.NO_HIT
    assert t.trace_hit_groups(rays, prog, ray_flags=T.RAY_FLAG_CULL_OPAQUE)["t"][0] == 2.0


def _two_quads(flags_near, flags_far):
    near_v, idx = ut_quad(depth=1.0)
    far_v, _ = ut_quad(depth=2.0)
    b = oracle.Blas([dict(vertices=near_v, stride=12, indices=idx, flags=flags_near),
                     dict(vertices=far_v, stride=12, indices=idx, flags=flags_far)])
    return oracle.Tlas([b], [scenes.IDENTITY_3X4], hit_groups=[0])


def test_any_hit_on_non_opaque_triangles():
    t = _two_quads(T.GEOMETRY_FLAG_NONE, T.GEOMETRY_FLAG_OPAQUE)
    rays = ut_rays()
    mult = 1  # record = geometry index
    base = t.trace_hit_groups(rays, [[0, 0], [0, 0]], geometry_multiplier=mult)
    np.testing.assert_array_equal(base["t"], np.float32(1.0))
    assert (base["leaf_slot"] >> 24 == T.HIT_KIND_TRIANGLE_FRONT_FACE).all()
    # IgnoreHit on the near (non-opaque) quad: the far quad is the closest hit
    ign = t.trace_hit_groups(rays, [[T.ANYHIT_IGNORE, 0], [T.ANYHIT_IGNORE, 0]], geometry_multiplier=mult)
    np.testing.assert_array_equal(ign["t"], np.float32(2.0))
    np.testing.assert_array_equal(ign["geometry_index"], 1)
    # a no-op any-hit shader (the application's ShadowAnyHit) changes nothing
    acc = t.trace_hit_groups(rays, [[T.ANYHIT_ACCEPT, 0], [T.ANYHIT_ACCEPT, 0]], geometry_multiplier=mult)
    np.testing.assert_array_equal(acc.view(np.uint8), base.view(np.uint8))
    # FORCE_OPAQUE silences the any-hit shader; the instance flag does the same
    fo = t.trace_hit_groups(rays, [[T.ANYHIT_IGNORE, 0]] * 2, ray_flags=T.RAY_FLAG_FORCE_OPAQUE, geometry_multiplier=mult)
    np.testing.assert_array_equal(fo["t"], np.float32(1.0))
    # cutout: a pattern over the barycentrics ignores part of the near quad
    cut = t.trace_hit_groups(rays, [[T.ANYHIT_CUTOUT, 0], [0, 0]], geometry_multiplier=mult)
    u, v = base["bary"][:, 0], base["bary"][:, 1]
    odd = ((np.floor(8 * u).astype(int) + np.floor(8 * v).astype(int)) & 1) == 1
    assert 0 < odd.sum() < len(rays)
    np.testing.assert_array_equal(cut["t"], np.where(odd, 2.0, 1.0).astype(np.float32))


def test_accept_first_hit_quirk_is_literal():
    # FL/TraverseFunction.hlsli:721 — under ACCEPT_FIRST_HIT_AND_END_SEARCH an IGNOREd candidate still ends the search
    t = _two_quads(T.GEOMETRY_FLAG_NONE, T.GEOMETRY_FLAG_NONE)
    rays = ut_rays()
    flags = T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH
    h = t.trace_hit_groups(rays, [[T.ANYHIT_IGNORE, 0]] * 2, ray_flags=flags, geometry_multiplier=1)
    assert (h["primitive_index"] == T.NO_HIT).all()
    # AcceptHitAndEndSearch commits the first candidate presented, whichever quad that is
    e = t.trace_hit_groups(rays, [[T.ANYHIT_END_SEARCH, 0]] * 2, geometry_multiplier=1)
    assert (e["primitive_index"] != T.NO_HIT).all()
    assert set(np.unique(e["t"])) <= {np.float32(1.0), np.float32(2.0)}


def test_mixed_triangle_and_procedural_geometry():
    mesh = scenes.bunny_scale(2)
    aabbs = _aabbs(64, 21, lo=-3, hi=3, size=(0.3, 0.9))
    b = oracle.Blas([dict(vertices=mesh.vertices, stride=24, indices=mesh.indices),
                     dict(aabbs=aabbs, flags=T.GEOMETRY_FLAG_OPAQUE)])
    d = T.parse_blas_blob(b.blob())
    ntri = mesh.indices.size // 3
    assert d["n"] == ntri + 64
    assert (d["prims"]["type"] == T.PRIMITIVE_TYPE_PROCEDURAL).sum() == 64
    leaf = (d["nodes"]["flags"] & T.LEAF_FLAG) != 0
    slots = d["nodes"]["flags"][leaf] & 0x00FFFFFF
    proc = (d["nodes"]["flags"][leaf] & T.PROCEDURAL_FLAG) != 0
    np.testing.assert_array_equal(proc, d["prims"]["type"][slots] == T.PRIMITIVE_TYPE_PROCEDURAL)
    assert (d["meta"]["geom"][d["prims"]["type"] == T.PRIMITIVE_TYPE_PROCEDURAL] == 1).all()
    t = oracle.Tlas([b], [scenes.IDENTITY_3X4], hit_groups=[0])
    rays = random_rays(3000, seed=4, lo=(-6, -1, -6), hi=(6, 8, 6), tmin=1e-4)
    with_spheres = t.trace_hit_groups(rays, [[0, 0], [0, T.INTERSECTION_SPHERE]], geometry_multiplier=1)
    tri_only = t.trace_hit_groups(rays, [[0, 0], [0, 0]], geometry_multiplier=1)
    plain = t.trace(rays)
    np.testing.assert_array_equal(tri_only["t"], plain["t"])
    sph = with_spheres["geometry_index"][with_spheres["primitive_index"] != T.NO_HIT] == 1
    assert 20 < sph.sum()
    assert (with_spheres["t"] <= tri_only["t"]).all()
