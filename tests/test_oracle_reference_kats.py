"""Pins the CPU oracle against the reference's OWN unit-test logic (CPU only, no GPU needed).

The reference (philcn/DXRExperiments) cannot run here (Windows/D3D12/closed dxrfallbackcompiler.dll), so the
oracle is pinned by re-running what the Fallback Layer's MSTest suite checks, with independent numpy
restatements of the test-side code ("UT" = externals/D3D12RaytracingFallback/src/FallbackLayerUnitTests/
fallbacklayerunittests.cpp):
  * scene AABB == CPU min/max, bit exact                               UT:2627-2701, 2795-2829
  * Morton codes == UT's GetMortonCodeFromUnitCoord (low 3 bits masked) UT:2569-2615, 2831-2835
  * sorted order == sort by code, index must match                      UT:2857-2888
  * BVH structural invariants (BvhValidator)                            FL/BVHValidator.cpp:58-176, UT:544-776
  * 6x4-ray hit/miss matrices: transforms, culling, masks, empty TLAS   UT:3889-4078
The rand() sequences of the MSVC CRT are not reproducible; fixtures use our own seeded generator with the same
shape (coordinates in [-500, 500)).
"""
import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T
from helpers import partition_transform, ut_quad, ut_rays


def ut_triangles(n, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.random((n, 3, 3), dtype=np.float32) * np.float32(1000) - np.float32(500)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ scene AABB
@pytest.mark.parametrize("n", [4, 50, 1000])
def test_scene_aabb_bit_exact_vs_cpu_minmax(n, orc):
    tris = ut_triangles(n)
    got = orc.scene_aabb(orc.prims_from_triangles(tris))
    want = np.concatenate([tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)])
    assert got.tobytes() == want.astype(np.float32).tobytes()


# ------------------------------------------------------------------------------------------------ Morton + sort
def ut_morton(tris, aabb):
    """Independent numpy restatement of the TEST's CPU code: CalculateMortonCode(Triangle&, AABB) UT:2604-2608 with
    GetMortonCodeFromUnitCoord UT:2569-2590 (float3 arithmetic in fp32, (UINT) truncation, axis order y,x,z)."""
    f32 = np.float32
    c = ((tris[:, 0] + tris[:, 1]) + tris[:, 2]) / f32(3.0)
    mn, mx = aabb[:3].astype(f32), aabb[3:].astype(f32)
    unit = (c - mn) / (mx - mn)
    adj = np.minimum(np.maximum(unit * f32(1024.0), f32(0)), f32(1023.0))
    q = adj.astype(np.uint32)
    coords = [q[:, 1], q[:, 0], q[:, 2]]
    code = np.zeros(tris.shape[0], np.uint32)
    for b in range(10):
        for a in range(3):
            code |= ((coords[a] >> np.uint32(b)) & np.uint32(1)) << np.uint32(b * 3 + a)
    return code


@pytest.mark.parametrize("n", [300, 5000])
def test_morton_codes_and_sort_vs_unit_test_restatement(n, orc):
    tris = ut_triangles(n, seed=n)
    prims = orc.prims_from_triangles(tris)
    aabb = orc.scene_aabb(prims)
    codes = orc.morton_codes(prims, aabb)
    want = ut_morton(tris, aabb)
    mask = np.uint32(~np.uint32(7))  # IsMortonCodeEqual: UT:2831-2835
    assert ((codes & mask) == (want & mask)).all()
    assert (codes < (1 << 30)).all()
    sorted_codes, perm = orc.sort_pairs(codes)
    order = np.argsort(codes, kind="stable")  # std::sort by code; ties resolved by index (BitonicSortCommon.hlsli:37-47)
    np.testing.assert_array_equal(perm, order.astype(np.uint32))
    np.testing.assert_array_equal(sorted_codes, codes[order])
    assert (np.diff(sorted_codes.astype(np.int64)) >= 0).all()


def test_sort_is_stable_on_duplicate_keys(orc):
    codes = np.array([5, 1, 5, 1, 0, 5, 1, 0], np.uint32)
    s, p = orc.sort_pairs(codes)
    np.testing.assert_array_equal(s, [0, 0, 1, 1, 1, 5, 5, 5])
    np.testing.assert_array_equal(p, [4, 7, 1, 3, 6, 0, 2, 5])


# ------------------------------------------------------------------------------------------------ BVH validator
def validate_bvh(blob, expected_tris, eps=1e-3):
    """BvhValidator::VerifyBVHOutput (FL/BVHValidator.cpp:58-176): breadth-first; children inside parents; at every
    level every expected triangle fits in some node; every triangle found in exactly one leaf."""
    b = T.parse_blas_blob(blob)
    nodes, prims = b["nodes"], b["prims"]
    lo = nodes["center"] - nodes["halfDim"]
    hi = nodes["center"] + nodes["halfDim"]
    remaining = {i: t for i, t in enumerate(expected_tris)}
    level = [0]
    seen_leaves = 0
    while level:
        nxt = []
        found = set()
        for ni in level:
            for i, t in remaining.items():
                if (t >= lo[ni] - eps).all() and (t <= hi[ni] + eps).all():
                    found.add(i)
            if nodes["flags"][ni] & T.LEAF_FLAG:
                slot = int(nodes["flags"][ni] & 0xFFFFFF)
                tri = prims["v"][slot].reshape(3, 3)
                match = [i for i, t in remaining.items() if np.array_equal(t, tri)]
                assert match, "leaf triangle is not one of the expected triangles"
                del remaining[match[-1]]
                found.discard(match[-1])
                seen_leaves += 1
            else:
                l, r = int(nodes["flags"][ni] & 0xFFFFFF), int(nodes["right"][ni])
                assert l != 0 and r != 0, "circular reference to the root"
                for c in (l, r):
                    assert (lo[c] >= lo[ni] - eps).all() and (hi[c] <= hi[ni] + eps).all(), "child box not inside parent"
                    nxt.append(c)
        assert all(i in found for i in remaining if True) or not remaining or nxt, "a level cannot contain a leaf"
        for i in remaining:
            assert i in found, "one of the BVH levels has AABBs that can't contain one of the triangles"
        level = nxt
    assert not remaining, "a triangle was never found in a leaf"
    return seen_leaves


def reference_vertices(n_tris):
    """ReferenceVerticies0/1 (UT:544-604): unit triangles stacked at z = 0, 1, 2, ..."""
    tris = []
    for k in range(n_tris):
        z = float(k % 3) + 3.0 * (k // 3)
        tris.append([[0.0, 1.0, z], [1.0, 0.0, z], [-1.0, 0.0, z]])
    return np.array(tris, np.float32)


@pytest.mark.parametrize("n", [1, 3, 6, 16])
def test_blas_passes_reference_validator(n, orc):
    tris = reference_vertices(n)
    blas = orc.Blas([dict(vertices=tris.reshape(-1, 3), stride=12, indices=None)])
    assert validate_bvh(blas.blob(), list(tris)) == n


def test_blas_validator_identical_and_stress(orc):
    same = np.repeat(reference_vertices(1), 16, axis=0)          # 16 identical triangles (UT:742-751)
    blas = orc.Blas([dict(vertices=same.reshape(-1, 3), stride=12, indices=None)])
    assert validate_bvh(blas.blob(), list(same)) == 16
    stress = ut_triangles(300, seed=7)                            # many small geometries (UT:753-776 shape)
    geoms = [dict(vertices=stress[i:i + 3].reshape(-1, 3), stride=12, indices=None) for i in range(0, 300, 3)]
    blas = orc.Blas(geoms)
    assert validate_bvh(blas.blob(), list(stress)) == 300
    meta = T.parse_blas_blob(blas.blob())["meta"]
    assert sorted(zip(meta["geom"].tolist(), meta["prim"].tolist())) == [(g, p) for g in range(100) for p in range(3)]


def test_hierarchy_is_a_proper_binary_radix_tree(orc):
    rng = np.random.Generator(np.random.PCG64(1))
    codes = np.sort(rng.integers(0, 1 << 30, size=2000, dtype=np.uint32))
    codes[100:140] = codes[100]  # duplicate run: ties resolved by index (BuildBVHSplits.hlsli:49-53)
    codes = np.sort(codes)
    n = codes.size
    h = orc.build_hierarchy(codes)
    covered = {}

    def span(v):
        if v >= n - 1:
            return v - (n - 1), v - (n - 1)
        l, r = int(h["left"][v]), int(h["right"][v])
        assert int(h["parent"][l]) == v and int(h["parent"][r]) == v
        a, b = span(l)
        c, d = span(r)
        assert b + 1 == c, "children must cover adjacent key ranges"
        covered[v] = (a, d)
        return a, d

    import sys
    sys.setrecursionlimit(10000)
    assert span(0) == (0, n - 1)
    assert len(covered) == n - 1


# ------------------------------------------------------------------------------------------------ tracing known answers
def _hit_grid(hits):
    return (hits["primitive_index"] != T.NO_HIT).reshape(4, 6)


def _scene(orc, specs):
    blases, xf, fl, mk = [], [], [], []
    for kind, winding, tr, flags, mask in specs:
        verts, idx = ut_quad(kind, winding)
        blases.append(orc.Blas([dict(vertices=verts, stride=12, indices=idx, flags=T.GEOMETRY_FLAG_OPAQUE)]))
        xf.append(tr), fl.append(flags), mk.append(mask)
    return orc.Tlas(blases, xf, masks=mk, flags=fl)


X = np.arange(6)[None, :].repeat(4, 0)
Y = np.arange(4)[:, None].repeat(6, 1)


@pytest.mark.parametrize("transform,want", [
    (scenes.IDENTITY_3X4, X < 3),                                                                # BasicTrace
    (np.array([-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), X >= 3),                       # ...WithInstanceFlip
    (np.array([np.cos(1.57), np.sin(1.57), 0, 0, -np.sin(1.57), np.cos(1.57), 0, 0, 0, 0, 1, 0], np.float32), Y >= 2),
    (np.array([1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), X >= 3),                        # ...WithInstanceTranslation
])
def test_basic_trace_known_answers(transform, want, orc):
    tlas = _scene(orc, [("left", "cw", transform, 0, 0xFF)])
    np.testing.assert_array_equal(_hit_grid(tlas.trace(ut_rays(), 0)), want)


@pytest.mark.parametrize("cull", [0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES])
def test_culling_known_answers(cull, orc):
    specs, want = [], []
    for i in range(6):
        winding = "ccw" if i < 3 else "cw"
        iflag = [0, T.INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE, T.INSTANCE_FLAG_TRIANGLE_CULL_DISABLE][i % 3]
        specs.append(("full", winding, partition_transform(i, 6), iflag, 0xFF))
        front = (winding == "cw" and not (iflag & 2)) or (winding == "ccw" and (iflag & 2))
        if iflag == T.INSTANCE_FLAG_TRIANGLE_CULL_DISABLE or cull == 0:
            want.append(True)
        elif front:
            want.append(not (cull & T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES))
        else:
            want.append(not (cull & T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES))
    tlas = _scene(orc, specs)
    np.testing.assert_array_equal(_hit_grid(tlas.trace(ut_rays(), cull)), np.array(want)[None, :].repeat(4, 0))


def test_instance_mask_and_empty_tlas_known_answers(orc):
    tlas = _scene(orc, [("full", "cw", partition_transform(i, 6), 0, 1 << i) for i in range(6)])
    want = np.array([bool((1 << i) & 0x23) for i in range(6)])[None, :].repeat(4, 0)
    np.testing.assert_array_equal(_hit_grid(tlas.trace(ut_rays(), 0, mask=0x23)), want)
    empty = orc.Tlas([], [])
    assert not _hit_grid(empty.trace(ut_rays(), 0)).any()


def test_traversal_equals_brute_force_closest_hit(orc):
    """Independent check of the traversal: the closest hit over ALL triangles tested one by one (each as its own
    single-triangle BLAS, i.e. no hierarchy involved) equals the BVH traversal result."""
    mesh = scenes.icosphere(2)
    tris = mesh.triangles()
    blas = orc.Blas.from_mesh(mesh)
    tlas = orc.Tlas([blas], [scenes.IDENTITY_3X4])
    from helpers import random_rays
    rays = np.concatenate([random_rays(300, seed=5, lo=(-0.5, -0.5, -0.5), hi=(0.5, 0.5, 0.5)),   # from inside: all hit
                           random_rays(300, seed=6, lo=(-2, -2, -2), hi=(2, 2, 2))])             # from outside: some miss
    hits = tlas.trace(rays, 0)
    best_t = np.full(rays.shape[0], np.inf, np.float32)
    best_p = np.full(rays.shape[0], T.NO_HIT, np.uint32)
    for k in range(tris.shape[0]):
        one = orc.Blas([dict(vertices=tris[k], stride=12, indices=None)])
        h = orc.Tlas([one], [scenes.IDENTITY_3X4]).trace(rays, 0)
        better = (h["primitive_index"] != T.NO_HIT) & (h["t"] < best_t)
        best_t[better] = h["t"][better]
        best_p[better] = k
    hit = best_p != T.NO_HIT
    assert hit.mean() > 0.3
    np.testing.assert_array_equal(hits["primitive_index"] != T.NO_HIT, hit)
    np.testing.assert_array_equal(hits["t"][hit], best_t[hit])
    same = hits["primitive_index"][hit] == best_p[hit]
    assert same.mean() > 0.995  # the rest are exact ties on shared edges (first found wins in both, in different orders)
