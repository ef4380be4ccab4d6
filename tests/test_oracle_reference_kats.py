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


# ------------------------------------------------------------------------------------------------ treelet pass
def _check_hierarchy(h, n):
    """The walk of TestTreeletReordering (UT:3041-3066): every child's ParentIndex is its parent, n leaves reachable.
    On the C++ side ParentIndex is a 31-bit field beside bCollapseChildren (FL/RayTracingHlslCompat.h:46-58), so the
    reference's comparison ignores bit 31."""
    stack, leaves, internal = [0], 0, 0
    while stack:
        v = stack.pop()
        if v >= n - 1:
            leaves += 1
            continue
        internal += 1
        l, r = int(h["left"][v]), int(h["right"][v])
        for c in (l, r):
            p = int(h["parent"][c])
            assert p & 0x7FFFFFFF == v, "incorrect parent index"
        stack += [l, r]
    assert leaves == n and internal == n - 1, "incorrectly constructed hierarchy"


@pytest.mark.parametrize("flag", [T.BUILD_FLAG_PREFER_FAST_TRACE, T.BUILD_FLAG_PREFER_FAST_BUILD, 0])
def test_treelet_reordering_reference_unit_test(flag, orc):
    """TreeletReorderingFastTrace / FastBuild (UT:2947-3068): 16 point triangles at x = -i (i < 8) or +i, a heap-shaped
    hierarchy (children 2i+1, 2i+2), then the parent/leaf-count walk."""
    n = 16
    tris = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        tris[i, :, 0] = -i if i < n // 2 else i
    h = np.zeros(2 * n - 1, T.HIER_DTYPE)
    for i in range(2 * n - 1):
        h["parent"][i] = (i - 1) // 2 if i else 0
        h["left"][i], h["right"][i] = 2 * i + 1, 2 * i + 2
    out = orc.treelet_optimise(h, orc.prims_from_triangles(tris), flag)
    _check_hierarchy(out, n)
    if flag == T.BUILD_FLAG_PREFER_FAST_BUILD:  # zero passes (FL/TreeletReorder.cpp:66-69)
        assert out.tobytes() == h.tobytes()
    else:
        assert out.tobytes() != h.tobytes(), "the interleaved point cloud must be re-partitioned"


def _leaf_aabb(tri):
    f32 = np.float32
    mn, mx = tri.min(0), tri.max(0)
    mn = np.minimum(mn, mx - f32(0.001))
    c = (mn + mx) * f32(0.5)
    hd = mx - c
    return c - hd, c + hd


def _area(mn, mx):
    d = (mx - mn).astype(np.float32)
    return np.float32(2.0) * ((d[0] * d[1] + d[0] * d[2]) + d[1] * d[2])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_treelet_of_seven_is_the_optimum_of_the_reference_cost_model(seed, orc):
    """n = 7: the root is the only treelet, so the output tree must minimise the reference's own cost function
    (TreeletReorder.hlsl:126-183, COMBINE_LEAF_NODES = 1), evaluated here by an independent memoised recursion."""
    f32 = np.float32
    rng = np.random.Generator(np.random.PCG64(seed))
    tris = (rng.random((7, 3, 3), dtype=np.float32) * f32(4.0)).astype(np.float32)
    prims = orc.prims_from_triangles(tris)
    codes, perm = orc.sort_pairs(orc.morton_codes(prims, orc.scene_aabb(prims)))
    h0 = orc.build_hierarchy(codes)
    out = orc.treelet_optimise(h0, prims[perm], 0)
    _check_hierarchy(out, 7)
    boxes = [_leaf_aabb(tris[perm[i]]) for i in range(7)]

    def union(mask):
        idx = [i for i in range(7) if mask >> i & 1]
        return np.min([boxes[i][0] for i in idx], 0), np.max([boxes[i][1] for i in idx], 0)

    root_area = _area(*union(127))
    memo = {}

    def best(mask):
        if mask in memo:
            return memo[mask]
        k = bin(mask).count("1")
        a = _area(*union(mask))
        if k == 1:
            r = f32(1.2) * a / root_area
        else:
            low, sub = None, (mask - 1) & mask
            while sub:
                c = best(sub) + best(mask ^ sub)
                low = c if low is None or c < low else low
                sub = (sub - 1) & mask
            r = min(f32(1.2) * a + low, f32(1.0) * a * f32(k))
        memo[mask] = f32(r)
        return memo[mask]

    def tree_cost(v):
        """cost of the oracle's output subtree under the same model; returns (cost, leaf mask)"""
        if v >= 6:
            m = 1 << (v - 6)
            return best(m), m
        (cl, ml), (cr, mr) = tree_cost(int(out["left"][v])), tree_cost(int(out["right"][v]))
        m = ml | mr
        a = _area(*union(m))
        return f32(min(f32(1.2) * a + (cl + cr), f32(1.0) * a * f32(bin(m).count("1")))), m

    got, mask = tree_cost(0)
    assert mask == 127
    assert got == best(127)


@pytest.mark.parametrize("flags", [0, T.BUILD_FLAG_PREFER_FAST_TRACE])
def test_blas_with_treelet_pass_passes_reference_validator_and_traces_identically(flags, orc):
    tris = ut_triangles(300, seed=11) * np.float32(0.01)
    geoms = [dict(vertices=tris.reshape(-1, 3), stride=12, indices=None)]
    plain = orc.Blas(geoms, T.BUILD_FLAG_PREFER_FAST_BUILD)
    opt = orc.Blas(geoms, flags)
    assert validate_bvh(opt.blob(), list(tris)) == 300
    _check_hierarchy(opt.hierarchy(), 300)
    assert opt.hierarchy().tobytes() != plain.hierarchy().tobytes()
    np.testing.assert_array_equal(opt.perm(), plain.perm())          # the pass touches only the topology
    rng = np.random.Generator(np.random.PCG64(5))
    rays = np.zeros(2000, T.RAY_DTYPE)
    rays["origin"] = rng.random((2000, 3), dtype=np.float32) * 10 - 5
    d = rng.standard_normal((2000, 3)).astype(np.float32)
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmax"] = 1e30
    a = orc.Tlas([plain], [scenes.IDENTITY_3X4]).trace(rays)
    b = orc.Tlas([opt], [scenes.IDENTITY_3X4]).trace(rays)
    np.testing.assert_array_equal(a["primitive_index"], b["primitive_index"])
    np.testing.assert_array_equal(a["t"], b["t"])


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


# ------------------------------------------------------------------------------------------------ ALLOW_UPDATE / PERFORM_UPDATE
def scrambled_vertices(tris):
    """RefitAABBsOnUpdate (UT:1277-1292): v' = ((v + vertexIndex) * (vertexIndex even ? -8 : 8)) * (triangle even ? -12 : 12)."""
    flat = tris.reshape(-1).astype(np.float32).copy()
    out = np.empty_like(flat)
    for i in range(flat.size):
        vi = i // 3
        ti = vi // 3
        out[i] = np.float32(np.float32(np.float32(flat[i] + np.float32(vi)) * np.float32(-8 if vi % 2 == 0 else 8)) *
                            np.float32(-12 if ti % 2 == 0 else 12))
    return out.reshape(tris.shape)


def test_store_sort_result_for_update(orc):
    """StoreSortResultForUpdate (UT:1089-1125): prims[sortCache[i]] is input triangle i."""
    tris = reference_vertices(6)  # ReferenceVerticies1 has 6 triangles
    blas = orc.Blas([dict(vertices=tris.reshape(-1, 3), stride=12, indices=None)], build_flags=T.BUILD_FLAG_ALLOW_UPDATE | T.BUILD_FLAG_PREFER_FAST_BUILD)
    prims = T.parse_blas_blob(blas.blob())["prims"]
    cache = blas.sort_cache()
    assert sorted(cache.tolist()) == list(range(6))
    for i in range(6):
        np.testing.assert_array_equal(prims["v"][cache[i]].reshape(3, 3), tris[i])
    np.testing.assert_array_equal(blas.perm()[cache], np.arange(6))


@pytest.mark.parametrize("n", [2, 6, 300])
def test_store_parent_indices_for_update(n, orc):
    """StoreParentIndicesForUpdate (UT:1127-1165): both children of every internal node name it as their parent."""
    tris = reference_vertices(n) if n <= 6 else ut_triangles(n, seed=11)
    blas = orc.Blas([dict(vertices=tris.reshape(-1, 3), stride=12, indices=None)], build_flags=T.BUILD_FLAG_ALLOW_UPDATE)
    nodes = T.parse_blas_blob(blas.blob())["nodes"]
    parents = blas.parents()
    assert parents.size == 2 * n - 1
    for p in range(2 * n - 1):
        if not (nodes["flags"][p] & T.LEAF_FLAG):
            assert parents[int(nodes["flags"][p] & 0xFFFFFF)] == p
            assert parents[int(nodes["right"][p])] == p


def test_refit_aabbs_on_update(orc):
    """RefitAABBsOnUpdate (UT:1167-1292): build, PERFORM_UPDATE with scrambled vertices, then the reference validator
    must accept the result for the UPDATED triangles; the topology (child links) must be the one of the first build."""
    tris = reference_vertices(6)
    moved = scrambled_vertices(tris)
    blas = orc.Blas([dict(vertices=tris.reshape(-1, 3), stride=12, indices=None)], build_flags=T.BUILD_FLAG_ALLOW_UPDATE | T.BUILD_FLAG_PREFER_FAST_BUILD)
    before = T.parse_blas_blob(blas.blob())
    blas.update([dict(vertices=moved.reshape(-1, 3), stride=12, indices=None)])
    after = T.parse_blas_blob(blas.blob())
    assert validate_bvh(blas.blob(), list(moved)) == 6
    np.testing.assert_array_equal(before["nodes"]["flags"], after["nodes"]["flags"])
    np.testing.assert_array_equal(before["nodes"]["right"], after["nodes"]["right"])
    assert not np.array_equal(before["nodes"]["center"], after["nodes"]["center"])
    with pytest.raises(ValueError):
        blas.update([dict(vertices=moved.reshape(-1, 3)[:9], stride=12, indices=None)])


def test_update_with_unchanged_input_is_the_identity_and_matches_a_rebuild_when_order_is_kept(orc):
    """Two properties of the refit that need no reference run: (i) updating with the original vertices reproduces the
    original blob bit for bit; (ii) a rigid translation keeps the Morton order, so update == fresh build of the moved mesh."""
    from dxrexperiments_b200 import scenes
    mesh = scenes.icosphere(3)
    g = [dict(vertices=mesh.vertices, stride=24, indices=mesh.indices)]
    blas = orc.Blas(g, build_flags=T.BUILD_FLAG_ALLOW_UPDATE)
    blob0 = blas.blob()
    blas.update(g)
    np.testing.assert_array_equal(blas.blob(), blob0)
    # power-of-two scale about the origin: every coordinate, centroid and box scales exactly, the Morton codes are equal.
    # PREFER_FAST_BUILD: the treelet pass's cost model mixes normalised and raw areas (TreeletReorder.hlsl:126-183), so
    # its topology is not scale invariant; the plain Karras tree is.
    fast = T.BUILD_FLAG_ALLOW_UPDATE | T.BUILD_FLAG_PREFER_FAST_BUILD
    blas = orc.Blas(g, build_flags=fast)
    moved = mesh.vertices.copy()
    moved["position"] *= np.float32(4.0)
    g2 = [dict(vertices=moved, stride=24, indices=mesh.indices)]
    fresh = orc.Blas(g2, build_flags=fast)
    np.testing.assert_array_equal(fresh.perm(), blas.perm())
    blas.update(g2)
    a, b = T.parse_blas_blob(blas.blob()), T.parse_blas_blob(fresh.blob())
    np.testing.assert_array_equal(a["prims"].view(np.uint8), b["prims"].view(np.uint8))
    np.testing.assert_array_equal(a["nodes"]["flags"], b["nodes"]["flags"])
    # leaf boxes carry the absolute 0.001 padding, so only un-padded internal extents scale exactly; all must contain
    np.testing.assert_allclose(a["nodes"]["center"], b["nodes"]["center"], rtol=0, atol=0)
    np.testing.assert_allclose(a["nodes"]["halfDim"], b["nodes"]["halfDim"], rtol=0, atol=0)


def test_tlas_update_refits_instances(orc):
    """TopLevel ..._WithUpdate (UT:913-935): rebuild with PERFORM_UPDATE and new transforms; metadata follows the cached
    order, every instance's world box is inside the root box, and traversal sees the moved instances."""
    from dxrexperiments_b200 import scenes
    from helpers import random_rays
    mesh = scenes.icosphere(1)
    ob = orc.Blas.from_mesh(mesh)
    n = 40
    xf0 = scenes.random_rigid_transforms(n, seed=3)
    xf1 = scenes.random_rigid_transforms(n, seed=4)
    t = orc.Tlas([ob] * n, xf0, build_flags=T.BUILD_FLAG_ALLOW_UPDATE)
    perm = t.perm()
    np.testing.assert_array_equal(perm[t.sort_cache()], np.arange(n))
    t.update(xf1)
    b = T.parse_tlas_blob(t.blob())
    np.testing.assert_array_equal(b["meta"]["instance_index"], perm)            # order of the ORIGINAL build
    np.testing.assert_array_equal(b["meta"]["o2w"], np.asarray(xf1, np.float32).reshape(n, 12)[perm])
    fresh = orc.Tlas([ob] * n, xf1)
    rays = random_rays(4000, 5, -60, 60)
    h_upd, h_new = t.trace(rays), fresh.trace(rays)
    for f in ("t", "primitive_index", "instance_index"):
        np.testing.assert_array_equal(h_upd[f], h_new[f])                        # same scene => same closest hits


def test_any_hit_visibility_does_not_depend_on_the_visiting_order(orc):
    """The production any-hit kernel visits the four children of a node in slot order instead of near-to-far
    (DESIGN 4.2).  That is only legitimate because an ACCEPT_FIRST_HIT search answers "is there ANY accepted hit in
    (tMin, tMax)": the reference-order search must agree with the existence of a closest hit for every ray, with and
    without culling flags, on single-level and instanced scenes."""
    from helpers import bunny_case, random_rays, two_material_case
    for case in (bunny_case(3), two_material_case()):
        otlas, _ = case.oracle(orc)
        rays = random_rays(6000, seed=23, lo=(-8, 0.05, -8), hi=(8, 9, 8), tmin=1e-4)
        rays["tmax"][::3] = 2.5  # a third of the rays are short, like point-light shadow rays
        for cull in (0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES):
            closest = otlas.trace(rays, cull, threads=4)
            anyhit = otlas.trace(rays, cull | T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | T.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER, threads=4)
            np.testing.assert_array_equal(anyhit["primitive_index"] != T.NO_HIT, closest["primitive_index"] != T.NO_HIT)
            # and what the any-hit search reports is a genuine hit no closer than the closest one
            hit = anyhit["primitive_index"] != T.NO_HIT
            assert (anyhit["t"][hit] >= closest["t"][hit]).all()
            assert (anyhit["t"][hit] < rays["tmax"][hit]).all() and (anyhit["t"][hit] > rays["tmin"][hit]).all()
