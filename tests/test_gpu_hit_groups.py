"""GPU parity for SURVEY.md 8f-4: procedural (AABB) geometry in the bottom-level build and traversal with the hit
groups' any-hit / intersection programs (rt_trace_rays_hit_groups through the C ABI) against the CPU oracle.

Integer / byte results (Morton codes, permutation, the whole reference-format blob, hit IDs, hit kinds) are bit-exact;
so are t and the attributes — both sides evaluate the same unfused fp32 expressions.
"""
import numpy as np
import pytest

from dxrexperiments_b200 import rtcore as rt, scenes, types as T
from helpers import random_rays, ut_quad, ut_rays

pytestmark = pytest.mark.gpu


def _aabbs(n, seed, lo=-10.0, hi=10.0, size=(0.2, 1.5)):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(lo, hi, size=(n, 3))
    h = rng.uniform(size[0], size[1], size=(n, 3)) * 0.5
    return np.concatenate([c - h, c + h], axis=1).astype(np.float32)


def _gpu_aabb_geom(ctx, aabbs, flags=T.GEOMETRY_FLAG_OPAQUE, stride=None):
    buf = ctx.upload(np.ascontiguousarray(aabbs, np.float32))
    return dict(aabbs=buf, aabb_count=aabbs.shape[0], stride=stride or aabbs.strides[0], flags=flags)


def _gpu_mesh_geom(ctx, mesh, flags=T.GEOMETRY_FLAG_OPAQUE):
    return dict(vertices=ctx.upload(mesh.vertices), vertex_count=mesh.vertices.shape[0], stride=24,
                indices=ctx.upload(mesh.indices), index_count=mesh.indices.size, index_format=32, flags=flags)


@pytest.mark.parametrize("n,build_flags", [(3, 0), (1, 0), (500, T.BUILD_FLAG_PREFER_FAST_BUILD), (5000, 0),
                                           (20000, T.BUILD_FLAG_PREFER_FAST_TRACE)])
def test_procedural_blas_blob_bit_exact(n, build_flags, ctx, orc):
    aabbs = _aabbs(n, 100 + n) if n > 3 else np.array([[-1, -1, -1, 1, 1, 1], [-1, -500, -1, 1, 2000, 1],
                                                       [1, 1, 1, 1, 1, 1]], np.float32)[:n]  # UT:2621-2624
    ob = orc.Blas([dict(aabbs=aabbs, flags=T.GEOMETRY_FLAG_NONE)], build_flags)
    gb = ctx.build_blas([_gpu_aabb_geom(ctx, aabbs, T.GEOMETRY_FLAG_NONE)], build_flags=build_flags, keep_scratch=True)
    ctx.sync()
    np.testing.assert_array_equal(gb.stage("morton_codes"), ob.morton())
    np.testing.assert_array_equal(gb.stage("sorted_indices"), ob.perm())
    np.testing.assert_array_equal(gb.blob(), ob.blob())
    d = T.parse_blas_blob(gb.blob())
    assert (d["prims"]["type"] == T.PRIMITIVE_TYPE_PROCEDURAL).all()
    leaf = (d["nodes"]["flags"] & T.LEAF_FLAG) != 0
    assert ((d["nodes"]["flags"][leaf] & T.PROCEDURAL_FLAG) != 0).all()
    info = gb.info()
    assert info.has_procedural == 1 and info.count == n


def test_mixed_geometry_blob_and_stride(ctx, orc):
    mesh = scenes.bunny_scale(3)
    aabbs = _aabbs(200, 7, lo=-3, hi=3, size=(0.3, 0.9))
    padded = np.full((200, 9), 123.0, np.float32)
    padded[:, :6] = aabbs
    og = [dict(vertices=mesh.vertices, stride=24, indices=mesh.indices), dict(aabbs=padded, stride=36, flags=T.GEOMETRY_FLAG_NONE)]
    gg = [_gpu_mesh_geom(ctx, mesh), _gpu_aabb_geom(ctx, padded, T.GEOMETRY_FLAG_NONE, stride=36)]
    for bf in (T.BUILD_FLAG_PREFER_FAST_BUILD, 0):
        ob = orc.Blas(og, bf)
        gb = ctx.build_blas(gg, build_flags=bf)
        np.testing.assert_array_equal(gb.blob(), ob.blob())


def test_procedural_blas_update_refit(ctx, orc):
    aabbs = _aabbs(3000, 31)
    moved = aabbs.copy()
    rng = np.random.Generator(np.random.PCG64(5))
    shift = rng.uniform(-0.5, 0.5, size=(3000, 3)).astype(np.float32)
    moved[:, :3] += shift
    moved[:, 3:] += shift
    bf = T.BUILD_FLAG_ALLOW_UPDATE
    ob = orc.Blas([dict(aabbs=aabbs)], bf)
    gb = ctx.build_blas([_gpu_aabb_geom(ctx, aabbs)], build_flags=bf)
    np.testing.assert_array_equal(gb.blob(), ob.blob())
    ob.update([dict(aabbs=moved)])
    ctx.update_blas(gb, [_gpu_aabb_geom(ctx, moved)])
    np.testing.assert_array_equal(gb.blob(), ob.blob())


def _both_tlas(ctx, orc, ogeoms, ggeoms, transforms, hit_groups, flags=None, build_flags=0):
    ob = [orc.Blas(g, build_flags) for g in ogeoms]
    gb = [ctx.build_blas(g, build_flags=build_flags) for g in ggeoms]
    ot = orc.Tlas(ob, transforms, hit_groups=hit_groups, flags=flags)
    gt = ctx.build_tlas(gb, transforms, hit_groups=hit_groups, flags=flags)
    return ot, gt


def _assert_hits_equal(hg, ho):
    miss = ho["primitive_index"] == T.NO_HIT
    np.testing.assert_array_equal(hg["primitive_index"], ho["primitive_index"])
    np.testing.assert_array_equal(hg["t"], ho["t"])
    for f in ("bary", "instance_index", "geometry_index", "instance_id", "leaf_slot"):
        np.testing.assert_array_equal(hg[f][~miss], ho[f][~miss], err_msg=f)


PROGRAM_TABLES = [
    [[T.ANYHIT_NONE, T.INTERSECTION_NONE], [T.ANYHIT_NONE, T.INTERSECTION_SPHERE]],
    [[T.ANYHIT_CUTOUT, T.INTERSECTION_NONE], [T.ANYHIT_CUTOUT, T.INTERSECTION_SPHERE]],
    [[T.ANYHIT_IGNORE, T.INTERSECTION_NONE], [T.ANYHIT_ACCEPT, T.INTERSECTION_BOX]],
    [[T.ANYHIT_END_SEARCH, T.INTERSECTION_NONE], [T.ANYHIT_END_SEARCH, T.INTERSECTION_SPHERE]],
]
RAY_FLAG_SETS = [0, T.RAY_FLAG_FORCE_NON_OPAQUE, T.RAY_FLAG_FORCE_OPAQUE, T.RAY_FLAG_CULL_OPAQUE, T.RAY_FLAG_CULL_NON_OPAQUE,
                 T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH, T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | T.RAY_FLAG_FORCE_NON_OPAQUE,
                 T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES | T.RAY_FLAG_FORCE_NON_OPAQUE]


@pytest.mark.parametrize("table", range(len(PROGRAM_TABLES)))
def test_trace_hit_groups_parity_mixed_scene(table, ctx, orc):
    """Two instances of a mixed (triangles non-opaque + spheres) BLAS and one all-procedural BLAS, rigid transforms,
    every ray-flag set, hit-group records = geometry index + 2 * instance."""
    mesh = scenes.bunny_scale(3)
    aabbs = _aabbs(150, 8, lo=-3, hi=3, size=(0.4, 1.2))
    aabbs[:, 1] += 3.0
    aabbs[:, 4] += 3.0
    only = _aabbs(400, 9, lo=-4, hi=4, size=(0.3, 1.0))
    og = [[dict(vertices=mesh.vertices, stride=24, indices=mesh.indices, flags=T.GEOMETRY_FLAG_NONE), dict(aabbs=aabbs)],
          [dict(aabbs=only, flags=T.GEOMETRY_FLAG_NONE), dict(aabbs=only[:50] + np.float32(0.25))]]
    gg = [[_gpu_mesh_geom(ctx, mesh, T.GEOMETRY_FLAG_NONE), _gpu_aabb_geom(ctx, aabbs)],
          [_gpu_aabb_geom(ctx, only, T.GEOMETRY_FLAG_NONE), _gpu_aabb_geom(ctx, only[:50] + np.float32(0.25))]]
    c, s = np.float32(np.cos(0.7)), np.float32(np.sin(0.7))
    transforms = [scenes.IDENTITY_3X4, np.array([c, 0, s, 9, 0, 1, 0, 0.5, -s, 0, c, -2], np.float32),
                  np.array([1.5, 0, 0, -9, 0, 1.5, 0, 2, 0, 0, 1.5, 3], np.float32)]
    inst_flags = [0, T.INSTANCE_FLAG_FORCE_NON_OPAQUE, T.INSTANCE_FLAG_FORCE_OPAQUE]
    ob = [orc.Blas(g) for g in og]
    gb = [ctx.build_blas(g) for g in gg]
    ot = orc.Tlas([ob[0], ob[0], ob[1]], transforms, hit_groups=[0, 2, 4], flags=inst_flags)
    gt = ctx.build_tlas([gb[0], gb[0], gb[1]], transforms, hit_groups=[0, 2, 4], flags=inst_flags)
    np.testing.assert_array_equal(gt.blob(), _fix_blas_pointers(ot, gt))
    assert gt.info().has_procedural == 1
    progs = (PROGRAM_TABLES[table] * 3)[:6]
    rays = random_rays(30000, seed=17 + table, lo=(-14, -2, -8), hi=(14, 9, 8), tmin=1e-4)
    for flags in RAY_FLAG_SETS:
        ho = ot.trace_hit_groups(rays, progs, ray_flags=flags, geometry_multiplier=1, threads=8)
        hg = ctx.trace_hit_groups(gt, rays, progs, ray_flags=flags, geometry_multiplier=1)
        _assert_hits_equal(hg, ho)
        if table == 0 and flags == 0:
            hit = ho["primitive_index"] != T.NO_HIT
            kinds = ho["leaf_slot"][hit] >> 24
            assert (kinds == T.HIT_KIND_TRIANGLE_FRONT_FACE).sum() > 500 and (kinds <= 1).sum() > 500


def _fix_blas_pointers(ot, gt):
    """The oracle's TLAS blob with its host BLAS addresses replaced by the GPU's (the only bytes allowed to differ)."""
    o = ot.blob().copy()
    g = gt.blob()
    d = T.parse_tlas_blob(o)
    off = d["header"][1]
    md_o = o[off:off + 116 * d["n"]].view(T.BVH_METADATA_DTYPE)
    md_g = np.ascontiguousarray(g[off:off + 116 * d["n"]]).view(T.BVH_METADATA_DTYPE)
    md_o["blas"] = md_g["blas"]
    return o


def test_mask_and_contribution_indexing(ctx, orc):
    near_v, idx = ut_quad(depth=1.0)
    far_v, _ = ut_quad(depth=2.0)
    og = [[dict(vertices=near_v, stride=12, indices=idx, flags=T.GEOMETRY_FLAG_NONE),
           dict(vertices=far_v, stride=12, indices=idx, flags=T.GEOMETRY_FLAG_NONE)]]
    gg = [[dict(vertices=ctx.upload(near_v), vertex_count=4, stride=12, indices=ctx.upload(idx), index_count=6, index_format=16,
                flags=T.GEOMETRY_FLAG_NONE),
           dict(vertices=ctx.upload(far_v), vertex_count=4, stride=12, indices=ctx.upload(idx), index_count=6, index_format=16,
                flags=T.GEOMETRY_FLAG_NONE)]]
    ot, gt = _both_tlas(ctx, orc, og, gg, [scenes.IDENTITY_3X4], [3])
    rays = ut_rays()
    # record = ray_contribution(1) + geometry * multiplier(2) + instance contribution(3): near -> 4, far -> 6
    progs = np.zeros((8, 2), np.uint32)
    progs[4, 0] = T.ANYHIT_IGNORE
    ho = ot.trace_hit_groups(rays, progs, ray_contribution=1, geometry_multiplier=2)
    hg = ctx.trace_hit_groups(gt, rays, progs, ray_contribution=1, geometry_multiplier=2)
    _assert_hits_equal(hg, ho)
    np.testing.assert_array_equal(hg["t"], np.float32(2.0))
    progs[6, 0] = T.ANYHIT_IGNORE
    hg = ctx.trace_hit_groups(gt, rays, progs, ray_contribution=1, geometry_multiplier=2)
    assert (hg["primitive_index"] == T.NO_HIT).all()
    # records beyond the table have no programs; an empty table behaves like rt_trace_rays
    hg = ctx.trace_hit_groups(gt, rays, progs[:4], ray_contribution=1, geometry_multiplier=2)
    np.testing.assert_array_equal(hg["t"], np.float32(1.0))
    plain = ctx.trace(gt, rays)
    hg = ctx.trace_hit_groups(gt, rays, np.zeros((0, 2), np.uint32))
    np.testing.assert_array_equal(hg["t"], plain["t"])
    np.testing.assert_array_equal(hg["leaf_slot"] & 0xFFFFFF, plain["leaf_slot"])
    # instance mask
    hg = ctx.trace_hit_groups(gt, rays, progs, mask=0x00)
    assert (hg["primitive_index"] == T.NO_HIT).all()


def test_unknown_program_is_rejected(ctx):
    aabbs = _aabbs(10, 1)
    gb = ctx.build_blas([_gpu_aabb_geom(ctx, aabbs)])
    gt = ctx.build_tlas([gb], [scenes.IDENTITY_3X4])
    rays = random_rays(8, seed=1, lo=(-1, -1, -1), hi=(1, 1, 1))
    with pytest.raises(rt.RtError):
        ctx.trace_hit_groups(gt, rays, [[0, 99]])
    with pytest.raises(rt.RtError):
        ctx.trace_hit_groups(gt, rays, [[99, 0]])
    with pytest.raises(rt.RtError):  # unrecognized geometry type: E_INVALIDARG (FL/LoadPrimitivesPass.cpp:124-127)
        g = _gpu_aabb_geom(ctx, aabbs)
        descs = rt._geometry_descs([g])
        descs[0].type = 7
        info = T.PrebuildInfo()
        rt.check(rt.lib.rt_blas_prebuild(ctx.handle, descs, 1, 0, rt.C.byref(info)))
        scratch, result = ctx.alloc(info.scratch_bytes), ctx.alloc(info.result_bytes)
        rt.check(rt.lib.rt_blas_build(ctx.handle, descs, 1, 0, scratch.ptr, scratch.nbytes, result.ptr, result.nbytes))


def test_triangle_pipelines_refuse_procedural_scenes(ctx):
    """The application's hit groups are of type TRIANGLES: a TLAS that reaches procedural primitives is a hit-group
    type mismatch — every ray misses and the status check fails loudly instead of shading AABB corners as triangles."""
    aabbs = _aabbs(100, 2, lo=-2, hi=2)
    gb = ctx.build_blas([_gpu_aabb_geom(ctx, aabbs)])
    gt = ctx.build_tlas([gb], [scenes.IDENTITY_3X4])
    rays = random_rays(2000, seed=3, lo=(-3, -3, -3), hi=(3, 3, 3))
    with pytest.raises(rt.RtError) as e:
        ctx.trace(gt, rays)
    assert e.value.code == -4  # RT_ERR_UNSUPPORTED
    ctx.status()  # not sticky
    hits = ctx.trace_hit_groups(gt, rays, [[0, T.INTERSECTION_BOX]])
    assert (hits["primitive_index"] != T.NO_HIT).sum() > 100


def test_large_procedural_scene_properties(ctx, orc):
    """1 M spheres: properties that do not need the oracle at full size, plus a sampled bit-exact comparison."""
    n = 1_000_000
    aabbs = _aabbs(n, 77, lo=-100, hi=100, size=(0.5, 2.0))
    gb = ctx.build_blas([_gpu_aabb_geom(ctx, aabbs)], build_flags=T.BUILD_FLAG_PREFER_FAST_BUILD)
    gt = ctx.build_tlas([gb], [scenes.IDENTITY_3X4])
    rays = random_rays(200000, seed=5, lo=(-100, -100, -100), hi=(100, 100, 100), tmin=1e-3)
    prog = [[T.ANYHIT_NONE, T.INTERSECTION_SPHERE]]
    h = ctx.trace_hit_groups(gt, rays, prog)
    hit = h["primitive_index"] != T.NO_HIT
    assert hit.mean() > 0.3
    # every reported hit lies on the sphere it names (|p - c| = r) and inside its AABB
    p = rays["origin"][hit].astype(np.float64) + rays["direction"][hit].astype(np.float64) * h["t"][hit][:, None]
    bb = aabbs[h["primitive_index"][hit]].astype(np.float64)
    c = (bb[:, :3] + bb[:, 3:]) * 0.5
    r = (bb[:, 3:] - c).min(axis=1)
    np.testing.assert_allclose(np.linalg.norm(p - c, axis=1), r, rtol=0, atol=5e-3)
    # any-hit visibility agrees with closest-hit visibility
    v = ctx.trace_hit_groups(gt, rays, prog, ray_flags=T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH)
    np.testing.assert_array_equal(v["primitive_index"] != T.NO_HIT, hit)
    # shortening the ray to just before its hit turns it into a miss
    short = rays.copy()
    short["tmax"] = np.where(hit, h["t"] * np.float32(0.999), rays["tmax"])
    s = ctx.trace_hit_groups(gt, short, prog)
    assert (s["primitive_index"][hit] == T.NO_HIT).mean() > 0.999
    # sampled oracle comparison on a 50 k subset of the same primitives
    sub = aabbs[:50000]
    ot = orc.Tlas([orc.Blas([dict(aabbs=sub)], T.BUILD_FLAG_PREFER_FAST_BUILD)], [scenes.IDENTITY_3X4])
    gs = ctx.build_tlas([ctx.build_blas([_gpu_aabb_geom(ctx, sub)], build_flags=T.BUILD_FLAG_PREFER_FAST_BUILD)], [scenes.IDENTITY_3X4])
    _assert_hits_equal(ctx.trace_hit_groups(gs, rays[:50000], prog), ot.trace_hit_groups(rays[:50000], prog, threads=8))
