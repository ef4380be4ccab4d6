"""GPU parity: two-level BVH2 traversal (rt_trace_rays through the C ABI) against the CPU oracle, plus the
reference's own TracingTests known answers (UT:3889-4078) run on the GPU path.

North-star criterion 2: primary-ray hit triangle IDs agree on >= 99.99 % of rays.
"""
import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T
from helpers import partition_transform, random_rays, ut_quad, ut_rays, bunny_case, cornell_case, two_material_case

pytestmark = pytest.mark.gpu


def _agreement(g, o):
    same = (g["primitive_index"] == o["primitive_index"]) & (g["instance_index"] == o["instance_index"]) | \
           ((g["primitive_index"] == T.NO_HIT) & (o["primitive_index"] == T.NO_HIT))
    return float(same.mean()), same


def _build_both(case, ctx, orc):
    otlas, _ = case.oracle(orc)
    blases = [ctx.build_blas_from_mesh(m) for m in case.meshes]
    gtlas = ctx.build_tlas(blases, case.transforms)
    return otlas, gtlas


@pytest.mark.parametrize("case_name,w,h", [("cornell", 256, 256), ("bunny", 480, 270), ("two", 320, 180)])
def test_primary_hit_ids(case_name, w, h, ctx, orc):
    case = {"cornell": cornell_case, "bunny": bunny_case, "two": two_material_case}[case_name]()
    otlas, gtlas = _build_both(case, ctx, orc)
    frame = scenes.make_frame(case.setup, w, h, 0, 0, jitter=(0.3 / w, -0.2 / h))
    rays_o = orc.primary_rays(frame, w, h, 30.0)
    rays_g = ctx.primary_rays(frame, w, h, 30.0)
    np.testing.assert_array_equal(rays_g.view(np.uint8), rays_o.view(np.uint8))  # RayGen is bit-exact
    flags = T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES
    ho = otlas.trace(rays_o, flags, threads=8)
    hg = ctx.trace(gtlas, rays_g, flags)
    frac, same = _agreement(hg, ho)
    assert frac >= 0.9999, f"hit-ID agreement {frac}"
    hit = same & (ho["primitive_index"] != T.NO_HIT)
    assert hit.sum() > 0.2 * w * h
    # where the IDs agree the hit itself is bit-identical (same unfused Woop arithmetic)
    np.testing.assert_array_equal(hg["t"][hit], ho["t"][hit])
    np.testing.assert_array_equal(hg["bary"][hit], ho["bary"][hit])
    np.testing.assert_array_equal(hg["leaf_slot"][hit], ho["leaf_slot"][hit])


@pytest.mark.parametrize("flags", [0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES,
                                   T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | T.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER])
def test_incoherent_rays_and_work_counters(flags, ctx, orc):
    case = bunny_case(4)
    otlas, gtlas = _build_both(case, ctx, orc)
    rays = random_rays(20000, seed=11, lo=(-8, 0.1, -8), hi=(8, 10, 8), tmin=1e-4)
    st_o = T.TraceStats()
    ho = otlas.trace(rays, flags, threads=8, stats=st_o)
    hg, st_g = ctx.trace(gtlas, rays, flags, stats=True)
    any_hit = bool(flags & T.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH)
    if any_hit:  # visibility must agree; which triangle terminated the search is traversal-order defined and equal too
        np.testing.assert_array_equal(hg["primitive_index"] == T.NO_HIT, ho["primitive_index"] == T.NO_HIT)
    frac, same = _agreement(hg, ho)
    assert frac >= 0.9999
    np.testing.assert_array_equal(hg["t"][same], ho["t"][same])
    # identical traversal order => identical work (this is what the roofline's bytes/ray are computed from)
    assert int(st_g[0]) == st_o.rays == rays.shape[0]
    assert int(st_g[1]) == st_o.internal_visits
    assert int(st_g[2]) == st_o.leaf_visits
    assert int(st_g[3]) == st_o.instance_visits
    assert int(st_g[4]) <= 64 and st_o.max_stack <= 64
    # the production kernel (persistent warps over 4-wide nodes) visits nodes in its own order: the closest hit is
    # the same triangle with the same bits; for any-hit rays only the visibility is order independent
    hp = ctx.trace(gtlas, rays, flags)
    if any_hit:
        np.testing.assert_array_equal(hp["primitive_index"] == T.NO_HIT, ho["primitive_index"] == T.NO_HIT)
    else:
        frac, same = _agreement(hp, ho)
        assert frac >= 0.9999
        np.testing.assert_array_equal(hp["t"][same], ho["t"][same])
        np.testing.assert_array_equal(hp["bary"][same], ho["bary"][same])
        np.testing.assert_array_equal(hp["leaf_slot"][same], ho["leaf_slot"][same])
        np.testing.assert_array_equal(hp["instance_id"][same], ho["instance_id"][same])
        np.testing.assert_array_equal(hp["geometry_index"][same], ho["geometry_index"][same])


def _ut_scene(ctx, orc, specs):
    """specs: list of (kind, winding, transform, flags, mask) -> (oracle tlas, gpu tlas)."""
    ob, gb, xf, fl, mk = [], [], [], [], []
    for kind, winding, tr, flags, mask in specs:
        verts, idx = ut_quad(kind, winding)
        ob.append(orc.Blas([dict(vertices=verts, stride=12, indices=idx, flags=T.GEOMETRY_FLAG_OPAQUE)]))
        dv, di = ctx.upload(verts), ctx.upload(idx)
        gb.append(ctx.build_blas([dict(vertices=dv, vertex_count=4, stride=12, indices=di, index_count=6, index_format=16,
                                       flags=T.GEOMETRY_FLAG_OPAQUE)]))
        xf.append(tr), fl.append(flags), mk.append(mask)
    return orc.Tlas(ob, xf, masks=mk, flags=fl), ctx.build_tlas(gb, xf, masks=mk, flags=fl)


def _hit_grid(hits, w=6, h=4):
    return (hits["primitive_index"] != T.NO_HIT).reshape(h, w)


IDENT = scenes.IDENTITY_3X4


@pytest.mark.parametrize("transform,expect", [
    (IDENT, "left"),                                                         # BasicTrace
    (np.array([-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), "right"),  # BasicTraceWithInstanceFlip
    (np.array([np.cos(1.57), np.sin(1.57), 0, 0, -np.sin(1.57), np.cos(1.57), 0, 0, 0, 0, 1, 0], np.float32), "bottom"),
    (np.array([1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), "right"),   # BasicTraceWithInstanceTranslation
])
def test_reference_basic_trace_known_answers(transform, expect, ctx, orc):
    """UT:3889-3933: a left-half-screen quad under an instance transform, 6x4 rays along +z."""
    ot, gt = _ut_scene(ctx, orc, [("left", "cw", transform, 0, 0xFF)])
    rays = ut_rays()
    x = np.arange(6)[None, :].repeat(4, 0)
    y = np.arange(4)[:, None].repeat(6, 1)
    want = {"left": x < 3, "right": x >= 3, "bottom": y >= 2}[expect]
    for hits in (ot.trace(rays, 0), ctx.trace(gt, rays, 0)):
        np.testing.assert_array_equal(_hit_grid(hits), want)


@pytest.mark.parametrize("cull", [0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES])
def test_reference_culling_known_answers(cull, ctx, orc):
    """TestCulling (UT:3952-3998): six screen strips {CCW x3, CW x3} x {NONE, FRONT_CCW, CULL_DISABLE}."""
    specs, want = [], []
    for i in range(6):
        winding = "ccw" if i < 3 else "cw"
        iflag = [T.INSTANCE_FLAG_NONE, T.INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE, T.INSTANCE_FLAG_TRIANGLE_CULL_DISABLE][i % 3]
        specs.append(("full", winding, partition_transform(i, 6), iflag, 0xFF))
        front = (winding == "cw" and not (iflag & T.INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE)) or \
                (winding == "ccw" and (iflag & T.INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE))
        if iflag == T.INSTANCE_FLAG_TRIANGLE_CULL_DISABLE or cull == 0:
            want.append(True)
        elif front:
            want.append(not (cull & T.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES))
        else:
            want.append(not (cull & T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES))
    ot, gt = _ut_scene(ctx, orc, specs)
    rays = ut_rays()
    expect = np.array(want)[None, :].repeat(4, 0)
    for hits in (ot.trace(rays, cull), ctx.trace(gt, rays, cull)):
        np.testing.assert_array_equal(_hit_grid(hits), expect)


def test_reference_instance_masks_known_answers(ctx, orc):
    """TraceInstanceMasks (UT:4027-4070): instance masks 1<<i against trace mask 0x23."""
    specs = [("full", "cw", partition_transform(i, 6), 0, 1 << i) for i in range(6)]
    ot, gt = _ut_scene(ctx, orc, specs)
    rays = ut_rays()
    expect = np.array([bool((1 << i) & 0x23) for i in range(6)])[None, :].repeat(4, 0)
    for hits in (ot.trace(rays, 0, mask=0x23), ctx.trace(gt, rays, 0, mask=0x23)):
        np.testing.assert_array_equal(_hit_grid(hits), expect)
    hits = ctx.trace(gt, rays, 0, mask=0x23)
    hit = hits["primitive_index"] != T.NO_HIT
    cols = np.arange(24) % 6
    np.testing.assert_array_equal(hits["instance_index"][hit], cols[hit])  # InstanceIndex() of each strip


def test_instanced_scene_hit_ids(ctx, orc):
    """TLAS over many transformed instances of one BLAS (config C4's shape, small): IDs and instance indices."""
    mesh = scenes.icosphere(3)
    n_inst = 200
    xf = scenes.random_rigid_transforms(n_inst, seed=10, extent=30.0)
    ob = orc.Blas.from_mesh(mesh)
    ot = orc.Tlas([ob] * n_inst, xf)
    gb = ctx.build_blas_from_mesh(mesh)
    gt = ctx.build_tlas([gb] * n_inst, xf)
    rays = random_rays(30000, seed=3, lo=(-35, -35, -35), hi=(35, 35, 35))
    ho = ot.trace(rays, 0, threads=8)
    hg = ctx.trace(gt, rays, 0)
    frac, same = _agreement(hg, ho)
    assert frac >= 0.9999
    assert (ho["primitive_index"] != T.NO_HIT).mean() > 0.05
    np.testing.assert_array_equal(hg["t"][same], ho["t"][same])
    np.testing.assert_array_equal(hg["instance_id"][same], ho["instance_id"][same])


def test_axis_aligned_rays_same_hits_without_walking_the_tree(ctx, orc):
    """Direction components that are exactly 0 turn the reference's slab test into NaN, which passes EVERY box: the
    reference (and the oracle, and the instrumented kernel) then walk the whole tree.  The production kernel evaluates
    the exact slab condition for such an axis instead (trace.cuh: ray_pre_box<ROBUST>); the hits must not change."""
    case = bunny_case(4)
    otlas, gtlas = _build_both(case, ctx, orc)
    rng = np.random.default_rng(5)
    n = 3000
    rays = np.zeros(n, T.RAY_DTYPE)
    rays["origin"] = rng.uniform((-8, -0.5, -8), (8, 14, 8), size=(n, 3)).astype(np.float32)
    axis = rng.integers(0, 3, size=n)
    d = np.zeros((n, 3), np.float32)
    d[np.arange(n), axis] = rng.choice([-1.0, 1.0], size=n)
    two = rng.random(n) < 0.5   # half of the rays keep one more non-zero component (a single zero component)
    d[two, (axis[two] + 1) % 3] = rng.uniform(-1, 1, size=int(two.sum())).astype(np.float32)
    d[rng.random(n) < 0.2] *= np.float32(-0.0) + 1  # keep; negative zeros appear through the sign choice above
    rays["direction"] = d
    rays["tmin"], rays["tmax"] = 1e-4, 1e38
    for flags in (0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES):
        st_o = T.TraceStats()
        ho = otlas.trace(rays, flags, threads=8, stats=st_o)
        hp = ctx.trace(gtlas, rays, flags)
        frac, same = _agreement(hp, ho)
        assert frac >= 0.9999, frac
        np.testing.assert_array_equal(hp["t"][same], ho["t"][same])
        np.testing.assert_array_equal(hp["bary"][same], ho["bary"][same])
        assert (ho["primitive_index"] != T.NO_HIT).sum() > n // 10
        assert st_o.internal_visits > 50 * n  # the literal arithmetic really does walk large parts of the tree
