"""GPU parity: the wavefront pipelines (rt_dispatch_rays) and the denoiser (rt_denoise) against the oracle.

North-star criterion 3: accumulated images after N spp are within 1e-3 relative RMSE of the oracle's.
The tolerance is written where it is checked (REL_RMSE_TOL).
"""
import numpy as np
import pytest

from conftest import rel_rmse
from dxrexperiments_b200 import scenes, types as T
from helpers import bunny_case, cornell_case, two_material_case

pytestmark = pytest.mark.gpu

REL_RMSE_TOL = 1e-3


def _render_both(case, ctx, rt, orc, kind, w, h, spp, options=None, seed=77):
    otlas, recs = case.oracle(orc)
    r = case.renderer(rt, ctx, kind, w, h)
    jit = scenes.jitter_sequence(seed, spp, w, h)
    acc = np.zeros((h, w, 4), np.float32)
    oc, aovs = T.RayCounts(), None
    ctx.ray_counts(reset=True)
    for s in range(spp):
        frame = scenes.make_frame(case.setup, w, h, frame_count=s, accum_count=s if kind == rt.PROGRESSIVE else 0,
                                  jitter=jit[s], options=options)
        if kind == rt.PROGRESSIVE:
            orc.render_progressive(otlas, recs, case.env, frame, w, h, acc, threads=8, counts=oc)
        else:
            aovs = orc.render_realtime(otlas, recs, case.env, frame, w, h, threads=8, counts=oc)
        r.dispatch(frame)
    ctx.status()
    gc = ctx.ray_counts()
    if kind == rt.PROGRESSIVE:
        return r.image(0), acc, gc, oc
    return (r.image(0), r.image(1)), aovs, gc, oc


@pytest.mark.parametrize("case_fn,w,h,spp", [(cornell_case, 256, 256, 1), (cornell_case, 128, 128, 8),
                                             (bunny_case, 320, 180, 4), (two_material_case, 240, 136, 4)])
def test_progressive_accumulation_matches_oracle(case_fn, w, h, spp, ctx, rt, orc):
    img, ref, gc, oc = _render_both(case_fn(), ctx, rt, orc, rt.PROGRESSIVE, w, h, spp)
    assert np.isfinite(img).all()
    assert (img[..., 3] == 1.0).all()
    err = rel_rmse(img[..., :3], ref[..., :3])
    assert err <= REL_RMSE_TOL, f"relative RMSE {err}"
    # the same rays were traced: primary exactly; secondary/shadow within the handful of pixels whose primary hit differs
    assert gc.primary == oc.primary == w * h * spp
    assert abs(int(gc.secondary) - int(oc.secondary)) <= 1e-4 * oc.secondary + 4
    assert abs(int(gc.shadow) - int(oc.shadow)) <= 1e-3 * oc.shadow + 8


OPTION_CASES = {
    "uniform_hemisphere": dict(cosineHemisphereSampling=0),
    "no_indirect_diffuse": dict(noIndirectDiffuse=1),
    "indirect_diffuse_only": dict(showIndirectDiffuseOnly=1),
    "indirect_specular_only": dict(showIndirectSpecularOnly=1),
    "ambient_occlusion_only": dict(showAmbientOcclusionOnly=1),
    "ambient_occlusion_uniform": dict(showAmbientOcclusionOnly=1, cosineHemisphereSampling=0),
    "albedo_only": dict(showGBufferAlbedoOnly=1),
    "direct_only": dict(showDirectLightingOnly=1),
    "fresnel_only": dict(showFresnelTerm=1),
    "one_light_sampling": dict(debug=2),
    "environment_strength": dict(environmentStrength=2.5),
}


@pytest.mark.parametrize("name", sorted(OPTION_CASES))
def test_progressive_debug_options(name, ctx, rt, orc):
    """Every DebugOptions switch of shade() (ProgressiveRaytracing.hlsl:80-148) changes both sides alike."""
    opt = scenes.default_options()
    for k, v in OPTION_CASES[name].items():
        setattr(opt, k, v)
    img, ref, _, _ = _render_both(cornell_case(), ctx, rt, orc, rt.PROGRESSIVE, 96, 96, 2, options=opt)
    err = rel_rmse(img[..., :3], ref[..., :3])
    assert err <= REL_RMSE_TOL, f"{name}: relative RMSE {err}"


def test_max_iterations_early_out(ctx, rt, orc):
    """RayGen returns without touching gOutput once accumCount >= maxIterations (ProgressiveRaytracing.hlsl:13-15)."""
    case = cornell_case()
    r = case.renderer(rt, ctx, rt.PROGRESSIVE, 64, 64)
    opt = scenes.default_options()
    opt.maxIterations = 1
    r.dispatch(scenes.make_frame(case.setup, 64, 64, 0, 0, options=opt))
    first = r.image(0).copy()
    r.dispatch(scenes.make_frame(case.setup, 64, 64, 1, 1, options=opt))
    np.testing.assert_array_equal(r.image(0), first)


def test_region_dispatch_tiles_equal_full_frame(ctx, rt, orc):
    """Screen-tile sharding: rendering the frame as four rectangles gives the full-frame bits."""
    case = cornell_case()
    w, h = 100, 76
    full = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
    tiled = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
    for s in range(2):
        frame = scenes.make_frame(case.setup, w, h, s, s, jitter=(0.001 * s, 0.0))
        full.dispatch(frame)
        for reg in [(0, 0, 37, 40), (37, 0, w, 40), (0, 40, 64, h), (64, 40, w, h)]:
            tiled.dispatch(frame, region=reg)
    np.testing.assert_array_equal(tiled.image(0), full.image(0))


@pytest.mark.parametrize("case_fn,w,h", [(cornell_case, 160, 120), (two_material_case, 192, 108)])
def test_realtime_aovs_and_denoise_match_oracle(case_fn, w, h, ctx, rt, orc):
    (direct, spec), (odirect, ospec), gc, oc = _render_both(case_fn(), ctx, rt, orc, rt.REALTIME, w, h, 1)
    assert rel_rmse(direct[..., :3], odirect[..., :3]) <= REL_RMSE_TOL
    assert rel_rmse(spec[..., :3], ospec[..., :3]) <= REL_RMSE_TOL
    assert gc.primary == oc.primary
    prm = T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0)  # src/DenoiseCompositor.cpp:45-50
    gout, gtmp = ctx.denoise(direct, spec, prm)
    oout, otmp = orc.denoise(direct, spec, prm, threads=8)
    # same inputs on both sides: the filter itself is the same sequence of IEEE operations
    np.testing.assert_allclose(gtmp, otmp, rtol=0, atol=0)
    np.testing.assert_allclose(gout, oout, rtol=0, atol=0)


@pytest.mark.parametrize("world,k", [(2, 12), (3, 12), (5, 20), (4, 3)])
def test_realtime_frame_sharded_by_row_bands_equals_the_full_frame(world, k, ctx, rt):
    """the realtime pipeline + DenoiseCompositor sharded by screen bands with the filter's reach as halo.  Every
    "rank" renders rows [r0, r1) of the AOVs, filters exactly those rows and keeps its core rows; the sum of the ranks'
    buffers (what rt_accum_reduce with weight 1 computes) is the full-frame result, bit for bit."""
    from dxrexperiments_b200 import sharding
    case = two_material_case()
    w, h = 200, 123  # ragged against the tile sizes of both filter passes
    f = scenes.make_frame(case.setup, w, h, 0, 0, jitter=(0.25, -0.15))
    prm = T.DenoiserParams(1.0, 2.2, 1, 0, k, 0)
    full = case.renderer(rt, ctx, rt.REALTIME, w, h)
    full.dispatch(f)
    ref, _ = ctx.denoise(full.image(0), full.image(1), prm)
    total = np.zeros((h, w, 4), np.float32)
    for rank in range(world):
        band = sharding.band_plan(rank, world, h, halo=k)
        r = case.renderer(rt, ctx, rt.REALTIME, w, h)
        tmp, final = ctx.alloc(16 * w * h).zero(), ctx.alloc(16 * w * h).zero()
        r.realtime_band(f, band, prm, tmp, final)
        img = final.download(np.float32).reshape(h, w, 4)
        assert not img[: band.y0].any() and not img[band.y1:].any()  # nothing outside the core rows
        total += img
    np.testing.assert_array_equal(total, ref)
    ctx.status()


@pytest.mark.parametrize("k,tonemap,gamma,dbg", [(12, 1, 0, 0), (1, 0, 0, 1), (20, 1, 1, 0), (25, 1, 0, 0), (7, 0, 1, 2), (5, 1, 0, 3)])
@pytest.mark.parametrize("w,h", [(64, 64), (130, 71), (1922 // 8, 1126 // 8)])
def test_denoise_parameters_and_ragged_sizes(k, tonemap, gamma, dbg, w, h, ctx, orc):
    rng = np.random.Generator(np.random.PCG64(k * 1000 + w))
    direct = rng.random((h, w, 4), dtype=np.float32)
    direct[:, : w // 2, :3] *= 0.05  # an edge for the range weight
    spec = (rng.random((h, w, 4), dtype=np.float32) ** 3).astype(np.float32)
    prm = T.DenoiserParams(1.3, 2.2, tonemap, gamma, k, dbg)
    gout, gtmp = ctx.denoise(direct, spec, prm)
    oout, otmp = orc.denoise(direct, spec, prm, threads=8)
    np.testing.assert_array_equal(gtmp, otmp)
    if gamma:  # powf differs by a few ulp between CUDA and glibc
        np.testing.assert_allclose(gout, oout, rtol=2e-6, atol=1e-7)
    else:
        np.testing.assert_array_equal(gout, oout)
