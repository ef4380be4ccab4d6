"""rt_set_render_options: the radiance-depth knob (MAX_RADIANCE_RAY_DEPTH 1 -> 2) and R16G16B16A16_FLOAT target emulation.

Depth 1 / fp32 is the default everywhere and the only setting the parity criteria are stated for; here the oracle follows the
same shader source with the constant changed (oracle.set_render_options) and the CUDA path must agree with it.
"""
import numpy as np
import pytest

from conftest import rel_rmse
from dxrexperiments_b200 import scenes, types as T

from helpers import bunny_case, cornell_case, two_material_case


@pytest.fixture
def options(orc):
    yield orc.set_render_options
    orc.set_render_options(1, False)


def _oracle_frames(orc, case, w, h, spp, realtime=False):
    tlas, recs = case.oracle(orc)
    jit = scenes.jitter_sequence(3, spp, w, h)
    if realtime:
        return orc.render_realtime(tlas, recs, case.env, scenes.make_frame(case.setup, w, h, 0, 0, jitter=jit[0]), w, h, threads=4)
    acc = np.zeros((h, w, 4), np.float32)
    for s in range(spp):
        orc.render_progressive(tlas, recs, case.env, scenes.make_frame(case.setup, w, h, s, s, jitter=jit[s]), w, h, acc, threads=4)
    return acc


def test_oracle_depth_two_adds_the_second_reflection_bounce(orc, options):
    case = cornell_case()
    w = h = 40
    a1 = _oracle_frames(orc, case, w, h, 2).copy()
    options(2, False)
    c = T.RayCounts()
    tlas, recs = case.oracle(orc)
    f = scenes.make_frame(case.setup, w, h, 0, 0)
    acc = np.zeros((h, w, 4), np.float32)
    orc.render_progressive(tlas, recs, case.env, f, w, h, acc, threads=1, counts=c)
    a2 = _oracle_frames(orc, case, w, h, 2)
    assert c.secondary > 2 * (acc[..., 3] > 0).sum() * 0.5  # more than the two depth-0 bounces: the depth-1 lobe rays are traced now
    assert np.isfinite(a2).all() and (a2[..., :3] >= a1[..., :3] - 1e-6).all()  # a non-negative term was added
    assert rel_rmse(a2[..., :3], a1[..., :3]) > 1e-3
    options(1, False)
    np.testing.assert_array_equal(_oracle_frames(orc, case, w, h, 2), a1)  # the knob returns to the reference's value


def test_oracle_half_targets_hold_fp16_values(orc, options):
    case = cornell_case()
    w = h = 32
    full = _oracle_frames(orc, case, w, h, 3).copy()
    options(1, True)
    half = _oracle_frames(orc, case, w, h, 3)
    np.testing.assert_array_equal(half, half.astype(np.float16).astype(np.float32))  # every stored value is an fp16 value
    assert 0 < rel_rmse(half[..., :3], full[..., :3]) < 2e-3
    d, s = _oracle_frames(orc, case, w, h, 1, realtime=True)
    np.testing.assert_array_equal(d, d.astype(np.float16).astype(np.float32))
    out, tmp = orc.denoise(d, s, T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0))
    np.testing.assert_array_equal(tmp[..., :3], tmp[..., :3].astype(np.float16).astype(np.float32))
    np.testing.assert_array_equal(out[..., :3], out[..., :3].astype(np.float16).astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("case_fn,w,h,spp", [(cornell_case, 128, 128, 3), (bunny_case, 240, 136, 2), (two_material_case, 200, 120, 2)])
def test_cuda_depth_two_matches_oracle(case_fn, w, h, spp, ctx, rt, orc, options):
    case = case_fn()
    options(2, False)
    ctx.set_render_options(2, False)
    try:
        want = _oracle_frames(orc, case, w, h, spp)
        r = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
        jit = scenes.jitter_sequence(3, spp, w, h)
        ctx.ray_counts(reset=True)
        for s in range(spp):
            r.dispatch(scenes.make_frame(case.setup, w, h, s, s, jitter=jit[s]))
        ctx.status()
        got = r.image(0)
        assert rel_rmse(got[..., :3], want[..., :3]) <= 1e-3
        # and the knob really changed the frame
        ctx.set_render_options(1, False)
        r1 = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
        for s in range(spp):
            r1.dispatch(scenes.make_frame(case.setup, w, h, s, s, jitter=jit[s]))
        assert rel_rmse(got[..., :3], r1.image(0)[..., :3]) > 1e-3
        # ray counts: the oracle's counters agree on the number of traced secondary rays at depth 2
        ctx.set_render_options(2, False)
        tlas, recs = case.oracle(orc)
        oc = T.RayCounts()
        f0 = scenes.make_frame(case.setup, w, h, 0, 0, jitter=jit[0])
        orc.render_progressive(tlas, recs, case.env, f0, w, h, np.zeros((h, w, 4), np.float32), threads=4, counts=oc)
        ctx.ray_counts(reset=True)
        case.renderer(rt, ctx, rt.PROGRESSIVE, w, h).dispatch(f0)
        gc = ctx.ray_counts(reset=True)
        assert gc.primary == oc.primary
        assert abs(gc.secondary - oc.secondary) <= 2e-3 * oc.secondary and abs(gc.shadow - oc.shadow) <= 2e-3 * oc.shadow
    finally:
        ctx.set_render_options(1, False)


@pytest.mark.gpu
def test_cuda_depth_two_realtime_and_two_band_dispatch(ctx, rt, orc, options):
    case = bunny_case(4)
    w, h = 640, 512  # two pixel bands
    options(2, False)
    ctx.set_render_options(2, False)
    try:
        f = scenes.make_frame(case.setup, w, h, 1, 0, jitter=(0.1 / w, -0.2 / h))
        tlas, recs = case.oracle(orc)
        d, s = orc.render_realtime(tlas, recs, case.env, f, w, h, threads=8)
        r = case.renderer(rt, ctx, rt.REALTIME, w, h)
        r.dispatch(f)
        ctx.status()
        assert rel_rmse(r.image(0)[..., :3], d[..., :3]) <= 1e-3
        assert rel_rmse(r.image(1)[..., :3], s[..., :3]) <= 1e-3
    finally:
        ctx.set_render_options(1, False)


@pytest.mark.gpu
def test_cuda_half_render_targets_match_oracle(ctx, rt, orc, options):
    case = cornell_case()
    w = h = 96
    options(1, True)
    ctx.set_render_options(1, True)
    try:
        want = _oracle_frames(orc, case, w, h, 4)
        r = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
        jit = scenes.jitter_sequence(3, 4, w, h)
        for s in range(4):
            r.dispatch(scenes.make_frame(case.setup, w, h, s, s, jitter=jit[s]))
        got = r.image(0)
        np.testing.assert_array_equal(got, got.astype(np.float16).astype(np.float32))
        assert rel_rmse(got[..., :3], want[..., :3]) <= 1e-3
        # one fp16 ulp at most wherever the two differ (the fp32 values feeding the rounding differ by ~1e-7)
        diff = np.abs(got[..., :3] - want[..., :3])
        assert (diff <= np.maximum(np.abs(want[..., :3]), 6.2e-5) * 2.0 ** -9).all()
        d, s = _oracle_frames(orc, case, w, h, 1, realtime=True)
        out, tmp = ctx.denoise(d, s, T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0))
        wout, wtmp = orc.denoise(d, s, T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0))
        np.testing.assert_array_equal(out[..., :3], out[..., :3].astype(np.float16).astype(np.float32))
        assert rel_rmse(out[..., :3], wout[..., :3]) <= 1e-3 and rel_rmse(tmp[..., :3], wtmp[..., :3]) <= 1e-3
    finally:
        ctx.set_render_options(1, False)
    with pytest.raises(rt.RtError):
        ctx.set_render_options(3, False)
