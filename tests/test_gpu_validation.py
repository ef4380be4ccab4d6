"""Argument and state validation at the C ABI that round 2 added: bad instances no longer fault the context, holes in
the hit-record table and out-of-range record indices are reported, the realtime program's second output pitch is checked, a TLAS
refit forgets procedural BLASes that were swapped out, and the denoiser works on a second device of the same process."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T

from helpers import bunny_case, random_rays

pytestmark = pytest.mark.gpu


def _tlas_from_descs(ctx, rt, descs: np.ndarray, n: int, flags=0, result=None):
    info = T.PrebuildInfo()
    rt.check(rt.lib.rt_tlas_prebuild(ctx.handle, n, flags, C.byref(info)))
    scratch = ctx.alloc(info.scratch_bytes)
    result = result or ctx.alloc(info.result_bytes)
    dev = ctx.upload(descs)
    rt.check(rt.lib.rt_tlas_build(ctx.handle, dev.ptr, n, flags, scratch.ptr, scratch.nbytes, result.ptr, result.nbytes))
    ctx.sync()
    return result, dev


def test_instance_with_null_or_foreign_blas_is_inactive_not_fatal(ctx, rt):
    mesh = scenes.icosphere(2)
    blas = ctx.build_blas_from_mesh(mesh)
    junk = ctx.alloc(4096).zero()  # 64-byte aligned device memory that is not an acceleration structure
    xf = [scenes.IDENTITY_3X4, np.array([1, 0, 0, 5, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), np.array([1, 0, 0, -5, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)]

    class Fake:  # stands in for an Accel in _instance_descs_bytes
        def __init__(self, ptr):
            self.result = type("R", (), {"ptr": ptr})()

    descs = rt._instance_descs_bytes([blas, Fake(0), Fake(junk.ptr)], xf)
    tlas, _keep = _tlas_from_descs(ctx, rt, descs, 3)
    with pytest.raises(rt.RtError) as e:
        ctx.status()
    assert e.value.code == -1 and "BLAS" in str(e.value)  # RT_ERR_INVALID_ARG, reported once
    ctx.status()
    # the good instance is still traceable; rays towards the two bad ones miss
    rays = np.zeros(3, T.RAY_DTYPE)
    rays["origin"] = [(0, 0, -4), (5, 0, -4), (-5, 0, -4)]
    rays["direction"] = (0, 0, 1)
    rays["tmax"] = 100.0
    acc = rt.Accel(ctx, tlas, 3, top=True)
    hits = ctx.trace(acc, rays)
    assert hits["primitive_index"][0] != T.NO_HIT and hits["instance_index"][0] == 0
    assert (hits["primitive_index"][1:] == T.NO_HIT).all()


def test_tlas_refit_forgets_a_procedural_blas_that_was_swapped_out(ctx, rt):
    rng = np.random.Generator(np.random.PCG64(4))
    lo = rng.uniform(-1, 1, size=(50, 3)).astype(np.float32)
    aabbs = np.concatenate([lo, lo + 0.2], axis=1).astype(np.float32)
    ab = ctx.upload(aabbs.reshape(-1))
    proc = ctx.build_blas([dict(aabbs=ab, aabb_count=50, stride=24)])
    tri = ctx.build_blas_from_mesh(scenes.icosphere(2))
    flags = T.BUILD_FLAG_ALLOW_UPDATE
    tlas = ctx.build_tlas([proc], [scenes.IDENTITY_3X4], build_flags=flags)
    assert tlas.info().has_procedural == 1
    tlas._keep[1] = [tri]  # the binding keeps the BLAS list of the build: swap the BLAS for the refit
    ctx.update_tlas(tlas, [scenes.IDENTITY_3X4])
    assert tlas.info().has_procedural == 0
    rays = random_rays(500, seed=2, lo=(-2, -2, -2), hi=(2, 2, 2))
    hits = ctx.trace(tlas, rays)  # no RT_ERR_UNSUPPORTED any more
    assert (hits["primitive_index"] != T.NO_HIT).sum() > 50


def test_hit_record_table_holes_and_out_of_range_records(ctx, rt):
    case = bunny_case(3)
    w, h = 96, 64
    r = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
    f = scenes.make_frame(case.setup, w, h, 0, 0)
    r.dispatch(f)
    ctx.status()
    # a record for instance 3 leaves instances 1..2 unbound: a hole inside the bound range
    b = r.blases[0]
    r.program.set_hit_record(0, 3, b.vb, b.ib, case.materials[0])
    with pytest.raises(rt.RtError) as e:
        r.dispatch(f)
    assert e.value.code == -1 and "unbound hit record" in str(e.value)
    # an instance whose InstanceContributionToHitGroupIndex points past the bound records: reported by the status check
    r2 = case.renderer(rt, ctx, rt.PROGRESSIVE, w, h)
    r2.tlas = ctx.build_tlas(r2.blases, case.transforms, hit_groups=[40])
    r2.dispatch(f)
    with pytest.raises(rt.RtError) as e:
        ctx.status()
    assert e.value.code == -1 and "record" in str(e.value)
    ctx.status()


def test_realtime_program_checks_the_second_output(ctx, rt):
    case = bunny_case(2)
    w, h = 64, 48
    r = case.renderer(rt, ctx, rt.REALTIME, w, h)
    f = scenes.make_frame(case.setup, w, h, 0, 0)
    rt.check(rt.lib.rt_set_output(ctx.handle, 0, r.out[0].ptr, 16 * w))
    rt.check(rt.lib.rt_set_output(ctx.handle, 1, r.out[1].ptr, 8 * w))  # too small a pitch for slot 1
    rt.check(rt.lib.rt_set_tlas(ctx.handle, r.tlas.result.ptr))
    rt.check(rt.lib.rt_set_frame_constants(ctx.handle, C.byref(f)))
    assert rt.lib.rt_dispatch_rays(ctx.handle, r.program.handle, w, h, 3) == -1
    r.dispatch(f)  # binds proper pitches again
    ctx.status()


def test_denoise_on_a_second_device_of_the_same_process(rt, orc):
    try:
        n = sum(1 for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.startswith("GPU "))
    except Exception:
        n = 0
    if n < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.Generator(np.random.PCG64(3))
    d = rng.random((70, 90, 4), dtype=np.float32)
    s = rng.random((70, 90, 4), dtype=np.float32)
    prm = T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0)
    want, _ = orc.denoise(d, s, prm)
    outs = []
    for dev in (0, 1, 0):  # the opt-in to > 48 KB of shared memory is per device
        c = rt.Context(dev)
        out, _ = c.denoise(d, s, prm)
        outs.append(out)
        c.close()
    for out in outs:
        np.testing.assert_allclose(out[..., :3], want[..., :3], rtol=1e-5, atol=1e-6)
