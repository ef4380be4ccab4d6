"""Closed-form answers for the shading path — expected images that do NOT come from the oracle.

The reference has no test for shading, RNG, accumulation or the denoiser, so the oracle's transcription of
assets/shaders/*.hlsl(i) is pinned here against analytic radiance values computed in float64 numpy straight from the
shader's formulas (cited per case), and the CUDA path is held to the same numbers on the GPU.  Each case is a scene in
which the Monte-Carlo estimator is a CONSTANT of the random numbers, so a single sample equals the expectation.
"""
import numpy as np

from dxrexperiments_b200 import scenes, types as T

M_PI_SHADER = float(np.float32(3.1415927))  # RaytracingUtils.hlsli:22


def camera_rays(setup, w, h):
    """RayGen (ProgressiveRaytracing.hlsl:17-31) in float64, jitter 0: per-pixel origin and unit direction."""
    u, v, fw = (np.asarray(a, np.float64) for a in setup.camera.uvw(w / h))
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    dx = (px + 0.5) / w * 2.0 - 1.0
    dy = (py + 0.5) / h * 2.0 - 1.0
    d = dx[..., None] * u - dy[..., None] * v + fw
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(np.asarray(setup.camera.eye, np.float64), d.shape)
    return o, d


class Case:
    def __init__(self, name, meshes, materials, setup, env, options, w, h, expected, mask, realtime=False, spp=1, tol=2e-5):
        self.name, self.meshes, self.materials, self.setup, self.env = name, meshes, materials, setup, env
        self.options, self.w, self.h, self.expected, self.mask = options, w, h, expected, mask
        self.realtime, self.spp, self.tol = realtime, spp, tol

    def frame(self, s):
        return scenes.make_frame(self.setup, self.w, self.h, s, 0 if self.realtime else s, options=self.options)

    def check(self, img, what=""):
        got = np.asarray(img, np.float64)[..., :3][self.mask]
        want = self.expected[self.mask]
        assert self.mask.sum() > 0.2 * self.w * self.h, "mask too small to mean anything"
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-12)
        assert err <= self.tol, f"{self.name} {what}: max error {err:.3e} relative to the largest expected value"


def const_env(rgb, size=8):
    e = np.zeros((6, size, size, 4), np.float32)
    e[..., :3] = np.asarray(rgb, np.float32)
    e[..., 3] = 1.0
    return e


def _options(**kw):
    o = scenes.default_options()
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def lit_plane(w=96, h=64):
    """Directional + point light on a diffuse plane, no indirect light: pixel = albedo * (dir + point) / pi with
    dir = color.rgb * color.a * sat(n.L)                     (RaytracingCommon.hlsli:126-134, L = normalize(-forwardDir))
    point = color.rgb * color.a * sat(n.L) / (2 pi dist^2)   (:136-147), every pixel unshadowed."""
    plane = scenes.quad((-50, -50, 0), (50, -50, 0), (50, 50, 0), (-50, 50, 0))  # normal +z
    albedo = np.array([0.8, 0.4, 0.2])
    mat = scenes.make_material(albedo=(*albedo, 1.0), type=0, reflectivity=0.0, emissive=(0, 0, 0, 0))
    setup = scenes.FrameSetup(camera=scenes.Camera(eye=(0.5, -0.25, 6.0), at=(0.0, 0.0, 0.0)), dir_light_color=(0.9, 0.7, 0.5, 1.5),
                              point_light_color=(0.2, 0.8, 0.6, 2.0), point_light_pos=(1.0, 0.5, 2.0, 1.0))
    o, d = camera_rays(setup, w, h)
    t = -o[..., 2] / d[..., 2]
    p = o + t[..., None] * d
    n = np.array([0.0, 0.0, 1.0])
    fwd = np.asarray(setup.directional_forward(), np.float64)[:3]
    L = -fwd / np.linalg.norm(fwd)
    dir_c = np.asarray(setup.dir_light_color[:3]) * setup.dir_light_color[3] * max(0.0, min(1.0, float(n @ L)))
    path = np.asarray(setup.point_light_pos[:3]) - p
    dist = np.linalg.norm(path, axis=-1)
    nol = np.clip((path / dist[..., None]) @ n, 0.0, 1.0)
    point_c = (np.asarray(setup.point_light_color[:3]) * setup.point_light_color[3])[None, None, :] * (nol / (2.0 * M_PI_SHADER * dist * dist))[..., None]
    expected = albedo * (dir_c + point_c) / M_PI_SHADER
    mask = np.ones((h, w), bool)
    return Case("lit_plane", [plane], [mat], setup, const_env((0, 0, 0)), _options(noIndirectDiffuse=1), w, h, expected, mask)


def shadowed_emitter(w=96, h=64):
    """An occluder between the plane and the directional light, point light switched off (alpha 0), no indirect light:
    inside the umbra the pixel is emissive.rgb * emissive.a exactly (ProgressiveRaytracing.hlsl:147, visibility 0 from
    shootShadowRay, RaytracingCommon.hlsli:84-96); outside it is emissive + albedo * dir / pi; missed pixels show the
    environment (PrimaryMiss :160-164)."""
    plane = scenes.quad((-6, -2.5, 0), (6, -2.5, 0), (6, 2.5, 0), (-6, 2.5, 0))
    occ = scenes.quad((0.5, -1.0, 1.0), (2.5, -1.0, 1.0), (2.5, 1.0, 1.0), (0.5, 1.0, 1.0))
    albedo, emis, ea = np.array([0.5, 0.6, 0.7]), np.array([0.1, 0.2, 0.3]), 1.5
    mat = scenes.make_material(albedo=(*albedo, 1.0), type=0, reflectivity=0.0, emissive=(*emis, ea))
    # elapsed_time -2.618 -> (0.3, -0.2, -1) rotated by about -45 degrees about Y: the light comes from -x, the shadow falls
    # towards +x, clear of the part of the plane the occluder hides from the camera
    setup = scenes.FrameSetup(camera=scenes.Camera(eye=(0.0, 0.0, 9.0), at=(0.0, 0.0, 0.0)), elapsed_time=-2.618,
                              dir_light_color=(1.0, 0.9, 0.8, 2.0), point_light_color=(1.0, 1.0, 1.0, 0.0), point_light_pos=(0.0, 0.0, 5.0, 1.0))
    env_rgb = np.array([0.3, 0.5, 0.7])
    o, d = camera_rays(setup, w, h)
    fwd = np.asarray(setup.directional_forward(), np.float64)[:3]
    L = -fwd / np.linalg.norm(fwd)
    n = np.array([0.0, 0.0, 1.0])
    # what the camera sees first: occluder (z = 1) or plane (z = 0) or nothing
    t1 = (1.0 - o[..., 2]) / d[..., 2]
    q = o + t1[..., None] * d
    on_occ = (q[..., 0] > 0.5) & (q[..., 0] < 2.5) & (np.abs(q[..., 1]) < 1.0)
    t0 = -o[..., 2] / d[..., 2]
    p = o + t0[..., None] * d
    on_plane = (np.abs(p[..., 0]) < 6.0) & (np.abs(p[..., 1]) < 2.5) & ~on_occ
    # shadow: the ray p + s L crosses z = 1 inside the occluder
    s = 1.0 / L[2]
    sx, sy = p[..., 0] + s * L[0], p[..., 1] + s * L[1]
    margin = 0.08  # stay clear of the penumbra-free but pixel-quantised shadow boundary
    in_shadow = (sx > 0.5 + margin) & (sx < 2.5 - margin) & (np.abs(sy) < 1.0 - margin)
    out_shadow = ~((sx > 0.5 - margin) & (sx < 2.5 + margin) & (np.abs(sy) < 1.0 + margin))
    lit = emis * ea + albedo * (np.asarray(setup.dir_light_color[:3]) * setup.dir_light_color[3] * float(n @ L)) / M_PI_SHADER
    expected = np.zeros((h, w, 3))
    expected[:] = env_rgb
    expected[on_plane & in_shadow] = emis * ea
    expected[on_plane & out_shadow] = lit
    expected[on_occ] = lit  # the occluder faces the light itself
    edge = 0.06
    clear_of_edges = (np.abs(np.abs(p[..., 0]) - 6.0) > edge) & (np.abs(np.abs(p[..., 1]) - 2.5) > edge) & \
                     (np.abs(q[..., 0] - 0.5) > edge) & (np.abs(q[..., 0] - 2.5) > edge) & (np.abs(np.abs(q[..., 1]) - 1.0) > edge)
    mask = clear_of_edges & ((on_plane & (in_shadow | out_shadow)) | on_occ | (~on_plane & ~on_occ))
    assert (on_plane & in_shadow & mask).sum() > 50 and (~on_plane & ~on_occ & mask).sum() > 50
    return Case("shadowed_emitter", [scenes.merge([plane, occ])], [mat], setup, const_env(env_rgb), _options(noIndirectDiffuse=1), w, h, expected, mask)


def furnace(w=80, h=60, spp=1, uniform=False):
    """Camera inside a closed room whose walls all emit E = emissive.rgb * emissive.a, lights off, environment = E as well:
    every secondary ray returns exactly E (shade() at depth 1: emissive + albedo * 0 / pi + reflectivity * (0 * brdf / pdf)
    * fresnel, ProgressiveRaytracing.hlsl:104-147; or the environment on a miss), so with cosine sampling
      indirect = E * pi (:70), diffuse = E, specular = E * brdf / pdf = E * (e + 2) / (e + 1), e = exp((1 - roughness) * 12)
      (RaytracingUtils.hlsli:101-123), fresnel = f0 + (1 - f0) * (1 - sat(dot(-D, n)))^5 (:126-130)
      pixel = E + albedo * E + reflectivity * E * (e + 2) / (e + 1) * fresnel             — a constant of the random numbers."""
    room = scenes.box((-4, -3, -5), (4, 3, 5), inward=True)
    albedo, emis, ea = np.array([0.6, 0.3, 0.1]), np.array([0.2, 0.3, 0.4]), 1.5
    refl, rough, f0 = 0.7, 0.5, 0.58
    mat = scenes.make_material(albedo=(*albedo, 1.0), specular=(f0, f0, f0, 1.0), emissive=(*emis, ea), reflectivity=refl, roughness=rough, type=1)
    setup = scenes.FrameSetup(camera=scenes.Camera(eye=(0.5, 0.2, 3.0), at=(-0.5, 0.0, -5.0)), dir_light_color=(1, 1, 1, 0.0),
                              point_light_color=(1, 1, 1, 0.0), point_light_pos=(0.0, 0.0, 0.0, 1.0))
    E = emis * ea
    o, d = camera_rays(setup, w, h)
    # first wall hit and its (inward) normal
    lo, hi = np.array([-4.0, -3.0, -5.0]), np.array([4.0, 3.0, 5.0])
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = (hi - o) / d
        t_lo = (lo - o) / d
    t_exit = np.where(d > 0, t_hi, t_lo)
    axis = np.argmin(t_exit, axis=-1)
    n = np.zeros_like(d)
    sign = -np.sign(np.take_along_axis(d, axis[..., None], -1))[..., 0]
    np.put_along_axis(n, axis[..., None], sign[..., None], -1)
    cos = np.clip(np.sum(-d * n, axis=-1), 0.0, 1.0)
    fres = f0 + (1.0 - f0) * (1.0 - cos) ** 5
    e = float(np.exp(np.float32((1.0 - rough) * 12.0)))
    expected = E + albedo * E + refl * E * ((e + 2.0) / (e + 1.0)) * fres[..., None]
    # pixels whose camera ray passes within a hair of a room edge may see either wall
    sorted_t = np.sort(t_exit, axis=-1)
    mask = (sorted_t[..., 1] - sorted_t[..., 0]) > 0.05
    opts = _options(cosineHemisphereSampling=0 if uniform else 1)
    return Case("furnace_uniform" if uniform else "furnace", [room], [mat], setup, const_env(E), opts, w, h, expected, mask, spp=spp,
                tol=2e-5 if not uniform else 3e-2)


def realtime_direct(w=96, h=64):
    """RealtimeRaytracing.hlsl:65-103: AOV 0 = albedo * direct / pi for the lit plane (same light formulas), AOV 1 = 0 for a
    material without reflection."""
    c = lit_plane(w, h)
    c.name, c.realtime = "realtime_direct", True
    c.options = _options()
    return c


CASES = [lit_plane, shadowed_emitter, furnace, realtime_direct]


# ---------------------------------------------------------------------------------------------- denoiser, from the text of SURVEY A5
def denoise_reference(direct, spec, k=12, exposure=1.0, tonemap=True, gamma_correct=False, gamma=2.2, mode=0):
    """DenoiseCompositor (BilateralFilter.hlsli:50-118, DenoiseCommon.hlsli:46-77) written from its specification with
    whole-image numpy operations in float64 — structurally unrelated to oracle/oracle_denoise.cpp's pixel loops.
    Returns (pass H result, final image)."""
    direct = np.asarray(direct, np.float64)[..., :3]
    spec = np.asarray(spec, np.float64)[..., :3]
    lut = np.array([1, 1, 0.9, 0.75, 0.6, 0.5, 0.0])
    radius = np.float32(k)
    den = np.float32(0.001) + np.float32(abs(radius * np.float32(0.8)))

    def w_s(i):
        idx = int(np.float32(abs(i) * 5) / den)
        return lut[min(max(idx, 0), 6)]

    def shifted(img, i, axis):
        out = np.zeros_like(img)  # reads outside the image return 0 (D3D out-of-bounds load)
        n = img.shape[axis]
        if abs(i) >= n:
            return out
        src = [slice(None)] * 3
        dst = [slice(None)] * 3
        if i >= 0:
            src[axis], dst[axis] = slice(i, n), slice(0, n - i)
        else:
            src[axis], dst[axis] = slice(0, n + i), slice(-i, n)
        out[tuple(dst)] = img[tuple(src)]
        return out

    def one_pass(inp, joint, axis):
        num = np.zeros_like(inp)
        wsum = np.zeros(inp.shape[:2])
        for i in range(-k, k + 1):
            jr = shifted(joint, i, axis)
            wr = 1.0 - np.clip(10.0 * np.abs(jr - joint).sum(axis=-1), 0.0, 1.0)
            wgt = w_s(i) * wr
            num += shifted(inp, i, axis) * wgt[..., None]
            wsum += wgt
        return num / wsum[..., None]

    hpass = one_pass(spec, direct, 1)
    v = one_pass(hpass, direct, 0)
    if mode == 0:
        c = v + direct
    elif mode == 1:
        c = v
    elif mode == 2:
        c = hpass  # raw input of the pass
    else:
        c = direct
    c = c * exposure
    if tonemap:
        lum = c @ np.array([0.299, 0.587, 0.114])
        with np.errstate(divide="ignore", invalid="ignore"):
            # a black pixel gives 0 * (0 / 1) / 0 = NaN, and HLSL's max(NaN, 0) is 0 (the non-NaN operand, as fmaxf / np.fmax)
            c = np.fmax(c * ((lum / (lum + 1.0)) / lum)[..., None], 0.0)
    if gamma_correct:
        c = np.clip(c ** (1.0 / gamma), 0.0, 1.0)
    return hpass, c
