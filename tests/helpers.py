"""Shared scene/fixture helpers: the same numpy inputs go to the oracle and to the CUDA path."""
import numpy as np

from dxrexperiments_b200 import scenes, types as T

QUAD_INDICES_CW = np.array([0, 1, 2, 2, 1, 3], dtype=np.uint16)  # UT:3634-3637


def ut_quad(kind="full", winding="cw", left=-1.0, right=1.0, top=-1.0, bottom=1.0, depth=1.0):
    """Geometry of TracingTests::BuildBottomLevelAccelerationStructure (UT:3630-3663): float3 vertices
    (stride 12), R16 indices."""
    l, r = left, right
    if kind == "left":
        r = (left + right) / 2.0
    elif kind == "right":
        l = (left + right) / 2.0
    verts = np.array([[l, top, depth], [l, bottom, depth], [r, top, depth], [r, bottom, depth]], dtype=np.float32)
    idx = QUAD_INDICES_CW.copy()
    if winding == "ccw":
        idx[[1, 2]] = idx[[2, 1]]
        idx[[4, 5]] = idx[[5, 4]]
    return verts, idx


def ut_rays(width=6, height=4, left=-1.0, top=-1.0, right=1.0, bottom=1.0):
    """RayGen of SimpleRayTracing.hlsl:28-46: origin lerped over the viewport at z=0, direction +z, tMax 1e4."""
    rays = np.zeros(width * height, dtype=T.RAY_DTYPE)
    for y in range(height):
        for x in range(width):
            lx, ly = np.float32((x + 0.5) / width), np.float32((y + 0.5) / height)
            ox = np.float32(left) + lx * (np.float32(right) - np.float32(left))
            oy = np.float32(top) + ly * (np.float32(bottom) - np.float32(top))
            rays[y * width + x] = ((ox, oy, 0.0), 0.0, (0.0, 0.0, 1.0), 10000.0)
    return rays


def partition_transform(i, total, left=-1.0, right=1.0):
    """TransformFromFullScreenToScreenPartition (UT:3935-3950)."""
    width = right - left
    xs = 1.0 / total
    pw = xs * width
    xo = pw * i + pw / 2.0
    return np.array([xs, 0, 0, xo + left, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)


def random_rays(n, seed, lo, hi, tmax=1e38, tmin=0.0):
    """Incoherent rays: origins uniform in the box [lo, hi], directions uniform on the sphere."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rays = np.zeros(n, dtype=T.RAY_DTYPE)
    rays["origin"] = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


class SceneCase:
    """One scene instantiated on both sides."""

    def __init__(self, meshes, transforms=None, materials=None, env_size=16, camera=None, setup=None):
        self.meshes = list(meshes)
        self.transforms = [scenes.IDENTITY_3X4] * len(self.meshes) if transforms is None else transforms
        self.materials = [scenes.make_material() for _ in self.meshes] if materials is None else materials
        self.env = scenes.sky_cube(env_size) if env_size else None
        self.setup = setup or scenes.FrameSetup(camera=camera or scenes.Camera())

    def oracle(self, orc):
        blases = [orc.Blas.from_mesh(m) for m in self.meshes]
        tlas = orc.Tlas(blases, self.transforms)
        recs = orc.Records(self.meshes, self.materials)
        return tlas, recs

    def renderer(self, rt, ctx, kind, width, height):
        return rt.Renderer(ctx, self.meshes, self.transforms, self.materials, self.env, kind, width, height)


def cornell_case():
    setup = scenes.FrameSetup(camera=scenes.Camera(eye=(0.0, 0.0, 3.5), at=(0.0, 0.0, 0.0)),
                              point_light_pos=(0.0, 0.5, 0.0, 1.0))
    return SceneCase([scenes.cornell_box()], setup=setup)


def bunny_case(subdiv=4):
    return SceneCase([scenes.bunny_scale(subdiv)], camera=scenes.BUNNY_CAMERA)


def two_material_case():
    """Two instances with different materials and transforms (exercises records, TLAS and instance transforms)."""
    ball = scenes.icosphere(3)
    ground = scenes.quad((-20, 0, 20), (20, 0, 20), (20, 0, -20), (-20, 0, -20))
    t_ball = np.array([3, 0, 0, 0, 0, 3, 0, 4, 0, 0, 3, 0], dtype=np.float32)
    mats = [scenes.make_material(), scenes.make_material(albedo=(0.2, 0.6, 0.9, 1.0), type=0, reflectivity=0.0,
                                                       emissive=(0.1, 0.1, 0.1, 1.0))]
    return SceneCase([ball, ground], transforms=[t_ball, scenes.IDENTITY_3X4], materials=mats)
