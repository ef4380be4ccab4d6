"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multi.py, one process per GPU).

Renders its shard of a progressive frame (strip group x sample group, dxrexperiments_b200.sharding.plan), sums the
ranks' buffers onto rank 0 with rt_accum_reduce (NCCL inside librt_core, the unique id exchanged through a file — no
torch.distributed), and rank 0 compares the reduced frame with the same frame rendered by ONE GPU in this process.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dxrexperiments_b200 import rtcore as rt, scenes, sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    strip_groups, spp = int(os.environ["MGPU_STRIP_GROUPS"]), int(os.environ["MGPU_SPP"])
    W, H = int(os.environ.get("MGPU_W", "640")), int(os.environ.get("MGPU_H", "360"))
    out_path, id_path = os.environ["MGPU_OUT"], os.environ["MGPU_ID_FILE"]
    ctx = rt.Context(int(os.environ.get("LOCAL_RANK", rank)))
    comm = rt.Comm(ctx, world, rank, rt.file_exchange(id_path))
    mesh = scenes.bunny_scale(4)
    ground = scenes.quad((-20, -1.2, 20), (20, -1.2, 20), (20, -1.2, -20), (-20, -1.2, -20))
    mats = [scenes.make_material(), scenes.make_material(albedo=(0.2, 0.6, 0.9, 1.0), type=0, reflectivity=0.0)]
    setup = scenes.FrameSetup(camera=scenes.BUNNY_CAMERA)
    env = scenes.sky_cube(16)
    jit = scenes.jitter_sequence(setup.seed, spp, W, H)
    if os.environ.get("MGPU_MODE") == "realtime":
        return realtime_bands(ctx, comm, rank, world, W, H, [mesh, ground], mats, env, setup, out_path)
    p = sharding.plan(rank, world, spp, strip_groups=strip_groups, strip_rows=16)
    r = rt.Renderer(ctx, [mesh, ground], [scenes.IDENTITY_3X4] * 2, mats, env, rt.PROGRESSIVE, W, H)
    for local, s in enumerate(p.samples):
        r.dispatch(scenes.make_frame(setup, W, H, s, local, jitter=jit[s]),
                   strips=(p.strip_rows, p.strip_groups, p.strip_group) if p.strip_groups > 1 else None)
    count = W * H * 4
    recv = ctx.alloc(4 * count).zero() if rank == 0 else None
    comm.reduce(r.out[0].ptr, recv.ptr if recv else None, count, p.weight, root=0)
    ctx.sync()
    ctx.status()
    if rank == 0:
        reduced = recv.download(np.float32).reshape(H, W, 4)
        single = rt.Renderer(ctx, [mesh, ground], [scenes.IDENTITY_3X4] * 2, mats, env, rt.PROGRESSIVE, W, H)
        for s in range(spp):
            single.dispatch(scenes.make_frame(setup, W, H, s, s, jitter=jit[s]))
        ref = single.image(0)
        a, b = reduced[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
        rel = float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2)))
        res = {"world": world, "strip_groups": strip_groups, "spp": spp, "rel_rmse": rel, "max_abs": float(np.abs(a - b).max()),
               "alpha_min": float(reduced[..., 3].min()), "alpha_max": float(reduced[..., 3].max()), "mean": float(b.mean()),
               "nccl_version": comm.nccl_version(), "version": rt.lib.rt_version().decode()}
        with open(out_path, "w") as f:
            json.dump(res, f)
    comm.close()
    ctx.close()


def realtime_bands(ctx, comm, rank, world, W, H, meshes, mats, env, setup, out_path):
    """one realtime frame (1 spp AOVs + DenoiseCompositor) sharded by row bands with the filter's reach as halo;
    the weight-1 rt_accum_reduce composites the bands on rank 0, which compares with the frame ONE GPU produces."""
    from dxrexperiments_b200 import types as T
    k = 12
    prm = T.DenoiserParams(1.0, 2.2, 1, 0, k, 0)  # src/DenoiseCompositor.cpp:45-50
    f = scenes.make_frame(setup, W, H, 0, 0, jitter=(0.2, -0.3))
    band = sharding.band_plan(rank, world, H, halo=k)
    r = rt.Renderer(ctx, meshes, [scenes.IDENTITY_3X4] * len(meshes), mats, env, rt.REALTIME, W, H)
    tmp, final = ctx.alloc(16 * W * H).zero(), ctx.alloc(16 * W * H).zero()
    r.realtime_band(f, band, prm, tmp, final)
    count = W * H * 4
    recv = ctx.alloc(4 * count).zero() if rank == 0 else None
    comm.reduce(final.ptr, recv.ptr if recv else None, count, 1.0, root=0)
    ctx.sync()
    ctx.status()
    if rank == 0:
        got = recv.download(np.float32).reshape(H, W, 4)
        single = rt.Renderer(ctx, meshes, [scenes.IDENTITY_3X4] * len(meshes), mats, env, rt.REALTIME, W, H)
        single.dispatch(f)
        ref, _ = ctx.denoise(single.image(0), single.image(1), prm)
        a, b = got.astype(np.float64), ref.astype(np.float64)
        res = {"world": world, "mode": "realtime row bands, halo 12", "rel_rmse": float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))),
               "max_abs": float(np.abs(a - b).max()), "bit_identical": bool(np.array_equal(got, ref)), "mean": float(b[..., :3].mean()),
               "alpha_min": float(got[..., 3].min()), "alpha_max": float(got[..., 3].max()), "strip_groups": 0, "spp": 1,
               "nccl_version": comm.nccl_version(), "version": rt.lib.rt_version().decode()}
        with open(out_path, "w") as fh:
            json.dump(res, fh)
    comm.close()
    ctx.close()


if __name__ == "__main__":
    main()
