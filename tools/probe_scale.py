#!/usr/bin/env python
"""probe_scale.py — build throughput (Mtri/s) versus triangle count and per-stage trace throughput on larger
scenes.  A development probe (run under gpurun); bench.py is the contract, this explores around it.

  python tools/probe_scale.py --build 100000,1000000,10000000 --render 8    # icosphere subdiv 8 = 1.3 M tris
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dxrexperiments_b200 import scenes, types as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", default="100000,1000000")
    ap.add_argument("--render", default="", help="comma list of icosphere subdivision levels to render at 1080p")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--flags", type=int, default=0, help="RT_BUILD_FLAG_* of the build probe (8 = PREFER_FAST_BUILD: no treelet pass)")
    args = ap.parse_args()
    import torch

    from dxrexperiments_b200 import rtcore as rt
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)

    for n in [int(x) for x in args.build.split(",") if x]:
        t0 = time.time()
        mesh = scenes.triangle_soup(n)
        tgen = time.time() - t0
        vb = ctx.upload(mesh.vertices)
        ib = ctx.upload(mesh.indices)
        import ctypes as C
        desc = (T.GeometryDesc * 1)()
        desc[0].vertex_buffer, desc[0].vertex_count, desc[0].vertex_stride_bytes = vb.ptr, mesh.vertices.shape[0], 24
        desc[0].index_buffer, desc[0].index_count, desc[0].index_format = ib.ptr, mesh.indices.size, 32
        desc[0].flags = T.GEOMETRY_FLAG_OPAQUE
        info = T.PrebuildInfo()
        rt.check(rt.lib.rt_blas_prebuild(ctx.handle, desc, 1, args.flags, C.byref(info)))
        scr, res = ctx.alloc(info.scratch_bytes), ctx.alloc(info.result_bytes)
        times = []
        for r in range(args.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rt.check(rt.lib.rt_blas_build(ctx.handle, desc, 1, args.flags, scr.ptr, scr.nbytes, res.ptr, res.nbytes))
            e1.record(stream)
            torch.cuda.synchronize()
            if r >= 2:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        print(json.dumps({"probe": "build", "triangles": n, "ms": ms, "mtri_per_s": n / ms / 1e3,
                          "algorithmic_gbs": n * 432 / ms / 1e6, "scratch_mb": info.scratch_bytes / 2**20,
                          "result_mb": info.result_bytes / 2**20, "gen_s": tgen}), flush=True)
        del scr, res, vb, ib

    for sub in [int(x) for x in args.render.split(",") if x]:
        mesh = scenes.bunny_scale(sub)
        setup = scenes.FrameSetup(camera=scenes.BUNNY_CAMERA)
        env = scenes.sky_cube(64)
        W, H = 1920, 1080
        jit = scenes.jitter_sequence(setup.seed, 64, W, H)
        r = rt.Renderer(ctx, [mesh], [scenes.IDENTITY_3X4], [scenes.make_material()], env, rt.PROGRESSIVE, W, H)

        def frames():
            for s in range(args.spp):
                r.dispatch(scenes.make_frame(setup, W, H, frame_count=s, accum_count=s, jitter=jit[s]))
        frames()
        ctx.enable_trace_stats(True)
        ctx.trace_stats(reset=True)
        frames()
        st = ctx.trace_stats(reset=True)
        ctx.enable_trace_stats(False)
        ctx.enable_stage_timing(True)
        ctx.stage_timing(reset=True)
        ctx.ray_counts(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.reps):
            frames()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tp, ts, tsh = ctx.stage_timing(reset=True)
        ctx.enable_stage_timing(False)
        rc = ctx.ray_counts(reset=True)
        out = {"probe": "render", "triangles": mesh.num_triangles, "ms_per_frame": ms / (args.reps * args.spp),
               "mrays_per_s_all": (rc.primary + rc.secondary + rc.shadow) / ms / 1e3}
        for name, s, t in zip(("primary", "secondary", "shadow"), st, (tp, ts, tsh)):
            rays = s.rays * args.reps
            out[name] = {"mrays_per_s": rays / t / 1e3 if t else None, "n_int": s.internal_visits / max(s.rays, 1),
                         "n_leaf": s.leaf_visits / max(s.rays, 1), "max_stack": int(s.max_stack), "ms_total": t}
        import hashlib
        out["image_sha1"] = hashlib.sha1(r.image(0).tobytes()).hexdigest()[:12]  # variants must not change the frame
        print(json.dumps(out), flush=True)
        ctx.status()


if __name__ == "__main__":
    main()
