import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dxrexperiments_b200 import scenes, rtcore as rt
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
wl = scenes.workload("C2")
env = scenes.sky_cube(64)
W,H = 1920,1080
jit = scenes.jitter_sequence(wl.setup.seed, 1024, W, H)
r = rt.Renderer(ctx, wl.meshes, wl.transforms, wl.materials, env, rt.PROGRESSIVE, W, H)
ctx.enable_stage_timing(True)
for rep in range(2):
    for s in range(16):
        ctx.stage_timing(reset=True)
        r.dispatch(scenes.make_frame(wl.setup, W, H, frame_count=s, accum_count=s, jitter=jit[s]))
        tp, ts, tsh = ctx.stage_timing(reset=True)
        if rep: print(s, "primary %.3f secondary %.3f shadow %.3f" % (tp, ts, tsh))
