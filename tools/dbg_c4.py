import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle as orc
from dxrexperiments_b200 import scenes, rtcore as rt, types as T
ctx = rt.Context(0)
wl = scenes.workload("C4"); W, H = wl.width, wl.height
env = scenes.sky_cube(64)
blases = [orc.Blas.from_mesh(m) for m in wl.meshes]
otlas = orc.Tlas([blases[k] for k in wl.instance_mesh], wl.transforms)
recs = orc.Records([wl.meshes[k] for k in wl.instance_mesh], [wl.materials[k] for k in wl.instance_mesh])
r = rt.Renderer(ctx, wl.meshes, wl.transforms, wl.materials, env, rt.PROGRESSIVE, W, H, instance_mesh=wl.instance_mesh)
f = scenes.make_frame(wl.setup, W, H, 0, 0)
acc = np.zeros((H, W, 4), np.float32)
oc = T.RayCounts()
orc.render_progressive(otlas, recs, env, f, W, H, acc, threads=16, counts=oc)
ctx.ray_counts(reset=True)
r.dispatch(f)
gc = ctx.ray_counts()
img = r.image(0)
d = np.abs(img[..., :3].astype(np.float64) - acc[..., :3]).max(-1)
print("counts gpu", gc.primary, gc.secondary, gc.shadow, "oracle", oc.primary, oc.secondary, oc.shadow)
print("pixels differing > 1e-4:", (d > 1e-4).sum(), " > 1e-2:", (d > 1e-2).sum(), "max", d.max(), "nonfinite", (~np.isfinite(img)).sum(), (~np.isfinite(acc)).sum())
idx = np.argsort(d.ravel())[-8:]
for i in idx:
    y, x = divmod(i, W); print((x, y), img[y, x, :3], acc[y, x, :3])
print("rmse", np.sqrt((d**2).mean()) , "ref rms", np.sqrt((acc[..., :3].astype(np.float64)**2).mean()))
