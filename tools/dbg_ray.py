import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from dxrexperiments_b200 import scenes, rtcore as rt, types as T
ctx = rt.Context(0)
wl = scenes.workload("C2")
blas = ctx.build_blas_from_mesh(wl.meshes[0])
tlas = ctx.build_tlas([blas], [scenes.IDENTITY_3X4])
def ray(o, d, tmin=1e-4, tmax=1e38):
    r = np.zeros(1, T.RAY_DTYPE); r["origin"] = o; r["direction"] = d; r["tmin"] = tmin; r["tmax"] = tmax; return r
cases = {"up-outside": ray((20, -0.5, 20), (0, 1, 0)), "up-under": ray((1, -0.5, 1), (0, 1, 0)),
         "up-under-tilt": ray((1, -0.5, 1), (1e-6, 1, 1e-6)), "x-only": ray((-20, 7, 0.3), (1, 0, 0)),
         "diag": ray((-20, 3, 0.3), (1, 0.2, 0.1))}
for name, r in cases.items():
    h, st = ctx.trace(tlas, r, 0, stats=True)
    t0 = time.time(); h2 = ctx.trace(tlas, r, 0); ctx.sync(); t1 = time.time() - t0
    t0 = time.time(); h3, _ = ctx.trace(tlas, r, 0, stats=True); ctx.sync(); t2 = time.time() - t0
    print(name, "bvh2 visits int %d leaf %d maxstack %d  prim %d/%d t %g/%g   wall: persistent-bvh4 %.2f ms, bvh2 %.2f ms" % (st[1], st[2], st[4], h["primitive_index"][0], h2["primitive_index"][0], h["t"][0], h2["t"][0], t1 * 1e3, t2 * 1e3))
ctx.status()
