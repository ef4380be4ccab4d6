"""Regenerates tests/golden/*.npz from the CPU oracle.

    python tests/golden/make_golden.py

The reference cannot be executed here (Windows/D3D12), so these vectors are NOT reference outputs: they freeze the
oracle's answers (which tests/test_oracle_reference_kats.py pins against the reference's own test logic) so that
(a) a later change to the oracle cannot silently move the goalposts and (b) the GPU box — which has no
/root/reference and may have a different libm — checks the CUDA path against bytes committed to the repo.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from dxrexperiments_b200 import scenes, types as T  # noqa: E402
from helpers import cornell_case  # noqa: E402


def rng_kats():
    rows = []
    for v0, v1 in [(0, 0), (12345, 7), (1920 * 1080 - 1, 15), (0xFFFFFFFF, 0xFFFFFFFF)]:
        s = oracle.init_rand(v0, v1)
        vals = []
        st = s
        for _ in range(4):
            v, st = oracle.next_rand(st)
            vals.append(v)
        rows.append((v0, v1, s, vals))
    return (np.array([[r[0], r[1], r[2]] for r in rows], np.uint64), np.array([r[3] for r in rows], np.float32))


def main():
    seeds, rands = rng_kats()
    # build: a seeded soup with duplicate codes
    soup = scenes.triangle_soup(2000, seed=99, extent=50.0, edge=2.0)
    blas = oracle.Blas.from_mesh(soup)
    build = dict(aabb=blas.scene_aabb(), morton=blas.morton(), perm=blas.perm())
    # hierarchy + fitted nodes per treelet-pass count: default flags (1 pass), PREFER_FAST_BUILD (0: the plain Karras
    # tree), PREFER_FAST_TRACE (3) — FL/TreeletReorder.cpp:66-80
    for suffix, flags in (("", 0), ("_fast_build", T.BUILD_FLAG_PREFER_FAST_BUILD), ("_fast_trace", T.BUILD_FLAG_PREFER_FAST_TRACE)):
        b = oracle.Blas.from_mesh(soup, build_flags=flags)
        build["hier" + suffix] = b.hierarchy().view(np.uint32).reshape(-1, 3)
        build["nodes" + suffix] = T.parse_blas_blob(b.blob())["nodes"].view(np.uint32).reshape(-1, 8)
    np.savez_compressed(os.path.join(HERE, "build_soup2000.npz"), **build)
    # trace + render: Cornell 48x48
    case = cornell_case()
    tlas, recs = case.oracle(oracle)
    w = h = 48
    frame = scenes.make_frame(case.setup, w, h, 0, 0)
    rays = oracle.primary_rays(frame, w, h)
    hits = tlas.trace(rays, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES)
    acc = np.zeros((h, w, 4), np.float32)
    jit = scenes.jitter_sequence(5, 4, w, h)
    for s in range(4):
        oracle.render_progressive(tlas, recs, case.env, scenes.make_frame(case.setup, w, h, s, s, jitter=jit[s]), w, h, acc)
    direct, spec = oracle.render_realtime(tlas, recs, case.env, scenes.make_frame(case.setup, w, h, 3, 0, jitter=jit[0]), w, h)
    prm = T.DenoiserParams(1.0, 2.2, 1, 0, 12, 0)
    den, _ = oracle.denoise(direct, spec, prm)
    np.savez_compressed(os.path.join(HERE, "cornell48.npz"), prim=hits["primitive_index"], t=hits["t"], bary=hits["bary"],
                        progressive4=acc, direct=direct, spec=spec, denoised=den)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), seeds=seeds, rands=rands)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
