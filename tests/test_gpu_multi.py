"""Multi-GPU accumulation through the product: N ranks, one process per GPU, strip x sample sharding,
rt_accum_reduce (NCCL inside librt_core) — the reduced frame equals the single-GPU frame up to fp32 summation order.

Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box.
The single-GPU half (strip-interleaved dispatches assemble the full frame bit for bit) runs everywhere.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from dxrexperiments_b200 import scenes

from helpers import bunny_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("groups,strip_rows", [(2, 16), (3, 8), (4, 64), (8, 4)])
def test_interleaved_strips_assemble_the_full_frame(ctx, rt, groups, strip_rows):
    """Every strip group rendered into the same buffer = the full-frame dispatch, bit for bit (ragged last strip included)."""
    case = bunny_case(4)
    W, H = 328, 203  # neither a multiple of the 8x4 pixel tile nor of the strip height
    f0 = scenes.make_frame(case.setup, W, H, 0, 0, jitter=(0.3, -0.2))
    f1 = scenes.make_frame(case.setup, W, H, 1, 1, jitter=(-0.1, 0.4))
    full = case.renderer(rt, ctx, rt.PROGRESSIVE, W, H)
    full.dispatch(f0), full.dispatch(f1)
    parts = case.renderer(rt, ctx, rt.PROGRESSIVE, W, H)
    for f in (f0, f1):
        for g in range(groups):
            parts.dispatch(f, strips=(strip_rows, groups, g))
    np.testing.assert_array_equal(parts.image(0), full.image(0))
    # one group alone touches only its own strips
    one = case.renderer(rt, ctx, rt.PROGRESSIVE, W, H)
    one.dispatch(f0, strips=(strip_rows, groups, 1))
    img = one.image(0)
    rows = np.arange(H)
    mine = (rows // strip_rows) % groups == 1
    assert not img[~mine].any()
    first = case.renderer(rt, ctx, rt.PROGRESSIVE, W, H)
    first.dispatch(f0)
    np.testing.assert_array_equal(img[mine], first.image(0)[mine])
    ctx.status()


def test_interleaved_dispatch_rejects_bad_arguments(ctx, rt):
    case = bunny_case(2)
    r = case.renderer(rt, ctx, rt.PROGRESSIVE, 64, 64)
    f = scenes.make_frame(case.setup, 64, 64, 0, 0)
    for bad in [(12, 2, 0), (2, 2, 0), (16, 2, 2), (16, 0, 0)]:
        with pytest.raises(rt.RtError):
            r.dispatch(f, strips=bad)


def _run_ranks(tmp_path, world, strip_groups, spp, mode="progressive"):
    n = _gpu_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, this box has {n}")
    out, idf = tmp_path / "result.json", tmp_path / "nccl_id"
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MGPU_STRIP_GROUPS=str(strip_groups),
                   MGPU_SPP=str(spp), MGPU_OUT=str(out), MGPU_ID_FILE=str(idf), MGPU_MODE=mode)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_realtime_row_bands_composited_by_the_reduce_equal_the_single_gpu_frame(tmp_path, world):
    """realtime AOVs + DenoiseCompositor sharded by row bands (halo = the filter's reach), composited on rank 0
    by the weight-1 rt_accum_reduce: x + 0 + ... + 0 is x, so the frame is the single-GPU frame bit for bit."""
    res = _run_ranks(tmp_path, world, 0, 1, mode="realtime")
    assert res["bit_identical"], res
    assert abs(res["alpha_min"] - 1.0) < 1e-6 and abs(res["alpha_max"] - 1.0) < 1e-6, res
    evidence = os.path.join(ROOT, "gpurun_out", f"mgpu_realtime_bands_w{world}.json")
    os.makedirs(os.path.dirname(evidence), exist_ok=True)
    json.dump(res, open(evidence, "w"))


@pytest.mark.parametrize("world,strip_groups,spp", [(2, 1, 8), (2, 2, 4), (4, 2, 8), (8, 2, 16)])
def test_nccl_reduced_frame_equals_single_gpu_frame(tmp_path, world, strip_groups, spp):
    n = _gpu_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, this box has {n}")
    out, idf = tmp_path / "result.json", tmp_path / "nccl_id"
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MGPU_STRIP_GROUPS=str(strip_groups),
                   MGPU_SPP=str(spp), MGPU_OUT=str(out), MGPU_ID_FILE=str(idf))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = json.load(open(out))
    assert res["rel_rmse"] <= 1e-6, res          # fp32 summation order only
    assert res["max_abs"] <= 1e-5 * max(res["mean"], 1.0) * 10, res
    assert abs(res["alpha_min"] - 1.0) < 1e-6 and abs(res["alpha_max"] - 1.0) < 1e-6, res
    evidence = os.path.join(ROOT, "gpurun_out", f"mgpu_parity_w{world}_g{strip_groups}.json")
    os.makedirs(os.path.dirname(evidence), exist_ok=True)
    json.dump(res, open(evidence, "w"))
