#!/bin/bash
# tools/run_scale8_short.sh <tag> — 8-GPU box, reduced pass: N-rank parity tests, then the bench at N = 1 and N = 8 (weak headline + strong block)
TAG=${1:-s}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py tests/test_gpu_validation.py -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_multi.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-target-scene --build-tris 0 --no-denoise > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_n8.err
cat gpurun_out/${TAG}_pytest_multi.log; tail -c 400 gpurun_out/${TAG}_n8.err
python - <<PY
import json
for f in ["gpurun_out/${TAG}_bench_n%d.json"%n for n in (1,8)]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); s=d["strong"]
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "strong ms", round(s["time_to_frame_ms"],2), s["shard_plan"], {k:round(v,3) for k,v in s["root_breakdown_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
