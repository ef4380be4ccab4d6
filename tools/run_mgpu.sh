#!/bin/bash
# tools/run_mgpu.sh <tag> <N> — multi-GPU evidence on an N-GPU box: N-rank parity tests, bench at N (weak headline + strong block).
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py tests/test_gpu_validation.py -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
$TR bench.py --gpus $N --steps 1 --warmup 3 --no-e2e --strong-strip-groups 2 > gpurun_out/${TAG}_bench_n${N}_strips2.json 2>> gpurun_out/${TAG}_bench_n${N}.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-target-scene --build-tris 0 --no-denoise > gpurun_out/${TAG}_bench_n1.json 2>> gpurun_out/${TAG}_bench_n${N}.err
cat gpurun_out/${TAG}_pytest_multi.log; tail -c 1500 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n1.json","gpurun_out/${TAG}_bench_n${N}.json","gpurun_out/${TAG}_bench_n${N}_strips2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); s=d["strong"]
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "strong ms", s["time_to_frame_ms"], s["shard_plan"], s["root_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
