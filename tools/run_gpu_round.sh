#!/bin/bash
# tools/run_gpu_round.sh <tag> — the standard evidence pass on a GPU box: GPU test suite, smoke, bench (ours + reference arm).
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2>> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; tail -c 600 gpurun_out/${TAG}_bench.err; head -c 1500 gpurun_out/${TAG}_bench.json
