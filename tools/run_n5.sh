#!/bin/bash
TAG=${1:-n5}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
TAIL=3 tools/ab.sh 6,8 > gpurun_out/${TAG}_ab.log 2>&1; cat gpurun_out/${TAG}_ab.log
python tools/probe_scale.py --build 1000000,10000000 --flags 8 --reps 7 2>&1 | grep '"probe"' | cut -c1-110
python tools/probe_scale.py --build 10000000 --flags 0 --reps 5 2>&1 | grep '"probe"' | cut -c1-110
ncu --set full --clock-control none --import-source on -k regex:"k_radix_scatter|k_lbvh_fit|k_lbvh_exits" -s 6 -c 4 -o gpurun_out/${TAG}_prof_build python tools/probe_scale.py --build 10000000 --flags 8 --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_*
