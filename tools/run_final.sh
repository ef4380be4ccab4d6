#!/bin/bash
# tools/run_final.sh <tag> — the round's evidence pass on one GPU: tests, smoke, bench (+ reference arm), ncu captures, build launch lists
TAG=${1:-q}
tools/run_gpu_round.sh $TAG
tools/run_ncu.sh $TAG > /dev/null 2>&1
unset RT_BANDS
for f in 8 0; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_build${f}_launches.csv python tools/probe_scale.py --build 10000000 --flags $f --reps 1 > /dev/null 2>&1
done
python tools/probe_scale.py --build 100000,1000000,10000000 --flags 8 --reps 9 2>&1 | grep '"probe"' > gpurun_out/${TAG}_build_probe.log
python tools/probe_scale.py --build 10000000 --flags 0 --reps 5 2>&1 | grep '"probe"' >> gpurun_out/${TAG}_build_probe.log
python tools/probe_scale.py --build 10000000 --flags 4 --reps 5 2>&1 | grep '"probe"' >> gpurun_out/${TAG}_build_probe.log
ncu --set full --clock-control none --import-source on -k regex:"k_lbvh_fit|k_lbvh_exits|k_rearrange|k_radix_scatter" -s 2 -c 7 -f -o gpurun_out/${TAG}_prof_build python tools/probe_scale.py --build 10000000 --flags 8 --reps 1 > /dev/null 2>&1
cat gpurun_out/${TAG}_build_probe.log | cut -c1-100
ls gpurun_out/${TAG}_*
