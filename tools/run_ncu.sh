#!/bin/bash
# tools/run_ncu.sh <tag> — ncu evidence (one GPU): launch list of a short default bench, full captures of the trace kernels on C2 / C1M / C4,
# of the denoiser and of the shading kernels.  Numbers printed under ncu are never bench values.
TAG=${1:-p}
mkdir -p gpurun_out
Q="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-strong --no-target-scene --build-tris 0"
export RT_BANDS=1   # whole-frame kernels (one pixel band per dispatch), so that per-launch counters are per frame
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $Q --spp 2 > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_denoise" -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_denoise python bench.py $Q --spp 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_persistent|k_primary|k_shade|k_resolve" -s 14 -c 7 -f -o gpurun_out/${TAG}_prof_C2 python bench.py $Q --spp 4 --no-denoise > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_persistent|k_primary" -s 8 -c 4 -f -o gpurun_out/${TAG}_prof_C1M python bench.py $Q --config C1M --spp 4 --no-denoise > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_persistent|k_primary" -s 8 -c 4 -f -o gpurun_out/${TAG}_prof_C4 python bench.py $Q --config C4 --spp 4 --no-denoise > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
