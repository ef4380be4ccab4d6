#!/bin/bash
TAG=${1:-n3}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py tests/test_independent_answers.py tests/test_render_options.py -m gpu -x -q -k "denoise or Denoise or c5 or realtime" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -m pytest tests/test_gpu_build.py -m gpu -x -q >> gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
{
python tools/probe_denoise.py 2>&1 | tail -1
RT_CORE_LIB=$PWD/build/rt_dnscalar.so python tools/probe_denoise.py 2>&1 | tail -1
python tools/probe_denoise.py 2>&1 | tail -1
RT_CORE_LIB=$PWD/build/rt_dnscalar.so python tools/probe_denoise.py 2>&1 | tail -1
python tools/probe_scale.py --build 1000000,10000000 --flags 8 --reps 7 2>&1 | grep '"probe"' | cut -c1-120
} > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_ab.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_build0_launches.csv python tools/probe_scale.py --build 10000000 --flags 0 --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_table.py gpurun_out/${TAG}_build0_launches.csv 27
ncu --set full --clock-control none --import-source on -k regex:"k_denoise" -s 4 -c 2 -o gpurun_out/${TAG}_prof_denoise python tools/probe_denoise.py > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_treelet_reorder|k_find_treelets|k_fit_local|k_fit_exits" -s 4 -c 4 -o gpurun_out/${TAG}_prof_treelet python tools/probe_scale.py --build 10000000 --flags 0 --reps 1 > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/${TAG}_*
