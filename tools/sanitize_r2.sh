#!/bin/bash
# tools/sanitize_r2.sh — compute-sanitizer over what round 2 added: the binning shade kernel (shared-memory counting sort), strip-interleaved
# dispatches, the depth-2 wave, half targets, the debug capture, closed-form scenes.  Small cases only.
mkdir -p gpurun_out
T="timeout 900 compute-sanitizer --error-exitcode 9"
$T --tool memcheck python -m pytest tests/test_gpu_stage_parity.py tests/test_render_options.py tests/test_independent_answers.py tests/test_gpu_multi.py -m gpu -k "not nccl" -x -q 2>&1 | tail -4 > gpurun_out/san_memcheck.log
$T --tool racecheck python -m pytest tests/test_gpu_stage_parity.py -m gpu -k "cornell or realtime" -x -q 2>&1 | tail -4 > gpurun_out/san_racecheck.log
$T --tool racecheck python -m pytest tests/test_render_options.py -m gpu -k "cornell" -x -q 2>&1 | tail -4 >> gpurun_out/san_racecheck.log
$T --tool synccheck python -m pytest tests/test_gpu_stage_parity.py tests/test_render_options.py -m gpu -k "cornell" -x -q 2>&1 | tail -4 > gpurun_out/san_synccheck.log
tail -n 5 gpurun_out/san_*.log
