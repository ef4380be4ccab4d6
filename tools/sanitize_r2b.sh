#!/bin/bash
# tools/sanitize_r2b.sh — compute-sanitizer over what the second half of round 2 added or rewrote: k_lbvh_fit / k_lbvh_exits (shared-memory
# rounds, 64-bit exchange climb), the radix scatter (shared atomics, five barriers per tile), the packed denoiser, the persistent trace
# kernels with two steps per phase selection.  Small cases only.
mkdir -p gpurun_out
T="timeout 900 compute-sanitizer --error-exitcode 9"
$T --tool memcheck python -m pytest tests/test_gpu_build.py tests/test_golden.py tests/test_gpu_trace.py -m gpu -k "not 65537 and not 300001 and not soup_70k" -x -q 2>&1 | tail -4 > gpurun_out/sanb_memcheck.log
$T --tool memcheck python -m pytest tests/test_gpu_render.py -m gpu -k "denoise" -x -q 2>&1 | tail -4 >> gpurun_out/sanb_memcheck.log
$T --tool racecheck python -m pytest tests/test_gpu_build.py -m gpu -k "fused and (257 or 513 or 4097) or boundaries and (511 or 1025)" -x -q 2>&1 | tail -4 > gpurun_out/sanb_racecheck.log
$T --tool racecheck python -m pytest tests/test_gpu_render.py -m gpu -k "denoise" -x -q 2>&1 | tail -4 >> gpurun_out/sanb_racecheck.log
$T --tool synccheck python -m pytest tests/test_gpu_build.py -m gpu -k "fused and (257 or 4097)" -x -q 2>&1 | tail -4 > gpurun_out/sanb_synccheck.log
tail -n 6 gpurun_out/sanb_*.log
