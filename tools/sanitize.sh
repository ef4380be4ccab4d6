#!/bin/bash
# tools/sanitize.sh — compute-sanitizer passes over the GPU parity tests (run under gpurun; small cases only, the
# sanitizer slows kernels 10-100x).  Last run: 0 memcheck errors, 0 racecheck hazards, 0 synccheck errors.
set -x
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -k "boundaries or degenerate or cornell or flat or one_tri or two_tri" -x -q
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hit_groups.py tests/test_gpu_trace.py tests/test_gpu_render.py -k "not large and not fullsize" -x -q
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -k "boundaries and 1025 or cornell-default or bunny4-fast_trace" -x -q
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_render.py -k "cornell or denoise" -x -q
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py tests/test_gpu_render.py -k "boundaries and 1025 or cornell-default" -x -q
