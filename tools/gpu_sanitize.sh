#!/bin/bash
# One GPU: compute-sanitizer memcheck over every kernel family, racecheck + synccheck over the render kernels (tools/sanitize_case.py, small shapes).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py all > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/san_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py render > gpurun_out/san_racecheck_render.log 2>&1; echo "racecheck exit $?" >> gpurun_out/san_racecheck_render.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_case.py render > gpurun_out/san_synccheck_render.log 2>&1; echo "synccheck exit $?" >> gpurun_out/san_synccheck_render.log
tail -5 gpurun_out/san_memcheck.log; tail -5 gpurun_out/san_racecheck_render.log; tail -5 gpurun_out/san_synccheck_render.log
