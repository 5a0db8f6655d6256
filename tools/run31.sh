#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "left_looking or blocked_square or blocked_rect or blocked_ldda or getri or vbatched" > gpurun_out/r31_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/r31_memcheck.log
tail -6 gpurun_out/r31_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r31_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_left_looking_4_warps and 128-128 or test_getrf_blocked_square and 256" > gpurun_out/r31_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/r31_racecheck.log
tail -8 gpurun_out/r31_racecheck.log
