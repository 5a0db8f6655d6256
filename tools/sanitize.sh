#!/bin/bash
# compute-sanitizer over the GPU parity tests of the blocked tier (memcheck) and two shapes under racecheck
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "left_looking or blocked or getri or nopiv or vbatched or mid_tier" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/memcheck.log
tail -5 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_left_looking_4_warps and 128-128 or test_getrf_blocked_square and 256 or test_left_looking_16_warps and 400" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/racecheck.log
tail -5 gpurun_out/racecheck.log
