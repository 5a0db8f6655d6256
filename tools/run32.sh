#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r32_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r32_tests.log
tail -12 gpurun_out/r32_tests.log
