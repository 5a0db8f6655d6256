#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on the blocked tier, vbatched, inverse, s/c/z and the round-2 kernels;
# racecheck (shared-memory hazards, named barriers, mbarriers) on one shape per kernel family
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "left_looking or blocked or getri or nopiv or vbatched or mid_tier or chain_panel or fused_tail or fused_tier or getrf_batched_prec or getrs_batched_prec or gesv_batched_prec or rbt or gemm or trsm or tall_panel or any_alignment" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/memcheck.log
tail -5 gpurun_out/memcheck.log
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_left_looking_4_warps and 128-128 or test_getrf_blocked_square and 256 or test_left_looking_16_warps and 400 or test_chain_panel_switch and 3-128-128 or test_fused_tail_switch and 2-128-128 or test_fused_tier_full_size_sample or test_getri_fused and 64-64-64 or test_getri_fused and 31-31 or test_getrf_batched_prec and 100-100 and z or test_getrf_batched_prec and 32-32-32 and c or test_getrs_batched_prec and 50-4-111-s or test_tall_panel_structured or test_any_alignment_tma_staging and 255-255-257" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/racecheck.log
tail -5 gpurun_out/racecheck.log
