timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_tall_panel_switch and 1-512-512 or test_tall_panel_switch and 1-200-130 or test_tall_panel_structured or test_any_alignment_tma_staging and 255-255-257 or test_any_alignment_tma_staging and 101-77-101" > gpurun_out/racecheck2.log 2>&1; echo "racecheck rc $?" >> gpurun_out/racecheck2.log
tail -4 gpurun_out/racecheck2.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "tall_panel or any_alignment" > gpurun_out/memcheck2.log 2>&1; echo "memcheck rc $?" >> gpurun_out/memcheck2.log
tail -4 gpurun_out/memcheck2.log
