set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2h_tests.log
python tools/stage_timing.py 150 2>&1 | tee gpurun_out/r2h_stage_timing_step150.log
python tools/stage_timing.py 161 2>&1 | tail -5 | tee -a gpurun_out/r2h_stage_timing_step150.log
