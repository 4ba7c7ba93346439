set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/v8f_tests.log 2>&1; tail -3 gpurun_out/v8f_tests.log
(python tools/ab_step.py cfg3 "CTDR_LANES=1" "CTDR_LANES=2" "CTDR_LANES=1" "CTDR_LANES=2" 2>&1 | tail -4
AB_ANGLES=90 python tools/ab_step.py cfg3 "CTDR_LANES=1" "CTDR_LANES=2" 2>&1 | tail -2
AB_ANGLES=180 python tools/ab_step.py cfg3 "CTDR_LANES=1" "CTDR_LANES=2" 2>&1 | tail -2
AB_ANGLES=90 python tools/ab_step.py northstar "CTDR_LANES=1" "CTDR_LANES=2" 2>&1 | tail -2) | tee gpurun_out/v8f_ab.log
