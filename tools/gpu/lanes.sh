set -x
(python tools/ab_step.py cfg3 "CTDR_LANES=2" "CTDR_LANES=3" "CTDR_LANES=4" "CTDR_LANES=2" 2>&1 | tail -4
AB_ANGLES=90 python tools/ab_step.py cfg3 "CTDR_LANES=2" "CTDR_LANES=3" "CTDR_LANES=4" 2>&1 | tail -3
AB_ANGLES=180 python tools/ab_step.py cfg3 "CTDR_LANES=2" "CTDR_LANES=3" "CTDR_LANES=4" 2>&1 | tail -3
AB_ANGLES=90 python tools/ab_step.py northstar "CTDR_LANES=2" "CTDR_LANES=3" "CTDR_LANES=4" 2>&1 | tail -3) | tee gpurun_out/lanes.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/lanes_tests.log 2>&1; tail -3 gpurun_out/lanes_tests.log
