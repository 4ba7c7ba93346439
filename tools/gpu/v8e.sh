set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/v8e_tests.log 2>&1; tail -3 gpurun_out/v8e_tests.log
(AB_ANGLES=90 python tools/ab_step.py cfg3 "CTDR_REPLAY_TILES=4" "" 2>&1 | tail -2
AB_ANGLES=180 python tools/ab_step.py cfg3 "CTDR_REPLAY_TILES=4" "" 2>&1 | tail -2
AB_ANGLES=90 python tools/ab_step.py northstar "CTDR_REPLAY_TILES=4" "" 2>&1 | tail -2
python tools/ab_step.py cfg3 "" 2>&1 | tail -1) | tee gpurun_out/v8e_ab.log
