set -x
(python tools/ab_step.py cfg3 "" "" 2>&1 | tail -2
AB_ANGLES=90 python tools/ab_step.py cfg3 "" 2>&1 | tail -1
AB_ANGLES=90 python tools/ab_step.py northstar "" 2>&1 | tail -1) | tee gpurun_out/v8h_ab.log
