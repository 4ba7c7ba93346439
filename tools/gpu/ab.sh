# usage: bash tools/gpu/ab.sh "<variant A>" "<variant B>" ...   (A/B stage timings on config 3, a 90-angle shard of it, and the north-star mesh)
set -x
python tools/ab_step.py cfg3 "$@" 2>&1 | tee gpurun_out/ab_cfg3.log
AB_ANGLES=90 python tools/ab_step.py cfg3 "$@" 2>&1 | tee gpurun_out/ab_cfg3_90.log
AB_ANGLES=90 python tools/ab_step.py northstar "$@" 2>&1 | tee gpurun_out/ab_northstar_90.log
