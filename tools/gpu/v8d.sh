set -x
for L in "" build/libctdr_short.so build/libctdr_short2.so; do
  echo "== lib: ${L:-in-tree}"
  CTDR_LIB=$L python tools/ab_step.py cfg3 "" "" 2>&1 | tail -2
  CTDR_LIB=$L AB_ANGLES=90 python tools/ab_step.py cfg3 "" "CTDR_REGION_TILES=8" 2>&1 | tail -2
  CTDR_LIB=$L AB_ANGLES=90 python tools/ab_step.py northstar "" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/v8d_ab.log
AB_ANGLES=180 python tools/ab_step.py cfg3 "" "CTDR_REGION_TILES=8" 2>&1 | tail -2 | tee -a gpurun_out/v8d_ab.log
CTDR_LIB=build/libctdr_short2.so timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/v8d_tests.log 2>&1; tail -3 gpurun_out/v8d_tests.log
