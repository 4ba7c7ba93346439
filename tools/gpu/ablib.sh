# usage: bash tools/gpu/ablib.sh <lib> [<lib> ...]   — stage timings of config 3 (720 and 90 angles) with alternative builds of the library ("" = in-tree)
for L in "" "$@"; do
  echo "== lib: ${L:-in-tree}"
  CTDR_LIB=$L python tools/ab_step.py cfg3 "" "" 2>&1 | tail -2
  CTDR_LIB=$L AB_ANGLES=90 python tools/ab_step.py cfg3 "" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/ablib.log
