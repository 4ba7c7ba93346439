set -x
(for L in "" build/libctdr_lb8.so; do
  CTDR_LIB=$L python tools/time_forward.py cfg3 2>&1 | tail -1
  CTDR_LIB=$L python tools/time_forward.py cfg3 90 2>&1 | tail -1
  CTDR_LIB=$L python tools/time_forward.py northstar 90 2>&1 | tail -1
  CTDR_LIB=$L python tools/time_forward.py cfg5 2>&1 | tail -1
  CTDR_LIB=$L python tools/ab_step.py cfg3 "" 2>&1 | tail -1
  CTDR_LIB=$L AB_ANGLES=90 python tools/ab_step.py northstar "" 2>&1 | tail -1
done) | tee gpurun_out/lb.log
