# usage: bash tools/gpu/run_full.sh <tag>   — GPU test suite, default bench line, launch list and one full ncu capture of a step
TAG=${1:-x}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit=$?" >> gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
NOX="--no-e2e --no-cpu-baseline --no-breakdown --no-optimise --no-reference-ext"
python bench.py --steps 5 --warmup 3 $NOX > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 5 --warmup 3 $NOX > gpurun_out/${TAG}_ncu_l.log 2>&1
# the full capture runs the step in ONE lane, so that a launch is the whole 720-angle grid the stage timings refer to
# (10 kernels per step: angle table, face normals, vertex transform, face assembly, super lists, forward sweep, replay,
#  face list, walker, vertex backward; 6 for the target's forward + 3 warm-up steps come first)
CTDR_LANES=1 python bench.py --steps 1 --warmup 3 $NOX > gpurun_out/${TAG}_plain1.log 2>&1 && \
CTDR_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 36 -c 10 -o gpurun_out/${TAG}_prof_cfg3 -f python bench.py --steps 1 --warmup 3 $NOX > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log; cut -c1-400 gpurun_out/${TAG}_bench.json
