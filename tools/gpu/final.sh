set -x
bash tools/gpu/run_full.sh r2v12
NOX="--no-cpu-baseline --no-optimise --no-reference-ext"
python bench.py --config northstar --steps 5 --warmup 3 $NOX > gpurun_out/v12s_northstar.json 2> gpurun_out/v12s_northstar.err; tail -c 200 gpurun_out/v12s_northstar.json
python bench.py --config cfg5 --steps 3 --warmup 3 > gpurun_out/v12s_cfg5.json 2> gpurun_out/v12s_cfg5.err; tail -c 300 gpurun_out/v12s_cfg5.json
python tools/time_forward.py cfg3 | tee gpurun_out/v12s_fwd.json
python tools/run_estimate_mu.py > gpurun_out/v12s_emu.json 2> gpurun_out/v12s_emu.err; tail -c 400 gpurun_out/v12s_emu.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
