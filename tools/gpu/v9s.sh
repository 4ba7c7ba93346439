set -x
NOX="--no-cpu-baseline --no-optimise --no-reference-ext"
python bench.py --config northstar --steps 5 --warmup 3 $NOX > gpurun_out/v9s_northstar.json 2> gpurun_out/v9s_northstar.err; tail -c 200 gpurun_out/v9s_northstar.json
python bench.py --config cfg4 --steps 2 --warmup 3 $NOX --no-e2e > gpurun_out/v9s_cfg4.json 2> gpurun_out/v9s_cfg4.err; tail -c 200 gpurun_out/v9s_cfg4.json
python bench.py --config cfg5 --steps 3 --warmup 3 > gpurun_out/v9s_cfg5.json 2> gpurun_out/v9s_cfg5.err; tail -c 300 gpurun_out/v9s_cfg5.json
python bench.py --config cfg2 > gpurun_out/v9s_cfg2.json 2> gpurun_out/v9s_cfg2.err; tail -c 400 gpurun_out/v9s_cfg2.json
python bench.py --config cfg1 --steps 20 --warmup 5 $NOX > gpurun_out/v9s_cfg1.json 2> gpurun_out/v9s_cfg1.err; tail -c 200 gpurun_out/v9s_cfg1.json
python tools/run_estimate_mu.py > gpurun_out/v9s_emu.json 2> gpurun_out/v9s_emu.err; tail -c 400 gpurun_out/v9s_emu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v9s_ref.json 2> gpurun_out/v9s_ref.err; tail -c 600 gpurun_out/v9s_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
