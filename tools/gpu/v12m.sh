set -x
run() { # N tag cfg steps
python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $1 --config $3 --steps $4 --warmup 3 --no-cpu-baseline --no-optimise --no-reference-ext > gpurun_out/${2}_n${1}_${3}.json 2> gpurun_out/${2}_n${1}_${3}.err
tail -c 300 gpurun_out/${2}_n${1}_${3}.json
}
run 8 v12m cfg3 20
run 8 v12m northstar 10
run 4 v12m cfg3 20
run 2 v12m cfg3 20
