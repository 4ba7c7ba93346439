# usage: bash tools/gpu/scale.sh <N> <tag> [config]   — one bench line on N GPUs with per-stage timings
N=$1; TAG=$2; CFG=${3:-cfg3}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --config $CFG --steps 20 --warmup 5 --no-cpu-baseline --no-optimise --no-reference-ext > gpurun_out/${TAG}_n${N}_${CFG}.json 2> gpurun_out/${TAG}_n${N}_${CFG}.err
tail -c 2500 gpurun_out/${TAG}_n${N}_${CFG}.json
