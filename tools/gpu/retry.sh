# usage: bash tools/gpu/retry.sh <gpus> <timeout_s> <max_tries> -- <command>   — retry a gpurun call while the pod answers busy (exit 3)
G=$1; T=$2; N=$3; shift 4
for i in $(seq 1 $N); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
