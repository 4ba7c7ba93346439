set -x
bash tools/gpu/scale.sh 8 v8c_peer > /dev/null
CTDR_PEER_ALLREDUCE=0 bash tools/gpu/scale.sh 8 v8c_nccl > /dev/null
bash tools/gpu/scale.sh 8 v8c_peer northstar > /dev/null
bash tools/gpu/scale.sh 4 v8c_peer > /dev/null
tail -2 gpurun_out/v8c_peer_n8_cfg3.err
