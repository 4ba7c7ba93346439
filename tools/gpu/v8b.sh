set -x
nvidia-smi topo -m 2>&1 | head -8
timeout 300 python -m pytest tests/test_gpu_multi_rank.py -x -q -s > gpurun_out/v8b_mr.log 2>&1; tail -12 gpurun_out/v8b_mr.log
bash tools/gpu/scale.sh 2 v8b_peer | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['config'].get('collective'), d['kernels']['zero_and_all_reduce_ms'])"
CTDR_PEER_ALLREDUCE=0 bash tools/gpu/scale.sh 2 v8b_nccl | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['config'].get('collective'), d['kernels']['zero_and_all_reduce_ms'])"
tail -3 gpurun_out/v8b_peer_n2_cfg3.err
