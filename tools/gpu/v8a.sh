set -x
timeout 200 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/v8a_peer.log 2>&1; tail -15 gpurun_out/v8a_peer.log
python tools/ab_step.py cfg3 "CTDR_FACE_NORMALS=0" "CTDR_FACE_NORMALS=1" "CTDR_FACE_NORMALS=0" "CTDR_FACE_NORMALS=1" 2>&1 | tee gpurun_out/v8a_ab.log | tail -5
AB_ANGLES=90 python tools/ab_step.py cfg3 "CTDR_FACE_NORMALS=0" "CTDR_FACE_NORMALS=1" 2>&1 | tee gpurun_out/v8a_ab90.log | tail -3
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/v8a_tests.log 2>&1; tail -5 gpurun_out/v8a_tests.log
