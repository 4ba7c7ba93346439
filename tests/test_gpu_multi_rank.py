"""GPU, two ranks over NCCL: the angle-sharded residual step + gradient all-reduce equals the one-rank step.

Needs two GPUs on the box (``gpurun --gpus 2``); skipped on a single-GPU box.  One process per GPU, launched with
torch.distributed.run like bench.py (rendezvous on 127.0.0.1), rank r owns the angles r, r+2, ...
(distributed.strided_angles); after ``GradientExchange.all_reduce`` every rank must hold the gradients, the loss
sum and the valid-pixel count of the full angle set, to fp32 reassociation (1e-5), and its sinogram shard must be
bit-identical to the corresponding slices of the one-rank sinogram.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from shapefromprojections_b200 import distributed, workloads
from shapefromprojections_b200.function import rasterizer as R
rank, world, local = distributed.init_from_env("nccl")
assert world == 2 and dist.get_backend() == "nccl"
dev = torch.device("cuda", local)
wl = workloads.make_workload("cfg3_small")                      # cone 48 x 512 x 512, torus 20k faces
gd = R.GeometryDevice(wl["geom"], dev)
v = torch.as_tensor(wl["verts"]).to(dev); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
lab = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
NA = wl["nangles"]
full = torch.arange(NA, device=dev)
target_full, _, _ = R.project_forward(gd, v * torch.as_tensor(workloads.TARGET_SCALE).to(dev), f32, lab, mus, full)
# one-rank reference, computed redundantly on every rank
ph1, gv1, gm1, o1 = R.project_residual_step(gd, v, f32, lab, mus, full, target_full)
# sharded: this rank's angles only, then ONE grouped all-reduce
mine = distributed.strided_angles(NA, rank, world).to(dev)
xch = distributed.GradientExchange(v.shape[0], mus.numel(), dev)
xch.zero_()
ph, _, _, _ = R.project_residual_step(gd, v, f32, lab, mus, mine, target_full[mine].contiguous(),
                                      grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
local_copy = xch.storage.clone()
xch.all_reduce()
torch.cuda.synchronize()
# the exchange ran as the library's own kernel over NVLink peer memory (CUDA IPC between the two processes) ...
assert xch.peer is not None, "peer all-reduce was not set up: " + xch.collective
assert xch.peer.status() == 0
# ... and agrees bit for bit with the NCCL all-reduce of the same local buffers (two ranks: a + b either way)
nccl = distributed.GradientExchange(v.shape[0], mus.numel(), dev, peer=False)
nccl.storage.copy_(local_copy)
nccl.all_reduce()
torch.cuda.synchronize()
assert torch.equal(nccl.storage, xch.storage), "peer all-reduce differs from NCCL"
for _ in range(3):                                              # repeated use: both payload slots, epochs in step
    xch.storage.copy_(local_copy); xch.all_reduce()
torch.cuda.synchronize()
assert torch.equal(nccl.storage, xch.storage) and xch.peer.status() == 0
assert torch.equal(ph, ph1[mine]), "sinogram shard differs from the one-rank slices"
assert xch.scalars[1].item() == o1[1].item(), (xch.scalars.tolist(), o1.tolist())
assert abs(xch.scalars[0].item() - o1[0].item()) <= 1e-9 * o1[0].item()
ev = ((xch.grad_verts - gv1).norm() / gv1.norm()).item(); em = ((xch.grad_mus - gm1).norm() / gm1.norm()).item()
assert ev <= 1e-5 and em <= 1e-5, (ev, em)
assert gv1.norm().item() > 0
gv, gm, loss = xch.mean_gradients()
assert abs(loss.item() - o1[0].item() / o1[1].item()) <= 1e-9 * loss.item()
xch.peer.close()
dist.barrier(); dist.destroy_process_group()
sys.stdout.write(f"rank {{rank}} ok grad_verts rel {{ev:.2e}} grad_mus rel {{em:.2e}}\n"); sys.stdout.flush()
"""


def test_two_rank_nccl_step_equals_one_rank(cuda_lib, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29741", str(script)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
    print(res.stdout[-400:])
