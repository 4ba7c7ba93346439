"""A/B timing of the residual step's stages under tuning knobs (CTDR_* environment, re-read via ctdr_tuning_reload).

    python tools/ab_step.py cfg3 "CTDR_REPLAY_STAGE=0" "CTDR_REPLAY_STAGE=1 CTDR_GRAD_PER_VERTEX=0" ...
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapefromprojections_b200 import _lib, workloads, distributed
from shapefromprojections_b200.function import rasterizer as R

name = sys.argv[1]
variants = sys.argv[2:] or [""]
dev = torch.device("cuda", 0)
wl = workloads.make_workload(name)
gd = R.GeometryDevice(wl["geom"], dev)
v = torch.as_tensor(wl["verts"]).to(dev); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
lab = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
nang = int(os.environ.get("AB_ANGLES", wl["nangles"]))
idx = distributed.strided_angles(wl["nangles"], 0, wl["nangles"] // nang).to(dev)
target, _, _ = R.project_forward(gd, v * torch.as_tensor(workloads.TARGET_SCALE).to(dev), f32, lab, mus, idx)
phat = torch.empty_like(target)
xch = distributed.GradientExchange(v.shape[0], mus.numel(), dev)
ref = None
for var in variants:
    keys = []
    for kv in var.split():
        k, val = kv.split("="); os.environ[k] = val; keys.append(k)
    _lib.reload_tuning()
    for _ in range(3):
        xch.zero_(); R.project_residual_step(gd, v, f32, lab, mus, idx, target, phat=phat, grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
    torch.cuda.synchronize()
    acc = {}
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        xch.zero_(); R.project_residual_step(gd, v, f32, lab, mus, idx, target, phat=phat, grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / n
    for _ in range(3):
        xch.zero_(); R.project_residual_step(gd, v, f32, lab, mus, idx, target, phat=phat, grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars, stage_ms=acc)
    gv = xch.grad_verts.clone()
    if ref is None:
        ref = gv
    rel = float((gv - ref).norm() / ref.norm())
    print(json.dumps({"variant": var, "step_ms": round(total, 3), **{k: round(val / 3, 3) for k, val in acc.items()}, "grad_rel_vs_first": rel}), flush=True)
    for k in keys:
        del os.environ[k]
_lib.reload_tuning()
