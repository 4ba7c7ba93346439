"""Diagnostic: raster backward vs the fp64 oracle on one cfg3 angle; where does the error sit?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import oracle
from shapefromprojections_b200 import workloads
from shapefromprojections_b200.function import rasterizer as R

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
angle = int(sys.argv[2]) if len(sys.argv) > 2 else 90
wl = workloads.make_workload(name)
gd = R.GeometryDevice(wl["geom"], "cuda")
v = torch.as_tensor(wl["verts"]).cuda(); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).cuda()
lab = torch.as_tensor(wl["labels"]).cuda(); mus = torch.as_tensor(wl["mus"]).cuda()
idx = torch.tensor([angle], device="cuda")
p6, len3, ff = R.vertex_forward(gd, v, f32, idx)
H, W, msv = gd.H, gd.W, gd.max_sino_val
im, cnt, st = R.raster_forward(p6, len3, ff, lab, mus, H, W, msv, want_cnt=True, want_stats=True)
g = torch.randn(1, H, W, 1, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
dp, dl, dm, sb = R.raster_backward(g, im, p6, len3, ff, lab, mus, msv, want_stats=True)
torch.cuda.synchronize()
print("fragments fwd", int(st[1]), "bwd", int(sb[1]), "invalid pixels", int(((im.double() < -1e-5) | (im > msv)).sum()))
c = dict(p6=p6.cpu().numpy(), len3=len3.cpu().numpy(), ff=ff.cpu().numpy(), labels=wl["labels"], mus=wl["mus"])
d = np.float64
fwd = oracle.rasterizer_forward(c["p6"].astype(d), c["len3"].astype(d), c["ff"].astype(d), c["labels"], c["mus"].astype(d), H, W, msv, True, fast=True)
gp64, gl64, gm64 = oracle.rasterizer_backward(g.cpu().numpy().astype(d), fwd, c["p6"].astype(d), c["len3"].astype(d), c["ff"].astype(d), c["labels"], c["mus"].astype(d))
fwd32 = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], H, W, msv, True, fast=True)
gp32, gl32, gm32 = oracle.rasterizer_backward(g.cpu().numpy(), fwd32, c["p6"], c["len3"], c["ff"], c["labels"], c["mus"])
print("same cover 32/64:", np.array_equal(fwd["cnt"], fwd32["cnt"]))
for nme, a, b64, b32 in (("grad_p6", dp, gp64, gp32), ("grad_len3", dl, gl64, gl32), ("grad_mus", dm, gm64, gm32)):
    a = a.cpu().numpy().astype(d)
    print(nme, "rel err vs 64:", np.linalg.norm(a - b64) / np.linalg.norm(b64), " oracle32 vs 64:", np.linalg.norm(b32 - b64) / np.linalg.norm(b64))
a = dp.cpu().numpy().astype(d)[0]; b = gp64[0]
e = np.linalg.norm(a - b, axis=1); n = np.linalg.norm(b, axis=1)
order = np.argsort(-e)[:12]
P = c["p6"][0]
print("total err^2", (e ** 2).sum(), "top-12 share", (e[order] ** 2).sum() / (e ** 2).sum())
for f in order:
    x = P[f, 0::2]; y = P[f, 1::2]
    m, p_, n_, q = x[1] - x[0], y[1] - y[0], x[2] - x[0], y[2] - y[0]
    k3 = m * q - n_ * p_
    print(f"face {f}: err {e[f]:.3e} |g| {n[f]:.3e} k3 {k3:.4g} bbox x[{x.min():.3f},{x.max():.3f}] y[{y.min():.3f},{y.max():.3f}] ours {a[f]} ref {b[f]}")
rel = e / np.maximum(n, 1e-30)
print("faces with rel err > 1e-3:", int((rel > 1e-3).sum()), "of", int((n > 0).sum()), "; median rel", np.median(rel[n > 0]))
