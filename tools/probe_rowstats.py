"""Diagnostic (needs a library built with -DCTDR_PROBE_ROWSTATS, passed via CTDR_LIB): share of (face, row) pairs of the
replay kernel that take the exact path, by cause."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapefromprojections_b200 import _lib, workloads, distributed
from shapefromprojections_b200.function import rasterizer as R
wl = workloads.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
dev = torch.device("cuda", 0)
gd = R.GeometryDevice(wl["geom"], dev)
v = torch.as_tensor(wl["verts"]).to(dev); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
lab = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
idx = distributed.strided_angles(wl["nangles"], 0, 8).to(dev)
target, _, _ = R.project_forward(gd, v * torch.as_tensor(workloads.TARGET_SCALE).to(dev), f32, lab, mus, idx)
h = ctypes.CDLL(_lib.LIB_PATH)
h.ctdr_debug_rowstats(None, 1)
R.project_residual_step(gd, v, f32, lab, mus, idx, target)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 4)()
h.ctdr_debug_rowstats(out, 0)
rows, doubt, margin, empty = list(out)
print("batches", rows, "irregular faces", doubt, "-", margin, "irregular batches", empty)
