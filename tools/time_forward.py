"""Device time of the fused forward projector alone (ctdr_project_forward: FP.forward without backprop, the simulation
scripts' path) on a named workload.   python tools/time_forward.py cfg3 [angles]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapefromprojections_b200 import workloads, distributed
from shapefromprojections_b200.function import rasterizer as R

name = sys.argv[1]
wl = workloads.make_workload(name)
dev = torch.device("cuda", 0)
gd = R.GeometryDevice(wl["geom"], dev)
v = torch.as_tensor(wl["verts"]).to(dev); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
lab = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
nang = int(sys.argv[2]) if len(sys.argv) > 2 else wl["nangles"]
idx = distributed.strided_angles(wl["nangles"], 0, wl["nangles"] // nang).to(dev)
for _ in range(3):
    im, _, _ = R.project_forward(gd, v, f32, lab, mus, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    im, _, _ = R.project_forward(gd, v, f32, lab, mus, idx)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"workload": name, "angles": nang, "lib": os.environ.get("CTDR_LIB", "in-tree"), "project_forward_ms": round(e0.elapsed_time(e1) / n, 4),
                  "checksum": float(im.double().sum())}))
