"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck` (tiny shapes)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from shapefromprojections_b200 import _lib  # noqa: E402
from shapefromprojections_b200 import cuda_diff_fp as compat  # noqa: E402
from shapefromprojections_b200.function import rasterizer as R  # noqa: E402

dev = lambda a, d=None: (torch.as_tensor(np.ascontiguousarray(a)).to(d) if d else torch.as_tensor(np.ascontiguousarray(a))).cuda()
for c in (cases.soup_case(0, B=2, F=700, H=70, W=101, n_mus=4), cases.soup_case(2, B=1, F=20000, H=64, W=64)):
    p6, len3, ff, lab, mus = dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"])
    im, cnt, _ = R.raster_forward(p6, len3, ff, lab, mus, c["H"], c["W"], c["msv"], want_cnt=True, want_stats=True)
    g = torch.randn_like(im)
    for flags in (0, _lib.FLAG_DETERMINISTIC, _lib.FLAG_NO_PREFILTER):
        R.raster_backward(g, im, p6, len3, ff, lab, mus, c["msv"], flags=flags)
    B, F = c["p6"].shape[:2]
    cum = torch.cumsum(cnt.reshape(-1), 0, dtype=torch.int32); vf = int(cum[-1])
    first = torch.zeros(B * c["H"] * c["W"], dtype=torch.int32, device="cuda"); first[1:] = cum[:-1]
    imwei = torch.zeros(3 * vf, device="cuda"); imidx = torch.zeros(vf, dtype=torch.int32, device="cuda")
    bbox = torch.zeros(B, F, 4, device="cuda")
    compat.forward(p6, ff, bbox, len3, lab, mus, imwei, imidx, cnt, first.reshape(B, c["H"], c["W"], 1), im, 0, c["msv"])
for kind, geom in (("par", cases.parallel_geometry(5, 72)), ("cone", cases.cone_geometry(5, 72))):
    verts, faces, labels, mus = cases.two_material_mesh(1)
    gd = R.GeometryDevice(geom, "cuda")
    v, f32, lab, m = dev(verts), dev(faces, torch.int32), dev(labels), dev(mus)
    idx = torch.arange(5, device="cuda")
    tgt = torch.rand(5, gd.H, gd.W, device="cuda")
    for flags in (0, _lib.FLAG_DETERMINISTIC):
        phat, gv, gm, out2 = R.project_residual_step(gd, v, f32, lab, m, idx, tgt, flags=flags)
        R.project_backward(gd, v, f32, lab, m, idx, tgt, phat, flags=flags)
torch.cuda.synchronize()
print("sanitize_small done")
