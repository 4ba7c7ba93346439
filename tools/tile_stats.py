"""Binning statistics of a workload for candidate tile shapes: (face, tile) pairs per face, faces per
non-empty tile, batches of 32 per angle and their fill — the quantities the per-batch overhead of
k_raster scales with.  Usage: python tools/tile_stats.py [cfg3|northstar|cfg1]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from shapefromprojections_b200.function import rasterizer as R
from shapefromprojections_b200.workloads import make_workload
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = make_workload(name)
dev = torch.device("cuda")
gd = R.GeometryDevice(w["geom"], dev)
v = torch.as_tensor(w["verts"]).to(dev); f = torch.as_tensor(w["faces"]).to(torch.int32).to(dev)
idx = torch.arange(0, w["nangles"], 10, device=dev)
p6, len3, ff = R.vertex_forward(gd, v, f, idx)
p = p6.view(idx.numel(), -1, 3, 2)
xmin, xmax = p[..., 0].min(-1).values, p[..., 0].max(-1).values
ymin, ymax = p[..., 1].min(-1).values, p[..., 1].max(-1).values
H = W = w["H"]
ok = (xmin >= 0) & (ymin >= 0) & (xmax <= W - 1) & (ymax <= H - 1)
x0, x1 = torch.ceil(xmin).long(), torch.floor(xmax).long()
y0, y1 = torch.ceil(ymin).long(), torch.floor(ymax).long()
ok &= (x0 <= x1) & (y0 <= y1)
B = idx.numel()
print(name, "angles", B, "faces ok per angle", ok.sum().item() / B, "bbox cand per face", ((x1 - x0 + 1) * (y1 - y0 + 1))[ok].float().mean().item())
for tw, th in ((16, 16), (32, 16), (32, 32), (64, 16)):
    nx = (x1 // tw - x0 // tw + 1); ny = (y1 // th - y0 // th + 1)
    pairs = (nx * ny)[ok].sum().item()
    # per-tile face counts
    ntx, nty = (W + tw - 1) // tw, (H + th - 1) // th
    counts = torch.zeros(B * ntx * nty, dtype=torch.int32, device=dev)
    bi = torch.arange(B, device=dev)[:, None].expand_as(x0)
    mx, my = int(nx[ok].max()), int(ny[ok].max())
    for dx in range(mx):
        for dy in range(my):
            sel = ok & (nx > dx) & (ny > dy)
            t = (bi[sel] * nty + (y0[sel] // th + dy)) * ntx + (x0[sel] // tw + dx)
            counts.index_add_(0, t, torch.ones_like(t, dtype=torch.int32))
    nz = counts[counts > 0].float()
    batches = torch.ceil(nz / 32).sum().item()
    print(f"tile {tw}x{th}: pairs/face {pairs / ok.sum().item():.3f} nonempty tiles/angle {nz.numel() / B:.0f} faces/tile {nz.mean().item():.1f} batches/angle {batches / B:.0f} fill {pairs / batches / 32:.3f}")
