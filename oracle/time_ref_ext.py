"""oracle/time_ref_ext.py — TEST INFRASTRUCTURE (context numbers, not the product).

Times the reference's own CUDA extension (oracle/_ref/cuda_diff_fp.so, the kernels of ctdr/cuda compiled
unmodified for sm_100a by oracle/build_ref.py) on one B200, through the host flow of
ctdr/function/rasterizer.py:23-105 (bbox, pass 1, cumsum + vf sync, CSR allocation, pass 2, backward),
next to this repository's kernels at the same Rasterizer boundary and on the same inputs.
Config 1 in full; config 3 on a subset of angles (the reference loops over all faces for every pixel:
7.6e13 face tests per pass for the full config).  Prints one JSON line per case.

    python -m oracle.time_ref_ext            (on a GPU box; needs oracle/_ref built)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import build_ref  # noqa: E402
from shapefromprojections_b200.function import rasterizer as R  # noqa: E402
from shapefromprojections_b200.workloads import make_workload  # noqa: E402


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def reference_fwd_bwd(m, p6, len3, ff, labels, mus, H, W, msv, g):
    B, F = p6.shape[:2]
    pv = p6.view(B, F, 3, 2)
    bbox = torch.cat([pv.min(dim=2)[0], pv.max(dim=2)[0]], dim=2).contiguous()      # rasterizer.py:23-32
    im = torch.zeros(B, H, W, 1, device="cuda"); cnt = torch.zeros(B, H, W, 1, dtype=torch.int32, device="cuda")
    e_f = torch.zeros(0, device="cuda"); e_i = torch.zeros(0, dtype=torch.int32, device="cuda")
    m.forward(p6, ff, bbox, len3, labels, mus, e_f, e_i, cnt, e_i, im, 1, float(msv))        # :51-55
    cumsum = torch.cumsum(cnt.reshape(-1), 0, dtype=torch.int32)                              # :63
    vf = int(cumsum[-1])                                                                      # :64 (host sync)
    imwei = torch.zeros(3 * vf, device="cuda"); imidx = torch.zeros(vf, dtype=torch.int32, device="cuda")
    first = torch.zeros(B * H * W, dtype=torch.int32, device="cuda"); first[1:] = cumsum[:-1]
    first = first.reshape(B, H, W, 1).contiguous()
    m.forward(p6, ff, bbox, len3, labels, mus, imwei, imidx, cnt, first, im, 0, float(msv))   # :76-80
    dldp = torch.zeros_like(p6); dldf = torch.zeros_like(len3); dldmu = torch.zeros_like(mus)
    m.backward(g, im, imidx, cnt, first, imwei, p6, len3, labels, mus, ff, dldp, dldf, dldmu)  # :101-105
    return im, dldp, dldf, dldmu


def ours_fwd_bwd(p6, len3, ff, labels, mus, H, W, msv, g):
    im, _, _ = R.raster_forward(p6, len3, ff, labels, mus, H, W, msv)
    dldp, dldf, dldmu = R.raster_backward(g, im, p6, len3, ff, labels, mus, msv)
    return im, dldp, dldf, dldmu


def run_cases(m, cases=(("cfg1", None, 5), ("cfg3", 4, 1))):
    """[(workload, angles or None = all, repetitions of the reference)] -> list of result dicts."""
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []
    for name, nang, reps in cases:
        wl = make_workload(name)
        gd = R.GeometryDevice(wl["geom"], dev)
        v = torch.as_tensor(wl["verts"]).to(dev); f32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
        labels = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
        nang = wl["nangles"] if nang is None else nang
        idx = torch.arange(0, wl["nangles"], wl["nangles"] // nang, device=dev)[:nang]
        p6, len3, ff = R.vertex_forward(gd, v, f32, idx)
        H, W, msv = gd.H, gd.W, gd.max_sino_val
        g = torch.randn(nang, H, W, 1, device=dev, generator=torch.Generator(dev).manual_seed(0))
        t_ref, o_ref = timed(lambda: reference_fwd_bwd(m, p6, len3, ff, labels, mus, H, W, msv, g), reps)
        t_new, o_new = timed(lambda: ours_fwd_bwd(p6, len3, ff, labels, mus, H, W, msv, g), max(reps, 5))
        pix = nang * H * W
        same = bool(torch.equal(o_ref[0], o_new[0]))
        gerr = float((o_ref[1] - o_new[1]).norm() / o_ref[1].norm())
        out.append({"case": f"{name}: {nang} angle(s) x {H} x {W}, F={f32.shape[0]}, Rasterizer boundary, fwd (pass 1 + pass 2) + bwd",
                    "reference_cuda_extension_ms": t_ref, "this_repo_ms": t_new, "ratio": t_ref / t_new,
                    "reference_Gpix_per_s": pix / t_ref / 1e6, "this_repo_Gpix_per_s": pix / t_new / 1e6,
                    "sinogram_bit_identical": same, "grad_p6_rel_diff": gerr})
    return out


def main():
    m = build_ref.load()
    if m is None:
        print(json.dumps({"unavailable": "oracle/_ref/cuda_diff_fp.so is not built"}))
        return
    for res in run_cases(m):
        print(json.dumps(res))


if __name__ == "__main__":
    main()
