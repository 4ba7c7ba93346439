"""oracle/ref_driver.py — TEST INFRASTRUCTURE.  Drive the reference's compiled CUDA extension.

Replays the host steps of ``ctdr/function/rasterizer.py:23-83,97-105`` (bbox, zero-initialised
buffers, pass 1, int32 cumsum -> firstidx, pass 2, backward) around ``oracle/_ref/cuda_diff_fp.so``
(the reference's own kernels, built by ``oracle/build_ref.py``).  Needs a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import build_ref

_mod = None


def module():
    global _mod
    if _mod is None:
        _mod = build_ref.load()
    return _mod


def available() -> bool:
    return torch.cuda.is_available() and module() is not None


def forward(p6, len3, ff, labels, mus, H, W, msv, dtype=torch.float32):
    """Returns dict of CUDA tensors: im, cnt, firstidx, imidx, imwei, bbox (+ inputs)."""
    m = module()
    def t(a, d=dtype):      # NumPy arrays or tensors (already on the device: no round trip through the host)
        a = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        return a.to(d).cuda().contiguous()
    p6, len3, ff, mus = t(p6), t(len3), t(ff), t(mus)
    labels = t(labels, torch.int32)
    B, F = p6.shape[:2]
    pv = p6.view(B, F, 3, 2)
    bbox = torch.cat([pv.min(dim=2)[0], pv.max(dim=2)[0]], dim=2).contiguous()      # :23-32
    im = torch.zeros(B, H, W, 1, dtype=dtype, device="cuda")                        # :36
    cnt = torch.zeros(B, H, W, 1, dtype=torch.int32, device="cuda")                 # :38
    e_f = torch.zeros(0, dtype=dtype, device="cuda")
    e_i = torch.zeros(0, dtype=torch.int32, device="cuda")
    m.forward(p6, ff, bbox, len3, labels, mus, e_f, e_i, cnt, e_i, im, 1, float(msv))   # :51-55
    cumsum = torch.cumsum(cnt.reshape(-1), 0, dtype=torch.int32)                    # :63
    vf = int(cumsum[-1])                                                            # :64
    imwei = torch.zeros(3 * vf, dtype=dtype, device="cuda")                         # :67-68
    imidx = torch.zeros(vf, dtype=torch.int32, device="cuda")
    first = torch.zeros(B * H * W, dtype=torch.int32, device="cuda")                # :71-73
    first[1:] = cumsum[:-1]
    first = first.reshape(B, H, W, 1).contiguous()
    m.forward(p6, ff, bbox, len3, labels, mus, imwei, imidx, cnt, first, im, 0, float(msv))  # :76-80
    torch.cuda.synchronize()
    return dict(im=im, cnt=cnt, firstidx=first, imidx=imidx, imwei=imwei, bbox=bbox, vf=vf,
                p6=p6, len3=len3, ff=ff, labels=labels, mus=mus)


def backward(fwd, grad_im):
    m = module()
    g = grad_im if torch.is_tensor(grad_im) else torch.as_tensor(np.ascontiguousarray(grad_im))
    g = g.to(fwd["im"].dtype).cuda().contiguous()
    dldp = torch.zeros_like(fwd["p6"]); dldf = torch.zeros_like(fwd["len3"]); dldmu = torch.zeros_like(fwd["mus"])
    m.backward(g, fwd["im"], fwd["imidx"], fwd["cnt"], fwd["firstidx"], fwd["imwei"], fwd["p6"], fwd["len3"],
               fwd["labels"], fwd["mus"], fwd["ff"], dldp, dldf, dldmu)               # :101-105
    torch.cuda.synchronize()
    return dldp, dldf, dldmu
