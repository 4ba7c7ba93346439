"""Config 5 of BASELINE.json: batched simulation — 64 distinct 20k-triangle meshes projected at 360
angles x 512 x 512 (run/simulation_cone_beam.py-style throughput), forward only, one GPU.
Meshes: torus(r_i, R_i, 100, 100) with r_i in [0.15, 0.3], R_i in [0.35, 0.55] and a random rigid
rotation from np.random.default_rng(i).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapefromprojections_b200.function import rasterizer as R  # noqa: E402
from shapefromprojections_b200.utils import meshgen, util_sino  # noqa: E402


def main(nmesh=64):
    dev = torch.device("cuda", 0)
    geom = util_sino.geom_2vec(util_sino.make_geometry("cone", 360, 512, 512, 3.0 / 512, dso=10.0, dod=4.0))
    gd = R.GeometryDevice(geom, dev)
    idx = torch.arange(360, device=dev)
    meshes = []
    for i in range(nmesh):
        rng = np.random.default_rng(i)
        v, f = meshgen.torus(rng.uniform(0.15, 0.3), rng.uniform(0.35, 0.55), 100, 100)
        v = v @ meshgen.rigid_rotation(i).T + 1e-4
        meshes.append((torch.as_tensor(v, dtype=torch.float32).to(dev), torch.as_tensor(f).to(torch.int32).to(dev)))
    labels = torch.as_tensor(np.tile(np.array([1, 0], np.int32), (20000, 1))).to(dev)
    mus = torch.tensor([0.0, 1.0], device=dev)
    out = torch.empty(360, 512, 512, device=dev)

    def sweep():
        total = 0.0
        for v, f in meshes:
            im, _, _ = R.project_forward(gd, v, f, labels, mus, idx)
            total += 0.0
        return im
    sweep(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); im = sweep(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    pix = nmesh * 360 * 512 * 512
    print(json.dumps({"config": "cfg5: 64 x (torus 20k triangles, cone 360 x 512 x 512), forward only, 1 GPU",
                      "ms_total": ms, "ms_per_mesh": ms / nmesh, "forward_Gpix_per_s": pix / ms / 1e6,
                      "last_sinogram_max": float(im.max())}))


if __name__ == "__main__":
    main()
