"""oracle/gen_ref_golden.py — TEST INFRASTRUCTURE.  Golden vectors from the REFERENCE kernels.

Run on a B200 box (``gpurun -- python oracle/gen_ref_golden.py``): executes the reference's own
compiled extension (oracle/_ref, built from /root/reference/ctdr/cuda by oracle/build_ref.py) on
the seeded cases of tests/cases.py and writes ``gpurun_out/golden/ref_raster_<case>.npz`` holding
its outputs (im, cnt, fragment list, float and double gradients).  Copy those files to
``tests/golden/``; tests/test_oracle_golden.py then pins the CPU oracle against them without a GPU.
Inputs are regenerated from the seeds by the tests, so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import ref_driver, vertex_stage  # noqa: E402

GOLDEN_CASES = {
    "soup": lambda: cases.soup_case(11, B=2, F=600, H=48, W=56, n_mus=3, reference_safe=True),
    "sphere": lambda: _mesh(cases.parallel_geometry(3, 64), cases.sphere_mesh(2)),
    "two_material_cone": lambda: _mesh(cases.cone_geometry(3, 72), cases.two_material_mesh(1)),
}


def _mesh(geom, mesh):
    verts, faces, labels, mus = mesh
    n = len(geom["ProjectionAngles"]) if "ProjectionAngles" in geom else geom["Vectors"].shape[0]
    p6, len3, ff, msv = vertex_stage.vertex_stage(geom, verts, faces, np.arange(n), np.float32)
    return dict(p6=p6, len3=len3, ff=ff, labels=labels, mus=mus, H=geom["DetectorRowCount"],
                W=geom["DetectorColCount"], msv=msv)


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    assert ref_driver.available(), "reference extension not built (oracle/_ref) or no GPU"
    for name, make in GOLDEN_CASES.items():
        c = make()
        g = np.random.default_rng(5).normal(size=(c["p6"].shape[0], c["H"], c["W"], 1)).astype(np.float32)
        f32 = ref_driver.forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"])
        b32 = ref_driver.backward(f32, g)
        f64 = ref_driver.forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"], torch.float64)
        b64 = ref_driver.backward(f64, g)
        cpu = lambda t: t.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, f"ref_raster_{name}.npz"),
                            im=cpu(f32["im"]), cnt=cpu(f32["cnt"]), firstidx=cpu(f32["firstidx"]),
                            imidx=cpu(f32["imidx"]), imwei=cpu(f32["imwei"]),
                            dldp=cpu(b32[0]), dldf=cpu(b32[1]), dldmu=cpu(b32[2]),
                            im64=cpu(f64["im"]), cnt64=cpu(f64["cnt"]),
                            dldp64=cpu(b64[0]), dldf64=cpu(b64[1]), dldmu64=cpu(b64[2]), grad_im=g)
        print(name, "vf", f32["vf"], "im max", float(f32["im"].max()))


if __name__ == "__main__":
    main()
