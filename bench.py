#!/usr/bin/env python
"""bench.py — fwd+bwd detector Gpix/s of the mesh projector (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]

One "step" = one projection + masked-L2 residual + backward sweep over ALL angles of the workload
(angles sharded over the N ranks, mesh replicated) followed by the all-reduce of the gradients —
the data-path of one optimiser iteration (ctdr/optimize.py:162-190).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd detector Gpix/s"
CPU_SAMPLE = dict(angles=2, rows=32)          # bounded sample of the workload for the CPU arm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-optimise", action="store_true")
    ap.add_argument("--optimise-iters", type=int, default=100)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle port; the extension has no CPU path)
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, repeats=1):
    """Time vertex stage + Rasterizer forward (pass 1 + pass 2) + backward of the reference
    algorithm (pixel-parallel, loop over all faces) on CPU_SAMPLE of the workload."""
    from oracle import oracle, vertex_stage
    na, rows = CPU_SAMPLE["angles"], CPU_SAMPLE["rows"]
    H, W = wl["H"], wl["W"]
    na = min(na, wl["nangles"]); rows = min(rows, H)
    idx = np.linspace(0, wl["nangles"], na, endpoint=False).astype(np.int64)
    y0 = (H - rows) // 2
    oracle.lib()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        geom = wl["geom"]
        full_idx = np.arange(wl["nangles"])
        if geom["type"][-3:] == "vec":      # fp_vecs needs the full angle set; project only the sample
            sub = dict(geom, Vectors=geom["Vectors"][idx])
            p6, len3, ff, msv = vertex_stage.vertex_stage(sub, wl["verts"], wl["faces"], np.arange(na), np.float32)
        else:
            p6, len3, ff, msv = vertex_stage.vertex_stage(geom, wl["verts"], wl["faces"], full_idx[idx], np.float32)
        bb = oracle.bbox(p6)
        im = np.zeros((na, H, W, 1), np.float32); cnt = np.zeros((na, H, W, 1), np.int32)
        oracle.forward_pass(p6, ff, bb, len3, wl["labels"], wl["mus"], None, None, cnt, None, im, 1, msv,
                            rows=(y0, y0 + rows))
        cumsum = np.cumsum(cnt.reshape(-1), dtype=np.int32)
        vf = int(cumsum[-1])
        imwei = np.zeros(3 * vf, np.float32); imidx = np.zeros(vf, np.int32)
        first = np.zeros(na * H * W, np.int32); first[1:] = cumsum[:-1]
        first = first.reshape(na, H, W, 1)
        oracle.forward_pass(p6, ff, bb, len3, wl["labels"], wl["mus"], imwei, imidx, cnt, first, im, 0, msv,
                            rows=(y0, y0 + rows))
        g = np.ones_like(im)
        fwd = dict(im=im, cnt=cnt, firstidx=first, imidx=imidx, imwei=imwei)
        oracle.rasterizer_backward(g, fwd, p6, len3, ff, wl["labels"], wl["mus"], rows=(y0, y0 + rows))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    pixels = na * rows * W
    return pixels / best / 1e9, best, f"{na} angles x {rows} detector rows x {W} cols x full mesh ({wl['faces'].shape[0]} faces)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from shapefromprojections_b200 import workloads
    wl = workloads.make_workload(args.config)
    cores = os.cpu_count()
    times = []
    for i in range(args.warmup + args.steps):
        gpix, dt, sample = cpu_reference_sample(wl)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    pixels = min(CPU_SAMPLE["angles"], wl["nangles"]) * min(CPU_SAMPLE["rows"], wl["H"]) * wl["W"]
    value = pixels / (ms / 1e3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {wl['desc']}",
                       "sample": f"reference algorithm (oracle port) on the host cores; each step = a bounded sample "
                                 f"({sample}); pixels are independent in that algorithm, so the sample rate is the full-job rate"},
            "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from shapefromprojections_b200 import _lib, distributed, workloads
    from shapefromprojections_b200.function import rasterizer as R

    rank, world, local_rank = distributed.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.lib()

    wl = workloads.make_workload(args.config)
    H, W, NA = wl["H"], wl["W"], wl["nangles"]
    V, F = wl["verts"].shape[0], wl["faces"].shape[0]
    emu = int(os.environ.get("CTDR_BENCH_EMULATE_WORLD", "0"))    # profiling aid: one rank's shard of an N-way split
    my_angles = distributed.strided_angles(NA, rank if not emu else 0, world if not emu else emu)
    Bl = int(my_angles.numel())
    gd = R.GeometryDevice(wl["geom"], dev)
    verts = torch.as_tensor(wl["verts"]).to(dev)
    faces32 = torch.as_tensor(wl["faces"]).to(torch.int32).to(dev)
    labels = torch.as_tensor(wl["labels"]).to(dev)
    mus = torch.as_tensor(wl["mus"]).to(dev)
    idx = my_angles.to(dev)

    # seeded target: projection of the ground-truth shape (the mesh scaled anisotropically)
    gt = verts * torch.as_tensor(workloads.TARGET_SCALE).to(dev)
    target, _, _ = R.project_forward(gd, gt, faces32, labels, mus, idx)
    phat = torch.empty_like(target)
    xch = distributed.GradientExchange(V, mus.numel(), dev)
    stats = None

    def step():
        xch.zero_()
        R.project_residual_step(gd, verts, faces32, labels, mus, idx, target, phat=phat,
                                grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
        xch.all_reduce()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    pixels = NA * H * W
    value = pixels / (ms_step / 1e3) / 1e9
    loss_sum, n_valid = (float(x) for x in xch.scalars.tolist())

    # ---- end to end: host buffers in, host results out, every step -------------------------------
    e2e = None
    if not args.no_e2e:
        h_verts = wl_pin(torch, wl["verts"]); h_mus = wl_pin(torch, wl["mus"])
        h_idx = my_angles.clone().pin_memory()
        h_target = torch.empty(target.shape, dtype=torch.float32).pin_memory()
        h_target.copy_(target)
        stepper = R.HostStepper(gd, faces32, labels, V, mus.numel(), Bl, windows=8)

        def reduce_fn(flat, out2):
            if world > 1:
                dist.all_reduce(flat); dist.all_reduce(out2)

        def timed_loop(fn, n):
            fn(); barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return 1e3 * float(dt.item()) / n

        n_e2e = max(2, min(args.steps, 5))
        # (1) every input from pinned host memory every step, gradients + loss back to the host
        e2e_ms = timed_loop(lambda: stepper.step(h_verts, h_mus, h_idx, h_target, reduce_fn), n_e2e)
        assert abs(float(stepper.h_out2[0]) - loss_sum) <= 1e-6 * abs(loss_sum) + 1e-9, "e2e result differs"
        h2d = world_sum(torch, dist, world, dev, stepper.h2d_bytes)
        d2h = world_sum(torch, dist, world, dev, stepper.d2h_bytes)

        # (2) the reference's full-batch flow: target uploaded once before the loop (optimize.py:125-127);
        # per step only the mesh parameters / angle indices go up and loss + gradients come back
        h_grad = torch.empty(xch.flat.shape, dtype=torch.float32).pin_memory()
        h_scal = torch.empty(2, dtype=torch.float64).pin_memory()

        def resident_step():
            v = h_verts.to(dev, non_blocking=True); m = h_mus.to(dev, non_blocking=True)
            ix = h_idx.to(dev, non_blocking=True)
            xch.zero_()
            R.project_residual_step(gd, v, faces32, labels, m, ix, target, phat=phat,
                                    grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
            xch.all_reduce()
            h_grad.copy_(xch.flat, non_blocking=True); h_scal.copy_(xch.scalars, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        res_ms = timed_loop(resident_step, n_e2e)
        e2e = {"value": pixels / (e2e_ms / 1e3) / 1e9, "unit": "Gpix/s", "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "what": "pinned host verts/mus/idx_angles/target -> device (target streamed in 8 angle windows on a "
                       "copy stream, overlapped with compute), fused step, all-reduce, gradients + loss -> pinned host",
               "resident_target": {"value": pixels / (res_ms / 1e3) / 1e9, "unit": "Gpix/s", "ms_per_step": res_ms,
                                   "what": "reference full-batch flow: target uploaded once (optimize.py:125-127); per "
                                           "step verts/mus/idx_angles up, gradients + loss down"}}
        del h_target, stepper

    # ---- per-kernel breakdown (staged entry points, CUDA events on the launching stream) ----------
    breakdown = None
    if not args.no_breakdown and rank == 0:
        breakdown = kernel_breakdown(torch, R, gd, verts, faces32, labels, mus, idx, target, phat)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    alg_bytes = workloads.algorithmic_bytes(NA, H, W, V, F, world)
    achieved = alg_bytes / (ms_step / 1e3) / 1e9 / world          # GB/s per GPU
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "what": "algorithmic bytes of one step (8 B/pixel + mesh term) / step time, per GPU; "
                        "dominant kernels listed in `kernels`"}
    line = {"metric": METRIC, "value": value, "unit": "Gpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {wl['desc']}", "angles_per_gpu": Bl, "V": V, "F": F,
                       "step": "projection + masked-L2 residual + backward over all angles + gradient all-reduce",
                       "l2_policy": "inputs larger than L2 (target + sinogram shards >> 126 MB)",
                       "parallelism": f"angles sharded over {world} GPU(s) (rank r owns angles r, r+N, ...), mesh replicated"},
            "roofline": roofline, "clocks": clocks, "gpu_launches": LAUNCHES_PER_STEP * args.steps,
            "loss": loss_sum / max(n_valid, 1.0), "n_valid": n_valid}
    if e2e is not None:
        line["e2e"] = e2e
    if breakdown is not None:
        line["kernels"] = breakdown
        if "raster_step_ms" in breakdown:
            # the two raster sweeps are the step's dominant kernels (42 % / 39 % of it in the ncu launch list);
            # both are timed live here through the staged C-ABI entry points (same kernels, unlinked: inside the
            # fused step the forward also records its face batches, +5 %, and the residual/backward sweep replays
            # them instead of walking its own lists, -25 %)
            tj = None
            tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
            if os.path.exists(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh).get(args.config)
                if not (tj and tj.get("angles") == Bl):
                    tj = None

            def kernel_roofline(name, ms, nbytes, key):
                return {"name": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9,
                        "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / peak, "traffic": tj.get(key) if tj else None,
                        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                          "(profiles/r1_ncu_cfg3_summary.md)"}
            fwd = kernel_roofline("k_raster<MODE_FWD> (forward sweep: 4 B sinogram store per pixel)",
                                  breakdown["raster_forward_ms"], 4 * Bl * H * W, "k_raster<MODE_FWD>")
            bwd = kernel_roofline("k_raster residual + backward sweep (4 B phat + 4 B target per pixel; timed as the staged "
                                  "MODE_BWD launch, same traversal)", breakdown["raster_step_ms"], 8 * Bl * H * W,
                                  "k_raster<MODE_STEP>")
            line["roofline"]["dominant_kernel"] = fwd
            line["roofline"]["second_kernel"] = bwd
            line["roofline"]["traffic"] = fwd["traffic"]
    if not args.no_optimise and world == 1:
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):       # the reference-style modules print; stdout carries the JSON line only
            line["optimise_iteration"] = optimise_iteration(torch, args.optimise_iters)
    if not args.no_cpu_baseline and world == 1:
        gpix, dt, sample = cpu_reference_sample(wl)
        line["cpu_baseline"] = {"value": gpix, "unit": "Gpix/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": sample, "seconds": dt}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


LAUNCHES_PER_STEP = 7   # per angle window: vertex transform, face assemble (+ face and chunk boxes), super-region lists, raster forward (records batches), raster residual/backward replay, its fallback twin, vertex backward


def optimise_iteration(torch, niter):
    """BASELINE.json's second metric, 's per optimise iteration', on config 2: 2bunnyA geometry
    (30 x 192 x 192), icosphere(4) initial mesh (V=2562, F=5120), synthetic star-shaped target with
    eta = 0.3 noise, the reference's default weights (parse_args.py:13-23).  Times ctdr.optimize.run
    (fused data term + sparse regularisers + Adam), one host sync per iteration like the reference."""
    import tempfile
    import types
    from shapefromprojections_b200 import optimize, workloads
    from shapefromprojections_b200.dataset.dataset import SinoDataset
    from shapefromprojections_b200.function import rasterizer as R
    from shapefromprojections_b200.model.vanilla import Model
    wl = workloads.make_workload("cfg1")                      # same geometry numbers as data/2bunnyA
    dev = torch.device("cuda", torch.cuda.current_device())
    gd = R.GeometryDevice(wl["geom"], dev)
    gt_v, gt_f = workloads.star_surface(4)
    labels = torch.as_tensor(np.tile(np.array([1, 0], np.int32), (gt_f.shape[0], 1))).to(dev)
    clean, _, _ = R.project_forward(gd, torch.as_tensor(gt_v).to(dev), torch.as_tensor(gt_f).to(torch.int32).to(dev),
                                    labels, torch.tensor([0.0, 1.0], device=dev), torch.arange(wl["nangles"], device=dev))
    ds = SinoDataset.from_arrays(wl["geom"], clean.cpu().numpy(), nmaterials=2, eta=0.3)
    labels_v = np.ones(wl["verts"].shape[0], dtype=np.int32)
    model = Model((wl["verts"], wl["faces"], labels_v, wl["labels"]), wl["geom"], 2, [0.0, 1.0], 1,
                  wlap=10.0, wflat=0.01).to(dev)
    model.set_mu0(ds.p.to(dev))
    with tempfile.TemporaryDirectory() as tmp:
        args = types.SimpleNamespace(data="2bunnyA", lr=0.01, b=0, wedge=2.0, wlap=10.0, wflat=0.01, verbose=0,
                                     dresult=tmp + "/")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            optimize.run(model, ds, 10, args, fused=True, log_every=10 ** 9)          # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            optimize.run(model, ds, 10 + niter, args, epoch_start=10, fused=True, log_every=10 ** 9)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            # same iteration captured in a CUDA graph (graph build outside the timed region)
            git = optimize.GraphedIteration(model, args, torch.arange(wl["nangles"]), ds.p)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(niter):
                last = git.step().tolist()           # one host sync per iteration, like the reference's loss.item()
            dt_graph = time.perf_counter() - t0
        log = open(tmp + "/log.txt").read().splitlines()
    l2 = [float(l.split("l2_loss:")[1].split()[0]) for l in log if l.startswith("~ ") and "l2_loss" in l]
    return {"workload": "cfg2: 2bunnyA geometry 30x192x192, icosphere(4) V=2562 F=5120, star-shaped target + eta=0.3 noise, "
                        "wlap=10 wedge=2 wflat=0.01 lr=0.01 (parse_args.py defaults)",
            "s_per_iteration": dt / niter, "s_per_iteration_cuda_graph": dt_graph / niter, "l2_after_graphed": last[0],
            "iterations": niter, "l2_first": l2[0] if l2 else None,
            "l2_last": l2[-1] if l2 else None, "mus": [float(x) for x in model.mus.detach().cpu()]}


def wl_pin(torch, arr):
    return torch.as_tensor(np.ascontiguousarray(arr)).pin_memory()


def world_sum(torch, dist, world, dev, v):
    t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


def kernel_breakdown(torch, R, gd, verts, faces32, labels, mus, idx, target, phat, reps=3):
    """Time the stages of one step through the staged C-ABI entry points (same kernels)."""
    from shapefromprojections_b200 import _lib
    out = {}

    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    try:
        p6, len3, ff = R.vertex_forward(gd, verts, faces32, idx)
        out["vertex_forward_ms"] = timed(lambda: R.vertex_forward(gd, verts, faces32, idx))
        out["raster_forward_ms"] = timed(lambda: R.raster_forward(p6, len3, ff, labels, mus, gd.H, gd.W, gd.max_sino_val))
        im, _, st = R.raster_forward(p6, len3, ff, labels, mus, gd.H, gd.W, gd.max_sino_val, want_stats=True)
        g = (2.0 * (im.squeeze(-1) - target)).unsqueeze(-1).contiguous()
        out["raster_backward_ms"] = timed(lambda: R.raster_backward(g, im, p6, len3, ff, labels, mus, gd.max_sino_val))
        out["raster_step_ms"] = out["raster_backward_ms"]
        st = st.tolist()
        out["fragments"] = st[_lib.STAT_FRAGMENTS]; out["candidates"] = st[_lib.STAT_CANDIDATES]
        out["nonpositive_len"] = st[_lib.STAT_NONPOSITIVE_LEN]; out["list_segments"] = st[_lib.STAT_LIST_SEGMENTS]
        out["face_batches"] = st[4]; out["chunk_hits"] = st[5]
        out["full_forward_ms"] = timed(lambda: R.project_forward(gd, verts, faces32, labels, mus, idx))
    except torch.cuda.OutOfMemoryError:
        out["note"] = "staged breakdown skipped (out of memory)"
    return out


if __name__ == "__main__":
    main()
