#!/usr/bin/env python
"""bench.py — fwd+bwd detector Gpix/s of the mesh projector (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]

Default (--config cfg3, the configuration BASELINE.json's metric is quoted on; also cfg1, cfg4, northstar,
cfg3_small): one "step" = one projection + masked-L2 residual + backward sweep over ALL angles of the workload
(angles sharded over the N ranks, mesh replicated) followed by the all-reduce of the gradients — the data path of
one optimiser iteration (ctdr/optimize.py:162-190).
--config cfg5: batched simulation, forward only (run/simulation_cone_beam.py:63-82), the 64 meshes sharded over
the ranks, no collective.  --config cfg2: BASELINE.json's second metric, s per optimise iteration, on the 2bunnyA
geometry for 500 iterations (one GPU).
Prints ONE JSON line on rank 0.  Every number is measured where its key says (no aliases): `kernels.*_ms` are the
stages of the step's kernel sequence over all of the rank's angles on ONE stream (CUDA events between the stages,
ctdr_project_residual_step_timed); the timed step itself runs that sequence in two lanes (two halves of the angles on
two streams), so `ms_per_step` is a little below the sum of the stages; `gpu_launches` is the
library's own launch counter over the timed region.
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
CPU_SAMPLE = dict(angles=2, rows=64)          # bounded sample of the workload for the CPU arm
STEP_CONFIGS = ("cfg1", "cfg3", "cfg3_small", "cfg4", "northstar")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3", choices=list(STEP_CONFIGS) + ["cfg2", "cfg5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-optimise", action="store_true")
    ap.add_argument("--no-reference-ext", action="store_true")
    ap.add_argument("--optimise-iters", type=int, default=100)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """Cores this process may use (affinity mask when the platform has one)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle port; the extension has no CPU path)
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, repeats=1):
    """Time vertex stage + Rasterizer forward (pass 1 + pass 2) + backward of the reference
    algorithm (pixel-parallel, loop over all faces) on CPU_SAMPLE of the workload.
    Returns (Gpix/s, seconds, sample description, OpenMP threads used)."""
    from oracle import oracle, vertex_stage
    lib = oracle.lib()
    # torchrun exports OMP_NUM_THREADS=1 to its workers: set the count explicitly and report what OpenMP uses
    lib.oracle_set_threads(host_threads())
    threads = int(lib.oracle_max_threads())
    na, rows = CPU_SAMPLE["angles"], CPU_SAMPLE["rows"]
    H, W = wl["H"], wl["W"]
    na = min(na, wl["nangles"]); rows = min(rows, H)
    idx = np.linspace(0, wl["nangles"], na, endpoint=False).astype(np.int64)
    y0 = (H - rows) // 2
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
    return (pixels / best / 1e9, best,
            f"{na} angles x {rows} detector rows x {W} cols x full mesh ({wl['faces'].shape[0]} faces)", threads)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from shapefromprojections_b200 import workloads
    name = "cfg1" if args.config == "cfg2" else args.config
    wl = workloads.make_workload(name)
    times, threads, sample = [], 1, ""
    for i in range(args.warmup + args.steps):
        gpix, dt, sample, threads = cpu_reference_sample(wl)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    pixels = min(CPU_SAMPLE["angles"], wl["nangles"]) * min(CPU_SAMPLE["rows"], wl["H"]) * wl["W"]
    value = pixels / (ms / 1e3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {wl['desc']}",
                       "sample": f"reference algorithm (oracle port, OpenMP, {threads} threads) on the host cores; each step = a "
                                 f"bounded sample ({sample}); pixels are independent in that algorithm, so the sample rate is the full-job rate"},
            "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": threads, "kind": "port", "sample": sample,
                             "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
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


class Dist:
    """The little of torch.distributed the bench needs, no-ops on one rank."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor([float(v) for v in values], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                        "sum": self.dist.ReduceOp.SUM}[op])
        return [float(x) for x in t.tolist()]

    def timed_device(self, fn, steps):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the current stream, max over ranks."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.reduce([e0.elapsed_time(e1)], "max")[0] / steps

    def timed_wall(self, fn, steps):
        fn(); self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return 1e3 * self.reduce([time.perf_counter() - t0], "max")[0] / steps


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from shapefromprojections_b200 import _lib, distributed

    rank, world, local_rank = distributed.init_from_env()
    # one process per GPU: keep the rank (and the pinned host buffers it allocates) on the GPU's own socket
    args.numa = distributed.bind_to_gpu_numa(local_rank) if world > 1 and args.impl != "reference" else "one rank: not bound"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.lib()
    D = Dist(torch, dist, world, dev)
    if args.config == "cfg5":
        line = bench_cfg5(args, torch, D, rank, world, local_rank, dev)
    elif args.config == "cfg2":
        line = bench_cfg2(args, torch, D, rank, world, local_rank, dev)
    else:
        line = bench_step(args, torch, dist, D, rank, world, local_rank, dev)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# the headline: fused projection + residual + backward step, angles sharded over the ranks
# ----------------------------------------------------------------------------------------------
def bench_step(args, torch, dist, D, rank, world, local_rank, dev):
    from shapefromprojections_b200 import _lib, distributed, workloads
    from shapefromprojections_b200.function import rasterizer as R
    L = _lib.lib()
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

    def step():
        xch.zero_()
        R.project_residual_step(gd, verts, faces32, labels, mus, idx, target, phat=phat,
                                grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
        xch.all_reduce()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    D.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.ctdr_launch_count()
    ms_step = D.timed_device(step, args.steps)
    launches = (L.ctdr_launch_count() - launches0) + args.steps * xch.kernels_per_exchange(world)
    if xch.peer is not None:
        assert xch.peer.status() == 0, "peer all-reduce gave up waiting for a rank: the sums are invalid"
    if world > 1:
        # the exchange really summed over the ranks: one more step WITHOUT it, local sums added up by a plain
        # all-reduce of a fresh tensor, against what the timed steps left in the exchange buffers
        reduced = xch.scalars.clone(); g_reduced = xch.flat.double().sum(); g_mag = xch.flat.double().abs().sum()
        xch.zero_()
        R.project_residual_step(gd, verts, faces32, labels, mus, idx, target, phat=phat,
                                grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
        chk = torch.cat([xch.scalars, xch.flat.double().sum().reshape(1)])
        dist.all_reduce(chk)
        assert chk[1].item() == reduced[1].item() and abs(chk[0].item() - reduced[0].item()) <= 1e-9 * abs(chk[0].item()), \
            f"gradient exchange did not sum over the ranks: {reduced.tolist()} vs {chk.tolist()}"
        assert abs(chk[2].item() - g_reduced.item()) <= 1e-5 * g_mag.item() + 1e-12, (chk[2].item(), g_reduced.item(), g_mag.item())
        xch.scalars.copy_(reduced)
    clocks = sampler.stop() if rank == 0 else None
    pixels = NA * H * W
    value = pixels / (ms_step / 1e3) / 1e9
    loss_sum, n_valid = xch.loss_sum_and_count()

    # ---- stages of the step as they run inside it (events between the stages), max / min over the ranks ----
    stages = None
    if not args.no_breakdown:
        acc, nrep = {}, 3
        for _ in range(nrep):
            xch.zero_()
            R.project_residual_step(gd, verts, faces32, labels, mus, idx, target, phat=phat, grad_verts=xch.grad_verts,
                                    grad_mus=xch.grad_mus, out2=xch.scalars, stage_ms=acc)
        names = list(_lib.STAGES)
        mine = [acc[n] / nrep for n in names]
        # the exchange on its own: fill + all-reduce behind a barrier (launch + NCCL latency, 0.6 MB at cfg3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        e0.record()
        for _ in range(10):
            xch.zero_(); xch.all_reduce()
        e1.record(); D.barrier()
        mine.append(e0.elapsed_time(e1) / 10); names.append("zero_and_all_reduce")
        mx, mn = D.reduce(mine, "max"), D.reduce(mine, "min")
        stages = {f"{n}_ms": {"max_over_ranks": a, "min_over_ranks": b} for n, a, b in zip(names, mx, mn)}
        stage_ms = dict(zip(names, mx))
        stages["sum_of_stage_max_ms"] = sum(mx)
        stages["note"] = ("stages timed on one stream over all of the rank's angles; the timed step runs the same sequence in "
                          f"{1 if os.environ.get('CTDR_LANES') == '1' else 2} lane(s), overlapping each kernel's tail with the other lane")

    # ---- end to end: host buffers in, host results out, every step -------------------------------
    e2e = None
    if not args.no_e2e:
        h_verts = wl_pin(torch, wl["verts"]); h_mus = wl_pin(torch, wl["mus"])
        h_idx = my_angles.clone().pin_memory()
        h_target = torch.empty(target.shape, dtype=torch.float32).pin_memory()
        h_target.copy_(target)
        stepper = R.HostStepper(gd, faces32, labels, V, mus.numel(), Bl, windows=8)

        def reduce_fn(flat, out2):
            xch.all_reduce_tensors(flat, out2)

        n_e2e = max(2, min(args.steps, 5))
        # (1) every input from pinned host memory every step, gradients + loss back to the host
        e2e_ms = D.timed_wall(lambda: stepper.step(h_verts, h_mus, h_idx, h_target, reduce_fn), n_e2e)
        assert abs(float(stepper.h_out2[0]) - loss_sum) <= 1e-6 * abs(loss_sum) + 1e-9, "e2e result differs"
        h2d = int(D.reduce([stepper.h2d_bytes], "sum")[0])
        d2h = int(D.reduce([stepper.d2h_bytes], "sum")[0])

        # (2) the reference's full-batch flow: target uploaded once before the loop (optimize.py:125-127);
        # per step only the mesh parameters / angle indices go up and loss + gradients come back
        h_grad = torch.empty(xch.flat.shape, dtype=xch.flat.dtype).pin_memory()

        def resident_step():
            v = h_verts.to(dev, non_blocking=True); m = h_mus.to(dev, non_blocking=True)
            ix = h_idx.to(dev, non_blocking=True)
            xch.zero_()
            R.project_residual_step(gd, v, faces32, labels, m, ix, target, phat=phat,
                                    grad_verts=xch.grad_verts, grad_mus=xch.grad_mus, out2=xch.scalars)
            xch.all_reduce()
            h_grad.copy_(xch.flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        res_ms = D.timed_wall(resident_step, n_e2e)
        e2e = {"value": pixels / (e2e_ms / 1e3) / 1e9, "unit": "Gpix/s", "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "what": "pinned host verts/mus/idx_angles/target -> device (target streamed in 8 angle windows on a "
                       "copy stream, overlapped with compute), fused step, all-reduce, gradients + loss -> pinned host",
               "resident_target": {"value": pixels / (res_ms / 1e3) / 1e9, "unit": "Gpix/s", "ms_per_step": res_ms,
                                   "what": "reference full-batch flow: target uploaded once (optimize.py:125-127); per "
                                           "step verts/mus/idx_angles up, gradients + loss down"}}
        del h_target, stepper

    # ---- staged entry points (Rasterizer boundary, unlinked kernels) + counters: context for the stage numbers ----
    staged = None
    if not args.no_breakdown and rank == 0:
        staged = staged_breakdown(torch, R, gd, verts, faces32, labels, mus, idx, target)

    collective = xch.collective
    if xch.peer is not None:
        assert xch.peer.status() == 0, "peer all-reduce gave up waiting for a rank"
        xch.peer.close(); xch.peer = None                     # collective: every rank unmaps before anyone frees
    if rank != 0:
        return None

    peak, peak_src = peaks()
    alg_bytes = workloads.algorithmic_bytes(NA, H, W, V, F, world)
    achieved = alg_bytes / (ms_step / 1e3) / 1e9 / world          # GB/s per GPU
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "what": "algorithmic bytes of one step (8 B/pixel + mesh term) / step time, per GPU; per-kernel "
                        "figures in dominant_kernel / second_kernel (stage times measured inside the step)"}
    line = {"metric": METRIC, "value": value, "unit": "Gpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {wl['desc']}", "angles_per_gpu": Bl, "V": V, "F": F,
                       "step": "projection + masked-L2 residual + backward over all angles + gradient all-reduce",
                       "l2_policy": "inputs larger than L2 (target + sinogram shards >> 126 MB)" if Bl * H * W * 8 > 2 * 126e6
                                    else "inputs smaller than L2: the step streams its own 2 x B x H x W x 4 B of sinogram/target, nothing else is reused between steps",
                       "parallelism": f"angles sharded over {world} GPU(s) (rank r owns angles r, r+N, ...), mesh replicated",
                       "collective": collective, "host_placement": args.numa},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches) * world,
            "gpu_launches_per_rank_per_step": launches / args.steps,
            "loss": loss_sum / max(n_valid, 1.0), "n_valid": n_valid}
    if e2e is not None:
        line["e2e"] = e2e
    if stages is not None:
        line["kernels"] = stages
        tj = traffic_table(args.config, Bl)

        def kernel_roofline(name, ms, nbytes, key):
            return {"name": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9,
                    "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / peak, "traffic": tj.get(key) if tj else None,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/)"}
        fwd = kernel_roofline("k_raster<MODE_FWD, linked> (forward sweep as it runs in the step: 4 B sinogram store per pixel)",
                              stage_ms["raster_forward"], 4 * Bl * H * W, "k_raster<MODE_FWD>")
        bwd = kernel_roofline("k_step_replay + k_raster<MODE_STEP> twin (residual + backward sweep as it runs in the step: "
                              "4 B phat + 4 B target per pixel)", stage_ms["raster_step"], 8 * Bl * H * W,
                              "k_step_replay")
        line["roofline"]["dominant_kernel"] = fwd
        line["roofline"]["second_kernel"] = bwd
        line["roofline"]["traffic"] = fwd["traffic"]
    if staged is not None:
        line["staged_entry_points"] = staged
    if world == 1:
        import contextlib
        if not args.no_reference_ext:
            with contextlib.redirect_stdout(sys.stderr):
                line["reference_cuda_ext"] = reference_cuda_ext(torch)
        if not args.no_optimise:
            with contextlib.redirect_stdout(sys.stderr):       # the reference-style modules print; stdout carries the JSON line only
                line["optimise_iteration"] = optimise_iteration(torch, args.optimise_iters)
        if not args.no_cpu_baseline:
            gpix, dt, sample, threads = cpu_reference_sample(wl)
            line["cpu_baseline"] = {"value": gpix, "unit": "Gpix/s", "cores": threads, "kind": "port",
                                    "sample": sample, "seconds": dt}
    return line


def traffic_table(config, angles):
    """ncu DRAM traffic per launch, kept per round under profiles/ (newest first)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as fh:
                tj = json.load(fh).get(config)
            if tj and tj.get("angles") == angles:
                return tj
    return None


def staged_breakdown(torch, R, gd, verts, faces32, labels, mus, idx, target, reps=3):
    """The same work through the staged (Rasterizer-boundary) C-ABI entry points: unlinked kernels that
    materialise p6/len3/ff; plus the fragment / candidate counters of the forward sweep."""
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
        out["ctdr_vertex_forward_ms"] = timed(lambda: R.vertex_forward(gd, verts, faces32, idx))
        out["ctdr_raster_forward_ms"] = timed(lambda: R.raster_forward(p6, len3, ff, labels, mus, gd.H, gd.W, gd.max_sino_val))
        im, _, st = R.raster_forward(p6, len3, ff, labels, mus, gd.H, gd.W, gd.max_sino_val, want_stats=True)
        g = (2.0 * (im.squeeze(-1) - target)).unsqueeze(-1).contiguous()
        out["ctdr_raster_backward_ms"] = timed(lambda: R.raster_backward(g, im, p6, len3, ff, labels, mus, gd.max_sino_val))
        st = st.tolist()
        out["fragments"] = st[_lib.STAT_FRAGMENTS]; out["candidates"] = st[_lib.STAT_CANDIDATES]
        out["nonpositive_len"] = st[_lib.STAT_NONPOSITIVE_LEN]; out["list_segments"] = st[_lib.STAT_LIST_SEGMENTS]
        out["face_batches"] = st[4]; out["chunk_hits"] = st[5]
        del p6, len3, ff, im, g
        out["ctdr_project_forward_ms"] = timed(lambda: R.project_forward(gd, verts, faces32, labels, mus, idx))
    except torch.cuda.OutOfMemoryError:
        out["note"] = "staged breakdown skipped (out of memory)"
    return out


# ----------------------------------------------------------------------------------------------
# the reference's own CUDA extension on this GPU (north_star: "plus the reference's own CUDA extension on
# one B200 where it builds"): context, like-for-like at the Rasterizer boundary
# ----------------------------------------------------------------------------------------------
def reference_cuda_ext(torch):
    try:
        from oracle import build_ref, time_ref_ext
        mod = build_ref.load()
    except Exception as exc:          # noqa: BLE001  (optional test infrastructure)
        return {"unavailable": f"oracle/_ref could not be loaded: {exc}"}
    if mod is None:
        return {"unavailable": "oracle/_ref/cuda_diff_fp.so is not built (python oracle/build_ref.py in the build container)"}
    return {"what": "reference kernels (ctdr/cuda/diff_fp_for.cu, diff_fp_back.cu compiled unmodified for sm_100a) through the "
                    "host flow of ctdr/function/rasterizer.py:23-105 vs this library, same inputs at the Rasterizer boundary",
            "cases": time_ref_ext.run_cases(mod)}


# ----------------------------------------------------------------------------------------------
# config 5: batched simulation, forward only, meshes sharded over the ranks
# ----------------------------------------------------------------------------------------------
def bench_cfg5(args, torch, D, rank, world, local_rank, dev):
    from shapefromprojections_b200 import _lib, workloads
    from shapefromprojections_b200.function import rasterizer as R
    L = _lib.lib()
    wl = workloads.make_workload("cfg5")
    H, W, NA, NM = wl["H"], wl["W"], wl["nangles"], wl["nmeshes"]
    gd = R.GeometryDevice(wl["geom"], dev)
    mine = list(range(rank, NM, world))                      # 8 meshes per GPU at N = 8; no collective
    meshes = []
    for i in mine:
        v, f = workloads.cfg5_mesh(i)
        meshes.append((torch.as_tensor(v, dtype=torch.float32).to(dev), torch.as_tensor(f).to(torch.int32).to(dev),
                       torch.as_tensor(np.ascontiguousarray(v, dtype=np.float32)).pin_memory()))
    F = int(meshes[0][1].shape[0]); V = int(meshes[0][0].shape[0])
    labels = torch.as_tensor(wl["labels"]).to(dev); mus = torch.as_tensor(wl["mus"]).to(dev)
    idx = torch.arange(NA, device=dev)
    keep = {}

    def step():
        for v, f, _ in meshes:
            keep["im"], _, _ = R.project_forward(gd, v, f, labels, mus, idx)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.ctdr_launch_count()
    ms_step = D.timed_device(step, args.steps)
    launches = L.ctdr_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    pixels = NM * NA * H * W
    value = pixels / (ms_step / 1e3) / 1e9
    e2e = None
    if not args.no_e2e:
        # simulation_cone_beam.py:63-82 per mesh: vertices from the host, sinogram back to the host (.cpu().numpy())
        h_out = torch.empty(NA, H, W, dtype=torch.float32).pin_memory()

        def host_step():
            for _, f, hv in meshes:
                v = hv.to(dev, non_blocking=True)
                im, _, _ = R.project_forward(gd, v, f, labels, mus, idx)
                h_out.copy_(im, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        n = max(2, min(args.steps, 3))
        e2e_ms = D.timed_wall(host_step, n)
        e2e = {"value": pixels / (e2e_ms / 1e3) / 1e9, "unit": "Gpix/s", "ms_per_step": e2e_ms, "steps": n,
               "h2d_bytes_per_step": NM * V * 12, "d2h_bytes_per_step": NM * NA * H * W * 4,
               "what": "per mesh: pinned host vertices -> device, fused forward projection, full sinogram -> pinned host "
                       "(simulation_cone_beam.py:82 copies every sinogram to the host): PCIe-bound"}
    if rank != 0:
        return None
    peak, peak_src = peaks()
    alg = 4 * pixels + NM * (12 * V + 12 * F + 8 * F) + 48 * NA * NM
    achieved = alg / (ms_step / 1e3) / 1e9 / world
    line = {"metric": "fwd detector Gpix/s (batched simulation, backprop=False)", "value": value, "unit": "Gpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg5: {wl['desc']}", "meshes_per_gpu": len(mine), "V": V, "F": F,
                       "step": "forward projection of all 64 meshes (one fused call per mesh), no collective",
                       "l2_policy": "each mesh writes a 377 MB sinogram (> 126 MB L2); meshes differ from call to call",
                       "parallelism": f"meshes sharded over {world} GPU(s) (rank r owns meshes r, r+N, ...)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "what": "4 B sinogram store per pixel + mesh term, per GPU"},
            "clocks": clocks, "gpu_launches": int(launches) * world, "gpu_launches_per_rank_per_step": launches / args.steps,
            "sinogram_max": float(keep["im"].max())}
    if e2e is not None:
        line["e2e"] = e2e
    return line


# ----------------------------------------------------------------------------------------------
# config 2: s per optimise iteration (BASELINE.json's second metric), 500 iterations by default
# ----------------------------------------------------------------------------------------------
def bench_cfg2(args, torch, D, rank, world, local_rank, dev):
    if rank != 0:
        return None
    import contextlib
    niter = 500 if args.steps == 10 else max(args.steps, 20)        # the configuration names 500 iterations
    sampler = ClockSampler(local_rank)
    sampler.start()
    with contextlib.redirect_stdout(sys.stderr):
        res = optimise_iteration(torch, niter)
    clocks = sampler.stop()
    return {"metric": "s per optimise iteration", "value": res["s_per_iteration_cuda_graph"], "unit": "s", "n_gpus": 1,
            "steps": niter, "warmup": 10, "ms_per_step": 1e3 * res["s_per_iteration_cuda_graph"], "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": res["workload"], "step": "one full optimiser iteration (fused data term + sparse regularisers + "
                       "Adam + mus clamp) replayed from a CUDA graph, one host sync per iteration for the loss (ctdr/optimize.py:193)",
                       "l2_policy": "the whole problem (30x192x192 sinogram, 2562 vertices) is L2-resident by nature: this "
                                    "metric is launch/latency-bound, not bandwidth-bound", "parallelism": "1 GPU"},
            "e2e": {"value": res["s_per_iteration"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "what": "ctdr.optimize.run (the public loop, eager launches): target uploaded once like the reference "
                            "(optimize.py:125-127), loss read back every iteration"},
            "clocks": clocks, "optimise_iteration": res, "gpu_launches": res["ctdr_launches_graphed"]}


def optimise_iteration(torch, niter):
    """BASELINE.json's second metric, 's per optimise iteration', on config 2: 2bunnyA geometry
    (30 x 192 x 192), icosphere(4) initial mesh (V=2562, F=5120), synthetic star-shaped target with
    eta = 0.3 noise, the reference's default weights (parse_args.py:13-23).  Times ctdr.optimize.run
    (fused data term + sparse regularisers + Adam), one host sync per iteration like the reference."""
    import tempfile
    import types
    from shapefromprojections_b200 import _lib, optimize, workloads
    from shapefromprojections_b200.dataset.dataset import SinoDataset
    from shapefromprojections_b200.function import rasterizer as R
    from shapefromprojections_b200.model.vanilla import Model
    wl = workloads.make_workload("cfg2")
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
            "l2_last": l2[-1] if l2 else None, "mus": [float(x) for x in model.mus.detach().cpu()],
            "ctdr_launches_graphed": int(git.ctdr_launches_per_iteration) * niter}


def wl_pin(torch, arr):
    return torch.as_tensor(np.ascontiguousarray(arr)).pin_memory()


if __name__ == "__main__":
    main()
