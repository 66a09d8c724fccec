#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the GeoSeries hot path.

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on): N_POINTS uniform random
points `.contains()`-joined against 10 000 64-vertex star polygons, one such
batch per GPU (weak scaling: every rank owns its own 100 M-point row range, the polygon side is
broadcast from rank 0 over NCCL — the only exchange step this path has).

One "step" = one pass of the hot path over one batch: [N>1: NCCL broadcast of the polygon coordinates]
+ polygon index build + the point-in-polygon kernel over all of the rank's points [+ N>1: per-polygon
hit histogram and its all-reduce].  `value` = points processed by all ranks / max-over-ranks device time,
inputs resident in HBM.  `e2e` = the same join through the C ABI from pinned HOST buffers (H2D of the
points and D2H of the ids inside the timed region, chunked and overlapped).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
  python bench.py --impl reference      # the CPU restatement of the reference path on the host cores

The CPU oracle (oracle/) is used here only for `cpu_baseline` and `--impl reference`; the measured GPU
path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "geometries/s"
N_POLYGONS = 10_000
POLY_GRID = 100
N_VERT = 64
COORD_BYTES = 16  # one f64 xy pair: the algorithmic bytes per point


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any"""
    p = os.path.join(ROOT, "profiles", "pip_query_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(n_sample: int, threads: int = 0):
    """Time the CPU restatement (oracle, OpenMP all cores = the 'Rayon path' stand-in) on a bounded sample."""
    from geopolars_b200 import synth
    from oracle import oracle as og

    xy, ro, go = synth.star_polygons(N_POLYGONS, POLY_GRID, 10.0, N_VERT)
    polys = og.OGArray(og.POLYGON, xy, geom_off=go, ring_off=ro)
    pts = og.gen_uniform_points(2, 0, n_sample, 1000.0)
    if threads <= 0:
        # every core this process may run on; torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = threads
    og.contains_join(polys, pts[: min(n_sample, 100_000)], True, threads)  # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    first, _ = og.contains_join(polys, pts, True, threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, cores, dt, int((first >= 0).sum())


def best_cpu_threads(probe_points: int = 4_000_000) -> int:
    """all logical CPUs or one per physical core, whichever runs the CPU join faster on a small probe
    (the ring walk is latency bound; hyper-threads can hurt)"""
    all_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, best_rate = all_threads, 0.0
    for th in sorted({all_threads, max(1, all_threads // 2)}, reverse=True):
        rate, _, _, _ = cpu_baseline_sample(probe_points, th)
        if rate > best_rate:
            best, best_rate = th, rate
    return best


def run_reference(args, out=sys.stdout):
    """--impl reference: the reference's CPU path (restated: oracle/geo_oracle.c, kind 'port') on host cores."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return 0
    n_sample = args.ref_sample
    best_threads = best_cpu_threads(min(n_sample, 4_000_000))
    for _ in range(args.warmup):
        cpu_baseline_sample(min(n_sample, 500_000), best_threads)
    vals, times = [], []
    cores = best_threads
    for _ in range(args.steps):
        v, cores, dt, _hits = cpu_baseline_sample(n_sample, best_threads)
        vals.append(v)
        times.append(dt)
    total_t = sum(times)
    value = n_sample * args.steps / total_t
    sample = f"{n_sample} of {args.points} points per step x {N_POLYGONS} polygons, bbox-grid candidates + exact test, OpenMP static row chunks"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "geometries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.points} random points .contains() x {N_POLYGONS} {N_VERT}-vertex polygons (BASELINE configs[1])",
                   "note": "reference arithmetic lives in un-vendored Rust crates (geo 0.27); this is the C restatement oracle/geo_oracle.c"},
        "cpu_baseline": {"value": value, "unit": "geometries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "geometries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


def main():
    # exactly ONE line goes to the real stdout (the JSON); anything libraries print (e.g. NCCL's version
    # banner when NCCL_DEBUG=VERSION) is sent to stderr instead
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        return _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=100_000_000, help="points per GPU (BASELINE configs[1]: 100M)")
    ap.add_argument("--ref-sample", type=int, default=64_000_000, help="points per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=32_000_000, help="points of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--prebuilt-index", action="store_true", help="exclude the polygon index build from the step")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as dist

    from geopolars_b200 import GeoArrowArray, GeometryType, synth
    from geopolars_b200 import engine as E

    world = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local = _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the geopolars_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.points
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = E.Context(local, stream.cuda_stream)
        # ---- inputs resident in HBM -------------------------------------------------------------
        pts = torch.empty((n, 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_gen_uniform_points(ctx._h, 2, rank * n, n, 1000.0, pts.data_ptr()))
        n_pc = N_POLYGONS * (N_VERT + 1)
        poly_xy = torch.empty((n_pc, 2), dtype=torch.float64, device=dev)
        ring_off = torch.arange(N_POLYGONS + 1, dtype=torch.int64, device=dev) * (N_VERT + 1)
        geom_off = torch.arange(N_POLYGONS + 1, dtype=torch.int64, device=dev)
        if rank == 0:
            xy, _, _ = synth.star_polygons(N_POLYGONS, POLY_GRID, 10.0, N_VERT)  # host: libm cos/sin, see synth.py
            poly_xy.copy_(torch.from_numpy(xy))
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        counts = torch.zeros(N_POLYGONS, dtype=torch.int64, device=dev)
        stream.synchronize()

        def make_index():
            polys = ctx.wrap_device(GeometryType.POLYGON, N_POLYGONS, n_pc, poly_xy.data_ptr(), geom_off_ptr=geom_off.data_ptr(),
                                    ring_off_ptr=ring_off.data_ptr(), n_rings=N_POLYGONS, keepalive=(poly_xy, geom_off, ring_off))
            return E.PipIndex(polys)

        kernel_ms = []
        state = {"idx": None}

        def step(timed_kernel: bool):
            if world > 1:
                dist.broadcast(poly_xy, src=0)  # the broadcast-join's one exchange step (NCCL over NVLink)
            if not args.prebuilt_index or state["idx"] is None:
                if state["idx"] is not None:
                    state["idx"].free()
                state["idx"] = make_index()
            idx = state["idx"]
            if timed_kernel:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            idx.query_device(pts.data_ptr(), n, ids.data_ptr())
            if timed_kernel:
                e1.record(stream)
                kernel_ms.append((e0, e1))
            if world > 1:
                counts.zero_()
                E.check(ctx.lib.gpl_join_histogram(ctx._h, ids.data_ptr(), n, counts.data_ptr(), N_POLYGONS, E.GPL_DEVICE))
                dist.all_reduce(counts)

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # sampled from the warm-up on: the timed region alone can be shorter than one sample period
        for _ in range(max(args.warmup, 3)):
            step(False)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = ctx.launch_count
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        for _ in range(args.steps):
            step(True)
        t_end.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = ctx.launch_count - launches0
        total_ms = t_start.elapsed_time(t_end)
        clocks = sampler.stop() if rank == 0 else None
        k_ms = [a.elapsed_time(b) for a, b in kernel_ms]
        hits = int((ids >= 0).sum().item())

        # ---- e2e: host buffers through the C ABI, copies inside the timed region --------------------
        e2e = None
        if not args.no_e2e:
            host_pts = torch.empty((n, 2), dtype=torch.float64, pin_memory=True)
            host_ids = torch.empty(n, dtype=torch.int32, pin_memory=True)
            host_pts.copy_(pts)
            torch.cuda.synchronize()
            poly_host = poly_xy.cpu().numpy()
            ro_h, go_h = ring_off.cpu().numpy(), geom_off.cpu().numpy()

            def e2e_step():
                arr = GeoArrowArray.polygons(poly_host, ro_h, go_h)
                d_polys = ctx.upload(arr)  # H2D of the polygon side (10.4 MB) from host memory
                idx = E.PipIndex(d_polys)
                idx.query_host_pipelined(host_pts.data_ptr(), n, host_ids.data_ptr())
                return idx

            for _ in range(2):
                e2e_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            assert int((host_ids >= 0).sum().item()) == hits, "e2e ids differ from the resident run"
            e2e = {"seconds": e2e_s, "h2d": n * COORD_BYTES + n_pc * COORD_BYTES + 8 * 2 * (N_POLYGONS + 1), "d2h": n * 4}

    # ---- reduce over ranks (max time) ------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, e2e["seconds"] if e2e else 0.0, float(launches), float(statistics.mean(k_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_max, launches, k_mean = t[0].item(), t[1].item(), int(t[2].item()), t[3].item()
        if e2e:
            e2e["seconds"] = e2e_max
    else:
        k_mean = statistics.mean(k_ms)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    peak, peak_src = measured_peaks()
    algo_bytes = n * COORD_BYTES + n_pc * COORD_BYTES  # coordinate bytes read by one launch
    achieved = algo_bytes / (k_mean * 1e-3) / 1e9
    traffic = recorded_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "geometries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{n} random points per GPU .contains() x {N_POLYGONS} {N_VERT}-vertex polygons (BASELINE configs[1]; SURVEY.md config 2 generator)",
            "step": ("NCCL broadcast of polygon coords + " if world > 1 else "") + ("" if args.prebuilt_index else "polygon index build + ")
                    + "point-in-polygon kernel over all points" + (" + hit histogram all-reduce" if world > 1 else ""),
            "l2": "inputs (1.6 GB of points per GPU) exceed the 126 MB L2; no flush needed",
            "parallelism": f"row-range partition of points over {world} GPU(s), polygon side replicated",
            "hit_rate": hits / n,
            "kernel_ms": k_mean,
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "kernel": "k_pip_stream<LEAN> (+ k_pip_deferred)",
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "read_plus_write_GBps": (algo_bytes + 4 * n) / (k_mean * 1e-3) / 1e9},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = {"value": world * n * args.steps / e2e["seconds"], "unit": "geometries/s", "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * e2e["seconds"] / args.steps,
                       "path": "gpl_array_from_buffers(host) + gpl_pip_index_build + gpl_contains_join_host (pinned host points -> pinned host ids)"}
    if not args.no_cpu:
        v, cores, dt, _ = cpu_baseline_sample(args.cpu_sample, best_cpu_threads())
        line["cpu_baseline"] = {"value": v, "unit": "geometries/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample} of {n} points x {N_POLYGONS} polygons in {dt:.2f} s, oracle/geo_oracle.c OpenMP (bbox grid + exact test)"}
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
