#!/usr/bin/env python
"""bench.py — Gcell-steps/s of the fused erosion step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload.  N=1: BASELINE config[1], the 4096^2 virtual-pipe + thermal grid, in its WET
variant (rain every 16 steps) so water, flux, sediment and both
thermal layers are live; a "step" is one iteration of the reference main loop
(src/main.cpp:310-321: rain when due, then Erosion::dispatch_grid).  N>1: the same
per-GPU work, weak scaling: a 4096-column map of 4096*N rows, one 4096-row slab per rank,
one NVLink halo push + device-side flag wait per step (no collective on the data path).
`value` is timed on the device with inputs resident in HBM (K steps between two CUDA events on
the step stream, hg_run_profiled; the same run carries one event pair around every fused step
kernel: roofline.kernel_ms is their average over the timed region); `e2e` runs every step through
the C ABI from pinned HOST buffers (upload H,F,S in the reference's RGBA32F texture
format, step, download H,F,S).  The working set (604 MB per plane set at 4096^2) exceeds
the 126 MB L2, so no flush is needed between timed steps.

--impl reference times the reference's CPU path.  Its numerics are GLSL and the llvmpipe
route BASELINE.json names cannot run here (no GL/EGL/Mesa on this image),
but its own compute shaders compile for the CPU through a C++ shim of the GLSL vocabulary
(oracle/refshader/ -> oracle/_ref/libhg_refshaders.so, kind "reference"): this arm times
them, driven like the reference's main loop, rows of each dispatch over all host threads.
Without that library it falls back to the CPU restatement (oracle/, kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1234.5
DT_TIME = 0.015
RAIN_PERIOD = 16
PREROLL = 64           # untimed setup steps that wet the terrain
ALG_BYTES_PER_CELL = 72    # 9 fp32 read + 9 fp32 written
METRIC = "Gcell-steps/s (fused erosion step)"


def read_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """dram bytes per fused launch from the newest committed ncu summary of the same workload
    (profiles/rNN*_fused_step.json, written by scripts/ncu_summary.py), or None."""
    import glob
    try:
        path = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fused_step.json")))[-1]
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms use every core this process may run on."""
    import ctypes
    n = len(os.sched_getaffinity(0))
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_port(width, rows, steps):
    """The oracle (CPU restatement, kind "port") on a bounded sample: `steps` wet steps of a
    width x rows band, all host threads.  Returns (Gcell-steps/s, threads, description)."""
    import oracle
    threads = _use_all_host_threads()
    w = oracle.World(width, rows, seed=SEED)
    w.gen_heightmap()
    w.rain.period = 4
    for s in range(1, 9):          # wet it (untimed)
        w.step(s * DT_TIME)
    w.rain.period = RAIN_PERIOD
    t0 = time.perf_counter()
    for s in range(steps):
        w.step((9 + s) * DT_TIME)
    dt = time.perf_counter() - t0
    w.close()
    return width * rows * steps / dt / 1e9, threads, f"{steps} wet steps of a {width}x{rows} band of the workload, oracle port, {threads} OpenMP threads"


REF_SAMPLE_N = 512      # the reference steps square maps only (src/main.cpp:210)


def cpu_reference(steps, warm=2):
    """The reference's OWN compute shaders compiled for the CPU (oracle/_ref/libhg_refshaders.so, built from
    /root/reference/glsl by oracle/refshader/build_ref.py; kind "reference"), driven like its main loop
    (oracle/refshaders.py), rows of each dispatch spread over all host threads, on a bounded sample of the
    workload: a 512x512 map of the same kind (generated terrain, pre-wetted, rain every 16 steps).
    Returns (Gcell-steps/s, threads, description) or None when the library is not there."""
    import ctypes
    from oracle import refshaders
    if not refshaders.available(build=False):
        return None
    import oracle
    threads = _use_all_host_threads()
    w = oracle.World(8)        # only for the default settings blocks
    ref = refshaders.RefWorld(REF_SAMPLE_N, oracle.ErosionData.from_buffer_copy(bytes(w.erosion)),
                              oracle.RainData.from_buffer_copy(bytes(w.rain)), oracle.MapSettingsData.from_buffer_copy(bytes(w.map)))
    w.close()
    ref.map.seed = SEED
    ref.gen_heightmap()
    ref.rain.period = 4
    for s in range(1, 9):          # wet it (untimed)
        ref.step(s, s * DT_TIME)
    ref.rain.period = RAIN_PERIOD
    for s in range(9, 9 + warm):
        ref.step(s, s * DT_TIME)
    t0 = time.perf_counter()
    for s in range(9 + warm, 9 + warm + steps):
        ref.step(s, s * DT_TIME)
    dt = time.perf_counter() - t0
    return (REF_SAMPLE_N * REF_SAMPLE_N * steps / dt / 1e9, threads,
            f"{steps} wet steps of a {REF_SAMPLE_N}x{REF_SAMPLE_N} map of the workload, the reference's own GLSL compute shaders compiled for "
            f"the CPU (oracle/_ref), {threads} OpenMP threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    got = cpu_reference(steps, warm=max(args.warmup, 0))
    kind = "reference"
    note = ("the reference's own shaders (GLSL) compiled for the CPU through oracle/refshader/glsl_shim.hpp; the route BASELINE.json names "
            "(the same shaders under Mesa llvmpipe) cannot run on this image (no GL)")
    if got is None:            # oracle/_ref was not built: the CPU restatement instead
        if args.warmup > 0:
            cpu_port(4096, 256, 2)
        got = cpu_port(4096, 256, steps)
        kind, note = "port", "oracle/_ref/libhg_refshaders.so is missing (built from /root/reference by __graft_entry__.build()); this is the CPU restatement"
    val, threads, sample = got
    cells = 4096 * 4096 * max(args.gpus, 1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gcell-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cells / (val * 1e9) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "Gcell-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Gcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": note}
    print(json.dumps(line), flush=True)


def workload_config(n, W=4096, rows=4096):
    mb = 36 * W * rows / 1e6
    return {"workload": f"{W}x{rows * n} virtual-pipe + thermal grid erosion, wet variant (rain period {RAIN_PERIOD}), "
                        f"{'single GPU' + (' (BASELINE config[1])' if (W, rows) == (4096, 4096) else '') if n == 1 else f'{n} row slabs of {rows} rows, NVLink halo push per step'}",
            "map": [W, rows * n], "rows_per_gpu": rows, "rain_period": RAIN_PERIOD, "seed": SEED,
            "l2": f"working set {mb:.0f} MB per plane set per GPU > 126 MB L2, no flush needed",
            "parallelism": f"row-slab x{n}" if n > 1 else "none"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # the other BASELINE configs, for manual runs (the driver uses the defaults): e.g. --width 16384 --rows-per-gpu 2048
    # --gpus 8 is 16384^2 strong-scaled over 8 slabs, --width 65536 --rows-per-gpu 8192 --gpus 8 --e2e-steps 0 is 65536^2
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--rows-per-gpu", type=int, default=4096)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    from hydro_gen_b200 import Context, PinnedBuffer, _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    n = world
    dist = None
    if n > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W, rows = args.width, args.rows_per_gpu
    H = rows * n
    ctx = Context(W, H, device=local, row0=rank * rows, rows=rows)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 4; ctx.set_rain(r)
    if n > 1:
        from hydro_gen_b200 import slabs
        slabs.connect_ring(ctx, dist, n, rank)
    ctx.gen_heightmap()
    if n > 1:
        dist.barrier()

    def barrier():
        ctx.sync()
        if n > 1:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    # untimed setup: wet the terrain (rain every 4 steps), then the benchmark's period
    ctx.run(PREROLL, DT_TIME, DT_TIME, True)
    r.period = RAIN_PERIOD; ctx.set_rain(r)
    t = (PREROLL + 1) * DT_TIME
    ctx.run(args.warmup, t, DT_TIME, True)
    t += args.warmup * DT_TIME
    ctx.far_fetch_count()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = ctx.launch_count
    # the timed region: K steps between two CUDA events on the step stream, plus one event pair around every fused step kernel
    ms, k_ms = ctx.run_profiled(args.steps, t, DT_TIME, True)
    barrier()
    launches = ctx.launch_count - launches0
    t += args.steps * DT_TIME
    clocks = sampler.stop() if sampler else None
    far = ctx.far_fetch_count()
    if n > 1:
        import torch
        tm = torch.tensor([ms], device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    ms_per_step = ms / args.steps
    cells = W * H
    value = cells / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel: its average duration over the timed region (CUDA events around every launch)
    if n > 1:
        import torch
        tk = torch.tensor([k_ms], device="cuda")
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        k_ms = float(tk.item())
    peak, peak_src = read_peak()
    achieved = ALG_BYTES_PER_CELL * W * rows / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "k_fused_ws", "kernel_ms": k_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_cell_step": ALG_BYTES_PER_CELL, "cells_per_launch": W * rows,
                "note": "the kernel is fp32-issue bound, not HBM bound (DESIGN.md §Roofline)"}

    # end to end through the C ABI with HOST buffers: every step uploads H, F, S from pinned RGBA32F
    # images (the reference's texture format), steps, and downloads H, F, S into a second set.
    # hg_step_host_async pipelines consecutive steps over copy streams (PCIe is full duplex).
    fields = (_lib.FIELD_HEIGHTMAP, _lib.FIELD_FLUX, _lib.FIELD_SEDIMENT)
    if args.e2e_steps <= 0:          # manual runs of the very large configs: no host images
        fields = ()
    pins = [PinnedBuffer((rows, W, 4)) for _ in fields]
    pouts = [PinnedBuffer((rows, W, 4)) for _ in fields]
    for f, p in zip(fields, pins):
        ctx.download(f, p.array)
    ins, outs = [p.array for p in pins], [p.array for p in pouts]
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e_ms = float("nan")
    if fields:
        for _ in range(2):      # warm the staging path
            ctx.step_host_async(ins, outs)
        barrier()
        ctx.timer_start()
        for _ in range(e2e_steps):
            ctx.step_host_async(ins, outs)
        e_ms = ctx.timer_stop()     # waits for the last download
        barrier()
    if n > 1:
        import torch
        te = torch.tensor([e_ms], device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
    e2e_val = cells / (e_ms / e2e_steps * 1e-3) / 1e9 if fields else None
    bytes_dir = len(fields) * rows * W * 16
    e2e = {"value": e2e_val, "unit": "Gcell-steps/s", "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
           "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps if fields else None,
           "what": "per step: hg_step_host_async = upload H,F,S from pinned RGBA32F host images, Erosion::dispatch_grid, "
                   "download H,F,S to a second pinned set; consecutive steps pipelined over 3 streams"}
    halo_errors = ctx.slab_errors()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Gcell-steps/s", "n_gpus": n, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n, W, rows),
                "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "far_fetch_cells_per_step": far / args.steps, "halo_errors": halo_errors}
        if n == 1 and not args.no_cpu_baseline:
            got = cpu_reference(150)        # ~10-20 s of CPU work
            if got is not None:
                v, threads, sample = got
                line["cpu_baseline"] = {"value": v, "unit": "Gcell-steps/s", "cores": threads, "kind": "reference", "sample": sample}
            v, threads, sample = cpu_port(W, 256, 100)
            port = {"value": v, "unit": "Gcell-steps/s", "cores": threads, "kind": "port", "sample": sample}
            if got is None:
                line["cpu_baseline"] = port
            else:
                line["cpu_port"] = port         # the optimised CPU restatement (oracle/), for scale
        print(json.dumps(line), flush=True)
    for p in pins + pouts:
        p.free()
    ctx.close()
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
