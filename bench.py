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

Beside the headline the line carries the other BASELINE configs, each measured the same way
(device-timed, state resident, own clock samples) after the headline's contexts are released:
  N=1: "north_star_16384" (config 3 at one GPU, where the north star states its roofline target) and
       "droplets_8192_4Mi" (config 4: Erosion::dispatch_particle with 4 Mi droplets on 8192^2, for the
       reference's default hmap_dims 1024 and for hmap_dims = map);
  N>1: "slab_parity" (every rank steps a small slab with NVLink halo pushes AND the whole map on its own
       GPU and compares its rows bit for bit: cross-process parity of the IPC path), "cfg3_16384_strong"
       (16384^2 split N ways) and, at N=8, "cfg5_65536" (65536^2 on 8 GPUs with t1 = ONE GPU stepping one
       65536x8192 slab alone, the weak-scaling reference SURVEY.md §8d defines).

--impl reference times the reference's CPU path.  Its numerics are GLSL and the llvmpipe
route BASELINE.json names cannot run here (no GL/EGL/Mesa on this image),
but its own compute shaders compile for the CPU through a C++ shim of the GLSL vocabulary
(oracle/refshader/ -> oracle/_ref/libhg_refshaders.so, kind "reference"): this arm times
them, driven like the reference's main loop, rows of each dispatch over all host threads, on
the 4096^2 map of config[1] itself (one rank's slab of the N-GPU workload; the reference steps
square maps only).  `ms_per_step` is what was measured on that map, never an extrapolation.
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
DROPLET_BYTES = 240        # per droplet-step: 80 B droplet state + 4 x 20 B gathered + 4 x 20 B reduced
DROPLET_GRID_BYTES = 56    # per cell-step of the grid tail in droplet mode: H.rgb 24 B + M 32 B
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


def ncu_traffic(W, rows):
    """dram bytes per fused launch from the newest committed ncu summary (profiles/rNN*_fused_step.json, written by
    scripts/ncu_summary.py from one `ncu --set full` capture of this command) and where it came from.  Only reported
    when that capture was taken on this run's map size; otherwise (None, why)."""
    import glob
    try:
        path = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fused_step.json")))[-1]
        with open(path) as f:
            js = json.load(f)
        cells = int(js.get("cells_per_launch", 4096 * 4096))     # summaries older than round 2 were all taken at 4096^2
        src = {"file": os.path.relpath(path, ROOT), "cells_per_launch": cells, "kernel": js.get("kernel")}
        if cells != W * rows:
            src["note"] = f"capture is of another map size ({cells} cells per launch, this run {W * rows}): not reported"
            return None, src
        return js.get("dram_bytes_per_launch"), src
    except Exception:
        return None, None


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


def cpu_reference(n, steps, warm=2):
    """The reference's OWN compute shaders compiled for the CPU (oracle/_ref/libhg_refshaders.so, built from
    /root/reference/glsl by oracle/refshader/build_ref.py; kind "reference"), driven like its main loop
    (oracle/refshaders.py), rows of each dispatch spread over all host threads, on an n x n map of the workload
    (generated terrain, pre-wetted, rain every 16 steps).  The untimed setup (terrain + 8 wetting steps) runs on the
    oracle port, which reproduces those shaders bit for bit (tests/test_refshaders.py) and is 6x faster.
    Returns (Gcell-steps/s, threads, description, seconds per step) or None when the library is not there."""
    from oracle import refshaders
    if not refshaders.available(build=False):
        return None
    import oracle
    threads = _use_all_host_threads()
    w = oracle.World(n, seed=SEED)
    w.gen_heightmap()
    w.rain.period = 4
    for s in range(1, 9):          # wet it (untimed)
        w.step(s * DT_TIME)
    ref = refshaders.RefWorld(n, oracle.ErosionData.from_buffer_copy(bytes(w.erosion)),
                              oracle.RainData.from_buffer_copy(bytes(w.rain)), oracle.MapSettingsData.from_buffer_copy(bytes(w.map)))
    for name, f in (("heightmap", oracle.FIELD_H), ("flux", oracle.FIELD_F), ("velocity", oracle.FIELD_V), ("sediment", oracle.FIELD_S)):
        getattr(ref, name).read[...] = w.field(f)
    w.close()
    ref.rain.period = RAIN_PERIOD
    for s in range(9, 9 + warm):
        ref.step(s, s * DT_TIME)
    t0 = time.perf_counter()
    for s in range(9 + warm, 9 + warm + steps):
        ref.step(s, s * DT_TIME)
    dt = time.perf_counter() - t0
    return (n * n * steps / dt / 1e9, threads,
            f"{steps} wet steps of a {n}x{n} map of the workload, the reference's own GLSL compute shaders compiled for "
            f"the CPU (oracle/_ref), {threads} OpenMP threads", dt / steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    # config[1]'s own 4096^2 map (about 1 s per step on 16 host cores) while the run stays within a few minutes;
    # a longer request steps a 1024^2 map of the same kind and says so
    n = 4096 if steps + max(args.warmup, 0) <= 100 else 1024
    got = cpu_reference(n, steps, warm=max(args.warmup, 0))
    kind = "reference"
    note = ("the reference's own shaders (GLSL) compiled for the CPU through oracle/refshader/glsl_shim.hpp; the route BASELINE.json names "
            "(the same shaders under Mesa llvmpipe) cannot run on this image (no GL)")
    if got is None:            # oracle/_ref was not built: the CPU restatement instead
        if args.warmup > 0:
            cpu_port(4096, 256, 2)
        v, threads, sample = cpu_port(4096, 256, steps)
        got = (v, threads, sample, None)
        n = 4096
        kind, note = "port", "oracle/_ref/libhg_refshaders.so is missing (built from /root/reference by __graft_entry__.build()); this is the CPU restatement"
    val, threads, sample, s_per_step = got
    cfg = workload_config(args.gpus)
    cfg["reference_map"] = [n, n]
    cfg["reference_map_note"] = (("config[1]'s own 4096x4096 map" if n == 4096 else "a 1024x1024 map of the same kind (steps + warmup > 100)")
                                 + ("" if args.gpus == 1 else f": ONE rank's slab of the {args.gpus}-GPU weak-scaling workload (the reference steps square maps on one device; its rate does not depend on the row count)"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gcell-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": s_per_step * 1e3 if s_per_step is not None else n * n / (val * 1e9) * 1e3,
            "ms_per_step_is": f"measured wall time per step of the {n}x{n} map this arm stepped",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
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


class Rig:
    """rank / world / device of this process and the few collectives the bench needs (NCCL, N > 1 only)"""

    def __init__(self, gpus):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != gpus and self.world != 1:
            raise SystemExit(f"--gpus {gpus} but WORLD_SIZE={self.world}")
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.sync()
        if self.dist:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        if not self.dist:
            return v
        import torch
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def make_grid(rig, W, rows, n_slabs=None, alone=False):
    """This rank's slab of a W x (rows * n) map: created, connected to its peers, terrain generated.
    alone=True: a whole W x rows map on this GPU, no peers (t1 of the weak-scaling configs)."""
    from hydro_gen_b200 import Context, slabs
    n = 1 if alone else (n_slabs or rig.world)
    rank = 0 if alone else rig.rank
    ctx = Context(W, rows * n, device=rig.local, row0=rank * rows, rows=rows)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 4; ctx.set_rain(r)
    if n > 1:
        slabs.connect_ring(ctx, rig.dist, n, rank)
    ctx.gen_heightmap()
    if n > 1:
        rig.barrier(ctx)
    return ctx


def time_grid(rig, ctx, W, rows, n, steps, warmup, sample_clocks=True, collective=True):
    """Pre-wet (PREROLL steps, rain every 4), warm up, then time `steps` main-loop iterations on the device:
    whole run between two CUDA events on the step stream, one event pair around every fused step kernel
    (hg_run_profiled); max over ranks.  Returns the figures of one bench entry."""
    ctx.run(PREROLL, DT_TIME, DT_TIME, True)
    r = ctx.get_rain(); r.period = RAIN_PERIOD; ctx.set_rain(r)
    t = (PREROLL + 1) * DT_TIME
    ctx.run(warmup, t, DT_TIME, True)
    t += warmup * DT_TIME
    ctx.far_fetch_count()
    sampler = ClockSampler(rig.local) if rig.rank == 0 and sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    rig.barrier(ctx) if collective else ctx.sync()
    launches0 = ctx.launch_count
    ms, k_ms = ctx.run_profiled(steps, t, DT_TIME, True)
    rig.barrier(ctx) if collective else ctx.sync()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    far = ctx.far_fetch_count()
    if collective:
        ms, k_ms = rig.reduce(ms), rig.reduce(k_ms)
    peak, peak_src = read_peak()
    cells = W * rows * n
    achieved = ALG_BYTES_PER_CELL * W * rows / (k_ms * 1e-3) / 1e9
    return {"value": cells / (ms / steps * 1e-3) / 1e9, "ms_per_step": ms / steps, "kernel_ms": k_ms, "steps": steps, "warmup": warmup,
            "achieved": achieved, "peak": peak, "peak_source": peak_src, "frac": achieved / peak, "launches": launches,
            "clocks": clocks, "far_per_step": far / steps, "t_end": t + steps * DT_TIME}


def extra_entry(W, rows, n, m, what):
    return {"workload": what, "map": [W, rows * n], "n_gpus": n, "value": m["value"], "unit": "Gcell-steps/s", "ms_per_step": m["ms_per_step"],
            "steps": m["steps"], "warmup": m["warmup"],
            "roofline": {"bound": "hbm", "achieved": m["achieved"], "peak": m["peak"], "unit": "GB/s", "frac": m["frac"], "kernel": "k_fused_ws",
                         "kernel_ms": m["kernel_ms"], "cells_per_launch": W * rows, "algorithmic_bytes_per_cell_step": ALG_BYTES_PER_CELL},
            "clocks": m["clocks"], "far_fetch_cells_per_step": m["far_per_step"]}


def droplet_config(rig, steps=40, warm=20):
    """BASELINE config 4: Erosion::dispatch_particle (droplet move + erode + thermal x2 + smoothing with the momentum
    map) with 4 Mi droplets on 8192^2, should_rain = 1, device-timed, for the reference's default hmap_dims (1024:
    every droplet lives in a 1022^2 corner, ~4 per cell) and for hmap_dims = map (sparse)."""
    from hydro_gen_b200 import Context, _lib
    N, COUNT = 8192, 4 * 1024 * 1024
    peak, _ = read_peak()
    out = {"workload": "Erosion::dispatch_particle, 4194304 droplets on 8192x8192 (BASELINE config 4), should_rain=1",
           "algorithmic_bytes": {"per_droplet_step": DROPLET_BYTES, "per_cell_step_of_the_grid_tail": DROPLET_GRID_BYTES,
                                 "per_dispatch": DROPLET_BYTES * COUNT + DROPLET_GRID_BYTES * N * N}}
    for key, hmap in (("hmap_dims_1024", 1024), ("hmap_dims_8192", 8192)):
        ctx = Context(N, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES, device=rig.local)
        m = ctx.get_map(); m.seed = SEED; m.hmap_dims[0], m.hmap_dims[1] = hmap, hmap; ctx.set_map(m)
        ctx.gen_heightmap()
        ctx.run(warm, DT_TIME, DT_TIME, True)
        ctx.sync()
        sampler = ClockSampler(rig.local)
        sampler.start()
        time.sleep(0.25)
        ctx.timer_start()
        ctx.run(steps, (warm + 1) * DT_TIME, DT_TIME, True)
        ms = ctx.timer_stop() / steps
        clocks = sampler.stop()
        gbs = (DROPLET_BYTES * COUNT + DROPLET_GRID_BYTES * N * N) / (ms * 1e-3) / 1e9
        out[key] = {"ms_per_dispatch": ms, "Mdroplet_steps_per_s": COUNT / ms / 1e3, "G_atomic_updates_per_s": 20 * COUNT / ms / 1e6,
                    "grid_tail_Gcell_steps_per_s": N * N / ms / 1e6, "achieved_GBps": gbs, "roofline_frac": gbs / peak, "steps": steps, "clocks": clocks}
        ctx.close()
    return out


def copy_ceiling(rig, bytes_dir, iters=6):
    """What the host link allows: the same bytes per step as the e2e leg (bytes_dir up + bytes_dir down, pinned memory,
    two streams so both directions overlap), with no kernel and no packing at all, all ranks at once.  e2e cannot be
    faster than this on this host; torch is plumbing here (pinned allocation, streams, events)."""
    import torch
    dev = torch.device("cuda", rig.local)
    n = bytes_dir // 4
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float32, device=dev); d_out = torch.zeros(n, dtype=torch.float32, device=dev)
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    def once():
        with torch.cuda.stream(up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(down):
            h_out.copy_(d_out, non_blocking=True)
    once(); torch.cuda.synchronize(dev)
    rig.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    up.wait_event(e0); down.wait_event(e0)
    for _ in range(iters):
        once()
    torch.cuda.current_stream(dev).wait_stream(up); torch.cuda.current_stream(dev).wait_stream(down)
    e1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    ms = rig.reduce(e0.elapsed_time(e1) / iters)
    rig.barrier()
    return {"ms_per_step_equivalent": ms, "GBps_per_direction_per_gpu": bytes_dir / ms / 1e6,
            "what": f"{bytes_dir} B host->device and {bytes_dir} B device->host per GPU, pinned, overlapped on two streams, no kernels; max over ranks"}


def contracted_build(args):
    """The same headline workload on the opt-in CONTRACTED build of the library (HG_FMAD=1: -fmad=true, results within
    the north star's tolerance of the reference instead of bit-identical; tests/test_fmad_build.py), in a child process:
    what bit-exactness costs.  The headline numbers of this line are the default, bit-exact library's."""
    cmd = [sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup", str(args.warmup), "--no-extras", "--no-cpu-baseline", "--e2e-steps", "0"]
    try:
        res = subprocess.run(cmd, env=dict(os.environ, HG_FMAD="1"), capture_output=True, text=True, timeout=600)
        js = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        return {"what": "libhydrogen_b200_fmad.so (-fmad=true), same workload, same timing; parity: <= 1e-5 per field after one step, not bit-identical",
                "value": js["value"], "unit": js["unit"], "ms_per_step": js["ms_per_step"], "kernel_ms": js["roofline"]["kernel_ms"],
                "roofline_frac": js["roofline"]["frac"], "clocks": js["clocks"]}
    except Exception as e:      # the headline must not depend on it
        return {"error": repr(e)[:300]}


def droplet_slabs_config(rig, steps=40, warm=20):
    """BASELINE config 4 on N GPUs (the reference is single-GPU; SURVEY.md §8e lists sharded droplets under "next"): the
    same 4 Mi droplets on 8192^2, hmap_dims = map, row slabs of 8192/N rows, droplets owned by the slab that holds them,
    hand-over and peer atomics over NVLink.  Device-timed, max over ranks."""
    from hydro_gen_b200 import Context, _lib, slabs
    N, COUNT, n = 8192, 4 * 1024 * 1024, rig.world
    rows = N // n
    ctx = Context(N, N, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES, device=rig.local, row0=rig.rank * rows, rows=rows)
    slabs.connect_ring(ctx, rig.dist, n, rig.rank)
    m = ctx.get_map(); m.seed = SEED; m.hmap_dims[0], m.hmap_dims[1] = N, N; ctx.set_map(m)
    ctx.gen_heightmap()
    rig.barrier(ctx)
    ctx.run(warm, DT_TIME, DT_TIME, True)
    rig.barrier(ctx)
    ctx.timer_start()
    ctx.run(steps, (warm + 1) * DT_TIME, DT_TIME, True)
    ms = rig.reduce(ctx.timer_stop() / steps)
    owned = rig.reduce(float(ctx.particle_owners().sum()), "sum")
    errs = rig.reduce(ctx.slab_errors(), "sum")
    rig.barrier(ctx)
    ctx.close()
    return {"workload": f"Erosion::dispatch_particle, 4194304 droplets on 8192x8192 over {n} row slabs of {rows} rows, hmap_dims = map",
            "ms_per_dispatch": ms, "Mdroplet_steps_per_s": COUNT / ms / 1e3, "steps": steps, "droplets_with_an_owner": owned, "halo_errors": errs,
            "speedup_note": "one GPU: droplets_8192_4Mi.hmap_dims_8192 of the N=1 line"}


def slab_parity(rig, W=1024, rows=512, steps=48):
    """scripts/mgpu_check.py inside the bench: every rank steps its slab of a W x (rows * N) map with NVLink halo
    pushes (CUDA IPC between the processes) AND the whole map on its own GPU, and compares its rows bit for bit;
    livelier water (d_t 0.01) so far fetches cross slab borders.  Untimed."""
    import numpy as np
    from hydro_gen_b200 import Context, slabs
    n = rig.world

    def setup(ctx):
        m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
        r = ctx.get_rain(); r.period = 8; ctx.set_rain(r)
        e = ctx.get_erosion(); e.d_t = 0.01; ctx.set_erosion(e)
        ctx.gen_heightmap()

    slab = Context(W, rows * n, device=rig.local, row0=rig.rank * rows, rows=rows)
    slabs.connect_ring(slab, rig.dist, n, rig.rank)
    setup(slab)
    rig.barrier(slab)
    slab.run(steps, DT_TIME, DT_TIME, True)
    slab.sync()
    whole = Context(W, rows * n, device=rig.local)
    setup(whole)
    whole.run(steps, DT_TIME, DT_TIME, True)
    ok = True
    for f in (0, 1, 3):
        a, b = slab.download(f), whole.download(f)[rig.rank * rows:(rig.rank + 1) * rows]
        ok = ok and bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    errs = slab.slab_errors()
    far = slab.far_fetch_count()
    rig.barrier(slab)
    slab.close(); whole.close()
    all_ok = rig.reduce(1.0 if ok and errs == 0 else 0.0, "min") == 1.0
    out = {"result": "bit-identical" if all_ok else "MISMATCH", "what": f"{n} slabs of {W}x{rows} vs the whole {W}x{rows * n} map stepped on every rank's own GPU, "
           f"{steps} main-loop steps (rain every 8), H/F/S compared bit for bit on every rank", "far_fetch_cells_all_ranks": rig.reduce(far, "sum")}
    out["droplets"] = droplet_slab_parity(rig)
    return out


def droplet_slab_parity(rig, W=2048, H=2048, count=64, steps=60):
    """The droplet mode on row slabs across processes (CUDA IPC): every rank steps its slab (hand-over of droplets,
    peer atomics on the neighbours' edge texels, image exchanges) AND the whole map on its own GPU; in the sparse regime
    (few droplets, result independent of their order) the droplets it owns, its rows of the heightmap and of the
    momentum map must equal the whole-map run bit for bit.  The map is the same 2048^2 for every N (so the droplets'
    hashed trajectories are too): 64 droplets that never share a texel in these 60 steps -- with colliding droplets the
    order of the additions, which is free in the reference as well, shows in the last bit (scripts/drops_slab_debug.py)."""
    import numpy as np
    from hydro_gen_b200 import Context, _lib, slabs
    n = rig.world
    rows = H // n

    def setup(ctx):
        m = ctx.get_map(); m.seed = SEED; m.hmap_dims[0], m.hmap_dims[1] = W, H; ctx.set_map(m)
        ctx.gen_heightmap()

    slab = Context(W, H, particle_count=count, erosion_type=_lib.HG_PARTICLES, device=rig.local, row0=rig.rank * rows, rows=rows)
    slabs.connect_ring(slab, rig.dist, n, rig.rank)
    setup(slab)
    whole = Context(W, H, particle_count=count, erosion_type=_lib.HG_PARTICLES, device=rig.local)
    setup(whole)
    rig.barrier(slab)
    own0 = slab.particle_owners()
    moved = 0
    for k in range(1, steps + 1):
        slab.dispatch_particle(k * DT_TIME, True)
        whole.dispatch_particle(k * DT_TIME, True)
        if k % 10 == 0:
            own = slab.particle_owners()
            moved += int((own != own0).sum())
            own0 = own
    slab.sync()
    own = slab.particle_owners().astype(bool)
    a, b = slab.download_particles(), whole.download_particles()
    ok = a[own].tobytes() == b[own].tobytes()
    for f in (0, 2):
        x, y = slab.download(f), whole.download(f)[rig.rank * rows:(rig.rank + 1) * rows]
        ok = ok and bool(np.array_equal(x.view(np.uint32), y.view(np.uint32)))
    errs = slab.slab_errors()
    owned_total = rig.reduce(float(own.sum()), "sum")
    rig.barrier(slab)
    slab.close(); whole.close()
    all_ok = rig.reduce(1.0 if ok and errs == 0 else 0.0, "min") == 1.0 and owned_total == count
    return {"result": "bit-identical" if all_ok else "MISMATCH", "what": f"{count} droplets on {n} slabs of {W}x{rows}, {steps} Erosion::dispatch_particle steps, owned droplets and "
            f"the slab's rows of H and M against the whole-map run on the same GPU", "droplets_with_exactly_one_owner": owned_total == count,
            "ownership_changes_seen": rig.reduce(moved, "sum")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (profiling runs)")
    # the other BASELINE configs by hand: e.g. --width 16384 --rows-per-gpu 2048 --gpus 8
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--rows-per-gpu", type=int, default=4096)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    from hydro_gen_b200 import PinnedBuffer, _lib
    rig = Rig(args.gpus)
    rank, n = rig.rank, rig.world
    W, rows = args.width, args.rows_per_gpu
    H = rows * n
    ctx = make_grid(rig, W, rows)
    m = time_grid(rig, ctx, W, rows, n, args.steps, args.warmup)
    ms_per_step, value, k_ms = m["ms_per_step"], m["value"], m["kernel_ms"]
    cells = W * H
    traffic, traffic_src = ncu_traffic(W, rows)
    roofline = {"bound": "hbm", "achieved": m["achieved"], "peak": m["peak"], "unit": "GB/s", "frac": m["frac"],
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_fused_ws", "kernel_ms": k_ms, "peak_source": m["peak_source"],
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
        rig.barrier(ctx)
        ctx.timer_start()
        for _ in range(e2e_steps):
            ctx.step_host_async(ins, outs)
        e_ms = ctx.timer_stop()     # waits for the last download
        rig.barrier(ctx)
        e_ms = rig.reduce(e_ms)
    e2e_val = cells / (e_ms / e2e_steps * 1e-3) / 1e9 if fields else None
    bytes_dir = len(fields) * rows * W * 16
    e2e = {"value": e2e_val, "unit": "Gcell-steps/s", "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
           "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps if fields else None,
           "what": "per step: hg_step_host_async = upload H,F,S from pinned RGBA32F host images, Erosion::dispatch_grid, "
                   "download H,F,S to a second pinned set; consecutive steps pipelined over 3 streams"}
    halo_errors = rig.reduce(ctx.slab_errors(), "sum")
    for p in pins + pouts:
        p.free()
    if fields and not args.no_extras:
        try:
            e2e["copy_ceiling"] = copy_ceiling(rig, bytes_dir)
            e2e["fraction_of_copy_ceiling"] = e2e["copy_ceiling"]["ms_per_step_equivalent"] / e2e["ms_per_step"]
        except Exception as ex:      # the headline must not depend on it
            e2e["copy_ceiling"] = {"error": repr(ex)[:200]}
    rig.barrier(ctx)
    ctx.close()

    line = {"metric": METRIC, "value": value, "unit": "Gcell-steps/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n, W, rows),
            "roofline": roofline, "e2e": e2e, "gpu_launches": m["launches"], "clocks": m["clocks"],
            "far_fetch_cells_per_step": m["far_per_step"], "halo_errors": halo_errors}

    # ---- the other BASELINE configs, on the same line (see the docstring)
    default_shape = (W, rows) == (4096, 4096)
    if not args.no_extras and default_shape:
        xs, xw = 40, 5
        if n == 1:
            c3 = make_grid(rig, 16384, 16384)
            line["north_star_16384"] = extra_entry(16384, 16384, 1, time_grid(rig, c3, 16384, 16384, 1, xs, xw),
                                                   "16384x16384 virtual-pipe + thermal, wet variant, one GPU (BASELINE config 3 at N=1; the north star's >= 60 % roofline target is stated here)")
            c3.close()
            line["droplets_8192_4Mi"] = droplet_config(rig)
        else:
            line["slab_parity"] = slab_parity(rig)
            if 8192 % (8 * n) == 0:
                line["droplets_8192_4Mi_slabs"] = droplet_slabs_config(rig)
            if 16384 % (8 * n) == 0:
                r3 = 16384 // n
                c3 = make_grid(rig, 16384, r3)
                line["cfg3_16384_strong"] = extra_entry(16384, r3, n, time_grid(rig, c3, 16384, r3, n, xs, xw),
                                                        f"16384x16384 virtual-pipe + thermal, wet variant, strong-scaled over {n} row slabs of {r3} rows (BASELINE config 3)")
                rig.barrier(c3)
                c3.close()
            if n == 8:
                c5 = make_grid(rig, 65536, 8192)
                e5 = extra_entry(65536, 8192, 8, time_grid(rig, c5, 65536, 8192, 8, 20, 3),
                                 "65536x65536 virtual-pipe + thermal, wet variant, 8 row slabs of 8192 rows (BASELINE config 5)")
                rig.barrier(c5)
                c5.close()
                # t1 of SURVEY.md §8d: ONE GPU stepping one 65536x8192 slab alone (a whole 65536x8192 map, no peers)
                rig.barrier()
                if rank == 0:
                    c1 = make_grid(rig, 65536, 8192, alone=True)
                    m1 = time_grid(rig, c1, 65536, 8192, 1, 20, 3, collective=False)
                    c1.close()
                    e5["t1_one_gpu_one_slab"] = {"ms_per_step": m1["ms_per_step"], "value": m1["value"], "kernel_ms": m1["kernel_ms"], "roofline_frac": m1["frac"], "clocks": m1["clocks"],
                                                 "what": "one GPU stepping a 65536x8192 map alone: the weak-scaling reference of SURVEY.md §8d"}
                    e5["weak_scaling_inputs"] = {"t1_ms": m1["ms_per_step"], "t8_ms": e5["ms_per_step"]}
                rig.barrier()
                line["cfg5_65536"] = e5

    if rank == 0 and n == 1 and not args.no_extras and default_shape and os.environ.get("HG_FMAD") != "1":
        line["contracted_build"] = contracted_build(args)
    if rank == 0:
        if n == 1 and not args.no_cpu_baseline:
            got = cpu_reference(512, 150)        # ~10-20 s of CPU work
            if got is not None:
                v, threads, sample, _ = got
                line["cpu_baseline"] = {"value": v, "unit": "Gcell-steps/s", "cores": threads, "kind": "reference", "sample": sample}
            v, threads, sample = cpu_port(W, 256, 100)
            port = {"value": v, "unit": "Gcell-steps/s", "cores": threads, "kind": "port", "sample": sample}
            if got is None:
                line["cpu_baseline"] = port
            else:
                line["cpu_port"] = port         # the optimised CPU restatement (oracle/), for scale
        print(json.dumps(line), flush=True)
    rig.close()


if __name__ == "__main__":
    main()
