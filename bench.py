#!/usr/bin/env python
"""bench.py - headline benchmark of the render hot path (BASELINE.json: Mrays/s, all bounces).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mesh100k|mesh1m|bundled|cornell] [--impl ptap|reference]

A "step" is one whole frame of the workload: Renderer::renderLoop for `spp` iterations (ray generation, closest hit,
shading + compaction, film accumulation for every bounce).  Rays = rays actually traced, summed over closest-hit launches
(a restored first-hit cache does not count).  N > 1: launched under torchrun, one rank per GPU, the scene
replicated, every rank renders its own sample range [rank*spp, (rank+1)*spp) of the same frame, and the per-rank films are
combined with ONE NCCL reduce inside the timed region; per-GPU work is fixed => "scaling": "weak".

Rank 0 prints ONE JSON line (contract in the task statement).  `--impl reference` times the reference's own CPU code
(oracle/_ref, or the C port when that library is absent) on a bounded sample of the same workload.
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

GOLDEN_SCENE = os.path.join(ROOT, "tests", "golden", "bundled_scene.npz")

# name -> (W, H, spp, depth, description)
WORKLOADS = {
    "mesh100k": (1920, 1080, 64, 5, "configs[1]: displaced icosphere 81,920 tris (DIFFUSE) in the bundled box with its 4 lights, 1920x1080, 64 spp, depth 5, BVH"),
    "mesh1m": (1920, 1080, 64, 5, "configs[3]: displaced icosphere 1,310,720 tris (DIFFUSE) in the bundled box, 1920x1080, 64 spp, depth 5, BVH"),
    "mesh5m": (1920, 1080, 64, 5, "configs[3] upper end: displaced icosphere 5,242,880 tris (DIFFUSE) in the bundled box, 1920x1080, 64 spp, depth 5"),
    "bundled": (2800, 2240, 64, 5, "configs[2]: the reference's coded scene (METAL/COAT/REFLECTIVE/DIFFUSE/EMISSIVE, 11 models), 2800x2240, 64 spp, depth 5, BVH"),
    "mesh1m4k": (3840, 2160, 128, 5, "configs[4]: the mesh1m scene at 3840x2160, 128 spp per GPU (1024 spp on 8 GPUs), sample-partitioned, one NCCL reduce of the film"),
    "cornell": (512, 512, 16, 8, "configs[0]: Cornell box from Input data, 512x512, 16 spp, depth 8, diffuse only"),
}


# ------------------------------------------------------------------------------------------------ scenes (host arrays)

def bundled_arrays():
    z = np.load(GOLDEN_SCENE)
    return {k: z[k] for k in ("models", "meshes", "vertices", "triangles")}


def build_scene(workload: str):
    """Returns (product Scene, dict of the four reference-layout arrays) for a workload."""
    from pathtracerap_b200 import DIFFUSE, EMISSIVE, Scene
    from pathtracerap_b200 import _native as N
    base = bundled_arrays()
    if workload == "bundled":
        s = Scene.from_arrays(base["models"], base["meshes"], base["vertices"], base["triangles"])
    elif workload == "cornell":
        m = base["models"][3:].copy()
        m["mat"]["type"] = np.where(m["mat"]["type"] == EMISSIVE, EMISSIVE, DIFFUSE)
        s = Scene.from_arrays(m, base["meshes"], base["vertices"], base["triangles"])
    else:
        level = {"mesh100k": 6, "mesh1m": 8, "mesh1m4k": 8, "mesh5m": 9}[workload]
        keep = [3, 7, 8, 9, 10]                      # box + four lights of Scene.cpp:114-124, 175-221
        s = Scene.from_arrays(base["models"][keep], base["meshes"], base["vertices"], base["triangles"])
        mi = s.add_icosphere(level, radius=1000.0, displacement=0.05, seed=1 if workload == "mesh100k" else 2)
        s.add_model(mi, translate=(25.0, 230.0, -50.0), rotate_y_degrees=30.0, scale=(0.25, 0.25, 0.25), material=DIFFUSE, color=(0.75, 0.6, 0.4))
    a = s.arrays()
    return s, {k: a[k] for k in ("models", "meshes", "vertices", "triangles")}


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """Polls nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arms (oracle/)

def cpu_reference_arm(arrays, depth, sample_w, sample_h, iters, threads=None):
    """The reference's own CPU implementation on a bounded sample: `iters` iterations at sample_w x sample_h of the same
    scene, full bounce loop, 25^3 grid as the reference builds it.  Returns (Mrays/s, rays, seconds, kind, cores, ms_per_iter)."""
    from oracle import ref
    # every host core this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently run the
    # reference on one thread and flatter the GPU arm
    threads = threads or len(os.sched_getaffinity(0))
    if ref.available():
        kind = "reference"
        if threads:
            ref.lib().ref_set_threads(threads)
        cores = ref.lib().ref_max_threads()
        scene = ref.RefScene.from_arrays(arrays["models"], arrays["meshes"], arrays["vertices"], arrays["triangles"])
        r = ref.RefRenderer(scene, sample_w, sample_h, depth)
        r.init_image()
        t0 = time.perf_counter()
        rays = sum(sum(r.run_iteration(it)) for it in range(iters))
        dt = time.perf_counter() - t0
        r.close(); scene.close()
    else:
        from oracle import port
        kind = "port"
        if threads:
            port.lib().oracle_set_threads(threads)
        cores = port.lib().oracle_max_threads()
        scene = port.OracleScene(arrays)
        w = port.OracleWavefront(scene, sample_w, sample_h, depth)
        w.init_image()
        t0 = time.perf_counter()
        rays = w.render(0, iters, first_hit_cache=False)
        dt = time.perf_counter() - t0
        w.close()
    return rays / dt / 1e6, rays, dt, kind, cores, dt / iters * 1e3


def sample_size(W, H, target_paths):
    """Largest W/k x H/k (k integer) with at most target_paths paths and a multiple of 32 (Renderer.cpp:573)."""
    k = 1
    while (W // k) * (H // k) > target_paths or ((W // k) * (H // k)) % 32:
        k += 1
    return W // k, H // k


# ------------------------------------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ptap", choices=["ptap", "reference"])
    ap.add_argument("--workload", default="mesh1m", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel of the workload")
    ap.add_argument("--accel", default="bvh", choices=["bvh", "lbvh", "grid"], help="bvh: host-built SAH tree (default); lbvh: tree built on the GPU at upload; grid: the reference's uniform grid")
    ap.add_argument("--grid-dim", type=int, default=25, help="voxels per axis of the uniform grid (--accel grid); the reference fixes 25 (Config.h:8-10)")
    ap.add_argument("--no-cache", action="store_true", help="disable the first-hit cache (Renderer.cpp:594-613)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    W, H, spp, depth, desc = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    metric, unit = "Mrays/s (all bounces)", "Mrays/s"

    # -------- reference arm: the reference's CPU implementation on a bounded sample; rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        _, arrays = build_scene(args.workload) if os.path.exists(os.path.join(ROOT, "pathtracerap_b200", "libptap.so")) else (None, bundled_arrays())
        sw, sh = sample_size(W, H, 140_000)
        ITERS = 8
        for _ in range(max(args.warmup, 1)):
            cpu_reference_arm(arrays, depth, sw, sh, 1)
        vals, rays_total, t_total = [], 0, 0.0
        for _ in range(args.steps):
            v, rays, dt, kind, cores, _ms = cpu_reference_arm(arrays, depth, sw, sh, ITERS)
            vals.append(v); rays_total += rays; t_total += dt
        value = rays_total / t_total / 1e6
        sample = f"{sw}x{sh} x {ITERS} iterations per step of the {args.workload} scene, full bounce loop, reference 25^3 grid, no first-hit cache"
        print(json.dumps({"impl": "reference", "metric": metric, "value": round(value, 4), "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": round(t_total / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": args.workload, "description": desc, "resolution": [W, H], "spp": spp, "depth": depth},
                          "cpu_baseline": {"value": round(value, 4), "unit": unit, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": round(value, 4), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # -------- our arm
    import torch
    import torch.distributed as dist
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE, ACCEL_GRID_COMPAT, Renderer

    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)
    torch.cuda.init()

    scene, arrays = build_scene(args.workload)
    accel = {"bvh": ACCEL_BVH, "lbvh": ACCEL_BVH_DEVICE, "grid": ACCEL_GRID_COMPAT}[args.accel]
    t0 = time.perf_counter()
    if accel == ACCEL_BVH:
        scene.build_bvh()                 # host-side, part of scene construction like the reference's addMeshesToGrid
    elif accel == ACCEL_BVH_DEVICE:
        pass                              # built on the GPU inside Renderer.allocateOnGPU / upload
    else:
        scene.build_grids(args.grid_dim, args.grid_dim, args.grid_dim)
    build_s = time.perf_counter() - t0
    ntris = len(arrays["triangles"])

    r = Renderer(device=dev, width=W, height=H, depth=depth, accel=accel, first_hit_cache=not args.no_cache, profile=True)
    r.allocateOnGPU(scene)

    from pathtracerap_b200 import multi_gpu
    # sample partition, weak scaling: every rank renders `spp` iterations, the union over ranks is one (world * spp)-sample frame
    it0, it1 = multi_gpu.iteration_range(rank, world, world * spp)
    lib_stream = torch.cuda.ExternalStream(r.stream_ptr(), device=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            dist.barrier()
        r.sync(); torch.cuda.synchronize()

    def step():
        r.frame_begin()
        r.render(it0, it1)
        if world > 1:
            # the collective is ordered on the LIBRARY's stream (NCCL's stream waits for it and it waits for NCCL), so the device
            # timer on that stream covers render + reduce and no host synchronisation separates them
            with torch.cuda.stream(lib_stream):
                multi_gpu.reduce_film(film_t, 0)

    # counting pass (not timed): traversal work per ray of this exact workload, for the algorithmic-bytes figure
    r.set_params(W, H, depth, first_hit_cache=not args.no_cache, count=True)
    r.render(it0, it0 + 1); r.sync()
    cst = r.stats()
    avg_nodes, avg_tris, avg_cells, avg_refs = cst["avg_nodes"], cst["avg_tris"], cst["avg_cells"], cst["avg_refs"]
    r.set_params(W, H, depth, first_hit_cache=not args.no_cache, profile=False)
    film_t = multi_gpu.film_tensor(r) if world > 1 else None      # zero-copy torch view of the library's film buffer

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    rays = launches = 0
    lanes = r.stats()["lanes"]
    t_wall0 = time.perf_counter()
    ms_dev = 0.0
    for _ in range(args.steps):
        r.timer_start()
        step()
        ms_dev += r.timer_stop()                 # device time on the launching stream (render + reduce; the lanes' streams join it)
        st = r.stats()
        rays += st["rays_traced"]; launches += st["kernel_launches"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    t_rank = ms_dev / 1e3                        # CUDA events on the library stream; max over ranks below
    if world > 1:
        tt = torch.tensor([t_rank], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
        rr = torch.tensor([float(rays)], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        rays_all = float(rr.item())
    else:
        t_max, rays_all = t_rank, float(rays)
    value = rays_all / t_max / 1e6

    # -------- end to end through the public API with HOST buffers: upload scene (H2D) + render + film read-back (D2H), every step
    film_host = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True).numpy()      # page-locked read-back target
    from pathtracerap_b200 import _native as N
    import ctypes as C
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        r.upload(scene)
        r.frame_begin()
        r.render(it0, it1)
        if world > 1:
            with torch.cuda.stream(lib_stream):
                multi_gpu.reduce_film(film_t, 0)
        N.lib().ptap_read_film(r.h, N.ptr(film_host))
        e2e_rays += r.stats()["rays_traced"]
    barrier()
    t_e2e = time.perf_counter() - t0
    h2d = r.stats()["scene_bytes"]
    d2h = film_host.nbytes
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=f"cuda:{dev}"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t_e2e = float(tt.item())
        rr = torch.tensor([float(e2e_rays)], dtype=torch.float64, device=f"cuda:{dev}"); dist.all_reduce(rr, op=dist.ReduceOp.SUM); e2e_rays = float(rr.item())
    e2e_value = e2e_rays / t_e2e / 1e6

    # -------- instrumented pass (not part of `value`): the same K steps on ONE lane with CUDA events around every launch, for the
    # per-kernel durations of the roofline.  (The timed region above pipelines iterations over PTAP_LANES streams; kernels of
    # different lanes overlap there, so a per-kernel duration is only defined when they run one after the other.)
    r.set_params(W, H, depth, first_hit_cache=not args.no_cache, profile=True)
    film_t = multi_gpu.film_tensor(r) if world > 1 else None
    step(); r.sync()
    trace_launches = 0
    rays_i = 0
    ms_trace = ms_shade = ms_gen = ms_instr = 0.0
    for _ in range(args.steps):
        r.timer_start()
        step()
        ms_instr += r.timer_stop()
        st = r.stats()
        rays_i += st["rays_traced"]; trace_launches += st["trace_launches"]
        ms_trace += st["ms_trace"]; ms_shade += st["ms_shade"]; ms_gen += st["ms_generate"]
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -------- roofline of the dominant kernel (closest hit), measured live in the instrumented pass
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if accel != ACCEL_GRID_COMPAT:
        bytes_per_ray = 48.0 + 128.0 * avg_nodes + 64.0 * avg_tris
    else:
        bytes_per_ray = 48.0 + 8.0 * avg_cells + 4.0 * avg_refs + 48.0 * avg_tris
    achieved = bytes_per_ray * rays_i / (ms_trace / 1e3) / 1e9 if ms_trace > 0 else None
    # measured DRAM bytes of one launch of the same kernel on the same workload, from the committed `ncu --set full` capture
    traffic, traffic_note = None, "no ncu capture committed for this workload"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(f"{args.workload}:{args.accel}")
        if t:
            traffic, traffic_note = t["dram_bytes_per_launch"], t["note"]
    roofline = {"bound": "hbm", "kernel": "k_trace_bvh" if accel != ACCEL_GRID_COMPAT else "k_trace_grid",
                "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "bytes_per_ray": round(bytes_per_ray, 1), "avg_nodes_per_ray": round(avg_nodes, 2), "avg_tris_per_ray": round(avg_tris, 2),
                "avg_cells_per_ray": round(avg_cells, 2), "avg_refs_per_ray": round(avg_refs, 2),
                "trace_launches": trace_launches, "avg_launch_ms": round(ms_trace / max(trace_launches, 1), 4),
                "trace_Mrays_per_s": round(rays_i / (ms_trace / 1e3) / 1e6, 1) if ms_trace > 0 else None,
                "share_of_step": {"trace": round(ms_trace / ms_instr, 3), "shade": round(ms_shade / ms_instr, 3), "generate": round(ms_gen / ms_instr, 3)},
                "timing": f"instrumented pass of the same {args.steps} steps on one lane, CUDA events around every launch ({ms_instr / args.steps:.3f} ms per step); "
                          "the timed region of `value` pipelines iterations over the library's lanes without per-kernel events"}

    cpu = None
    if not args.no_cpu_baseline:
        sw, sh = sample_size(W, H, 140_000)
        cpu_reference_arm(arrays, depth, sw, sh, 1)          # warm the page cache / OpenMP pool
        n_it = 2
        v, rays_c, dt, kind, cores, _ms = cpu_reference_arm(arrays, depth, sw, sh, n_it)
        if dt < 5.0:
            n_it = int(min(64, max(2, 12.0 / (dt / n_it))))
            v, rays_c, dt, kind, cores, _ms = cpu_reference_arm(arrays, depth, sw, sh, n_it)
        cpu = {"value": round(v, 4), "unit": unit, "cores": cores, "kind": kind,
               "sample": f"{sw}x{sh} x {n_it} iterations of the same scene ({rays_c} rays, {dt:.1f} s), full bounce loop, reference 25^3 grid, no first-hit cache"}

    out = {"metric": metric, "value": round(value, 2), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(t_max / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "description": desc, "resolution": [W, H], "spp_per_gpu": spp, "depth": depth,
                      "triangles": ntris, "models": len(arrays["models"]), "accel": args.accel if args.accel == "bvh" else f"grid {args.grid_dim}^3", "first_hit_cache": not args.no_cache, "lanes": lanes,
                      "parallelism": f"sample-partition x{world}" + (" + 1 NCCL reduce of the film" if world > 1 else ""),
                      "l2": "wavefront state (6 float4 queues + hits, >300 MB at 1080p) exceeds L2 every bounce; the scene is small and L2-resident by design",
                      "host_accel_build_s": round(build_s, 3), "device_bvh_build_ms": round(r.stats()["ms_build"], 3) if accel == ACCEL_BVH_DEVICE else None},
           "rays_per_step": int(rays_all / args.steps), "ms_per_frame_device": round(ms_dev / args.steps, 3), "wall_s_timed_region": round(t_wall, 3),
           "e2e": {"value": round(e2e_value, 2), "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": round(t_e2e / args.steps * 1e3, 3), "what": "Renderer.upload(scene) from host arrays + renderLoop + film read-back to host, per step"},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
