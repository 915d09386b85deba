#!/usr/bin/env python
"""Kernel-level A/B of the closest-hit kernel: the bounce-1 rays of a real 1920x1080 iteration of a workload (taken through
ptap_render_probe) traced `reps` times back to back by ptap_bench_trace (CUDA events, inputs resident), plus a short whole-frame rate.
The library under test is chosen with PTAP_LIB (tools/build_variants.sh); one process per variant.

    PTAP_LIB=pathtracerap_b200/variants/libptap_s0.so python tools/trace_micro.py [workload] [tag]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pathtracerap_b200 import ACCEL_BVH, Renderer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
tag = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(os.environ.get("PTAP_LIB", "default"))
W, H, spp, depth, _ = bench.WORKLOADS[workload]
scene, arrays = bench.build_scene(workload)
scene.build_bvh()
r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
r.allocateOnGPU(scene)
out = {"tag": tag, "workload": workload}
for rnd in (0, 1, 2):
    rays, _, _ = r.render_probe(0, rnd)
    ms = r.bench_trace(rays, reps=30)
    out[f"round{rnd}"] = {"rays": len(rays), "ms": round(ms, 4), "Mrays_s": round(len(rays) / ms / 1e3, 1)}
r.frame_begin()
r.render(0, 8); r.sync()
r.timer_start(); r.frame_begin(); r.render(0, 16); ms = r.timer_stop()
out["frame16spp"] = {"ms": round(ms, 3), "Mrays_s": round(r.stats()["rays_traced"] / ms / 1e3, 1)}
print(json.dumps(out))
