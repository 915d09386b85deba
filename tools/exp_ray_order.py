#!/usr/bin/env python
"""Experiment (GPU box): does the ORDER of an incoherent wavefront matter to k_trace_bvh?  Builds bounce-1 rays of the bench scene
(camera hits + cosine-distributed directions), then times ptap_bench_trace on the same rays in several orders."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pathtracerap_b200 import ACCEL_BVH, Renderer

workload = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
scene, arrays = bench.build_scene(workload)
scene.build_bvh()
r = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH)
r.allocateOnGPU(scene)
W, H = 1920, 1080
i = np.arange(W * H)
cam = np.zeros((W * H, 6), np.float32)
cam[:, 2] = 920.0
cam[:, 3] = (-10.0 + ((i % W).astype(np.float32) * np.float32(20.0 / W)).astype(np.float64)).astype(np.float32)
cam[:, 4] = (-4.0 + ((i // W).astype(np.float32) * np.float32(16.0 / H)).astype(np.float64)).astype(np.float32)
cam[:, 5] = -20.0
h = r.trace(cam)
ok = h["model"] >= 0
d = cam[ok, 3:] / np.linalg.norm(cam[ok, 3:], axis=1, keepdims=True)
p = cam[ok, :3] + d * h["dist"][ok, None]
n = h["normal"][ok]
rs = np.random.RandomState(1)
u1, u2 = rs.rand(len(p)), rs.rand(len(p))
up, over, ar = np.sqrt(u1), np.sqrt(1 - u1), 2 * np.pi * u2
a = np.where(np.abs(n[:, :1]) < 0.577, [[1, 0, 0]], np.where(np.abs(n[:, 1:2]) < 0.577, [[0, 1, 0]], [[0, 0, 1]]))
t1 = np.cross(n, a); t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
t2 = np.cross(n, t1)
nd = n * up[:, None] + t1 * (np.cos(ar) * over)[:, None] + t2 * (np.sin(ar) * over)[:, None]
rays = np.concatenate([p + 0.1 * n, nd], 1).astype(np.float32)
print("bounce-1 rays:", len(rays))

def timeit(name, order):
    ms = min(r.bench_trace(rays[order], reps=10) for _ in range(2))
    print(f"{name:28s} {ms:8.3f} ms  {len(rays) / ms / 1e3:8.1f} Mrays/s", flush=True)

idx = np.arange(len(rays))
timeit("camera rays (coherent)", idx[:0] if False else idx)      # placeholder to warm up
ms = min(r.bench_trace(cam, reps=10) for _ in range(2)); print(f"{'primary rays':28s} {ms:8.3f} ms  {len(cam) / ms / 1e3:8.1f} Mrays/s")
timeit("slot order", idx)
octant = (rays[:, 3] < 0) * 1 + (rays[:, 4] < 0) * 2 + (rays[:, 5] < 0) * 4
timeit("octant buckets (stable)", np.argsort(octant, kind="stable"))
lo, hi = rays[:, :3].min(0), rays[:, :3].max(0)
q = np.clip(((rays[:, :3] - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023)
def spread(v):
    v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
    return v
morton = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
timeit("morton(origin)", np.argsort(morton, kind="stable"))
timeit("octant, then morton", np.lexsort((morton, octant)))
timeit("morton>>15, then octant", np.lexsort((octant, morton >> 15)))
dq = np.clip(((rays[:, 3:] * 0.5 + 0.5) * 15).astype(np.int64), 0, 15)
dkey = dq[:, 0] | (dq[:, 1] << 4) | (dq[:, 2] << 8)
timeit("direction cell, then morton", np.lexsort((morton, dkey)))
timeit("random permutation", rs.permutation(len(rays)))
