python -m pytest tests/test_gpu_scenes.py -m gpu -x -q 2>&1 | tail -15
PTAP_UPLOAD_TIMING=1 python - <<'PY' 2>&1 | grep -E "BVH checks|upload" | tail -6
import os, sys, time
sys.path.insert(0, ".")
import bench
from pathtracerap_b200 import ACCEL_BVH, Renderer
scene, arrays = bench.build_scene("mesh1m")
scene.build_bvh()
r = Renderer(width=1920, height=1080, depth=5, accel=ACCEL_BVH)
r.allocateOnGPU(scene)
for k in range(3):
    t0 = time.perf_counter(); r.upload(scene); r.sync(); t1 = time.perf_counter()
    print(f"upload {k}: {(t1 - t0) * 1e3:.2f} ms wall", file=sys.stderr)
PY
