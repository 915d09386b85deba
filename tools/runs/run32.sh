python - <<'PY' 2>&1 | tail -30
import os, sys, time
os.environ["PTAP_UPLOAD_TIMING"] = "1"
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
r.render(0, 2); r.sync()
t0 = time.perf_counter(); f = r.film(); t1 = time.perf_counter()
print(f"film read-back: {(t1 - t0) * 1e3:.2f} ms wall", file=sys.stderr)
PY
