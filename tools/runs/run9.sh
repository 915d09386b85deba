python -m pytest tests/test_gpu_device_bvh.py tests/test_gpu_trace.py tests/test_gpu_production.py tests/test_gpu_scenes.py tests/test_gpu_device_grid.py -m gpu -x -q > gpurun_out/r2_tests9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests9.log; tail -5 gpurun_out/r2_tests9.log
PTAP_DEVICE_BUILDER=lbvh python -m pytest tests/test_gpu_device_bvh.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B --workload mesh1m --accel lbvh > gpurun_out/r9_ploc_mesh1m.json 2>gpurun_out/r9.err; show gpurun_out/r9_ploc_mesh1m.json
python -c "
import sys; sys.path.insert(0,'.')
import bench
from pathtracerap_b200 import ACCEL_BVH_DEVICE, Renderer
for w in ('mesh1m','mesh5m'):
    s,a = bench.build_scene(w)
    r = Renderer(width=64,height=32,depth=5,accel=ACCEL_BVH_DEVICE); r.allocateOnGPU(s); r.upload(s); print(w, 'device build', r.build_stats()); r.free()
"
PTAP_DEVICE_BUILDER=lbvh $B --workload mesh1m --accel lbvh > gpurun_out/r9_lbvh_mesh1m.json 2>>gpurun_out/r9.err; show gpurun_out/r9_lbvh_mesh1m.json
$B --workload mesh5m --accel lbvh > gpurun_out/r9_ploc_mesh5m.json 2>>gpurun_out/r9.err; show gpurun_out/r9_ploc_mesh5m.json
$B --workload mesh100k --accel lbvh > gpurun_out/r9_ploc_mesh100k.json 2>>gpurun_out/r9.err; show gpurun_out/r9_ploc_mesh100k.json
$B --workload bundled --accel grid > gpurun_out/r9_grid_bundled.json 2>>gpurun_out/r9.err; show gpurun_out/r9_grid_bundled.json
for v in g7 g8; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B --workload bundled --accel grid > gpurun_out/r9_grid_bundled_$v.json 2>>gpurun_out/r9.err; show gpurun_out/r9_grid_bundled_$v.json; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload mesh1m4k --scaling strong > gpurun_out/scale/strong_n1.json 2>>gpurun_out/r9.err; show gpurun_out/scale/strong_n1.json
tail -3 gpurun_out/r9.err
