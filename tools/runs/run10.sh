mkdir -p gpurun_out/scale
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for v in p4 p16 p32; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B --workload mesh1m --accel lbvh > gpurun_out/r10_$v.json 2>>gpurun_out/r10.err; show gpurun_out/r10_$v.json
PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so python -c "
import sys; sys.path.insert(0,'.')
import bench
from pathtracerap_b200 import ACCEL_BVH_DEVICE, Renderer
s,a = bench.build_scene('mesh1m')
r = Renderer(width=64,height=32,depth=5,accel=ACCEL_BVH_DEVICE); r.allocateOnGPU(s); r.upload(s); print('$v device build', r.build_stats()); r.free()
"; done
$B --workload bundled --accel grid > gpurun_out/r10_grid_bundled.json 2>>gpurun_out/r10.err; show gpurun_out/r10_grid_bundled.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload mesh1m4k --scaling strong > gpurun_out/scale/strong_n1.json 2>>gpurun_out/r10.err; show gpurun_out/scale/strong_n1.json
tail -3 gpurun_out/r10.err
