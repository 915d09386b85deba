show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for w in mesh1m bundled mesh5m; do
$B --workload $w > gpurun_out/r39_${w}_default.json 2>>gpurun_out/r39.err; show gpurun_out/r39_${w}_default.json
PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_kx.so $B --workload $w > gpurun_out/r39_${w}_kx.json 2>>gpurun_out/r39.err; show gpurun_out/r39_${w}_kx.json
done
PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_kx.so python -m pytest tests/test_gpu_trace.py tests/test_gpu_scenes.py -m gpu -x -q 2>&1 | tail -2
tail -2 gpurun_out/r39.err
