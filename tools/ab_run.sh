# A/B measurements on one GPU box: parity subset first, then bench lines of the default library and of the variant builds
python -m pytest tests/test_gpu_trace.py tests/test_gpu_scenes.py tests/test_gpu_production.py tests/test_gpu_device_bvh.py tests/test_gpu_wavefront.py -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log; tail -4 gpurun_out/r2_tests5.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'], 'trace1lane', r['trace_Mrays_per_s'])" 2>&1 | tail -1; }
$B --workload mesh1m > gpurun_out/ab_h_mesh1m.json 2>gpurun_out/ab.err; show gpurun_out/ab_h_mesh1m.json
for v in pathtracerap_b200/variants/*.so; do n=$(basename $v .so); PTAP_LIB=$PWD/$v $B --workload mesh1m > gpurun_out/ab_${n}_mesh1m.json 2>>gpurun_out/ab.err; show gpurun_out/ab_${n}_mesh1m.json; done
$B --workload mesh5m > gpurun_out/ab_h_mesh5m.json 2>>gpurun_out/ab.err; show gpurun_out/ab_h_mesh5m.json
$B --workload bundled > gpurun_out/ab_h_bundled.json 2>>gpurun_out/ab.err; show gpurun_out/ab_h_bundled.json
$B --workload mesh100k > gpurun_out/ab_h_mesh100k.json 2>>gpurun_out/ab.err; show gpurun_out/ab_h_mesh100k.json
$B --workload mesh1m --accel lbvh > gpurun_out/ab_h_mesh1m_lbvh.json 2>>gpurun_out/ab.err; show gpurun_out/ab_h_mesh1m_lbvh.json
tail -3 gpurun_out/ab.err
