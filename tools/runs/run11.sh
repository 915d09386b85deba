# round 2, run 11: k_trace_grid with model loop / empty-space steps / mailbox: parity first, then A/B of the three devices
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trace.py tests/test_gpu_device_grid.py tests/test_gpu_production.py tests/test_gpu_wavefront.py tests/test_gpu_scenes.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r11_tests.log; tail -5 gpurun_out/r11_tests.log
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r.get('avg_cells_per_ray'), r.get('avg_refs_per_ray'), r.get('avg_tris_per_ray'))" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B --workload bundled --accel grid > gpurun_out/r11_grid_bundled.json 2>gpurun_out/r11.err; show gpurun_out/r11_grid_bundled.json
for v in goff gmb0 gbl0 gml0 gmb2 gc7; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B --workload bundled --accel grid > gpurun_out/r11_grid_bundled_$v.json 2>>gpurun_out/r11.err; show gpurun_out/r11_grid_bundled_$v.json; done
$B --workload cornell --accel grid > gpurun_out/r11_grid_cornell.json 2>>gpurun_out/r11.err; show gpurun_out/r11_grid_cornell.json
tail -3 gpurun_out/r11.err
