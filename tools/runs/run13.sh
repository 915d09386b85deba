# round 2, run 15: PTAP_ACCEL_GRID_EMULATED - parity first, then the bundled scene through the walk and through its emulation
mkdir -p gpurun_out
python -m pytest tests/test_gpu_emulated.py tests/test_gpu_production.py tests/test_gpu_trace.py tests/test_gpu_device_grid.py -m gpu -x -q > gpurun_out/r15_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r15_tests.log; tail -25 gpurun_out/r15_tests.log
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r.get('avg_nodes_per_ray'), 'cells', r.get('avg_cells_per_ray'), 'tris', r.get('avg_tris_per_ray'))" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B --workload bundled --accel emu > gpurun_out/r15_emu_bundled.json 2>gpurun_out/r15.err; show gpurun_out/r15_emu_bundled.json
$B --workload bundled --accel grid > gpurun_out/r15_grid_bundled.json 2>>gpurun_out/r15.err; show gpurun_out/r15_grid_bundled.json
$B --workload bundled --accel bvh > gpurun_out/r15_bvh_bundled.json 2>>gpurun_out/r15.err; show gpurun_out/r15_bvh_bundled.json
$B --workload cornell --accel emu > gpurun_out/r15_emu_cornell.json 2>>gpurun_out/r15.err; show gpurun_out/r15_emu_cornell.json
tail -5 gpurun_out/r15.err
