# dense-mesh parity of the emulation, then BASELINE configs[3]'s "BVH vs grid" on the 1.3 M-triangle mesh with the grid emulated
python -m pytest tests/test_gpu_emulated.py -m gpu -x -q -k dense 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'nodes', r.get('avg_nodes_per_ray'), 'cells', r.get('avg_cells_per_ray'), 'tris', r.get('avg_tris_per_ray'))" 2>&1 | tail -1; }
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload mesh1m"
$B --accel emu --grid-dim 25 --spp 8 > gpurun_out/r25_mesh1m_emu25.json 2>gpurun_out/r25.err; show gpurun_out/r25_mesh1m_emu25.json
$B --accel emu --grid-dim 160 --spp 8 > gpurun_out/r25_mesh1m_emu160.json 2>>gpurun_out/r25.err; show gpurun_out/r25_mesh1m_emu160.json
$B --accel grid --grid-dim 160 --spp 4 > gpurun_out/r25_mesh1m_grid160.json 2>>gpurun_out/r25.err; show gpurun_out/r25_mesh1m_grid160.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload mesh100k --accel emu --spp 16 > gpurun_out/r25_mesh100k_emu25.json 2>>gpurun_out/r25.err; show gpurun_out/r25_mesh100k_emu25.json
tail -3 gpurun_out/r25.err
