show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload bundled --accel emu"
for l in 4 6 8; do PTAP_LANES=$l $B > gpurun_out/r40_lanes$l.json 2>>gpurun_out/r40.err; show gpurun_out/r40_lanes$l.json; done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload mesh1m --accel emu --spp 16"
for l in 4 8; do PTAP_LANES=$l $B > gpurun_out/r40_mesh1m_lanes$l.json 2>>gpurun_out/r40.err; show gpurun_out/r40_mesh1m_lanes$l.json; done
