# correctly-rounded reciprocal instead of IEEE division by 1: full parity, then the headline workloads
python -m pytest tests -m gpu -x -q > gpurun_out/r38_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r38_tests.log; tail -4 gpurun_out/r38_tests.log
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
$B --workload mesh1m > gpurun_out/r38_mesh1m.json 2>gpurun_out/r38.err; show gpurun_out/r38_mesh1m.json
$B --workload bundled > gpurun_out/r38_bundled.json 2>>gpurun_out/r38.err; show gpurun_out/r38_bundled.json
$B --workload bundled --accel emu > gpurun_out/r38_bundled_emu.json 2>>gpurun_out/r38.err; show gpurun_out/r38_bundled_emu.json
$B --workload bundled --accel grid --steps 2 > gpurun_out/r38_bundled_grid.json 2>>gpurun_out/r38.err; show gpurun_out/r38_bundled_grid.json
tail -2 gpurun_out/r38.err
