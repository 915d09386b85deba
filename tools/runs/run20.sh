# persistent replay kernel: parity, then the bundled scene
python -m pytest tests/test_gpu_emulated.py tests/test_gpu_production.py -m gpu -x -q 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload bundled --accel emu"
$B > gpurun_out/r20_base.json 2>gpurun_out/r20.err; show gpurun_out/r20_base.json
PTAP_EMU_REPLAY_CTAS=4 $B > gpurun_out/r20_replay4.json 2>>gpurun_out/r20.err; show gpurun_out/r20_replay4.json
PTAP_EMU_REPLAY_CTAS=8 $B > gpurun_out/r20_replay8.json 2>>gpurun_out/r20.err; show gpurun_out/r20_replay8.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload cornell --accel emu > gpurun_out/r20_cornell.json 2>>gpurun_out/r20.err; show gpurun_out/r20_cornell.json
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:'k_trace_emu|k_emu_replay|k_trace_grid' -c 12 --csv --log-file gpurun_out/r20_launches.csv $B --spp 2 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r20_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); idi=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[idi], r[ki].split('(')[0][:24]), {})[r[mi]] = r[vi]
for (i,k),m in cur.items(): print(i,k,m)
PY
tail -3 gpurun_out/r20.err
