python -m pytest tests/test_gpu_emulated.py tests/test_gpu_production.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|emulated walk|Error|assert" | head -12
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'cells', r.get('avg_cells_per_ray'))" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B --workload bundled --accel emu > gpurun_out/r28_bundled.json 2>gpurun_out/r28.err; show gpurun_out/r28_bundled.json
$B --workload mesh1m --accel emu --grid-dim 25 --spp 8 --steps 1 > gpurun_out/r28_mesh1m_emu25.json 2>>gpurun_out/r28.err; show gpurun_out/r28_mesh1m_emu25.json
$B --workload mesh100k --accel emu --spp 16 --steps 1 > gpurun_out/r28_mesh100k_emu25.json 2>>gpurun_out/r28.err; show gpurun_out/r28_mesh100k_emu25.json
for w in bundled mesh1m; do
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:'k_trace_emu|k_emu_|k_trace_grid' -s 5 -c 5 --csv --log-file gpurun_out/r28_launches_$w.csv $B --workload $w --accel emu --spp 2 --steps 1 > /dev/null 2>&1
python - $w <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(f'gpurun_out/r28_launches_{sys.argv[1]}.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); idi=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[idi], r[ki].split('(')[0][:24]), {})[r[mi]] = r[vi]
for (i,k),m in cur.items(): print(sys.argv[1], i,k,m)
PY
done
tail -3 gpurun_out/r28.err
