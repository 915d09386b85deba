# launch list of the emulated walk on the bundled scene (per-launch durations; serialised, cold-cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r14_emu_launches.csv python bench.py --workload bundled --accel emu --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r14_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r14_emu_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1000.0 if u in ('nsecond', 'ns') else v if u in ('usecond', 'us') else v * 1000.0 if u in ('msecond', 'ms') else v
    name = r[ki].split('(')[0][:40]
    agg[name][0] += 1; agg[name][1] += us; seq.append((name, us))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f'{k:42s} n={n:4d} total={t/1000:9.3f} ms avg={t/n:9.1f} us')
print('first 40:'); [print(f'  {n:40s} {u:9.1f}') for n, u in seq[:40]]
PY
