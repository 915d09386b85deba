# full GPU suite, smoke, default bench line (what the driver runs at round end)
python -m pytest tests -m gpu -x -q > gpurun_out/r47_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r47_tests.log; tail -6 gpurun_out/r47_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r47_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r47_smoke.log
python bench.py > gpurun_out/r47_bench.json 2> gpurun_out/r47_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r47_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], json.dumps(d.get('extras',{}).get('ms_per_frame_2800x2240_64spp')))"
tail -3 gpurun_out/r47_bench.err
