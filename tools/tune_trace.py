#!/usr/bin/env python
"""Sweeps the work-stealing knobs of k_trace_bvh on a GPU box (PTAP_REFILL / PTAP_BATCH / PTAP_TRACE_CTAS are read by
ptap_create) by running bench.py once per setting.  Usage: python tools/tune_trace.py [workload] > gpurun_out/tune.txt"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
grid = {"PTAP_VOTE_TRI": ["4", "8", "12", "16"], "PTAP_VOTE_INST": ["4", "8", "16"], "PTAP_VOTE_REFILL": ["8"], "PTAP_BATCH": ["32"]}
if len(sys.argv) > 2:
    grid = json.loads(sys.argv[2])
keys = sorted(grid)
for combo in itertools.product(*(grid[k] for k in keys)):
    env = dict(os.environ, **dict(zip(keys, combo)))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-extras"],
                       env=env, capture_output=True, text=True)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    try:
        j = json.loads(line)
        print(dict(zip(keys, combo)), "value", j["value"], "ms_per_step", j["ms_per_step"], "trace_Mrays/s", j["roofline"].get("trace_Mrays_per_s"), flush=True)
    except Exception:
        print(dict(zip(keys, combo)), "FAILED", p.stderr[-400:], flush=True)
