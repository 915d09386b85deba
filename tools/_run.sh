python -m pytest tests/test_gpu_wavefront.py tests/test_gpu_scenes.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/sort_tests.log
PTAP_SHADE_SORT=0 python -m pytest tests/test_gpu_wavefront.py -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/sort_tests.log
for w in bundled mesh1m cornell; do
  PTAP_SHADE_SORT=1 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sort1_$w.json 2> gpurun_out/sort1_$w.err
  PTAP_SHADE_SORT=0 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sort0_$w.json 2> gpurun_out/sort0_$w.err
done
