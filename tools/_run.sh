python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/full_tests.log
export PATH=/usr/local/cuda/bin:$PATH
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wavefront.py -m gpu -x -q -k "shade_and or supersampled or bmp_matches or iteration_ranges" > gpurun_out/san_mem.log 2>&1; echo "rc=$?" >> gpurun_out/san_mem.log
PTAP_SHADE_SORT=1 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wavefront.py -m gpu -x -q -k "shade_and or iteration_ranges" > gpurun_out/san_mem_sort.log 2>&1; echo "rc=$?" >> gpurun_out/san_mem_sort.log
PTAP_SHADE_SORT=1 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_wavefront.py -m gpu -x -q -k "shade_and" > gpurun_out/san_race.log 2>&1; echo "rc=$?" >> gpurun_out/san_race.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_trace.py -m gpu -x -q -k "edge or golden" > gpurun_out/san_trace.log 2>&1; echo "rc=$?" >> gpurun_out/san_trace.log
