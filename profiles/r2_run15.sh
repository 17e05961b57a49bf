set -x
mkdir -p gpurun_out
timeout 300 python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2g_loop.log 2>&1; cat gpurun_out/r2g_loop.log
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2g_gpu_tests.log 2>&1; tail -12 gpurun_out/r2g_gpu_tests.log
