set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2d_gpu_tests.log 2>&1; tail -12 gpurun_out/r2d_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cut -c1-300 gpurun_out/r2d_bench.json; grep "profiled pass\|per step\|identical" gpurun_out/r2d_bench.err; grep -o '"parity_checked": {[^}]*}' gpurun_out/r2d_bench.json | cut -c1-80
