set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2b_gpu_tests.log 2>&1; tail -15 gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json; tail -14 gpurun_out/r2b_bench.err
