set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/t_loop_gpu.py quick > gpurun_out/r2_loop_memcheck.log 2>&1; tail -12 gpurun_out/r2_loop_memcheck.log
timeout 900 python profiles/t_loop_gpu.py > gpurun_out/r2_loop.log 2>&1; cat gpurun_out/r2_loop.log
