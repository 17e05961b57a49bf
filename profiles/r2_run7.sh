set -x
mkdir -p gpurun_out
python profiles/t_loop_gpu.py quick > gpurun_out/r2_loop_quick.log 2>&1; cat gpurun_out/r2_loop_quick.log
python profiles/loop_profile.py nt 16000 200 > gpurun_out/r2_loop_prof_nt.log 2>&1; cat gpurun_out/r2_loop_prof_nt.log
python profiles/loop_profile.py aa 20000 1287 > gpurun_out/r2_loop_prof_aa.log 2>&1; cat gpurun_out/r2_loop_prof_aa.log
python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2_host_prof_aa.log 2>&1; cat gpurun_out/r2_host_prof_aa.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; cat gpurun_out/r2_bench_try.json; tail -5 gpurun_out/r2_bench_try.err
