set -x
mkdir -p gpurun_out
python profiles/t_loop_gpu.py quick > gpurun_out/r2_loop_quick.log 2>&1; cat gpurun_out/r2_loop_quick.log
python profiles/loop_profile.py nt 16000 200 > gpurun_out/r2_loop_prof_nt.log 2>&1; cat gpurun_out/r2_loop_prof_nt.log
python profiles/loop_profile.py aa 8000 1287 > gpurun_out/r2_loop_prof_aa.log 2>&1; cat gpurun_out/r2_loop_prof_aa.log
