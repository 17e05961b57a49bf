set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_parity.log 2>&1; tail -5 gpurun_out/r2_parity.log
timeout 300 python profiles/t_loop_gpu.py quick > gpurun_out/r2_loop_quick.log 2>&1; cat gpurun_out/r2_loop_quick.log
timeout 300 python profiles/loop_profile.py nt 16000 200 > gpurun_out/r2_loop_prof_nt.log 2>&1; cat gpurun_out/r2_loop_prof_nt.log
timeout 300 python profiles/loop_profile.py aa 20000 1287 > gpurun_out/r2_loop_prof_aa.log 2>&1; cat gpurun_out/r2_loop_prof_aa.log
timeout 300 python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2_host_prof_aa.log 2>&1; cat gpurun_out/r2_host_prof_aa.log
VFT_NO_STAGING=1 timeout 300 python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2_host_prof_aa_nostage.log 2>&1; cat gpurun_out/r2_host_prof_aa_nostage.log
