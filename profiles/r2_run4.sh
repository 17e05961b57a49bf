set -x
mkdir -p gpurun_out
python profiles/loop_profile.py nt 16000 200 > gpurun_out/r2_loop_prof_nt.log 2>&1; cat gpurun_out/r2_loop_prof_nt.log
python profiles/loop_profile.py aa 8000 1287 > gpurun_out/r2_loop_prof_aa.log 2>&1; cat gpurun_out/r2_loop_prof_aa.log
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_nj_step -s 6000 -c 2 -o gpurun_out/prof_nj_step_nt -f python profiles/loop_profile.py nt 8000 200 > gpurun_out/prof_nj_step_nt.log 2>&1; tail -3 gpurun_out/prof_nj_step_nt.log
ls -la gpurun_out | tail -5
