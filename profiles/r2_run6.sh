set -x
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none --cache-control none --warp-sampling-interval 0 -k regex:k_nj_step -s 6000 -c 16 -o /tmp/ncu/step -f python profiles/loop_profile.py nt 8000 200 > gpurun_out/prof_nj_step_nt2.log 2>&1; tail -3 gpurun_out/prof_nj_step_nt2.log
python profiles/ncu_lines.py /tmp/ncu/step.ncu-rep gpurun_out/r2_k_nj_step_lines.txt 90
ncu -i /tmp/ncu/step.ncu-rep --page raw --csv > gpurun_out/r2_k_nj_step_raw.csv 2>/dev/null
head -60 gpurun_out/r2_k_nj_step_lines.txt
du -sh gpurun_out
