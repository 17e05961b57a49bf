set -x
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none --cache-control none --sampling-interval 0 -k regex:k_nj_step -s 6000 -c 40 -o gpurun_out/prof_nj_step_nt2 -f python profiles/loop_profile.py nt 8000 200 > gpurun_out/prof_nj_step_nt2.log 2>&1; tail -3 gpurun_out/prof_nj_step_nt2.log
ls -la gpurun_out | tail -3
