set -x
mkdir -p gpurun_out /tmp/ncu
cap() { # name regex skip count
  timeout 300 ncu --set full --import-source on --clock-control none --kill 1 -k regex:$2 -s $3 -c $4 -o /tmp/ncu/$1 -f python profiles/loop_profile.py aa 20000 1287 0 > /tmp/ncu/$1.log 2>&1
  python profiles/ncu_stalls.py /tmp/ncu/$1.ncu-rep gpurun_out/r2f_$1_stalls.txt
  python profiles/ncu_lines.py /tmp/ncu/$1.ncu-rep gpurun_out/r2f_$1_lines.txt 45
}
cap sweep k_sweep20 700 4
cap wide k_eval_wide 3000 2
cap avg k_average 3000 2
cat gpurun_out/r2f_sweep_stalls.txt gpurun_out/r2f_wide_stalls.txt gpurun_out/r2f_avg_stalls.txt
timeout 600 python -m pytest tests/test_ingest.py -q -m gpu 2>&1 | tail -3
du -sh gpurun_out
