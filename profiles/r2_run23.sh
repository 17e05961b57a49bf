mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_gpu_tests.log 2>&1; tail -4 gpurun_out/r2o_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; cut -c1-330 gpurun_out/r2o_bench.json; grep "profiled pass\|per step" gpurun_out/r2o_bench.err; grep -o '"e2e": {[^}]*}' gpurun_out/r2o_bench.json; grep -o '"cpu_baseline": {"value": [0-9.]*' gpurun_out/r2o_bench.json; grep -o '"identical": [a-z]*' gpurun_out/r2o_bench.json
cap() { # name regex skip count
  timeout 300 ncu --set full --import-source on --clock-control none --kill 1 -k regex:$2 -s $3 -c $4 -o /tmp/ncu/$1 -f python profiles/loop_profile.py aa 20000 1287 0 > /tmp/ncu/$1.log 2>&1
  python profiles/ncu_stalls.py /tmp/ncu/$1.ncu-rep gpurun_out/r2o_$1_stalls.txt
  python profiles/ncu_lines.py /tmp/ncu/$1.ncu-rep gpurun_out/r2o_$1_lines.txt 40
}
cap sweep k_sweep20 700 6
cap wide k_eval_wide 3000 2
cap avg k_average 3000 2
grep "== \|gpu__time\|dram__bytes\|issue_active\|warps_active" gpurun_out/r2o_sweep_stalls.txt | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400000 --csv --log-file /tmp/ncu/launches.csv python bench.py --steps 1 --warmup 0 --skip-cpu > /tmp/ncu/launches.log 2>&1
python profiles/ncu_launch_table.py /tmp/ncu/launches.csv gpurun_out/r2o_launches_bench.md "python bench.py --steps 1 --warmup 0 --skip-cpu"; cat gpurun_out/r2o_launches_bench.md
du -sh gpurun_out
