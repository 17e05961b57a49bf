set -x
mkdir -p gpurun_out
for v in "VFT_SWEEP=0" "VFT_SWEEP_ROWS=8" "VFT_SWEEP_ROWS=16" "VFT_SWEEP_ROWS=32" "VFT_SWEEP_ROWS=8 VFT_AVG_SPLIT=0"; do
  echo "=== $v" >> gpurun_out/r2e_variants.log
  env $v timeout 300 python profiles/loop_profile.py aa 20000 1287 0 >> gpurun_out/r2e_variants.log 2>&1
done
cat gpurun_out/r2e_variants.log
mkdir -p /tmp/ncu
timeout 300 ncu --set full --import-source on --clock-control none --kill 1 -k regex:k_sweep20 -s 700 -c 6 -o /tmp/ncu/sweep -f python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2e_ncu_sweep.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --kill 1 -k regex:k_eval_wide -s 3000 -c 2 -o /tmp/ncu/wide -f python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2e_ncu_wide.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --kill 1 -k regex:k_average -s 3000 -c 2 -o /tmp/ncu/avg -f python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2e_ncu_avg.log 2>&1
cp /tmp/ncu/*.ncu-rep gpurun_out/; ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2e_ncu_sweep.log
