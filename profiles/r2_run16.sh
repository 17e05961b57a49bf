set -x
mkdir -p gpurun_out
for v in "VFT_SWEEP_MODES=7" "VFT_SWEEP_MODES=4" "VFT_SWEEP_MODES=6"; do
  echo "=== $v" >> gpurun_out/r2h_variants.log
  env $v timeout 300 python profiles/loop_profile.py aa 20000 1287 0 >> gpurun_out/r2h_variants.log 2>&1
done
echo "=== nt 16000 (sweeps not used: no matrix)" >> gpurun_out/r2h_variants.log
timeout 300 python profiles/loop_profile.py nt 16000 200 0 >> gpurun_out/r2h_variants.log 2>&1
cat gpurun_out/r2h_variants.log
