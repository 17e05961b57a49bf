mkdir -p gpurun_out
for v in "VFT_SELF_DEFER=1" "VFT_SELF_DEFER=0"; do
  echo "=== $v" >> gpurun_out/r2k_variants.log
  env $v timeout 300 python profiles/loop_profile.py aa 20000 1287 0 >> gpurun_out/r2k_variants.log 2>&1
done
cat gpurun_out/r2k_variants.log
timeout 300 python profiles/loop_profile.py nt 16000 200 0 2>&1 | head -3
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_gpu_tests.log 2>&1; tail -5 gpurun_out/r2k_gpu_tests.log
