mkdir -p gpurun_out
for v in "VFT_SELF_DEFER=1" "VFT_SELF_DEFER=0"; do
  echo "=== $v" >> gpurun_out/r2j_durations.log
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kill 1 -k regex:"k_average|k_eval_wide|k_self_sum" -s 6000 -c 12 --csv python profiles/loop_profile.py aa 20000 1287 0 2>/dev/null | python profiles/ncu_durations.py >> gpurun_out/r2j_durations.log
done
cat gpurun_out/r2j_durations.log
