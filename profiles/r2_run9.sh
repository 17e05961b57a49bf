set -x
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2_gpu_tests.log 2>&1; tail -15 gpurun_out/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; cat gpurun_out/r2_bench_try.json; tail -14 gpurun_out/r2_bench_try.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_try.json 2> gpurun_out/r2_bench_ref_try.err; cat gpurun_out/r2_bench_ref_try.json
# ncu: the refresh sweeps and batches of the C3-shaped workload (full set, a few launches each, summarised here)
timeout 900 ncu --set full --clock-control none -k regex:"k_one_vs_all_warp|k_out_distance_all|k_topk_select|k_merge" -s 300 -c 24 -o /tmp/ncu/sweeps -f python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2_ncu_sweeps.log 2>&1
python profiles/ncu_kernel_table.py /tmp/ncu/sweeps.ncu-rep gpurun_out/r2_ncu_sweeps_aa20k.md "refresh sweeps, aa 20 000 x 1287 fp32 (host-driven loop)"
timeout 900 ncu --set full --clock-control none -k regex:"k_eval" -s 4000 -c 40 -o /tmp/ncu/evals -f python profiles/loop_profile.py aa 20000 1287 0 > gpurun_out/r2_ncu_evals.log 2>&1
python profiles/ncu_kernel_table.py /tmp/ncu/evals.ncu-rep gpurun_out/r2_ncu_evals_aa20k.md "k_eval launches (per-join lists and refresh batches), aa 20 000 x 1287 fp32 (host-driven loop)"
cat gpurun_out/r2_ncu_sweeps_aa20k.md gpurun_out/r2_ncu_evals_aa20k.md
du -sh gpurun_out
