#!/bin/bash
# Run under gpurun (one GPU): the likelihood kernels behind the lock-step optimisers (csrc/ml_opt.cpp).
#   ml_opt_4000.log      rounds of optimizeAllBranchLengths at 4 000 x 1287 aa (level schedule), kernel times per round
#   ml_sweep_20k.log     recomputeMLProfiles + treeLogLk at 20 000 x 1287 aa (whole-tree batches)
#   prof_loglk_*.ncu-rep --set full of k_pair_loglk in both regimes: ~130 items x 8 warps (a Brent round) / one item per warp
set -x
mkdir -p gpurun_out
python profiles/ml_opt.py 4000 1287 ${1:-0} > gpurun_out/ml_opt_4000.log 2>&1
python profiles/ml_sweep.py 20000 1287 > gpurun_out/ml_sweep_20k.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_pair_loglk -s 600 -c 2 -o gpurun_out/prof_loglk_round -f python profiles/ml_opt.py 4000 1287 0 > gpurun_out/prof_loglk_round.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_pair_loglk -c 2 -o gpurun_out/prof_loglk_tree -f python profiles/ml_sweep.py 20000 1287 > gpurun_out/prof_loglk_tree.log 2>&1
ls -la gpurun_out | tail -8
# (added later in the round) one tree level of recomputeMLProfiles, and the rebuilt out-profile kernel on a C2-sized step
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_posterior -s 3 -c 1 -o gpurun_out/prof_posterior -f python profiles/ml_sweep.py 20000 1287 > gpurun_out/prof_posterior.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_outprofile_rebuild -s 2 -c 1 -o gpurun_out/prof_rebuild -f python profiles/one_step.py 16000 > gpurun_out/prof_rebuild.log 2>&1
