#!/bin/bash
# Run under gpurun (one GPU).  Raw ncu output goes to gpurun_out/; the summaries are committed under profiles/.
#   1. launch list: every kernel of one full C2-shaped step (4 000 taxa) with its device time
#      (cold-cache, serialised: compare SHARES, not absolutes)
#   2. --set full captures: the per-join k_eval launches of a 16 000-taxon step (warm cache), one large
#      nt batch and one large aa batch (bandwidth regime), the CTA-per-pair kernel on aa lists
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/launches.csv python profiles/one_step.py 4000 > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_eval -s 20000 -c 3 -o gpurun_out/prof_eval_small -f python profiles/one_step.py 16000 > gpurun_out/prof_eval_small.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_eval -c 1 -o gpurun_out/prof_batch_nt -f python profiles/big_batch.py 16000 nt 200 262144 1 > gpurun_out/prof_batch_nt.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_eval -c 1 -o gpurun_out/prof_batch_aa -f python profiles/big_batch.py 20000 aa 1287 131072 1 > gpurun_out/prof_batch_aa.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_eval_wide -s 3000 -c 2 -o gpurun_out/prof_wide_aa -f python profiles/scale_aa.py 4000 1287 > gpurun_out/prof_wide_aa.log 2>&1
ls -la gpurun_out
