#!/bin/bash
# Run under gpurun (one GPU).  Writes raw ncu output to gpurun_out/; summaries are committed under profiles/.
#   1. launch list: every kernel of one full step with its device time (cold-cache, serialised: SHARES only)
#   2. full-metric capture of the dominant kernel (k_eval) for dram bytes / stall reasons
set -x
TAXA=${TAXA:-4000}
mkdir -p gpurun_out
cat > /tmp/one_step.py <<PY
import sys
sys.path.insert(0, '.')
from veryfasttree_b200 import api, synth
chars = synth.make_alignment($TAXA, 200, 'nt', 1)
chars = chars[synth.unique_rows(chars)]
t = api.nj_build(api.encode(chars, 'nt'), 4, 32, trace=False)
print('taxa', chars.shape[0], 'launches', t.stats['counters']['launches'], 'device ms', t.stats['deviceMsResident'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/launches_r1.csv python /tmp/one_step.py > gpurun_out/launches_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3000 -c 3 -o gpurun_out/prof_k_eval_r1 -f python /tmp/one_step.py > gpurun_out/prof_k_eval_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_one_vs_all_warp -s 2 -c 2 -o gpurun_out/prof_k_ova_r1 -f python /tmp/one_step.py > gpurun_out/prof_k_ova_r1.log 2>&1
ls -la gpurun_out
