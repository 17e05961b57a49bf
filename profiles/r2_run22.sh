mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -s 2>&1 | grep "sharded\|passed\|failed" | tail -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2n_c3_n2.json 2> gpurun_out/r2n_c3_n2.err; cut -c1-250 gpurun_out/r2n_c3_n2.json; grep "per step\|identical" gpurun_out/r2n_c3_n2.err; grep -o '"trees_identical_across_ranks": [a-z]*' gpurun_out/r2n_c3_n2.json
