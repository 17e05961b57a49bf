set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc; nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2c_sharded_tests.log; cat gpurun_out/r2c_sharded_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err; cat gpurun_out/r2c_bench1.json | cut -c1-400; grep "profiled pass" gpurun_out/r2c_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err; cat gpurun_out/r2c_bench2.json; grep "profiled pass\|last timed" gpurun_out/r2c_bench2.err
VFT_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2c_bench2_nccl.json 2> gpurun_out/r2c_bench2_nccl.err; cat gpurun_out/r2c_bench2_nccl.json | cut -c1-1200; tail -5 gpurun_out/r2c_bench2_nccl.err
