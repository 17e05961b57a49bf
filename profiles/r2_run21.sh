mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 2 > gpurun_out/r2m_c3_n4.json 2> gpurun_out/r2m_c3_n4.err; cut -c1-1000 gpurun_out/r2m_c3_n4.json; grep "profiled pass\|per step" gpurun_out/r2m_c3_n4.err; tail -3 gpurun_out/r2m_c3_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2m_c3_n2.json 2> gpurun_out/r2m_c3_n2.err; cut -c1-300 gpurun_out/r2m_c3_n2.json; grep "per step" gpurun_out/r2m_c3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --multi replicas --steps 3 --warmup 2 > gpurun_out/r2m_c3_n4_replicas.json 2> gpurun_out/r2m_c3_n4_replicas.err; cut -c1-300 gpurun_out/r2m_c3_n4_replicas.json; grep "per step" gpurun_out/r2m_c3_n4_replicas.err
