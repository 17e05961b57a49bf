set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|avx512" | head -5
grep -o -m1 'avx512[a-z]*' /proc/cpuinfo | sort -u | head
timeout 1500 python -m pytest tests/test_gpu_at_size.py -x -q -m gpu --durations=10 > gpurun_out/r2_at_size.log 2>&1; tail -25 gpurun_out/r2_at_size.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_parity.log 2>&1; tail -5 gpurun_out/r2_parity.log
timeout 900 python profiles/ref_aa_timing.py 8000,20000 8,16,32 > gpurun_out/r2_ref_aa_timing.log 2>&1; cat gpurun_out/r2_ref_aa_timing.log
