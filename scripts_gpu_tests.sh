#!/bin/bash
# Runs the GPU test groups as separate, individually time-limited processes (a deadlocked kernel only loses its group).
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${TMO:-240} python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 25 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,memory.total --format=csv
run glue tests/test_kernels_gpu.py -k "cfg_euler or ln_modulate or gemv or timestep or rmsnorm or qk_norm_rope"
run gemm tests/test_kernels_gpu.py -k "gemm"
run attn tests/test_kernels_gpu.py -k "attention"
run forward tests/test_forward_gpu.py
run parallel tests/test_parallel_gpu.py
