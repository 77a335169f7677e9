#!/bin/bash
# One gpurun call: run the GPU test tiers with individual timeouts, keep every log under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== debug_gemm tn"; timeout 300 python scripts/debug_gemm.py tn > gpurun_out/debug_tn.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/debug_tn.log
echo "== debug_gemm nt"; timeout 300 python scripts/debug_gemm.py nt > gpurun_out/debug_nt.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/debug_nt.log
echo "== pytest gemm"; timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gemm.log
echo "== pytest mlp"; timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu > gpurun_out/pytest_mlp.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_mlp.log
