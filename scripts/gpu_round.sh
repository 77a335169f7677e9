#!/bin/bash
# One gpurun call: run the GPU test tiers with individual timeouts, keep every log under gpurun_out/.
# usage: bash scripts/gpu_round.sh [tests] [bench] [ncu]
mkdir -p gpurun_out
WHAT="${@:-tests bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
if [[ "$WHAT" == *debug* ]]; then
  echo "== debug_gemm"; timeout 600 python scripts/debug_gemm.py all > gpurun_out/debug_gemm.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/debug_gemm.log
fi
if [[ "$WHAT" == *tests* ]]; then
  echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_gpu.log
  echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/smoke.log
fi
if [[ "$WHAT" == *bench* ]]; then
  echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
  echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
fi
if [[ "$WHAT" == *ncu* ]]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 225 -c 52 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 4 --warmup 3 --cpu-sample 256 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_bench.log
  echo "== ncu full (top kernels)"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 180 -c 20 -o gpurun_out/prof_gemm -f \
      python bench.py --steps 4 --warmup 3 --cpu-sample 256 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
