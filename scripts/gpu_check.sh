#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list of one eager step, ncu --set full of the top kernels.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_check.sh [tests] [bench] [launches] [full]   (default: all)
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches full}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
if [[ $what == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ $what == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json
  TPGSR_CONV_MODE=bf16 TPGSR_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
  cat gpurun_out/bench_bf16.json
fi
if [[ $what == *refarm* ]]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  cat gpurun_out/bench_ref.json
fi
if [[ $what == *launches* ]]; then
  # eager (non-graphed) step so every launch is listed; skip the warm-up/accounting passes (first ~3 steps)
  TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e > gpurun_out/launches.log 2>&1
  echo "launches rc=$?"; wc -l gpurun_out/launches.csv
fi
if [[ $what == *full* ]]; then
  TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'conv_tc_kernel|conv_wgrad_tc_kernel|gru_fwd_kernel|gru_bwd_kernel' -s 40 -c 12 -o gpurun_out/prof_top -f \
      python bench.py --steps 1 --warmup 3 --no-graph --no-e2e > gpurun_out/full.log 2>&1
  echo "full rc=$?"; ls -la gpurun_out/
fi
