#!/bin/bash
# One gpurun call: probes, GPU parity tests, bench (both arms), ncu launch list of one eager step, ncu --set full of the top kernels.
# Usage (repo root, on the GPU box): bash scripts/gpu_check.sh [flat] [timing] [tests] [smoke] [bench] [refarm] [launches] [full]
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches full}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
if [[ $what == *flat* ]]; then
  timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -x > gpurun_out/pytest_flat.log 2>&1; echo "pytest flat rc=$?"
  tail -15 gpurun_out/pytest_flat.log
fi
if [[ $what == *timing* ]]; then
  timeout 300 python scripts/time_conv.py all > gpurun_out/time_conv.log 2>&1; echo "timing rc=$?"; cat gpurun_out/time_conv.log
fi
if [[ $what == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
if [[ $what == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
  TPGSR_CONV_MODE=bf16 TPGSR_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
  cat gpurun_out/bench_bf16.json
fi
if [[ $what == *refarm* ]]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  cat gpurun_out/bench_ref.json
fi
if [[ $what == *launches* ]]; then
  # eager (non-graphed) step so every launch is listed
  # only the last ~1.1 eager steps are instrumented (-s skips the accounting pass and the warm-ups): listing all ~2800 launches costs 8 min
  TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LAUNCH_SKIP:-2100} -c 700 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e > gpurun_out/launches.log 2>&1
  echo "launches rc=$?"; wc -l gpurun_out/launches.csv
fi
if [[ $what == *full* ]]; then
  TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'conv_flat_kernel|conv_wgrad_flat_kernel|gru_fwd_kernel|gru_bwd_kernel|gru_dwhh_kernel' -s ${FULL_SKIP:-70} -c ${FULL_COUNT:-10} -o gpurun_out/prof_top -f \
      python bench.py --steps 1 --warmup 3 --no-graph --no-e2e > gpurun_out/full.log 2>&1
  echo "full rc=$?"; ls -la gpurun_out/
fi
