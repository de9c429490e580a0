#!/bin/bash
# One short gpurun call: the whole GPU suite with persistent packed weights switched on, then the bench with and without them.
set -u
mkdir -p gpurun_out
TPGSR_PREPACK=1 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_prepack.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_prepack.log
tail -8 gpurun_out/pytest_prepack.log
TPGSR_PREPACK=1 TPGSR_BENCH_NO_CPU=1 timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_prepack1.json 2> gpurun_out/bench_prepack1.err; echo "bench1 rc=$?"
cat gpurun_out/bench_prepack1.json
TPGSR_PREPACK=0 TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 200 python bench.py --no-e2e --steps 40 --warmup 5 > gpurun_out/bench_prepack0.json 2> gpurun_out/bench_prepack0.err; echo "bench0 rc=$?"
cat gpurun_out/bench_prepack0.json
