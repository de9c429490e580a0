#!/bin/bash
# Round 2, GPU call 9: operand planes from the BatchNorm apply passes (stand-alone finalize kernels) A/B against separate split passes.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_prior.py tests/test_gpu_cascade.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $O/pytest_gpu.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 30 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1"
BARGS="--config 2" run_bench c2 $QUICK
BARGS="--config 2" run_bench c2_noplanes $QUICK TPGSR_BN_PLANES=0
BARGS="--config 3" run_bench c3 $QUICK
BARGS="--config 3" run_bench c3_noplanes $QUICK TPGSR_BN_PLANES=0
BARGS="--config 2" run_bench c2_again $QUICK
BARGS="--config 3" run_bench c3_again $QUICK
du -sh $O
