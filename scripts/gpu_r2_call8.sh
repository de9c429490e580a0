#!/bin/bash
# Round 2, GPU call 8: fused BatchNorm passes (last-block finalize, planes from the apply pass): GPU suite, A/B on configs 2 / 3,
# graph-node overhead microbenchmark.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_stats.jsonl
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 $O/pytest_gpu.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 20 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
BARGS="--config 2" run_bench c2 $QUICK
BARGS="--config 2" run_bench c2_bn3 $QUICK TPGSR_BN_FUSED=0
BARGS="--config 3" run_bench c3 $QUICK
BARGS="--config 3" run_bench c3_bn3 $QUICK TPGSR_BN_FUSED=0
timeout 200 python scripts/graph_overhead.py > $O/graph_overhead.log 2>&1; cat $O/graph_overhead.log
du -sh $O
