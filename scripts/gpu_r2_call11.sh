#!/bin/bash
# Round 2, GPU call 11: LSTM stores behind the cluster arrive, teacher packs kept across steps, dead-branch removal, smoke() with the cascade step:
# GPU suite, smoke, default bench line (config 3, all legs) and config 2 / 4 / 5 lines.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_stats.jsonl
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1"
BARGS="" run_bench default A=1
BARGS="--config 2 --steps 30 --warmup 5" run_bench c2 $QUICK
BARGS="--config 4 --steps 20 --warmup 5" run_bench c4 $QUICK
BARGS="--config 5 --steps 30 --warmup 5" run_bench c5 $QUICK
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/bench_ref.json
timeout 200 python scripts/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.log 2>&1; grep -i "lstm\|tps\|pixelshuffle" $O/hbm_kernels.log
du -sh $O
