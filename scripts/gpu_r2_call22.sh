#!/bin/bash
# Round 2, GPU call 22: the record run after the recurrence / pooling / conv0 work: GPU suite, smoke, the default bench line with all legs,
# configs 2 / 4 / 5, the reference arm.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"; python scripts/bench_summary.py $O/bench_default.json default
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1"
for c in 2 4 5; do env $QUICK timeout 500 python bench.py --config $c --steps 30 --warmup 5 > $O/bench_c$c.json 2> $O/bench_c$c.err; python scripts/bench_summary.py $O/bench_c$c.json c$c | cut -c1-330; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/bench_ref.json
du -sh $O
