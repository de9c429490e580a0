#!/bin/bash
# Round 2, GPU call 24: final validation of the committed code: GPU suite, smoke, default bench line (all legs), reference arm.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"; python scripts/bench_summary.py $O/bench_default.json default
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-160 $O/bench_ref.json
du -sh $O
