#!/bin/bash
# Round 2, GPU call 25 (2 GPUs): final validation of the committed code: the whole GPU suite including the 2-rank NCCL parity test,
# smoke, the default bench line at N = 1 and N = 2.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 timeout 400 python bench.py --steps 30 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; python scripts/bench_summary.py $O/bench_n1.json n1 | cut -c1-330
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?"; python scripts/bench_summary.py $O/bench_n2.json n2 | cut -c1-120; grep -c "NCCL INFO" $O/bench_n2.err; tail -c 1500 $O/bench_n2.err > $O/bench_n2.err.tail; rm -f $O/bench_n2.err
du -sh $O
