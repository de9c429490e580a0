#!/bin/bash
# Round 2, GPU call 17 (8 GPUs): the default bench line (config 3) at N = 8 and N = 4 exactly as the driver launches it, config 4 at N = 8,
# and the 2-rank NCCL parity test on this box.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
run_n() {  # n cfg port
  NCCL_DEBUG=INFO TPGSR_BENCH_CONFIG=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 \
      bench.py --gpus $1 --steps 20 --warmup 5 > $O/bench_n$1_c$2.json 2> $O/bench_n$1_c$2.err
  echo "bench n$1 c$2 rc=$?"; python scripts/bench_summary.py $O/bench_n$1_c$2.json n$1_c$2 | cut -c1-200; grep -c "NCCL INFO" $O/bench_n$1_c$2.err
  grep -m3 "NVLS\|nranks" $O/bench_n$1_c$2.err | cut -c1-220
  tail -c 2000 $O/bench_n$1_c$2.err > $O/bench_n$1_c$2.err.tail; rm -f $O/bench_n$1_c$2.err
}
run_n 8 3 29521
run_n 4 3 29522
run_n 8 4 29523
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi.log
TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1 timeout 400 python bench.py --config 3 --steps 20 --warmup 5 > $O/bench_n1_c3.json 2> $O/bench_n1_c3.err; python scripts/bench_summary.py $O/bench_n1_c3.json n1_c3 | cut -c1-120
du -sh $O
