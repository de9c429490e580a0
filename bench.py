#!/usr/bin/env python
"""bench.py — SR images/sec of the TPGSR hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 20 --warmup 5            # our CUDA path (default arm)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle port) on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                # one rank per GPU, weak scaling (48 images / GPU)

Workload at N=1 (BASELINE.json configs[1]): TSRN single-stage, STN on, mask on, batch 48, 16x64 -> 32x128, one TRAIN
step = forward + ImageLoss(gradient)*100 + backward + clip_grad_norm_(0.25) + Adam(1e-3, (0.5, 0.999)); synthetic
U[0,1) crops, random-init weights of the reference architecture.

  value    images/s with the step's inputs already resident in HBM (CUDA-graphed step, CUDA-event timing, max over ranks)
  e2e      images/s through the public API with HOST (pinned) input buffers: H2D of LR+HR and D2H of the loss inside the
           timed region, every step
  roofline dominant kernel (3x3 64->64 conv, the shape that is 11 of the 18 conv launches per forward), timed live with
           CUDA events on the launch stream; achieved = algorithmic FLOPs / launch time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the same step on the box's host cores (bounded sample), rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
PER_GPU_BATCH = 48
METRIC = 'SR images/sec (16x64->32x128), TSRN train step'
# one description of the workload for both arms (the driver compares the two lines' metric / config)
WORKLOAD = ('BASELINE configs[1]: TSRN single-stage TRAIN step, STN on, mask, ImageLoss(gradient,[1,1e-4])*100, '
            'clip_grad_norm 0.25, Adam(1e-3,(0.5,0.999))')
# SURVEY §8d / BASELINE.md §3: TSRN train-mode forward = 1.8445 GFLOP/img, train step = 3x forward
GFLOP_PER_IMG_STEP = 5.53


# DRAM traffic per launch of the dominant kernel (3x3, 64->64, B=48) from the committed ncu --set full captures (profiles/r01/ncu_*.csv):
# 13.8 MB read + 0 written inside the kernel's window (12.6 MB of operand planes; the 12.6 MB fp32 output stays in the 126 MB L2)
NCU_TRAFFIC_BYTES = {('conv_flat_kernel', 3): 13.77e6, ('conv_tc_kernel', 3): 12.81e6}


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_burst': d['bf16_tflops'], 'tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(self.rows[0][1]) if self.rows[0][1].replace('.', '').isdigit() else None,
                'power_w_max': max(float(r[2]) for r in self.rows if r[2].replace('.', '').isdigit()) if self.rows else None,
                'samples': len(self.rows), 'reasons': sorted(reasons)}


def synth_batches(n_batches: int, batch: int, seed: int, device, pin: bool):
    from numpy.random import RandomState
    out = []
    for i in range(n_batches):
        r = RandomState(seed + i)
        lr = torch.from_numpy(r.rand(batch, 4, 16, 64).astype('float32'))
        hr = torch.from_numpy(r.rand(batch, 4, 32, 128).astype('float32'))
        lr[:, 3] = (lr[:, 3] > 0.5).float()
        hr[:, 3] = (hr[:, 3] > 0.5).float()
        if pin:
            out.append((lr.pin_memory(), hr.pin_memory()))
        else:
            out.append((lr.to(device), hr.to(device)))
    return out


# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_images_per_s(sample_batch: int, steps: int, warmup: int, threads: int):
    """The reference's step on host cores: oracle port of TSRN(STN)+ImageLoss+clip+Adam (oracle/step.py)."""
    from oracle import step as ostep
    from oracle.weights import synth_images, synth_state_dict
    from oracle.ops import tps_buffers
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))['tsrn_train']
    p = synth_state_dict(shapes, 0)
    for k, v in tps_buffers(16, 64).items():
        p['tps.' + k] = v
    opt = ostep.AdamState()
    # pick the intra-op thread count that is fastest for this op mix (many small GEMMs: all 128 hyper-threads of the
    # GPU box are ~100x slower than 16); candidates are timed on a 2-image step
    best = None
    for th in sorted({t for t in (8, 16, 32, 64, threads) if t <= threads}):
        torch.set_num_threads(th)
        lr, hr = synth_images(2, seed=99)
        ostep.tsrn_train_step(p, lr, hr, None)
        t0 = time.perf_counter()
        ostep.tsrn_train_step(p, lr, hr, None)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, th)
        if dt > 20:
            break
    torch.set_num_threads(best[1])
    cpu_port_images_per_s.threads_used = best[1]
    times = []
    for i in range(warmup + steps):
        lr, hr = synth_images(sample_batch, seed=1234 + i)
        t0 = time.perf_counter()
        ostep.tsrn_train_step(p, lr, hr, opt)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return sample_batch * len(times) / total, total / len(times)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 12
    ips, sec = cpu_port_images_per_s(sample, args.steps, args.warmup, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': PER_GPU_BATCH, 'lr_hw': [16, 64], 'hr_hw': [32, 128],
                   'implementation': 'CPU oracle port of the same step (oracle/step.py), torch CPU fp32', 'sample_batch_per_step': sample},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cpu_port_images_per_s.threads_used, 'host_cpus': threads, 'kind': 'port',
                         'sample': f'{args.steps} timed steps of {sample} images (1/4 of the 48-image batch), torch {torch.__version__} CPU fp32'},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def time_dominant_kernel(device, batch: int, peaks):
    """3x3 64->64 conv forward on [batch,16,64,64] NHWC through the tcgen05 kernel (tpgsr_conv2d_tc) in the active
    CONV_MODE: CUDA-event timing on the launch stream (torch's current stream = the stream the C ABI is handed), L2
    flushed between launches.  achieved = ALGORITHMIC flops (2*M*N*K, one product — the 3 MMAs of split-bf16 are not
    counted three times) / launch time."""
    from tpgsr_b200 import ops
    x = torch.randn(batch, 16, 64, 64, device=device)
    w = torch.randn(64, 64, 3, 3, device=device) * 0.05
    b = torch.zeros(64, device=device)
    need_lo = ops.CONV_MODE == 'bf16x3'
    (wf, _) = ops.pack_conv_weights_bf16(w, False, need_lo)
    y = torch.empty(batch, 16, 64, 64, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flat = ops.flat_supported(batch, 16, 64, 1, 64, 64, 3, 3, 1, 1)
    if flat:
        fp = ops.split_bf16_flat(x, 1, need_lo)
        run = lambda: ops.conv_flat(fp, wf, b, 64, 3, 3, 1, 1, out=y)
    else:
        xp = ops.split_bf16(x, need_lo)
        run = lambda: ops.conv_tc(xp, wf, b, batch, 16, 64, 64, 64, 3, 3, 1, 1, out=y)
    for _ in range(3):
        run()
    torch.cuda.synchronize(device)
    durs = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        e1.synchronize()
        durs.append(e0.elapsed_time(e1) * 1e-3)
    dur = sum(durs) / len(durs)
    flops = 2.0 * batch * 16 * 64 * 64 * 576
    ach = flops / dur / 1e12
    nprod = 3 if need_lo else 1
    kname = 'conv_flat_kernel' if flat else 'conv_tc_kernel'
    return {'bound': 'tensor', 'kernel': '%s<%d> (tcgen05+TMA implicit GEMM, 3x3, 64->64, M=%d, N=64, K=576, mode %s)' % (kname, nprod, batch * 1024, ops.CONV_MODE),
            'achieved': ach, 'peak': peaks['tflops_burst'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tflops_burst'],
            'traffic': NCU_TRAFFIC_BYTES.get((kname, nprod)), 'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/r01',
            'launch_us': dur * 1e6, 'launch_us_min': min(durs) * 1e6, 'algorithmic_flops_per_launch': flops,
            'mma_flops_per_launch': flops * nprod, 'tensor_pipe_tflops': ach * nprod,
            'peak_source': peaks['source'] + ', bf16 burst (kernel timed alone)'}


def run_ours(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the tpgsr_b200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        # rank 0 must print exactly one JSON line on stdout, and NCCL logs there (its version banner at any NCCL_DEBUG level, WARN included):
        # send NCCL's log to stderr, drop the debug level, and park stdout on stderr until the JSON line is ready
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if not os.environ.get('TPGSR_KEEP_NCCL_DEBUG'):
            os.environ.pop('NCCL_DEBUG', None)
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
    from tpgsr_b200 import ops
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import TSRNTrainStep
    torch.manual_seed(0)
    net = TSRN(**KW).to(dev).train()
    B = args.batch
    ts = TSRNTrainStep(net, batch=B, use_graph=not args.no_graph)
    dev_batches = synth_batches(8, B, 1234 + 100 * rank, dev, pin=False)
    host_batches = synth_batches(8, B, 1234 + 100 * rank, dev, pin=True)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # launches per step (eager accounting pass; the graph replays exactly these kernels)
    l0 = ops.LAUNCHES
    ts.load_batch(*dev_batches[0])
    ts._fwd_bwd()
    ts._optim()
    launches_per_step = ops.LAUNCHES - l0
    torch.cuda.synchronize(dev)

    def timed(batches, e2e: bool):
        losses = []
        for i in range(args.warmup):
            ts(*batches[i % len(batches)])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            loss = ts(*batches[i % len(batches)])
            if e2e:
                losses.append(loss.to('cpu', non_blocking=False))      # D2H of the step's result, every step
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, 0.0)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t) * 1e-3, wall, (float(losses[-1]) if losses else float(ts.loss))

    with ClockSampler(local) as clocks:
        sec, wall, last_loss = timed(dev_batches, False)
    if args.no_e2e:          # developer switch for profiler runs (ncu launch lists): one timed leg only
        sec_e2e, wall_e2e = sec, wall
    else:
        sec_e2e, wall_e2e, _ = timed(host_batches, True)
    sec_e2e = max(sec_e2e, wall_e2e) if world == 1 else sec_e2e      # host-visible time when the D2H sync dominates

    imgs = B * world * args.steps
    value = imgs / sec
    e2e_val = imgs / sec_e2e
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peaks = _peaks()
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if ops.CONV_MODE != 'bf16' else 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': B, 'global_batch': B * world,
                   'lr_hw': [16, 64], 'hr_hw': [32, 128], 'parallelism': f'dp{world}', 'cuda_graph': not args.no_graph,
                   'precision': {'bf16x3': 'fp32 tensors everywhere; conv / 1x1 / linear contractions as split-bf16 (hi*hi + hi*lo + lo*hi) on tcgen05 with '
                                           'fp32 TMEM accumulation = fp32-class accuracy (the parity mode, SR output within 1e-3 of the fp32 reference); '
                                           'GRU / BN / TPS / losses / Adam in fp32',
                                 'bf16': 'single bf16 product per contraction (throughput mode; misses the 1e-3 parity budget, labelled)',
                                 'fp32': 'CUDA-core fp32 contractions'}[ops.CONV_MODE],
                   'l2': 'no explicit flush: the step touches >1 GB of activations/stashes per replay (>> 126 MB L2); inputs rotate over 8 distinct batches',
                   'last_loss': last_loss,
                   # non-default (experimental) switches in force, so that A/B lines describe themselves
                   'switches': {k: True for k, on in (('TPGSR_PREPACK', ops.PREPACK), ('TPGSR_FLAT_OCC2', ops.FLAT_OCC2),
                                                      ('TPGSR_BWD_PRODUCTS=2', ops.BWD2), ('TPGSR_WGRAD_PRODUCTS=1', ops.WGRAD1)) if on}},
        'e2e': {'value': e2e_val, 'unit': 'images/s', 'h2d_bytes_per_step': B * (4 * 16 * 64 + 4 * 32 * 128) * 4, 'd2h_bytes_per_step': 4,
                'ms_per_step': sec_e2e / args.steps * 1e3, 'api': 'tpgsr_b200.train.TSRNTrainStep(images_lr_pinned, images_hr_pinned) -> loss.cpu()'},
        'gpu_launches': launches_per_step * args.steps,
        'launches_per_step': launches_per_step,
        'clocks': clocks.summary(),
        'step_tflops': value * GFLOP_PER_IMG_STEP / 1e3,
        'step_frac_of_sustained_bf16_peak': value * GFLOP_PER_IMG_STEP / 1e3 / peaks['tflops_sustained'],
    }
    if world == 1 and not os.environ.get('TPGSR_BENCH_NO_ROOFLINE'):
        line['roofline'] = time_dominant_kernel(dev, B, peaks)
    if world == 1 and not os.environ.get('TPGSR_BENCH_NO_CPU'):     # (developer switch: skip the ~60 s CPU leg)
        threads = os.cpu_count() or 1
        ips, sec_cpu = cpu_port_images_per_s(12, 2, 1, threads)
        line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': cpu_port_images_per_s.threads_used, 'host_cpus': threads, 'kind': 'port',
                                'sample': '2 timed steps of 12 images (1/4 batch) of the same train step, oracle port, torch CPU fp32'}
    if world == 1 and os.environ.get('TPGSR_BENCH_INFER') == '1':
        # optional extra (SURVEY §8d also asks for inference images/s): eval-mode forward through the public nn.Module API (no STN, running BN
        # statistics), device-resident inputs, eager launches.  Off by default: the contract line is the train step.
        try:
            net.eval()
            x = dev_batches[0][0]
            with torch.no_grad():
                for _ in range(args.warmup):
                    net(x)
                torch.cuda.synchronize(dev)
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                i0.record()
                for i in range(args.steps):
                    net(dev_batches[i % len(dev_batches)][0])
                i1.record()
                torch.cuda.synchronize(dev)
            line['inference'] = {'value': B * args.steps / (i0.elapsed_time(i1) * 1e-3), 'unit': 'images/s', 'batch': B,
                                 'api': 'TSRN.eval()(images_lr) under torch.no_grad(), eager'}
        except Exception as e:      # never let the optional leg cost the contract line
            line['inference'] = {'error': repr(e)}
        finally:
            net.train()
    if world > 1:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)                     # anything NCCL says while shutting down goes to stderr as well
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=PER_GPU_BATCH)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='developer switch: skip the host-buffer leg (profiler runs)')
    args = ap.parse_args()
    if os.environ.get('TPGSR_BENCH_TRACE'):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['TPGSR_BENCH_TRACE']), exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
