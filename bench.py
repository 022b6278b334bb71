#!/usr/bin/env python
"""bench.py -- X-UNet DDPM training throughput (BASELINE.json metric: train images/sec) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload small64|small128|full128|full64]
                    [--batch B] [--dtype bf16|fp32]

One "step" = one optimisation step (pinned H2D staging -> forward -> backward -> [NCCL all-reduce] -> Adam) on a
synthetic SRN-shaped batch of B (source,target) pairs PER GPU (weak scaling).  Default workload = BASELINE.json
configs[1]: small X-UNet (ch=32, ch_mult=(1,2), emb_ch=32, nrb=2, attn_res=(8,16,32), heads=4), 64x64, bf16, B=8.

Printed JSON (rank 0, one line):
  value  : images/s with the inputs already resident in HBM (device tensors), CUDA-graph replayed step
  e2e    : images/s through the public API (TrainStep.__call__) with HOST numpy inputs: pinned staging + H2D copies and a
           D2H read of the loss inside the timed region, every step
  roofline : the dominant kernel of the step, timed alone with CUDA events on its launching stream
  cpu_baseline : the CPU oracle (restatement of the JAX reference; JAX is not installable here) on the host cores
--impl reference times that CPU oracle alone with the same metric/unit (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #            model preset, side, default per-GPU batch, fwd FLOPs/sample
    'small64': ('small', 64, 8, 13_997_445_120),
    'small128': ('small', 128, 8, 27_804_045_312),
    'full128': ('full', 128, 2, 3_155_647_528_960),
    'full64': ('full', 64, 4, 881_256_824_832),
}
METRIC, UNIT = 'xunet_train_images_per_sec', 'images/s'


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def make_host_batches(n, B, S, seed):
    """n synthetic SRN-shaped batches as the data loader would hand them over (data_loader.py:102-113: x float32; z, noise
    float64; logsnr float64)."""
    from novel_view_synthesis_3d_b200.synthetic import synthetic_batch
    return [synthetic_batch(B, S, seed=seed + i) for i in range(n)]


def cpu_baseline(preset, S, B_sample, iters, threads=None):
    """CPU restatement of the JAX reference (oracle, fp32), one training step = fwd + bwd (autograd) + Adam on all leaves."""
    import torch
    from oracle import xunet_ref as R
    avail = len(os.sched_getaffinity(0))
    # the small model's CPU kernels stop scaling (and then regress) beyond a few dozen threads; cap and report it
    cores = threads or min(avail, 32)
    torch.set_num_threads(cores)
    cfg = R.SMALL if preset == 'small' else R.FULL
    params = R.init_params(cfg, S, seed=0, zero_init=True, dtype=torch.float32)
    flat = R.flatten(params)
    m = {k: torch.zeros_like(v) for k, v in flat.items()}
    v_ = {k: torch.zeros_like(v) for k, v in flat.items()}
    batch, noise = R.synthetic_batch(B_sample, S, seed=1234, dtype=torch.float32)
    cond = torch.ones(B_sample)
    rng = np.random.RandomState(0)
    mask_fn = lambda idx, shape: torch.from_numpy(rng.random_sample(shape) >= cfg.dropout)
    times = []
    t_begin = time.perf_counter()
    for it in range(iters + 1):
        if it > 1 and time.perf_counter() - t_begin > 25.0:     # bounded sample: ~10-30 s of CPU work
            break
        t0 = time.perf_counter()
        loss, grads, _ = R.loss_and_grads(R.nest(flat), batch, noise, cond, cfg, train=True, drop_mask_fn=mask_fn)
        for k in flat:
            flat[k], m[k], v_[k] = R.adam_update(flat[k], grads[k], m[k], v_[k], it + 1)
        dt = time.perf_counter() - t0
        if it > 0:          # first iteration = warm-up
            times.append(dt)
    med = float(np.median(times))
    return dict(value=B_sample / med, unit=UNIT, cores=cores, cores_available=avail, kind='port',
                sample=f'{len(times)} timed train steps (fwd+bwd+Adam, fp32 torch-CPU oracle) of batch {B_sample} at {S}x{S}, '
                       f'median {med * 1e3:.0f} ms/step; JAX itself is not installable offline')


def run_reference(args, preset, S, B):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    Bs = 2 if preset == 'small' else 1
    iters = max(1, min(args.steps, 5 if preset == 'small' else 1))
    cb = cpu_baseline(preset, S, Bs, iters)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': iters,
            'warmup': 1, 'ms_per_step': 1e3 * Bs / cb['value'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'model': preset, 'side': S, 'per_step_sample_batch': Bs, 'arm_per_gpu_batch': B},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=30, warm=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def dominant_kernel_roofline(P, model, B, S, peaks):
    """Times the step's heaviest kernels ALONE (CUDA events on the launching stream, >=3 warm-ups) through the operator-level
    C-ABI at the exact shapes and with the exact implementations (tcgen05) the step uses, and returns the roofline of the
    one with the largest share of the step.  Algorithmic FLOPs per launch are SURVEY 8(d)'s (2*MAC; backward := 2 x forward)."""
    import torch
    from novel_view_synthesis_3d_b200 import _lib
    lib = _lib.load()
    os.environ['XUNET_OP_CACHE_SHADOW'] = '1'     # time the conv kernel alone (the step converts weights once per forward)
    cfg = model.config
    tc = cfg.dtype == 'bf16'
    dt = _lib.DTYPE_BF16 if tc else _lib.DTYPE_F32
    impl = 1 if tc else 0
    tdt = torch.bfloat16 if tc else torch.float32
    st = torch.cuda.current_stream().cuda_stream
    N, dev, out = 2 * B, 'cuda', []
    nrb, L = cfg.num_res_blocks, len(cfg.ch_mult)

    def conv_case(H, Ci, Co, count):
        x = torch.randn(N, H, H, Ci, device=dev).to(tdt)
        w = torch.randn(9 * Ci * Co, device=dev) * 0.05
        b = torch.zeros(Co, device=dev)
        y = torch.empty(N, H, H, Co, device=dev, dtype=tdt)
        dw = torch.zeros(9 * Ci * Co, device=dev)
        flops = 2.0 * N * H * H * 9 * Ci * Co
        fn = lambda: lib.xunet_op_conv(dt, impl, x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), N, H, H, Ci, Co, 3, 1, 1, 1.0, st)
        assert fn() == 0, lib.xunet_last_error()
        out.append(dict(kernel=f'conv3x3 fwd {Ci}->{Co} @{H} ({"tcgen05" if tc else "simt"})', seconds=time_kernel(fn), flops=flops, count=count))
        fw = lambda: lib.xunet_op_conv_wgrad(dt, impl, x.data_ptr(), y.data_ptr(), dw.data_ptr(), b.data_ptr(), N, H, H, Ci, Co, 3, 1, 1, 1.0, st)
        assert fw() == 0, lib.xunet_last_error()
        out.append(dict(kernel=f'conv3x3 wgrad {Ci}->{Co} @{H} ({"tcgen05" if tc else "simt"})', seconds=time_kernel(fw), flops=flops, count=count))

    def attn_case(Lq, C, heads, count):
        qkv = torch.randn(N, Lq, 3 * C, device=dev).to(tdt)
        res = torch.randn(N, Lq, C, device=dev).to(tdt)
        o = torch.empty(N, Lq, C, device=dev, dtype=tdt)
        lse = torch.empty(N, heads, Lq, device=dev)
        dscr = torch.empty(N * Lq * (heads + C), device=dev)
        dqkv = torch.empty_like(qkv)
        flops = 4.0 * N * heads * Lq * Lq * (C // heads)
        fn = lambda: lib.xunet_op_attention(dt, impl, qkv.data_ptr(), res.data_ptr(), o.data_ptr(), lse.data_ptr(), N, Lq, C, heads, 0, st)
        assert fn() == 0, lib.xunet_last_error()
        out.append(dict(kernel=f'attention fwd L={Lq} hd={C // heads}', seconds=time_kernel(fn), flops=flops, count=count))
        fb = lambda: lib.xunet_op_attention_bwd(dt, impl, qkv.data_ptr(), res.data_ptr(), o.data_ptr(), res.data_ptr(), lse.data_ptr(),
                                                dscr.data_ptr(), dqkv.data_ptr(), N, Lq, C, heads, 0, st)
        assert fb() == 0, lib.xunet_last_error()
        # algorithmic backward work = 5 GEMMs (S, dP, dV, dK, dQ) = 2.5x the forward's two
        out.append(dict(kernel=f'attention bwd (prep + fused dK/dV/dQ + store) L={Lq} hd={C // heads}', seconds=time_kernel(fb), flops=2.5 * flops, count=count))

    feat = [cfg.ch * m for m in cfg.ch_mult]
    conv_case(S, feat[0], feat[0], 2 * nrb + 2)
    if L > 1:
        conv_case(S // 2, feat[1], feat[1], 2 * nrb + 4)
        if (S // 2) in cfg.attn_resolutions:
            attn_case((S // 2) ** 2, feat[1], cfg.attn_heads, 2 * (2 * nrb + 2))
    top = max(out, key=lambda r: r['seconds'] * r['count'])
    peak = peaks['tf_burst']
    achieved = top['flops'] / top['seconds'] / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')   # dram bytes/launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top['kernel'])
    return {'bound': 'tensor', 'kernel': top['kernel'], 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': achieved / peak, 'traffic': traffic, 'peak_source': peaks['source'] + ' (burst, kernel timed alone)',
            'per_launch_us': top['seconds'] * 1e6, 'algorithmic_flops_per_launch': top['flops'],
            'candidates': [{'kernel': r['kernel'], 'us': r['seconds'] * 1e6, 'tflops': r['flops'] / r['seconds'] / 1e12,
                            'launches_per_step': r['count']} for r in out]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='small64', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: workload preset)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--sampler-steps', type=int, default=256, help='extra: DDPM sampler views/s at N=1 (0 = skip)')
    args = ap.parse_args()
    preset, S, B0, fwd_flops = WORKLOADS[args.workload]
    B = args.batch or B0
    if args.impl == 'reference':
        run_reference(args, preset, S, B)
        return

    import torch
    import novel_view_synthesis_3d_b200 as P
    from novel_view_synthesis_3d_b200 import dist as xdist
    if args.warmup < 3:
        args.warmup = 3
    local = xdist.init_from_env('nccl')
    world, rank = xdist.world_size(), xdist.rank()
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)'
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()

    cfgk = dict(dtype=args.dtype)
    model = P.XUNet(**cfgk) if preset == 'small' else P.XUNet.from_config(P.XUNetConfig(**{**P.FULL_3DIM.__dict__, **cfgk}))
    state = P.create_train_state(0, 1, 1e-4, B, S, model=model)            # Flax-style init, rank-0 params broadcast
    step = P.TrainStep(state, use_graph=not args.no_graph)
    eng = step.eng
    n_host = 4
    host = make_host_batches(n_host, B, S, seed=1234 + 100 * rank)
    devb = [({k: torch.as_tensor(v, dtype=torch.float32).to(dev) for k, v in nb.items()},
             torch.as_tensor(nz, dtype=torch.float32).to(dev)) for nb, nz in host]
    rng = np.random.RandomState(rank)
    masks = [np.where(rng.random_sample(B) > 0.1, 1, 0).astype(np.float32) for _ in range(n_host)]
    dmasks = [torch.as_tensor(m).to(dev) for m in masks]

    def sync_all():
        xdist.barrier()
        torch.cuda.synchronize(dev)

    def timed(run_one, K):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            run_one(i)
        e1.record()
        sync_all()
        return xdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    # ---- value: inputs resident in HBM -----------------------------------------------------------------------
    def dev_step(i):
        b, nz = devb[i % n_host]
        step(b, nz, cond_mask=dmasks[i % n_host])

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(args.warmup):
        dev_step(i)
    t_dev = timed(dev_step, args.steps)
    # ---- e2e: host numpy inputs through the public API, loss read back every step -------------------------------
    # the loss of every step is read back device->host inside the timed region, asynchronously into pinned memory (a real
    # loop logs step i-1 while step i runs); the final synchronize of the timed region covers the last copy
    loss_ring = torch.zeros(args.steps + 8, dtype=torch.float32).pin_memory()

    def host_step(i):
        nb, nz = host[i % n_host]
        loss = step(nb, nz, cond_mask=masks[i % n_host])
        loss_ring[i % loss_ring.numel()].copy_(loss, non_blocking=True)

    for i in range(3):
        host_step(i)
    h2d = step.h2d_bytes
    t_e2e = timed(host_step, args.steps)
    losses = [float(v) for v in loss_ring[:args.steps]]
    clk = clocks.stop() if rank == 0 else None

    nf, nb_ = eng.count_kernels(state.params.flat)
    launches = (nf + nb_ + 1) * args.steps
    roof = dominant_kernel_roofline(P, model, B, S, peaks) if rank == 0 else None
    sampler = None
    if rank == 0 and world == 1 and args.sampler_steps > 0:
        # BASELINE.json's second metric: views/sec of the 256-step ancestral sampler with classifier-free guidance
        # (sampling.py:119-151; cond + uncond evaluated as ONE forward of a 2B batch, CUDA-graph replayed)
        sampler = {'steps': args.sampler_steps, 'guidance_w': 3.0, 'model': preset, 'side': S, 'runs': []}
        for Bs in (1, B):
            smp = P.Sampler(model, state.params, Bs, S, steps=args.sampler_steps, w=3.0, use_graph=True)
            sb = {k: v[:Bs] for k, v in host[0][0].items()}
            smp.sample(sb, seed=0)                      # warm-up + graph capture
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smp.sample(sb, seed=1)
            e1.record()
            torch.cuda.synchronize(dev)
            ts = e0.elapsed_time(e1) * 1e-3
            sampler['runs'].append({'views_in_flight': Bs, 'views_per_sec': Bs / ts, 'ms_per_sampler_step': ts / len(smp.sched) * 1e3})
            del smp
    cb = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cb = cpu_baseline(preset, S, 2 if preset == 'small' else 1, 4 if preset == 'small' else 1)
    if rank == 0:
        imgs = B * world * args.steps
        value = imgs / t_dev
        train_flops = 3.0 * fwd_flops * B
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.dtype == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'model': preset, 'side': S, 'per_gpu_batch': B, 'global_batch': B * world,
                       'parallelism': f'dp{world}', 'optimizer': 'adam lr1e-4', 'loss': 'frobenius (train.py:67)',
                       'cuda_graph': not args.no_graph,
                       'l2': 'no explicit flush: one step touches ~%.1f GB of activations+grads (>> 126 MB L2) and rotates over %d '
                             'different input batches' % (eng.ws_bytes / 1e9, n_host)},
            'e2e': {'value': imgs / t_e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4,
                    'ms_per_step': t_e2e / args.steps * 1e3},
            'gpu_launches': int(launches), 'kernels_per_step': {'forward': nf, 'backward': nb_, 'adam': 1},
            'step_tflops': train_flops / (t_dev / args.steps) / 1e12,
            'step_tensor_frac_of_sustained': train_flops / (t_dev / args.steps) / 1e12 / peaks['tf_sustained'],
            'roofline': roof, 'sampler': sampler, 'cpu_baseline': cb, 'clocks': clk, 'final_loss': losses[-1] if losses else None,
        }
        print(json.dumps(line))
    if xdist.is_dist():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
