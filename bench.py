#!/usr/bin/env python
"""bench.py -- anytime voxel reconstructions/sec on N B200s (BASELINE.json metric, config 2 workload).

A step = one pass of the anytime hot path over one batch: for each missing rate in {25, 50, 75 %}, 256 partially
received ModelNet latents (D = 64) are completed with K = 16 Philox prior samples, every completed latent is decoded to
a 64^3 occupancy grid, the K grids are averaged, thresholded (>= 0.5) and scored (TP/FP/FN) against synthetic targets:
768 objects = 12,288 decodes per step per GPU (weak scaling: every rank processes its own 768 objects; the only
collective is one all-reduce of the integer counts).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl a3d|reference]

`value`  : objects/s with inputs already resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`    : same metric through the host-buffer C-ABI call a3d_anytime_eval_host (numpy in, counts out; H2D + D2H
           inside the timed region).
`roofline`: the dominant kernel (128->64 transposed-conv implicit GEMM) against the measured bf16 tensor peak.
`oracle/` is used here in two ways only: as the timed CPU baseline (below), and as the seeded numpy generator of the
           synthetic inputs / random-init weights (no device compute goes through it).
`cpu_baseline` / `--impl reference`: the torch-CPU-fp32 oracle (the reference itself needs TensorFlow, which is not
           installable offline -- see DESIGN.md) on a bounded sample, on the host cores.
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
sys.path.insert(0, ROOT)

B_PER_RATE = 256
RATES = (0.25, 0.50, 0.75)
K = 16
D = 64
NCAT = 40
L4_MACS = 1_952_382_976          # exact MACs of the 128->64 layer per decode
FLOP_PER_DECODE = 6.663830528e9
WORKLOAD = (f'ModelNet VAE_dr anytime decode: {len(RATES)} missing rates (25/50/75%) x {B_PER_RATE} objects x K={K} '
            f'prior samples per GPU per step, D={D}, Keras-default random-init weights, synthetic ellipsoid targets')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='a3d', choices=['a3d', 'reference'])
    ap.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('bf16_tflops_sustained', 1365.9), d.get('hbm_gbs', 6549.1), 'measured'
    return 1400.0, 6650.0, 'fallback'


def synth(rank: int, seed: int = 1235):
    """Synthetic inputs of config 2: seeded PCG64, derived once on the host."""
    from oracle import anytime_ref as ar, decoder_ref as dr
    rng = np.random.Generator(np.random.PCG64(seed + 1000 * rank))
    mu = rng.standard_normal((NCAT, D)).astype(np.float32)
    batches = []
    for r in RATES:
        z = dr.round_bf16(rng.standard_normal((B_PER_RATE, D)).astype(np.float32))
        mask = ar.bernoulli_mask(rng, B_PER_RATE, D, r)
        batches.append((z, mask))
    tgt = ar.make_targets(rng, 8)
    bits8 = ar.pack_bits(tgt)
    bits = np.tile(bits8, (B_PER_RATE // 8, 1))
    return mu, batches, bits


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            p = [x.strip() for x in ln.split(',')]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_rate(n_obj: int, reps: int = 1):
    """Oracle (torch CPU fp32, all host threads) on a bounded sample of the same workload; objects/s."""
    import torch
    from oracle import anytime_ref as ar, decoder_ref as dr
    MODELNET_DECODER = dr.MODELNET_DECODER
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.Generator(np.random.PCG64(99))
    ws = dr.keras_default_weights(MODELNET_DECODER, 3)
    mu = rng.standard_normal((NCAT, D)).astype(np.float32)
    z = rng.standard_normal((n_obj, D)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, n_obj, D, 0.5)
    tgt = ar.make_targets(rng, n_obj)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        zc, _ = ar.impute(z, mask, mu, K, seed=1, fill='prior_sample')
        ar.anytime_eval(MODELNET_DECODER, ws, zc, tgt, 0.5, batch=16)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_obj / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow (the reference's only backend) is
    not installed, so this is the oracle port, all host threads, on a bounded sample per step."""
    if rank != 0:
        return
    n_obj = 8
    for _ in range(max(args.warmup, 0)):
        cpu_reference_rate(n_obj)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(n_obj)
    dt = time.perf_counter() - t0
    val = n_obj * args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'anytime voxel reconstructions/sec', 'value': val, 'unit': 'objects/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'reference_sample': f'bounded sample of {n_obj} objects ({n_obj * K} decodes) '
                                                              f'of that workload per step on the host CPU'},
        'cpu_baseline': {'value': val, 'unit': 'objects/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': f'{n_obj} objects x K={K} per step, torch CPU fp32 oracle (reference needs '
                                   f'TensorFlow, unavailable offline)'},
        'e2e': {'value': val, 'unit': 'objects/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def config3_aux(a3d, dev, dtype, B=128, size=256, D16=16, steps=5):
    """Secondary measurement (not the headline metric): BASELINE config 3 -- Pascal3D image encoder + voxel decoder on
    synthetic RGB crops, batch 128, one GPU: Darknet19 + head2D -> mean / clipped logvar -> sampling -> decoder -> counts
    (K = 1, full latent), random-init weights.  `value` with the images resident in HBM, `e2e` from pinned host images
    (100 MB H2D per step inside the timed region) to counts on the host."""
    import torch
    from oracle import anytime_ref as ar, decoder_ref as dr, encoder2d_ref as er
    layers = er.layer_list()
    enc = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(size, size), max_batch=B, operand_dtype=dtype)
    enc.set_weights(er.keras_default_weights(layers, 3, seed=77))
    dec = a3d.decoder3D(a3d.presets.PASCAL_DECODER, max_chunk=B, operand_dtype=dtype)
    dec.set_weights(dr.keras_default_weights(a3d.presets.PASCAL_DECODER, 78))
    rng = np.random.Generator(np.random.PCG64(1237))
    x_host = rng.uniform(0, 1, (B, size, size, 3)).astype(np.float32)
    bits = torch.from_numpy(np.tile(ar.pack_bits(ar.make_targets(rng, 8)), (B // 8, 1))).to(dev)
    x = torch.from_numpy(x_host).to(dev)
    ones = torch.ones((B, D16), device=dev)

    def step(inp, i):
        _, _, z = enc.encode(inp, D16, seed=100 + i)
        return a3d.anytime_eval(dec, z, ones, None, bits, K=1, seed=i, fill='normal')['counts']

    def timed(fn):
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            c = fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, c

    ms_enc, _ = timed(lambda i: enc(x))
    ms_all, c = timed(lambda i: step(x, i))
    # e2e: images start in pinned host memory (bench contract); a second handle with max_batch = 32 lets the host path
    # (a3d_enc2d_forward_host) overlap the H2D copy of chunk i+1 with the forward of chunk i
    x_pin = torch.from_numpy(x_host).pin_memory()
    enc_h = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(size, size), max_batch=32, operand_dtype=dtype)
    enc_h.set_weights(enc.get_weights())

    def step_host(i):
        _, _, z = enc_h.encode(x_pin, D16, seed=100 + i)
        return a3d.anytime_eval(dec, z, ones, None, bits, K=1, seed=i, fill='normal')['counts']
    step_host(0).cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        ce = step_host(i).cpu()
    e2e_s = (time.perf_counter() - t0) / steps
    # the same step from the loader's raw bytes (uint8 NHWC, `image / 255.` of pascal3D.py:242 applied on the device):
    # the H2D copy carries 25 MB instead of 100 MB per 128 crops (tests/tools/sweep_host_images.py: chunk-size sweep)
    x_u8 = torch.from_numpy(np.clip(np.rint(x_host * 255.0), 0, 255).astype(np.uint8)).pin_memory()

    def step_host_u8(i):   # one chunk of 128 on the resident handle: the 25 MB copy is shorter than a chunk's forward
        _, _, z = enc.encode(x_u8, D16, seed=100 + i)
        return a3d.anytime_eval(dec, z, ones, None, bits, K=1, seed=i, fill='normal')['counts']
    step_host_u8(0).cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step_host_u8(i).cpu()
    e2e_u8_s = (time.perf_counter() - t0) / steps
    enc_h.close()
    alg, dense = er.encoder_macs(layers, size, size, 3)
    out = {'workload': f'Pascal3D multi-modal image encoder (Darknet19 + head2D, {size}x{size} RGB) + voxel decoder '
                       f'(D={D16}), batch {B}, K=1, synthetic crops, Keras-default random-init weights',
           'value': B / (ms_all * 1e-3), 'unit': 'objects/s', 'ms_per_step': ms_all,
           'encoder_ms_per_step': ms_enc, 'encoder_images_per_s': B / (ms_enc * 1e-3),
           'encoder_tflops_algorithmic': 2.0 * alg * B / (ms_enc * 1e-3) / 1e12,
           'encoder_macs_per_image': {'algorithmic': alg, 'dense': dense},
           'e2e': {'value': B / e2e_s, 'unit': 'objects/s', 'h2d_bytes_per_step': int(x_host.nbytes), 'd2h_bytes_per_step': B * 24},
           'e2e_uint8_images': {'value': B / e2e_u8_s, 'unit': 'objects/s', 'h2d_bytes_per_step': int(x_u8.numel()),
                                'd2h_bytes_per_step': B * 24},
           'gpu_launches_per_step': None, 'counts_tp_fp_fn': [int(v) for v in c.sum(0).tolist()]}
    l0 = enc.launch_count + dec.launch_count
    step(x, 0)
    out['gpu_launches_per_step'] = int(enc.launch_count + dec.launch_count - l0)
    enc.close()
    dec.close()
    return out


def _stdout_to_stderr():
    """Route fd 1 to stderr until the result line is printed: libraries (NCCL prints its version banner on stdout at
    communicator creation) must not add lines to the one-JSON-line contract.  Returns the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _restore_stdout(saved):
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


def main():
    args = parse()
    saved_stdout = _stdout_to_stderr()
    try:
        _main(args, saved_stdout)
    finally:
        sys.stdout.flush()


def _main(args, saved_stdout):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        _restore_stdout(saved_stdout)
        run_reference(args, rank, world)
        return
    import torch
    import a3d
    from a3d.presets import MODELNET_DECODER
    from oracle import decoder_ref as dr

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=B_PER_RATE * K, operand_dtype=args.dtype, device=local_rank)
    dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1234))
    mu, batches, bits = synth(rank)
    mu_d = torch.from_numpy(mu).to(dev)
    bits_d = torch.from_numpy(bits).to(dev)
    dev_batches = [(torch.from_numpy(z).to(dev), torch.from_numpy(m).to(dev)) for z, m in batches]
    total = torch.zeros(3, dtype=torch.int64, device=dev)

    def step(i):
        acc = torch.zeros(3, dtype=torch.int64, device=dev)
        for j, (z, m) in enumerate(dev_batches):
            r = a3d.anytime_eval(dec, z, m, mu_d, bits_d, K=K, seed=1000 + i, fill='prior_sample',
                                 obj_offset=(rank * len(RATES) + j) * B_PER_RATE)
            acc += r['counts'].sum(0)
        if dist is not None:
            dist.all_reduce(acc)       # the path's only collective: [TP, FP, FN] int64 over NVLink
        return acc

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = dec.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        total = step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = dec.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    obj_per_step = B_PER_RATE * len(RATES) * world
    value = obj_per_step * args.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call
    host_bits = bits
    def e2e_step(i):
        out = np.zeros(3, np.int64)
        for j, (z, m) in enumerate(batches):
            c = a3d.anytime_eval_host(dec, z, m, mu, host_bits, K=K, seed=1000 + i, fill='prior_sample',
                                      obj_offset=(rank * len(RATES) + j) * B_PER_RATE)
            out += c.sum(0)
        return out
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_counts = e2e_step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = obj_per_step * args.steps / e2e_s
    h2d = len(RATES) * (2 * B_PER_RATE * D * 4 + NCAT * D * 4 + B_PER_RATE * 32768)
    d2h = len(RATES) * B_PER_RATE * 24

    # ---- per-stage device times (CUDA events around each kernel, separate pass so the sync does not pollute `value`)
    dec.set_profiling(True)
    stage = {}
    nprof = 0
    for i in range(2):
        for j, (z, m) in enumerate(dev_batches):
            a3d.anytime_eval(dec, z, m, mu_d, bits_d, K=K, seed=7 + i, fill='prior_sample')
            for k_, v in dec.stage_times_ms().items():
                stage[k_] = stage.get(k_, 0.0) + v
            nprof += 1
    dec.set_profiling(False)
    stage = {k_: v / nprof for k_, v in stage.items()}
    decodes_per_launch = B_PER_RATE * K
    tpeak, hpeak, src = peaks()
    l4_tflops = 2.0 * L4_MACS * decodes_per_launch / (stage['l4'] * 1e-3) / 1e12
    # algorithmic bytes of the fused tail per launch (SURVEY 8d variant B): K * 4 MiB of bf16/fp16 activations + target
    # bits + counts per object
    tail_bytes = B_PER_RATE * (K * 2_097_152 * 2 + 262_144 // 8 + 24)
    tail_gbs = tail_bytes / (stage['tail'] * 1e-3) / 1e9
    roofline = {'bound': 'tensor', 'kernel': 'convt_l4_ws_kernel (128->64 ConvT, 2-CTA weight-stationary)',
                'achieved': l4_tflops, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': l4_tflops / tpeak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this size, from the
                # ncu --set full capture committed as profiles/r01_l4_ws2cta_ncu_full.txt (algorithmic: 21.5e9)
                'traffic': 30.17e9, 'traffic_unit': 'bytes per launch (ncu, profiles/r01_l4_ws2cta_ncu_full.txt)',
                'peak_source': f'{src} sustained bf16', 'stage_ms': stage,
                'decoder_tflops_all_stages': FLOP_PER_DECODE * decodes_per_launch / (sum(stage.values()) * 1e-3) / 1e12,
                'fused_tail': {'bound': 'hbm', 'kernel': 'tail_pair_kernel<MODE_HCOL> (final ConvT + sigmoid + K-mean + threshold + counts)',
                               'achieved': tail_gbs, 'peak': hpeak, 'unit': 'GB/s', 'frac': tail_gbs / hpeak,
                               'traffic': 17.20e9,   # dram bytes of one launch, profiles/r01_tail_hcol_final_ncu_full.txt
                               'algorithmic_bytes_per_launch': tail_bytes}}

    aux = None
    if rank == 0 and world == 1:   # secondary measurement and CPU baseline: single-process runs only
        try:
            aux = config3_aux(a3d, dev, args.dtype)
        except Exception as e:   # the secondary measurement must never take the headline line down
            aux = {'error': repr(e)[:300]}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rate, secs = cpu_reference_rate(40)     # ~10 s of CPU work on the 16-thread host (bounded sample)
            cpu = {'value': rate, 'unit': 'objects/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': f'40 objects x K={K} = 640 decodes of the same workload in {secs:.1f} s, torch CPU fp32 '
                             f'oracle, {os.cpu_count()} threads'}
        line = {
            'metric': 'anytime voxel reconstructions/sec', 'value': value, 'unit': 'objects/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype + ' operands, f32 accumulate',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'decodes_per_step_per_gpu': B_PER_RATE * K * len(RATES), 'parallelism': f'objects sharded x{world}',
                       'l2_policy': 'activation working set per step (65 GB/GPU) is streamed through HBM, >> 126 MB L2'},
            'decodes_per_s': value * K,
            'e2e': {'value': e2e_val, 'unit': 'objects/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
            'aux_config3_images_to_voxels': aux,
            'counts_tp_fp_fn': [int(v) for v in total.tolist()], 'e2e_counts': [int(v) for v in e2e_counts.tolist()],
        }
        _restore_stdout(saved_stdout)
        print(json.dumps(line), flush=True)
        saved_stdout = _stdout_to_stderr()   # anything printed during teardown stays off stdout
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
