#!/usr/bin/env python
"""bench.py -- anytime voxel reconstructions/sec on N B200s (BASELINE.json metric, config 2 workload).

A step = one pass of the anytime hot path over one batch: for each missing rate in {25, 50, 75 %}, 256 partially
received ModelNet latents (D = 64) are completed with K = 16 Philox prior samples, every completed latent is decoded to
a 64^3 occupancy grid, the K grids are averaged, thresholded (>= 0.5) and scored (TP/FP/FN) against synthetic targets:
768 objects = 12,288 decodes per step per GPU (weak scaling: every rank processes its own 768 objects; counts are
accumulated on the device and all-reduced ONCE, after the last step, inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl a3d|reference]

`value`      : objects/s with inputs already resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`        : same metric through the host-buffer C-ABI call a3d_anytime_eval_host (numpy in, counts out; H2D + D2H
               inside the timed region).
`roofline`   : the dominant kernel (128->64 transposed-conv implicit GEMM) against the measured bf16 tensor peak, the
               fused tail against the measured HBM peak; `traffic` comes from profiles/roofline_traffic.json (written by
               tools/ncu_summary.py from the committed ncu captures), never from a literal.
`parity_spot`: two objects of the timed step decoded through the CPU oracle on the box (max |dp|, flipped voxels,
               count differences), for the Keras-default weights of the headline run and for a trained-like weight set.
`tf_available`: importlib.util.find_spec('tensorflow') on this box (the reference's own framework).
aux keys     : config 1 (B = 32 / K = 32 single calls), decoder(z) numpy->numpy grid return, config 3 (images), the
               same step on trained-like weights, configs 4 and 5 (sharded over the ranks when N > 1).
`oracle/` is used here in three ways only: the timed CPU baseline, the on-box parity spot check, and as the seeded numpy
generator of the synthetic inputs / random-init weights (no device compute goes through it).
`cpu_baseline` / `--impl reference`: the torch-CPU-fp32 oracle (the reference itself needs TensorFlow, which is not
installable offline -- see DESIGN.md) on a bounded sample (40 objects x K = 16; the reference arm shrinks its per-step
sample, never below 8 objects, so that K + W steps end within ~3 minutes), compute only, all host threads.
"""
from __future__ import annotations

import argparse
import importlib.util
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
CPU_SAMPLE_OBJECTS = 40
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
    ap.add_argument('--no-aux', action='store_true', help='headline measurement only (profiling runs)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('bf16_tflops_sustained', 1365.9), d.get('hbm_gbs', 6549.1), 'measured'
    return 1400.0, 6650.0, 'fallback'


def ncu_traffic():
    """DRAM bytes per launch of the two roofline kernels, from the committed ncu captures (tools/ncu_summary.py)."""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def tf_available() -> bool:
    return importlib.util.find_spec('tensorflow') is not None


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
    def __init__(self, gpu_id: str):
        self.proc = None
        self.lines = []
        self.gpu = gpu_id            # UUID ("GPU-...") or index as nvidia-smi -i takes it

    def start(self):
        # started BEFORE the warm-up steps (nvidia-smi needs ~1 s to come up); stop(t0, t1) keeps only the samples whose
        # own timestamp falls inside the timed region
        q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
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

    def stop(self, t0: float, t1: float):
        """t0, t1: time.time() at the start / end of the timed region."""
        import datetime
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons, n_all = [], [], [], set(), 0
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            p = [x.strip() for x in ln.split(',')]
            if len(p) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                vals = (float(p[1]), float(p[2]), float(p[3]))
            except ValueError:
                continue
            n_all += 1
            if not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            sm.append(vals[0]); mx.append(vals[1]); pw.append(vals[2])
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_median': float(np.median(pw)) if pw else None, 'reasons': sorted(reasons),
                'samples_in_timed_region': len(sm), 'samples': n_all, 'gpu': self.gpu}


# ---------------------------------------------------------------------------------------------------- CPU arm
class CpuSample:
    """Bounded sample of the bench workload for the CPU arm: weights, latents, masks and targets are generated ONCE
    (outside every timed region); `run()` times the compute only (imputation + K-sample decode + mean + counts)."""

    def __init__(self, n_obj: int = CPU_SAMPLE_OBJECTS):
        import torch
        from oracle import anytime_ref as ar, decoder_ref as dr
        torch.set_num_threads(os.cpu_count() or 1)
        self.ar, self.st = ar, dr.MODELNET_DECODER
        rng = np.random.Generator(np.random.PCG64(99))
        self.n_obj = n_obj
        self.ws = dr.keras_default_weights(self.st, 3)
        self.mu = rng.standard_normal((NCAT, D)).astype(np.float32)
        self.z = rng.standard_normal((n_obj, D)).astype(np.float32)
        self.mask = ar.bernoulli_mask(rng, n_obj, D, 0.5)
        self.tgt = ar.make_targets(rng, n_obj)

    def run(self) -> float:
        t0 = time.perf_counter()
        zc, _ = self.ar.impute(self.z, self.mask, self.mu, K, seed=1, fill='prior_sample')
        self.ar.anytime_eval(self.st, self.ws, zc, self.tgt, 0.5, batch=16)
        return time.perf_counter() - t0

    def describe(self, secs: float) -> str:
        return (f'{self.n_obj} objects x K={K} = {self.n_obj * K} decodes of the same workload in {secs:.1f} s, compute '
                f'only (inputs and weights generated beforehand), torch CPU fp32 oracle, {os.cpu_count()} threads')


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow (the reference's only backend) is
    not installed, so this is the oracle port, all host threads; every step is the same bounded sample the
    `cpu_baseline` leg of the main arm times."""
    if rank != 0:
        return
    # the whole --steps K --warmup W run must end within a few minutes: a calibration pass on 8 objects gives the host's
    # seconds per object, the per-step sample is then the cpu_baseline leg's 40 objects, or fewer when K + W steps of 40
    # would take longer than ~3 minutes (never fewer than 8)
    cal = CpuSample(8)
    sec_per_obj = cal.run() / 8.0
    budget_s = 180.0
    n_obj = int(budget_s / (max(args.steps + max(args.warmup, 0), 1) * sec_per_obj))
    n_obj = max(8, min(CPU_SAMPLE_OBJECTS, n_obj))
    s = CpuSample(n_obj)
    for _ in range(max(args.warmup, 0)):
        s.run()
    secs = [s.run() for _ in range(args.steps)]
    dt = float(np.sum(secs))
    val = s.n_obj * args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'anytime voxel reconstructions/sec', 'value': val, 'unit': 'objects/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'reference_sample': f'bounded sample of {s.n_obj} objects ({s.n_obj * K} decodes) '
                                                              f'of that workload per step on the host CPU'},
        'cpu_baseline': {'value': val, 'unit': 'objects/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': s.describe(dt / args.steps) + ' per step (reference needs TensorFlow, unavailable offline)'},
        'e2e': {'value': val, 'unit': 'objects/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'tf_available': tf_available(),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- helpers
def cuda_timed(torch, fn, reps, warm=2):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for i in range(reps):
        out = fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def parity_spot(a3d, dec, st, ws, z, mask, mu, bits_rows, seed, obj_offset, pick, full_counts=None):
    """Decode `pick` objects of a timed batch through the CPU oracle on this box and compare: the GPU re-runs just those
    objects with their global object ids (same Philox draws), the oracle decodes the same completed latents."""
    from oracle import anytime_ref as ar
    out = {'objects': [int(p) for p in pick], 'K': K, 'max_dp': 0.0, 'flips': 0, 'voxels': 0, 'dcounts': 0}
    got_c, ref_c = [], []
    for p in pick:
        r = a3d.anytime_eval(dec, z[p:p + 1], mask[p:p + 1], mu, bits_rows[p:p + 1], K=K, seed=seed, fill='prior_sample',
                             obj_offset=obj_offset + p, return_grid=True)
        zc = r['z_completed'].cpu().numpy()
        tgt = np.unpackbits(bits_rows[p:p + 1], axis=1, bitorder='little').reshape(1, 64, 64, 64, 1).astype(np.float32)
        ref_mp, ref_cnt = ar.anytime_eval(st, ws, zc, tgt, 0.5, batch=16)
        mp = r['mean_prob'].cpu().numpy()
        cnt = r['counts'].cpu().numpy()
        out['max_dp'] = max(out['max_dp'], float(np.abs(mp - ref_mp).max()))
        out['flips'] += int(((mp >= 0.5) != (ref_mp >= 0.5)).sum())
        out['voxels'] += int(mp.size)
        out['dcounts'] += int(np.abs(cnt - ref_cnt).sum())
        got_c.append(cnt[0].astype(np.float64))
        ref_c.append(np.asarray(ref_cnt)[0].astype(np.float64))
        if full_counts is not None:      # the counts-only call of the timed batch: same integers up to threshold ties
            out['dcounts_vs_timed_batch'] = out.get('dcounts_vs_timed_batch', 0) + int(np.abs(full_counts[p] - cnt[0]).sum())
    out['flips_pct'] = 100.0 * out['flips'] / max(out['voxels'], 1)
    # IoU = TP / (TP + FP + FN): mean over the spot objects and the global ratio, GPU minus oracle
    g, r = np.array(got_c), np.array(ref_c)
    iou = lambda c: c[:, 0] / np.maximum(c.sum(1), 1.0)
    out['iou_delta_mean'] = float(iou(g).mean() - iou(r).mean())
    out['iou_delta_global'] = float(g[:, 0].sum() / max(g.sum(), 1.0) - r[:, 0].sum() / max(r.sum(), 1.0))
    out['ok'] = bool(out['max_dp'] < 1e-2 and out['flips_pct'] < 0.1)
    return out


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

    ms_enc, _ = cuda_timed(torch, lambda i: enc(x), steps)
    ms_all, c = cuda_timed(torch, lambda i: step(x, i), steps)
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
        step_host(i).cpu()
    e2e_s = (time.perf_counter() - t0) / steps
    x_u8 = torch.from_numpy(np.clip(np.rint(x_host * 255.0), 0, 255).astype(np.uint8)).pin_memory()

    def step_host_u8(i):
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
           'e2e': {'value': B / e2e_s, 'unit': 'objects/s', 'h2d_bytes_per_step': int(x_host.nbytes), 'd2h_bytes_per_step': B * 24},
           'e2e_uint8_images': {'value': B / e2e_u8_s, 'unit': 'objects/s', 'h2d_bytes_per_step': int(x_u8.numel()),
                                'd2h_bytes_per_step': B * 24},
           'counts_tp_fp_fn': [int(v) for v in c.sum(0).tolist()]}
    enc.close()
    dec.close()
    return out


def config1_aux(a3d, torch, dev, ws, dtype):
    """BASELINE config 1 and the reference's own call shapes: B = 32 objects x K = 1 full latents (test_modelnet_VAE),
    ONE object x K = 32 samples (nolbo_test.py:167-177) and the plain 32-latent decoder(z) call, each as a single call:
    device-resident (CUDA events, eager launches and a CUDA-graph replay) and end to end from numpy."""
    from oracle import anytime_ref as ar
    dec = a3d.decoder3D(a3d.presets.MODELNET_DECODER, max_chunk=96, operand_dtype=dtype)
    dec.set_weights(ws)
    rng = np.random.Generator(np.random.PCG64(4321))
    out = {}
    for name, B_, K_ in (('B32_K1', 32, 1), ('B1_K32', 1, 32), ('B72_K1', 72, 1)):
        z = rng.standard_normal((B_, D)).astype(np.float32)
        mask = np.ones((B_, D), np.float32) if K_ == 1 else ar.bernoulli_mask(rng, B_, D, 0.5)
        mu = rng.standard_normal((NCAT, D)).astype(np.float32)
        bits = ar.pack_bits(ar.make_targets(rng, B_))
        zd, md, mud, bd = (torch.from_numpy(a).to(dev) for a in (z, mask, mu, bits))
        ms, _ = cuda_timed(torch, lambda i: a3d.anytime_eval(dec, zd, md, mud, bd, K=K_, seed=i)['counts'], 50, warm=5)
        a3d.anytime_eval_host(dec, z, mask, mu, bits, K=K_, seed=0)
        t0 = time.perf_counter()
        for i in range(50):
            a3d.anytime_eval_host(dec, z, mask, mu, bits, K=K_, seed=i)
        e2e_s = (time.perf_counter() - t0) / 50
        n_dec = B_ * K_
        out[name] = {'objects': B_, 'K': K_, 'us_per_call_device': 1e3 * ms, 'objects_per_s': B_ / (ms * 1e-3),
                     'tflops': n_dec * FLOP_PER_DECODE / (ms * 1e-3) / 1e12,
                     'e2e': {'us_per_call': 1e6 * e2e_s, 'objects_per_s': B_ / e2e_s,
                             'h2d_bytes': int(z.nbytes + mask.nbytes + mu.nbytes + bits.nbytes), 'd2h_bytes': B_ * 24}}
    # decoder(z) with 32 latents, device in / device out (the fp32 grid is written to HBM): eager and graph replay
    z32 = torch.randn(32, D, device=dev)
    ms_eager, _ = cuda_timed(torch, lambda i: dec(z32), 50, warm=5)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dec(z32)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        dec(z32)
    ms_graph, _ = cuda_timed(torch, lambda i: g.replay(), 50, warm=5)
    out['decoder_call_32'] = {'us_eager': 1e3 * ms_eager, 'us_graph_replay': 1e3 * ms_graph,
                              'tflops_graph': 32 * FLOP_PER_DECODE / (ms_graph * 1e-3) / 1e12}
    del g
    dec.close()
    return out


def decode_grid_aux(a3d, torch, dev, ws, dtype):
    """decoder(z) the way the reference's scripts use it (test_modelnet_VAE_dr.py:128-130): numpy latents in, the whole
    occupancy grid back on the host, through a3d_decode_host with a page-locked result buffer; against the pinned
    device->host copy rate measured on this box."""
    src = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
    dst = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    d2h = 4 * (1 << 28) / (time.perf_counter() - t0)
    del src, dst
    dec = a3d.decoder3D(a3d.presets.MODELNET_DECODER, max_chunk=256, operand_dtype=dtype)
    dec.set_weights(ws)
    out = {'pinned_d2h_GBps': d2h / 1e9, 'cases': {}}
    for B_ in (72, 4096):
        z = np.random.default_rng(B_).standard_normal((B_, D)).astype(np.float32)
        for dt, npdt, per in (('f32', np.float32, 262144 * 4), ('f16', np.float16, 262144 * 2), ('bits', np.uint8, 32768)):
            res = a3d.pinned_empty((B_, 32768) if dt == 'bits' else (B_, 64, 64, 64, 1), npdt)
            reps = 2 if B_ > 1000 else 20
            dec(z, out=res, out_dtype=dt)
            t0 = time.perf_counter()
            for _ in range(reps):
                dec(z, out=res, out_dtype=dt)
            secs = (time.perf_counter() - t0) / reps
            out['cases'][f'B{B_}_{dt}'] = {'ms_per_call': 1e3 * secs, 'decodes_per_s': B_ / secs,
                                           'h2d_bytes': int(z.nbytes), 'd2h_bytes': int(B_ * per),
                                           'frac_of_pcie_bound': (B_ * per / d2h) / secs}
            del res
    dec.close()
    return out


def config4_aux(a3d, torch, dist, dev, ws, rank, world, dtype, n_objects=1024):
    """BASELINE config 4: every latent prefix length 1..64 of every object (K = 1 prior-sample fill), objects sharded
    over the ranks (shard_range), ONE all-reduce of the [64, 3] int64 per-length counts; strong scaling."""
    from oracle import anytime_ref as ar, decoder_ref as dr
    lo, hi = a3d.shard_range(n_objects, rank, world)
    nb = hi - lo
    dec = a3d.decoder3D(a3d.presets.MODELNET_DECODER, max_chunk=4096, operand_dtype=dtype, device=dev.index)
    dec.set_weights(ws)
    rng = np.random.Generator(np.random.PCG64(1238))
    z_all = dr.round_bf16(rng.standard_normal((n_objects, D)).astype(np.float32))
    mu = rng.standard_normal((NCAT, D)).astype(np.float32)
    tgt8 = ar.make_targets(rng, 8)
    z = np.repeat(z_all[lo:hi], 64, axis=0)                                  # (object, prefix) pairs, prefix minor
    mask = np.tile(ar.prefix_mask(64, D, np.arange(1, 65)), (nb, 1))
    bits = torch.from_numpy(ar.pack_bits(tgt8)[(np.arange(lo, hi) % 8).repeat(64)]).to(dev)
    zd, md, mud = (torch.from_numpy(a).to(dev) for a in (z, mask, mu))

    def run(i):
        r = a3d.anytime_eval(dec, zd, md, mud, bits, K=1, seed=11, obj_offset=lo * 64)
        per_len = r['counts'].view(nb, 64, 3).sum(0)
        a3d.allreduce_counts(per_len)
        return per_len
    ms, per_len = cuda_timed(torch, run, 2, warm=1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    c = per_len.cpu().numpy().astype(np.float64)
    iou = c[:, 0] / np.maximum(c.sum(1), 1)
    dec.close()
    return {'workload': f'anytime arrival sweep: {n_objects} objects x 64 prefix lengths, K=1, sharded over {world} GPU(s)',
            'scaling': 'strong', 'ms': ms, 'value': n_objects / (ms * 1e-3), 'unit': 'objects/s (64 prefix decodes each)',
            'decodes_per_s': n_objects * 64 / (ms * 1e-3), 'collective': 'one all-reduce of [64,3] int64',
            'iou_at_prefix_1_16_32_48_64': [float(iou[i]) for i in (0, 15, 31, 47, 63)],
            'counts_checksum': int(per_len.sum().item())}


def config5_aux(a3d, torch, dist, dev, ws, rank, world, dtype, per_rank=8192):
    """BASELINE config 5: large batch x K = 32; 8192 objects per rank (65,536 objects = 2,097,152 decodes at N = 8),
    device-timed max over ranks, ONE all-reduce of the [3] int64 counts."""
    from oracle import anytime_ref as ar, decoder_ref as dr
    dec = a3d.decoder3D(a3d.presets.MODELNET_DECODER, max_chunk=4096, operand_dtype=dtype, device=dev.index)
    dec.set_weights(ws)
    rng = np.random.Generator(np.random.PCG64(1239 + rank))
    z = dr.round_bf16(rng.standard_normal((per_rank, D)).astype(np.float32))
    mask = ar.bernoulli_mask(rng, per_rank, D, 0.5)
    mu = np.random.Generator(np.random.PCG64(77)).standard_normal((NCAT, D)).astype(np.float32)
    tgt8 = ar.make_targets(np.random.Generator(np.random.PCG64(78)), 8)
    bits = torch.from_numpy(np.tile(ar.pack_bits(tgt8), (per_rank // 8, 1))).to(dev)
    zd, md, mud = (torch.from_numpy(a).to(dev) for a in (z, mask, mu))
    a3d.anytime_eval(dec, zd[:128], md[:128], mud, bits[:128], K=32, seed=13)       # warm-up on one chunk

    def run(i):
        r = a3d.anytime_eval(dec, zd, md, mud, bits, K=32, seed=13, obj_offset=rank * per_rank)
        tot = r['counts'].sum(0)
        a3d.allreduce_counts(tot)
        return tot
    ms, tot = cuda_timed(torch, run, 1, warm=0)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    dec.close()
    n = per_rank * world
    return {'workload': f'large batch: {n} objects x K=32 = {n * 32} decodes over {world} GPU(s) (8192 objects per rank; '
                        f'N=8 is BASELINE config 5), p_missing 0.5, NCCL count reduce',
            'scaling': 'weak', 'ms': ms, 'value': n / (ms * 1e-3), 'unit': 'objects/s', 'decodes_per_s': n * 32 / (ms * 1e-3),
            'frac_of_sustained_bf16_peak_per_gpu': n * 32 / (ms * 1e-3) * FLOP_PER_DECODE / 1e12 / peaks()[0] / world,
            'counts_tp_fp_fn': [int(v) for v in tot.tolist()]}


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


def _guard(fn, *a, **kw):
    """Secondary measurements must never take the headline line down."""
    try:
        return fn(*a, **kw)
    except Exception as e:   # noqa: BLE001
        return {'error': repr(e)[:300]}


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
    ws_default = dr.keras_default_weights(MODELNET_DECODER, 1234)
    dec.set_weights(ws_default)
    mu, batches, bits = synth(rank)
    mu_d = torch.from_numpy(mu).to(dev)
    bits_d = torch.from_numpy(bits).to(dev)
    dev_batches = [(torch.from_numpy(z).to(dev), torch.from_numpy(m).to(dev)) for z, m in batches]
    base_off = rank * len(RATES) * B_PER_RATE

    def step(i, acc, keep=None):
        for j, (z, m) in enumerate(dev_batches):
            r = a3d.anytime_eval(dec, z, m, mu_d, bits_d, K=K, seed=1000 + i, fill='prior_sample',
                                 obj_offset=base_off + j * B_PER_RATE)
            acc += r['counts'].sum(0)
            if keep is not None:
                keep.append(r['counts'])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    wall = [0.0, 0.0]      # wall-clock bounds of the last timed region (for the clock sampler)

    def timed_steps(n_steps, seed0=0, keep=None):
        """n_steps steps, counts accumulated on the device, ONE all-reduce at the end; CUDA events, max over ranks."""
        acc = torch.zeros(3, dtype=torch.int64, device=dev)
        barrier()
        wall[0] = time.time()          # after the barrier: the first one of a run creates the NCCL communicator (idle GPU)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            step(seed0 + i, acc, keep if i == n_steps - 1 else None)
        if dist is not None:
            dist.all_reduce(acc)       # the path's only collective: [TP, FP, FN] int64 over NVLink, once per run
        e1.record()
        torch.cuda.synchronize()
        wall[1] = time.time()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc

    # nvidia-smi numbers the physical GPUs of the box; CUDA's index is relative to CUDA_VISIBLE_DEVICES: address by UUID
    try:
        gpu_id = 'GPU-' + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:   # noqa: BLE001
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id)
    if rank == 0:
        sampler.start()
    warm = torch.zeros(3, dtype=torch.int64, device=dev)
    for i in range(args.warmup):
        step(i, warm)
    l0 = dec.launch_count
    last_counts = []
    ms, total = timed_steps(args.steps, keep=last_counts)
    t_wall0, t_wall1 = wall
    launches = dec.launch_count - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    obj_per_step = B_PER_RATE * len(RATES) * world
    value = obj_per_step * args.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call
    def e2e_step(i):
        out = np.zeros(3, np.int64)
        for j, (z, m) in enumerate(batches):
            c = a3d.anytime_eval_host(dec, z, m, mu, bits, K=K, seed=1000 + i, fill='prior_sample',
                                      obj_offset=base_off + j * B_PER_RATE)
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
    traffic = ncu_traffic()
    l4_tflops = 2.0 * L4_MACS * decodes_per_launch / (stage['l4'] * 1e-3) / 1e12
    # algorithmic bytes of the fused tail per launch (SURVEY 8d variant B): K * 4 MiB of 16-bit activations + target
    # bits + counts per object
    tail_bytes = B_PER_RATE * (K * 2_097_152 * 2 + 262_144 // 8 + 24)
    tail_gbs = tail_bytes / (stage['tail'] * 1e-3) / 1e9
    t_l4, t_tail = traffic.get('l4', {}), traffic.get('tail', {})
    roofline = {'bound': 'tensor', 'kernel': 'convt_l4_sw_kernel (128->64 ConvT, 2-CTA weight-stationary w-sweep)',
                'achieved': l4_tflops, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': l4_tflops / tpeak,
                'traffic': t_l4.get('dram_bytes'), 'traffic_unit': 'bytes per launch (ncu dram__bytes_read + write)',
                'traffic_source': t_l4.get('source'), 'algorithmic_bytes_per_launch': decodes_per_launch * 5_242_880,
                'peak_source': f'{src} sustained bf16', 'stage_ms': stage,
                'decoder_tflops_all_stages': FLOP_PER_DECODE * decodes_per_launch / (sum(stage.values()) * 1e-3) / 1e12,
                'fused_tail': {'bound': 'hbm', 'kernel': 'tail_hcol_kernel (final ConvT + sigmoid + K-mean + threshold + counts)',
                               'achieved': tail_gbs, 'peak': hpeak, 'unit': 'GB/s', 'frac': tail_gbs / hpeak,
                               'traffic': t_tail.get('dram_bytes'), 'traffic_source': t_tail.get('source'),
                               'algorithmic_bytes_per_launch': tail_bytes}}

    # ---- on-box parity spot check of the timed step (rank 0): first object of the 25 % batch, last of the 75 % batch
    spot = None
    if rank == 0:
        seed_last = 1000 + args.steps - 1
        z0, m0 = batches[0]
        z2, m2 = batches[2]
        a = _guard(parity_spot, a3d, dec, MODELNET_DECODER, ws_default, z0, m0, mu, bits, seed_last, base_off, [0],
                   last_counts[0].cpu().numpy())
        b = _guard(parity_spot, a3d, dec, MODELNET_DECODER, ws_default, z2, m2, mu, bits, seed_last,
                   base_off + 2 * B_PER_RATE, [B_PER_RATE - 1], last_counts[2].cpu().numpy())
        if 'error' in a or 'error' in b:
            spot = {'error': a.get('error') or b.get('error')}
        else:
            spot = {'weights': 'Keras-default init (probabilities within 3e-4 of 0.5: the flip test is the binding one)',
                    'objects': a['objects'] + [2 * B_PER_RATE + v for v in b['objects']], 'K': K,
                    'max_dp': max(a['max_dp'], b['max_dp']), 'flips': a['flips'] + b['flips'],
                    'flips_pct': 100.0 * (a['flips'] + b['flips']) / (a['voxels'] + b['voxels']),
                    'dcounts': a['dcounts'] + b['dcounts'],
                    'iou_delta_mean': 0.5 * (a['iou_delta_mean'] + b['iou_delta_mean']),
                    'iou_delta_per_object': [a['iou_delta_mean'], b['iou_delta_mean']],
                    'dcounts_vs_timed_batch': a['dcounts_vs_timed_batch'] + b['dcounts_vs_timed_batch'],
                    'ok': a['ok'] and b['ok']}

    # ---- the same step on trained-like weights (logit sigma ~ 3, occupancy ~ 10 %): value + parity spot
    aux_trained = None
    if not args.no_aux:
        ws_tr = dr.trained_like_weights(MODELNET_DECODER, 102)
        dec.set_weights(ws_tr)
        warm.zero_()
        step(0, warm)
        keep_tr = []
        ms_tr, tot_tr = timed_steps(3, keep=keep_tr)
        aux_trained = {'weights': 'trained-like generator (oracle/decoder_ref.py), same inputs as the headline step',
                       'value': obj_per_step * 3 / (ms_tr * 1e-3), 'unit': 'objects/s', 'steps': 3,
                       'counts_tp_fp_fn': [int(v) for v in tot_tr.tolist()]}
        if rank == 0:
            z0, m0 = batches[0]
            aux_trained['parity_spot'] = _guard(parity_spot, a3d, dec, MODELNET_DECODER, ws_tr, z0, m0, mu, bits, 1000 + 2, base_off,
                                                [0, B_PER_RATE - 1], keep_tr[0].cpu().numpy())
        dec.set_weights(ws_default)

    aux3 = aux1 = aux_grid = aux4 = aux5 = None
    if not args.no_aux:
        if rank == 0 and world == 1:   # single-call and host-return measurements: single-process runs only
            aux3 = _guard(config3_aux, a3d, dev, args.dtype)
            aux1 = _guard(config1_aux, a3d, torch, dev, ws_default, args.dtype)
            aux_grid = _guard(decode_grid_aux, a3d, torch, dev, ws_default, args.dtype)
        dec.close()                    # release the 22 GB arena before the aux handles allocate theirs
        aux4 = _guard(config4_aux, a3d, torch, dist, dev, ws_default, rank, world, args.dtype)
        aux5 = _guard(config5_aux, a3d, torch, dist, dev, ws_default, rank, world, args.dtype)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            s = CpuSample(CPU_SAMPLE_OBJECTS)     # ~10 s of CPU work on the 16-thread host (bounded sample)
            secs = s.run()
            cpu = {'value': s.n_obj / secs, 'unit': 'objects/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': s.describe(secs)}
        line = {
            'metric': 'anytime voxel reconstructions/sec', 'value': value, 'unit': 'objects/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype + ' operands, f32 accumulate',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'decodes_per_step_per_gpu': B_PER_RATE * K * len(RATES), 'parallelism': f'objects sharded x{world}',
                       'collective': 'one all-reduce of the [3] int64 counts after the last step (inside the timed region)',
                       'l2_policy': 'activation working set per step (65 GB/GPU) is streamed through HBM, >> 126 MB L2'},
            'decodes_per_s': value * K,
            'e2e': {'value': e2e_val, 'unit': 'objects/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
            'parity_spot': spot, 'tf_available': tf_available(),
            'aux_trained_like_weights': aux_trained,
            'aux_config1_single_calls': aux1, 'aux_decode_grid_e2e': aux_grid,
            'aux_config3_images_to_voxels': aux3, 'aux_config4_prefix_sweep': aux4, 'aux_config5_large_batch': aux5,
            'counts_tp_fp_fn': [int(v) for v in total.tolist()], 'e2e_counts': [int(v) for v in e2e_counts.tolist()],
        }
        _restore_stdout(saved_stdout)
        print(json.dumps(line), flush=True)
        saved_stdout = _stdout_to_stderr()   # anything printed during teardown stays off stdout
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
