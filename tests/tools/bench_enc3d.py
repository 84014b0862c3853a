"""Throughput of the CUDA voxel encoder (encoder3D) on synthetic occupancy grids: objects/s and TFLOP/s.

    python tests/tools/bench_enc3d.py [--batch 256] [--steps 10]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import a3d  # noqa: E402
from oracle import anytime_ref as ar, encoder3d_ref as e3  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--dtype', default='fp16')
    args = ap.parse_args()
    st = a3d.presets.MODELNET_ENCODER
    alg, dense = e3.encoder_macs(st)
    enc = a3d.encoder3D(st, max_batch=args.batch, operand_dtype=args.dtype)
    enc.set_weights(e3.keras_default_weights(st, 3))
    x8 = ar.make_targets(np.random.default_rng(1), 8)
    x = torch.from_numpy(x8).cuda().repeat((args.batch + 7) // 8, 1, 1, 1, 1)[:args.batch].contiguous()
    for _ in range(3):
        enc(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        enc(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(f'batch {args.batch}: {ms:.3f} ms/step, {args.batch / ms * 1e3:.0f} objects/s, '
          f'{2 * alg * args.batch / ms / 1e9:.1f} TFLOP/s algorithmic ({alg / 1e9:.4f} GMAC/object; dense {dense / 1e9:.4f})')


if __name__ == '__main__':
    main()
