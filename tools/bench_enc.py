"""Throughput of the CUDA image encoder (Darknet19 + head2D) on synthetic RGB crops: images/s and TFLOP/s.

    python tools/bench_enc.py [--batch 128] [--size 256] [--max-batch 128] [--steps 10]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import a3d  # noqa: E402
from a3d.encoder2d import darknet19_layers, head2d_layers  # noqa: E402


def encoder_macs(layers, H, W, C):
    """(algorithmic MACs: taps that land inside the image, dense MACs incl. the zero padding) per image."""
    alg = dense = 0
    for l in layers:
        if l['kind'] == 'conv':
            k = l['ksize']
            kept_h = k * H - (2 if k == 3 else 0)
            kept_w = k * W - (2 if k == 3 else 0)
            alg += kept_h * kept_w * C * l['filters']
            dense += k * k * H * W * C * l['filters']
            C = l['filters']
        elif l['kind'] == 'maxpool':
            H, W = H // 2, W // 2
    return alg, dense


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--max-batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--dtype', default='fp16')
    args = ap.parse_args()
    layers = darknet19_layers() + head2d_layers(32, [], [], 'max')
    alg, dense = encoder_macs(layers, args.size, args.size, 3)
    enc = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(args.size, args.size),
                            max_batch=args.max_batch, operand_dtype=args.dtype)
    rng = np.random.Generator(np.random.PCG64(5))
    ws = []
    for shp in enc.weight_shapes():
        if len(shp) == 4:
            lim = np.sqrt(6.0 / (shp[0] * shp[1] * (shp[2] + shp[3])))
            ws.append(rng.uniform(-lim, lim, shp).astype(np.float32))
        else:
            ws.append(np.ones(shp, np.float32))
    enc.set_weights(ws)
    x = torch.rand((args.batch, args.size, args.size, 3), device='cuda')
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
    print(f'batch {args.batch} size {args.size} max_batch {args.max_batch}: {ms:.3f} ms/step, '
          f'{args.batch / ms * 1e3:.0f} images/s, {2 * alg * args.batch / ms / 1e9:.1f} TFLOP/s algorithmic '
          f'({alg / 1e9:.4f} GMAC/image; dense {dense / 1e9:.4f})')


if __name__ == '__main__':
    main()
