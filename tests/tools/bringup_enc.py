"""Bring-up / diagnostic: per-layer comparison of the CUDA image encoder against the torch-CPU oracle.

    python tests/tools/bringup_enc.py [--size 64] [--n 3] [--dtype fp16] [--init trained|keras]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import a3d  # noqa: E402
from oracle import encoder2d_ref as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=64)
    ap.add_argument('--n', type=int, default=3)
    ap.add_argument('--dtype', default='fp16')
    ap.add_argument('--init', default='trained')
    ap.add_argument('--max-batch', type=int, default=8)
    args = ap.parse_args()
    layers = E.layer_list()
    if args.init == 'trained':
        ws = E.trained_like_weights(layers, 3, seed=7, hw=min(args.size, 64))
    else:
        ws = E.keras_default_weights(layers, 3, seed=7)
    rng = np.random.Generator(np.random.PCG64(11))
    x = rng.uniform(0, 1, (args.n, args.size, args.size, 3)).astype(np.float32)
    ref, ref_layers = E.forward(layers, ws, x, return_layers=True)
    enc = a3d.image_encoder(a3d.presets.PASCAL_ENCODER_HEAD, input_size=(args.size, args.size),
                            max_batch=args.max_batch, operand_dtype=args.dtype)
    enc.set_weights(ws)
    out = enc(x)
    worst = 0.0
    for li, l in enumerate(layers):
        if l['kind'] in ('global_max', 'global_avg'):
            continue
        fused = l['kind'] == 'conv' and li + 1 < len(layers) and layers[li + 1]['kind'] == 'maxpool'
        if fused:
            continue   # stored pooled: compared at the pool's index
        if li == len(layers) - 2:
            continue   # final conv before the global pool: fp32 buffer [pixels, C] (checked through the output)
        got = enc.debug_layer(li, min(args.n, args.max_batch))
        want = ref_layers[li].numpy()[-got.shape[0]:] if args.n > args.max_batch else ref_layers[li].numpy()[:got.shape[0]]
        if args.n > args.max_batch:
            last = args.n % args.max_batch or args.max_batch
            got = got[:last]
            want = ref_layers[li].numpy()[args.n - last:]
        err = np.abs(got - want).max()
        scale = np.abs(want).max()
        worst = max(worst, err / max(scale, 1e-6))
        print(f'layer {li:2d} {l["kind"]:8s} shape {got.shape} max|ref| {scale:8.4f} max err {err:.3e} rel {err / max(scale, 1e-6):.3e}')
    err = np.abs(out - ref.numpy()).max()
    print(f'output {out.shape} max|ref| {np.abs(ref.numpy()).max():.4f} max err {err:.3e}; launches {enc.launch_count}; '
          f'arena {enc.workspace_bytes() / 2**20:.1f} MiB; worst rel {worst:.3e}')


if __name__ == '__main__':
    main()
