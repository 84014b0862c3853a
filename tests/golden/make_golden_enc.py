"""Generates tests/golden/golden_enc_v1.npz from the image-encoder oracle (oracle/encoder2d_ref.py).  The reference
cannot run here (TensorFlow missing) and ships no goldens -> PARITY UNPINNED; the fixture pins the oracle against
regressions and gives the GPU tests fixed inputs / outputs.  Weights are regenerated from seeds and pinned by checksums.

    python tests/golden/make_golden_enc.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import encoder2d_ref as er  # noqa: E402


def main():
    out = {}
    layers = er.layer_list()
    for tag, size, wseed in (('s64', 64, 301), ('s256', 256, 302)):
        ws = er.trained_like_weights(layers, 3, seed=wseed, hw=size)
        rng = np.random.Generator(np.random.PCG64(9000 + size))
        x = rng.uniform(0, 1, (2, size, size, 3)).astype(np.float32)
        y, outs = er.forward(layers, ws, x, return_layers=True)
        out[f'{tag}_wsum'] = np.array([np.float64(np.asarray(w, np.float64).sum()) for w in ws])
        out[f'{tag}_wabs'] = np.array([np.float64(np.abs(np.asarray(w, np.float64)).sum()) for w in ws])
        out[f'{tag}_x_sum'] = np.float64(x.astype(np.float64).sum())
        out[f'{tag}_out'] = y.numpy()
        out[f'{tag}_layer_sum'] = np.array([float(o.double().sum()) for o in outs])
        out[f'{tag}_layer_abs'] = np.array([float(o.double().abs().sum()) for o in outs])
    enc_out = out['s256_out']
    mean, logvar, z = er.split_sample(enc_out, 16, seed=0xC0FFEE, obj_offset=5)
    out['split_mean'], out['split_logvar'], out['split_z'] = mean, logvar, z
    out['latent_normals'] = er.latent_normals(99, np.array([0, 7, 2 ** 33 + 1], np.uint64), 16)
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_enc_v1.npz')
    np.savez_compressed(p, **out)
    print('wrote', p, os.path.getsize(p), 'bytes')


if __name__ == '__main__':
    main()
