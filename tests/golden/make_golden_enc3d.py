"""Generates tests/golden/golden_enc3d_v1.npz from the voxel-encoder oracle (oracle/encoder3d_ref.py).  The reference
cannot run here (TensorFlow missing) and ships no goldens -> PARITY UNPINNED; the fixture pins the oracle against
regressions and gives the GPU tests fixed inputs / outputs.  Weights and voxel grids are regenerated from seeds.

    python tests/golden/make_golden_enc3d.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import anytime_ref as ar, encoder3d_ref as e3  # noqa: E402


def main():
    st = e3.MODELNET_ENCODER
    ws = e3.trained_like_weights(st, 401)
    x = ar.make_targets(np.random.Generator(np.random.PCG64(402)), 3)
    y, outs = e3.forward(st, ws, x, return_layers=True)
    out = {
        'wsum': np.array([np.float64(np.asarray(w, np.float64).sum()) for w in ws]),
        'x_occupancy': x.reshape(3, -1).sum(1).astype(np.int64),
        'out': y.numpy(),
        'layer_abs': np.array([float(o.double().abs().sum()) for o in outs]),
        'layer_samples': np.stack([o.numpy().reshape(3, -1)[:, ::max(1, o[0].numel() // 64)][:, :64] for o in outs[:4]]),
    }
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_enc3d_v1.npz')
    np.savez_compressed(p, **out)
    print('wrote', p, os.path.getsize(p), 'bytes')


if __name__ == '__main__':
    main()
