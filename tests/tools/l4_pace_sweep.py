"""Sweep of the pacing distance of the 128->64 kernel (A3D_L4_PACE, read at handle creation): layer time per 4096 decodes.
Usage: python tests/tools/l4_pace_sweep.py [deltas...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr
deltas = [int(v) for v in sys.argv[1:]] or [0, 4, 8, 12, 16]
n = 4096
ws = dr.keras_default_weights(MODELNET_DECODER, 1)
zc = torch.from_numpy(np.random.default_rng(0).standard_normal((n // 16, 16, 64)).astype(np.float32)).cuda()
bits = torch.zeros((n // 16, 32768), dtype=torch.uint8, device='cuda')
for rep in range(2):
    for dl in deltas:
        os.environ['A3D_L4_PACE'] = str(dl)
        d = a3d.decoder3D(MODELNET_DECODER, max_chunk=n)
        d.set_weights(ws)
        d.set_profiling(True)
        rec = []
        for i in range(7):
            a3d.anytime_eval(d, None, None, None, bits, z_completed=zc)
            rec.append(d.stage_times_ms()['l4'])
        print(f'pace {dl:3d}: l4 median {np.median(rec[2:]):.3f} ms  min {min(rec[2:]):.3f}', flush=True)
        d.close()
