"""Per-stage device times of one anytime_eval chunk (diagnostic).  Usage: python tests/tools/stage_times.py [n_obj] [K] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=B * K)
dec.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))
rng = np.random.default_rng(0)
zc = torch.from_numpy(rng.standard_normal((B, K, 64)).astype(np.float32)).cuda()
bits = torch.zeros((B, 32768), dtype=torch.uint8, device='cuda')
dec.set_profiling(True)
rec = []
for i in range(reps):
    a3d.anytime_eval(dec, None, None, None, bits, z_completed=zc)
    rec.append(dec.stage_times_ms())
    if i == reps - 1 or reps <= 3:
        print(os.environ.get('A3D_DEBUG_FLAGS', '0'), {k: round(v, 3) for k, v in rec[-1].items()}, flush=True)
if reps > 3:
    med = {k: round(float(np.median([r[k] for r in rec[1:]])), 3) for k in rec[0]}
    print('median', med, 'total', round(sum(med.values()), 2), flush=True)
