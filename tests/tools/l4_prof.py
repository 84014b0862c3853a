"""One full chunk (default 4096 decodes) through the decoder a few times: the command ncu profiles for the 128->64
layer (`-k regex:convt_l4`).  Usage: python tests/tools/l4_prof.py [n] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = a3d.decoder3D(MODELNET_DECODER, max_chunk=n)
d.set_weights(dr.keras_default_weights(MODELNET_DECODER, 1))
rng = np.random.default_rng(0)
zc = torch.from_numpy(rng.standard_normal((n // 16, 16, 64)).astype(np.float32)).cuda()
bits = torch.zeros((n // 16, 32768), dtype=torch.uint8, device='cuda')
for i in range(reps):
    r = a3d.anytime_eval(d, None, None, None, bits, z_completed=zc)
torch.cuda.synchronize()
print('ok', r['counts'].sum(0).tolist())
