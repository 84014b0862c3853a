"""Tiny end-to-end case for compute-sanitizer (decode + anytime_eval, both L4 kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr, anytime_ref as ar
ws = dr.keras_default_weights(MODELNET_DECODER, 1)
dec = a3d.decoder3D(MODELNET_DECODER, max_chunk=32)
dec.set_weights(ws)
rng = np.random.default_rng(0)
z = rng.standard_normal((3, 64)).astype(np.float32)
mask = ar.bernoulli_mask(rng, 3, 64, 0.5)
mu = rng.standard_normal((40, 64)).astype(np.float32)
tgt = ar.make_targets(rng, 3)
r = a3d.anytime_eval(dec, z, mask, mu, tgt, K=2, seed=1, return_grid=True)
p = dec(z)
torch.cuda.synchronize()
print('ok', r['counts'].sum(0).tolist(), float(p.mean()))
