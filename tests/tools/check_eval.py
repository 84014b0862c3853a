"""GPU diagnostic: anytime_eval (K-mean + counts) vs the oracle on the GPU-completed latents.
Usage: python tests/tools/check_eval.py [tcgen05|simt] [B] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr, anytime_ref as ar
impl = sys.argv[1] if len(sys.argv) > 1 else 'tcgen05'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
st = MODELNET_DECODER
ws = dr.trained_like_weights(st, 11)
rng = np.random.default_rng(5)
z = dr.round_bf16(rng.standard_normal((B, 64)).astype(np.float32))
mu = rng.standard_normal((40, 64)).astype(np.float32)
mask = ar.bernoulli_mask(rng, B, 64, 0.5)
tgt = ar.make_targets(rng, B)
dec = a3d.decoder3D(st, max_chunk=32, impl=impl)
dec.set_weights(ws)
r = a3d.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=3, return_grid=True)
torch.cuda.synchronize()
zc = r['z_completed'].cpu().numpy()
zr, cs = ar.impute(z, mask, mu, K, seed=3)
print('impute maxerr', np.abs(zc - zr).max(), 'cstar equal', (cs == r['cstar'].cpu().numpy()).all())
ref, rc = ar.anytime_eval(st, ws, zc, tgt)
got = r['mean_prob'].cpu().numpy()
err = np.abs(got - ref)
print('mean_prob maxerr', err.max(), 'flips %', 100 * ((got >= .5) != (ref >= .5)).mean())
print('counts gpu', r['counts'].cpu().numpy().tolist()); print('counts ref', rc.tolist())
if err.max() > 1e-2:
    bad = err > 1e-2
    idx = np.argwhere(bad[..., 0])
    print('bad frac', bad.mean(), 'first bad idx', idx[:10].tolist())
    for ax in (1, 2, 3):
        print('bad by coord axis', ax, np.unique(idx[:, ax])[:70].tolist())
# decode path (K=1, full grid)
p = dec(zc[:, 0, :])
pr = dr.decoder_forward(st, ws, zc[:, 0, :]).numpy()
print('decode maxerr', np.abs(p - pr).max())
