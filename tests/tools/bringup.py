"""GPU bring-up diagnostic: per-layer comparison of the CUDA path with the torch-fp32 oracle.
Usage: python tests/tools/bringup.py [simt|tcgen05] [fp16|bf16] [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr

impl = sys.argv[1] if len(sys.argv) > 1 else 'tcgen05'
dt = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
wkind = sys.argv[4] if len(sys.argv) > 4 else 'trained'
st = MODELNET_DECODER
ws = dr.trained_like_weights(st, 11) if wkind == 'trained' else dr.keras_default_weights(st, 11)
rng = np.random.default_rng(5)
z = dr.round_bf16(rng.standard_normal((n, 64)).astype(np.float32))
t0 = time.time()
ref, layers = dr.decoder_forward(st, ws, z, return_layers=True)
print(f'oracle {time.time()-t0:.2f}s', flush=True)
dec = a3d.decoder3D(st, max_chunk=max(32, (n + 31) // 32 * 32), operand_dtype=dt, impl=impl)
dec.set_weights(ws)
t0 = time.time()
out = dec(z)
torch.cuda.synchronize()
print(f'gpu {time.time()-t0:.3f}s impl={impl} dtype={dt} weights={wkind}', flush=True)
for li in range(5):
    g = dec.debug_layer(li, n)
    r = layers[li].numpy()
    err = np.abs(g - r)
    print(f'layer {li} shape {g.shape} max|ref| {np.abs(r).max():.4f} maxerr {err.max():.5f} meanerr {err.mean():.6f} '
          f'relRMS {np.sqrt((err**2).mean())/np.sqrt((r**2).mean()):.2e}', flush=True)
    if err.max() > 0.05 * max(1.0, np.abs(r).max()):
        idx = np.unravel_index(err.argmax(), err.shape)
        print('   worst at', idx, 'gpu', g[idx], 'ref', r[idx])
        bad = err > 0.05 * max(1.0, np.abs(r).max())
        print('   bad fraction', bad.mean(), 'by n', bad.reshape(n, -1).mean(1)[:8],
              'by parity(d,h,w)', [float(bad[:, pd::2, ph::2, pw::2].mean()) for pd in (0, 1) for ph in (0, 1) for pw in (0, 1)] if li > 1 else '')
r = ref.numpy()
err = np.abs(out - r)
flips = ((out >= 0.5) != (r >= 0.5)).mean()
print(f'prob maxerr {err.max():.5f} meanerr {err.mean():.2e} flips {100*flips:.4f}% occupancy {(r>=0.5).mean():.3f}')
