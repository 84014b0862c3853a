"""Cross-check of the 128->64 layer kernels on the device (diagnostic).  The layer implementation is chosen at handle
creation by A3D_L4_IMPL (unset = w-sweep kernel convt_l4_sw.cu, 'generic' = 1-CTA kernel of
convt_tc.cu); this script decodes the same latents with each, compares the 32^3 x 64 activations with each other and with
the oracle, localises mismatches (by decode, parity class, d, h, w, channel) and times the layer on a full chunk.

Usage: python tests/tools/l4_check.py [n_small] [n_big] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER
from oracle import decoder_ref as dr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
st = MODELNET_DECODER
ws = dr.trained_like_weights(st, 11)
rng = np.random.default_rng(5)
z = dr.round_bf16(rng.standard_normal((n, 64)).astype(np.float32))


def make(impl, max_chunk):
    if impl:
        os.environ['A3D_L4_IMPL'] = impl
    else:
        os.environ.pop('A3D_L4_IMPL', None)
    d = a3d.decoder3D(st, max_chunk=max_chunk)
    d.set_weights(ws)
    return d


def localise(tag, g, r):
    err = np.abs(g - r)
    tol = 0.02 * max(1.0, float(np.abs(r).max()))
    bad = err > tol
    print(f'{tag}: max|ref| {np.abs(r).max():.4f} maxerr {err.max():.5f} relRMS '
          f'{np.sqrt((err ** 2).mean()) / np.sqrt((r ** 2).mean()):.2e} bad {bad.mean():.5f}', flush=True)
    if bad.any():
        print('   by decode', np.round(bad.reshape(bad.shape[0], -1).mean(1), 3).tolist())
        print('   by class (pd,ph,pw)', [round(float(bad[:, pd::2, ph::2, pw::2].mean()), 3) for pd in (0, 1) for ph in (0, 1) for pw in (0, 1)])
        print('   by d', np.round(bad.mean((0, 2, 3, 4)), 2).tolist())
        print('   by h', np.round(bad.mean((0, 1, 3, 4)), 2).tolist())
        print('   by w', np.round(bad.mean((0, 1, 2, 4)), 2).tolist())
        print('   by channel/8', np.round(bad.mean((0, 1, 2, 3)).reshape(8, 8).mean(1), 2).tolist())
        idx = np.unravel_index(err.argmax(), err.shape)
        print('   worst at', idx, 'gpu', g[idx], 'ref', r[idx])
    return not bad.any()


_, layers = dr.decoder_forward(st, ws, z, return_layers=True)
ref = layers[4].numpy()
outs = {}
ok = True
for impl in (None, 'generic'):
    d = make(impl, 32)
    d(torch.from_numpy(z).cuda())     # one chunk through a3d_decode (the numpy path decodes in sub-chunks)
    torch.cuda.synchronize()
    outs[impl] = d.debug_layer(4, n)
    ok &= localise(f'L4 impl={impl or "sw"} vs oracle (n={n})', outs[impl], ref)
    del d
ok &= localise('L4 sw vs generic', outs[None], outs['generic'])
print('exact equal fraction sw vs generic', float((outs[None] == outs['generic']).mean()))

if n_big > 0:
    zc = torch.from_numpy(rng.standard_normal((n_big, 1, 64)).astype(np.float32)).cuda()
    bits = torch.zeros((n_big, 32768), dtype=torch.uint8, device='cuda')
    for impl in (None, 'generic'):
        d = make(impl, n_big)
        d.set_profiling(True)
        rec = []
        for i in range(reps + 1):
            a3d.anytime_eval(d, None, None, None, bits, z_completed=zc)
            rec.append(d.stage_times_ms())
        med = {k: round(float(np.median([r[k] for r in rec[1:]])), 3) for k in rec[0]}
        tf = 2 * 1952382976 * n_big / (med['l4'] * 1e-3) / 1e12 if 'l4' in med else float('nan')
        print(f'stage ms per {n_big} decodes impl={impl or "sw"}: {med}  l4 = {tf:.0f} TFLOP/s', flush=True)
        a4 = d.debug_layer(4, 8)
        del d
print('L4_CHECK', 'OK' if ok else 'MISMATCH')
