"""Ragged / border cases of the decoder path, run through whichever build A3D_LIB selects; prints one JSON object of
SHA-1 digests of every result.  tests/test_gpu_checked_build.py runs it on the release library and on the bounds-checked
build (liba3d_checked.so: device-side range checks of every hand-computed index, csrc/internal.h) and compares: no check
may fire, and the checks may not change a single bit.  Usage: [A3D_LIB=liba3d_checked.so] python tests/tools/checked_cases.py"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import a3d
from a3d.presets import MODELNET_DECODER, PASCAL_DECODER
from oracle import decoder_ref as dr, anytime_ref as ar

out = {'lib': os.path.basename(a3d._capi.LIB_PATH)}


def dig(*arrs):
    h = hashlib.sha1()
    for a in arrs:
        a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def make(st, ws, max_chunk=32, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    d = a3d.decoder3D(st, max_chunk=max_chunk, **kw)
    for k in (env or {}):
        os.environ.pop(k)
    d.set_weights(ws)
    return d


rng = np.random.default_rng(11)
ws = dr.trained_like_weights(MODELNET_DECODER, 5)
mu = rng.standard_normal((40, 64)).astype(np.float32)
variants = {'default': {}, 'l4_generic': {'env': {'A3D_L4_IMPL': 'generic'}}, 'tail_simt': {'env': {'A3D_TAIL_IMPL': 'simt'}},
            'pair1': {'env': {'A3D_CONV_PAIR': '1'}}, 'pair2': {'env': {'A3D_CONV_PAIR': '2'}},
            'pair3': {'env': {'A3D_CONV_PAIR': '3'}}, 'simt': {'impl': 'simt'},
            'bf16': {'operand_dtype': 'bf16'}}
for name, kw in variants.items():
    rng = np.random.default_rng(12)                   # every variant sees the same inputs (the test compares their digests)
    dec = make(MODELNET_DECODER, ws, **kw)
    sizes = (1, 5, 33) if name in ('simt', 'tail_simt') else (1, 5, 21, 33, 40, 72)
    for n in sizes:                                   # decoder(z): ragged n, several chunks of 32
        z = rng.standard_normal((n, 64)).astype(np.float32)
        out[f'{name}/decode/{n}'] = dig(dec(torch.from_numpy(z).cuda()))
    for B, K in ((3, 1), (2, 3), (5, 2), (1, 16), (33, 1)):    # fused evaluation: K = 1 / odd / even, ragged chunks
        if name == 'simt' and B * K > 16:
            continue
        z = rng.standard_normal((B, 64)).astype(np.float32)
        mask = ar.bernoulli_mask(rng, B, 64, 0.5)
        tgt = ar.make_targets(rng, B)
        for fill in ('prior_sample', 'mean', 'normal'):
            r = a3d.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=3, fill=fill, return_grid=(fill == 'prior_sample'))
            out[f'{name}/eval/{B}x{K}/{fill}'] = dig(r['counts'], r['z_completed'], *([r['mean_prob']] if 'mean_prob' in r else []))
        r = a3d.anytime_eval(dec, z, mask, mu, tgt, K=K, seed=3, return_loss=True, gamma=0.97)
        out[f'{name}/loss/{B}x{K}'] = dig(r['counts'], r['loss'])
    dec.close()
# host-return path (a3d_decode_host: sub-chunks, converted outputs), the other activations, the Pascal decoder (D = 16)
dec = make(MODELNET_DECODER, ws, max_chunk=64)
z = rng.standard_normal((37, 64)).astype(np.float32)
for dt in ('f32', 'f16', 'bits'):
    out[f'host/{dt}'] = dig(dec(z, out_dtype=dt))
dec.close()
for act, fin in (('relu', 'sigmoid'), ('lrelu', 'sigmoid'), ('elu', 'None')):
    st = dict(MODELNET_DECODER, activation=act, final_activation=fin)
    dec = make(st, ws)
    out[f'act/{act}/{fin}'] = dig(dec(torch.from_numpy(z[:9]).cuda()))
    dec.close()
dec = make(PASCAL_DECODER, dr.trained_like_weights(PASCAL_DECODER, 6))
out['pascal/decode/7'] = dig(dec(torch.from_numpy(rng.standard_normal((7, 16)).astype(np.float32)).cuda()))
dec.close()
torch.cuda.synchronize()
print(json.dumps(out))
