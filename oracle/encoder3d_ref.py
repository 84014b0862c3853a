"""torch-CPU restatement of the reference voxel encoder ``encoder3D`` (TEST INFRASTRUCTURE, see oracle/__init__.py; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this).

Follows /root/reference/src/net_core/autoencoder3D.py:

* ``conv3DEnc``  :26-39   Conv3D(filters, k, strides, 'same', use_bias=False) -> BatchNormalization -> activation
                          ('lrelu' = LeakyReLU() with the Keras default alpha 0.3, unlike darknet.py's 0.1)
* ``encoder3D``  :72-102  conv3DEnc for every entry of the lists but the last, then a bare Conv3D (no BN, no
                          activation), reduce_mean / reduce_max over (D, H, W) (``final_pool``), optional sigmoid

and the callers (src/module/nolbo.py:1463-1470): ``mean = out[..., :D]``, ``logvar = clip(out[..., D:2D], -10, 10)``,
``z = sampling(mean, logvar)`` -- shared with the image encoder (oracle/encoder2d_ref.split_sample).

Keras semantics restated (tf.keras 2.x, un-vendored, unpinned): Conv3D kernel variable ``[kd, kh, kw, Cin, Cout]``,
cross-correlation, channels_last; 'same' with stride s: out = ceil(in / s), pad_total = max((out-1)*s + k - in, 0),
pad_before = pad_total // 2, pad_after = the rest (k = 4: s = 2 -> 1/1, s = 1 -> 1/2); BN eps 1e-3; ELU alpha 1.
Variable order of ``get_weights()``: per conv3DEnc (kernel, gamma, beta, moving_mean, moving_variance), final kernel.

PARITY UNPINNED: the reference holds no goldens for this path and TensorFlow is not installable here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .decoder_ref import BN_EPS, round_bf16

# test_modelnet_VAE_dr.py:172-181 with latent_dim = 64
MODELNET_ENCODER = {
    'name': 'encoder3D',
    'input_shape': [64, 64, 64, 1],
    'filter_num_list': [64, 128, 256, 512, 128],
    'filter_size_list': [4, 4, 4, 4, 4],
    'strides_list': [2, 2, 2, 2, 1],
    'final_pool': 'average',
    'activation': 'elu',
    'final_activation': 'None',
}

_ACT = {'elu': lambda t: F.elu(t, alpha=1.0), 'relu': F.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.3),
        None: lambda t: t, 'None': lambda t: t}


def same_pads(n_in: int, k: int, s: int) -> tuple[int, int]:
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return total // 2, total - total // 2


def weight_shapes(structure: dict) -> list[tuple[str, tuple[int, ...]]]:
    out, c = [], structure['input_shape'][-1]
    n = len(structure['filter_num_list'])
    for i, (f, k) in enumerate(zip(structure['filter_num_list'], structure['filter_size_list'])):
        out.append((f'conv{i}/kernel', (k, k, k, c, f)))
        if i < n - 1:
            for nm in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                out.append((f'bn{i}/{nm}', (f,)))
        c = f
    return out


def keras_default_weights(structure: dict, seed: int, bf16_kernels: bool = True) -> list[np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    ws = []
    for name, shape in weight_shapes(structure):
        if name.endswith('/kernel'):
            rec = shape[0] * shape[1] * shape[2]
            lim = np.sqrt(6.0 / (rec * shape[3] + rec * shape[4]))     # glorot_uniform, Keras fans
            w = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            ws.append(round_bf16(w) if bf16_kernels else w)
        elif name.endswith('/gamma') or name.endswith('/moving_variance'):
            ws.append(np.ones(shape, np.float32))
        else:
            ws.append(np.zeros(shape, np.float32))
    return ws


def conv3d_same(x: torch.Tensor, k: torch.Tensor, stride: int) -> torch.Tensor:
    """x [N,C,D,H,W], k Keras layout [kd,kh,kw,Cin,Cout] -> Conv3D(strides=stride, padding='same')."""
    pads = []
    for dim in (4, 3, 2):                                   # F.pad takes the last dimension first
        pb, pa = same_pads(x.shape[dim], k.shape[0], stride)
        pads += [pb, pa]
    return F.conv3d(F.pad(x, pads), k.permute(4, 3, 0, 1, 2).contiguous(), stride=stride)


def forward(structure: dict, weights, voxels, dtype=torch.float32, return_layers: bool = False, calibrate=None):
    """voxels [N,64,64,64,1] (NDHWC) -> [N, filters[-1]] (after the final pool; NDHWC grid if final_pool is None)."""
    x = torch.as_tensor(np.asarray(voxels), dtype=dtype).permute(0, 4, 1, 2, 3).contiguous()
    ws = [torch.as_tensor(np.asarray(w), dtype=dtype) for w in weights]
    it = iter(range(len(ws)))
    n = len(structure['filter_num_list'])
    outs = []
    for i, s in enumerate(structure['strides_list']):
        x = conv3d_same(x, ws[next(it)], s)                                   # autoencoder3D.py:27-30 / :87-89
        if i < n - 1:
            gi, bi, mi, vi = next(it), next(it), next(it), next(it)
            if calibrate is not None:
                upd = calibrate(i, x)
                if upd is not None:
                    ws[mi], ws[vi] = upd
                    weights[mi], weights[vi] = upd[0].numpy(), upd[1].numpy()
            v = lambda t: t.view(1, -1, 1, 1, 1)
            x = v(ws[gi]) * (x - v(ws[mi])) / torch.sqrt(v(ws[vi]) + BN_EPS) + v(ws[bi])   # :31
            x = _ACT[structure['activation']](x)                              # :33-38
        if return_layers:
            outs.append(x.permute(0, 2, 3, 4, 1).contiguous())
    fp = structure.get('final_pool')
    if fp == 'average':
        x = x.mean(dim=(2, 3, 4))                                             # :91-92
    elif fp == 'max':
        x = x.amax(dim=(2, 3, 4))                                             # :93-94
    else:
        x = x.permute(0, 2, 3, 4, 1).contiguous()
    if structure.get('final_activation') == 'sigmoid':
        x = torch.sigmoid(x)                                                  # :98-99
    return (x, outs) if return_layers else x


def trained_like_weights(structure: dict, seed: int, calib_voxels=None) -> list[np.ndarray]:
    """Glorot kernels (bf16-representable), randomised gamma / beta, BN moving statistics calibrated on synthetic
    occupancy grids so hidden activations are ~N(0,1) before the non-linearity; the final (BN-less) kernel is rescaled
    so the pooled output is O(1) like the (mean, logvar) of a trained VAE encoder."""
    from .anytime_ref import make_targets
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    ws = keras_default_weights(structure, seed)
    names = [n for n, _ in weight_shapes(structure)]
    for i, n in enumerate(names):
        if n.endswith('/gamma'):
            ws[i] = rng.uniform(0.6, 1.4, ws[i].shape).astype(np.float32)
        elif n.endswith('/beta'):
            ws[i] = (0.25 * rng.standard_normal(ws[i].shape)).astype(np.float32)
    x = make_targets(rng, 2) if calib_voxels is None else calib_voxels

    def cal(j, pre):
        mean = pre.mean(dim=(0, 2, 3, 4))
        var = pre.var(dim=(0, 2, 3, 4), unbiased=False)
        jit = torch.from_numpy(rng.uniform(0.8, 1.25, mean.shape).astype(np.float32))
        m = mean + 0.1 * var.sqrt() * torch.from_numpy(rng.standard_normal(mean.shape).astype(np.float32))
        return m.float(), (var * jit + 1e-6).float()

    forward(structure, ws, x, calibrate=cal)
    pre = forward(dict(structure, final_pool=None, final_activation='None'), ws, x)
    ws[-1] = round_bf16((ws[-1] / max(float(pre.std()), 1e-6)).astype(np.float32))
    return [np.ascontiguousarray(w, dtype=np.float32) for w in ws]


def conv3d_same_definition(x: np.ndarray, k: np.ndarray, stride: int) -> np.ndarray:
    """fp64 Conv3D('same') straight from the definition: y[n,o,co] = sum_{t,ci} x[n, o*s + t - pad_before, ci] k[t,ci,co]
    with zero outside the grid.  Pins conv3d_same above (small cases only)."""
    x = np.asarray(x, np.float64)
    k = np.asarray(k, np.float64)
    N, Dn, Hn, Wn, _ = x.shape
    kk = k.shape[0]
    od, oh, ow = -(-Dn // stride), -(-Hn // stride), -(-Wn // stride)
    pd, ph, pw = same_pads(Dn, kk, stride)[0], same_pads(Hn, kk, stride)[0], same_pads(Wn, kk, stride)[0]
    y = np.zeros((N, od, oh, ow, k.shape[4]))
    for a in range(od):
        for b in range(oh):
            for c in range(ow):
                for td in range(kk):
                    i = a * stride + td - pd
                    if i < 0 or i >= Dn:
                        continue
                    for th in range(kk):
                        j = b * stride + th - ph
                        if j < 0 or j >= Hn:
                            continue
                        for tw in range(kk):
                            l = c * stride + tw - pw
                            if l < 0 or l >= Wn:
                                continue
                            y[:, a, b, c, :] += x[:, i, j, l, :] @ k[td, th, tw]
    return y


def encoder_macs(structure: dict) -> tuple[int, int]:
    """(algorithmic, dense) MACs per object: algorithmic counts only taps that land inside the grid."""
    g, c = structure['input_shape'][0], structure['input_shape'][-1]
    alg = dense = 0
    for f, k, s in zip(structure['filter_num_list'], structure['filter_size_list'], structure['strides_list']):
        out = -(-g // s)
        pb = same_pads(g, k, s)[0]
        kept = sum(1 for o in range(out) for t in range(k) if 0 <= o * s + t - pb < g)
        alg += kept ** 3 * c * f
        dense += (out * k) ** 3 * c * f
        g, c = out, f
    return alg, dense
