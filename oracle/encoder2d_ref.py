"""torch-CPU restatement of the reference image encoder: Darknet19 backbone + head2D (TEST INFRASTRUCTURE, see
oracle/__init__.py; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this).

Follows /root/reference/src/net_core/darknet.py:

* ``Darknet19Conv``  :83-94    Conv2D(filters, k, strides=1, 'same', use_bias=False) -> BatchNormalization -> act
* ``Darknet19``      :96-133   18 convs (3x3 / 1x1), MaxPool2D(2, 2, 'same') after convs 1, 2, 5, 8, 13
* ``convHead``       :135-147  Conv2D('same', no bias) -> BN -> act
* ``head2D``         :149-168  [convHead]* -> Conv2D(output_dim, 1, no bias, no BN, no activation) -> reduce_max /
                               reduce_mean over (H, W) (``last_pooling``)

and the latent split of the callers (src/module/nolbo.py:869-875): ``mean = out[..., :D]``,
``logvar = clip(out[..., D:2D], -10, 10)``, ``z = sampling(mean, logvar)``.

Keras semantics restated (tf.keras 2.x, un-vendored, unpinned): Conv2D kernel variable ``[kh, kw, Cin, Cout]``
(cross-correlation, SAME = pad (k-1)/2 on each side for odd k and stride 1); BN inference
``gamma*(x-mean)/sqrt(var+1e-3)+beta``; ELU alpha 1; LeakyReLU(alpha=0.1); MaxPool2D(2,2,'same') on even sizes is the
plain 2x2/stride-2 max; variable order of ``get_weights()``: per layer in creation order (kernel, gamma, beta,
moving_mean, moving_variance).  Default init: glorot_uniform with fans (kh*kw*Cin, kh*kw*Cout).

PARITY UNPINNED: the reference holds no goldens for this path and TensorFlow is not installable here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .decoder_ref import BN_EPS, round_bf16

# (filters, kernel size) of the 18 Darknet19 convolutions and the conv indices followed by a 2x2 max-pool
DARKNET19_CONVS = [(32, 3), (64, 3), (128, 3), (64, 1), (128, 3), (256, 3), (128, 1), (256, 3),
                   (512, 3), (256, 1), (512, 3), (256, 1), (512, 3),
                   (1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3)]
DARKNET19_POOL_AFTER = (0, 1, 4, 7, 12)

# encoder head of test_pascal_VAE_dr.py:186-195 (latent 16 -> output_dim 32, no hidden head convs, max pooling; the
# pooling mode is set where the model is built, src/module/nolbo.py, ``last_pooling='max'``)
PASCAL_HEAD = {'name': 'nolbo_head', 'output_dim': 32, 'filter_num_list': [], 'filter_size_list': [],
               'activation': 'elu', 'last_pooling': 'max'}


def layer_list(head: dict | None = PASCAL_HEAD, backbone: bool = True, activation: str = 'elu') -> list[dict]:
    """Flat layer list of backbone (+ head): dicts with kind in {'conv', 'maxpool', 'global_max', 'global_avg'}."""
    L = []
    if backbone:
        for i, (f, k) in enumerate(DARKNET19_CONVS):
            L.append({'kind': 'conv', 'filters': f, 'ksize': k, 'bn': True, 'act': activation})
            if i in DARKNET19_POOL_AFTER:
                L.append({'kind': 'maxpool'})
    if head is not None:
        for f, k in zip(head['filter_num_list'], head['filter_size_list']):
            L.append({'kind': 'conv', 'filters': f, 'ksize': k, 'bn': True, 'act': head['activation']})
        L.append({'kind': 'conv', 'filters': head['output_dim'], 'ksize': 1, 'bn': False, 'act': None})
        lp = head.get('last_pooling')
        if lp == 'max':
            L.append({'kind': 'global_max'})
        elif lp == 'average':
            L.append({'kind': 'global_avg'})
    return L


def weight_shapes(layers: list[dict], in_ch: int) -> list[tuple[str, tuple[int, ...]]]:
    out = []
    c = in_ch
    j = 0
    for l in layers:
        if l['kind'] != 'conv':
            continue
        out.append((f'conv{j}/kernel', (l['ksize'], l['ksize'], c, l['filters'])))
        if l['bn']:
            for nm in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                out.append((f'bn{j}/{nm}', (l['filters'],)))
        c = l['filters']
        j += 1
    return out


def keras_default_weights(layers, in_ch: int, seed: int, bf16_kernels: bool = True) -> list[np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    ws = []
    for name, shape in weight_shapes(layers, in_ch):
        if name.endswith('/kernel'):
            rec = shape[0] * shape[1]
            lim = np.sqrt(6.0 / (rec * shape[2] + rec * shape[3]))
            w = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            ws.append(round_bf16(w) if bf16_kernels else w)
        elif name.endswith('/gamma') or name.endswith('/moving_variance'):
            ws.append(np.ones(shape, np.float32))
        else:
            ws.append(np.zeros(shape, np.float32))
    return ws


_ACT = {'elu': lambda t: F.elu(t, alpha=1.0), 'relu': F.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.1), None: lambda t: t}


def forward(layers, weights, x_nhwc, dtype=torch.float32, return_layers: bool = False, calibrate=None):
    """Run the layer list on NHWC input ``x_nhwc`` ([N,H,W,C]).  Returns NHWC (or [N,C] after a global pool).

    ``calibrate(conv_index, pre_bn_tensor_nchw) -> (mean, var)`` lets the weight generator set BN statistics in place.
    """
    x = torch.as_tensor(np.asarray(x_nhwc), dtype=dtype).permute(0, 3, 1, 2).contiguous()
    ws = [torch.as_tensor(np.asarray(w), dtype=dtype) for w in weights]
    it = iter(range(len(ws)))
    outs = []
    j = 0
    for l in layers:
        kind = l['kind']
        if kind == 'conv':
            k = ws[next(it)]                                     # [kh,kw,Cin,Cout]
            p = (l['ksize'] - 1) // 2
            x = F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), padding=p)   # darknet.py:84-85 / :136-138 / :155-157
            if l['bn']:
                gi, bi, mi, vi = next(it), next(it), next(it), next(it)
                if calibrate is not None:
                    upd = calibrate(j, x)
                    if upd is not None:
                        ws[mi], ws[vi] = upd
                        weights[mi], weights[vi] = upd[0].numpy(), upd[1].numpy()
                v = lambda t: t.view(1, -1, 1, 1)
                x = v(ws[gi]) * (x - v(ws[mi])) / torch.sqrt(v(ws[vi]) + BN_EPS) + v(ws[bi])   # :86 / :139
            x = _ACT[l['act']](x)                                # :87-92
            j += 1
        elif kind == 'maxpool':
            x = F.max_pool2d(x, 2, 2)                            # MaxPool2D(2, 2, 'same') on even sizes, :100 ...
        elif kind == 'global_max':
            x = x.amax(dim=(2, 3))                               # tf.reduce_max(axis=[1,2]) :159-160
        elif kind == 'global_avg':
            x = x.mean(dim=(2, 3))                               # :162-163
        if return_layers:
            outs.append(x.permute(0, 2, 3, 1).contiguous() if x.dim() == 4 else x.clone())
    y = x.permute(0, 2, 3, 1).contiguous() if x.dim() == 4 else x
    return (y, outs) if return_layers else y


def trained_like_weights(layers, in_ch: int, seed: int, hw: int = 64, calib: int = 2) -> list[np.ndarray]:
    """Glorot kernels (bf16-representable), randomised gamma/beta, BN moving statistics calibrated layer by layer on
    random U[0,1] images so every hidden activation is ~N(0,1) before the non-linearity (like a trained net)."""
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    ws = keras_default_weights(layers, in_ch, seed, bf16_kernels=True)
    names = [n for n, _ in weight_shapes(layers, in_ch)]
    for i, n in enumerate(names):
        if n.endswith('/gamma'):
            ws[i] = rng.uniform(0.6, 1.4, ws[i].shape).astype(np.float32)
        elif n.endswith('/beta'):
            ws[i] = (0.25 * rng.standard_normal(ws[i].shape)).astype(np.float32)
    x = rng.uniform(0, 1, (calib, hw, hw, in_ch)).astype(np.float32)

    def cal(j, pre):
        mean = pre.mean(dim=(0, 2, 3))
        var = pre.var(dim=(0, 2, 3), unbiased=False)
        jit = torch.from_numpy(rng.uniform(0.8, 1.25, mean.shape).astype(np.float32))
        m = mean + 0.1 * var.sqrt() * torch.from_numpy(rng.standard_normal(mean.shape).astype(np.float32))
        return m.float(), (var * jit + 1e-6).float()

    forward(layers, ws, x, calibrate=cal)
    # the last convolution of a head has no BN: rescale its kernel so the pre-pool output is ~N(0, 1) like the
    # (mean, logvar) of a trained VAE encoder (keeps logvar away from the +-10 clip)
    conv_idx = [i for i, n in enumerate(names) if n.endswith('/kernel')]
    last = conv_idx[-1]
    if last == len(names) - 1:
        pre = forward([l for l in layers if l['kind'] not in ('global_max', 'global_avg')], ws, x)
        ws[last] = round_bf16((ws[last] / max(float(pre.std()), 1e-6)).astype(np.float32))
    return [np.ascontiguousarray(w, dtype=np.float32) for w in ws]


def split_latent(enc_out: np.ndarray, D: int, clip: float = 10.0):
    """nolbo.py:869-873: mean = out[..., :D]; logvar = clip(out[..., D:2D], -10, 10)."""
    enc_out = np.asarray(enc_out, np.float32)
    return enc_out[..., :D].copy(), np.clip(enc_out[..., D:2 * D], -clip, clip)


LATENT_STREAM = 0x5A4D504C   # counter word 1 of the encoder's sampling() draws (include/a3d.h, a3d_enc2d_split_sample)


def latent_normals(seed: int, obj_ids: np.ndarray, D: int) -> np.ndarray:
    """N(0,1) draws [B, D] of the encoder's latent sampler: Philox4x32-10, counter (dim/4, LATENT_STREAM, obj lo, obj hi),
    key = seed, Box-Muller on fp32 uniforms (same contract as oracle.anytime_ref.philox_normals)."""
    from .anytime_ref import philox4x32_10
    nq = (D + 3) // 4
    obj = np.asarray(obj_ids, dtype=np.uint64)
    ctr = np.zeros((len(obj), nq, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(nq, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.uint32(LATENT_STREAM)
    ctr[..., 2] = (obj & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[..., 3] = (obj >> np.uint64(32)).astype(np.uint32)[:, None]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    w = philox4x32_10(ctr, key.reshape(1, 1, 2))
    u = (w.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float64)
    out = np.empty(w.shape, dtype=np.float64)
    for a in (0, 2):
        r = np.sqrt(-2.0 * np.log(u[..., a]))
        out[..., a] = r * np.cos(2.0 * np.pi * u[..., a + 1])
        out[..., a + 1] = r * np.sin(2.0 * np.pi * u[..., a + 1])
    return out.reshape(len(obj), nq * 4)[:, :D]


def split_sample(enc_out: np.ndarray, D: int, seed: int, obj_offset: int = 0, clip: float = 10.0):
    """nolbo.py:869-875 + function.py:35-38 with the seeded draws above: returns (mean, logvar, z), float64 math."""
    mean, logvar = split_latent(enc_out, D, clip)
    eps = latent_normals(seed, np.arange(len(mean), dtype=np.uint64) + np.uint64(obj_offset), D)
    z = mean.astype(np.float64) + np.sqrt(np.exp(logvar.astype(np.float64))) * eps
    return mean, logvar, z.astype(np.float32)


def conv2d_same_definition(x: np.ndarray, k: np.ndarray) -> np.ndarray:
    """fp64 Conv2D(strides=1, 'same') straight from the definition (cross-correlation, zero padding (k-1)/2):
    y[n,h,w,co] = sum_{dy,dx,ci} x[n, h+dy-p, w+dx-p, ci] * k[dy,dx,ci,co].  Pins the torch restatement above."""
    x = np.asarray(x, np.float64)
    k = np.asarray(k, np.float64)
    N, H, W, _ = x.shape
    kh, kw, _, co = k.shape
    p = (kh - 1) // 2
    y = np.zeros((N, H, W, co))
    for dy in range(kh):
        for dx in range(kw):
            for h in range(H):
                hh = h + dy - p
                if hh < 0 or hh >= H:
                    continue
                for w in range(W):
                    ww = w + dx - p
                    if ww < 0 or ww >= W:
                        continue
                    y[:, h, w, :] += x[:, hh, ww, :] @ k[dy, dx]
    return y


def encoder_macs(layers, H: int, W: int, C: int):
    """(algorithmic, dense) MACs per image: algorithmic counts only taps that land inside the image."""
    alg = dense = 0
    for l in layers:
        if l['kind'] == 'conv':
            k = l['ksize']
            alg += (k * H - (2 if k == 3 else 0)) * (k * W - (2 if k == 3 else 0)) * C * l['filters']
            dense += k * k * H * W * C * l['filters']
            C = l['filters']
        elif l['kind'] == 'maxpool':
            H, W = H // 2, W // 2
    return alg, dense
