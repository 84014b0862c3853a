"""torch-CPU restatement of the reference 3-D voxel decoder (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/src/net_core/autoencoder3D.py line by line:

* ``decoder3D``        :104-139  graph builder (structure dict -> Keras model)
* ``linearTransform``  :56-70    reshape -> Dense(bias) -> BatchNormalization -> ELU
* ``conv3DDec``        :41-54    Conv3DTranspose(k, s, 'same', no bias) -> BN -> ELU
* final layer          :129-136  Conv3DTranspose(64->1, no bias, no BN) -> tf.sigmoid

Keras semantics restated here (TensorFlow 2.x tf.keras, un-vendored, version unpinned):

* ``BatchNormalization()`` at inference: ``gamma*(x-mean)/sqrt(var+1e-3)+beta`` (epsilon default 1e-3).
* ``ELU()``: alpha = 1.
* ``Conv3DTranspose(padding='same')``: output length = input*stride; it is the adjoint (input-gradient) of a
  SAME-padded forward convolution, i.e. ``out[o] = sum_{i,t : o = i*s + t - pad_before} in[i] * W[t]`` with
  ``pad_total = k - s``, ``pad_before = pad_total // 2``  (k=4: s=2 -> 1/1, s=1 -> 1 before / 2 after).
  Kernel variable layout ``[kd, kh, kw, Cout, Cin]``; no kernel flip relative to ``torch.conv_transpose3d``.
* variable order of ``model.get_weights()``: per layer in creation order, trainable before non-trainable.
* default initialisers: ``glorot_uniform`` kernels (conv fans = receptive*shape[-2], receptive*shape[-1]),
  zero bias, BN gamma=1 beta=0 moving_mean=0 moving_variance=1.

PARITY UNPINNED (no reference goldens exist); oracle/numpy_ref.py pins this file from the definition.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon

# ModelNet decoder structure, test_modelnet_VAE_dr.py:176-185 of the reference
MODELNET_DECODER = {
    'name': 'decoder',
    'input_dim': 64,
    'output_shape': [64, 64, 64, 1],
    'filter_num_list': [512, 256, 128, 64, 1],
    'filter_size_list': [4, 4, 4, 4, 4],
    'strides_list': [1, 2, 2, 2, 2],
    'activation': 'elu',
    'final_activation': 'sigmoid',
}
# Pascal3D decoder structure, test_pascal_VAE_dr.py:196-205 (latent dim 16)
PASCAL_DECODER = dict(MODELNET_DECODER, input_dim=16)


def parse_structure(structure: dict) -> dict:
    """Shape arithmetic of autoencoder3D.py:113-120, including its float quirks."""
    output_shape = structure['output_shape']
    strides = structure['strides_list']
    filters = structure['filter_num_list']
    # :115  list / np.int64 -> float array [4., 4., 4.]
    conv_input_dim_wo_ch = np.asarray(output_shape[:-1], dtype=np.float64) / np.prod(strides)
    # :116-118
    conv_input_ch = filters[0] / 64
    if conv_input_ch < 8:
        conv_input_ch = 8
    linear_output_dim = np.prod(conv_input_dim_wo_ch) * conv_input_ch  # :120
    return {
        'input_dim': int(structure['input_dim']),
        'grid0': [int(v) for v in conv_input_dim_wo_ch],
        'ch0': int(conv_input_ch),
        'dense_units': int(linear_output_dim),
        'filters': [int(f) for f in filters],
        'ksizes': [int(k) for k in structure['filter_size_list']],
        'strides': [int(s) for s in strides],
        'activation': structure['activation'],
        'final_activation': structure['final_activation'],
    }


def weight_shapes(structure: dict) -> list[tuple[str, tuple[int, ...]]]:
    """(name, shape) of the Keras variables in get_weights() order (27 arrays for the stock decoder)."""
    s = parse_structure(structure)
    out = [('dense/kernel', (s['input_dim'], s['dense_units'])), ('dense/bias', (s['dense_units'],))]
    for nm in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
        out.append((f'bn0/{nm}', (s['dense_units'],)))
    cin = s['ch0']
    nl = len(s['filters'])
    for i, (f, k) in enumerate(zip(s['filters'], s['ksizes'])):
        out.append((f'convT{i + 1}/kernel', (k, k, k, f, cin)))
        if i < nl - 1:
            for nm in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                out.append((f'bn{i + 1}/{nm}', (f,)))
        cin = f
    return out


def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round fp32 values to the nearest bf16-representable fp32 value (round-to-nearest-even)."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


def keras_default_weights(structure: dict, seed: int, bf16_kernels: bool = True) -> list[np.ndarray]:
    """Keras-default initialisation of every decoder variable (what a freshly built reference model holds)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ws = []
    for name, shape in weight_shapes(structure):
        if name.endswith('/kernel'):
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:
                rec = int(np.prod(shape[:-2]))
                fan_in, fan_out = shape[-2] * rec, shape[-1] * rec
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            ws.append(round_bf16(w) if bf16_kernels else w)
        elif name.endswith('/gamma') or name.endswith('/moving_variance'):
            ws.append(np.ones(shape, np.float32))
        else:
            ws.append(np.zeros(shape, np.float32))
    return ws


def _elu(x):
    return F.elu(x, alpha=1.0)


def conv3d_transpose_same(x: torch.Tensor, w_keras: torch.Tensor, stride: int) -> torch.Tensor:
    """Keras Conv3DTranspose(padding='same', use_bias=False) on NCDHW ``x`` with a Keras-layout kernel.

    autoencoder3D.py:42-45 / :129-132.  ``w_keras``: [kd,kh,kw,Cout,Cin].
    """
    k = w_keras.shape[0]
    w_t = w_keras.permute(4, 3, 0, 1, 2).contiguous()  # torch layout [Cin, Cout, kd, kh, kw]
    pad_total = max(k - stride, 0)
    pb = pad_total // 2
    full = F.conv_transpose3d(x, w_t, stride=stride)  # length (in-1)*s + k
    n_out = x.shape[2] * stride
    return full[:, :, pb:pb + n_out, pb:pb + n_out, pb:pb + n_out]


def decoder_forward(structure: dict, weights: list, z, dtype=torch.float32, return_layers: bool = False,
                    final_logits: bool = False):
    """``decoder(z, training=False)`` of the reference.  z: [B, D] -> [B, 64, 64, 64, 1] (NDHWC, like Keras).

    ``return_layers`` additionally returns the post-activation NDHWC tensor of every hidden layer
    (dense, convT1..convT4) and the final logits, for per-layer parity tests.
    """
    s = parse_structure(structure)
    ws = [torch.as_tensor(np.asarray(w), dtype=dtype) for w in weights]
    x = torch.as_tensor(np.asarray(z), dtype=dtype).reshape(-1, s['input_dim'])  # linearTransform :58
    it = iter(ws)
    layers = []

    def bn(x, ch_axis):
        g, b, m, v = next(it), next(it), next(it), next(it)
        shape = [1] * x.dim()
        shape[ch_axis] = -1
        return g.view(shape) * (x - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) + b.view(shape)

    act = {'elu': _elu, 'relu': F.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.3)}.get(s['activation'], lambda t: t)
    kern, bias = next(it), next(it)
    x = x @ kern + bias                       # Dense :59-61
    x = act(bn(x, 1))                         # BN + ELU :62-68
    g0 = s['grid0']
    x = x.reshape(-1, g0[0], g0[1], g0[2], s['ch0'])   # tf.reshape :125 (NDHWC)
    layers.append(x.clone())
    x = x.permute(0, 4, 1, 2, 3).contiguous()          # -> NCDHW for torch
    nl = len(s['filters'])
    for i in range(nl - 1):                   # conv3DDec :127-128
        x = conv3d_transpose_same(x, next(it), s['strides'][i])
        x = act(bn(x, 1))
        layers.append(x.permute(0, 2, 3, 4, 1).contiguous())
    x = conv3d_transpose_same(x, next(it), s['strides'][-1])   # final ConvT :129-132
    logits = x.permute(0, 2, 3, 4, 1).contiguous()
    layers.append(logits)
    out = logits if (final_logits or s['final_activation'] != 'sigmoid') else torch.sigmoid(logits)  # :134-136
    if return_layers:
        return out, layers
    return out


def trained_like_weights(structure: dict, seed: int, calib: int = 4, logit_std: float = 3.0,
                         logit_mean: float = -3.3, bf16_kernels: bool = True) -> list[np.ndarray]:
    """Deterministic "trained-like" weights: Glorot kernels, BN moving statistics calibrated on random
    latents (so every hidden layer is ~N(0,1) before ELU with randomised gamma/beta), and a final kernel
    scaled/offset so logits have std ~= ``logit_std`` and mean ~= ``logit_mean`` (occupancy ~ 10-15 %).
    Kernels are rounded to bf16-representable values; BN vectors stay fp32 (applied in fp32 on both sides).
    ``bf16_kernels=False`` keeps full fp32 kernels, i.e. what a trained reference checkpoint holds (the GPU then sees
    the 16-bit rounding of the operands while the oracle computes with the fp32 originals).
    """
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    ws = keras_default_weights(structure, seed, bf16_kernels=False)
    names = [n for n, _ in weight_shapes(structure)]
    s = parse_structure(structure)
    z = rng.standard_normal((calib, s['input_dim'])).astype(np.float32)
    # randomise gamma/beta, then set moving stats layer by layer from the calibration batch
    for i, n in enumerate(names):
        if n.endswith('/gamma'):
            ws[i] = rng.uniform(0.6, 1.4, ws[i].shape).astype(np.float32)
        elif n.endswith('/beta'):
            ws[i] = (0.25 * rng.standard_normal(ws[i].shape)).astype(np.float32)
        elif n == 'dense/bias':
            ws[i] = (0.05 * rng.standard_normal(ws[i].shape)).astype(np.float32)
        elif n.endswith('/kernel') and bf16_kernels:
            ws[i] = round_bf16(ws[i])
    bn_layers = [i for i, n in enumerate(names) if n.endswith('/moving_mean')]
    for li, mi in enumerate(bn_layers):
        # pre-BN statistics of this layer with the statistics fixed so far
        x = torch.from_numpy(z)
        it = iter([torch.from_numpy(w) for w in ws])
        kern, bias = next(it), next(it)
        pre = x @ kern + bias
        cur = 0
        while True:
            g, b, m, v = next(it), next(it), next(it), next(it)
            if cur == li:
                red = tuple(d for d in range(pre.dim()) if d != 1)
                mean = pre.mean(dim=red)
                var = pre.var(dim=red, unbiased=False)
                jit = torch.from_numpy(rng.uniform(0.8, 1.25, mean.shape).astype(np.float32))
                ws[mi] = (mean + 0.1 * var.sqrt() * torch.from_numpy(
                    rng.standard_normal(mean.shape).astype(np.float32))).numpy().astype(np.float32)
                ws[mi + 1] = (var * jit + 1e-6).numpy().astype(np.float32)
                break
            shape = [1, -1] + [1] * (pre.dim() - 2)
            h = g.view(shape) * (pre - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) + b.view(shape)
            h = _elu(h)
            if cur == 0:
                g0 = s['grid0']
                h = h.reshape(-1, g0[0], g0[1], g0[2], s['ch0']).permute(0, 4, 1, 2, 3).contiguous()
            pre = conv3d_transpose_same(h, next(it), s['strides'][cur])
            cur += 1
    # final kernel: scale + constant offset to hit the logit statistics
    k5 = len(ws) - 1
    base = ws[k5].copy()
    ws[k5] = base
    l0 = decoder_forward(structure, ws, z, final_logits=True)
    ones = [w.copy() for w in ws]
    ones[k5] = np.ones_like(base)
    s0 = decoder_forward(structure, ones, z, final_logits=True)
    a = logit_std / float(l0.std())
    # mean(a*l0 + b*s0) = logit_mean
    b = (logit_mean - a * float(l0.mean())) / float(s0.mean())
    # keep the offset from dominating the variance
    tot = a * l0 + b * s0
    a *= logit_std / float(tot.std())
    b = (logit_mean - a * float(l0.mean())) / float(s0.mean())
    ws[k5] = (a * base + b).astype(np.float32)
    if bf16_kernels:
        ws[k5] = round_bf16(ws[k5])
    return ws
