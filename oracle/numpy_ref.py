"""fp64 numpy restatement written from the DEFINITIONS (TEST INFRASTRUCTURE, see oracle/__init__.py).

Independent of torch's conv kernels: the transposed convolution is built as the literal adjoint of a
TensorFlow SAME-padded strided forward convolution, which is how ``tf.nn.conv3d_transpose`` /
``tf.keras.layers.Conv3DTranspose`` (called at /root/reference/src/net_core/autoencoder3D.py:42-45,129-132)
are defined.  Used to pin oracle/decoder_ref.py on small shapes; pure numpy einsum, no loops over voxels.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3


def same_forward_relation(big: int, k: int, s: int) -> np.ndarray:
    """R[o, t, j] = 1 iff the SAME-padded forward conv (input length ``big``, kernel k, stride s) reads input
    sample j with tap t to produce output o.  TF SAME rule: out = ceil(big/s),
    pad_total = max((out-1)*s + k - big, 0), pad_before = pad_total // 2."""
    out = -(-big // s)
    pad_total = max((out - 1) * s + k - big, 0)
    pb = pad_total // 2
    R = np.zeros((out, k, big), dtype=np.float64)
    for o in range(out):
        for t in range(k):
            j = o * s + t - pb
            if 0 <= j < big:
                R[o, t, j] = 1.0
    return R


def conv3d_transpose_same(x: np.ndarray, w: np.ndarray, stride: int) -> np.ndarray:
    """x: [N, d, h, w, Cin] (NDHWC), w: Keras kernel [k,k,k,Cout,Cin] -> [N, d*s, h*s, w*s, Cout] in fp64.

    Adjoint of  y[o, cin] = sum_{t, cout} xbig[o*s + t - pb, cout] * W[t, cout, cin]  per axis.
    """
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    k = w.shape[0]
    n, d, h, wd, _ = x.shape
    Rd = same_forward_relation(d * stride, k, stride)
    Rh = same_forward_relation(h * stride, k, stride)
    Rw = same_forward_relation(wd * stride, k, stride)
    assert Rd.shape[0] == d and Rh.shape[0] == h and Rw.shape[0] == wd
    out = np.zeros((n, d * stride, h * stride, wd * stride, w.shape[3]), np.float64)
    # y[n, D, H, W, co] = sum x[n,a,b,c,ci] Rd[a,p,D] Rh[b,q,H] Rw[c,r,W] w[p,q,r,co,ci], one tap (p,q,r) at a time:
    # the relation tensors give, per tap, the (input index, output index) pairs that the forward conv couples.
    for p in range(k):
        ad, jd = np.nonzero(Rd[:, p, :])
        for q in range(k):
            ah, jh = np.nonzero(Rh[:, q, :])
            for r in range(k):
                aw, jw = np.nonzero(Rw[:, r, :])
                if len(ad) == 0 or len(ah) == 0 or len(aw) == 0:
                    continue
                y = x[:, ad][:, :, ah][:, :, :, aw] @ w[p, q, r].T      # [n, |ad|, |ah|, |aw|, co]
                out[np.ix_(np.arange(n), jd, jh, jw, np.arange(w.shape[3]))] += y
    return out


def batchnorm(x, gamma, beta, mean, var):
    g, b, m, v = (np.asarray(a, np.float64) for a in (gamma, beta, mean, var))
    return g * (x - m) / np.sqrt(v + BN_EPS) + b


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def decoder_forward(weights: list, z: np.ndarray, strides, grid0=(4, 4, 4), ch0=8, return_layers=False):
    """fp64 decoder(z, training=False); same graph as decoder_ref.decoder_forward (autoencoder3D.py:104-139)."""
    ws = [np.asarray(w, np.float64) for w in weights]
    it = iter(ws)
    x = np.asarray(z, np.float64) @ next(it) + next(it)
    x = elu(batchnorm(x, next(it), next(it), next(it), next(it)))
    x = x.reshape(-1, grid0[0], grid0[1], grid0[2], ch0)
    layers = [x]
    for s in strides[:-1]:
        x = conv3d_transpose_same(x, next(it), s)
        x = elu(batchnorm(x, next(it), next(it), next(it), next(it)))
        layers.append(x)
    logits = conv3d_transpose_same(x, next(it), strides[-1])
    layers.append(logits)
    out = 1.0 / (1.0 + np.exp(-logits))
    return (out, layers) if return_layers else out


def voxel_precision_recall(x_target: np.ndarray, x_pred: np.ndarray, prob: float = 0.5):
    """function.py:100-115: yPred = (xPred >= prob); TP, FP, FN per object as float sums."""
    b = x_target.shape[0]
    yt = np.asarray(x_target, np.float32).reshape(b, -1)
    yp = (np.asarray(x_pred).reshape(b, -1) >= prob).astype(np.float32)
    tp = (yt * yp).sum(-1)
    fp = ((1.0 - yt) * yp).sum(-1)
    fn = (yt * (1.0 - yp)).sum(-1)
    return tp, fp, fn
