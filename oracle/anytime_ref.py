"""CPU restatement of the anytime imputation / K-mean / scoring steps (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference call sites restated here:

* ``sampling``                      /root/reference/src/module/function.py:35-38
* Bernoulli mask + mean fill        /root/reference/src/module/nolbo.py:1472-1486
* nearest-prior "corrected" fill    /root/reference/src/module/nolbo.py:1505-1510
* N(0,1) fill                       /root/reference/src/module/nolbo.py:431-439
* K-sample mean of sigmoid grids    /root/reference/src/module/nolbo_test.py:167-177
* ``voxelPrecisionRecall``          /root/reference/src/module/function.py:100-115

RNG contract (ours; the reference draws from unseeded np.random / tf.random, so sample-level parity with it is
undefined): Philox4x32-10 (Salmon et al., SC'11; Random123), key = (seed_lo, seed_hi),
counter = (dim // 4, sample k, object_id_lo, object_id_hi); the 4 output words give 4 normals for
dims 4q..4q+3 by Box-Muller on (w0,w1) and (w2,w3):  u = (w + 0.5) * 2^-32,
n_even = sqrt(-2 ln u0) * cos(2 pi u1), n_odd = sqrt(-2 ln u0) * sin(2 pi u1).
"""
from __future__ import annotations

import numpy as np
import torch

from . import decoder_ref

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
FILL_MODES = {'prior_sample': 0, 'mean': 1, 'normal': 2}


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Vectorised Philox4x32-10.  ctr: [..., 4] uint32, key: [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    for _ in range(10):
        p0 = c[0].astype(np.uint64) * PHILOX_M0
        p1 = c[2].astype(np.uint64) * PHILOX_M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over='ignore'):
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_words(seed: int, obj_ids: np.ndarray, K: int, D: int) -> np.ndarray:
    """uint32 words [B, K, D] per the counter contract above."""
    nq = (D + 3) // 4
    obj = np.asarray(obj_ids, dtype=np.uint64)
    B = obj.shape[0]
    ctr = np.zeros((B, K, nq, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(nq, dtype=np.uint32)[None, None, :]
    ctr[..., 1] = np.arange(K, dtype=np.uint32)[None, :, None]
    ctr[..., 2] = (obj & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None, None]
    ctr[..., 3] = (obj >> np.uint64(32)).astype(np.uint32)[:, None, None]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    w = philox4x32_10(ctr, key.reshape(1, 1, 1, 2))
    return w.reshape(B, K, nq * 4)[:, :, :D]


def philox_normals(seed: int, obj_ids: np.ndarray, K: int, D: int) -> np.ndarray:
    """Standard normals [B, K, D] (float64 evaluation of the fp32-defined uniforms)."""
    nq = (D + 3) // 4
    w = philox_words(seed, obj_ids, K, nq * 4).reshape(len(obj_ids), K, nq, 4)
    # the uniforms are defined in float32: fma(float(w), 2^-32, 2^-33)
    u = (w.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)
    u = u.astype(np.float64)
    out = np.empty(w.shape, dtype=np.float64)
    for a in (0, 2):
        r = np.sqrt(-2.0 * np.log(u[..., a]))
        out[..., a] = r * np.cos(2.0 * np.pi * u[..., a + 1])
        out[..., a + 1] = r * np.sin(2.0 * np.pi * u[..., a + 1])
    return out.reshape(len(obj_ids), K, nq * 4)[:, :, :D]


def sampling(mu: np.ndarray, log_var: np.ndarray, eps: np.ndarray) -> np.ndarray:
    """function.py:35-38 with the noise passed in: mu + sqrt(exp(logVar)) * eps."""
    return mu + np.sqrt(np.exp(log_var)) * eps


def bernoulli_mask(rng: np.random.Generator, B: int, D: int, missing_prob: float) -> np.ndarray:
    """nolbo.py:1475-1476: np.random.choice(2, B*D, p=[missing, 1-missing]) reshaped [B, D] float32 (1 = received)."""
    if missing_prob <= 0:
        return np.ones((B, D), np.float32)   # nolbo.py:1485-1486
    m = rng.choice(2, B * D, p=[missing_prob, 1.0 - missing_prob])
    return m.reshape(B, D).astype(np.float32)


def prefix_mask(B: int, D: int, length) -> np.ndarray:
    """Anytime arrival: the first ``length`` latent dims received (config 4 of BASELINE.json)."""
    length = np.broadcast_to(np.asarray(length), (B,))
    return (np.arange(D)[None, :] < length[:, None]).astype(np.float32)


def impute(z: np.ndarray, mask: np.ndarray, mu_table: np.ndarray, K: int, seed: int, obj_offset: int = 0,
           fill: str = 'prior_sample'):
    """Complete partially received latents.  Returns (z_out [B,K,D] float32, cstar [B] int32).

    mean          nolbo.py:1477-1482: z*mask, then where(z == 0) <- mean over categories of the prior means.
    prior_sample  nolbo.py:1505-1510: on top of the mean fill, c* = argmin_c sum_d mask*(z-mu_c)^2 and the
                  missing dims (mask == 0) get N(mu_c*, 1) draws.
    normal        nolbo.py:431-439:   where(z*mask == 0) <- N(0, 1).
    none          nolbo.py:1485-1486: the missing_prob == 0 branch: mask of ones, no fill at all (z * mask).
    """
    z = np.asarray(z, np.float32)
    mask = np.asarray(mask, np.float32)
    B, D = z.shape
    obj = np.arange(B, dtype=np.uint64) + np.uint64(obj_offset)
    zm = z * mask
    out = np.empty((B, K, D), np.float32)
    cstar = np.full((B,), -1, np.int32)
    if fill == 'none':
        out[:] = zm[:, None, :]
        return out, cstar
    if fill == 'normal':
        eps = philox_normals(seed, obj, K, D).astype(np.float32)
        out[:] = np.where(zm[:, None, :] == 0, eps, zm[:, None, :])
        return out, cstar
    mu = np.asarray(mu_table, np.float32)
    prior_mean = mu.astype(np.float64).mean(axis=0).astype(np.float32)  # reduce_mean over categories :1473
    zf = np.where(zm == 0, prior_mean[None, :], zm).astype(np.float32)
    dist = (mask[:, None, :].astype(np.float64) * (zf[:, None, :].astype(np.float64) - mu[None].astype(np.float64)) ** 2).sum(-1)
    cstar = dist.argmin(-1).astype(np.int32)
    if fill == 'mean':
        out[:] = zf[:, None, :]
        return out, cstar
    assert fill == 'prior_sample'
    eps = philox_normals(seed, obj, K, D)
    prior = (mu[cstar][:, None, :].astype(np.float64) + eps).astype(np.float32)   # sampling(mu_c*, logVar=0)
    out[:] = np.where(mask[:, None, :] == 0, prior, zf[:, None, :])
    return out, cstar


def nearest_prior(z: np.ndarray, mu_table: np.ndarray, category_list: np.ndarray | None = None):
    """nolbo.py:1488-1494: argmin_c sum_d (z - mu_c)^2 over all dims; acc_cat = mean(argmin == argmax(category_list))."""
    z = np.asarray(z, np.float64)
    mu = np.asarray(mu_table, np.float64)
    idx = ((z[:, None, :] - mu[None]) ** 2).sum(-1).argmin(-1).astype(np.int32)
    acc = None if category_list is None else float((idx == np.asarray(category_list).argmax(-1)).mean())
    return idx, acc


def make_targets(rng: np.random.Generator, B: int, G: int = 64) -> np.ndarray:
    """Synthetic voxel targets: union of 1-3 random axis-aligned ellipsoids, occupancy ~5-20 %.
    Layout of the reference loaders: float32 {0,1} [B, 64, 64, 64, 1] (pascal3D.py:149-152)."""
    ax = np.arange(G, dtype=np.float32)
    out = np.zeros((B, G, G, G, 1), np.float32)
    for b in range(B):
        occ = np.zeros((G, G, G), bool)
        for _ in range(int(rng.integers(1, 4))):
            c = rng.uniform(0.3 * G, 0.7 * G, 3)
            r = rng.uniform(0.12 * G, 0.3 * G, 3)
            d = ((ax[:, None, None] - c[0]) / r[0]) ** 2 + ((ax[None, :, None] - c[1]) / r[1]) ** 2 + \
                ((ax[None, None, :] - c[2]) / r[2]) ** 2
            occ |= d <= 1.0
        out[b, ..., 0] = occ
    return out


def pack_bits(grid01: np.ndarray) -> np.ndarray:
    """[B, ...V] {0,1} -> [B, V/8] uint8, voxel v in bit (v % 8) of byte v // 8 (little-endian bit order)."""
    b = grid01.shape[0]
    return np.packbits(np.asarray(grid01).reshape(b, -1).astype(np.uint8), axis=1, bitorder='little')


def counts(x_target: np.ndarray, x_pred: np.ndarray, prob: float = 0.5) -> np.ndarray:
    """voxelPrecisionRecall (function.py:100-115) as exact integers: [B, 3] int64 = TP, FP, FN (>= threshold)."""
    b = x_target.shape[0]
    yt = np.asarray(x_target).reshape(b, -1) > 0.5
    yp = np.asarray(x_pred).reshape(b, -1) >= prob
    tp = (yt & yp).sum(-1)
    fp = (~yt & yp).sum(-1)
    fn = (yt & ~yp).sum(-1)
    return np.stack([tp, fp, fn], -1).astype(np.int64)


def binary_loss(x_pred: np.ndarray, x_target: np.ndarray, gamma: float = 0.5, epsilon: float = 1e-7) -> np.ndarray:
    """function.py:73-82 with b_range = False: -sum_v gamma*y*log(p) + (1-gamma)*(1-y)*log(1-p), p clipped in fp32
    like tf.clip_by_value; the sum is evaluated in fp64.  Returns [B]."""
    b = x_pred.shape[0]
    p = np.clip(np.asarray(x_pred, np.float32).reshape(b, -1), np.float32(epsilon), np.float32(1.0) - np.float32(epsilon))
    y = np.asarray(x_target, np.float64).reshape(b, -1)
    p = p.astype(np.float64)
    return -(gamma * y * np.log(p) + (1.0 - gamma) * (1.0 - y) * np.log(1.0 - p)).sum(-1)


def counts_sweep(x_target: np.ndarray, x_pred: np.ndarray, thresholds, strict: bool = True) -> np.ndarray:
    """modelnetAE3.ipynb cell 2: yPred = (xPred > prob) per threshold -> [B, T, 3] int64."""
    b = x_target.shape[0]
    yt = np.asarray(x_target).reshape(b, -1) > 0.5
    xp = np.asarray(x_pred, np.float32).reshape(b, -1)
    out = np.zeros((b, len(thresholds), 3), np.int64)
    for i, th in enumerate(np.asarray(thresholds, np.float32)):
        yp = xp > th if strict else xp >= th
        out[:, i, 0] = (yt & yp).sum(-1)
        out[:, i, 1] = (~yt & yp).sum(-1)
        out[:, i, 2] = (yt & ~yp).sum(-1)
    return out


def iou_from_counts(c: np.ndarray):
    """IoU = TP / (TP+FP+FN) (derived; the reference reports pr/rc from the same counts, nolbo.py:1499-1501).
    Returns (mean over objects of per-object IoU, global sum-then-ratio IoU)."""
    c = np.asarray(c, np.float64)
    den = c.sum(-1)
    per = np.where(den > 0, c[:, 0] / np.maximum(den, 1), 1.0)
    return float(per.mean()), float(c[:, 0].sum() / max(c.sum(), 1.0))


def precision_recall(c: np.ndarray):
    """nolbo.py:1499-1501: pr = mean_b TP/(TP+FP+1e-10), rc = mean_b TP/(TP+FN+1e-10)."""
    c = np.asarray(c, np.float64)
    pr = (c[:, 0] / (c[:, 0] + c[:, 1] + 1e-10)).mean()
    rc = (c[:, 0] / (c[:, 0] + c[:, 2] + 1e-10)).mean()
    return float(pr), float(rc)


def anytime_eval(structure: dict, weights: list, z_bkd: np.ndarray, targets: np.ndarray, threshold: float = 0.5,
                 batch: int = 16):
    """K-sample anytime reconstruction of completed latents: decode all B*K, mean of the K post-sigmoid grids
    (nolbo_test.py:176), threshold >= and TP/FP/FN.  Returns (mean_prob [B,64,64,64,1] f32, counts [B,3] i64)."""
    z_bkd = np.asarray(z_bkd, np.float32)
    B, K, D = z_bkd.shape
    flat = z_bkd.reshape(B * K, D)
    outs = []
    with torch.no_grad():
        for i in range(0, B * K, batch):
            outs.append(decoder_ref.decoder_forward(structure, weights, flat[i:i + batch]).numpy())
    p = np.concatenate(outs, 0).reshape(B, K, *outs[0].shape[1:])
    mean_prob = p.astype(np.float32).mean(axis=1, dtype=np.float32)  # tf.reduce_mean(axis=0) in fp32
    return mean_prob, counts(targets, mean_prob, threshold)
