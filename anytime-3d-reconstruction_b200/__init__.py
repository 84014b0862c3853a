"""a3d -- B200-native anytime voxel-decoder hot path (drop-in for the reference's net_core decoder path).

Host-side mirror of the reference interface for this path (same names, argument meaning, error behaviour):

* ``decoder3D(structure)``                 <- src/net_core/autoencoder3D.py:104-139 (returns a callable model)
* ``model(z, training=False)``             <- call sites src/module/nolbo.py:1496,1520; nolbo_test.py:176,180
* ``model.set_weights / get_weights / load_weights / save_weights``  <- nolbo.py:1568-1592 (Keras variable order)
* ``sampling(mu, logVar)``                 <- src/module/function.py:35-38
* ``voxelPrecisionRecall(xTarget, xPred, prob)``  <- src/module/function.py:100-115
* ``Darknet19(...)`` / ``head2D(...)``     <- src/net_core/darknet.py:96-133,149-168 (image encoder, encoder2d.py)
* ``encoder3D(structure)``                 <- src/net_core/autoencoder3D.py:72-102 (voxel encoder, encoder3d.py)
* ``anytime_eval(...)`` / ``getEval(...)`` <- the imputation + decode + score sequence of nolbo.py:1449-1528 with the
  K-sample mean of nolbo_test.py:167-177

Everything numerical runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/a3d.h``
(``liba3d.so``, loaded with ctypes).  PyTorch is used only for device memory, streams and torch.distributed.
There is no CPU / PyTorch fallback: calls raise ``RuntimeError`` if the extension or a B200 is missing.

The directory name contains '-', so import it as ``import a3d`` (alias module at the repo root) or with
``importlib.import_module('anytime-3d-reconstruction_b200')``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi, presets
from ._capi import FILL, OUT, VOXELS
from .encoder2d import Darknet19, Encoder2D, head2D, image_encoder
from .encoder3d import Encoder3D, encoder3D

__all__ = ['decoder3D', 'Decoder3D', 'sampling', 'voxelPrecisionRecall', 'voxelPrecisionRecallSweep', 'binary_loss',
           'anytime_eval', 'anytime_eval_host', 'impute', 'getEval', 'pack_targets', 'iou_from_counts', 'shard_range',
           'nearest_prior', 'pinned_empty', 'getPredShapes',
           'allreduce_counts', 'Darknet19', 'head2D', 'image_encoder', 'Encoder2D', 'getEvalImages', 'encoder3D', 'Encoder3D', 'getEvalVoxels']


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError('a3d needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch


def _stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _as_dev_f32(x, torch, device):
    """numpy / torch (cpu or cuda) -> contiguous fp32 CUDA tensor on ``device`` (plumbing only)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)


def _parse_structure(structure: dict) -> dict:
    """Shape arithmetic of decoder3D, autoencoder3D.py:105-120 (including the list / np.int64 float quirk)."""
    for key in ('name', 'input_dim', 'output_shape', 'filter_num_list', 'filter_size_list', 'strides_list',
                'activation', 'final_activation'):
        if key not in structure:
            raise KeyError(key)   # the reference indexes the dict directly and raises KeyError too
    out_shape = structure['output_shape']
    strides = list(structure['strides_list'])
    filters = list(structure['filter_num_list'])
    ksizes = list(structure['filter_size_list'])
    grid0 = np.asarray(out_shape[:-1], dtype=np.float64) / np.prod(strides)
    ch0 = filters[0] / 64
    if ch0 < 8:
        ch0 = 8
    if not (len(filters) == len(ksizes) == len(strides)):
        raise ValueError('filter_num_list, filter_size_list and strides_list must have the same length')
    return dict(name=structure['name'], input_dim=int(structure['input_dim']), out_grid=int(out_shape[0]),
                out_shape=[int(v) for v in out_shape], grid0=[int(g) for g in grid0], ch0=int(ch0),
                dense_units=int(np.prod(grid0) * ch0), filters=[int(f) for f in filters],
                ksizes=[int(k) for k in ksizes], strides=[int(s) for s in strides],
                activation=structure['activation'], final_activation=structure['final_activation'])


class Decoder3D:
    """Callable stand-in for the ``tf.keras.Model`` returned by the reference's ``decoder3D(structure)``."""

    def __init__(self, structure: dict, max_chunk: int = 256, operand_dtype: str = 'fp16', impl: str = 'tcgen05',
                 device: int | None = None):
        self.structure = dict(structure)
        self._s = _parse_structure(structure)
        self.name = self._s['name']
        s = self._s
        if s['activation'] not in _capi.ACT or s['final_activation'] not in _capi.FINAL:
            raise ValueError(f"unsupported activation {s['activation']!r} / {s['final_activation']!r}")
        torch = _require_cuda()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)
        d = _capi.Desc()
        d.abi_version = _capi.A3D_ABI_VERSION
        d.latent_dim = s['input_dim']
        d.num_layers = len(s['filters'])
        if d.num_layers > _capi.A3D_MAX_LAYERS:
            raise ValueError('too many layers')
        for i in range(d.num_layers):
            d.filters[i], d.ksizes[i], d.strides[i] = s['filters'][i], s['ksizes'][i], s['strides'][i]
        d.out_grid = s['out_grid']
        d.activation = _capi.ACT[s['activation']]
        d.final_activation = _capi.FINAL[s['final_activation']]
        d.device = self.device_index
        d.max_chunk = int(max_chunk)
        d.operand_dtype = _capi.DTYPE[operand_dtype]
        d.impl = _capi.IMPL[impl]
        self.max_chunk = int(max_chunk)
        self.operand_dtype = operand_dtype
        self.impl = impl
        self._lib = _capi.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_create(C.byref(d), C.byref(h)), 'a3d_create')
        self._h = h
        self.input_dim = s['input_dim']

    # ---- lifetime
    def close(self):
        if getattr(self, '_h', None):
            self._lib.a3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights (Keras get_weights()/set_weights() order and layouts)
    @property
    def num_weights(self) -> int:
        return int(self._lib.a3d_num_weights(self._h))

    def weight_shapes(self) -> list[tuple[int, ...]]:
        s = self._s
        shapes = [(s['input_dim'], s['dense_units']), (s['dense_units'],)] + [(s['dense_units'],)] * 4
        cin = s['ch0']
        n = len(s['filters'])
        for i, (f, k) in enumerate(zip(s['filters'], s['ksizes'])):
            shapes.append((k, k, k, f, cin))
            if i < n - 1:
                shapes += [(f,)] * 4
            cin = f
        return shapes

    def set_weights(self, weights) -> None:
        weights = list(weights)
        shapes = self.weight_shapes()
        if len(weights) != len(shapes):
            raise ValueError(f'You called `set_weights(weights)` with a weight list of length {len(weights)}, '
                             f'but the layer was expecting {len(shapes)} weights.')
        for i, (w, shp) in enumerate(zip(weights, shapes)):
            a = np.ascontiguousarray(np.asarray(w), dtype=np.float32)
            if tuple(a.shape) != tuple(shp):
                raise ValueError(f'Layer weight shape {tuple(shp)} not compatible with provided weight shape '
                                 f'{tuple(a.shape)} (variable {i})')
            _capi.check(self._lib.a3d_set_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes), 'a3d_set_weight')

    def get_weights(self) -> list[np.ndarray]:
        out = []
        for i, shp in enumerate(self.weight_shapes()):
            a = np.empty(shp, np.float32)
            _capi.check(self._lib.a3d_get_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes), 'a3d_get_weight')
            out.append(a)
        return out

    def _layer_var_names(self) -> list[list[str]]:
        """Variable names per weighted Keras layer in get_weights() order (autoencoder3D.py:122-132)."""
        bn = ['gamma', 'beta', 'moving_mean', 'moving_variance']
        names = [['kernel', 'bias'], bn]
        n = len(self._s['filters'])
        for i in range(n):
            names.append(['kernel'])
            if i < n - 1:
                names.append(bn)
        return names

    def save_weights(self, path: str, save_format: str | None = None) -> None:
        """``model.save_weights(path)`` as the reference calls it (nolbo.py:1572-1574, a bare prefix): like Keras, a path
        without a recognised suffix is written as a TensorFlow tensor-bundle checkpoint (``path.index`` +
        ``path.data-00000-of-00001`` with the Keras object graph, tf_checkpoint.py); a ``.npz`` path (or
        ``save_format='npz'``) stores the Keras-order arrays in a numpy archive."""
        from . import tf_checkpoint
        if tf_checkpoint.wants_tf_format(path, save_format):
            tf_checkpoint.save_keras_weights(path, self.get_weights(), self._layer_var_names())
            return
        np.savez(path if path.endswith('.npz') else path + '.npz', *self.get_weights())

    def load_weights(self, path: str) -> None:
        """nolbo.py:1585-1592: accepts the checkpoint prefix Keras ``save_weights`` wrote (TF tensor bundle, read without
        TensorFlow by tf_checkpoint.py) or an .npz of Keras-order arrays."""
        from . import tf_checkpoint
        if tf_checkpoint.is_checkpoint(path):
            self.set_weights(tf_checkpoint.load_keras_weights(path))
            return
        p = path if os.path.exists(path) else path + '.npz'
        with np.load(p) as f:
            self.set_weights([f[f'arr_{i}'] for i in range(len(f.files))])

    # ---- forward
    def __call__(self, latents, training: bool = False, out=None, out_dtype: str = 'f32', threshold: float = 0.5):
        """decoder(z, training=False) -> [B, 64, 64, 64, 1] float32 (NDHWC).  numpy in -> numpy out (through the
        host-buffer entry a3d_decode_host: device->host copies overlapped with the decode); torch in -> CUDA torch
        tensor out (a3d_decode on the current stream).

        Extras for the numpy path (not in the reference): ``out`` = a preallocated result array, ideally from
        ``a3d.pinned_empty`` (page-locked memory: the copy then runs at the PCIe rate; a fresh pageable array is
        allocated otherwise); ``out_dtype`` = 'f32' (the Keras output), 'f16' (half the bytes) or 'bits'
        ([B, 32768] uint8, bit = p >= threshold, packed like the targets)."""
        if training:
            raise NotImplementedError('a3d implements the inference path only (training=False)')
        torch = _torch()
        if not isinstance(latents, torch.Tensor):
            return self.decode_host(latents, out=out, out_dtype=out_dtype, threshold=threshold)
        z = _as_dev_f32(latents, torch, self.device).reshape(-1, self.input_dim)
        n = z.shape[0]
        res = torch.empty((n, 64, 64, 64, 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_decode(self._h, z.data_ptr(), n, res.data_ptr(), _stream_ptr(torch)),
                        'a3d_decode')
        return res

    predict = __call__

    def decode_host(self, latents, out=None, out_dtype: str = 'f32', threshold: float = 0.5) -> np.ndarray:
        """numpy latents -> numpy grid through a3d_decode_host (see __call__)."""
        _require_cuda()
        z = np.ascontiguousarray(np.asarray(latents), dtype=np.float32).reshape(-1, self.input_dim)
        n = z.shape[0]
        code = OUT[out_dtype]
        shape, dt = (((n, 64, 64, 64, 1), np.float32), ((n, 64, 64, 64, 1), np.float16), ((n, VOXELS // 8), np.uint8))[code]
        if out is None:
            out = np.empty(shape, dt)
        elif out.dtype != dt or out.size != int(np.prod(shape)) or not out.flags['C_CONTIGUOUS']:
            raise ValueError(f'out must be a C-contiguous {np.dtype(dt).name} array with {int(np.prod(shape))} elements')
        torch = _torch()
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_decode_host(self._h, z.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p),
                                                  code, float(threshold)), 'a3d_decode_host')
        return out.reshape(shape)

    # ---- diagnostics
    def debug_layer(self, layer: int, n: int) -> np.ndarray:
        s = self._s
        grids = [s['grid0'][0]]
        chans = [s['ch0']]
        for f, st in zip(s['filters'][:-1], s['strides'][:-1]):
            grids.append(grids[-1] * st)
            chans.append(f)
        g, c = grids[layer], chans[layer]
        a = np.empty((n, g, g, g, c), np.float32)
        _capi.check(self._lib.a3d_debug_read_layer(self._h, layer, n, a.ctypes.data_as(C.c_void_p), a.nbytes),
                    'a3d_debug_read_layer')
        return a

    def set_profiling(self, on: bool) -> None:
        self._lib.a3d_set_profiling(self._h, int(on))

    def stage_times_ms(self) -> dict:
        buf = (C.c_float * 5)()
        n = self._lib.a3d_stage_times_ms(self._h, buf, 5)
        return dict(zip(['dense_l1', 'l2', 'l3', 'l4', 'tail'][:n], [float(v) for v in buf[:n]]))

    @property
    def launch_count(self) -> int:
        return int(self._lib.a3d_launch_count(self._h))

    def workspace_bytes(self, n: int = 0) -> int:
        return int(self._lib.a3d_workspace_bytes(self._h, n))


def decoder3D(structure: dict, **kw) -> Decoder3D:
    """Same call as the reference's ``src.net_core.autoencoder3D.decoder3D(structure)`` (autoencoder3D.py:104).
    Keyword extras (not in the reference): max_chunk, operand_dtype ('fp16' | 'bf16'), impl, device."""
    return Decoder3D(structure, **kw)


# ------------------------------------------------------------------------------------------------ free functions
def impute(decoder: Decoder3D, z, mask, category_vectors, K: int = 1, seed: int = 0, obj_offset: int = 0,
           fill: str = 'prior_sample'):
    """Complete partially received latents on the GPU (nolbo.py:1472-1486,1505-1510,431-439).
    Returns (z_out [B,K,D] CUDA tensor, cstar [B] int32 CUDA tensor)."""
    torch = _require_cuda()
    dev = decoder.device
    z = _as_dev_f32(z, torch, dev)
    mask = _as_dev_f32(mask, torch, dev)
    B, D = z.shape
    mu = None if category_vectors is None else _as_dev_f32(category_vectors, torch, dev)
    Cn = 0 if mu is None else mu.shape[0]
    out = torch.empty((B, K, D), dtype=torch.float32, device=dev)
    cstar = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(decoder.device_index):
        _capi.check(decoder._lib.a3d_impute(decoder._h, z.data_ptr(), mask.data_ptr(),
                                            0 if mu is None else mu.data_ptr(), Cn, B, K, seed, obj_offset, FILL[fill],
                                            out.data_ptr(), cstar.data_ptr(), _stream_ptr(torch)), 'a3d_impute')
    return out, cstar


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory (for the ``out=`` / host-buffer calls: copies then run at the PCIe rate).
    The memory is owned by a torch tensor kept alive by the array's base."""
    torch = _require_cuda()
    t = torch.empty(tuple(int(v) for v in np.atleast_1d(shape)), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t.numpy()


def sampling(mu, logVar, seed: int | None = None, decoder: Decoder3D | None = None, obj_offset: int = 0):
    """function.py:35-38: mu + sqrt(exp(logVar)) * eps with eps ~ N(0,1) (Philox4x32-10 + Box-Muller on the GPU,
    a3d_sampling: one kernel, no handle).  The reference is unseeded; pass ``seed`` for reproducible draws (row b uses
    the k = 0 stream of object ``obj_offset + b`` of the imputation sampler).  ``decoder`` only selects the device."""
    torch = _require_cuda()
    is_np = not isinstance(mu, torch.Tensor)
    dev = torch.device('cuda', torch.cuda.current_device()) if decoder is None else decoder.device
    mu_t = _as_dev_f32(mu, torch, dev)
    lv_t = _as_dev_f32(logVar, torch, dev)
    if mu_t.shape != lv_t.shape:
        raise ValueError('mu and logVar must have the same shape')
    D = int(mu_t.shape[-1])
    n = mu_t.numel() // max(D, 1)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), 'little')
    out = torch.empty_like(mu_t)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_sampling(mu_t.data_ptr(), lv_t.data_ptr(), n, D, seed, obj_offset, out.data_ptr(),
                                             _stream_ptr(torch)), 'a3d_sampling')
    return out.cpu().numpy() if is_np else out


def nearest_prior(z, category_vectors, category_list=None, device=None):
    """Nearest-prior classification of getEval (nolbo.py:1488-1494): argmin_c ||z_b - mu_c||^2 over all dims and, with
    the one-hot ``category_list``, acc_cat = mean(argmin == argmax(category_list)).  Returns (idx int32 CUDA tensor [B],
    acc float or None).  ``z`` may be [B, D] or [B, K, D] (the first of the K completed latents is classified)."""
    torch = _require_cuda()
    dev = z.device if isinstance(z, torch.Tensor) and z.is_cuda else (
        torch.device('cuda', torch.cuda.current_device()) if device is None else device)
    z_t = _as_dev_f32(z, torch, dev)
    B, D = z_t.shape[0], z_t.shape[-1]
    stride = z_t.numel() // max(B, 1)
    mu = _as_dev_f32(category_vectors, torch, dev)
    lab = None if category_list is None else _as_dev_f32(category_list, torch, dev)
    idx = torch.empty((B,), dtype=torch.int32, device=dev)
    hits = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_nearest_prior(z_t.data_ptr(), stride, mu.data_ptr(), mu.shape[0], D,
                                                  0 if lab is None else lab.data_ptr(), B, idx.data_ptr(),
                                                  0 if lab is None else hits.data_ptr(), _stream_ptr(torch)),
                    'a3d_nearest_prior')
    return idx, (None if lab is None else hits.item() / max(B, 1))


def pack_targets(decoder: Decoder3D | None, targets):
    """fp32 {0,1} voxel targets [B,64,64,64,1] (loader layout) -> bit-packed [B, 32768] uint8 CUDA tensor."""
    torch = _require_cuda()
    dev, hnd = _dev_and_handle(decoder, torch)
    t = _as_dev_f32(targets, torch, dev)
    B = t.shape[0]
    V = t.numel() // B
    bits = torch.empty((B, V // 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_pack_targets(hnd, t.data_ptr(), B, V, bits.data_ptr(), _stream_ptr(torch)),
                    'a3d_pack_targets')
    return bits


def voxelPrecisionRecall(xTarget, xPred, prob: float = 0.5, decoder: Decoder3D | None = None):
    """function.py:100-115.  Returns (TP, FP, FN), each [B] float32 like the reference's float sums."""
    torch = _require_cuda()
    is_np = not isinstance(xPred, torch.Tensor)
    dev, hnd = _dev_and_handle(decoder, torch)
    t = _as_dev_f32(xTarget, torch, dev)
    p = _as_dev_f32(xPred, torch, dev)
    B = p.shape[0]
    V = p.numel() // max(B, 1)
    if t.numel() != p.numel():
        raise ValueError('xTarget and xPred must have the same number of voxels')
    cnt = torch.empty((B, 3), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_counts(hnd, t.data_ptr(), p.data_ptr(), B, V, float(prob), cnt.data_ptr(),
                                            _stream_ptr(torch)), 'a3d_counts')
    f = cnt.to(torch.float32)
    tp, fp, fn = f[:, 0], f[:, 1], f[:, 2]
    if is_np:
        return tp.cpu().numpy(), fp.cpu().numpy(), fn.cpu().numpy()
    return tp, fp, fn


def _dev_and_handle(decoder, torch):
    """The stand-alone scoring helpers need no decoder handle (the C entry points accept NULL)."""
    if decoder is None:
        return torch.device('cuda', torch.cuda.current_device()), None
    return decoder.device, decoder._h


def voxelPrecisionRecallSweep(xTarget, xPred, thresholds, strict: bool = True, decoder: Decoder3D | None = None):
    """Threshold sweep of the evaluation notebooks (modelnetAE3.ipynb cell 2: ``yPred > prob`` for a list of
    thresholds).  Returns int64 counts [B, T, 3] (TP, FP, FN) as a CUDA tensor (numpy if the inputs were numpy)."""
    torch = _require_cuda()
    is_np = not isinstance(xPred, torch.Tensor)
    dev, hnd = _dev_and_handle(decoder, torch)
    t = _as_dev_f32(xTarget, torch, dev)
    p = _as_dev_f32(xPred, torch, dev)
    B = p.shape[0]
    V = p.numel() // max(B, 1)
    thr = np.ascontiguousarray(thresholds, np.float32).reshape(-1)
    cnt = torch.empty((B, len(thr), 3), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_counts_sweep(hnd, t.data_ptr(), p.data_ptr(), B, V,
                                                  thr.ctypes.data_as(C.c_void_p), len(thr), int(strict), cnt.data_ptr(),
                                                  _stream_ptr(torch)), 'a3d_counts_sweep')
    return cnt.cpu().numpy() if is_np else cnt


def binary_loss(xPred, xTarget, epsilon: float = 1e-7, gamma: float = 0.5, b_range: bool = False,
                decoder: Decoder3D | None = None):
    """function.py:73-82 (same argument order and defaults): weighted BCE summed over voxels, one value per object.
    ``epsilon`` is fixed at 1e-7 and ``b_range`` at False (what every call site of the reference uses)."""
    if b_range or abs(epsilon - 1e-7) > 1e-12:
        raise NotImplementedError('binary_loss: only epsilon=1e-7, b_range=False (the reference call sites) are built')
    torch = _require_cuda()
    is_np = not isinstance(xPred, torch.Tensor)
    dev, hnd = _dev_and_handle(decoder, torch)
    p = _as_dev_f32(xPred, torch, dev)
    t = _as_dev_f32(xTarget, torch, dev)
    B = p.shape[0]
    V = p.numel() // max(B, 1)
    loss = torch.empty((B,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().a3d_binary_loss(hnd, p.data_ptr(), t.data_ptr(), B, V, float(gamma),
                                                 loss.data_ptr(), _stream_ptr(torch)), 'a3d_binary_loss')
    out = loss.to(torch.float32)
    return out.cpu().numpy() if is_np else out


def iou_from_counts(counts):
    """IoU = TP / (TP + FP + FN): (mean over objects, global sum-then-ratio)."""
    c = np.asarray(counts, dtype=np.float64).reshape(-1, 3)
    den = c.sum(-1)
    per = np.where(den > 0, c[:, 0] / np.maximum(den, 1), 1.0)
    return float(per.mean()) if len(per) else 0.0, float(c[:, 0].sum() / max(c.sum(), 1.0))


def anytime_eval(decoder: Decoder3D, z, mask, category_vectors, targets, K: int = 16, seed: int = 0,
                 fill: str = 'prior_sample', threshold: float = 0.5, return_grid: bool = False, obj_offset: int = 0,
                 z_completed=None, return_loss: bool = False, gamma: float = 0.6):
    """Anytime reconstruction of a batch of partially received latents, all on the GPU:
    K-sample imputation -> decoder -> mean of the K occupancy grids -> threshold -> TP/FP/FN.

    z, mask: [B, D]; category_vectors: [C, D]; targets: fp32 {0,1} [B,64,64,64,1] or bit-packed uint8 [B,32768].
    Returns dict(counts [B,3] int64 CUDA tensor, z_completed [B,K,D], cstar, mean_prob (if return_grid),
    loss [B] float64 = weighted BCE of function.py:73-82 with ``gamma`` (if return_loss)).
    ``z_completed`` (already imputed [B,K,D]) skips the imputation step."""
    torch = _require_cuda()
    dev = decoder.device
    if z_completed is None:
        zc, cstar = impute(decoder, z, mask, category_vectors, K=K, seed=seed, obj_offset=obj_offset, fill=fill)
    else:
        zc = _as_dev_f32(z_completed, torch, dev)
        cstar = None
        K = zc.shape[1]
    B = zc.shape[0]
    if isinstance(targets, torch.Tensor) and targets.dtype == torch.uint8:
        bits = targets.to(dev).contiguous()
    elif isinstance(targets, np.ndarray) and targets.dtype == np.uint8:
        bits = torch.from_numpy(np.ascontiguousarray(targets)).to(dev)
    else:
        bits = pack_targets(decoder, targets)
    counts = torch.empty((B, 3), dtype=torch.int64, device=dev)
    grid = torch.empty((B, 64, 64, 64, 1), dtype=torch.float32, device=dev) if return_grid else None
    loss = torch.empty((B,), dtype=torch.float64, device=dev) if return_loss else None
    with torch.cuda.device(decoder.device_index):
        if return_loss:
            _capi.check(decoder._lib.a3d_anytime_eval_loss(
                decoder._h, zc.data_ptr(), B, K, bits.data_ptr(), float(threshold), float(gamma), counts.data_ptr(),
                loss.data_ptr(), 0 if grid is None else grid.data_ptr(), _stream_ptr(torch)), 'a3d_anytime_eval_loss')
        else:
            _capi.check(decoder._lib.a3d_anytime_eval(decoder._h, zc.data_ptr(), B, K, bits.data_ptr(),
                                                      float(threshold), counts.data_ptr(),
                                                      0 if grid is None else grid.data_ptr(), _stream_ptr(torch)),
                        'a3d_anytime_eval')
    out = {'counts': counts, 'z_completed': zc, 'cstar': cstar}
    if return_loss:
        out['loss'] = loss
    if return_grid:
        out['mean_prob'] = grid
    return out


def anytime_eval_host(decoder: Decoder3D, z, mask, category_vectors, target_bits, K: int = 16, seed: int = 0,
                      fill: str = 'prior_sample', threshold: float = 0.5, obj_offset: int = 0, return_grid: bool = False):
    """Same as anytime_eval but through the host-buffer C-ABI entry (a3d_anytime_eval_host): numpy in, numpy out,
    host<->device copies inside the call.  This is the call bench.py times for its ``e2e`` number."""
    z = np.ascontiguousarray(z, np.float32)
    mask = np.ascontiguousarray(mask, np.float32)
    mu = None if category_vectors is None else np.ascontiguousarray(category_vectors, np.float32)
    bits = np.ascontiguousarray(target_bits, np.uint8)
    B, D = z.shape
    counts = np.empty((B, 3), np.int64)
    grid = np.empty((B, 64, 64, 64, 1), np.float32) if return_grid else None
    torch = _require_cuda()
    with torch.cuda.device(decoder.device_index):
        _capi.check(decoder._lib.a3d_anytime_eval_host(
            decoder._h, z.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p),
            None if mu is None else mu.ctypes.data_as(C.c_void_p), 0 if mu is None else mu.shape[0], B, K, seed,
            obj_offset, FILL[fill], bits.ctypes.data_as(C.c_void_p), float(threshold),
            counts.ctypes.data_as(C.c_void_p), None if grid is None else grid.ctypes.data_as(C.c_void_p)),
            'a3d_anytime_eval_host')
    return (counts, grid) if return_grid else counts


def getEval(decoder: Decoder3D, inputs, category_vectors, training: bool = False, missing_prob: float = 0.0,
            K: int = 1, seed: int = 0, mask=None, rng: np.random.Generator | None = None):
    """Latent-space equivalent of ``nolboSingleObject_modelnet_category_VAE.getEval`` (nolbo.py:1449-1528).

    ``inputs = (z, output_images, category_list)`` where ``z`` are the encoder latents (the reference samples them
    from the encoder at :1463-1470; the encoder is outside this path).  Returns the reference's 10-tuple
    ``(pred, loss_shape, pr, rc, acc_cat, pred_corr, loss_corr, pr_corr, rc_corr, acc_cat_corr)``; ``loss_shape`` is the
    batch mean of binary_loss(gamma=0.60) (nolbo.py:1497-1498), fused into the tail kernel; the corrected branch is
    K-sample averaged when K > 1."""
    if training:
        raise NotImplementedError('inference only')
    torch = _require_cuda()
    z, output_images, category_list = inputs
    z_t = _as_dev_f32(z, torch, decoder.device)
    B, D = z_t.shape
    mu = _as_dev_f32(category_vectors, torch, decoder.device)
    if missing_prob > 0:
        if mask is None:
            rng = rng or np.random.default_rng()
            mask = rng.choice(2, B * D, p=[missing_prob, 1. - missing_prob]).reshape(B, D).astype('float32')  # :1475
        fill0 = 'mean'
    else:
        mask = np.ones((B, D), np.float32)   # :1485-1486: no mask and no where(z == 0) fill on this branch
        fill0 = 'none'
    bits = pack_targets(decoder, output_images)
    cat = None if category_list is None else _as_dev_f32(category_list, torch, decoder.device)

    def branch(fill, k):
        r = anytime_eval(decoder, z_t, mask, mu, bits, K=k, seed=seed, fill=fill, return_grid=True, return_loss=True,
                         gamma=0.60)
        loss = r['loss'].mean().item()                                  # reduce_mean(axis=0) :1498
        c = r['counts'].to(torch.float64)
        pr = (c[:, 0] / (c[:, 0] + c[:, 1] + 1e-10)).mean().item()   # :1499-1501
        rc = (c[:, 0] / (c[:, 0] + c[:, 2] + 1e-10)).mean().item()
        acc = None
        if cat is not None:                                             # :1488-1494 / :1511-1518
            _, acc = nearest_prior(r['z_completed'], mu, cat)
        return r['mean_prob'], loss, pr, rc, acc

    pred, loss, pr, rc, acc = branch(fill0, 1)
    if missing_prob == 0.0:
        return pred, loss, pr, rc, acc, 0, 0, 0, 0, 0   # :1502-1503
    pred_c, loss_c, pr_c, rc_c, acc_c = branch('prior_sample', K)
    return pred, loss, pr, rc, acc, pred_c, loss_c, pr_c, rc_c, acc_c


def getEvalImages(encoder: Encoder2D, decoder: Decoder3D, inputs, category_vectors, training: bool = False,
                  missing_prob: float = 0.0, K: int = 1, seed: int = 0, mask=None,
                  rng: np.random.Generator | None = None):
    """``nolboSingleObject_pascal_category_VAE.getEval`` (nolbo.py:855-935) end to end: ``inputs = (input_images,
    output_images, category_list)`` with RGB crops [B,H,W,3] in [0,1]; the encoder (Darknet19 + head2D, one handle from
    ``image_encoder``) produces mean / clipped logvar / z on the GPU (nolbo.py:869-875), then the latent-space
    ``getEval`` above runs unchanged.  Returns the same 10-tuple."""
    if training:
        raise NotImplementedError('inference only')
    input_images, output_images, category_list = inputs
    _, _, z = encoder.encode(input_images, decoder.input_dim, seed=seed ^ 0x656E63)
    return getEval(decoder, (z, output_images, category_list), category_vectors, missing_prob=missing_prob, K=K,
                   seed=seed, mask=mask, rng=rng)


def getEvalVoxels(encoder: Encoder3D, decoder: Decoder3D, inputs, category_vectors, training: bool = False,
                  missing_prob: float = 0.0, K: int = 1, seed: int = 0, mask=None,
                  rng: np.random.Generator | None = None):
    """``nolboSingleObject_modelnet_category_VAE.getEval`` (nolbo.py:1449-1528) end to end from voxel grids:
    ``inputs = (input_images, output_images, category_list)`` with ``input_images`` [B,64,64,64,1]; the voxel encoder
    produces mean / clipped logvar / z on the GPU (nolbo.py:1463-1470), then the latent-space ``getEval`` runs."""
    if training:
        raise NotImplementedError('inference only')
    input_images, output_images, category_list = inputs
    _, _, z = encoder.encode(input_images, decoder.input_dim, seed=seed ^ 0x656E63)
    return getEval(decoder, (z, output_images, category_list), category_vectors, missing_prob=missing_prob, K=K,
                   seed=seed, mask=mask, rng=rng)


# ------------------------------------------------------------------------------------------------ multi-GPU plumbing
def getPredShapes(decoder: Decoder3D, inst_mean, inst_log_var, is_sampling: bool = True, sampling_num: int = 32,
                  seed: int | None = None, obj_offset: int = 0):
    """The 3-D shape block of getPred (src/module/nolbo_test.py:167-182) for ALL selected instances in one call:
    ``is_sampling``: every instance's (mean, logVar) is stacked ``sampling_num`` times, ``sampling`` draws the latents, the
    decoder runs on all of them and the ``sampling_num`` post-sigmoid grids of an instance are averaged
    (``tf.reduce_mean(self._decoder(latents, training=False), axis=0)``); otherwise ``decoder(inst_mean)``.
    The reference loops over the instances (32 decodes per decoder call); here the draws (a3d_sampling), the decodes and
    the K-sample mean (fused into the tail kernel) of all instances are one launch chain.
    Returns float32 [N, 64, 64, 64] (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out); N = 0 gives an empty array."""
    torch = _require_cuda()
    is_np = not isinstance(inst_mean, torch.Tensor)
    dev = decoder.device
    mean = _as_dev_f32(inst_mean, torch, dev).reshape(-1, decoder.input_dim)
    n = mean.shape[0]
    if n == 0:
        out = torch.empty((0, 64, 64, 64), dtype=torch.float32, device=dev)
        return out.cpu().numpy() if is_np else out
    if not is_sampling:
        out = decoder(mean).reshape(n, 64, 64, 64)
        return out.cpu().numpy() if is_np else out
    logvar = _as_dev_f32(inst_log_var, torch, dev).reshape(-1, decoder.input_dim)
    if logvar.shape != mean.shape:
        raise ValueError('inst_mean and inst_log_var must have the same shape')
    K = int(sampling_num)
    zc = sampling(mean.repeat_interleave(K, dim=0), logvar.repeat_interleave(K, dim=0), seed=seed, decoder=decoder,
                  obj_offset=obj_offset * K).reshape(n, K, decoder.input_dim)
    grid = torch.empty((n, 64, 64, 64, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(decoder.device_index):
        _capi.check(decoder._lib.a3d_anytime_eval(decoder._h, zc.data_ptr(), n, K, 0, 0.5, 0, grid.data_ptr(),
                                                  _stream_ptr(torch)), 'a3d_anytime_eval')
    out = grid.reshape(n, 64, 64, 64)
    return out.cpu().numpy() if is_np else out


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous object shard [lo, hi) of rank ``rank``: objects (with all their K samples) never cross GPUs."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def allreduce_counts(counts, group=None):
    """Sum an int64 count tensor over ranks (the only cross-GPU traffic of the path): NCCL over NVLink on GPUs,
    gloo in the CPU tests.  Integer sums => results are bit-identical for any world size."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
