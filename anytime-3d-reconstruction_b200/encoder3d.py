"""Host mirror of the reference's voxel encoder interface, SURVEY.md section 8 row f2:

* ``encoder3D(structure)``  <- src/net_core/autoencoder3D.py:72-102 (same ``structure`` dict: name, input_shape,
  filter_num_list, filter_size_list, strides_list, final_pool, activation, final_activation) -> callable model
* ``model(voxels, training=False)`` <- src/module/nolbo.py:1463; ``set_weights / get_weights / load_weights /
  save_weights`` in Keras variable order (per conv3DEnc: kernel [kd,kh,kw,Cin,Cout], gamma, beta, moving_mean,
  moving_variance; last: the bare Conv3D kernel)
* ``model.encode(voxels, z_dim)`` -> (mean, clipped logvar, z): the latent split + sampling of nolbo.py:1464-1470.

All arithmetic runs in liba3d (csrc/conv3d_tc.cu: tcgen05 implicit GEMMs fed by element-strided 5-D TMA boxes); there
is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi


def _torch():
    import torch
    return torch


class Encoder3D:
    """Callable stand-in for the ``tf.keras.Model`` returned by the reference's ``encoder3D(structure)``."""

    def __init__(self, structure: dict, max_batch: int = 64, operand_dtype: str = 'fp16', device: int | None = None):
        for key in ('name', 'input_shape', 'filter_num_list', 'filter_size_list', 'strides_list', 'final_pool',
                    'activation', 'final_activation'):
            if key not in structure:
                raise KeyError(key)   # the reference indexes the dict directly
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('a3d needs a CUDA device (sm_100a); there is no CPU fallback')
        self.structure = dict(structure)
        self.name = structure['name']
        shp = list(structure['input_shape'])
        if len(shp) != 4 or shp[0] is None or not (shp[0] == shp[1] == shp[2]) or shp[3] != 1:
            raise ValueError(f'input_shape must be a fixed cubic one-channel grid [G,G,G,1], got {shp}')
        self.filters = [int(f) for f in structure['filter_num_list']]
        ks = [int(k) for k in structure['filter_size_list']]
        st = [int(s) for s in structure['strides_list']]
        if not (len(self.filters) == len(ks) == len(st)) or len(ks) > _capi.A3D_MAX_LAYERS:
            raise ValueError('filter_num_list, filter_size_list and strides_list must have the same length (<= 8)')
        if structure['activation'] not in _capi.ACT or structure['final_activation'] not in _capi.FINAL \
                or structure['final_pool'] not in _capi.POOL:
            raise ValueError('unsupported activation / final_activation / final_pool')
        self.grid = int(shp[0])
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)
        self.max_batch = int(max_batch)
        d = _capi.Enc3dDesc()
        d.abi_version, d.in_grid, d.num_layers = _capi.A3D_ABI_VERSION, self.grid, len(ks)
        for i in range(len(ks)):
            d.filters[i], d.ksizes[i], d.strides[i] = self.filters[i], ks[i], st[i]
        d.final_pool = _capi.POOL[structure['final_pool']]
        d.activation = _capi.ACT[structure['activation']]
        d.final_activation = _capi.FINAL[structure['final_activation']]
        d.device, d.max_batch, d.operand_dtype = self.device_index, self.max_batch, _capi.DTYPE[operand_dtype]
        self._pooled = d.final_pool != 0
        self._lib = _capi.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc3d_create(C.byref(d), C.byref(h)), 'a3d_enc3d_create')
        self._h = h
        g = self.grid
        self._grids = []
        for s in st:
            g //= s
            self._grids.append(g)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.a3d_enc3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def output_shape(self):
        g, c = self._grids[-1], self.filters[-1]
        return (None, c) if self._pooled else (None, g, g, g, c)

    # ---- weights
    def weight_shapes(self) -> list[tuple[int, ...]]:
        shapes, c = [], 1
        n = len(self.filters)
        for i, f in enumerate(self.filters):
            shapes.append((4, 4, 4, c, f))
            if i < n - 1:
                shapes += [(f,)] * 4
            c = f
        return shapes

    def _layer_var_names(self) -> list[list[str]]:
        names = []
        n = len(self.filters)
        for i in range(n):
            names.append(['kernel'])
            if i < n - 1:
                names.append(['gamma', 'beta', 'moving_mean', 'moving_variance'])
        return names

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
            _capi.check(self._lib.a3d_enc3d_set_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes),
                        'a3d_enc3d_set_weight')

    def get_weights(self) -> list[np.ndarray]:
        out = []
        for i, shp in enumerate(self.weight_shapes()):
            a = np.empty(shp, np.float32)
            _capi.check(self._lib.a3d_enc3d_get_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes),
                        'a3d_enc3d_get_weight')
            out.append(a)
        return out

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
        from . import tf_checkpoint
        if tf_checkpoint.is_checkpoint(path):
            self.set_weights(tf_checkpoint.load_keras_weights(path))
            return
        p = path if os.path.exists(path) else path + '.npz'
        with np.load(p) as f:
            self.set_weights([f[f'arr_{i}'] for i in range(len(f.files))])

    # ---- forward
    def __call__(self, voxels, training: bool = False):
        """encoder(voxels, training=False): [N,G,G,G,1] float32 -> [N, filters[-1]] float32.  numpy in -> numpy out."""
        if training:
            raise NotImplementedError('a3d implements the inference path only (training=False)')
        torch = _torch()
        is_np = not isinstance(voxels, torch.Tensor)
        x = torch.from_numpy(np.ascontiguousarray(voxels, dtype=np.float32)) if is_np else voxels
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        g = self.grid
        if x.numel() % (g * g * g) != 0 or (x.dim() >= 4 and tuple(x.shape[1:4]) != (g, g, g)):
            raise ValueError(f'expected voxels [N,{g},{g},{g},1], got {tuple(x.shape)}')
        n = x.numel() // (g * g * g)
        go, c = self._grids[-1], self.filters[-1]
        out = torch.empty((n, c) if self._pooled else (n, go, go, go, c), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc3d_forward(self._h, x.data_ptr(), n, out.data_ptr(),
                                                    int(torch.cuda.current_stream().cuda_stream)), 'a3d_enc3d_forward')
        return out.cpu().numpy() if is_np else out

    predict = __call__

    def split_sample(self, enc_out, D: int, seed: int | None = None, obj_offset: int = 0, clip: float = 10.0):
        """nolbo.py:1464-1470: mean / clip(logvar, +-10) / sampling; CUDA tensors (mean, logvar, z)."""
        torch = _torch()
        e = enc_out if isinstance(enc_out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(enc_out, np.float32))
        e = e.to(device=self.device, dtype=torch.float32).contiguous()
        n, stride = e.shape
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')
        mean = torch.empty((n, D), dtype=torch.float32, device=self.device)
        logvar, z = torch.empty_like(mean), torch.empty_like(mean)
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc3d_split_sample(
                self._h, e.data_ptr(), n, D, stride, float(clip), 1, seed, obj_offset, mean.data_ptr(),
                logvar.data_ptr(), z.data_ptr(), int(torch.cuda.current_stream().cuda_stream)), 'a3d_enc3d_split_sample')
        return mean, logvar, z

    def encode(self, voxels, z_dim: int, seed: int | None = None, obj_offset: int = 0):
        torch = _torch()
        x = voxels if isinstance(voxels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(voxels, np.float32))
        return self.split_sample(self(x), z_dim, seed=seed, obj_offset=obj_offset)

    # ---- diagnostics
    def debug_layer(self, layer: int, n: int) -> np.ndarray:
        g, c = self._grids[layer], self.filters[layer]
        a = np.empty((n, g, g, g, c), np.float32)
        _capi.check(self._lib.a3d_enc3d_debug_read_layer(self._h, layer, n, a.ctypes.data_as(C.c_void_p), a.nbytes),
                    'a3d_enc3d_debug_read_layer')
        return a

    @property
    def launch_count(self) -> int:
        return int(self._lib.a3d_enc3d_launch_count(self._h))

    def workspace_bytes(self) -> int:
        return int(self._lib.a3d_enc3d_workspace_bytes(self._h))


def encoder3D(structure: dict, **kw) -> Encoder3D:
    """Same call as the reference's ``src.net_core.autoencoder3D.encoder3D(structure)`` (autoencoder3D.py:72).
    Keyword extras (not in the reference): max_batch, operand_dtype ('fp16' | 'bf16'), device."""
    return Encoder3D(structure, **kw)
