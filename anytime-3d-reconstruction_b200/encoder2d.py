"""Host mirror of the reference's image encoder interface (src/net_core/darknet.py), SURVEY.md section 8 row f1.

* ``Darknet19(name=None, activation='elu')``                       <- darknet.py:96-133 (returns a callable model)
* ``head2D(name, input_shape, output_dim, filter_num_list, filter_size_list, last_pooling=None, activation='elu')``
                                                                   <- darknet.py:149-168
* ``model(x, training=False)``, ``model.output_shape``, ``set_weights / get_weights / load_weights / save_weights``
  in Keras variable order (per conv: kernel [kh,kw,Cin,Cout], then gamma, beta, moving_mean, moving_variance)
* ``image_encoder(head_structure, ...)``: backbone + head in ONE handle (what ``head(backbone(images))`` computes,
  src/module/nolbo.py:869), with ``encode(images) -> (mean, logvar, z)`` doing the latent split, the +-10 clip and
  ``sampling`` of nolbo.py:869-875 on the GPU.

All arithmetic runs in liba3d (tcgen05 implicit-GEMM convolutions, see csrc/conv2d_tc.cu); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi

# (filters, kernel size) of the 18 Darknet19 convolutions, darknet.py:99-131; pools follow convs 0, 1, 4, 7, 12
_DARKNET19 = [(32, 3), (64, 3), (128, 3), (64, 1), (128, 3), (256, 3), (128, 1), (256, 3),
              (512, 3), (256, 1), (512, 3), (256, 1), (512, 3),
              (1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3)]
_DARKNET19_POOL_AFTER = (0, 1, 4, 7, 12)


def darknet19_layers(activation: str = 'elu') -> list[dict]:
    L = []
    for i, (f, k) in enumerate(_DARKNET19):
        L.append({'kind': 'conv', 'filters': f, 'ksize': k, 'bn': True, 'act': activation})
        if i in _DARKNET19_POOL_AFTER:
            L.append({'kind': 'maxpool'})
    return L


def head2d_layers(output_dim, filter_num_list, filter_size_list, last_pooling=None, activation='elu') -> list[dict]:
    L = []
    for f, k in zip(filter_num_list, filter_size_list):
        L.append({'kind': 'conv', 'filters': int(f), 'ksize': int(k), 'bn': True, 'act': activation})
    L.append({'kind': 'conv', 'filters': int(output_dim), 'ksize': 1, 'bn': False, 'act': None})
    if last_pooling == 'max':
        L.append({'kind': 'global_max'})
    elif last_pooling == 'average':
        L.append({'kind': 'global_avg'})
    return L


def _torch():
    import torch
    return torch


class Encoder2D:
    """Callable stand-in for the ``tf.keras.Model`` objects the reference's Darknet19 / head2D return."""

    def __init__(self, layers: list[dict], input_shape, name: str | None = None, max_batch: int = 128,
                 operand_dtype: str = 'fp16', device: int | None = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('a3d needs a CUDA device (sm_100a); there is no CPU fallback')
        self.name = name
        self.layers = [dict(l) for l in layers]
        self.input_shape = tuple(int(v) for v in input_shape)          # (H, W, C)
        if len(self.layers) > _capi.A3D_ENC_MAX_LAYERS:
            raise ValueError('too many layers')
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)
        self.operand_dtype = operand_dtype
        self._tdtype = torch.float16 if _capi.DTYPE[operand_dtype] == 0 else torch.bfloat16
        self.max_batch = int(max_batch)
        self.uint8_scale = 1.0 / 255.0     # uint8 images are scaled like the loader does (pascal3D.py:242)
        d = _capi.Enc2dDesc()
        d.abi_version = _capi.A3D_ABI_VERSION
        d.in_h, d.in_w, d.in_ch = self.input_shape
        d.num_layers = len(self.layers)
        for i, l in enumerate(self.layers):
            d.layers[i].kind = _capi.L2D[l['kind']]
            if l['kind'] == 'conv':
                if l['act'] not in _capi.ACT2D:
                    raise ValueError(f"unsupported activation {l['act']!r}")
                d.layers[i].filters, d.layers[i].ksize = int(l['filters']), int(l['ksize'])
                d.layers[i].batch_norm, d.layers[i].activation = int(bool(l['bn'])), _capi.ACT2D[l['act']]
        d.device, d.max_batch, d.operand_dtype = self.device_index, self.max_batch, _capi.DTYPE[operand_dtype]
        self._lib = _capi.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc2d_create(C.byref(d), C.byref(h)), 'a3d_enc2d_create')
        self._h = h
        dims = (C.c_int32 * 3)()
        _capi.check(self._lib.a3d_enc2d_output_shape(self._h, dims), 'a3d_enc2d_output_shape')
        self._out_hwc = tuple(int(v) for v in dims)
        self._pooled = self.layers[-1]['kind'] in ('global_max', 'global_avg')

    # ---- lifetime
    def close(self):
        if getattr(self, '_h', None):
            self._lib.a3d_enc2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def output_shape(self):
        """Keras-style: (None, h, w, c), or (None, c) after a global pool (darknet.py:158-163)."""
        return (None, self._out_hwc[2]) if self._pooled else (None,) + self._out_hwc

    # ---- weights
    def weight_shapes(self) -> list[tuple[int, ...]]:
        shapes, c = [], self.input_shape[2]
        for l in self.layers:
            if l['kind'] != 'conv':
                continue
            shapes.append((l['ksize'], l['ksize'], c, l['filters']))
            if l['bn']:
                shapes += [(l['filters'],)] * 4
            c = l['filters']
        return shapes

    @property
    def num_weights(self) -> int:
        return int(self._lib.a3d_enc2d_num_weights(self._h))

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
            _capi.check(self._lib.a3d_enc2d_set_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes),
                        'a3d_enc2d_set_weight')

    def get_weights(self) -> list[np.ndarray]:
        out = []
        for i, shp in enumerate(self.weight_shapes()):
            a = np.empty(shp, np.float32)
            _capi.check(self._lib.a3d_enc2d_get_weight(self._h, i, a.ctypes.data_as(C.c_void_p), a.nbytes),
                        'a3d_enc2d_get_weight')
            out.append(a)
        return out

    def _layer_var_names(self) -> list[list[str]]:
        names = []
        for l in self.layers:
            if l['kind'] == 'conv':
                names.append(['kernel'])
                if l['bn']:
                    names.append(['gamma', 'beta', 'moving_mean', 'moving_variance'])
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
        from . import tf_checkpoint
        if tf_checkpoint.is_checkpoint(path):
            self.set_weights(tf_checkpoint.load_keras_weights(path))
            return
        p = path if os.path.exists(path) else path + '.npz'
        with np.load(p) as f:
            self.set_weights([f[f'arr_{i}'] for i in range(len(f.files))])

    # ---- forward
    def __call__(self, x, training: bool = False, out_dtype: str = 'fp32'):
        """model(x, training=False).  x: [N,H,W,C] NHWC, numpy / torch; fp32 (or the operand dtype for feature maps).
        Returns fp32 NHWC ([N,C] after a global pool); ``out_dtype=operand dtype`` keeps 16-bit features for chaining
        (head(backbone(x, out_dtype='fp16')) skips the fp32 round trip).  numpy in -> numpy out."""
        if training:
            raise NotImplementedError('a3d implements the inference path only (training=False)')
        torch = _torch()
        is_np = not isinstance(x, torch.Tensor)
        if is_np:
            x = np.asarray(x)
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint8 if x.dtype == np.uint8 else np.float32))
        if x.dtype == torch.uint8:
            # raw image bytes: the loader's `image / 255.` (pascal3D.py:242) runs on the device
            if x.device.type == 'cpu' and out_dtype in ('fp32', 'f32', 'float32'):
                return self._call_host(x.contiguous(), is_np)
            x = x.to(self.device).to(torch.float32) * self.uint8_scale
        if x.device.type == 'cpu' and out_dtype in ('fp32', 'f32', 'float32'):
            return self._call_host(x.to(torch.float32).contiguous(), is_np)
        if x.dtype not in (torch.float32, self._tdtype) or self.input_shape[2] == 3:
            x = x.to(torch.float32)
        x = x.to(self.device).contiguous()
        if tuple(x.shape[1:]) != self.input_shape:
            raise ValueError(f'expected input [N,{self.input_shape}], got {tuple(x.shape)}')
        n = x.shape[0]
        in_io = 2 if x.dtype == torch.float32 else _capi.DTYPE[self.operand_dtype]
        out_io = _capi.IO[out_dtype]
        odt = torch.float32 if out_io == 2 else self._tdtype
        shape = (n, self._out_hwc[2]) if self._pooled else (n,) + self._out_hwc
        out = torch.empty(shape, dtype=odt, device=self.device)
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc2d_forward(self._h, x.data_ptr(), in_io, n, out.data_ptr(), out_io,
                                                    int(torch.cuda.current_stream().cuda_stream)), 'a3d_enc2d_forward')
        return out.cpu().numpy() if is_np else out

    predict = __call__

    def _call_host(self, x, is_np: bool):
        """Host images (numpy / CPU tensor, pinned or pageable; fp32, or uint8 bytes scaled by ``uint8_scale`` on the
        device) through a3d_enc2d_forward_host[_u8]: chunks of max_batch with
        the H2D copy of chunk i+1 overlapping the forward of chunk i.  numpy in -> numpy out, CPU tensor in -> CUDA out."""
        torch = _torch()
        if tuple(x.shape[1:]) != self.input_shape:
            raise ValueError(f'expected input [N,{self.input_shape}], got {tuple(x.shape)}')
        n = x.shape[0]
        shape = (n, self._out_hwc[2]) if self._pooled else (n,) + self._out_hwc
        if is_np:
            out = np.empty(shape, np.float32)
            out_dev, out_host = None, out.ctypes.data_as(C.c_void_p)
        else:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
            out_dev, out_host = out.data_ptr(), None
        with torch.cuda.device(self.device_index):
            torch.cuda.current_stream().synchronize()      # the call runs on the handle's own streams
            if x.dtype == torch.uint8:
                _capi.check(self._lib.a3d_enc2d_forward_host_u8(self._h, x.data_ptr(), self.uint8_scale, n, out_dev,
                                                                out_host), 'a3d_enc2d_forward_host_u8')
            else:
                _capi.check(self._lib.a3d_enc2d_forward_host(self._h, x.data_ptr(), n, out_dev, out_host),
                            'a3d_enc2d_forward_host')
        return out

    def split_sample(self, enc_out, D: int, seed: int | None = None, obj_offset: int = 0, clip: float = 10.0):
        """nolbo.py:869-875: mean = out[:, :D]; logvar = clip(out[:, D:2D], -10, 10); z = sampling(mean, logvar).
        ``seed=None`` draws a fresh seed (the reference is unseeded); returns CUDA tensors (mean, logvar, z)."""
        torch = _torch()
        e = enc_out if isinstance(enc_out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(enc_out, np.float32))
        e = e.to(device=self.device, dtype=torch.float32).contiguous()
        n, stride = e.shape
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')
        mean = torch.empty((n, D), dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mean)
        z = torch.empty_like(mean)
        with torch.cuda.device(self.device_index):
            _capi.check(self._lib.a3d_enc2d_split_sample(
                self._h, e.data_ptr(), n, D, stride, float(clip), 1, seed, obj_offset, mean.data_ptr(),
                logvar.data_ptr(), z.data_ptr(), int(torch.cuda.current_stream().cuda_stream)), 'a3d_enc2d_split_sample')
        return mean, logvar, z

    def encode(self, images, z_dim: int, seed: int | None = None, obj_offset: int = 0):
        """images -> (mean, logvar, z) CUDA tensors: forward + split_sample (the encoder half of getEval)."""
        torch = _torch()
        if not isinstance(images, torch.Tensor):
            images = np.asarray(images)
            images = torch.from_numpy(np.ascontiguousarray(images, np.uint8 if images.dtype == np.uint8 else np.float32))
        return self.split_sample(self(images), z_dim, seed=seed, obj_offset=obj_offset)

    # ---- diagnostics
    def debug_layer(self, layer: int, n: int) -> np.ndarray:
        """fp32 NHWC output of layer ``layer`` for the first n images of the last chunk, real channels only.
        A conv directly followed by a max-pool is stored pooled (the pool is fused): read the pool's index."""
        dims = (C.c_int32 * 4)()
        _capi.check(self._lib.a3d_enc2d_layer_shape(self._h, layer, dims), 'a3d_enc2d_layer_shape')
        hh, ww, c, cp = (int(v) for v in dims)
        a = np.empty((n, hh, ww, cp), np.float32)
        _capi.check(self._lib.a3d_enc2d_debug_read_layer(self._h, layer, n, a.ctypes.data_as(C.c_void_p), a.nbytes),
                    'a3d_enc2d_debug_read_layer')
        return a[..., :c]

    @property
    def launch_count(self) -> int:
        return int(self._lib.a3d_enc2d_launch_count(self._h))

    def workspace_bytes(self) -> int:
        return int(self._lib.a3d_enc2d_workspace_bytes(self._h))


def Darknet19(name=None, activation='elu', input_size=(256, 256), **kw) -> Encoder2D:
    """Same call as ``src.net_core.darknet.Darknet19(name, activation)`` (darknet.py:96).  The Keras model accepts
    any image size; this build fixes it at construction (``input_size``, multiples of 32; the reference evaluates
    256 x 256 crops, test_pascal_VAE_dr.py:52).  Extras: max_batch, operand_dtype, device."""
    return Encoder2D(darknet19_layers(activation), (int(input_size[0]), int(input_size[1]), 3), name=name, **kw)


def head2D(name, input_shape, output_dim, filter_num_list, filter_size_list, last_pooling=None, activation='elu',
           **kw) -> Encoder2D:
    """Same call as ``src.net_core.darknet.head2D`` (darknet.py:149); ``input_shape`` = backbone.output_shape[1:]."""
    return Encoder2D(head2d_layers(output_dim, filter_num_list, filter_size_list, last_pooling, activation),
                     tuple(input_shape), name=name, **kw)


def image_encoder(head_structure: dict, activation: str = 'elu', input_size=(256, 256), last_pooling: str = 'max',
                  **kw) -> Encoder2D:
    """Darknet19 + head2D fused into one handle: ``head(backbone(images))`` of nolbo.py:869 with the head built as at
    nolbo.py:778-783 (``last_pooling='max'``).  ``head_structure`` is the reference's ``config['encoder_head']`` dict
    (test_pascal_VAE_dr.py:193-201).  Weights: backbone.get_weights() + head.get_weights()."""
    layers = darknet19_layers(activation) + head2d_layers(
        head_structure['output_dim'], head_structure['filter_num_list'], head_structure['filter_size_list'],
        last_pooling, head_structure.get('activation', 'elu'))
    return Encoder2D(layers, (int(input_size[0]), int(input_size[1]), 3), name=head_structure.get('name'), **kw)
