"""ctypes binding of the liba3d C ABI (include/a3d.h).  No compute happens in Python; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# A3D_LIB selects another build of the same library in the package directory (liba3d_checked.so: the bounds-checked
# build of tests/test_gpu_checked_build.py); there is still no non-CUDA implementation behind this binding.
LIB_PATH = os.path.join(_HERE, os.path.basename(os.environ.get('A3D_LIB', 'liba3d.so')))
A3D_ABI_VERSION = 1
A3D_MAX_LAYERS = 8
VOXELS = 262144

ACT = {'None': 0, None: 0, 'linear': 0, 'elu': 1, 'relu': 2, 'lrelu': 3}
# darknet.py:87-92 / :140-145: 'lrelu' there is LeakyReLU(alpha=0.1)
ACT2D = {'None': 0, None: 0, 'linear': 0, 'elu': 1, 'relu': 2, 'lrelu': 4}
A3D_ENC_MAX_LAYERS = 40
L2D = {'conv': 0, 'maxpool': 1, 'global_max': 2, 'global_avg': 3}
IO = {'fp16': 0, 'f16': 0, 'float16': 0, 'bf16': 1, 'bfloat16': 1, 'fp32': 2, 'f32': 2, 'float32': 2}
FINAL = {'None': 0, None: 0, 'linear': 0, 'sigmoid': 1}
DTYPE = {'fp16': 0, 'f16': 0, 'float16': 0, 'bf16': 1, 'bfloat16': 1}
IMPL = {'tcgen05': 0, 'simt': 1}
FILL = {'prior_sample': 0, 'mean': 1, 'normal': 2, 'none': 3}
OUT = {'f32': 0, 'fp32': 0, 'float32': 0, 'f16': 1, 'fp16': 1, 'float16': 1, 'bits': 2}


class Desc(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32), ('latent_dim', C.c_int32), ('num_layers', C.c_int32),
        ('filters', C.c_int32 * A3D_MAX_LAYERS), ('ksizes', C.c_int32 * A3D_MAX_LAYERS),
        ('strides', C.c_int32 * A3D_MAX_LAYERS), ('out_grid', C.c_int32), ('activation', C.c_int32),
        ('final_activation', C.c_int32), ('device', C.c_int32), ('max_chunk', C.c_int32),
        ('operand_dtype', C.c_int32), ('impl', C.c_int32),
    ]


class Layer2d(C.Structure):
    _fields_ = [('kind', C.c_int32), ('filters', C.c_int32), ('ksize', C.c_int32), ('batch_norm', C.c_int32),
                ('activation', C.c_int32)]


class Enc2dDesc(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('in_h', C.c_int32), ('in_w', C.c_int32), ('in_ch', C.c_int32),
                ('num_layers', C.c_int32), ('layers', Layer2d * A3D_ENC_MAX_LAYERS), ('device', C.c_int32),
                ('max_batch', C.c_int32), ('operand_dtype', C.c_int32)]


class Enc3dDesc(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('in_grid', C.c_int32), ('num_layers', C.c_int32),
                ('filters', C.c_int32 * A3D_MAX_LAYERS), ('ksizes', C.c_int32 * A3D_MAX_LAYERS),
                ('strides', C.c_int32 * A3D_MAX_LAYERS), ('final_pool', C.c_int32), ('activation', C.c_int32),
                ('final_activation', C.c_int32), ('device', C.c_int32), ('max_batch', C.c_int32),
                ('operand_dtype', C.c_int32)]


POOL = {'None': 0, None: 0, 'average': 1, 'max': 2}

# name -> (restype, argtypes); must list every symbol include/a3d.h declares (tests check this against the header)
SIGNATURES = {
    'a3d_create': (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_void_p)]),
    'a3d_destroy': (None, [C.c_void_p]),
    'a3d_num_weights': (C.c_int, [C.c_void_p]),
    'a3d_weight_numel': (C.c_int64, [C.c_void_p, C.c_int]),
    'a3d_set_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_get_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_decode': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'a3d_decode_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float]),
    'a3d_sampling': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    'a3d_nearest_prior': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    'a3d_impute': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                             C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'a3d_anytime_eval': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    'a3d_anytime_eval_loss': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'a3d_binary_loss': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_void_p]),
    'a3d_counts_sweep': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p]),
    'a3d_counts': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p,
                             C.c_void_p]),
    'a3d_pack_targets': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    'a3d_anytime_eval_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                        C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                        C.c_void_p]),
    'a3d_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'a3d_debug_read_layer': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_size_t]),
    'a3d_debug_time_tail': (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    'a3d_launch_count': (C.c_int64, [C.c_void_p]),
    'a3d_set_profiling': (C.c_int, [C.c_void_p, C.c_int]),
    'a3d_stage_times_ms': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'a3d_enc2d_create': (C.c_int, [C.POINTER(Enc2dDesc), C.POINTER(C.c_void_p)]),
    'a3d_enc2d_destroy': (None, [C.c_void_p]),
    'a3d_enc2d_num_weights': (C.c_int, [C.c_void_p]),
    'a3d_enc2d_weight_numel': (C.c_int64, [C.c_void_p, C.c_int]),
    'a3d_enc2d_set_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_enc2d_get_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_enc2d_output_shape': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    'a3d_enc2d_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    'a3d_enc2d_forward_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'a3d_enc2d_forward_host_u8': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]),
    'a3d_enc2d_split_sample': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int,
                                         C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'a3d_enc2d_layer_shape': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    'a3d_enc2d_debug_read_layer': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_size_t]),
    'a3d_enc2d_launch_count': (C.c_int64, [C.c_void_p]),
    'a3d_enc2d_workspace_bytes': (C.c_size_t, [C.c_void_p]),
    'a3d_enc3d_create': (C.c_int, [C.POINTER(Enc3dDesc), C.POINTER(C.c_void_p)]),
    'a3d_enc3d_destroy': (None, [C.c_void_p]),
    'a3d_enc3d_num_weights': (C.c_int, [C.c_void_p]),
    'a3d_enc3d_weight_numel': (C.c_int64, [C.c_void_p, C.c_int]),
    'a3d_enc3d_set_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_enc3d_get_weight': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    'a3d_enc3d_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'a3d_enc3d_split_sample': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int,
                                         C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'a3d_enc3d_debug_read_layer': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_size_t]),
    'a3d_enc3d_launch_count': (C.c_int64, [C.c_void_p]),
    'a3d_enc3d_workspace_bytes': (C.c_size_t, [C.c_void_p]),
    'a3d_crc32c': (C.c_uint32, [C.c_void_p, C.c_size_t, C.c_uint32]),
    'a3d_last_error': (C.c_char_p, []),
    'a3d_abi_version': (C.c_int, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load liba3d.so; fail loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: the CUDA extension is not built (run `python -c "import __graft_entry__ as g; '
            f'g.build()"` in the repo root).  This package has no CPU or PyTorch fallback.')
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if l.a3d_abi_version() != A3D_ABI_VERSION:
        raise RuntimeError('liba3d ABI version mismatch; rebuild the extension')
    _lib = l
    return l


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().a3d_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (status {rc}): {msg}')
