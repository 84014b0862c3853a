"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol of include/a3d.h, and the host
mirror of the reference interface behaves like the reference's (names, arguments, errors).  No compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'a3d.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(a3d_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import a3d
    from a3d import _capi
    lib = _capi.lib()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/a3d.h but not exported'
        assert s in _capi.SIGNATURES, f'{s} has no ctypes signature'
    assert set(_capi.SIGNATURES) == set(syms)
    assert lib.a3d_abi_version() == 1


def test_desc_struct_matches_header_layout():
    from a3d import _capi
    assert C.sizeof(_capi.Desc) == 4 * (3 + 3 * 8 + 7)


def test_no_cpu_fallback():
    import a3d
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU'):
        a3d.decoder3D(a3d.presets.MODELNET_DECODER)
    from a3d import _capi
    lib = _capi.lib()
    d = _capi.Desc()
    d.abi_version, d.latent_dim, d.num_layers, d.out_grid, d.max_chunk = 1, 64, 5, 64, 32
    for i, (f, s) in enumerate(zip([512, 256, 128, 64, 1], [1, 2, 2, 2, 2])):
        d.filters[i], d.ksizes[i], d.strides[i] = f, 4, s
    d.activation, d.final_activation = 1, 1
    h = C.c_void_p()
    rc = lib.a3d_create(C.byref(d), C.byref(h))
    assert rc == -3 and b'no CPU path' in lib.a3d_last_error()       # A3D_ERR_NO_DEVICE
    d.filters[1] = 300
    assert lib.a3d_create(C.byref(d), C.byref(h)) == -1               # unsupported structure -> A3D_ERR_INVALID
    assert b'unsupported decoder structure' in lib.a3d_last_error()


def test_structure_dict_schema_and_errors():
    import a3d
    s = a3d._parse_structure(a3d.presets.MODELNET_DECODER)
    assert s['grid0'] == [4, 4, 4] and s['ch0'] == 8 and s['dense_units'] == 512 and s['input_dim'] == 64
    bad = dict(a3d.presets.MODELNET_DECODER)
    del bad['strides_list']
    with pytest.raises(KeyError):
        a3d._parse_structure(bad)      # the reference indexes structure['strides_list'] directly


def test_shard_range_partitions_objects():
    import a3d
    for n, w in [(10, 3), (1024, 8), (5, 8), (0, 2)]:
        parts = [a3d.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(lo <= hi for lo, hi in parts)


def test_iou_from_counts():
    import a3d
    m, g = a3d.iou_from_counts(np.array([[5, 10, 5], [0, 0, 0]]))
    assert g == pytest.approx(0.25) and m == pytest.approx((0.25 + 1.0) / 2)


def test_encoder_handles_validate_structure_before_touching_a_device():
    """a3d_enc2d_create / a3d_enc3d_create reject unsupported structures with A3D_ERR_INVALID (-1) and a message, and
    report A3D_ERR_NO_DEVICE (-3) for a valid structure when no GPU is present (no CPU fallback)."""
    from a3d import _capi
    from a3d.encoder2d import darknet19_layers, head2d_layers
    lib = _capi.lib()

    def enc2d_desc(layers, h, w, c):
        d = _capi.Enc2dDesc()
        d.abi_version, d.in_h, d.in_w, d.in_ch, d.num_layers = 1, h, w, c, len(layers)
        for i, l in enumerate(layers):
            d.layers[i].kind = _capi.L2D[l['kind']]
            if l['kind'] == 'conv':
                d.layers[i].filters, d.layers[i].ksize = l['filters'], l['ksize']
                d.layers[i].batch_norm, d.layers[i].activation = int(l['bn']), _capi.ACT2D[l['act']]
        d.max_batch, d.operand_dtype = 4, 0
        return d

    h = C.c_void_p()
    full = darknet19_layers() + head2d_layers(32, [], [], 'max')
    assert C.sizeof(_capi.Enc2dDesc) == 4 * (5 + 5 * 40 + 3)
    rc = lib.a3d_enc2d_create(C.byref(enc2d_desc(full, 256, 256, 3)), C.byref(h))
    assert rc == (0 if torch.cuda.is_available() else -3)
    if rc == 0:
        lib.a3d_enc2d_destroy(h)
    assert lib.a3d_enc2d_create(C.byref(enc2d_desc(full, 200, 200, 3)), C.byref(h)) == -1        # 25 x 25 at the 4th pool
    assert b'odd size' in lib.a3d_last_error()
    assert lib.a3d_enc2d_create(C.byref(enc2d_desc(full, 256, 256, 5)), C.byref(h)) == -1
    assert b'in_ch must be 3' in lib.a3d_last_error()
    bad = [dict(full[0], ksize=5)] + full[1:]
    assert lib.a3d_enc2d_create(C.byref(enc2d_desc(bad, 256, 256, 3)), C.byref(h)) == -1
    assert lib.a3d_enc2d_create(C.byref(enc2d_desc([{'kind': 'global_max'}], 8, 8, 64)), C.byref(h)) == -1
    assert b'global pool' in lib.a3d_last_error()

    d3 = _capi.Enc3dDesc()
    d3.abi_version, d3.in_grid, d3.num_layers, d3.final_pool, d3.activation, d3.max_batch = 1, 64, 5, 1, 1, 4
    for i, (f, s) in enumerate(zip([64, 128, 256, 512, 128], [2, 2, 2, 2, 1])):
        d3.filters[i], d3.ksizes[i], d3.strides[i] = f, 4, s
    assert C.sizeof(_capi.Enc3dDesc) == 4 * (3 + 3 * 8 + 6)
    rc = lib.a3d_enc3d_create(C.byref(d3), C.byref(h))
    assert rc == (0 if torch.cuda.is_available() else -3)
    if rc == 0:
        lib.a3d_enc3d_destroy(h)
    d3.strides[3] = 1
    assert lib.a3d_enc3d_create(C.byref(d3), C.byref(h)) == -1
    assert b'unsupported encoder3D structure' in lib.a3d_last_error()
    d3.strides[3] = 2
    d3.filters[2] = 200
    assert lib.a3d_enc3d_create(C.byref(d3), C.byref(h)) == -1


def test_python_mirrors_of_the_encoders_refuse_to_run_without_a_gpu():
    import a3d
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU'):
        a3d.Darknet19()
    with pytest.raises(RuntimeError, match='no CPU'):
        a3d.encoder3D(a3d.presets.MODELNET_ENCODER)
    with pytest.raises(KeyError):
        a3d.encoder3D({'name': 'x'})


def test_header_is_plain_c(tmp_path):
    """include/a3d.h is the C-ABI contract: it must compile as C99 with no C++ or CUDA types in the signatures."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if not gcc:
        pytest.skip('gcc not available')
    src = tmp_path / 'use_a3d.c'
    src.write_text('#include "a3d.h"\n'
                   'int main(void) { a3d_desc d; a3d_enc2d_desc e; a3d_enc3d_desc v;\n'
                   '  int (*f0)(void) = a3d_abi_version;\n'
                   '  int (*f1)(a3d_enc2d*, const void*, int, int64_t, void*, int, void*) = a3d_enc2d_forward;\n'
                   '  int (*f2)(a3d_enc3d*, const float*, int64_t, float*, void*) = a3d_enc3d_forward;\n'
                   '  (void)d; (void)e; (void)v; (void)f0; (void)f1; (void)f2; return 0; }\n')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
