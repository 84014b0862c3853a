"""CPU tests of the TensorFlow tensor-bundle checkpoint reader / writer (a3d.tf_checkpoint).
PARITY UNPINNED against TensorFlow itself (not installable here): the reader is checked against the writer, against
hand-built table / snappy / protobuf bytes and against the published CRC32C and varint known answers."""
import os
import struct

import numpy as np
import pytest

import a3d  # noqa: F401
from a3d import tf_checkpoint as tc
from oracle import decoder_ref as dr


def test_crc32c_and_varint_known_answers():
    assert tc.crc32c(b'123456789') == 0xE3069283                      # CRC-32C (Castagnoli) check value
    assert tc.crc32c(b'\x00' * 32) == 0x8A9136AA                      # RFC 3720 B.4
    assert tc._mask_crc(0) == 0xa282ead8
    assert tc._put_varint(300) == b'\xac\x02' and tc._varint(b'\xac\x02', 0) == (300, 2)


def test_crc32c_numpy_fallback_matches_the_byte_loop_and_the_native_routine():
    """Without liba3d (a host that only converts checkpoints) large tensors are checksummed by a chunk-parallel numpy
    evaluation; it must agree with the per-byte table loop, with the native slice-by-8 routine and with chained start
    values, for lengths around the chunking boundaries."""
    rng = np.random.default_rng(5)
    assert tc._crc32c_numpy(b'123456789', 0) == 0xE3069283          # short input: falls through to the byte loop
    for n in ((1 << 20), (1 << 20) + 1, (1 << 20) + 16383, 3_000_017):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        want = tc._crc_bytes(d, 0xFFFFFFFF) ^ 0xFFFFFFFF if n < 1_200_000 else None
        got = tc._crc32c_numpy(d, 0)
        if want is not None:
            assert got == want
        assert got == tc.crc32c(d)                                  # native routine when the library is built
        mid = n // 3
        assert tc._crc32c_numpy(d[mid:], tc._crc32c_numpy(d[:mid], 0)) == got or mid < (1 << 20)
        assert tc._crc32c_numpy(d[mid:], tc.crc32c(d[:mid])) == got


def test_snappy_decoder_on_hand_built_stream():
    # literal "abcd", copy(offset 4, len 8) -> "abcdabcdabcd", literal "Z"
    stream = tc._put_varint(13) + bytes([(4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([0]) + b'Z'
    assert tc._snappy_decompress(stream) == b'abcdabcdabcdZ'
    with pytest.raises(ValueError):
        tc._snappy_decompress(tc._put_varint(5) + bytes([(4 - 1) << 2]) + b'abcd')


def test_round_trip_many_tensors_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f'layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE': rng.standard_normal((3, i + 1)).astype(np.float32)
               for i in range(150)}                                    # > 1 data block, shared key prefixes
    tensors['step'] = np.array(7, np.int64)
    tensors['half'] = rng.standard_normal(5).astype(np.float16)
    tensors['empty'] = np.zeros((0, 4), np.float32)
    prefix = str(tmp_path / 'ckpt' / 'decoder')
    tc.save_checkpoint(prefix, tensors, block_entries=16)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    got = tc.load_checkpoint(prefix)
    assert set(got) == set(tensors)
    for k in tensors:
        assert got[k].dtype == tensors[k].dtype and got[k].shape == tensors[k].shape
        assert np.array_equal(got[k], tensors[k])
    index = tc.read_index(prefix + '.index')
    assert b'' in index and struct.unpack('<Q', open(prefix + '.index', 'rb').read()[-8:])[0] == 0xdb4775248b80fb57


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'c')
    tc.save_checkpoint(prefix, {'a': np.arange(8, dtype=np.float32)})
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[5] ^= 0xFF
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        tc.load_checkpoint(prefix)
    assert tc.load_checkpoint(prefix, verify=False)['a'].shape == (8,)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        tc.load_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(b'not a table')
    with pytest.raises(ValueError, match='bad magic'):
        tc.load_checkpoint(prefix)


def test_bfloat16_and_string_tensors(tmp_path):
    """A bf16 entry decodes to fp32; a DT_STRING entry (the object graph Keras stores) is skipped."""
    prefix = str(tmp_path / 'b')
    tc.save_checkpoint(prefix, {'x': np.zeros(2, np.float32)})
    index = tc.read_index(prefix + '.index')
    vals = np.array([1.0, -2.5, 3.140625], np.float32)
    raw = (vals.view(np.uint32) >> 16).astype('<u2').tobytes()
    entry = (tc._proto_varint(1, 14) + tc._proto_bytes(2, tc._proto_bytes(2, tc._proto_varint(1, 3))) +
             tc._proto_varint(4, 8) + tc._proto_varint(5, len(raw)))
    sentry = tc._proto_varint(1, 7) + tc._proto_bytes(2, b'') + tc._proto_varint(4, 8 + len(raw)) + tc._proto_varint(5, 4)
    with open(prefix + '.data-00000-of-00001', 'ab') as f:
        f.write(raw + b'\x03abc')
    items = sorted(list(index.items()) + [(b'bf', entry), (b'_CHECKPOINTABLE_OBJECT_GRAPH', sentry)])
    table = bytearray()
    blk = tc._build_block(items)
    table += blk + b'\x00' + struct.pack('<I', tc._mask_crc(tc.crc32c(blk + b'\x00')))
    h = tc._put_varint(0) + tc._put_varint(len(blk))
    off = len(table)
    iblk = tc._build_block([(items[-1][0], h)], 1)
    table += iblk + b'\x00' + struct.pack('<I', tc._mask_crc(tc.crc32c(iblk + b'\x00')))
    footer = tc._put_varint(off) + tc._put_varint(len(iblk)) + tc._put_varint(off) + tc._put_varint(len(iblk))
    table += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    open(prefix + '.index', 'wb').write(bytes(table))
    got = tc.load_checkpoint(prefix)
    assert set(got) == {'x', 'bf'} and np.array_equal(got['bf'], vals)


def test_keras_order_is_numeric_by_layer_then_variable(tmp_path):
    """27 decoder variables written under Keras object-graph keys come back in get_weights() order
    (layer_with_weights-10 sorts after -9, gamma/beta/moving_* in Keras order)."""
    ws = dr.keras_default_weights(dict(dr.MODELNET_DECODER, input_dim=4), 3)
    rng = np.random.default_rng(1)
    ws = [w if w.ndim > 1 else rng.standard_normal(w.shape).astype(np.float32) for w in ws]
    bn = ['gamma', 'beta', 'moving_mean', 'moving_variance']
    names = [['kernel', 'bias'], bn] + sum([[['kernel'], bn] for _ in range(4)], []) + [['kernel']]
    assert sum(len(n) for n in names) == 27
    prefix = str(tmp_path / 'decoder')
    tc.save_keras_weights(prefix, ws, names)
    assert tc.is_checkpoint(prefix)
    back = tc.load_keras_weights(prefix)
    assert len(back) == 27 and all(np.array_equal(a, b) for a, b in zip(back, ws))
    with pytest.raises(ValueError, match='not a Keras'):
        tc.keras_weight_list({'foo': np.zeros(1)})


def _literal_snappy(raw: bytes) -> bytes:
    """A valid snappy stream made of literals (<= 60 bytes each) plus one back-reference copy when the data allows it."""
    out = bytearray(tc._put_varint(len(raw)))
    pos = 0
    while pos < len(raw):
        if pos >= 8 and raw[pos:pos + 4] == raw[pos - 4:pos] and pos + 4 <= len(raw):      # copy(offset 4, len 4)
            out += bytes([((4 - 4) << 2) | 1, 4])
            pos += 4
            continue
        n = min(60, len(raw) - pos)
        out += bytes([(n - 1) << 2]) + raw[pos:pos + n]
        pos += n
    return bytes(out)


def _table(blocks, compress):
    """LevelDB table with one data block per entry list; `compress` -> snappy (type 1) data blocks."""
    table = bytearray()
    handles = []
    for entries in blocks:
        blk = tc._build_block(entries)
        body, ctype = (_literal_snappy(blk), b'\x01') if compress else (blk, b'\x00')
        handles.append((entries[-1][0], tc._put_varint(len(table)) + tc._put_varint(len(body))))
        table += body + ctype + struct.pack('<I', tc._mask_crc(tc.crc32c(body + ctype)))
    off = len(table)
    iblk = tc._build_block(handles, 1)
    table += iblk + b'\x00' + struct.pack('<I', tc._mask_crc(tc.crc32c(iblk + b'\x00')))
    footer = tc._put_varint(off) + tc._put_varint(len(iblk)) + tc._put_varint(off) + tc._put_varint(len(iblk))
    table += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    return bytes(table)


def test_hand_assembled_bundle_snappy_multi_shard_and_object_graph(tmp_path):
    """What a TensorFlow-written Keras checkpoint can contain beyond this package's own writer: snappy-compressed index
    blocks, tensors spread over two data shards, and the _CHECKPOINTABLE_OBJECT_GRAPH string tensor."""
    rng = np.random.default_rng(4)
    names = [['kernel', 'bias'], ['gamma', 'beta', 'moving_mean', 'moving_variance'], ['kernel']]
    ws = [rng.standard_normal(s).astype(np.float32) for s in [(3, 4), (4,), (4,), (4,), (4,), (4,), (2, 2, 2, 5, 4)]]
    graph = tc.object_graph_proto(names)
    shards = [bytearray(), bytearray()]
    entries = []
    it = iter(ws)
    for i, vs in enumerate(names):
        for v in vs:
            a = next(it)
            raw = a.tobytes()
            sid = (i + len(v)) & 1
            shape = b''.join(tc._proto_bytes(2, tc._proto_varint(1, int(d))) for d in a.shape)
            e = (tc._proto_varint(1, 1) + tc._proto_bytes(2, shape) + (tc._proto_varint(3, sid) if sid else b'') +
                 (tc._proto_varint(4, len(shards[sid])) if len(shards[sid]) else b'') + tc._proto_varint(5, len(raw)) +
                 tc._put_varint((6 << 3) | 5) + struct.pack('<I', tc._mask_crc(tc.crc32c(raw))))
            entries.append((f'layer_with_weights-{i}/{v}/.ATTRIBUTES/VARIABLE_VALUE'.encode(), e))
            shards[sid] += raw
    sraw, c = tc._string_tensor_bytes(graph)
    entries.append((b'_CHECKPOINTABLE_OBJECT_GRAPH', tc._proto_varint(1, 7) + tc._proto_bytes(2, b'') + tc._proto_varint(3, 1) +
                    tc._proto_varint(4, len(shards[1])) + tc._proto_varint(5, len(sraw)) + tc._put_varint((6 << 3) | 5) +
                    struct.pack('<I', tc._mask_crc(c))))
    shards[1] += sraw
    header = (b'', tc._proto_varint(1, 2) + tc._proto_bytes(3, tc._proto_varint(1, 1)))      # num_shards = 2
    entries = sorted([header] + entries)
    prefix = str(tmp_path / 'decoder')
    open(prefix + '.index', 'wb').write(_table([entries[:4], entries[4:]], compress=True))
    for i in range(2):
        open(f'{prefix}.data-{i:05d}-of-00002', 'wb').write(bytes(shards[i]))
    back = tc.load_keras_weights(prefix)
    assert len(back) == 7 and all(np.array_equal(a, b) for a, b in zip(back, ws))
    got = tc.load_checkpoint(prefix, strings=True)
    nodes = tc.parse_object_graph(got['_CHECKPOINTABLE_OBJECT_GRAPH'])
    assert list(nodes[0]['children']) == ['layer_with_weights-0', 'layer_with_weights-1', 'layer_with_weights-2']
    bn = nodes[nodes[0]['children']['layer_with_weights-1']]
    assert list(bn['children']) == names[1]
    var = nodes[bn['children']['moving_variance']]
    assert var['attributes'] == [('VARIABLE_VALUE', 'layer_1/moving_variance:0',
                                  'layer_with_weights-1/moving_variance/.ATTRIBUTES/VARIABLE_VALUE')]
    # a corrupted string tensor is detected
    bad = bytearray(shards[1]); bad[-3] ^= 0x55
    open(f'{prefix}.data-00001-of-00002', 'wb').write(bytes(bad))
    with pytest.raises(ValueError, match='checksum'):
        tc.load_checkpoint(prefix, strings=True)


def test_writer_emits_the_object_graph_and_keras_format_rule(tmp_path):
    names = [['kernel', 'bias'], ['kernel']]
    ws = [np.ones((2, 3), np.float32), np.zeros(3, np.float32), np.full((3, 1), 2.0, np.float32)]
    prefix = str(tmp_path / 'm')
    tc.save_keras_weights(prefix, ws, names)
    got = tc.load_checkpoint(prefix, strings=True)
    assert tc.parse_object_graph(got['_CHECKPOINTABLE_OBJECT_GRAPH'])[0]['children'] == {'layer_with_weights-0': 1,
                                                                                          'layer_with_weights-1': 4}
    assert all(np.array_equal(a, b) for a, b in zip(tc.load_keras_weights(prefix), ws))
    # Keras: no recognised suffix -> TensorFlow checkpoint format (nolbo.py:1572-1574 passes a bare prefix)
    assert tc.wants_tf_format('weights/decoder', None) and tc.wants_tf_format('x.ckpt', None)
    assert not tc.wants_tf_format('weights/decoder.npz', None) and not tc.wants_tf_format('d', 'npz')
    assert tc.wants_tf_format('d.npz', 'tf')
    with pytest.raises(NotImplementedError):
        tc.wants_tf_format('decoder.h5', None)
