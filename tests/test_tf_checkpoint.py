"""TensorFlow V2 checkpoint (tensor bundle) reader / writer without TensorFlow (cape_b200/tf_checkpoint.py)."""
import os
import struct

import numpy as np
import pytest

from cape_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                     # the CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43               # 32 bytes of ones
    assert T.crc32c(bytes(range(32))) == 0x46DD794E                 # ascending bytes
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283      # incremental
    # LevelDB's mask is a rotation plus a constant, invertible
    m = T.mask_crc(0xE3069283)
    rot = (m - 0xa282ead8) & 0xFFFFFFFF
    assert ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF == 0xE3069283


def _tensors(rng):
    from cape_b200.params import NZ64_AFFINE, init_params, param_specs
    specs = param_specs(NZ64_AFFINE, [6890, 6890, 3445, 3445, 1723, 1723, 862, 862, 862], [6890, 3445, 1723, 862, 431])
    small = {k: v for k, v in specs.items() if int(np.prod(v)) < 200000}       # keep the CPU test quick
    vals = init_params(small, 1)
    vals["global_step"] = np.asarray(12345, np.int64)
    vals["generator/encoder/encoder_conv1/weights/Momentum"] = rng.normal(size=(6, 64)).astype(np.float32)
    vals["some/double"] = rng.normal(size=(3, 1, 2))
    vals["some/empty"] = np.zeros((0, 4), np.float32)
    vals["some/int32"] = np.arange(7, dtype=np.int32)
    return vals


def test_roundtrip_and_listing(tmp_path):
    rng = np.random.RandomState(0)
    vals = _tensors(rng)
    prefix = str(tmp_path / "model.ckpt-12345")
    T.write_checkpoint(prefix, vals)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    assert T.is_checkpoint(prefix) and T.latest_checkpoint(str(tmp_path)) == prefix
    listed = T.list_variables(prefix)
    assert [n for n, _, _ in listed] == sorted(vals, key=lambda s: s.encode())   # table keys are sorted
    got = T.read_checkpoint(prefix)
    assert set(got) == set(vals)
    for k, v in vals.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape, k
        assert np.array_equal(got[k], v), k
    one = T.read_checkpoint(prefix, names=["global_step"])
    assert int(one["global_step"]) == 12345
    with pytest.raises(KeyError):
        T.read_checkpoint(prefix, names=["nope"])


def test_file_layout_matches_the_published_format(tmp_path):
    """Footer magic, block trailers and the header entry, byte by byte."""
    prefix = str(tmp_path / "m")
    T.write_checkpoint(prefix, {"a": np.float32([1, 2, 3]), "b/c": np.float32([[4.0]])})
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57
    entries = T._read_table(prefix + ".index")
    assert entries[0][0] == b"" and [k for k, _ in entries[1:]] == [b"a", b"b/c"]
    hdr = T._parse_proto(entries[0][1])
    assert hdr[1] == [1]                                              # num_shards
    e = T._parse_entry(entries[2][1])
    assert e["dtype"] == 1 and e["shape"] == (1, 1) and e["offset"] == 12 and e["size"] == 4
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert np.array_equal(np.frombuffer(data, "<f4"), [1, 2, 3, 4])
    assert e["crc32c"] == T.mask_crc(T.crc32c(data[12:16]))


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m")
    T.write_checkpoint(prefix, {"w": np.arange(100, dtype=np.float32)})
    fn = prefix + ".data-00000-of-00001"
    b = bytearray(open(fn, "rb").read())
    b[17] ^= 0x40
    open(fn, "wb").write(bytes(b))
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)
    assert T.read_checkpoint(prefix, verify=False)["w"].shape == (100,)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)


def test_snappy_blocks_and_prefix_compression(tmp_path):
    """Blocks written by other producers may be snappy-compressed and always use key prefix sharing."""
    assert T._snappy_decompress(bytes([11, 0x28]) + b"hello world") == b"hello world"
    # literal "ab" + copy(offset 2, length 6) -> "abababab"
    assert T._snappy_decompress(bytes([8, 0x04]) + b"ab" + bytes([((6 - 4) << 2) | 1, 2])) == b"abababab"
    items = [(("layer%03d/weights" % i).encode(), bytes([i])) for i in range(40)]
    blk = T._build_block(items, restart_interval=16)
    assert list(T._block_entries(blk)) == items
    assert len(blk) < sum(len(k) + 4 for k, _ in items)              # shared prefixes were elided
