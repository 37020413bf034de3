"""Reader / writer for TensorFlow "tensor bundle" checkpoints (tf.train.Saver V2: `<prefix>.index` +
`<prefix>.data-00000-of-00001`) without TensorFlow.

The reference saves and restores its weights with `tf.train.Saver` (lib/models.py:351, 209-215) and publishes
pretrained models in that format (README.md:104).  cape_b200 names its parameters exactly like the reference's TF
variables (cape_b200/params.py), so a checkpoint maps onto the engine by name:

    from cape_b200 import tf_checkpoint
    values = tf_checkpoint.read_checkpoint("checkpoints/CAPE_nz64/model.ckpt-12345")     # {name: ndarray}
    model.load_weights(values)                       # or CAPE.restore(), which finds TF checkpoints by itself

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, not part of the reference tree; restated
from its published description):
  * `.index` is an SSTable (LevelDB table): data blocks of prefix-compressed (key, value) entries with restart points,
    an index block mapping last-keys to block handles, a 48-byte footer (metaindex handle, index handle, magic
    0xdb4775248b80fb57); every block is followed by a 1-byte compression type (0 none, 1 snappy) and a masked CRC32C.
    Key "" holds a BundleHeaderProto (num_shards, endianness, version), every other key is a tensor name holding a
    BundleEntryProto (dtype, shape, shard_id, offset, size, crc32c).
  * `.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at those offsets.
The writer produces uncompressed single-shard bundles the same way, so weights trained here can be handed back to the
reference (`saver.restore`).  It has been checked against this reader only -- TensorFlow is not installable in the build
container -- and against the format constants above.
"""
import os
import re
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}

# ---------------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked as in LevelDB / TensorFlow
# ---------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = [int(v) for v in t]
    return _CRC_TABLE


def crc32c(data, crc=0):
    t = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------
# varints / minimal protobuf
# ---------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """{field number: [values]}; varint -> int, 64-bit -> bytes(8), length-delimited -> bytes, 32-bit -> bytes(4)."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            v, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(fn, []).append(v)
    return out


def _field(fn, wt, payload):
    return _put_varint((fn << 3) | wt) + payload


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
    p = _parse_proto(buf)
    shape = []
    if 2 in p:
        for dim in _parse_proto(p[2][0]).get(2, []):           # TensorShapeProto.dim
            shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
    return dict(dtype=p.get(1, [0])[0], shape=tuple(shape), shard_id=p.get(3, [0])[0], offset=p.get(4, [0])[0],
                size=p.get(5, [0])[0], crc32c=struct.unpack("<I", p[6][0])[0] if 6 in p else None, sliced=7 in p)


def _make_entry(dtype_id, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(s)) for s in shape))
    return (_field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims) +
            _field(4, 0, _put_varint(offset)) + _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc)))


# ---------------------------------------------------------------------------------------------------
# snappy (raw format) -- index blocks are normally stored uncompressed, but the table format allows it
# ---------------------------------------------------------------------------------------------------
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):                       # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block")
    return bytes(out)


# ---------------------------------------------------------------------------------------------------
# SSTable
# ---------------------------------------------------------------------------------------------------
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    body, ctype, crc = raw[:size], raw[size], struct.unpack("<I", raw[size + 1:size + 5])[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != crc:
        raise ValueError("checksum mismatch in table block at offset %d" % offset)
    if ctype == 1:
        body = _snappy_decompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression type %d" % ctype)
    return body


def _block_entries(block):
    nrestarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_table(path, verify=True):
    with open(path, "rb") as f:
        f.seek(0, 2)
        n = f.tell()
        if n < 48:
            raise ValueError("%s is too short to be a checkpoint index" % path)
        f.seek(n - 48)
        footer = f.read(48)
        if struct.unpack("<Q", footer[40:])[0] != MAGIC:
            raise ValueError("%s is not a TensorFlow checkpoint index (bad magic)" % path)
        _, pos = _get_varint(footer, 0)                  # metaindex handle (unused)
        _, pos = _get_varint(footer, pos)
        ioff, pos = _get_varint(footer, pos)
        isize, pos = _get_varint(footer, pos)
        out = []
        for _, handle in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, p2 = _get_varint(handle, 0)
            bsize, _ = _get_varint(handle, p2)
            out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
        return out


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    out += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def _write_table(path, items, block_size=4096):
    """items: sorted list of (key bytes, value bytes)."""
    with open(path, "wb") as f:
        def emit(block):
            off = f.tell()
            f.write(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
            return _put_varint(off) + _put_varint(len(block))

        index, cur, cur_size = [], [], 0
        for k, v in items:
            cur.append((k, v))
            cur_size += len(k) + len(v)
            if cur_size >= block_size:
                index.append((cur[-1][0], emit(_build_block(cur))))
                cur, cur_size = [], 0
        if cur:
            index.append((cur[-1][0], emit(_build_block(cur))))
        meta = emit(_build_block([]))
        idx = emit(_build_block(index, restart_interval=1))
        footer = meta + idx
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))


# ---------------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------------
def is_checkpoint(prefix):
    return os.path.exists(str(prefix) + ".index")


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by `<directory>/checkpoint`, else the highest-numbered *.index."""
    state = os.path.join(directory, "checkpoint")
    if os.path.exists(state):
        m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), flags=re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(directory, p)
            if is_checkpoint(p):
                return p
            p = os.path.join(directory, os.path.basename(p))        # directories get moved around
            if is_checkpoint(p):
                return p
    best = None
    if os.path.isdir(directory):
        for fn in os.listdir(directory):
            if fn.endswith(".index"):
                m = re.search(r"-(\d+)\.index$", fn)
                key = (int(m.group(1)) if m else -1, fn)
                if best is None or key > best[0]:
                    best = (key, os.path.join(directory, fn[:-6]))
    return best[1] if best else None


def list_variables(prefix, verify=True):
    """[(name, shape, numpy dtype)] like tf.train.list_variables."""
    out = []
    for k, v in _read_table(str(prefix) + ".index", verify):
        if k == b"":
            continue
        e = _parse_entry(v)
        out.append((k.decode(), e["shape"], _DTYPES.get(e["dtype"])))
    return out


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} for every (or the named) numeric tensor of the bundle `<prefix>.index/.data-*`.
    verify: check the table-block checksums and every tensor's CRC32C (pure Python: ~1 s per 10 MB; pass False to skip
    the tensor checksums)."""
    prefix = str(prefix)
    entries, header = {}, None
    for k, v in _read_table(prefix + ".index", True):
        if k == b"":
            header = _parse_proto(v)
        else:
            entries[k.decode()] = _parse_entry(v)
    if header is None:
        raise ValueError("%s.index has no bundle header" % prefix)
    num_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise NotImplementedError("big-endian checkpoint")
    out, files = {}, {}
    try:
        for name, e in entries.items():
            if names is not None and name not in names:
                continue
            if e["sliced"]:
                raise NotImplementedError("partitioned variable %s" % name)
            dt = _DTYPES.get(e["dtype"])
            if dt is None:
                continue                                   # strings / resources: nothing the engine can use
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb")
            f = files[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("truncated data file for %s" % name)
            if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise ValueError("checksum mismatch for tensor %s" % name)
            out[name] = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder("<")).reshape(e["shape"]).astype(dt)
    finally:
        for f in files.values():
            f.close()
    if names is not None:
        missing = sorted(set(names) - set(out))
        if missing:
            raise KeyError("not in checkpoint %s: %s" % (prefix, missing))
    return out


def write_checkpoint(prefix, tensors, update_state_file=True):
    """Write {name: ndarray} as a single-shard, uncompressed V2 bundle (`<prefix>.index`, `<prefix>.data-00000-of-00001`)
    and, like tf.train.Saver, point `<dir>/checkpoint` at it."""
    prefix = str(prefix)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
    # BundleHeaderProto: num_shards = 1, (endianness = LITTLE is the default 0), version { producer: 1 }
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            if a.dtype not in _DTYPE_IDS:
                raise TypeError("unsupported dtype %s for %s" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes(order="C")
            f.write(raw)
            items.append((name.encode(), _make_entry(_DTYPE_IDS[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    _write_table(prefix + ".index", items)
    if update_state_file:
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
            b = os.path.basename(prefix)
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (b, b))
    return prefix
