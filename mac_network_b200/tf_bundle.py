"""TensorFlow-free reader / writer of TensorFlow's checkpoint format ("tensor bundle", what `tf.train.Saver().save/restore`
of the reference reads and writes: `/root/reference/main.py:163-201`), so that real trained MAC weights
(`weights{epoch}.ckpt.index` + `weights{epoch}.ckpt.data-00000-of-00001`) load into `MACParams` and checkpoints written here
restore in the reference.

Format, restated from TensorFlow's public sources (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{table,format,
block}.cc -- the LevelDB table format -- and tensor_bundle.proto); TensorFlow itself is not installable in this image, so the
restatement is checked by round trips and structural tests only (tests/test_tf_bundle.py):

  <prefix>.index   an SSTable: data blocks of prefix-compressed (key, value) entries
                       varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value
                   each block ending with its restart offsets (uint32 LE) and their count, followed by a 5-byte trailer
                   (compression type, masked crc32c of block + type); a meta-index block; an index block mapping a separator
                   key of every data block to its BlockHandle (varint64 offset, varint64 size); a 48-byte footer (the two
                   handles, zero padding, magic 0xdb4775248b80fb57).
                   key ""            -> BundleHeaderProto { num_shards = 1, endianness = LITTLE, version { producer = 1 } }
                   key <tensor name> -> BundleEntryProto  { dtype, shape, shard_id, offset, size, crc32c (masked) }
  <prefix>.data-00000-of-00001   the tensors' raw little-endian bytes, concatenated in key order.
"""
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
       17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _DT.items()}

# ------------------------------------------------------------------ crc32c (Castagnoli), masked as leveldb / TF store it
_POLY = 0x82F63B78
_TABLE = np.zeros(256, dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if (_c & 1) else 0)
    _TABLE[_i] = _c
_TABLE_L = [int(x) for x in _TABLE]


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object (table driven; the tensors themselves go through the C library when it is loaded)."""
    lib = _clib()
    if lib is not None and len(data) >= 4096:
        buf = np.frombuffer(data, dtype=np.uint8)
        return int(lib.mac_host_crc32c(buf.ctypes.data, buf.size, int(crc)))
    c = crc ^ 0xFFFFFFFF
    t = _TABLE_L
    for b in bytes(data):
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _clib():
    try:
        import ctypes
        from . import _lib
        lib = ctypes.CDLL(_lib.LIB_PATH)
        lib.mac_host_crc32c.restype = ctypes.c_uint32
        lib.mac_host_crc32c.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_uint32]
        return lib
    except Exception:
        return None


# ------------------------------------------------------------------ varints / minimal protobuf
def _put_varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not (b & 0x80):
            return val, pos
        shift += 7


def _pb_fields(buf):
    """Yield (field number, wire type, value) of one protobuf message (varint, 64-bit, length-delimited, 32-bit)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, pos = _get_varint(buf, pos)
        elif w == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif w == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif w == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % w)
        yield f, w, v


def _entry_proto(dtype_enum, shape, offset, size, crc_masked):
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    msg = b"\x08" + _put_varint(dtype_enum) + b"\x12" + _put_varint(len(dims)) + dims
    # shard_id (field 3) = 0 is the proto3 default and is omitted, like offset 0
    if offset:
        msg += b"\x20" + _put_varint(offset)
    msg += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc_masked)
    return msg


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for f, w, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:                      # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            raise NotImplementedError("partitioned (sliced) variables are not used by the reference")
    return e


# ------------------------------------------------------------------ SSTable blocks
def _read_block(buf, offset, size, verify):
    body = buf[offset:offset + size]
    ctype = buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if mask_crc(crc32c(bytes(body) + bytes([ctype]))) != stored:
            raise ValueError("index block at %d: checksum mismatch" % offset)
    if ctype != 0:
        raise NotImplementedError("compressed index blocks (type %d); TensorFlow writes checkpoints uncompressed" % ctype)
    nrestart = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * nrestart
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(body, pos)
        non_shared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + bytes(body[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(body[pos:pos + vlen])))
        pos += vlen
    return out


def _build_block(entries, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(body))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


# ------------------------------------------------------------------ public API
def read_tensor_bundle(prefix, verify=True, names=None):
    """{variable name: numpy array} of a TensorFlow V2 checkpoint `<prefix>.index` + `<prefix>.data-00000-of-00001`."""
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    if len(idx) < 48 or struct.unpack_from("<Q", idx, len(idx) - 8)[0] != MAGIC:
        raise ValueError("%s.index is not a TensorFlow checkpoint index (bad magic)" % prefix)
    footer = idx[len(idx) - 48:]
    _, p = _get_varint(footer, 0)                      # metaindex handle (offset, size): unused
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries = []
    for _, handle in _read_block(idx, ioff, isize, verify):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        entries += _read_block(idx, boff, bsize, verify)
    header = dict((f, v) for f, _, v in _pb_fields(entries[0][1])) if entries and entries[0][0] == b"" else {}
    if header.get(1, 1) != 1:
        raise NotImplementedError("checkpoints sharded over %d data files" % header.get(1))
    if header.get(2, 0) != 0:
        raise NotImplementedError("big-endian checkpoints")
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    out = {}
    for key, val in entries:
        if key == b"":
            continue
        name = key.decode()
        if names is not None and name not in names:
            continue
        e = _parse_entry(val)
        if e["dtype"] not in _DT:
            raise NotImplementedError("%s: tensor dtype enum %d" % (name, e["dtype"]))
        raw = np.asarray(data[e["offset"]:e["offset"] + e["size"]])
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw.tobytes())) != e["crc32c"]:
            raise ValueError("%s: tensor checksum mismatch" % name)
        out[name] = raw.view(_DT[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_tensor_bundle(prefix, tensors):
    """Write {name: array} as a TensorFlow V2 checkpoint (one shard, little endian, uncompressed index).  Returns the names."""
    names = sorted(tensors, key=lambda s: s.encode())
    entries = [(b"", b"\x08\x01" + b"\x1a\x02\x08\x01")]       # num_shards = 1, (endianness LITTLE omitted), version.producer = 1
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])
            if not a.flags.c_contiguous:               # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
                a = np.ascontiguousarray(a)
            if a.dtype not in _DT_OF:
                raise NotImplementedError("%s: dtype %s" % (n, a.dtype))
            raw = a.tobytes()
            f.write(raw)
            entries.append((n.encode(), _entry_proto(_DT_OF[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    out = bytearray()
    data_block = _build_block(entries)
    data_handle = _put_varint(len(out)) + _put_varint(len(data_block))
    out += _with_trailer(data_block)
    meta_block = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta_block))
    out += _with_trailer(meta_block)
    index_block = _build_block([(entries[-1][0] + b"\x00", data_handle)], restart_interval=1)   # separator >= the last key
    index_handle = _put_varint(len(out)) + _put_varint(len(index_block))
    out += _with_trailer(index_block)
    footer = meta_handle + index_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    return names
