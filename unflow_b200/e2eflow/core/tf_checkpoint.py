"""TensorFlow checkpoint ("tensor bundle", the V2 Saver format) import / export.

The reference saves and restores its networks with ``tf.train.Saver`` (src/e2eflow/core/train.py:
23-65 ``restore_networks``, :258-259 ``saver.save(sess, save_path, global_step=i)``) and locates
them through the ``checkpoint`` state file (``tf.train.get_checkpoint_state``, src/e2eflow/util.py:
75-85, train.py:124).  The released UnFlow models (README.md:116-128) are such files:

    <dir>/checkpoint                          text proto: model_checkpoint_path / all_model_checkpoint_paths
    <dir>/model.ckpt-<iter>.index             table (sorted string -> string) of BundleEntryProto
    <dir>/model.ckpt-<iter>.data-00000-of-00001   raw little-endian tensor bytes

This module reads and writes that format without TensorFlow so checkpoints move between the
reference and this implementation in both directions.

Format, restated from the TensorFlow sources (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc},
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/io/{format,block,table_builder}.cc
-- the LevelDB table format); TensorFlow itself is not in /root/reference and not installed here,
so the restatement is anchored on the format's published constants (table magic
0xdb4775248b80fb57, CRC-32C check value 0xe3069283 for "123456789", the LevelDB CRC mask delta
0xa282ead8) and on a write -> read round trip; NO file written by TensorFlow was available to test
against:

  * ``.index`` = data blocks, an (empty) metaindex block, an index block, and a 48-byte footer:
    two block handles (varint64 offset, varint64 size) padded to 40 bytes + the 8-byte magic.
  * every block is followed by a 5-byte trailer: compression type (0 none, 1 snappy) and the masked
    CRC-32C of contents+type.  Block contents = prefix-compressed entries
    ``varint32 shared | varint32 non_shared | varint32 value_len | key delta | value`` followed by
    the uint32 restart offsets and their count.
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness (0 little), 3: version {1: producer}};
    every other key is a variable name -> BundleEntryProto {1: dtype, 2: shape {2: dim {1: size}},
    3: shard_id, 4: offset, 5: size, 6: fixed32 masked crc32c of the bytes, 7: slices}.

Variable names are the TF scopes the network code creates (``flownet_c/conv4/weights`` ...);
conv kernels are stored HWIO and transposed-conv kernels ``[kh, kw, out, in]`` -- exactly what
``FlowNetVariables.to_tf_dict`` / ``load_tf_dict`` produce and consume.  Adam's slots are the
variables ``<name>/Adam`` (m) and ``<name>/Adam_1`` (v).
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8
BLOCK_SIZE = 262144          # table::Options::block_size default
RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'),
           5: np.dtype('<i2'), 6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'),
           17: np.dtype('<u2'), 19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ---- checksums -------------------------------------------------------------------------------
def crc32c(data, crc=0):
    """CRC-32C through the native library (host code in csrc/checksum.cu)."""
    from ... import _native
    if isinstance(data, np.ndarray):
        if not data.flags.c_contiguous:
            data = np.ascontiguousarray(data)
        return int(_native.lib().unflow_crc32c(data.ctypes.data, data.nbytes, crc))
    data = bytes(data)
    return int(_native.lib().unflow_crc32c(data, len(data), crc))


def mask_crc(crc):
    """LevelDB's masking: rotate right by 15 and add a constant (crcs of crcs stay well behaved)."""
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ---- varints / protobuf wire format ------------------------------------------------------------
def _put_varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while n >= 0x80:
        out.append((n & 0x7f) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        byte = buf[pos]
        pos += 1
        result |= (byte & 0x7f) << shift
        if byte < 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _fields(buf):
    """Yield (field number, wire type, value) of one protobuf message; nested messages as bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        number, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _get_varint(buf, pos)
        elif wire == 1:
            value = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wire == 2:
            size, pos = _get_varint(buf, pos)
            value = bytes(buf[pos:pos + size])
            if len(value) != size:
                raise CheckpointError("truncated length-delimited field")
            pos += size
        elif wire == 5:
            value = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wire)
        yield number, wire, value


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _tag(number, wire):
    return _put_varint((number << 3) | wire)


def _bytes_field(number, payload):
    return _tag(number, 2) + _put_varint(len(payload)) + payload


def _encode_entry(dtype_id, shape, shard_id, offset, size, masked_crc):
    dims = b''.join(_bytes_field(2, _tag(1, 0) + _put_varint(int(d))) for d in shape)
    msg = _tag(1, 0) + _put_varint(dtype_id) + _bytes_field(2, dims)
    if shard_id:
        msg += _tag(3, 0) + _put_varint(shard_id)
    if offset:
        msg += _tag(4, 0) + _put_varint(offset)
    if size:
        msg += _tag(5, 0) + _put_varint(size)
    return msg + _tag(6, 5) + struct.pack('<I', masked_crc)


def _decode_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': 0, 'slices': 0}
    for number, wire, value in _fields(buf):
        if number == 1:
            e['dtype'] = value
        elif number == 2:
            for n2, _, dim in _fields(value):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _fields(dim):
                        if n3 == 1:
                            size = _signed64(v3)
                    e['shape'].append(size)
                elif n2 == 3 and dim:
                    raise CheckpointError("tensor of unknown rank in checkpoint")
        elif number == 3:
            e['shard_id'] = value
        elif number == 4:
            e['offset'] = value
        elif number == 5:
            e['size'] = value
        elif number == 6:
            e['crc32c'] = value
        elif number == 7:
            e['slices'] += 1
    return e


# ---- snappy (index blocks may be compressed by other writers) --------------------------------
def _snappy_uncompress(buf):
    total, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            length = tag >> 2
            if length >= 60:
                extra = length - 59
                length = int.from_bytes(buf[pos:pos + extra], 'little')
                pos += extra
            length += 1
            out += buf[pos:pos + length]
            pos += length
            continue
        if kind == 1:
            length = ((tag >> 2) & 7) + 4
            offset = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            length = (tag >> 2) + 1
            offset = int.from_bytes(buf[pos:pos + 2], 'little')
            pos += 2
        else:
            length = (tag >> 2) + 1
            offset = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if offset == 0 or offset > len(out):
            raise CheckpointError("corrupt snappy block")
        for _ in range(length):               # copies may overlap their own output
            out.append(out[-offset])
    if len(out) != total:
        raise CheckpointError("corrupt snappy block (length)")
    return bytes(out)


# ---- table blocks -----------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    end = offset + size
    if end + 5 > len(buf):
        raise CheckpointError("block handle points outside the index file")
    contents, kind = buf[offset:end], buf[end]
    if verify:
        want = unmask_crc(struct.unpack_from('<I', buf, end + 1)[0])
        if crc32c(buf[offset:end + 1]) != want:
            raise CheckpointError("index block checksum mismatch")
    if kind == 1:
        contents = _snappy_uncompress(contents)
    elif kind != 0:
        raise CheckpointError("unknown block compression type %d" % kind)
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("table block too small")
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 * (n_restarts + 1)
    if limit < 0:
        raise CheckpointError("corrupt table block")
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        value_len, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + value_len > limit:
            raise CheckpointError("corrupt table entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + value_len])
        pos += value_len


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b''

    def add(self, key, value):
        shared = 0
        if self.count % RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            limit = min(len(key), len(self.last))
            while shared < limit and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * (len(self.restarts) + 1)

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + \
            struct.pack('<I', len(self.restarts))


def _write_table(path, items):
    """items: (key bytes, value bytes) sorted by key."""
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block)
        out.extend(b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return handle

    index, data = _BlockBuilder(), _BlockBuilder()
    for key, value in items:
        if data.count and key <= data.last:
            raise CheckpointError("table keys must be strictly increasing")
        data.add(key, value)
        if data.size() >= BLOCK_SIZE:
            index.add(data.last, emit(data.finish()))
            data = _BlockBuilder()
    if data.count:
        index.add(data.last, emit(data.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(out)


def _read_table(path, verify=True):
    with open(path, 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise CheckpointError("%s is not a TensorFlow checkpoint index (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused by the bundle)
    _, pos = _get_varint(footer, pos)
    index_off, pos = _get_varint(footer, pos)
    index_size, pos = _get_varint(footer, pos)
    items = []
    for _, handle in _block_entries(_read_block(buf, index_off, index_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        items.extend(_block_entries(_read_block(buf, off, size, verify)))
    return items


# ---- the bundle -------------------------------------------------------------------------------
def _data_path(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


class BundleReader:
    """Read-only view of one checkpoint: ``BundleReader(prefix).tensor(name)``."""

    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        items = _read_table(prefix + '.index', verify)
        if not items or items[0][0] != b'':
            raise CheckpointError("checkpoint index has no header entry")
        self.num_shards, endianness = 1, 0
        for number, _, value in _fields(items[0][1]):
            if number == 1:
                self.num_shards = value
            elif number == 2:
                endianness = value
        if endianness != 0:
            raise CheckpointError("big-endian checkpoints are not supported")
        self.entries = {key.decode('utf-8'): _decode_entry(value) for key, value in items[1:]}

    def variables(self):
        """name -> (numpy dtype, shape)"""
        return {k: (_DTYPES.get(e['dtype']), tuple(e['shape'])) for k, e in self.entries.items()}

    def __contains__(self, name):
        return name in self.entries

    def tensor(self, name):
        if name not in self.entries:
            raise KeyError("%s not found in checkpoint %s" % (name, self.prefix))
        e = self.entries[name]
        if e['slices']:
            raise CheckpointError("%s is a partitioned variable (not produced by the reference)" % name)
        if e['dtype'] not in _DTYPES:
            raise CheckpointError("%s: unsupported dtype enum %d" % (name, e['dtype']))
        dtype = _DTYPES[e['dtype']]
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if count * dtype.itemsize != e['size']:
            raise CheckpointError("%s: %d bytes stored for shape %s" % (name, e['size'], e['shape']))
        arr = np.empty(e['shape'], dtype=dtype)
        with open(_data_path(self.prefix, e['shard_id'], self.num_shards), 'rb') as f:
            f.seek(e['offset'])
            if f.readinto(memoryview(arr).cast('B')) != e['size']:
                raise CheckpointError("%s: data file truncated" % name)
        if self.verify and crc32c(arr) != unmask_crc(e['crc32c']):
            raise CheckpointError("%s: tensor checksum mismatch" % name)
        return arr


def write_bundle(prefix, tensors):
    """Write ``{name: array}`` as ``prefix.index`` + ``prefix.data-00000-of-00001``."""
    header = _tag(1, 0) + _put_varint(1) + _bytes_field(3, _tag(1, 0) + _put_varint(1))
    items, offset = [(b'', header)], 0
    tmp = _data_path(prefix, 0, 1) + '.tmp'
    with open(tmp, 'wb') as f:
        for name in sorted(tensors, key=lambda s: s.encode('utf-8')):
            if not name:
                raise CheckpointError("empty variable name")
            arr = np.asarray(tensors[name])
            if not arr.flags.c_contiguous:      # (ascontiguousarray would turn scalars into [1])
                arr = np.ascontiguousarray(arr)
            if arr.dtype.byteorder == '>':
                arr = arr.astype(arr.dtype.newbyteorder('<'))
            if arr.dtype not in _DTYPE_IDS:
                raise CheckpointError("%s: dtype %s cannot be stored" % (name, arr.dtype))
            dtype_id = _DTYPE_IDS[arr.dtype]
            f.write(arr.tobytes())
            items.append((name.encode('utf-8'),
                          _encode_entry(dtype_id, arr.shape, 0, offset, arr.nbytes, mask_crc(crc32c(arr)))))
            offset += arr.nbytes
    os.replace(tmp, _data_path(prefix, 0, 1))
    _write_table(prefix + '.index', items)


# ---- the ``checkpoint`` state file ---------------------------------------------------------------
def get_checkpoint_state(ckpt_dir):
    """``tf.train.get_checkpoint_state``: (newest checkpoint prefix, [all prefixes]) or None."""
    path = os.path.join(ckpt_dir, 'checkpoint')
    if not os.path.isfile(path):
        return None
    latest, every = None, []
    with open(path) as f:
        for line in f:
            m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths)\s*:\s*"(.*)"\s*$', line)
            if not m:
                continue
            p = m.group(2)
            if not os.path.isabs(p):
                p = os.path.join(ckpt_dir, p)
            if m.group(1) == 'model_checkpoint_path':
                latest = p
            else:
                every.append(p)
    if latest is None:
        return None
    return latest, every


def update_checkpoint_state(ckpt_dir, prefix):
    state = get_checkpoint_state(ckpt_dir)
    every = [p for p in (state[1] if state else []) if p != prefix] + [prefix]
    rel = lambda p: os.path.relpath(p, ckpt_dir) if os.path.dirname(os.path.abspath(p)) == \
        os.path.abspath(ckpt_dir) else p
    with open(os.path.join(ckpt_dir, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % rel(prefix))
        for p in every:
            f.write('all_model_checkpoint_paths: "%s"\n' % rel(p))


def checkpoint_iteration(prefix):
    """train.py:130-131: the iteration is parsed from the file name ``model.ckpt-<iter>``."""
    return int(os.path.basename(prefix).split('-')[-1])


# ---- networks <-> checkpoints -------------------------------------------------------------------
def net_names(flownet_spec):
    """train.py:29: the scope prefixes the Savers filter on."""
    return ['flownet_c'] + ['stack_%d_flownet' % (i + 1) for i in range(len(flownet_spec) - 1)]


def restore_variables(variables, prefix, nets=None, allow_partial=True, verify=True):
    """Load the variables of the networks ``nets`` (indices into the stack; None = all) from a TF
    checkpoint into a ``FlowNetVariables``.  Like the reference's fallback (train.py:52-61), a
    checkpoint trained without ``full_res`` may miss the two full-resolution up-convolutions; with
    ``allow_partial`` those stay at their initial values.  Returns the restored names."""
    reader = BundleReader(prefix, verify)
    scopes = list(variables.kinds) if nets is None else \
        [s for i in nets for s in variables.scopes_of_net(i)]
    found, missing = {}, []
    for scope in scopes:
        for suffix in ('/weights', '/biases'):
            name = scope + suffix
            if name in reader:
                found[name] = reader.tensor(name)
            else:
                missing.append(name)
    # the extra up-convolutions live in the variable scope 'full_res' (flownet.py:133-153); the
    # reference's fallback drops every variable with 'full_res' in its name (train.py:57-59)
    if missing and not (allow_partial and all('full_res' in m for m in missing)):
        raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[0]))
    variables.load_tf_dict(found, strict=False)
    return sorted(found)


def save_variables(variables, prefix, nets=None, adam_slots=None):
    """Write the networks ``nets`` of a ``FlowNetVariables`` (and optionally Adam's slots,
    ``{name: (m, v)}`` in TF layout) as a TF checkpoint and update the ``checkpoint`` state file."""
    tensors = variables.to_tf_dict()
    if nets is not None:
        keep = {s for i in nets for s in variables.scopes_of_net(i)}
        tensors = {k: v for k, v in tensors.items() if k.rsplit('/', 1)[0] in keep}
    out = {k: np.asarray(v, dtype=np.float32) for k, v in tensors.items()}
    for name, (m, v) in (adam_slots or {}).items():
        if name in out:
            out[name + '/Adam'] = np.asarray(m, dtype=np.float32)
            out[name + '/Adam_1'] = np.asarray(v, dtype=np.float32)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    write_bundle(prefix, out)
    update_checkpoint_state(os.path.dirname(os.path.abspath(prefix)), prefix)
    return prefix
