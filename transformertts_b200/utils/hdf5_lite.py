"""Minimal pure-python HDF5 reader / writer for Keras ``model_weights.hdf5`` files (reference: model/models.py:619,637 --
``save_weights`` / ``load_weights`` through h5py, which is not installable here).

Implements the subset of the HDF5 file format (HDF Group "HDF5 File Format Specification", version 0 superblock family)
that h5py's default ``libver='earliest'`` produces for such files:

  * superblock version 0 / 1, 8-byte offsets and lengths;
  * "old style" groups: version-1 object headers with a Symbol Table message, version-1 group B-trees, SNOD symbol nodes,
    local heaps; header continuation blocks;
  * datasets with contiguous (or compact) layout, no filters: little/big-endian IEEE floats and fixed-point integers, scalar
    or N-d simple dataspaces;
  * attributes (version 1 / 2 / 3 messages) holding numbers or fixed-length strings (numpy ``S`` arrays: Keras' ``layer_names``,
    ``weight_names``, ``backend``, ``keras_version``).

Not implemented (raises ``Hdf5Error``): chunked / filtered datasets, new-style groups (link messages, fractal heaps), version-2
object headers, variable-length strings, superblock versions 2 and 3.  The writer emits exactly the subset the reader takes
(one B-tree node + one symbol node per group, sized through the superblock's group-leaf K).

    tree = read_hdf5(path)          # Group: .attrs {name: ndarray}, .children {name: Group | ndarray}
    write_hdf5(path, tree)
"""
from __future__ import annotations

import struct
from typing import Dict, Union

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF
MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = \
    0x0000, 0x0001, 0x0003, 0x0004, 0x0005, 0x0008, 0x000C, 0x0010, 0x0011
LEAF_K = 64          # writer: a symbol node holds up to 2 * LEAF_K links (stored in the superblock, honoured by readers)
INTERNAL_K = 16


class Hdf5Error(RuntimeError):
    pass


class Group:
    def __init__(self):
        self.attrs: Dict[str, np.ndarray] = {}
        self.children: Dict[str, Union['Group', np.ndarray]] = {}

    def __getitem__(self, path: str):
        node = self
        for part in path.strip('/').split('/'):
            if not isinstance(node, Group) or part not in node.children:
                raise KeyError(path)
            node = node.children[part]
        return node

    def require_group(self, path: str) -> 'Group':
        node = self
        for part in path.strip('/').split('/'):
            nxt = node.children.get(part)
            if nxt is None:
                nxt = node.children[part] = Group()
            if not isinstance(nxt, Group):
                raise Hdf5Error(f'{part} is a dataset')
            node = nxt
        return node

    def set_dataset(self, path: str, value):
        parts = path.strip('/').split('/')
        parent = self.require_group('/'.join(parts[:-1])) if len(parts) > 1 else self
        parent.children[parts[-1]] = np.asarray(value)


# ======================================================================================================================
# reader
# ======================================================================================================================
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != SIGNATURE:
            raise Hdf5Error('not an HDF5 file (signature missing at offset 0; user blocks are not supported)')
        ver = data[8]
        if ver not in (0, 1):
            raise Hdf5Error(f'superblock version {ver} is not supported (h5py libver="earliest" writes version 0)')
        if data[13] != 8 or data[14] != 8:
            raise Hdf5Error('only 8-byte offsets / lengths are supported')
        pos = 24 + (4 if ver == 1 else 0)
        self.base = self.u64(pos)
        self.root_entry = pos + 32

    def u16(self, o):
        return struct.unpack_from('<H', self.d, o)[0]

    def u32(self, o):
        return struct.unpack_from('<I', self.d, o)[0]

    def u64(self, o):
        return struct.unpack_from('<Q', self.d, o)[0]

    # ---- object headers
    def messages(self, addr: int):
        """[(type, flags, bytes)] of a version-1 object header at `addr`, continuation blocks included."""
        d = self.d
        if d[addr:addr + 4] == b'OHDR':
            raise Hdf5Error('version-2 object headers are not supported (file written with libver="latest")')
        if d[addr] != 1:
            raise Hdf5Error(f'object header version {d[addr]} at {addr:#x}')
        n_msgs = self.u16(addr + 2)
        blocks = [(addr + 16, self.u32(addr + 8))]
        out = []
        while blocks and len(out) < n_msgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = self.u16(pos), self.u16(pos + 2), d[pos + 4]
                body = d[pos + 8:pos + 8 + msize]
                if flags & 0x02:
                    raise Hdf5Error('shared header messages are not supported')
                if mtype == MSG_CONTINUATION:
                    blocks.append((self.base + struct.unpack_from('<Q', body, 0)[0], struct.unpack_from('<Q', body, 8)[0]))
                out.append((mtype, flags, body))
                pos += 8 + msize
        return out

    # ---- groups
    def heap_name(self, heap_addr: int, offset: int) -> str:
        if self.d[heap_addr:heap_addr + 4] != b'HEAP':
            raise Hdf5Error('bad local heap signature')
        seg = self.base + self.u64(heap_addr + 24)
        end = self.d.index(b'\x00', seg + offset)
        return self.d[seg + offset:end].decode('utf-8')

    def btree_entries(self, addr: int, heap: int):
        d = self.d
        if d[addr:addr + 4] != b'TREE':
            raise Hdf5Error('bad B-tree signature')
        if d[addr + 4] != 0:
            raise Hdf5Error('not a group B-tree')
        level, used = d[addr + 5], self.u16(addr + 6)
        out = []
        for i in range(used):
            child = self.base + self.u64(addr + 24 + 8 + i * 16)
            if level > 0:
                out.extend(self.btree_entries(child, heap))
            else:
                if d[child:child + 4] != b'SNOD':
                    raise Hdf5Error('bad symbol node signature')
                for s in range(self.u16(child + 6)):
                    e = child + 8 + s * 40
                    out.append((self.heap_name(heap, self.u64(e)), self.base + self.u64(e + 8)))
        return out

    # ---- datatypes / dataspaces
    @staticmethod
    def parse_dtype(b: bytes):
        cls, ver = b[0] & 0x0F, b[0] >> 4
        bits0 = b[1]
        size = struct.unpack_from('<I', b, 4)[0]
        order = '>' if bits0 & 1 else '<'
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f'float size {size}')
            return np.dtype(f'{order}f{size}')
        if cls == 0:
            return np.dtype(f'{order}{"i" if bits0 & 0x08 else "u"}{size}')
        if cls == 3:
            return np.dtype(f'S{size}')
        if cls == 9:
            raise Hdf5Error('variable-length datatypes are not supported')
        raise Hdf5Error(f'datatype class {cls} (version {ver}) is not supported')

    @staticmethod
    def parse_space(b: bytes):
        ver, rank, flags = b[0], b[1], b[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
            if b[3] == 2:
                raise Hdf5Error('null dataspace')
        else:
            raise Hdf5Error(f'dataspace version {ver}')
        return tuple(struct.unpack_from('<Q', b, off + 8 * i)[0] for i in range(rank))

    def parse_attribute(self, b: bytes):
        ver = b[0]
        name_sz, dt_sz, sp_sz = struct.unpack_from('<HHH', b, 2)
        if ver == 1:
            pad = lambda n: (n + 7) // 8 * 8  # noqa: E731
            pos = 8
        elif ver in (2, 3):
            pad = lambda n: n  # noqa: E731
            pos = 8 + (1 if ver == 3 else 0)
            if b[1] & 0x03:
                raise Hdf5Error('shared attribute datatype / dataspace')
        else:
            raise Hdf5Error(f'attribute message version {ver}')
        name = b[pos:pos + name_sz].split(b'\x00')[0].decode('utf-8')
        pos += pad(name_sz)
        dtype = self.parse_dtype(b[pos:pos + dt_sz])
        pos += pad(dt_sz)
        shape = self.parse_space(b[pos:pos + sp_sz])
        pos += pad(sp_sz)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dtype, count=n, offset=pos).reshape(shape)
        return name, arr.copy()

    # ---- objects
    def read_object(self, addr: int):
        msgs = self.messages(addr)
        attrs = {}
        sym = dtype = shape = layout = None
        for mtype, _, body in msgs:
            if mtype == MSG_SYMBOL_TABLE:
                sym = (self.base + struct.unpack_from('<Q', body, 0)[0], self.base + struct.unpack_from('<Q', body, 8)[0])
            elif mtype == MSG_DATATYPE:
                dtype = self.parse_dtype(body)
            elif mtype == MSG_DATASPACE:
                shape = self.parse_space(body)
            elif mtype == MSG_LAYOUT:
                layout = body
            elif mtype == MSG_ATTRIBUTE:
                try:
                    k, v = self.parse_attribute(body)
                    attrs[k] = v
                except Hdf5Error:
                    pass   # an attribute in an unsupported encoding (e.g. variable-length string) is skipped, not fatal
            elif mtype in (0x0002, 0x0006):
                raise Hdf5Error('new-style groups (link messages) are not supported: write the file with libver="earliest"')
        if sym is not None:
            g = Group()
            g.attrs = attrs
            for name, child in self.btree_entries(*sym):
                g.children[name] = self.read_object(child)
            return g
        if layout is None or dtype is None or shape is None:
            raise Hdf5Error(f'object at {addr:#x} is neither an old-style group nor a dataset')
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 1:
                a = struct.unpack_from('<Q', layout, 2)[0]
                if a == UNDEF:
                    arr = np.zeros(shape, dtype=dtype)
                else:
                    arr = np.frombuffer(self.d, dtype=dtype, count=n, offset=self.base + a).reshape(shape)
            elif cls == 0:
                sz = struct.unpack_from('<H', layout, 2)[0]
                arr = np.frombuffer(layout[4:4 + sz], dtype=dtype, count=n).reshape(shape)
            else:
                raise Hdf5Error('chunked datasets are not supported (Keras weight files are contiguous)')
        elif ver in (1, 2):
            rank, cls = layout[1], layout[2]
            if cls != 1:
                raise Hdf5Error('only contiguous version-1/2 layouts are supported')
            a = struct.unpack_from('<Q', layout, 8)[0]
            arr = np.frombuffer(self.d, dtype=dtype, count=n, offset=self.base + a).reshape(shape)
        else:
            raise Hdf5Error(f'data layout message version {ver}')
        arr = arr.astype(arr.dtype.newbyteorder('=')) if arr.dtype.byteorder == '>' else arr.copy()
        return arr

    def root(self) -> Group:
        obj = self.read_object(self.base + self.u64(self.root_entry + 8))
        if not isinstance(obj, Group):
            raise Hdf5Error('root object is not a group')
        return obj


def read_hdf5(path) -> Group:
    with open(path, 'rb') as f:
        return _Reader(f.read()).root()


# ======================================================================================================================
# writer
# ======================================================================================================================
def _pad8(b: bytes) -> bytes:
    return b + b'\x00' * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == 'f':
        size = dt.itemsize
        exp_bits, mant_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        head = bytes([0x11, 0x20, size * 8 - 1, 0x00]) + struct.pack('<I', size)
        props = struct.pack('<HHBBBBI', 0, size * 8, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1)
        return head + props
    if dt.kind in 'iu':
        head = bytes([0x10, 0x08 if dt.kind == 'i' else 0x00, 0x00, 0x00]) + struct.pack('<I', dt.itemsize)
        return head + struct.pack('<HH', 0, dt.itemsize * 8)
    if dt.kind == 'S':
        return bytes([0x13, 0x01, 0x00, 0x00]) + struct.pack('<I', dt.itemsize)   # fixed length, null padded, ASCII
    raise Hdf5Error(f'cannot write dtype {dt}')


def _space_msg(shape) -> bytes:
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b''.join(struct.pack('<Q', int(s)) for s in shape)


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack('<HHBBBB', mtype, len(body), 0, 0, 0, 0) + body


def _attr_msg(name: str, value) -> bytes:
    arr = np.asarray(value)
    if arr.dtype.kind == 'U':
        arr = np.char.encode(arr, 'utf-8')
    if arr.dtype.kind == 'S' and arr.dtype.itemsize == 0:
        arr = arr.astype('S1')
    nm = name.encode('utf-8') + b'\x00'
    dt, sp = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + np.asarray(arr, order='C').tobytes()
    if len(body) > 64000:
        raise Hdf5Error(f'attribute {name} exceeds the object-header limit (Keras splits such lists into chunks)')
    return _message(MSG_ATTRIBUTE, body)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)     # superblock placeholder

    def alloc(self, data: bytes) -> int:
        self.buf += b'\x00' * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, msgs) -> int:
        body = b''.join(msgs)
        head = struct.pack('<BBHII', 1, 0, len(msgs), 1, len(body)) + b'\x00' * 4
        return self.alloc(head + body)

    def write_dataset(self, arr: np.ndarray, attrs) -> int:
        arr = np.asarray(arr, order='C')   # (np.ascontiguousarray would turn a scalar dataset into shape (1,))
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [_message(MSG_DATASPACE, _space_msg(arr.shape)), _message(MSG_DATATYPE, _dtype_msg(arr.dtype)),
                _message(MSG_FILL, bytes([2, 2, 2, 0])),
                _message(MSG_LAYOUT, bytes([3, 1]) + struct.pack('<QQ', data_addr, len(raw)))]
        msgs += [_attr_msg(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)

    def write_group(self, g: Group):
        """-> (object header address, btree address, heap address)"""
        names = sorted(g.children, key=lambda s: s.encode('utf-8'))
        if len(names) > 2 * LEAF_K:
            raise Hdf5Error(f'a group with {len(names)} links needs more than one symbol node')
        child_addr = {}
        for n in names:
            c = g.children[n]
            child_addr[n] = self.write_group(c) if isinstance(c, Group) else (self.write_dataset(c, {}), None, None)
        # local heap: offset 0 holds the empty string, names follow 8-byte aligned
        heap = bytearray(8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode('utf-8') + b'\x00')
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b'HEAP' + bytes(4) + struct.pack('<QQQ', len(heap), 1, heap_data))
        # one symbol node with every link, one level-0 B-tree node above it
        snod = bytearray(b'SNOD' + bytes([1, 0]) + struct.pack('<H', len(names)))
        for n in names:
            oh, bt, hp = child_addr[n]
            if bt is not None:
                snod += struct.pack('<QQII', offs[n], oh, 1, 0) + struct.pack('<QQ', bt, hp)
            else:
                snod += struct.pack('<QQII', offs[n], oh, 0, 0) + bytes(16)
        snod += bytes(8 + 2 * LEAF_K * 40 - len(snod))
        snod_addr = self.alloc(bytes(snod))
        tree = bytearray(b'TREE' + bytes([0, 0]) + struct.pack('<H', 1 if names else 0) + struct.pack('<QQ', UNDEF, UNDEF))
        if names:
            tree += struct.pack('<QQQ', 0, snod_addr, offs[names[-1]])
        tree += bytes(24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8 - len(tree))
        tree_addr = self.alloc(bytes(tree))
        msgs = [_message(MSG_SYMBOL_TABLE, struct.pack('<QQ', tree_addr, heap_addr))] + [_attr_msg(k, v) for k, v in g.attrs.items()]
        return self.object_header(msgs), tree_addr, heap_addr

    def finish(self, root: Group) -> bytes:
        oh, bt, hp = self.write_group(root)
        self.buf += b'\x00' * (-len(self.buf) % 8)
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack('<HHI', LEAF_K, INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, oh, 1, 0) + struct.pack('<QQ', bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(path, root: Group):
    data = _Writer().finish(root)
    with open(path, 'wb') as f:
        f.write(data)
