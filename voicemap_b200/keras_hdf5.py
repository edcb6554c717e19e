"""Minimal pure-Python reader for Keras 2.2.x full-model HDF5 checkpoints, so that weights
saved by the reference's ``ModelCheckpoint`` (experiments/train_siamese.py:81-87) load into the B200 models without
h5py.  Supports exactly what those files use: superblock version 0, old-style groups (v1 B-trees + local heaps +
symbol-table nodes), version-1 object headers with continuation blocks, contiguous little-endian float/int
datasets, and attributes holding fixed-length / variable-length strings or string arrays.

    f = KerasH5(path)
    f.attr("/", "model_config")                 -> JSON text
    f.datasets("/model_weights")                -> {"conv1d_1/conv1d_1/kernel:0": ndarray, ...}
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class KerasH5:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise NotImplementedError("only superblock version 0 with 8-byte offsets/lengths is supported")
        self.root = struct.unpack_from("<Q", b, 64)[0]  # root symbol-table entry: object header address

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """Yield (type, payload bytes) of a version-1 object header, following continuation messages."""
        b = self.b
        if b[addr] != 1:
            raise NotImplementedError("only version-1 object headers are supported")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        size = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                out.append((mtype, data))
        return out

    # ------------------------------------------------------------------ groups
    def _heap_string(self, heap_addr, offset):
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        s = data_addr + offset
        return b[s:b.index(b"\x00", s)].decode()

    def _btree_entries(self, btree, heap):
        b = self.b
        assert b[btree:btree + 4] == b"TREE"
        level, used = b[btree + 5], struct.unpack_from("<H", b, btree + 6)[0]
        pos = btree + 24  # keys and children alternate: key0, child0, key1, ...
        children = [struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0] for i in range(used)]
        out = {}
        for child in children:
            if level > 0:
                out.update(self._btree_entries(child, heap))
                continue
            assert b[child:child + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            for i in range(nsym):
                e = child + 8 + 40 * i
                name_off, obj = struct.unpack_from("<QQ", b, e)
                out[self._heap_string(heap, name_off)] = obj
        return out

    def children(self, addr):
        """name -> object header address for a group (empty dict for a dataset)."""
        for mtype, data in self._messages(addr):
            if mtype == 0x11:  # symbol table message
                btree, heap = struct.unpack_from("<QQ", data, 0)
                return self._btree_entries(btree, heap)
        return {}

    def resolve(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            kids = self.children(addr)
            if part not in kids:
                raise KeyError(path)
            addr = kids[part]
        return addr

    # ------------------------------------------------------------------ datatypes / dataspaces
    @staticmethod
    def _dtype(data):
        cls = data[0] & 0x0F
        size = struct.unpack_from("<I", data, 4)[0]
        if cls == 1:
            return np.dtype(f"<f{size}"), size
        if cls == 0:
            signed = (data[1] >> 3) & 1
            return np.dtype(f"<{'i' if signed else 'u'}{size}"), size
        if cls == 3:
            return ("str", size), size
        if cls == 9:  # variable length; base type follows (we only need vlen strings)
            return ("vlen", size), size
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    @staticmethod
    def _shape(data):
        ver, rank = data[0], data[1]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))

    def _global_heap_object(self, collection, index):
        b = self.b
        assert b[collection:collection + 4] == b"GCOL"
        pos = collection + 16
        while True:
            idx, _ref, _, size = struct.unpack_from("<HHIQ", b, pos)
            if idx == index:
                return b[pos + 16:pos + 16 + size]
            if idx == 0:
                raise KeyError("global heap object not found")
            pos += 16 + (size + 7) // 8 * 8

    def _decode(self, dtype, shape, raw):
        kind = dtype[0]
        n = int(np.prod(shape)) if shape else 1
        if isinstance(kind, np.dtype):
            arr = np.frombuffer(raw, dtype=kind, count=n).copy()
            return arr.reshape(shape) if shape else arr[0]
        (tag, size) = kind
        vals = []
        for i in range(n):
            item = raw[i * size:(i + 1) * size]
            if tag == "str":
                vals.append(item.split(b"\x00")[0].decode())
            else:  # vlen string: length(4) + global heap collection address(8) + object index(4)
                _ln, coll, idx = struct.unpack_from("<IQI", item, 0)
                vals.append(self._global_heap_object(coll, idx).decode())
        return vals if shape else vals[0]

    # ------------------------------------------------------------------ attributes / datasets
    def attrs(self, path):
        out = {}
        for mtype, data in self._messages(self.resolve(path)):
            if mtype != 0x0C:
                continue
            ver = data[0]
            name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", data, 2)
            pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
            pos = 8
            name = data[pos:pos + name_sz].split(b"\x00")[0].decode()
            pos += pad(name_sz)
            dt = data[pos:pos + dt_sz]
            pos += pad(dt_sz)
            ds = data[pos:pos + ds_sz]
            pos += pad(ds_sz)
            out[name] = self._decode(self._dtype(dt), self._shape(ds), data[pos:])
        return out

    def attr(self, path, name):
        return self.attrs(path)[name]

    def dataset(self, addr_or_path):
        addr = self.resolve(addr_or_path) if isinstance(addr_or_path, str) else addr_or_path
        dtype = shape = None
        data_addr = size = None
        for mtype, data in self._messages(addr):
            if mtype == 0x03:
                dtype = self._dtype(data)
            elif mtype == 0x01:
                shape = self._shape(data)
            elif mtype == 0x08:
                ver = data[0]
                if ver == 3:
                    if data[1] != 1:
                        raise NotImplementedError("only contiguous dataset layout is supported")
                    data_addr, size = struct.unpack_from("<QQ", data, 2)
                else:
                    raise NotImplementedError(f"data layout message version {ver}")
        if dtype is None or shape is None or data_addr is None:
            raise KeyError("not a dataset")
        if data_addr == UNDEF:
            return np.zeros(shape, dtype=dtype[0])
        return self._decode(dtype, shape, self.b[data_addr:data_addr + size])

    def datasets(self, path="/"):
        """Recursively collect all datasets below ``path``: relative name -> ndarray."""
        out = {}

        def walk(addr, prefix):
            kids = self.children(addr)
            if not kids:
                try:
                    out[prefix] = self.dataset(addr)
                except KeyError:
                    pass
                return
            for name, child in kids.items():
                walk(child, f"{prefix}/{name}" if prefix else name)

        walk(self.resolve(path), "")
        return out


def load_keras_weights(path):
    """(model_config dict, ordered list of (layer name, [(weight name, ndarray), ...])) of a Keras 2.2.x HDF5
    checkpoint, in ``model.get_weights()`` order."""
    import json
    f = KerasH5(path)
    cfg = json.loads(f.attr("/", "model_config"))
    layers = []
    for lname in f.attr("/model_weights", "layer_names"):
        wnames = f.attr(f"/model_weights/{lname}", "weight_names")
        layers.append((lname, [(w, f.dataset(f"/model_weights/{lname}/{w}")) for w in wnames]))
    return cfg, layers


def load_training_config(path):
    """The ``training_config`` attribute of a full-model checkpoint (loss, metrics, optimizer_config), or None --
    what keras.models.load_model uses to hand back a COMPILED model."""
    import json
    f = KerasH5(path)
    try:
        raw = f.attr("/", "training_config")
    except (KeyError, ValueError):
        return None
    return json.loads(raw) if raw else None


# ======================================================================================================================
# Writer: the same subset of HDF5 the reader understands, laid out the way libhdf5 1.8/1.10 writes a Keras 2.2.x
# ``model.save`` file (superblock v0, old-style groups = v1 B-tree + local heap + one symbol-table node, version-1 object
# headers in a single chunk, contiguous little-endian float32 datasets, fixed-length null-padded string attributes).
# Structure of a checkpoint (keras/engine/saving.py of Keras 2.2.2, the version the reference pins):
#   /            attrs keras_version, backend, model_config (JSON), training_config (JSON, optional)
#   /model_weights            attrs layer_names [..], backend, keras_version
#   /model_weights/<layer>    attr  weight_names ['conv1d_1/kernel:0', ...]; datasets at <layer>/<weight_name>
# No libhdf5 / h5py exists in the build image: files are verified by reading them back with ``KerasH5`` (which was
# written against, and is tested on, a real Keras file) and by structural checks in tests/test_host.py.
# ======================================================================================================================
class _Group:
    def __init__(self):
        self.attrs = []        # (name, value): str -> scalar string, list[str] -> 1-D string array
        self.children = {}     # name -> _Group | np.ndarray


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype, payload):
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), 0) + payload


def _attr_message(name, value):
    nm = name.encode() + b"\x00"
    if isinstance(value, (list, tuple)):
        items = [v.encode() for v in value]
        size = max([len(i) for i in items] + [1])
        space = struct.pack("<BBB5xQQ", 1, 1, 1, len(items), len(items))   # rank 1, max dims present
        data = b"".join(i.ljust(size, b"\x00") for i in items)
    else:
        raw = value.encode()
        size = max(len(raw), 1)
        space = struct.pack("<BBB5x", 1, 0, 0)                             # scalar
        data = raw.ljust(size, b"\x00")
    dtype = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, size)                  # class 3 string v1, null-padded, ASCII
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype), len(space)) + _pad8(nm) + _pad8(dtype) + _pad8(space) + data
    if len(body) > 0xFFF0:
        raise ValueError(f"attribute {name!r} does not fit a version-1 object-header message (64 KB)")
    return _message(0x0C, body)


class _H5Writer:
    INTERNAL_K = 16

    def __init__(self, root):
        self.buf = bytearray(96)            # superblock is written last
        widest = [1]

        def scan(g):
            widest[0] = max(widest[0], len(g.children))
            for c in g.children.values():
                if isinstance(c, _Group):
                    scan(c)
        scan(root)
        self.leaf_k = max(4, (widest[0] + 1) // 2)   # one symbol-table node (2K entries) holds any group of this file
        hdr, btree, heap = self._group(root)
        eof = len(self.buf)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)   # root symbol-table entry
        assert len(sb) == 96
        self.buf[:96] = sb

    def _alloc(self, data):
        assert len(self.buf) % 8 == 0
        addr = len(self.buf)
        self.buf += _pad8(bytes(data))
        return addr

    def _dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        dims = arr.shape
        space = struct.pack("<BBB5x", 1, len(dims), 1 if dims else 0)
        space += b"".join(struct.pack("<Q", d) for d in dims) * (2 if dims else 0)   # dims then max dims
        # IEEE little-endian float32: class 1 v1; bit offset 0, precision 32, exponent at 23 (8 bits), mantissa 23 bits
        dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)            # v2: late allocation, fill written if set, size 0
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        msgs = _message(0x01, space) + _message(0x03, dtype) + _message(0x05, fill) + _message(0x08, layout)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, 4, 1, len(msgs)) + msgs)

    def _group(self, g):
        names = sorted(g.children, key=lambda s: s.encode())
        entries = []
        for name in names:
            child = g.children[name]
            if isinstance(child, _Group):
                entries.append((name, *self._group(child)))
            else:
                entries.append((name, self._dataset(child), None, None))
        # local heap: "" at offset 0, the link names, one trailing free block {next = 1 (none), size}
        seg = bytearray(8)
        offsets = []
        for name in names:
            offsets.append(len(seg))
            seg += _pad8(name.encode() + b"\x00")
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 32) + bytes(16)
        seg_addr = self._alloc(seg)
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for (name, hdr, bt, hp), off in zip(entries, offsets):
            if bt is None:
                snod += struct.pack("<QQII16x", off, hdr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", off, hdr, 1, 0, bt, hp)
        snod += bytes(8 + 2 * self.leaf_k * 40 - len(snod))
        snod_addr = self._alloc(snod)
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            tree += struct.pack("<QQQ", 0, snod_addr, offsets[-1])     # key0 = "", child, key1 = largest name
        tree += bytes(24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8 - len(tree))
        btree = self._alloc(tree)
        msgs = _message(0x11, struct.pack("<QQ", btree, heap))
        for name, value in g.attrs:
            msgs += _attr_message(name, value)
        nmsg = 1 + len(g.attrs)
        hdr = self._alloc(struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)
        return hdr, btree, heap


def save_keras_weights(path, model_config, layers, training_config=None, keras_version="2.2.2", backend="tensorflow"):
    """Write a Keras-2.2.x-style full-model checkpoint.  ``model_config`` / ``training_config``: JSON-serialisable
    dicts; ``layers``: [(layer name, [(weight name such as 'conv1d_1/kernel:0', ndarray), ...]), ...] in
    ``model.layers`` order (weight-less layers carry an empty list), i.e. what ``load_keras_weights`` returns."""
    import json
    root = _Group()
    root.attrs = [("keras_version", keras_version), ("backend", backend), ("model_config", json.dumps(model_config))]
    if training_config is not None:
        root.attrs.append(("training_config", json.dumps(training_config)))
    mw = _Group()
    mw.attrs = [("layer_names", [n for n, _ in layers]), ("backend", backend), ("keras_version", keras_version)]
    for lname, weights in layers:
        lg = _Group()
        lg.attrs = [("weight_names", [w for w, _ in weights])]
        for wname, arr in weights:
            node = lg
            parts = wname.split("/")
            for part in parts[:-1]:
                node = node.children.setdefault(part, _Group())
            node.children[parts[-1]] = np.asarray(arr)
        mw.children[lname] = lg
    root.children["model_weights"] = mw
    w = _H5Writer(root)
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))
