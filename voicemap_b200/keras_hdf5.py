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
