"""Minimal pure-Python HDF5 writer (and a reader for exactly what it writes): no h5py in the build image.

The reference logs episodes with h5py (reference gym_kmanip/log_h5py.py:13-61: groups, float32 / uint8 datasets, scalar and
list attributes).  This module writes the same tree in the oldest, most widely readable layout of the HDF5 File Format
Specification (version 0 superblock, version 1 object headers, symbol-table groups = v1 B-tree + local heap + one symbol
node per group, contiguous datasets, version 1 attribute messages) -- the layout every libhdf5 release understands.  The
build image has neither h5py nor libhdf5, so here the files are checked by the reader below and by byte-level tests of
the structures the specification prescribes (tests/test_hdf5_min.py); the h5py read-back test runs wherever h5py imports:

    write(path, {"action": ndarray, "observations": {"qpos": ndarray, "images": {}}, "metadata": {}},
          attrs={"": {"sim": True}, "metadata": {"q_len": 9, "name": "KManipSoloArm"}})

Groups are dicts, datasets are numpy arrays (float32/64, int8..64, uint8..64), attributes are bool / int / float / str or
lists / arrays of numbers; attribute dict keys are group paths ("" = root).  Every group holds at most 2 * LEAF_K entries
(one symbol node; LEAF_K is recorded in the superblock, so the limit is a property of the file, not of the format).
Not supported, on purpose: chunking, compression, variable-length types, links, resizing.
"""
import struct
from typing import Any, Dict, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 32, 16            # group leaf / internal node K of this file (superblock fields)
FREE_NULL = 1                          # libhdf5's H5HL_FREE_NULL: end of a local heap's free list


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------ datatype / dataspace
def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # class 1 (floating point), version 1; little endian, implied-msb mantissa normalisation, sign bit location
        eloc, esz, msz, bias = (23, 8, 23, 127) if dt.itemsize == 4 else (52, 11, 52, 1023)
        return struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, 0, msz, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        # class 0 (fixed point), version 1; little endian, bit 3 = signed
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        # class 3 (string), version 1; null padded, UTF-8
        return struct.pack("<BBBBI", 0x13, 0x11, 0, 0, dt.itemsize)
    raise TypeError(f"unsupported dtype {dt}")


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _attr_value(v: Any) -> np.ndarray:
    if isinstance(v, (bool, np.bool_)):
        return np.array(int(v), dtype=np.uint8)          # (h5py would write an enum; readers test truthiness)
    if isinstance(v, str):
        raw = v.encode("utf-8")
        return np.array(raw, dtype=f"S{max(1, len(raw))}")
    a = np.asarray(v)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype.kind == "i":
        a = a.astype(np.int64)
    if a.dtype.kind not in "fiuS":
        raise TypeError(f"unsupported attribute value {v!r}")
    return a


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _attribute(name: str, value: Any) -> bytes:
    a = _attr_value(value)
    nm = name.encode("utf-8") + b"\0"
    dt, ds = _datatype(a.dtype), _dataspace(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.astype(a.dtype.newbyteorder("<")).tobytes()
    return _message(0x000C, body)


def _object_header(messages) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


# ------------------------------------------------------------------------------------------------ writer
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)        # superblock, filled in at the end

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [_message(0x0001, _dataspace(arr.shape)),
                _message(0x0003, _datatype(arr.dtype)),
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),      # fill value v2: late allocation, write if set, default (size 0)
                _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]   # layout v3, contiguous
        return self.alloc(_object_header(msgs))

    def group(self, tree: Dict[str, Any], attrs: Dict[str, Dict[str, Any]], path: str) -> Tuple[int, int, int]:
        """Writes the children, then heap, symbol node, B-tree and object header of this group; returns (header, btree, heap)."""
        names = sorted(tree, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * LEAF_K:
            raise ValueError(f"group '{path}' has {len(names)} entries; this writer holds at most {2 * LEAF_K} per group")
        entries = []
        for nm in names:
            child, cpath = tree[nm], f"{path}/{nm}" if path else nm
            if isinstance(child, dict):
                entries.append((nm,) + self.group(child, attrs, cpath))
            else:
                entries.append((nm, self.dataset(np.asarray(child)), None, None))
        # local heap data segment: "" at offset 0, the names (null terminated, 8-byte aligned), one free block at the end
        seg, offs = bytearray(8), {}
        for nm in names:
            offs[nm] = len(seg)
            seg += _pad8(nm.encode("utf-8") + b"\0")
        free_off = len(seg)
        seg += struct.pack("<QQ", FREE_NULL, 32) + b"\0" * 16
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # symbol node: entries sorted by name, all 2 * LEAF_K slots allocated
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)))
        for nm, hdr, bt, hp in entries:
            if bt is None:
                snod += struct.pack("<QQII16x", offs[nm], hdr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", offs[nm], hdr, 1, 0, bt, hp)    # cache type 1: B-tree and heap of the sub-group
        snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
        snod_addr = self.alloc(bytes(snod))
        # B-tree v1, group node type, level 0: keys are heap offsets; child i holds names in (key[i], key[i + 1]]
        bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF))
        if entries:
            bt += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])
        bt += b"\0" * (24 + (2 * INTERNAL_K) * 8 + (2 * INTERNAL_K + 1) * 8 - len(bt))
        bt_addr = self.alloc(bytes(bt))
        msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_attribute(k, v) for k, v in attrs.get(path, {}).items()]
        return self.alloc(_object_header(msgs)), bt_addr, heap_addr

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        hdr, bt, hp = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, bt, hp)                      # root group symbol table entry
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write(path: str, tree: Dict[str, Any], attrs: Dict[str, Dict[str, Any]] = None) -> None:
    w = _Writer()
    blob = w.finish(w.group(tree, attrs or {}, ""))
    with open(path, "wb") as f:
        f.write(blob)


# ------------------------------------------------------------------------------------------------ reader (of the subset above)
def _read_dtype(b: bytes) -> np.dtype:
    cls, ver = b[0] & 0x0F, b[0] >> 4
    size = struct.unpack_from("<I", b, 4)[0]
    if ver != 1:
        raise ValueError("datatype version")
    if cls == 1:
        return np.dtype(f"<f{size}")
    if cls == 0:
        return np.dtype(("<i" if b[1] & 0x08 else "<u") + str(size))
    if cls == 3:
        return np.dtype(f"S{size}")
    raise ValueError(f"datatype class {cls}")


def _read_space(b: bytes) -> Tuple[int, ...]:
    ver, rank = b[0], b[1]
    if ver != 1:
        raise ValueError("dataspace version")
    return tuple(struct.unpack_from("<Q", b, 8 + 8 * i)[0] for i in range(rank))


def _messages(buf: bytes, addr: int):
    ver, _, n, _refs, size = struct.unpack_from("<BBHII", buf, addr)
    if ver != 1:
        raise ValueError("object header version")
    p, end = addr + 16, addr + 16 + size
    for _ in range(n):
        mtype, msize, _flags = struct.unpack_from("<HHB", buf, p)
        yield mtype, buf[p + 8: p + 8 + msize]
        p += 8 + msize
    assert p == end


def _read_object(buf: bytes, addr: int):
    """Returns (kind, value, attrs): kind 'group' -> value = {name: object address}; 'dataset' -> value = ndarray."""
    attrs, space, dtype, layout, sym = {}, None, None, None, None
    for mtype, d in _messages(buf, addr):
        if mtype == 0x0001:
            space = _read_space(d)
        elif mtype == 0x0003:
            dtype = _read_dtype(d)
        elif mtype == 0x0008:
            ver, cls, daddr, dsize = struct.unpack_from("<BBQQ", d, 0)
            assert ver == 3 and cls == 1
            layout = (daddr, dsize)
        elif mtype == 0x0011:
            sym = struct.unpack_from("<QQ", d, 0)
        elif mtype == 0x000C:
            ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", d, 0)
            assert ver == 1
            p = 8
            name = d[p: p + nsz - 1].decode("utf-8"); p += (nsz + 7) // 8 * 8
            adt = _read_dtype(d[p: p + tsz]); p += (tsz + 7) // 8 * 8
            ash = _read_space(d[p: p + ssz]); p += (ssz + 7) // 8 * 8
            cnt = int(np.prod(ash)) if ash else 1
            val = np.frombuffer(d, dtype=adt, count=cnt, offset=p).reshape(ash)
            attrs[name] = val.item() if not ash else val.copy()
            if isinstance(attrs[name], bytes):
                attrs[name] = attrs[name].decode("utf-8")
    if sym is not None:
        bt, hp = sym
        assert buf[hp: hp + 4] == b"HEAP"
        seg_size, _free, seg_addr = struct.unpack_from("<QQQ", buf, hp + 8)
        assert buf[bt: bt + 4] == b"TREE" and buf[bt + 4] == 0 and buf[bt + 5] == 0
        used = struct.unpack_from("<H", buf, bt + 6)[0]
        children = {}
        for i in range(used):
            snod = struct.unpack_from("<Q", buf, bt + 24 + 16 * i + 8)[0]
            assert buf[snod: snod + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", buf, snod + 6)[0]
            for k in range(nsym):
                noff, oaddr = struct.unpack_from("<QQ", buf, snod + 8 + 40 * k)
                end = buf.index(b"\0", seg_addr + noff)
                assert end < seg_addr + seg_size
                children[buf[seg_addr + noff: end].decode("utf-8")] = oaddr
        return "group", children, attrs
    daddr, dsize = layout
    cnt = int(np.prod(space)) if space else 1
    assert dsize == cnt * dtype.itemsize
    arr = np.frombuffer(buf, dtype=dtype, count=cnt, offset=daddr).reshape(space).copy() if cnt else np.zeros(space, dtype)
    return "dataset", arr, attrs


def read(path: str) -> Tuple[Dict[str, Any], Dict[str, Dict[str, Any]]]:
    """Inverse of write(): (tree, attrs) with groups as dicts and datasets as arrays."""
    buf = open(path, "rb").read()
    assert buf[:8] == SIGNATURE and buf[8] == 0 and buf[13] == 8 and buf[14] == 8
    assert struct.unpack_from("<Q", buf, 40)[0] == len(buf)            # end-of-file address
    root = struct.unpack_from("<Q", buf, 56 + 8)[0]
    attrs: Dict[str, Dict[str, Any]] = {}

    def walk(addr, path):
        kind, val, a = _read_object(buf, addr)
        if a:
            attrs[path] = a
        if kind == "dataset":
            return val
        return {nm: walk(child, f"{path}/{nm}" if path else nm) for nm, child in val.items()}

    return walk(root, ""), attrs
