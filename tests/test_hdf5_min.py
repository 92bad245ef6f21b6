"""The minimal pure-Python HDF5 writer behind the episode logger (gym_kmanip_b200/hdf5_min.py; reference
gym_kmanip/log_h5py.py:13-61 uses h5py, which is not in the build image).  Read back with h5py iff it imports; otherwise
with the module's own reader plus byte-level checks of the structures the HDF5 File Format Specification prescribes for
this layout (version 0 superblock, version 1 object headers, symbol-table groups, contiguous datasets)."""
import struct

import numpy as np
import pytest

from gym_kmanip_b200 import hdf5_min


def _tree(rng):
    return {"action": rng.standard_normal((64, 3)).astype(np.float32),
            "observations": {"qpos": rng.standard_normal((64, 10)).astype(np.float32), "qvel": rng.standard_normal((64, 10)).astype(np.float32),
                             "images": {"head": rng.integers(0, 255, (2, 4, 6, 3), dtype=np.uint8)}},
            "metadata": {"head": {}}, "counts": np.arange(5, dtype=np.int64), "f64": rng.standard_normal(7)}


ATTRS = {"": {"sim": True}, "metadata": {"q_len": 10, "a_len": 3, "dt": 0.02, "name": "KManipSoloArm", "mask": [1, 2, 3]},
         "metadata/head": {"resolution": [640, 480], "focal_length": 448.0, "principal_point": [320.0, 240.0]}}


def test_round_trip_and_structure(tmp_path):
    rng = np.random.default_rng(0)
    tree, path = _tree(rng), str(tmp_path / "episode_1.hdf5")
    hdf5_min.write(path, tree, ATTRS)
    got, attrs = hdf5_min.read(path)

    def same(a, b):
        if isinstance(a, dict):
            assert isinstance(b, dict) and sorted(a) == sorted(b)
            for k in a:
                same(a[k], b[k])
        else:
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    same(tree, got)
    assert attrs[""]["sim"] == 1 and attrs["metadata"]["q_len"] == 10 and attrs["metadata"]["dt"] == 0.02
    assert attrs["metadata"]["name"] == "KManipSoloArm" and list(attrs["metadata"]["mask"]) == [1, 2, 3]
    assert list(attrs["metadata/head"]["resolution"]) == [640, 480] and attrs["metadata/head"]["focal_length"] == 448.0

    buf = open(path, "rb").read()
    # superblock v0: signature, versions 0, 8-byte offsets and lengths, K values, base address 0, end-of-file address
    assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8:13] == b"\0\0\0\0\0" and buf[13] == 8 and buf[14] == 8
    leaf_k, int_k, flags = struct.unpack_from("<HHI", buf, 16)
    assert (leaf_k, int_k, flags) == (hdf5_min.LEAF_K, hdf5_min.INTERNAL_K, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", buf, 24)
    assert base == 0 and free == hdf5_min.UNDEF and eof == len(buf) and drv == hdf5_min.UNDEF and len(buf) % 8 == 0
    name_off, root_hdr, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", buf, 56)
    assert name_off == 0 and cache == 1 and root_hdr % 8 == 0 and bt % 8 == 0 and hp % 8 == 0
    # root object header v1: its first message is the symbol-table message pointing at the same B-tree and heap
    ver, _, nmsg, refs, hsize = struct.unpack_from("<BBHII", buf, root_hdr)
    assert ver == 1 and refs == 1 and nmsg == 2 and hsize % 8 == 0              # symbol table + the "sim" attribute
    mtype, msize = struct.unpack_from("<HH", buf, root_hdr + 16)
    assert mtype == 0x0011 and msize == 16 and struct.unpack_from("<QQ", buf, root_hdr + 24) == (bt, hp)
    # B-tree node: group type, leaf level, one child; key 0 is the empty string, key 1 the largest name of the child
    assert buf[bt:bt + 4] == b"TREE" and buf[bt + 4] == 0 and buf[bt + 5] == 0 and struct.unpack_from("<H", buf, bt + 6)[0] == 1
    assert struct.unpack_from("<QQ", buf, bt + 8) == (hdf5_min.UNDEF, hdf5_min.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", buf, bt + 24)
    assert buf[hp:hp + 4] == b"HEAP" and buf[hp + 4] == 0
    seg_size, free_off, seg = struct.unpack_from("<QQQ", buf, hp + 8)
    assert key0 == 0 and buf[seg] == 0                                         # "" at heap offset 0
    # free list of the heap: one block inside the segment, terminated by H5HL_FREE_NULL
    assert free_off + 16 <= seg_size and struct.unpack_from("<QQ", buf, seg + free_off) == (1, seg_size - free_off)
    # symbol node: version 1, names strictly increasing (the library searches it by bisection)
    assert buf[snod:snod + 4] == b"SNOD" and buf[snod + 4] == 1
    nsym = struct.unpack_from("<H", buf, snod + 6)[0]
    names = []
    for k in range(nsym):
        off, hdr, ctype = struct.unpack_from("<QQI", buf, snod + 8 + 40 * k)
        end = buf.index(b"\0", seg + off)
        names.append(buf[seg + off:end])
        assert off % 8 == 0 and hdr % 8 == 0 and ctype == (1 if isinstance(tree[names[-1].decode()], dict) else 0)
    assert names == sorted(names) == sorted(k.encode() for k in tree) and len(set(names)) == len(names)
    end = buf.index(b"\0", seg + key1)
    assert buf[seg + key1:end] == names[-1]
    # a dataset header: dataspace v1, datatype (IEEE float32 little endian), fill value v2, contiguous layout v3 -> raw data
    a_hdr = [struct.unpack_from("<QQ", buf, snod + 8 + 40 * k)[1] for k in range(nsym) if names[k] == b"action"][0]
    msgs = list(hdf5_min._messages(buf, a_hdr))
    assert [m[0] for m in msgs] == [0x0001, 0x0003, 0x0005, 0x0008]
    assert msgs[0][1][:2] == b"\x01\x02" and struct.unpack_from("<QQ", msgs[0][1], 8) == (64, 3)
    assert msgs[1][1][:8] == bytes([0x11, 0x20, 31, 0, 4, 0, 0, 0]) and struct.unpack_from("<HHBBBBI", msgs[1][1], 8) == (0, 32, 23, 8, 0, 23, 127)
    ver, cls, addr, size = struct.unpack_from("<BBQQ", msgs[3][1], 0)
    assert (ver, cls, size) == (3, 1, 64 * 3 * 4) and addr % 8 == 0
    assert np.array_equal(np.frombuffer(buf, "<f4", 64 * 3, addr).reshape(64, 3), tree["action"])


def test_h5py_reads_it_when_available(tmp_path):
    h5py = pytest.importorskip("h5py")
    rng = np.random.default_rng(1)
    tree, path = _tree(rng), str(tmp_path / "episode_2.hdf5")
    hdf5_min.write(path, tree, ATTRS)
    with h5py.File(path, "r") as f:
        assert np.array_equal(f["action"][:], tree["action"]) and np.array_equal(f["observations/qpos"][:], tree["observations"]["qpos"])
        assert np.array_equal(f["observations/images/head"][:], tree["observations"]["images"]["head"])
        assert f.attrs["sim"] and f["metadata"].attrs["q_len"] == 10 and list(f["metadata/head"].attrs["resolution"]) == [640, 480]
        assert sorted(f["observations"].keys()) == ["images", "qpos", "qvel"]


def test_limits():
    with pytest.raises(ValueError):
        hdf5_min._Writer().group({f"d{i}": np.zeros(1) for i in range(2 * hdf5_min.LEAF_K + 1)}, {}, "")
    with pytest.raises(TypeError):
        hdf5_min._attribute("x", {"nested": 1})
