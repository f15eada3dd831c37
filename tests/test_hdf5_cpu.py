"""Checkpoint container: the built-in HDF5 writer/reader behind mlf_hdf5_createFile/openFile/pushState
(src/mlf_hdf5.f90:153-225,671-737).  An independent pure-Python walk of the file (written from the HDF5 File Format
Specification, not from the C++ code) checks the on-disk structure: superblock v0, v1 object headers, symbol-table
root group, contiguous little-endian int64/float64 datasets.  Interoperability with stock libhdf5 cannot be checked
in this image (no libhdf5, no h5py) and is stated as unverified in DESIGN.md."""
import struct

import numpy as np

from mlfortran_b200 import _lib, mlf_hdf5_file

UNDEF = 0xFFFFFFFFFFFFFFFF


def _u(b, off, n):
    return int.from_bytes(b[off:off + n], "little")


def parse_hdf5(path):
    """Independent reader: returns {name: (dtype, dims, ndarray)}."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8] == 0 and b[13] == 8 and b[14] == 8          # superblock v0, 8-byte offsets/lengths
    leaf_k, internal_k = _u(b, 16, 2), _u(b, 18, 2)
    assert _u(b, 24, 8) == 0                                  # base address
    assert _u(b, 40, 8) == len(b)                             # end-of-file address
    root_oh, cache_type = _u(b, 64, 8), _u(b, 72, 4)
    assert cache_type == 1
    btree, heap = _u(b, 80, 8), _u(b, 88, 8)
    # root object header: version 1, one symbol-table message pointing at the same B-tree / heap
    assert b[root_oh] == 1 and _u(b, root_oh + 2, 2) == 1
    assert _u(b, root_oh + 16, 2) == 0x0011
    assert (_u(b, root_oh + 24, 8), _u(b, root_oh + 32, 8)) == (btree, heap)
    assert b[heap:heap + 4] == b"HEAP"
    heap_size, free_off, heap_data = _u(b, heap + 8, 8), _u(b, heap + 16, 8), _u(b, heap + 24, 8)
    assert free_off < heap_size and _u(b, heap_data + free_off, 8) == 1  # one trailing free block, end of list
    assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[btree + 5] == 0
    used = _u(b, btree + 6, 2)
    assert _u(b, btree + 8, 8) == UNDEF and _u(b, btree + 16, 8) == UNDEF
    out, names_in_order = {}, []
    p = btree + 24
    for i in range(used):
        child = _u(b, p + 8, 8)
        key_next = _u(b, p + 16, 8)
        p += 16
        assert b[child:child + 4] == b"SNOD" and b[child + 4] == 1
        ns = _u(b, child + 6, 2)
        assert 1 <= ns <= 2 * leaf_k
        last_name = None
        for s in range(ns):
            e = child + 8 + 40 * s
            noff, oh = _u(b, e, 8), _u(b, e + 8, 8)
            nm = b[heap_data + noff:b.index(b"\0", heap_data + noff)].decode()
            names_in_order.append(nm)
            last_name = noff
            assert oh % 8 == 0 and b[oh] == 1
            nmsg, hsize = _u(b, oh + 2, 2), _u(b, oh + 8, 4)
            q, end = oh + 16, oh + 16 + hsize
            dims = dtype = addr = size = None
            for _ in range(nmsg):
                mtype, msize = _u(b, q, 2), _u(b, q + 2, 2)
                assert msize % 8 == 0
                m = q + 8
                if mtype == 0x0001:
                    assert b[m] == 1
                    dims = [_u(b, m + 8 + 8 * j, 8) for j in range(b[m + 1])]
                elif mtype == 0x0003:
                    cls, ver = b[m] & 0xF, b[m] >> 4
                    assert ver == 1 and _u(b, m + 4, 4) == 8 and (b[m + 1] & 1) == 0
                    if cls == 0:
                        assert b[m + 1] & 0x08 and _u(b, m + 10, 2) == 64   # signed, 64-bit precision
                        dtype = np.int64
                    else:
                        assert cls == 1 and b[m + 2] == 63                     # sign bit position
                        assert (b[m + 12], b[m + 13], b[m + 14], b[m + 15]) == (52, 11, 0, 52) and _u(b, m + 16, 4) == 1023
                        dtype = np.float64
                elif mtype == 0x0008:
                    assert b[m] == 3 and b[m + 1] == 1
                    addr, size = _u(b, m + 2, 8), _u(b, m + 10, 8)
                q = m + msize
            assert q == end
            cnt = int(np.prod(dims))
            assert size == 8 * cnt and addr % 8 == 0
            out[nm] = (dtype, dims, np.frombuffer(b[addr:addr + size], dtype=dtype).reshape(dims))
        assert key_next == last_name                         # key i+1 = largest name in child i
    assert names_in_order == sorted(names_in_order)           # group entries are in strcmp order
    return out


def test_roundtrip_and_structure(tmp_path):
    p = str(tmp_path / "ck.h5")
    h = mlf_hdf5_file()
    assert h.createFile(p) == 0
    rng = np.random.default_rng(0)
    arrays = {"iPar": np.array([2 ** 63 - 1], dtype=np.int64), "rPar": np.array([1.5, np.finfo(float).max]),
              "iVar": np.array([50, 0], dtype=np.int64), "rVar": rng.random(3), "Mu": rng.random((3, 2)),
              "Cov": rng.random((3, 2, 2)), "lambda": rng.random(3)}
    for k, v in arrays.items():
        assert h.put(k, v) == 0
    assert h.put("Mu", np.zeros((3, 2))) == -6           # exists, no override (H5Dcreate_f would fail)
    assert h.put("Mu", arrays["Mu"], override=True) == 0
    h.finalize()
    parsed = parse_hdf5(p)
    assert set(parsed) == set(arrays)
    for k, v in arrays.items():
        assert parsed[k][0] == v.dtype and parsed[k][1] == list(v.shape)
        np.testing.assert_array_equal(parsed[k][2], v)
    g = mlf_hdf5_file()
    assert g.openFile(p, rw=False) == 0
    assert sorted(g.names()) == sorted(arrays)
    for k, v in arrays.items():
        np.testing.assert_array_equal(g.get(k), v)
    assert g.put("x", np.zeros(1)) == -6                 # read-only handle
    g.finalize()


def test_many_datasets_span_several_symbol_nodes(tmp_path):
    p = str(tmp_path / "many.h5")
    h = mlf_hdf5_file()
    assert h.createFile(p) == 0
    for i in range(37):
        assert h.put(f"d{i:03d}", np.full((i % 4 + 1, 2), float(i))) == 0
    h.finalize()
    parsed = parse_hdf5(p)
    assert len(parsed) == 37
    for i in range(37):
        np.testing.assert_array_equal(parsed[f"d{i:03d}"][2], np.full((i % 4 + 1, 2), float(i)))
    g = mlf_hdf5_file()
    assert g.openFile(p) == 0 and len(g.names()) == 37
    g.finalize()


def test_create_excl_and_open_missing(tmp_path):
    lib = _lib.load()
    p = str(tmp_path / "a.h5")
    h = lib.mlf_hdf5_createFile(p.encode(), 1)
    assert h
    assert lib.mlf_dealloc(h) == 0
    assert lib.mlf_hdf5_createFile(p.encode(), 0) is None   # H5F_ACC_EXCL on an existing file (src/mlf_hdf5.f90:178-185)
    assert lib.mlf_hdf5_openFile(str(tmp_path / "missing.h5").encode(), 1) is None
    open(str(tmp_path / "junk.h5"), "wb").write(b"not hdf5" * 20)
    assert lib.mlf_hdf5_openFile(str(tmp_path / "junk.h5").encode(), 0) is None
    h2 = lib.mlf_hdf5_createFile(p.encode(), 1)               # truncates
    assert h2 and lib.mlf_b200_h5_count(h2) == 0
    lib.mlf_dealloc(h2)
