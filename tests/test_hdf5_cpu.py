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


# ---- groups, stock-libhdf5 input, rewrite safety (round 2) -----------------------------------------------------------
def walk_hdf5(path):
    """Second independent reader (recursive): {path: ndarray} for every contiguous 8-byte dataset, groups, base address.
    Handles a user block in front of the superblock, header continuation blocks and layout versions 1-3."""
    b = open(path, "rb").read()
    sb = 0
    while b[sb:sb + 8] != b"\x89HDF\r\n\x1a\n":
        sb = 512 if sb == 0 else sb * 2
        assert sb < len(b)
    base = sb
    assert b[sb + 8] == 0
    out, groups, attrs = {}, [], {}

    def messages(oh):
        oh += base
        assert b[oh] == 1
        nmsg, hsize = _u(b, oh + 2, 2), _u(b, oh + 8, 4)
        chunks, msgs, ci = [(oh + 16, hsize)], [], 0
        while ci < len(chunks) and len(msgs) < nmsg:
            q, end = chunks[ci][0], chunks[ci][0] + chunks[ci][1]
            ci += 1
            while q + 8 <= end and len(msgs) < nmsg:
                t, sz = _u(b, q, 2), _u(b, q + 2, 2)
                if t == 0x10:
                    chunks.append((base + _u(b, q + 8, 8), _u(b, q + 16, 8)))
                msgs.append((t, q + 8, sz))
                q += 8 + sz
        return msgs

    def group(btree, heap, prefix):
        heap += base
        assert b[heap:heap + 4] == b"HEAP"
        hd = base + _u(b, heap + 24, 8)

        def node(n):
            n += base
            assert b[n:n + 4] == b"TREE"
            lvl, used, p = b[n + 5], _u(b, n + 6, 2), n + 24
            for _ in range(used):
                child = _u(b, p + 8, 8)
                p += 16
                if lvl > 0:
                    node(child)
                    continue
                c = child + base
                assert b[c:c + 4] == b"SNOD"
                for s in range(_u(b, c + 6, 2)):
                    e = c + 8 + 40 * s
                    noff, oh = _u(b, e, 8), _u(b, e + 8, 8)
                    obj(oh, prefix + b[hd + noff:b.index(b"\0", hd + noff)].decode())
        node(btree)

    def obj(oh, name):
        dims = dtype = addr = None
        na = 0
        for t, m, sz in messages(oh):
            if t == 0x11:
                groups.append(name)
                group(_u(b, m, 8), _u(b, m + 8, 8), name + "/")
                return
            if t == 1:
                p = m + (8 if b[m] == 1 else 4)
                dims = [_u(b, p + 8 * j, 8) for j in range(b[m + 1])]
            elif t == 3:
                dtype = {0: np.int64, 1: np.float64}[b[m] & 0xF]
                assert _u(b, m + 4, 4) == 8
            elif t == 8:
                addr = _u(b, m + 2, 8) if b[m] == 3 else _u(b, m + 8, 8)
            elif t == 0xC:
                na += 1
        n = int(np.prod(dims))
        out[name] = np.frombuffer(b[base + addr:base + addr + 8 * n], dtype=dtype).reshape(dims)
        attrs[name] = na

    for t, m, sz in messages(_u(b, sb + 64, 8)):
        if t == 0x11:
            group(_u(b, m, 8), _u(b, m + 8, 8), "")
    return out, groups, base, attrs


def test_groups_nest_and_roundtrip(tmp_path):
    """Sub-object checkpoints live in groups named after the sub-object (src/mlf_hdf5.f90:106-151,657-669)."""
    p = str(tmp_path / "g.h5")
    h = mlf_hdf5_file()
    assert h.createFile(p) == 0
    rng = np.random.default_rng(1)
    a, bb, c = rng.random((2, 3)), rng.integers(0, 9, 4).astype(np.int64), rng.random(5)
    assert h.put("top", a) == 0
    assert h.open_group("km") is None                      # H5Gopen_f on a missing group
    g = h.open_group("km", create=True)
    assert g is not None
    assert h.open_group("km", create=True) is None         # H5Gcreate_f on an existing group
    assert g.put("Mu", bb) == 0
    sub = g.getSubObject("inner")                          # open, else create
    assert sub is not None and sub.put("v", c) == 0
    again = g.getSubObject("inner")
    np.testing.assert_array_equal(again.get("v"), c)
    assert h.open_group("top") is None                     # a dataset, not a group
    assert h.put("km", a) != 0                             # a group of that name exists
    e = h.open_group("empty", create=True)
    for x in (again, sub, g, e):
        x.finalize()
    h.finalize()
    out, groups, base, _ = walk_hdf5(p)
    assert base == 0 and sorted(groups) == ["empty", "km", "km/inner"]
    assert sorted(out) == ["km/Mu", "km/inner/v", "top"]
    np.testing.assert_array_equal(out["top"], a)
    np.testing.assert_array_equal(out["km/Mu"], bb)
    np.testing.assert_array_equal(out["km/inner/v"], c)
    r = mlf_hdf5_file()
    assert r.openFile(p, rw=True) == 0 and r.unsupported() == 0
    assert sorted(r.names()) == ["km/Mu", "km/inner/v", "top"]
    np.testing.assert_array_equal(r.get("km/inner/v"), c)
    k = r.open_group("km")
    assert sorted(k.names()) == ["Mu", "inner/v"]
    np.testing.assert_array_equal(k.get("Mu"), bb)
    assert r.open_group("empty") is not None               # empty groups survive the rewrite cycle
    k.finalize()
    r.finalize()


STOCK = None
try:
    import scipy.io.matlab as _m
    import os as _os
    _cand = _os.path.join(_os.path.dirname(_m.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    STOCK = _cand if _os.path.exists(_cand) else None
except Exception:  # noqa: BLE001
    pass


def test_reads_a_file_written_by_stock_libhdf5(tmp_path):
    """The only libhdf5-written file in this image: SciPy ships MATLAB's v7.3 test file (HDF5 1.6-era writer: 512-byte
    user block, superblock v0, symbol-table root group, version-1 layout message, one attribute).  The built-in reader
    must return its dataset, report the content it cannot represent, and refuse a read-write open (a rewrite would
    drop the user block and the attribute)."""
    import pytest
    import shutil

    if STOCK is None:
        pytest.skip("scipy's MATLAB v7.3 test file is not installed")
    p = str(tmp_path / "stock.h5")
    shutil.copy(STOCK, p)
    before = open(p, "rb").read()
    want, groups, base, attrs = walk_hdf5(p)
    assert base == 512 and groups == [] and list(want) == ["testdouble"] and attrs["testdouble"] == 1
    np.testing.assert_allclose(want["testdouble"].ravel(), np.arange(9) * np.pi / 4, rtol=1e-15)
    r = mlf_hdf5_file()
    assert r.openFile(p, rw=False) == 0
    assert r.names() == ["testdouble"]
    got = r.get("testdouble")
    assert got.shape == (9, 1) and got.dtype == np.float64
    np.testing.assert_array_equal(got, want["testdouble"])
    assert r.unsupported() == 2                            # user block + the dataset's attribute
    assert r.put("x", np.zeros(1)) == -6                   # read-only
    r.finalize()
    assert mlf_hdf5_file().openFile(p, rw=True) != 0       # refused: nothing may be lost by a rewrite
    assert open(p, "rb").read() == before                  # and nothing was touched


def test_open_close_does_not_rewrite_and_writes_are_atomic(tmp_path):
    import os

    p = str(tmp_path / "ck.h5")
    h = mlf_hdf5_file()
    assert h.createFile(p) == 0 and h.put("a", np.arange(3.0)) == 0
    h.finalize()
    before = open(p, "rb").read()
    os.chmod(p, 0o444)
    os.chmod(str(tmp_path), 0o555)                          # any rewrite (temporary + rename) would fail loudly here
    try:
        r = mlf_hdf5_file()
        assert r.openFile(p, rw=True) == 0
        np.testing.assert_array_equal(r.get("a"), np.arange(3.0))
        r.finalize()                                        # no put since open: not dirty, nothing written
    finally:
        os.chmod(str(tmp_path), 0o755)
        os.chmod(p, 0o644)
    assert open(p, "rb").read() == before
    assert [f for f in os.listdir(tmp_path) if ".tmp." in f] == []
    r = mlf_hdf5_file()
    assert r.openFile(p, rw=True) == 0 and r.put("b", np.ones(2)) == 0
    r.finalize()
    assert [f for f in os.listdir(tmp_path) if ".tmp." in f] == []     # temporary renamed over the target
    out, _, _, _ = walk_hdf5(p)
    assert sorted(out) == ["a", "b"]


def test_malformed_files_are_rejected_not_trusted(tmp_path):
    """Every truncation and a sweep of single-byte corruptions of a valid checkpoint: the reader either rejects the file
    or returns bounded data; it never reads out of bounds (the process survives) and never reports absurd sizes."""
    lib = _lib.load()
    p = str(tmp_path / "ok.h5")
    h = mlf_hdf5_file()
    assert h.createFile(p) == 0
    assert h.put("Mu", np.arange(6.0).reshape(3, 2)) == 0 and h.put("iPar", np.array([7], dtype=np.int64)) == 0
    g = h.open_group("sub", create=True)
    assert g.put("v", np.arange(4.0)) == 0
    g.finalize()
    h.finalize()
    good = open(p, "rb").read()
    q = str(tmp_path / "bad.h5")
    for cut in list(range(0, len(good), 37)) + [len(good) - 1]:
        open(q, "wb").write(good[:cut])
        hh = lib.mlf_hdf5_openFile(q.encode(), 0)
        if hh:
            assert lib.mlf_b200_h5_count(hh) <= 3
            lib.mlf_dealloc(hh)
    rng = np.random.default_rng(5)
    for _ in range(400):
        bad = bytearray(good)
        for pos in rng.integers(8, len(good) - 56, size=int(rng.integers(1, 4))):   # metadata and data alike
            bad[int(pos)] = int(rng.integers(0, 256))
        open(q, "wb").write(bytes(bad))
        hh = lib.mlf_hdf5_openFile(q.encode(), 0)
        if hh:
            n = lib.mlf_b200_h5_count(hh)
            assert 0 <= n <= 3
            lib.mlf_dealloc(hh)
    # hostile dims: 2^61 elements would wrap count*8
    bad = bytearray(good)
    i = good.index(struct.pack("<QQ", 3, 2))              # Mu's dataspace dims
    bad[i:i + 8] = struct.pack("<Q", 1 << 61)
    open(q, "wb").write(bytes(bad))
    assert lib.mlf_hdf5_openFile(q.encode(), 0) is None
