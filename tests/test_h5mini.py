"""The minimal HDF5 reader / writer behind the save-file support (opendxmc_b200/csrc/h5mini.*).

No HDF5 library exists in this image, so
  * the READER is pinned on a genuine libhdf5-written file that ships with scipy (a MATLAB 7.3 file: 512-byte user block,
    version-0 superblock, old-style root group, a 9 x 1 array of doubles whose values scipy's own test-suite documents);
  * the WRITER is pinned on the reader: every object kind OpenDXMC's save files hold (R:src/libopendxmc/hdf5wrapper.cpp:
    384-628) is written, re-read and compared, and the file bytes are checked against the format specification where
    that is cheap (signature, superblock fields, structure signatures, the deflate stream of a chunk)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from opendxmc_b200 import _capi as K

F64, U64, U8, STRING = 1, 2, 3, 4
NP = {F64: np.float64, U64: np.uint64, U8: np.uint8}


def _open(path):
    h = K.VP()
    assert K.load().dxb_h5_open(C.byref(h), str(path).encode()) == K.DXB_OK
    return h


def _dataset(h, path):
    lib = K.load()
    t, r, d, z = C.c_int(), C.c_int(), (C.c_uint64 * 8)(), C.c_int()
    assert lib.dxb_h5_dataset_info(h, path.encode(), C.byref(t), C.byref(r), d, C.byref(z)) == K.DXB_OK, path
    dims = [int(v) for v in list(d)[:r.value]]
    n = int(np.prod(dims)) if dims else 1
    if t.value == STRING:
        return [lib.dxb_h5_dataset_string(h, path.encode(), i).decode() for i in range(n)], dims, bool(z.value)
    a = np.zeros(n, dtype=NP[t.value])
    assert lib.dxb_h5_dataset_read(h, path.encode(), a.ctypes.data_as(K.VP), a.nbytes) == K.DXB_OK
    return a.reshape(dims) if dims else a[0], dims, bool(z.value)


def _attr(h, group, name):
    lib = K.load()
    t, n = C.c_int(), C.c_int64()
    assert lib.dxb_h5_attribute_info(h, group.encode(), name.encode(), C.byref(t), C.byref(n)) == K.DXB_OK, (group, name)
    a = np.zeros(max(1, n.value), dtype=NP[t.value])
    assert lib.dxb_h5_attribute_read(h, group.encode(), name.encode(), a.ctypes.data_as(K.VP), a.nbytes) == K.DXB_OK
    return (a[0] if n.value < 0 else a), n.value


def _listing(h, group):
    out = {"g": [], "d": [], "a": []}
    for line in K.load().dxb_h5_list(h, group.encode()).decode().splitlines():
        out[line[0]].append(line[2:])
    return out


def test_reader_on_a_genuine_libhdf5_file():
    """scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat was written by MATLAB 7.4 through libhdf5 (1.6 era): user block,
    superblock version 0, symbol-table root group, version-1 object header, version-2 contiguous data layout.  scipy's
    test-suite states its content: testdouble = pi / 4 * arange(9)."""
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's test data is not installed")
    assert open(path, "rb").read(6) == b"MATLAB"          # the user block
    h = _open(path)
    assert _listing(h, "/")["d"] == ["testdouble"]
    a, dims, z = _dataset(h, "/testdouble")
    assert dims == [9, 1] and not z
    assert np.array_equal(a.reshape(-1), np.pi / 4 * np.arange(9.0))
    K.load().dxb_h5_close(h)


def _put(h, path, a, deflate):
    a = np.ascontiguousarray(a)
    code = {np.dtype(np.float64): F64, np.dtype(np.uint64): U64, np.dtype(np.uint8): U8}[a.dtype]
    dims = (C.c_uint64 * max(1, a.ndim))(*a.shape)
    assert K.load().dxb_h5_put_dataset(h, path.encode(), code, a.ndim, dims, a.ctypes.data_as(K.VP), 1 if deflate else 0) == K.DXB_OK


def test_write_then_read_every_object_kind(tmp_path):
    lib = K.load()
    rng = np.random.default_rng(1)
    h = lib.dxb_h5_create()
    dose = rng.random((7, 5, 3))                                  # z-y-x order: dims reversed (R:...hdf5wrapper.cpp:121-124)
    count = rng.integers(0, 2 ** 40, (7, 5, 3)).astype(np.uint64)
    mat = rng.integers(0, 5, (7, 5, 3)).astype(np.uint8)
    big = np.repeat(rng.random(1000), 50)                         # compressible, > 128 elements
    _put(h, "/dosearray", dose, True)
    _put(h, "/doseeventcountarray", count, True)
    _put(h, "/materialarray", mat, True)
    _put(h, "/dimensions", np.array([3, 5, 7], dtype=np.uint64), False)
    _put(h, "/spacing", np.array([0.1, 0.2, 0.3]), False)
    _put(h, "/deep/er/group/big", big, True)
    _put(h, "/empty", np.zeros(0), False)
    names = ["Air", "Soft tissue, a longer name with, commas", "", "Bone"]
    arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
    assert lib.dxb_h5_put_strings(h, b"/materialnames", len(names), arr) == K.DXB_OK
    for i in range(1, 4):
        g = f"/beams/CTSpiralBeams/{i}".encode()
        assert lib.dxb_h5_make_group(h, g) == K.DXB_OK
        v = np.array([1.5 * i, -2.0, 3.25])
        assert lib.dxb_h5_put_attribute(h, g, b"start_position", F64, 3, v.ctypes.data_as(K.VP)) == K.DXB_OK
        s = np.array([0.125 * i])
        assert lib.dxb_h5_put_attribute(h, g, b"pitch", F64, -1, s.ctypes.data_as(K.VP)) == K.DXB_OK
        n = np.array([10 ** 6 + i], dtype=np.uint64)
        assert lib.dxb_h5_put_attribute(h, g, b"particles_per_exposure", U64, -1, n.ctypes.data_as(K.VP)) == K.DXB_OK
    many = f"/many".encode()
    for i in range(150):                                          # more links than one symbol node holds
        _put(h, f"/many/d{i:03d}", np.array([float(i)]), False)
    path = tmp_path / "scene.h5"
    assert lib.dxb_h5_save(h, str(path).encode()) == K.DXB_OK, lib.dxb_h5_error(h)
    lib.dxb_h5_close(h)

    raw = open(path, "rb").read()
    # format checks that need no library: signature, superblock version 0, 8-byte offsets / lengths, end-of-file address
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert int.from_bytes(raw[40:48], "little") == len(raw)
    for tag in (b"TREE", b"HEAP", b"SNOD", b"GCOL"):
        assert tag in raw
    # the chunk of /deep/er/group/big is a plain zlib stream of the row-major doubles
    z = zlib.compress(big.tobytes(), 6)
    assert z in raw and len(z) < big.nbytes // 4

    r = _open(path)
    top = _listing(r, "/")
    assert sorted(top["g"]) == ["beams", "deep", "many"]
    assert sorted(top["d"]) == sorted(["dosearray", "doseeventcountarray", "materialarray", "dimensions", "spacing", "empty", "materialnames"])
    for p, a, deflated in (("/dosearray", dose, True), ("/doseeventcountarray", count, True), ("/materialarray", mat, False),
                           ("/deep/er/group/big", big, True)):
        got, dims, z = _dataset(r, p)
        assert dims == list(a.shape) and np.array_equal(got, a), p
        assert z == (a.size > 0 and (deflated or a.dtype == np.uint8)), p
    assert np.array_equal(_dataset(r, "/dimensions")[0], [3, 5, 7])
    assert np.array_equal(_dataset(r, "/spacing")[0], [0.1, 0.2, 0.3])
    assert _dataset(r, "/empty")[1] == [0]
    assert _dataset(r, "/materialnames")[0] == names
    assert _listing(r, "/beams/CTSpiralBeams")["g"] == ["1", "2", "3"]
    for i in range(1, 4):
        g = f"/beams/CTSpiralBeams/{i}"
        assert _listing(r, g)["a"] == ["start_position", "pitch", "particles_per_exposure"]   # creation order
        v, n = _attr(r, g, "start_position")
        assert n == 3 and np.array_equal(v, [1.5 * i, -2.0, 3.25])
        assert _attr(r, g, "pitch") == (0.125 * i, -1)
        assert _attr(r, g, "particles_per_exposure") == (10 ** 6 + i, -1)
    assert _listing(r, "/many")["d"] == [f"d{i:03d}" for i in range(150)]
    assert _dataset(r, "/many/d149")[0][0] == 149.0
    assert lib.dxb_h5_exists(r, b"/beams/CTSpiralBeams/2") == 1 and lib.dxb_h5_exists(r, b"/spacing") == 2 and lib.dxb_h5_exists(r, b"/nope") == 0
    lib.dxb_h5_close(r)


def test_not_an_hdf5_file(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5" * 100)
    h = K.VP()
    assert K.load().dxb_h5_open(C.byref(h), str(p).encode()) == K.DXB_EINVAL
    assert K.load().dxb_h5_open(C.byref(h), str(tmp_path / "missing.h5").encode()) == K.DXB_EINVAL
