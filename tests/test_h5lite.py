"""exp_b200.h5lite: the HDF5 subset of the reference's table caches (SLGridSph::ReadH5Cache /
WriteH5Cache, exputil/SLGridMP2.cc:490-696; EmpCylSL::ReadH5Cache / WriteH5Cache,
exputil/EmpCylSL.cc:7378-7680) without an HDF5 library: the reader against a file written by a real
libhdf5, writer -> reader round trips of both cache layouts, and the structure of the written file."""
import os
import struct

import numpy as np
import pytest

import halo_fixture as H
from exp_b200 import cyltable, h5lite, slgrid


def test_reads_a_real_libhdf5_file():
    """a MATLAB v7.3 file from scipy's test data: 512-byte user block, superblock 0, old-style group,
    fixed-length string attribute, contiguous float64 dataset, all written by libhdf5 itself"""
    scipy = pytest.importorskip("scipy")
    path = os.path.join(os.path.dirname(scipy.__file__), "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    f = h5lite.File(path)
    assert f.keys() == ["testdouble"] and f.sb == 512
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.attrs["MATLAB_class"] == "double"
    assert np.allclose(d.read()[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_slgrid_cache_round_trip(tmp_path):
    fx = H.load()
    tab = fx["table"]
    p = str(tmp_path / "SLGridSph.cache.run0")
    h5lite.write_slgrid_cache(p, tab, model="SLGridSph.model")
    c = h5lite.read_slgrid_cache(p)
    assert (c["geometry"], c["forceID"], c["model"]) == ("sphere", "SLGridSph", "SLGridSph.model")
    assert (c["lmax"], c["nmax"], c["numr"], c["cmap"], c["diverge"]) == (2, 10, 4000, 1, 0)
    assert (c["rmin"], c["rmax"], c["rmapping"], c["dfac"]) == (1e-4, 1.95, 0.0667, 1.0)
    assert c["ev"].tobytes() == np.ascontiguousarray(tab.ev).tobytes()
    assert c["ef"].tobytes() == np.ascontiguousarray(tab.ef).tobytes()
    model = slgrid.TableModel.from_arrays(fx["r"], fx["rho"], fx["m"], fx["u"])
    t2 = slgrid.table_from_cache(p, model)
    assert t2.p0.tobytes() == np.ascontiguousarray(tab.p0).tobytes()
    assert t2.ef.tobytes() == np.ascontiguousarray(tab.ef).tobytes() and (t2.rmin, t2.rmax, t2.rmap) == (tab.rmin, tab.rmax, tab.rmap)


def test_written_file_structure(tmp_path):
    """superblock 0 with 8-byte offsets/lengths, end-of-file address = file size, root symbol-table
    entry caches B-tree and heap, every group is HEAP + SNOD + TREE, strings live in a global heap"""
    tab = H.load()["table"]
    p = str(tmp_path / "c.h5")
    h5lite.write_slgrid_cache(p, tab)
    b = open(p, "rb").read()
    assert b[:8] == h5lite.SIG and b[8] == 0 and b[13] == 8 and b[14] == 8
    base, _, eof, _ = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and eof == len(b)
    name_off, ohdr, cache_type, _, bt, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert cache_type == 1 and b[bt:bt + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP" and b[ohdr] == 1
    assert b.count(b"SNOD") == 2 + 3 and b.count(b"GCOL") == 1        # root, Harmonic, three l groups
    f = h5lite.File(p)
    assert f.keys() == ["Harmonic"] and f["Harmonic"].keys() == ["0", "1", "2"]
    assert f["Harmonic/1"].keys() == ["ef", "ev"] and f["Harmonic/1/ef"].shape == (10, 4000)


def test_cylinder_cache_round_trip(tmp_path):
    tab = cyltable.synthetic_cyl_table(mmax=3, norder=12, numx=16, numy=8)     # 12 > 8: names sort as strings
    p = str(tmp_path / "cyl.cache")
    h5lite.write_cylinder_cache(p, tab, cmass=0.05)
    t2 = cyltable.table_from_cache(p)
    for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS"):
        assert getattr(t2, k).tobytes() == np.ascontiguousarray(getattr(tab, k)).tobytes(), k
    assert (t2.mmax, t2.norder, t2.numx, t2.numy, t2.cmapr, t2.cmapz) == (3, 12, 16, 8, 1, 1)
    assert (t2.ascale, t2.hscale, t2.rmin, t2.rmax) == (tab.ascale, tab.hscale, tab.rmin, tab.rmax)
    f = h5lite.File(p)
    assert f.attrs["cmass"] == 0.05 and f.attrs["model"] == "Exponential"
    assert sorted(f["Cosine"].keys(), key=int) == ["0", "1", "2", "3"] and "0" not in f["Sine"]
    assert f["Cosine/2"].keys() == sorted(str(n) for n in range(12))


def test_chunked_deflate_dataset_is_read():
    """a hand-assembled chunked + shuffle + deflate dataset (what h5py/HighFive write when compression
    is requested): exercises the chunk B-tree and the filter pipeline of the reader"""
    import zlib
    a = np.arange(30, dtype="<f8").reshape(5, 6) * 1.5
    w = h5lite.Writer()
    w.dataset(w.root, "x", a)                      # contiguous twin for the header layout
    # build the chunked variant by patching a written file: easier to assemble by hand
    O = 8
    out = bytearray(b"\0" * 96)
    def alloc(b):
        while len(out) % 8: out.append(0)
        p = len(out); out.extend(b); return p
    cshape = (4, 4)
    chunks = []
    for i in range(0, 5, 4):
        for j in range(0, 6, 4):
            c = np.zeros(cshape, dtype="<f8")
            blk = a[i:i + 4, j:j + 4]
            c[:blk.shape[0], :blk.shape[1]] = blk
            raw = np.frombuffer(c.tobytes(), dtype=np.uint8).reshape(-1, 8).T.tobytes()    # shuffle
            z = zlib.compress(raw)
            chunks.append(((i, j), alloc(z), len(z)))
    node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(chunks), h5lite.UNDEF, h5lite.UNDEF))
    for (i, j), addr, n in chunks:
        node += struct.pack("<IIQQQ", n, 0, i, j, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 5, 6, 0)
    bt = alloc(bytes(node))
    space = struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", 5, 6)
    dtype = h5lite.Writer._dtype_msg(np.float64)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 4, 4, 8)
    filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 8, 0) \
        + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 6, 0)
    body = bytearray()
    for t, pl in ((1, space), (3, dtype), (8, layout), (0x0B, filt)):
        pl = pl + b"\0" * (-len(pl) % 8)
        body += struct.pack("<HHB3x", t, len(pl), 0) + pl
    dhdr = alloc(struct.pack("<BBHII4x", 1, 0, 4, 1, len(body)) + bytes(body))
    heap = bytearray(b"\0" * 8 + b"x\0" + b"\0" * 6 + struct.pack("<QQ", 1, 32) + b"\0" * 16)
    hd = alloc(bytes(heap))
    hp = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 16, hd))
    sn = alloc(b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII16x", 8, dhdr, 0, 0) + b"\0" * (31 * 40))
    gt = alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, h5lite.UNDEF, h5lite.UNDEF) + struct.pack("<QQQ", 0, sn, 8) + b"\0" * 512)
    gbody = struct.pack("<HHB3x", 0x11, 16, 0) + struct.pack("<QQ", gt, hp)
    root = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(gbody)) + gbody)
    sb = h5lite.SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 16, 16, 0) \
        + struct.pack("<QQQQ", 0, h5lite.UNDEF, len(out), h5lite.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", gt, hp)
    out[0:len(sb)] = sb
    f = h5lite.File(bytes(out))
    assert np.array_equal(f["x"].read(), a)


# ---- coefficient files (CoefClasses::SphCoefs / CylCoefs layout) --------------------------------
def test_coefficient_file_create_extend_read(tmp_path):
    """dump_coefs_h5: first call = Coefs::WriteH5Coefs, later calls = ExtendH5Coefs (reference
    expui/Coefficients.cc:3100-3205); the snapshot group grows past one symbol node and past one
    B-tree node (32 and 1024 entries)"""
    rng = np.random.default_rng(5)
    p = str(tmp_path / "outcoef.halo.h5")
    lmax, nmax = 4, 7
    rows = (lmax + 1) * (lmax + 2) // 2
    want = []
    for k in range(40):
        c = rng.normal(size=(rows, nmax)) + 1j * rng.normal(size=(rows, nmax))
        want.append(c)
        n = h5lite.dump_coefs_h5(p, "sphere", "dark halo", "SphereSL", 0.01 * k, c, lmax=lmax, scale=0.05,
                                 config="id: SphereSL\nparameters: {Lmax: 4}\n", center=(0.1 * k, 0.0, -1.0))
        assert n == k + 1
    f = h5lite.File(p)
    assert f.attrs["CoefficientOutputVersion"] == "1.0" and f.attrs["geometry"] == "sphere"
    assert f.attrs["name"] == "dark halo" and f.attrs["forceID"] == "SphereSL"
    assert f.attrs["lmax"] == lmax and f.attrs["nmax"] == nmax and f.attrs["scale"] == 0.05
    assert int(f["count"].read()) == 40 and f["count"].shape == ()
    u = f["Units"].read()
    assert u.dtype.names == ("name", "unit", "value") and u.dtype.itemsize == 36
    assert [x.decode() for x in u["name"]] == ["length", "mass", "time"]
    d = f["snapshots/00000039/coefficients"]
    assert d.shape == (rows, nmax) and d.dtype.names == ("r", "i") and d.dtype.itemsize == 16
    r = h5lite.read_coefs_h5(p)
    assert np.array_equal(r["coefs"], np.array(want))
    assert np.allclose(r["times"], 0.01 * np.arange(40)) and np.array_equal(r["centers"][3], [0.1 * 3, 0.0, -1.0])
    r2 = h5lite.read_coefs_h5(p, stride=4, tmin=0.1)
    assert np.array_equal(r2["coefs"], np.array(want)[12::4])
    # CheckH5Params: a different basis does not extend the file
    with pytest.raises(h5lite.H5Error):
        h5lite.dump_coefs_h5(p, "sphere", "dark halo", "SphereSL", 1.0, want[0], lmax=lmax, scale=0.06)
    with pytest.raises(h5lite.H5Error):
        h5lite.dump_coefs_h5(p, "sphere", "dark halo", "SphereSL", 1.0, want[0][:, :5].copy(), lmax=lmax, scale=0.05)


def test_coefficient_file_many_snapshots_and_cylinder(tmp_path):
    p = str(tmp_path / "outcoef.disk.h5")
    w = h5lite.Writer()
    a = w.root["attrs"]
    a["CoefficientOutputVersion"], a["geometry"], a["name"], a["config"] = "1.0", "cylinder", "disk", ""
    a["mmax"], a["nmax"], a["forceID"] = 3, 5, "Cylinder"
    snaps = w.group(w.root, "snapshots")
    rng = np.random.default_rng(6)
    want = rng.normal(size=(1100, 4, 5)) + 1j * rng.normal(size=(1100, 4, 5))
    cnt = 0
    for c in want:
        cnt = h5lite._add_snapshot(w, snaps, cnt, 0.5 * cnt, np.zeros(3), np.eye(3), c)
    w.dataset(w.root, "count", np.uint32(cnt))
    w.write(p)
    r = h5lite.read_coefs_h5(p)
    assert r["mmax"] == 3 and r["coefs"].shape == (1100, 4, 5) and np.array_equal(r["coefs"], want)
    assert h5lite.dump_coefs_h5(p, "cylinder", "disk", "Cylinder", 999.0, want[0], mmax=3) == 1101
    assert h5lite.read_coefs_h5(p, tmin=900.0)["times"].tolist() == [999.0]


def test_coefficient_file_legacy_order(tmp_path):
    """files without CoefficientOutputVersion hold the transposed legacy order (reference
    expui/Coefficients.cc:248-251, 300-310)"""
    p = str(tmp_path / "legacy.h5")
    c = np.arange(12.0).reshape(3, 4) + 0j
    w = h5lite.Writer()
    a = w.root["attrs"]
    a["geometry"], a["name"], a["config"], a["mmax"], a["nmax"], a["forceID"] = "cylinder", "d", "", 2, 4, "Cylinder"
    s = w.group(w.group(w.root, "snapshots"), "00000000")
    s["attrs"]["Time"] = 0.0
    w.dataset(s, "coefficients", c)
    w.dataset(w.root, "count", np.uint32(1))
    w.write(p)
    got = h5lite.read_coefs_h5(p)["coefs"][0]
    assert np.array_equal(got, c.reshape(-1).reshape((3, 4), order="F"))
