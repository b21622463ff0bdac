"""On-disk formats: the HDF5 trajectory export of examples/freeflyerSE2.ipynb cell 6 and the run-time reader of
environment/iss_corner.mat.  No HDF5 library exists in this image, so the writer is checked structurally against the file-format
specification (signature, superblock fields, object addresses, end-of-file address) and through the independent reader."""
import json
import os
import struct

import numpy as np
import pytest

from util import gb, ROOT

REF_MAT = "/root/reference/src/environment/iss_corner.mat"


def test_trajectory_h5_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    N = 200
    X = rng.normal(size=(N, 6)); U = rng.normal(size=(N, 3))
    path = str(tmp_path / "predefined_trajectory_example.h5")
    ind_x = {"x": 0, "y": 1, "theta": 2, "vx": 3, "vy": 4, "omega": 5}
    ind_u = {"Fx": 0, "Fy": 1, "M": 2}
    n = gb.trajio.write_trajectory_h5(path, X, U, 110.0, ind_x, ind_u)
    assert os.path.getsize(path) == n
    d = gb.trajio.read_h5(path)
    assert sorted(d) == ["ind_u", "ind_x", "traj"] and sorted(d["traj"]) == ["t_traj", "u_traj", "x_traj"]
    assert np.array_equal(d["traj"]["x_traj"], X) and np.array_equal(d["traj"]["u_traj"], U)
    assert d["traj"]["x_traj"].shape == (N, 6)            # = the (6, N) Julia array of the notebook in HDF5.jl's dimension order
    t = d["traj"]["t_traj"]
    assert t.shape == (N,) and t[0] == 0.0 and abs(t[-1] - 110.0) < 1e-12 and np.allclose(np.diff(t), 110.0 / (N - 1))
    assert {k: int(v) for k, v in d["ind_x"].items()} == ind_x and {k: int(v) for k, v in d["ind_u"].items()} == ind_u
    assert d["ind_x"]["theta"].shape == () and d["ind_x"]["theta"].dtype == np.int64


def test_h5_file_structure_follows_the_format_specification(tmp_path):
    path = str(tmp_path / "s.h5")
    gb.trajio.write_h5(path, {"traj": {"x_traj": np.arange(6.0).reshape(3, 2)}})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    sb_ver, fs_ver, root_ver, _, shm_ver, so, sl, _ = struct.unpack_from("<8B", b, 8)
    assert (sb_ver, fs_ver, root_ver, shm_ver, so, sl) == (0, 0, 0, 0, 8, 8)
    base, freesp, eof, drv = struct.unpack_from("<4Q", b, 24)
    assert base == 0 and eof == len(b) and freesp == drv == 0xFFFFFFFFFFFFFFFF
    name_off, hdr, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", b, 56)
    assert cache == 1 and b[bt:bt + 4] == b"TREE" and b[hp:hp + 4] == b"HEAP" and hdr % 8 == 0
    ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", b, hdr)
    assert ver == 1 and nmsg == 1 and refc == 1 and size == 24
    mtype, msize = struct.unpack_from("<HH", b, hdr + 16)
    assert mtype == 0x0011 and msize == 16 and struct.unpack_from("<QQ", b, hdr + 24) == (bt, hp)
    snod = struct.unpack_from("<Q", b, bt + 32)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    # the dataset: IEEE little-endian double, contiguous, data exactly where the layout message says
    d = gb.trajio.read_h5(path)
    assert np.array_equal(d["traj"]["x_traj"], np.arange(6.0).reshape(3, 2))
    raw = np.arange(6.0).tobytes()
    assert b.count(raw) == 1 and b.index(raw) % 8 == 0


def test_reader_rejects_foreign_bytes(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all, sorry..........")
    with pytest.raises(ValueError):
        gb.trajio.read_h5(str(p))


@pytest.mark.skipif(not os.path.exists(REF_MAT), reason="the reference tree is only present in the build container")
def test_runtime_mat_reader_equals_the_packaged_geometry():
    """iss_corner.jl:11-23: reading the reference's own .mat at run time gives exactly the packaged JSON table (26 keep-out,
    16 keep-in zones, 4 boxes, 2 spheres), and the environment built from it is identical."""
    d = gb.trajio.load_iss_corner_mat(REF_MAT)
    j = json.load(open(os.path.join(ROOT, "gusto.jl_b200", "data", "iss_corner.json")))
    for k in ("keepin_zones", "keepout_zones", "obstacle_rectangles", "obstacle_spheres"):
        assert d[k] == j[k], k
    assert (len(d["keepin_zones"]), len(d["keepout_zones"]), len(d["obstacle_rectangles"]), len(d["obstacle_spheres"])) == (16, 26, 4, 2)
    e1, e2 = gb.models.ISSCorner(add_obstacles=True), gb.models.ISSCorner(add_obstacles=True, mat_path=REF_MAT)
    for a, b in zip(e1.keepout_zones + e1.keepin_zones, e2.keepout_zones + e2.keepin_zones):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert len(e1.obstacle_set) == len(e2.obstacle_set) == 6
