"""On-disk formats either side of the hot path.

* `write_trajectory_h5` / `read_h5`: the HDF5 trajectory export of /root/reference/examples/freeflyerSE2.ipynb cell 6
  (`h5open(...)`: group `traj` with `x_traj`, `u_traj`, `t_traj`, and the index groups `ind_x`, `ind_u` of scalar Int64
  datasets).  Neither libhdf5 nor h5py exists in this image, so the writer emits the file format directly -- the subset
  HDF5.jl / h5py read back without options: superblock version 0, version-1 object headers, one symbol-table group per
  level (B-tree + local heap + symbol-table node), contiguous little-endian Float64 / Int64 datasets ("HDF5 File Format
  Specification Version 1.1").  HDF5.jl stores a Julia (n, N) column-major array as a dataset of shape (N, n): that is exactly
  the memory of our row-major X[N][n], so arrays are written as they are and read back by HDF5.jl with the reference's shape.
  `read_h5` is a reader for the same subset (used by the tests and for re-loading predefined trajectories).
* `load_iss_corner_mat`: the keep-in / keep-out boxes of `src/environment/iss_corner.mat`, read at run time with
  scipy.io.loadmat exactly as `ISSCorner{T}()` does with `matread` (/root/reference/src/environment/iss_corner.jl:11-23).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 16, 16                       # symbol-table nodes hold up to 2 * LEAF_K entries


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def _datatype(arr):
    if arr.dtype == np.float64:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if arr.dtype == np.int64:
        return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)
    raise TypeError(f"unsupported dtype {arr.dtype} (Float64 / Int64 only)")


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                  # superblock, patched at the end

    def alloc(self, data):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, value):
        arr = np.asarray(value)
        if arr.dtype.kind in "iub":
            arr = arr.astype(np.int64)
        elif arr.dtype.kind == "f":
            arr = arr.astype(np.float64)
        raw = arr.tobytes(order="C")                               # (ascontiguousarray would turn a scalar into shape (1,))
        daddr = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill = struct.pack("<BBBB", 2, 1, 0, 0)                                    # version 2, allocate early, fill at allocation, undefined
        layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))                       # version 3, contiguous
        return self.alloc(_object_header([_msg(0x0001, space, 1), _msg(0x0003, _datatype(arr), 1), _msg(0x0005, fill, 1), _msg(0x0008, layout)]))

    def group(self, members):
        """members: dict name -> dict (sub-group) | array-like (dataset).  Returns (header, btree, heap) addresses."""
        if len(members) > 2 * LEAF_K:
            raise ValueError(f"at most {2 * LEAF_K} members per group in this writer")
        names = sorted(members, key=lambda s: s.encode())
        entries = []
        for n in names:
            v = members[n]
            entries.append((n, self.group(v)) if isinstance(v, dict) else (n, (self.dataset(v), None, None)))
        heap_data = bytearray(8)                                                   # offset 0: the empty name
        offs = []
        for n in names:
            offs.append(len(heap_data))
            heap_data += _pad8(n.encode() + b"\0")
        daddr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, daddr))      # no free block (H5HL_FREE_NULL = 1)
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for (n, (hdr, bt, hp)), off in zip(entries, offs):
            if bt is None:
                snod += struct.pack("<QQII16x", off, hdr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", off, hdr, 1, 0, bt, hp)
        snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
        saddr = self.alloc(bytes(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        tree += struct.pack("<QQQ", 0, saddr, offs[-1] if offs else 0)             # key 0 (empty name), child 0, key 1 (largest name)
        tree += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree))
        btree = self.alloc(bytes(tree))
        hdr = self.alloc(_object_header([_msg(0x0011, struct.pack("<QQ", btree, heap))]))
        return hdr, btree, heap

    def finish(self, root):
        hdr, bt, hp = root
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, tree):
    """Write a nested dict (groups) of array-likes (datasets) as an HDF5 file."""
    w = _Writer()
    data = w.finish(w.group(tree))
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def write_trajectory_h5(path, X, U, tf, ind_x=None, ind_u=None):
    """examples/freeflyerSE2.ipynb cell 6: traj/x_traj, traj/u_traj, traj/t_traj = collect(0:dt:Tf) with dt = Tf/(N-1)
    (Trajectory, types.jl:214-217), plus the zero-indexed index groups when given (e.g. {"x": 0, "y": 1, ...}).
    X[N, n_x], U[N, n_u] of ONE instance (row-major here = the reference's (n, N) column-major array on disk)."""
    X = np.asarray(X, dtype=np.float64); U = np.asarray(U, dtype=np.float64)
    if X.ndim != 2 or U.ndim != 2 or X.shape[0] != U.shape[0]:
        raise ValueError("X[N, n_x] and U[N, n_u] of one instance expected")
    N = X.shape[0]
    tree = {"traj": {"x_traj": X, "u_traj": U, "t_traj": np.linspace(0.0, float(tf), N)}}
    if ind_x:
        tree["ind_x"] = {k: np.int64(v) for k, v in ind_x.items()}
    if ind_u:
        tree["ind_u"] = {k: np.int64(v) for k, v in ind_u.items()}
    return write_h5(path, tree)


# ------------------------------------------------------------------------------------------------ reader
def _read_header(buf, addr):
    ver, _, nmsg, _, size = struct.unpack_from("<BBHII", buf, addr)
    if ver != 1:
        raise ValueError("only version-1 object headers are supported")
    p, end, out = addr + 16, addr + 16 + size, {}
    while p < end and len(out) < nmsg:
        mtype, msize, _ = struct.unpack_from("<HHB", buf, p)
        out[mtype] = bytes(buf[p + 8:p + 8 + msize])
        p += 8 + msize
    return out


def _read_group(buf, btree, heap):
    if buf[heap:heap + 4] != b"HEAP" or buf[btree:btree + 4] != b"TREE":
        raise ValueError("corrupt group")
    _, _, daddr = struct.unpack_from("<QQQ", buf, heap + 8)
    ntype, level, used = struct.unpack_from("<BBH", buf, btree + 4)
    if ntype != 0:
        raise ValueError("not a group B-tree")
    children = [struct.unpack_from("<Q", buf, btree + 24 + 8 + 16 * i)[0] for i in range(used)]
    out = {}
    for ch in children:
        if level > 0:
            raise ValueError("multi-level group B-trees are not supported")
        if buf[ch:ch + 4] != b"SNOD":
            raise ValueError("corrupt symbol-table node")
        nsym = struct.unpack_from("<H", buf, ch + 6)[0]
        for i in range(nsym):
            off, hdr, cache = struct.unpack_from("<QQI", buf, ch + 8 + 40 * i)
            end = buf.index(b"\0", daddr + off)
            out[bytes(buf[daddr + off:end]).decode()] = _read_object(buf, hdr)
    return out


def _read_object(buf, hdr):
    m = _read_header(buf, hdr)
    if 0x0011 in m:
        bt, hp = struct.unpack("<QQ", m[0x0011][:16])
        return _read_group(buf, bt, hp)
    space, dtype, layout = m[0x0001], m[0x0003], m[0x0008]
    rank = space[1]
    dims = struct.unpack_from(f"<{rank}Q", space, 8) if rank else ()
    cls, size = dtype[0] & 0x0F, struct.unpack_from("<I", dtype, 4)[0]
    np_dtype = {(1, 8): np.float64, (0, 8): np.int64, (1, 4): np.float32, (0, 4): np.int32}.get((cls, size))
    if np_dtype is None or layout[0] != 3 or layout[1] != 1:
        raise ValueError("unsupported dataset (contiguous little-endian float / int only)")
    addr, nbytes = struct.unpack_from("<QQ", layout, 2)
    arr = np.frombuffer(buf, dtype=np_dtype, count=nbytes // size, offset=addr).reshape(dims).copy() if nbytes else np.zeros(dims, np_dtype)
    return arr


def read_h5(path):
    """Nested dict of arrays of a file of the subset write_h5 emits (also what HDF5.jl's default writer produces for it)."""
    buf = open(path, "rb").read()
    if buf[:8] != SIG or buf[8] != 0:
        raise ValueError("not an HDF5 file with a version-0 superblock")
    eof = struct.unpack_from("<Q", buf, 40)[0]
    if eof != len(buf):
        raise ValueError("end-of-file address does not match the file size")
    hdr = struct.unpack_from("<Q", buf, 64)[0]
    return _read_object(buf, hdr)


# ------------------------------------------------------------------------------------------ environments
def load_iss_corner_mat(path):
    """The geometry of iss_corner.mat as ISSCorner{T}() / add_obstacles! build it (iss_corner.jl:11-23,52-63): every zone is a
    HyperRectangle(Vec3f0(corner1), Vec3f0(corner2 - corner1)), i.e. origin and widths rounded to Float32 and the maximum
    corner evaluated in Float32 (quirk q7); spheres are HyperSphere(Point3f0(center), Float32(radius)).  Returns the dict
    layout of data/iss_corner.json: {"keepin_zones": [{"lo", "hi"}...], "keepout_zones", "obstacle_rectangles",
    "obstacle_spheres": [{"center", "radius"}...]}.  Needs SciPy (the file is a MATLAB v5 blob)."""
    from scipy.io import loadmat
    m = loadmat(path, squeeze_me=True, struct_as_record=False)

    def box(z):
        c1 = np.asarray(z.corner1, dtype=np.float64).ravel()
        c2 = np.asarray(z.corner2, dtype=np.float64).ravel()
        origin = c1.astype(np.float32)
        widths = (c2 - c1).astype(np.float32)
        hi = (origin + widths).astype(np.float32)
        return {"lo": np.minimum(origin, hi).astype(np.float64).tolist(), "hi": np.maximum(origin, hi).astype(np.float64).tolist()}

    out = {"keepin_zones": [box(z) for z in np.atleast_1d(m["keepin_zones"])],
           "keepout_zones": [box(z) for z in np.atleast_1d(m["keepout_zones"])],
           "obstacle_rectangles": [box(z) for z in np.atleast_1d(m["rectangles"])] if "rectangles" in m else [],
           "obstacle_spheres": [{"center": np.asarray(z.center, dtype=np.float32).astype(np.float64).ravel().tolist(), "radius": float(np.float32(z.radius))}
                                for z in np.atleast_1d(m["spheres"])] if "spheres" in m else []}
    return out
