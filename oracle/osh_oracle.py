"""ORACLE (test infrastructure, not product): numpy/zlib restatement of the Omega_h
`.osh` (binary format version 9) directory reader and writer, plus the
tri->vert / tet->vert derivation that `Mesh::ask_elem_verts()` performs.

PARITY UNPINNED: the reference repository holds no golden vectors for this path
(its CTests check exit status only, /root/reference/CMakeLists.txt:66-120), and
Omega_h itself is an un-vendored dependency (README.md:151 / :547, tag v9.32.1),
so the format below restates Omega_h's published stream layout and is anchored
on the reference's own fixture files (24k.osh, 23elmMesh/, xgc-picparts/), which
must decode to the last byte, and on the reference call sites
ex12.cpp:813,818 (`Omega_h::binary::read`) and ex12.cpp:591,634,784
(`ask_elem_verts`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
import os
import struct
import zlib

import numpy as np

MAGIC = b"\xa1\x1a"
# Omega_h tag type codes (Omega_h_defines.h: OMEGA_H_I8=0, I32=2, I64=3, F64=5)
_TYPES = {0: np.dtype("<i1"), 2: np.dtype("<i4"), 3: np.dtype("<i8"), 5: np.dtype("<f8")}
_CODES = {np.dtype("int8"): 0, np.dtype("int32"): 2, np.dtype("int64"): 3, np.dtype("float64"): 5}


class _Stream:
    def __init__(self, data):
        self.b = data
        self.o = 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def array(self, dtype, compressed):
        n = self.take("i")
        if compressed:
            zbytes = self.take("q")
            raw = zlib.decompress(self.b[self.o:self.o + zbytes])
            self.o += zbytes
        else:
            nb = n * dtype.itemsize
            raw = self.b[self.o:self.o + nb]
            self.o += nb
        a = np.frombuffer(raw, dtype=dtype)
        if a.size != n:
            raise ValueError("osh array length mismatch: header %d, payload %d" % (n, a.size))
        return a.copy()


def read_osh_stream(data):
    """Decode one `<rank>.osh` stream. Returns a dict:
    family, dim, comm_size, comm_rank, parting, nghost, hints, nverts,
    down[d] (int32, d=1..dim), codes[d] (int8, d=2..dim),
    tags[d] = {name: (ncomps, array)}, owners[d] = (ranks, idxs) if comm_size>1.
    """
    s = _Stream(data)
    if s.b[:2] != MAGIC:
        raise ValueError("not an Omega_h binary stream (bad magic)")
    s.o = 2
    m = {}
    compressed = bool(s.take("b"))
    m["family"] = s.take("b")
    m["dim"] = dim = s.take("b")
    m["comm_size"] = s.take("i")
    m["comm_rank"] = s.take("i")
    m["parting"] = s.take("b")
    m["nghost"] = s.take("i")
    have_hints = s.take("b")
    hints = None
    if have_hints:
        naxes = s.take("i")
        hints = [s.take("ddd") for _ in range(naxes)]
    m["hints"] = hints
    m["nverts"] = s.take("i")
    m["down"], m["codes"] = {}, {}
    for d in range(1, dim + 1):
        m["down"][d] = s.array(_TYPES[2], compressed)
        if d > 1:
            m["codes"][d] = s.array(_TYPES[0], compressed)
    m["tags"] = {}
    m["owners"] = {}
    for d in range(0, dim + 1):
        ntags = s.take("i")
        tags = {}
        for _ in range(ntags):
            ln = s.take("i")
            name = s.b[s.o:s.o + ln].decode()
            s.o += ln
            ncomps = s.take("b")
            typ = s.take("b")
            tags[name] = (ncomps, s.array(_TYPES[typ], compressed))
        m["tags"][d] = tags
        if m["comm_size"] > 1:
            ranks = s.array(_TYPES[2], compressed)
            idxs = s.array(_TYPES[2], compressed)
            m["owners"][d] = (ranks, idxs)
    # trailer: class sets (i32 count, none in the fixtures) + i8 "has parents"
    m["nsets"] = s.take("i")
    if m["nsets"] != 0:
        raise ValueError("class sets not supported by the oracle reader")
    m["trailer"] = s.take("b")
    m["consumed"] = s.o
    m["total"] = len(s.b)
    return m


def read_osh(path, rank=0):
    """Read `<path>/<rank>.osh` of an Omega_h mesh directory (ex12.cpp:813,818)."""
    with open(os.path.join(path, "version")) as f:
        version = int(f.read().split()[0])
    with open(os.path.join(path, "nparts")) as f:
        nparts = int(f.read().split()[0])
    with open(os.path.join(path, "%d.osh" % rank), "rb") as f:
        m = read_osh_stream(f.read())
    m["version"] = version
    m["nparts"] = nparts
    return m


def _elem_verts_from_down(m):
    """Restates Omega_h's derivation of element->vertex connectivity from the
    stored one-level downward adjacencies + alignment codes
    (`Mesh::ask_elem_verts`, reached from ex12.cpp:591,634,784).

    code layout (Omega_h_align.hpp): bit0 = flipped, bits1-2 = rotation, bits3+ = which_down.
    For a triangle, vertex i is the first vertex of its edge i as seen from the
    triangle: edge i of a triangle runs (vert i -> vert i+1); the stored code says how
    the edge's own (v0,v1) is rotated to obtain that, so vertex i = edgeVerts[rot].
    """
    dim = m["dim"]
    ev = m["down"][1].reshape(-1, 2)
    if dim == 1:
        return ev.copy()
    te = m["down"][2].reshape(-1, 3)
    tc = m["codes"][2].reshape(-1, 3).astype(np.int32)
    rot = (tc >> 1) & 3
    if np.any(rot > 1):
        raise ValueError("edge rotation > 1 in tri->edge codes")
    tri = np.take_along_axis(ev[te], rot[:, :, None], axis=2)[:, :, 0]
    if dim == 2:
        return np.ascontiguousarray(tri, dtype=np.int32)
    # dim 3: tet -> tri (4 per tet) with codes; simplex_down_template(3,2):
    # face 0 = (0,2,1), 1 = (0,1,3), 2 = (1,2,3), 3 = (2,0,3)
    # (Omega_h_simplex.hpp). Tet verts 0..2 come from face 0, vert 3 from face 1.
    tt = m["down"][3].reshape(-1, 4)
    cc = m["codes"][3].reshape(-1, 4).astype(np.int32)

    def aligned(face_col):
        f = tt[:, face_col]
        c = cc[:, face_col]
        flip = (c & 1).astype(bool)
        r = (c >> 1) & 3
        v = tri[f]  # (n,3) face's own vertices
        # align_adj(code, ...): rotate by r then flip swaps entries 1 and 2
        idx = (np.arange(3)[None, :] + (3 - r)[:, None]) % 3
        out = np.take_along_axis(v, idx, axis=1)
        out[flip] = out[flip][:, [0, 2, 1]]
        return out

    f0 = aligned(0)  # (0,2,1)
    f1 = aligned(1)  # (0,1,3)
    tet = np.stack([f0[:, 0], f0[:, 2], f0[:, 1], f1[:, 2]], axis=1)
    return np.ascontiguousarray(tet, dtype=np.int32)


def elem_verts(m):
    return _elem_verts_from_down(m)


def coords(m):
    nc, a = m["tags"][0]["coordinates"]
    return a.reshape(-1, nc)


# ---------------------------------------------------------------------------
# writer (used to build synthetic .osh inputs and re-encoded fixtures)
# ---------------------------------------------------------------------------

def _w_array(out, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    z = zlib.compress(a.tobytes(), 6)
    out.append(struct.pack("<iq", a.size, len(z)))
    out.append(z)


def edges_from_tris(tris):
    """Build edge->vert, tri->edge and alignment codes for triangles so that the
    reader's derivation reproduces `tris` exactly. Edge k of triangle (a,b,c) is
    (a,b),(b,c),(c,a); an edge is stored with the orientation of its first use."""
    tris = np.asarray(tris, dtype=np.int64)
    n = tris.shape[0]
    a = tris[:, [0, 1, 2]].reshape(-1)
    b = tris[:, [1, 2, 0]].reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * (tris.max() + 1) + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    ev = np.stack([a[first], b[first]], axis=1).astype(np.int32)
    te = inv.reshape(n, 3).astype(np.int32)
    # rot=0 when the triangle sees the edge as stored, rot=1 when reversed
    rot = (ev[inv, 0] != a).astype(np.int8)
    codes = (rot << 1).astype(np.int8).reshape(n, 3)
    return ev, te, codes


def faces_from_tets(tets):
    """Build tri->vert (stored faces), tet->tri and alignment codes for tetrahedra so that the reader's
    derivation (faces 0 = (0,2,1), 1 = (0,1,3), 2 = (1,2,3), 3 = (2,0,3) of the simplex down template;
    code bit 0 = flip, bits 1-2 = rotation) reproduces `tets` exactly.  A face is stored with the vertex
    order of its first use; every later use gets the (rotation, flip) that maps the stored triple onto it."""
    tets = np.asarray(tets, dtype=np.int64)
    template = ((0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3))
    faces, index = [], {}
    tt = np.empty((tets.shape[0], 4), dtype=np.int32)
    cc = np.empty((tets.shape[0], 4), dtype=np.int8)
    for t, tet in enumerate(tets):
        for f, tpl in enumerate(template):
            used = tuple(int(tet[i]) for i in tpl)
            key = tuple(sorted(used))
            if key not in index:
                index[key] = len(faces)
                faces.append(used)
            fid = index[key]
            stored = faces[fid]
            code = None
            for flip in (0, 1):
                for r in range(3):
                    out = [stored[(k + 3 - r) % 3] for k in range(3)]
                    if flip:
                        out[1], out[2] = out[2], out[1]
                    if tuple(out) == used:
                        code = (r << 1) | flip
                        break
                if code is not None:
                    break
            assert code is not None
            tt[t, f] = fid
            cc[t, f] = code
    return np.asarray(faces, dtype=np.int32), tt, cc


def write_osh(path, dim, coords_arr, elems, tags=None, rank=0, nparts=1):
    """Write a serial `.osh` directory (version 9) for a triangle (dim=2) or tetrahedron (dim=3) mesh.
    `tags` = {d: {name: (ncomps, array)}} adds extra tags."""
    if dim not in (2, 3):
        raise NotImplementedError("oracle writer covers triangles and tetrahedra")
    os.makedirs(path, exist_ok=True)
    tt = cc = None
    tris = elems
    if dim == 3:
        tris, tt, cc = faces_from_tets(elems)
    ev, te, codes = edges_from_tris(tris)
    nverts = int(np.asarray(coords_arr).reshape(-1, dim).shape[0])
    out = [MAGIC, struct.pack("<bbbiibib", 1, 0, dim, 1, 0, 0, 0, 0), struct.pack("<i", nverts)]
    _w_array(out, ev.reshape(-1), "<i4")
    _w_array(out, te.reshape(-1), "<i4")
    _w_array(out, codes.reshape(-1), "<i1")
    if dim == 3:
        _w_array(out, tt.reshape(-1), "<i4")
        _w_array(out, cc.reshape(-1), "<i1")
    tags = tags or {}
    for d in range(dim + 1):
        dt = dict(tags.get(d, {}))
        if d == 0:
            dt = {"coordinates": (dim, np.asarray(coords_arr, dtype=np.float64).reshape(-1)), **dt}
        out.append(struct.pack("<i", len(dt)))
        for name, (ncomps, arr) in dt.items():
            arr = np.asarray(arr)
            nb = name.encode()
            out.append(struct.pack("<i", len(nb)))
            out.append(nb)
            out.append(struct.pack("<bb", ncomps, _CODES[arr.dtype]))
            _w_array(out, arr.reshape(-1), arr.dtype.newbyteorder("<"))
    out.append(struct.pack("<ib", 0, 0))
    with open(os.path.join(path, "%d.osh" % rank), "wb") as f:
        f.write(b"".join(out))
    with open(os.path.join(path, "nparts"), "w") as f:
        f.write("%d\n" % nparts)
    with open(os.path.join(path, "version"), "w") as f:
        f.write("9\n")
