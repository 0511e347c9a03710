"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/ohp.h declares; the host-only entry points (.osh reader) agree with the
oracle reader on the reference's fixture files; compute entry points fail LOUDLY without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import omegahpetsc_b200 as ohp
from conftest import GOLDEN, ROOT, golden_mesh_path

MESH_DIRS = sorted(os.listdir(os.path.join(GOLDEN, "meshes")))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(ohp.LIB_PATH):
        ohp.build()
    return ohp.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ohp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ohp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(L):
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.ohp_version() == 1


def test_library_is_sm100a_only(L):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", ohp.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("name", MESH_DIRS)
def test_osh_reader_matches_oracle_reader(L, name):
    from oracle import osh_oracle as oo
    o = ohp.Osh(golden_mesh_path(name))
    m = oo.read_osh(golden_mesh_path(name))
    assert (o.dim, o.nverts, o.version, o.nparts) == (m["dim"], m["nverts"], 9, 1)
    assert np.array_equal(o.elem_verts(), oo.elem_verts(m))
    assert np.array_equal(o.coords(), oo.coords(m))
    for d in range(o.dim + 1):
        for tag, (nc, arr) in m["tags"][d].items():
            got = o.tag(d, tag)
            assert got.dtype == arr.dtype and np.array_equal(got.reshape(-1), arr.reshape(-1)), (d, tag)
    assert o.nents(1) == m["down"][1].size // 2
    o.close()


def test_osh_reader_on_written_mesh(L, tmp_path):
    import oracle
    from oracle import osh_oracle as oo
    cells, coords = oracle.box2d(9, 4)
    oo.write_osh(str(tmp_path / "b.osh"), 2, coords, cells)
    o = ohp.Osh(str(tmp_path / "b.osh"))
    assert np.array_equal(o.elem_verts(), cells) and np.array_equal(o.coords(), coords)


def test_osh_reader_3d_roundtrip(L, tmp_path):
    """There is no 3-D `.osh` fixture in the reference: the tet path (tet -> tri -> edge -> vert with
    alignment codes) is checked by round trip through the oracle writer plus self-consistency (positive volumes,
    every tet recovered exactly), for the C++ reader and the oracle reader alike."""
    import oracle
    from oracle import osh_oracle as oo
    cells, coords = oracle.box3d(3, 2, 4)
    rng = np.random.default_rng(2)
    perm = rng.permutation(cells.shape[0])   # break the generator's regular face-sharing order
    cells = cells[perm]
    oo.write_osh(str(tmp_path / "tets.osh"), 3, coords, cells)
    m = oo.read_osh(str(tmp_path / "tets.osh"))
    assert m["consumed"] == m["total"] and m["dim"] == 3
    assert np.array_equal(oo.elem_verts(m), cells)
    o = ohp.Osh(str(tmp_path / "tets.osh"))
    assert (o.dim, o.nverts, o.nelems) == (3, coords.shape[0], cells.shape[0])
    ev = o.elem_verts()
    assert np.array_equal(ev, cells) and np.array_equal(o.coords(), coords)
    X = coords[ev]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6.0
    assert (vol > 0).all() and vol.sum() == pytest.approx(1.0, rel=1e-12)
    assert o.nents(2) == m["down"][2].size // 3 and o.nents(1) == m["down"][1].size // 2


def test_osh_reader_errors(L, tmp_path):
    with pytest.raises(ohp.OhpError) as e:
        ohp.Osh(str(tmp_path / "missing.osh"))
    assert e.value.code == 65
    d = tmp_path / "bad.osh"
    d.mkdir()
    (d / "version").write_text("9\n")
    (d / "nparts").write_text("1\n")
    (d / "0.osh").write_bytes(b"\x00\x01garbage")
    with pytest.raises(ohp.OhpError) as e:
        ohp.Osh(str(d))
    assert e.value.code == 79 and "magic" in str(e.value)
    good = open(os.path.join(golden_mesh_path("23elm.osh"), "0.osh"), "rb").read()
    (d / "0.osh").write_bytes(good[: len(good) // 2])
    with pytest.raises(ohp.OhpError):
        ohp.Osh(str(d))
    (d / "0.osh").write_bytes(good + b"\x00")
    with pytest.raises(ohp.OhpError) as e:
        ohp.Osh(str(d))
    assert "trailing" in str(e.value)
    with pytest.raises(ohp.OhpError) as e:
        ohp.Osh(golden_mesh_path("23elm.osh"), part=1)
    assert e.value.code == 63


def test_no_gpu_means_loud_failure(L):
    """The product has no CPU path: without a device ohp_init must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ohp.OhpError) as e:
        ohp.Context(0)
    assert e.value.code == 97 and "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing under the package, include/ or the facade may reference oracle/."""
    bad = []
    for base in ("omegahpetsc_b200", "include", "examples"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"\boracle\b", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_default_options(L):
    o = ohp.KspOptions()
    assert L.ohp_ksp_default_options(ctypes.byref(o)) == 0
    assert (o.rtol, o.abstol, o.dtol, o.max_it, o.pc) == (1e-5, 1e-50, 1e5, 10000, 0)
    s = ohp.SnesOptions()
    assert L.ohp_snes_default_options(ctypes.byref(s)) == 0
    assert (s.rtol, s.abstol, s.stol, s.max_it) == (1e-8, 1e-50, 1e-8, 50)


def test_ctypes_mirrors_match_the_header(tmp_path):
    """The Python harness mirrors ohp.h's structs by hand (ctypes): sizes and field offsets must equal what a C compiler
    lays out from the header itself (fields were appended in round 2: ksp_type, algo, the patch tables of the view)."""
    import ctypes
    import subprocess

    import omegahpetsc_b200 as ohp
    mirrors = {"ohp_ksp_options": ohp.KspOptions, "ohp_ksp_result": ohp.KspResult, "ohp_snes_options": ohp.SnesOptions,
               "ohp_snes_result": ohp.SnesResult}
    lines = []
    for cname, cls in mirrors.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "ohp.h"\nint main(void) {\n%s\nreturn 0; }\n' % "\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
