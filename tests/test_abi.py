"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, the
ctypes signatures cover the header, and the host-only entry points (no kernel launch) behave.
No compute calls are made: there is no GPU here and the library has no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "derfinder_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from derfinder_b200 import _lib
    return _lib.load()


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    from derfinder_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (dfb_[a-z0-9_]+)", out.stdout))
    decl = declared_symbols()
    assert len(decl) >= 45
    missing = [s for s in decl if s not in exported]
    assert not missing, f"declared but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == decl, "ctypes signatures and header disagree"


def test_library_is_sm100a_and_torch_free(lib):
    from derfinder_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libc10" not in ldd


def test_no_gpu_fails_loudly(lib):
    if lib.dfb_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.dfb_create(0, None, C.byref(h))
    assert rc < 0 and b"no CPU fallback" in lib.dfb_last_error()
    import derfinder_b200 as dfb
    with pytest.raises(RuntimeError):
        dfb.Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "derfinder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_chunk_lengths(lib):
    # R/preprocessCoverage.R:171-188, 218-227
    def chunks(N, cs, cores=1):
        k = lib.dfb_chunk_lengths(N, float(cs), cores, None, 0)
        out = np.zeros(k, dtype=np.int64)
        lib.dfb_chunk_lengths(N, float(cs), cores, out.ctypes.data_as(C.POINTER(C.c_int64)), k)
        return list(out)
    assert chunks(1434, 1e3) == [1000, 434]
    assert chunks(2000, 1000) == [1000, 1000]
    assert chunks(999, 1000) == [999]
    assert chunks(1000, 1000) == [1000]
    assert chunks(10, 0, 3) == [4, 4, 2]
    assert chunks(48129895, 5e6) == [5000000] * 9 + [3129895]


def test_cluster_maker_known_answers(lib):
    # tests/testthat/test_findRegions.R:5-11
    import derfinder_b200 as dfb
    pos = (np.array([1, 0, 1, 0, 1], bool), np.array([1000, 1000, 1000, 200, 1000]))
    ids, cnt = dfb.clusterMakerRle(pos)
    assert list(ids) == [1, 2] and list(cnt) == [1000, 2000]
    st, en = dfb.clusterMakerRle(pos, ranges=True)
    assert list(st) == [1, 1001] and list(en) == [1000, 3000]
    ids, cnt = dfb.clusterMakerRle(pos, maxGap=999)
    assert list(cnt) == [1000, 2000]
    ids, cnt = dfb.clusterMakerRle(pos, maxGap=1000)
    assert list(ids) == [1] and list(cnt) == [3000]
    from oracle import derfinder_oracle as O
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 50, size=400)
    vals = np.arange(400) % 2 == 1
    for g in (0, 3, 20, 49, 1000):
        a = dfb.clusterMakerRle((vals, lens), maxGap=g)
        b = O.cluster_maker_rle((vals, lens), maxGap=g)
        assert np.array_equal(a[1], b[1])
        a = dfb.clusterMakerRle((vals, lens), maxGap=g, ranges=True)
        b = O.cluster_maker_rle((vals, lens), maxGap=g, ranges=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_model_basis_host(lib, golden):
    from derfinder_b200 import _lib as L
    mod, mod0 = np.asfortranarray(golden.mod), np.asfortranarray(golden.mod0)
    q = np.zeros((4, 31))
    cen = C.c_int()
    rc = lib.dfb_model_basis(mod.ctypes.data_as(L.c_f64p), 31, 4, mod0.ctypes.data_as(L.c_f64p), 3,
                             q.ctypes.data_as(L.c_f64p), C.byref(cen))
    assert rc == 0 and cen.value == 1
    Q = q.T
    assert np.allclose(Q.T @ Q, np.eye(4), atol=1e-12)
    # first 3 columns span mod0, all 4 span mod
    P0 = Q[:, :3] @ Q[:, :3].T
    assert np.allclose(P0 @ golden.mod0, golden.mod0, atol=1e-9)
    assert np.allclose(Q @ Q.T @ golden.mod, golden.mod, atol=1e-9)
    # the basis reproduces the golden F-statistics through the C oracle of the engine arithmetic
    from oracle import c_oracle
    from oracle import derfinder_oracle as O
    F = c_oracle.fstats_basis(O.log2_libm(golden.cov + 32.0), Q.copy(), 3, None, True)
    assert np.max(np.abs(F - golden.fstats) / golden.fstats) < 1e-7
    Fp = c_oracle.fstats_projector(O.log2_libm(golden.cov + 32.0), golden.mod, golden.mod0)
    assert np.max(np.abs(Fp - golden.fstats) / golden.fstats) < 1e-7
    # error behaviour: rank deficiency / non-nested null
    bad = mod.copy()
    bad[:, 1] = bad[:, 0]
    assert lib.dfb_model_basis(bad.ctypes.data_as(L.c_f64p), 31, 4, mod0.ctypes.data_as(L.c_f64p), 3, None, None) < 0
    assert b"rank deficient" in lib.dfb_last_error() or b"nested" in lib.dfb_last_error()


def test_python_mirror_argument_checks(golden):
    import derfinder_b200 as dfb
    with pytest.raises(ValueError):
        dfb.filterData(golden.raw_rle, filter="two")
    with pytest.raises(ValueError):
        dfb.preprocessCoverage(dict(coverage=golden.cov))
    with pytest.raises(ValueError):
        dfb.calculateStats(dict(coverageProcessed=None), dict(mod=golden.mod, mod0=golden.mod0))
    with pytest.raises(NotImplementedError):
        dfb.findRegions(golden.position, golden.fstats, smooth=True)
    with pytest.raises(ValueError):
        dfb.regionMatrix({"chr21": golden.raw_rle}, cutoff=None, L=36)
    with pytest.warns(UserWarning):
        try:
            dfb.findRegions(golden.position, golden.fstats, maxClusterGap=1, maxRegionGap=5, cutoff=1)
        except RuntimeError:
            pass  # no GPU here: the call fails loudly after the reference's warning


def test_r_shim_typechecks():
    """r/src/rglue.c (the .Call layer of INTEGRATION.md) must agree with include/derfinder_b200.h.
    There is no R here, so it is compiled -fsyntax-only against stub R headers."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror=incompatible-pointer-types", "-Werror=implicit-function-declaration",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "r_stub"),
                        os.path.join(ROOT, "r", "src", "rglue.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every .Call the R glue makes is registered by the shim
    import re
    glue = open(os.path.join(ROOT, "r", "R", "engine.R")).read()
    shim = open(os.path.join(ROOT, "r", "src", "rglue.c")).read()
    for name in set(re.findall(r"\.Call\((dfbR_\w+)", glue)):
        assert '{"%s"' % name in shim, name
