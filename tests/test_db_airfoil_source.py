"""The device source of the database-airfoil model, compiled for the host (tests/native/db_airfoil_host.cpp includes
machupx_b200/csrc/db_airfoil.cuh, which is plain C++ on purpose), against the oracle's numpy restatement and the host getters of
machupx_b200.airfoil: values, central-difference slopes, zero-lift angle and the out-of-table flag, on the reference's two
(alpha, Rey) tables and on synthetic tables with one and three variables.  This is how the arithmetic of the table walk is
verified where no GPU exists; tests/test_database_airfoils.py runs the same code inside the kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from machupx_b200 import layout as L
from machupx_b200.airfoil import Airfoil
from oracle import lls_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("dbhost") / "libdbhost.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "native", "db_airfoil_host.cpp"), "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.db_host_evaluate.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, dp, dp, dp, dp, dp, ctypes.c_int, dp, ip]
    lib.db_host_evaluate.restype = None
    return lib


def _device_source(lib, af, what, alpha, Re, Mach, defl, cf):
    """db_evaluate() of the device source on the pool record the host ships to the GPU."""
    rec = np.ascontiguousarray(af.lls_pool(), dtype=np.float64)
    kinds = np.array([float(L.POLY_DOF_KINDS.index(d)) for d in af.dofs])
    n = len(alpha)
    arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))) for a in (alpha, Re, Mach, defl, cf)]
    out, oob = np.zeros(n), np.zeros(n, dtype=np.int32)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.db_host_evaluate(rec.ctypes.data_as(dp), len(af.dofs), kinds.ctypes.data_as(dp), n, *[a.ctypes.data_as(dp) for a in arrs], what,
                         out.ctypes.data_as(dp), oob.ctypes.data_as(ip))
    return out, oob.astype(bool)


class _T:
    """The slice of oracle.Tables the section functions of the oracle read."""

    def __init__(self, af, n, defl, cf):
        blk = af.lls_params()
        blk[L.AF_POLY_OFF] = float(L.AF_STRIDE)
        flat = np.concatenate((blk, af.lls_pool()))
        self.af = np.concatenate((flat, np.zeros((-flat.size) % L.AF_STRIDE))).reshape(-1, L.AF_STRIDE)
        self.n = n
        self.FLAP_DEFL, self.FLAP_CF = np.broadcast_to(defl, (n,)).astype(float), np.broadcast_to(cf, (n,)).astype(float)


def _compare(lib, af, alpha, Re, Mach, defl=0.0, cf=0.0, tol=1e-12):
    n = len(alpha)
    T = _T(af, n, defl, cf)
    mask = np.ones(n, dtype=bool)
    kw = dict(alpha=alpha, Rey=Re, Mach=Mach, trailing_flap_deflection=T.FLAP_DEFL, trailing_flap_fraction=T.FLAP_CF)
    for what in (0, 1, 2):
        got, oob = _device_source(lib, af, what, alpha, Re, Mach, defl, cf)
        ref = O._db_one(T, 0, mask, alpha, Re, Mach, what, None)
        host, found = af._db_lookup(what, kw)
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), what
        assert np.abs(got - host).max() <= tol * max(1.0, np.abs(ref).max()), what
        assert (oob == ~found).all()
    inside = ~_device_source(lib, af, 0, alpha, Re, Mach, defl, cf)[1]
    for dkind in (0, 1, 2):
        got, _ = _device_source(lib, af, 3 + dkind, alpha, Re, Mach, defl, cf)
        ref = O._db_one(T, 0, mask, alpha, Re, Mach, 0, dkind)
        assert np.abs(got - ref)[inside].max() <= 1e-9 * max(1.0, np.abs(ref[inside]).max()), dkind
    got, _ = _device_source(lib, af, 6, alpha, Re, Mach, defl, cf)
    ref = O._db_aL0(T, 0, mask, Re, Mach)
    ok = np.isfinite(ref)
    assert np.abs(got - ref)[ok].max() <= 1e-12
    return inside


@pytest.mark.skipif(not os.path.isfile(os.path.join(STAGED, "test", "NACA 2414.txt")), reason="reference test tables not staged")
@pytest.mark.parametrize("table", ["NACA 2414.txt", "NACA 0012.txt"])
def test_device_source_on_the_reference_tables(host_lib, table):
    af = Airfoil("t", {"type": "database", "input_file": os.path.join(STAGED, "test", table)})
    rng = np.random.default_rng(1)
    n = 400
    alpha = rng.uniform(-0.2, 0.3, n)
    Re = rng.uniform(5e5, 3e6, n)
    Re[:40] = rng.uniform(2e5, 4e6, 40)            # some points below / above the table
    alpha[40:60] = rng.uniform(-0.6, 0.6, 20)
    alpha[60:80] = np.radians(np.arange(-5.0, 5.0, 0.5))      # on grid lines
    Re[60:70] = 1e6                                           # on data points
    inside = _compare(host_lib, af, alpha, Re, np.full(n, 0.1))
    assert 300 < inside.sum() < n
    # the slope at an interior point is the plain central difference with airfoil_db's step
    a0, r0 = np.array([0.03]), np.array([8.1e5])
    cla, _ = _device_source(host_lib, af, 3, a0, r0, [0.0], 0.0, 0.0)
    hi, _ = _device_source(host_lib, af, 0, a0 + 0.001, r0, [0.0], 0.0, 0.0)
    lo, _ = _device_source(host_lib, af, 0, a0 - 0.001, r0, [0.0], 0.0, 0.0)
    assert abs(cla[0] - (hi[0] - lo[0]) / 0.002) < 1e-10


def _synthetic(tmp_path, dofs, rng, npts):
    cols = {"alpha": rng.uniform(-0.3, 0.3, npts), "Rey": rng.uniform(1e5, 1e6, npts), "Mach": rng.uniform(0.0, 0.6, npts),
            "trailing_flap_deflection": rng.uniform(-0.3, 0.3, npts), "trailing_flap_fraction": rng.uniform(0.0, 0.4, npts)}
    X = np.stack([cols[d] for d in dofs], axis=1)
    a = cols["alpha"] if "alpha" in dofs else np.zeros(npts)
    CL = 5.5 * a + 0.2 + (1e-7 * cols["Rey"] if "Rey" in dofs else 0.0) + (0.3 * cols["Mach"] ** 2 if "Mach" in dofs else 0.0)
    path = tmp_path / ("db_%d.txt" % len(dofs))
    with open(path, "w") as fh:
        fh.write("# " + "\t".join(list(dofs) + ["CL", "CD", "Cm"]) + "\n")
        for row, cl in zip(X, CL):
            fh.write("\t".join("%.12e" % v for v in list(row) + [cl, 0.01 + 0.05 * cl * cl, -0.05 + 0.01 * cl]) + "\n")
    return Airfoil("s", {"type": "database", "input_file": str(path)})


def test_device_source_with_one_and_three_variables(host_lib, tmp_path):
    rng = np.random.default_rng(2)
    one = _synthetic(tmp_path, ("alpha",), rng, 60)
    n = 200
    alpha = rng.uniform(-0.35, 0.35, n)
    inside = _compare(host_lib, one, alpha, np.full(n, 3e5), np.full(n, 0.2))
    assert 100 < inside.sum() < n
    expect = np.interp(alpha, np.sort(one._db_points[:, 0]), one._db_values[np.argsort(one._db_points[:, 0]), 0])
    got, _ = _device_source(host_lib, one, 0, alpha, 3e5, 0.2, 0.0, 0.0)
    assert np.abs(got - expect)[inside].max() < 1e-12
    three = _synthetic(tmp_path, ("alpha", "Mach", "trailing_flap_deflection"), rng, 300)
    defl = rng.uniform(-0.25, 0.25, n)
    inside = _compare(host_lib, three, rng.uniform(-0.25, 0.25, n), np.full(n, 3e5), rng.uniform(0.05, 0.5, n), defl=defl, cf=0.2)
    assert inside.sum() > 50
