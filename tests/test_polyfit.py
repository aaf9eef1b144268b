"""poly_fit airfoils on the host and in the oracle (CPU).  The device path is tests/test_polyfit_gpu.py."""
import copy

import numpy as np
import pytest

from golden_utils import load
from machupx_b200 import Scene, layout as L
from machupx_b200.airfoil import Airfoil
from machupx_b200.tables import build_aircraft_state
from oracle import lls_oracle as O
from polyfit_utils import linear_as_poly, with_poly_airfoils

RE_FIT = {"type": "poly_fit", "degrees_of_freedom": ["alpha", "Rey", "Mach"],
          "limits": [[-0.4, 0.4], [1e4, 1e8], [0.0, 0.6]],
          "fit_degrees": {"CL": [3, 1, 1], "CD": [2, 1, 0], "Cm": [1, 0, 1]},
          "fit_coefs": {"CL": [0.05, 0.02, 1e-9, 2e-9, 6.0, 0.4, 3e-8, 1e-8, 0.1, 0.0, 0.0, 0.0, -4.0, 0.0, 0.0, 0.0],
                        "CD": [0.008, -2e-10, 0.001, 0.0, 0.09, 1e-9], "Cm": [-0.02, 0.01, 0.03, 0.0]}}


def test_host_getters_match_hand_expansion():
    a = Airfoil("fit", RE_FIT)
    al, Re, M = np.array([-0.1, 0.0, 0.2]), np.array([5e5, 1e6, 3e6]), np.array([0.1, 0.2, 0.3])
    c = np.array(RE_FIT["fit_coefs"]["CL"]).reshape(4, 2, 2)
    CL = sum(c[i, j, k] * al ** i * Re ** j * M ** k for i in range(4) for j in range(2) for k in range(2))
    CLa = sum(i * c[i, j, k] * al ** (i - 1) * Re ** j * M ** k for i in range(1, 4) for j in range(2) for k in range(2))
    CLRe = sum(c[i, 1, k] * al ** i * M ** k for i in range(4) for k in range(2))
    CLM = sum(c[i, j, 1] * al ** i * Re ** j for i in range(4) for j in range(2))
    kw = dict(alpha=al, Rey=Re, Mach=M)
    assert np.allclose(a.get_CL(**kw), CL, rtol=1e-14) and np.allclose(a.get_CLa(**kw), CLa, rtol=1e-14)
    assert np.allclose(a.get_CLRe(**kw), CLRe, rtol=1e-14) and np.allclose(a.get_CLM(**kw), CLM, rtol=1e-14)
    aL0 = a.get_aL0(Rey=Re, Mach=M)
    assert np.abs(a.get_CL(alpha=aL0, Rey=Re, Mach=M)).max() < 1e-15
    assert a.out_of_bounds(alpha=np.array([0.0, 0.5]), Rey=1e6, Mach=0.1).tolist() == [False, True]
    pool = a.lls_pool()
    assert pool[6] == 16 and list(pool[7:10]) == [3, 1, 1] and len(pool) == 6 + (1 + 3 + 16) + (1 + 3 + 6) + (1 + 3 + 4)


@pytest.mark.parametrize("extra", [(), (("Rey", (1e3, 1e9)), ("trailing_flap_fraction", (0.0, 1.0)))])
def test_oracle_with_linear_equivalent_fits_reproduces_the_reference(extra):
    tables, ac_state, g, meta = load("testplane_nonlinear")
    scene = Scene(with_poly_airfoils(meta["input"], extra))
    assert (scene._tables["airfoils"][:scene._tables["n_airfoils"], L.AF_TYPE] == 1.0).all()
    T = O.Tables(scene._tables, build_aircraft_state(scene))
    sp = meta["solver"]
    res = O.solve(T, solver_type=sp["type"], convergence=sp["convergence"], relaxation=sp["relaxation"], max_iterations=sp["max_iterations"])
    assert res["iterations"] == meta["iterations"]
    assert np.abs(res["gamma"] - g["gamma"]).max() / np.abs(g["gamma"]).max() < 1e-11
    from golden_utils import segment_sums
    ref = segment_sums(g, tables)
    assert np.abs(res["seg_fm"] - ref).max() / np.abs(ref).max() < 1e-11


def test_linear_as_poly_matches_linear_airfoil_pointwise():
    spec = {"type": "linear", "aL0": -0.0368, "CLa": 6.1976, "CmL0": -0.0525, "Cma": 0.0326, "CD0": 0.00569, "CD1": -0.0045, "CD2": 0.0104}
    lin, fit = Airfoil("l", spec), Airfoil("p", linear_as_poly(spec))
    al = np.linspace(-0.15, 0.2, 9)
    for fn in ("get_CL", "get_CD", "get_Cm", "get_CLa"):
        assert np.allclose(getattr(lin, fn)(alpha=al), getattr(fit, fn)(alpha=al), rtol=0, atol=1e-15)
    assert abs(fit.get_aL0() - spec["aL0"]) < 1e-16
