"""Host geometry of this package (Scene / Airplane / WingSegment written from scratch) against the tables the
golden generator extracted from the unmodified reference objects (oracle/gen_golden.py runs
machupx_b200.tables.build_tables on machupX.Scene).  No GPU needed: constructing a Scene only builds O(N) tables."""
import copy

import numpy as np
import pytest

from golden_utils import CASES, load
from machupx_b200 import Scene, layout as L
from machupx_b200.tables import build_aircraft_state


@pytest.mark.parametrize("name", CASES)
def test_tables_match_reference_objects(name):
    tables, ac_state, g, meta = load(name)
    scene = Scene(copy.deepcopy(meta["input"]))
    mine = scene._tables
    n = tables["n"]
    assert mine["n"] == n and mine["ld"] == tables["ld"]
    assert mine["n_segments"] == tables["n_segments"] and mine["n_aircraft"] == tables["n_aircraft"]
    assert mine["flags"] == tables["flags"]
    assert [list(x) for x in mine["segment_names"]] == [list(x) for x in meta["segment_names"]]
    np.testing.assert_array_equal(mine["itab"][:, :n], tables["itab"][:, :n])
    ref, got = tables["tab"][:, :n], mine["tab"][:, :n]
    # geometry is O(1-100) ft; the host path repeats the reference's procedures (same quadrature, same gradient),
    # so agreement is at rounding level; rotations go through a matrix here and a quaternion product there
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
    err = np.abs(got - ref) / scale
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < 5e-14, "field %d cp %d: %.3e" % (worst[0], worst[1], err.max())
    np.testing.assert_allclose(mine["airfoils"], tables["airfoils"], rtol=0, atol=0)
    from machupx_b200 import layout as L
    mine = build_aircraft_state(scene)
    np.testing.assert_allclose(mine[:, :L.AC_ZF], ac_state[:, :L.AC_ZF], rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(np.linalg.norm(mine[:, L.AC_ZF:L.AC_ZF + 3], axis=1), 1.0, rtol=1e-14)   # body z axis (constrain_vortex_sheet)


def test_state_and_control_changes_reach_the_tables():
    tables, ac_state, g, meta = load("testplane_nonlinear")
    scene = Scene(copy.deepcopy(meta["input"]))
    base = scene._tables["tab"].copy()
    scene.set_aircraft_control_state({"elevator": 4.0})
    from machupx_b200.tables import flap_rows
    rows = flap_rows(scene)
    assert np.abs(rows[L.F_FLAP_DEFL]).max() == pytest.approx(np.radians(4.0))
    assert scene._current_controls_key() != scene._controls_key
    scene.set_aircraft_state({"position": [0, 0, 1000], "velocity": 100, "alpha": 2.0, "orientation": [10.0, 5.0, 0.0]})
    moved = scene._tables["tab"]
    assert np.abs(moved[L.F_PC:L.F_PC + 3] - base[L.F_PC:L.F_PC + 3]).max() > 1e-3     # rotated geometry rows
    np.testing.assert_allclose(moved[L.F_DS], base[L.F_DS])                             # scalars unchanged
