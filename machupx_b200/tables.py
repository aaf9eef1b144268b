"""Flattens a Scene (aircraft -> wings -> segments -> control points) into the O(N) tables liblls consumes.

This is the whole host->device data contract of the hot path: everything the reference
materialises as N x N x 3 arrays (``Airplane._calculate_geometry`` airplane.py:569-689,
``Scene._perform_geometry_and_atmos_calcs`` scene.py:431-515) is recomputed per pair on
the device from these per-control-point rows.

``build_tables`` is duck-typed on purpose: it only reads attributes that the reference's
``Scene``/``Airplane``/``WingSegment`` also expose, so the golden-vector generator
(oracle/gen_golden.py) can run it on the unmodified reference objects and the tests can
compare the result with what this package's own geometry code produces.
"""
import numpy as np

from . import layout as L
from .rotations import body_to_earth


def airfoil_param_block(airfoil):
    """AF_STRIDE doubles describing one airfoil for the device section model."""
    if hasattr(airfoil, "lls_params"):
        return np.asarray(airfoil.lls_params(), dtype=np.float64)
    # oracle harness: the airfoil_db stand-in used with the reference (oracle/ref_stubs)
    blk = np.zeros(L.AF_STRIDE)
    blk[L.AF_TYPE] = 0.0
    blk[L.AF_AL0] = airfoil._aL0
    blk[L.AF_CLA] = airfoil._CLa
    blk[L.AF_CML0] = airfoil._CmL0
    blk[L.AF_CMA] = airfoil._Cma
    blk[L.AF_CD0] = airfoil._CD0
    blk[L.AF_CD1] = airfoil._CD1
    blk[L.AF_CD2] = airfoil._CD2
    blk[L.AF_CLMAX] = airfoil._CL_max
    return blk


def flap_terms(delta, c_f):
    """Flap contributions of the linear section model (see machupx_b200/airfoil.py)."""
    from .airfoil import flap_effectiveness
    eps, dcm = flap_effectiveness(delta, c_f)
    return delta * eps, delta * dcm, 0.002 * np.abs(delta) * 180.0 / np.pi


def flap_rows(scene):
    """The table rows that change with the control state only: {field index: (n,) array} in scene order."""
    airplanes = list(scene._airplane_objects)
    n = int(sum(a.N for a in airplanes))
    rows = {f: np.zeros(n) for f in (L.F_FLAP_DA, L.F_FLAP_DCM, L.F_FLAP_DCD, L.F_FLAP_DEFL, L.F_FLAP_CF)}
    k = 0
    for ap in airplanes:
        for seg in ap.segments:
            sl = slice(k, k + seg.N)
            delta = np.asarray(seg._delta_flap, dtype=float)
            c_f = np.asarray(seg._cp_c_f, dtype=float)
            da, dcm, dcd = flap_terms(delta, c_f)
            rows[L.F_FLAP_DA][sl] = da
            rows[L.F_FLAP_DCM][sl] = dcm
            rows[L.F_FLAP_DCD][sl] = dcd
            rows[L.F_FLAP_DEFL][sl] = delta
            rows[L.F_FLAP_CF][sl] = c_f
            k += seg.N
    return rows


def solver_flags(scene):
    f = 0
    if scene._use_swept_sections:
        f |= L.FLAG_SWEPT_SECTIONS
    if scene._use_total_velocity:
        f |= L.FLAG_TOTAL_VELOCITY
    if scene._use_in_plane:
        f |= L.FLAG_IN_PLANE
    if scene._match_machup_pro:
        f |= L.FLAG_MATCH_MACHUP_PRO
    if getattr(scene, "_constrain_vortex_sheet", False):
        f |= L.FLAG_CONSTRAIN_VORTEX_SHEET
    return f


def build_tables(scene):
    """Returns dict(tab (NF, ld) f64, itab (NI, ld) i32, airfoils (n_af, AF_STRIDE) f64, n, ld,
    n_aircraft, n_segments, segment_names, segment_slices, aircraft_slices)."""
    airplanes = list(scene._airplane_objects)
    n = int(sum(a.N for a in airplanes))
    ld = L.padded_ld(n)
    tab = np.zeros((L.NF, ld))
    itab = np.zeros((L.NI, ld), dtype=np.int32)

    # benign padding: unit vectors / positive scalars so padded lanes never divide by zero
    tab[L.F_UA, n:] = 1.0
    tab[L.F_UN + 1, n:] = 1.0
    tab[L.F_US + 2, n:] = 1.0
    for f in (L.F_CBAR, L.F_DS, L.F_CSWI, L.F_RHO, L.F_NU, L.F_SOS):
        tab[f, n:] = 1.0
    tab[L.F_PC, n:] = 1e30
    itab[L.I_WING_BEGIN, n:] = -1
    itab[L.I_WING_END, n:] = -1
    itab[L.I_AIRCRAFT, n:] = -1

    airfoil_ids = {}
    airfoil_blocks = []
    airfoil_pools = []

    def af_id(obj):
        key = id(obj)
        if key not in airfoil_ids:
            airfoil_ids[key] = len(airfoil_blocks)
            airfoil_blocks.append(airfoil_param_block(obj))
            airfoil_pools.append(np.asarray(obj.lls_pool(), dtype=np.float64) if hasattr(obj, "lls_pool") and
                                 getattr(obj, "_type", "linear") in ("poly_fit", "database") else None)
        return airfoil_ids[key]

    swept = bool(scene._use_swept_sections)
    seg_names, seg_slices, ac_slices = [], [], []
    base = 0
    seg_counter = 0
    for ia, ap in enumerate(airplanes):
        na = ap.N
        sl = slice(base, base + na)
        ac_slices.append(sl)
        q = np.asarray(ap.q, dtype=float)

        def put3(f, body_vecs):
            tab[f:f + 3, sl] = body_to_earth(q, np.asarray(body_vecs)).T

        put3(L.F_PC, ap.PC)
        put3(L.F_P0, ap.P0)
        put3(L.F_P1, ap.P1)
        put3(L.F_P0J, ap.P0_joint)
        put3(L.F_P1J, ap.P1_joint)
        u_s_body = np.asarray(ap.u_s)
        pc_deriv = u_s_body / np.linalg.norm(u_s_body[:, 1:], axis=1, keepdims=True)  # airplane.py:564
        put3(L.F_PCD, pc_deriv)
        put3(L.F_UAU, ap.u_a_unswept)
        put3(L.F_UNU, ap.u_n_unswept)
        if swept:  # scene.py:451-458
            put3(L.F_UA, ap.u_a)
            put3(L.F_UN, ap.u_n)
            put3(L.F_US, ap.u_s)
        else:
            put3(L.F_UA, ap.u_a_unswept)
            put3(L.F_UN, ap.u_n_unswept)
            put3(L.F_US, ap.u_s_unswept)
        put3(L.F_DL, ap.dl)
        put3(L.F_RCG, ap.PC_CG)
        tab[L.F_SPC, sl] = ap.PC_span_locs
        tab[L.F_SP0, sl] = ap.P0_span_locs
        tab[L.F_SP1, sl] = ap.P1_span_locs
        sweep = np.asarray(ap.section_sweep)
        if swept:  # scene.py:423-426
            c_lambda = np.cos(sweep)
            tab[L.F_CBAR, sl] = np.asarray(ap.c_bar) * c_lambda
            tab[L.F_CSWI, sl] = 1.0 / c_lambda
        else:
            tab[L.F_CBAR, sl] = ap.c_bar
            tab[L.F_CSWI, sl] = 1.0
        tab[L.F_DS, sl] = ap.dS
        itab[L.I_AIRCRAFT, sl] = ia
        for ws in ap.wing_slices:
            itab[L.I_WING_BEGIN, base + ws.start:base + ws.stop] = base + ws.start
            itab[L.I_WING_END, base + ws.start:base + ws.stop] = base + ws.stop

        k = base
        for seg in ap.segments:
            m = seg.N
            ssl = slice(k, k + m)
            seg_names.append((ap.name, seg.name))
            seg_slices.append(ssl)
            itab[L.I_SEGMENT, ssl] = seg_counter
            itab[L.I_REID, ssl] = 1 if seg.reid_corr else 0
            tab[L.F_SIGMA, ssl] = (2.0 / (seg.b * seg.blend_dist * np.cos(seg.sweep_cp))) ** 2  # airplane.py:513
            c_node = np.asarray(seg.c_node)
            tab[L.F_CDJ0, ssl] = c_node[:-1] * seg.delta_joint
            tab[L.F_CDJ1, ssl] = c_node[1:] * seg.delta_joint
            # section model
            delta = np.asarray(seg._delta_flap, dtype=float)
            c_f = np.asarray(seg._cp_c_f, dtype=float)
            da, dcm, dcd = flap_terms(delta, c_f)
            tab[L.F_FLAP_DA, ssl] = da
            tab[L.F_FLAP_DCM, ssl] = dcm
            tab[L.F_FLAP_DCD, ssl] = dcd
            tab[L.F_FLAP_DEFL, ssl] = delta
            tab[L.F_FLAP_CF, ssl] = c_f
            if seg._num_airfoils == 1:
                a0 = af_id(seg._airfoils[0])
                itab[L.I_AFA, ssl] = a0
                itab[L.I_AFB, ssl] = a0
                tab[L.F_WB, ssl] = 0.0
            else:  # wing_segment.py:1062-1072
                spans = np.asarray(seg._airfoil_spans, dtype=float)
                cps = np.asarray(seg.cp_span_locs, dtype=float)
                j = np.searchsorted(spans, cps) - 1
                j = np.where(j < 0, 0, j)
                d = (cps - spans[j]) / (spans[j + 1] - spans[j])
                ids = np.array([af_id(a) for a in seg._airfoils], dtype=np.int32)
                itab[L.I_AFA, ssl] = ids[j]
                itab[L.I_AFB, ssl] = ids[j + 1]
                tab[L.F_WB, ssl] = d
            k += m
            seg_counter += 1
        base += na

    # atmosphere at the control points (scene.py:544-547): evaluated at earth positions
    tab[L.F_RHO, :n] = np.broadcast_to(np.asarray(scene._rho, dtype=float), (n,))
    tab[L.F_NU, :n] = np.broadcast_to(np.asarray(scene._nu, dtype=float), (n,))
    tab[L.F_SOS, :n] = np.broadcast_to(np.asarray(scene._a, dtype=float), (n,))
    tab[L.F_VWIND:L.F_VWIND + 3, :n] = np.asarray(scene._v_wind, dtype=float).reshape(n, 3).T

    # airfoil buffer: one parameter block per airfoil, then the records of the poly_fit / database airfoils (offsets patched into their
    # blocks), padded to a whole number of blocks so the buffer still reshapes to (-1, AF_STRIDE)
    blocks = np.asarray(airfoil_blocks, dtype=np.float64).reshape(-1, L.AF_STRIDE)
    pool, offset = [], blocks.size
    for a, rec in enumerate(airfoil_pools):
        if rec is not None:
            blocks[a, L.AF_POLY_OFF] = float(offset)
            pool.append(rec)
            offset += rec.size
    flat = np.concatenate([blocks.ravel()] + pool) if pool else blocks.ravel()
    flat = np.concatenate((flat, np.zeros((-flat.size) % L.AF_STRIDE)))
    flags = solver_flags(scene)
    if (blocks[:, L.AF_TYPE] > 1.5).any():
        flags |= L.FLAG_DATABASE_AIRFOILS
    return {
        "tab": tab, "itab": itab, "airfoils": flat.reshape(-1, L.AF_STRIDE), "n_airfoils": blocks.shape[0],
        "n": n, "ld": ld, "n_aircraft": len(airplanes), "n_segments": seg_counter,
        "segment_names": seg_names, "segment_slices": seg_slices, "aircraft_slices": ac_slices,
        "flags": flags,
    }


def build_aircraft_state(scene):
    """(n_aircraft, AC_STRIDE) block: what changes with the flight state (scene.py:562-582)."""
    airplanes = list(scene._airplane_objects)
    st = np.zeros((len(airplanes), L.AC_STRIDE))
    for ia, ap in enumerate(airplanes):
        q = np.asarray(ap.q, dtype=float)
        st[ia, L.AC_VTRANS:L.AC_VTRANS + 3] = -np.asarray(ap.v, dtype=float)
        st[ia, L.AC_W:L.AC_W + 3] = body_to_earth(q, np.asarray(ap.w, dtype=float))
        st[ia, L.AC_POS:L.AC_POS + 3] = np.asarray(ap.p_bar, dtype=float)
        st[ia, L.AC_CGR:L.AC_CGR + 3] = body_to_earth(q, np.asarray(ap.CG, dtype=float))
        st[ia, L.AC_ZF:L.AC_ZF + 3] = body_to_earth(q, np.array([0.0, 0.0, 1.0]))      # scene.py:601
    return st
