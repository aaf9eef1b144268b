"""CPU oracle: numpy restatement of MachUpX's lifting-line solve path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import
this module; the product (machupx_b200) never does and has no CPU fallback.

It consumes the same O(N) tables as liblls (machupx_b200/tables.py, include/lls.h) and
restates, function by function, what the reference computes (file:line under the reference):

  effective_rows()   Airplane._calculate_geometry, effective-LAC/joint loop   airplane.py:569-689
                     + cross-aircraft blocks                                    scene.py:469-515
  flow_state()       Scene._calc_invariant_flow_properties (O(N) part)          scene.py:557-611, 636-666
  influence_rows()   bound + joint segments, trailing legs, 1/4pi               scene.py:522-541, 621-634
  solve_linear()     Scene._solve_linear                                        scene.py:800-829
  residual()         _lifting_line_residual/_calc_v_i/_get_section_lift         scene.py:705-797
  jacobian()         Jacobian block of _solve_nonlinear                         scene.py:862-893
  solve_nonlinear()  Newton loop incl. exit/iteration bookkeeping               scene.py:851-915
  integrate()        _integrate_forces_and_moments (numeric part)               scene.py:941-1046, 1108-1117
  section_*()        airfoil_db linear model (third party, airfoil_db>=1.4.3; see oracle/ref_stubs)

Unlike the reference it is row-blocked (never holds more than a block of the per-pair
geometry), so it also runs at sizes where the reference exhausts memory.

PINNING: oracle/gen_golden.py runs the unmodified reference (with oracle/ref_stubs standing
in for airfoil_db/matplotlib/numpy-stl) and stores its V_ji, gamma, residual history and
force totals; tests/test_oracle_golden.py checks this module against those vectors and
against the reference's own hard-coded test goldens.
"""
import numpy as np

from machupx_b200 import layout as L

FOUR_PI_INV = 1.0 / (4.0 * np.pi)


class Tables:
    """Thin accessor over the table dict produced by machupx_b200.tables.build_tables."""

    def __init__(self, tables, ac_state):
        self.n = int(tables["n"])
        n = self.n
        tab = np.asarray(tables["tab"])
        itab = np.asarray(tables["itab"])
        self.flags = int(tables["flags"])
        self.af = np.asarray(tables["airfoils"]).reshape(-1, L.AF_STRIDE)
        self.ac = L.pad_ac_state(ac_state)
        v3 = lambda f: np.ascontiguousarray(tab[f:f + 3, :n].T)
        s1 = lambda f: np.ascontiguousarray(tab[f, :n])
        self.PC, self.P0, self.P1 = v3(L.F_PC), v3(L.F_P0), v3(L.F_P1)
        self.P0J, self.P1J = v3(L.F_P0J), v3(L.F_P1J)
        self.PCD, self.UAU, self.UNU = v3(L.F_PCD), v3(L.F_UAU), v3(L.F_UNU)
        self.UA, self.UN, self.US = v3(L.F_UA), v3(L.F_UN), v3(L.F_US)
        self.DL, self.RCG, self.VWIND = v3(L.F_DL), v3(L.F_RCG), v3(L.F_VWIND)
        self.SPC, self.SP0, self.SP1 = s1(L.F_SPC), s1(L.F_SP0), s1(L.F_SP1)
        self.SIGMA, self.CDJ0, self.CDJ1 = s1(L.F_SIGMA), s1(L.F_CDJ0), s1(L.F_CDJ1)
        self.CBAR, self.DS, self.CSWI = s1(L.F_CBAR), s1(L.F_DS), s1(L.F_CSWI)
        self.RHO, self.NU, self.SOS = s1(L.F_RHO), s1(L.F_NU), s1(L.F_SOS)
        self.WB = s1(L.F_WB)
        self.FLAP_DA, self.FLAP_DCM, self.FLAP_DCD = s1(L.F_FLAP_DA), s1(L.F_FLAP_DCM), s1(L.F_FLAP_DCD)
        self.FLAP_DEFL, self.FLAP_CF = s1(L.F_FLAP_DEFL), s1(L.F_FLAP_CF)
        self.wing_begin = itab[L.I_WING_BEGIN, :n].astype(np.int64)
        self.wing_end = itab[L.I_WING_END, :n].astype(np.int64)
        self.aircraft = itab[L.I_AIRCRAFT, :n].astype(np.int64)
        self.reid = itab[L.I_REID, :n].astype(bool)
        self.segment = itab[L.I_SEGMENT, :n].astype(np.int64)
        self.afa = itab[L.I_AFA, :n].astype(np.int64)
        self.afb = itab[L.I_AFB, :n].astype(np.int64)
        self.n_segments = int(tables["n_segments"])
        self.section_bounds = False     # set by the section model when a database airfoil is asked for a point outside its table
        self.swept = bool(self.flags & L.FLAG_SWEPT_SECTIONS)
        self.total_velocity = bool(self.flags & L.FLAG_TOTAL_VELOCITY)
        self.in_plane = bool(self.flags & L.FLAG_IN_PLANE)
        self.match_pro = bool(self.flags & L.FLAG_MATCH_MACHUP_PRO)
        self.constrain_sheet = bool(self.flags & L.FLAG_CONSTRAIN_VORTEX_SHEET)
        # earth-fixed absolute positions, as the reference forms them (scene.py:445-447, 472-475)
        pos = self.ac[self.aircraft, L.AC_POS:L.AC_POS + 3]
        self.PC_abs = pos + self.PC
        self.P0_abs, self.P1_abs = pos + self.P0, pos + self.P1
        self.P0J_abs, self.P1J_abs = pos + self.P0J, pos + self.P1J


# ------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------
def _gradient_rows(f, x):
    """numpy.gradient(f, x, edge_order=2, axis=0) applied independently to each leading row.

    f: (R, M, 3), x: (R, M).  Same three-point non-uniform formulas numpy uses
    (interior second order; one-sided second order at both ends), airplane.py:628, 640."""
    dx = np.diff(x, axis=1)
    out = np.empty_like(f)
    hs, hd = dx[:, :-1], dx[:, 1:]
    a = -hd / (hs * (hd + hs))
    b = (hd - hs) / (hd * hs)
    c = hs / (hd * (hd + hs))
    out[:, 1:-1] = a[..., None] * f[:, :-2] + b[..., None] * f[:, 1:-1] + c[..., None] * f[:, 2:]
    h0, h1 = dx[:, 0], dx[:, 1]
    a = -(2.0 * h0 + h1) / (h0 * (h0 + h1))
    b = (h0 + h1) / (h0 * h1)
    c = -h0 / (h1 * (h0 + h1))
    out[:, 0] = a[:, None] * f[:, 0] + b[:, None] * f[:, 1] + c[:, None] * f[:, 2]
    hs, hd = dx[:, -2], dx[:, -1]
    a = hd / (hs * (hd + hs))
    b = -(hd + hs) / (hd * hs)
    c = (2.0 * hd + hs) / (hd * (hd + hs))
    out[:, -1] = a[:, None] * f[:, -3] + b[:, None] * f[:, -2] + c[:, None] * f[:, -1]
    return out


def _joint_points(Pe, u_a, cdj):
    """Jointed-vortex corner from the blended line Pe (R, M, 3): airplane.py:623-635 / 637-647."""
    d = np.diff(Pe, axis=1)
    dt = np.zeros(Pe.shape[:2])
    dt[:, 1:] = np.cumsum(np.linalg.norm(d, axis=2), axis=1)
    T = _gradient_rows(Pe, dt)
    T = T / np.linalg.norm(T, axis=2)[..., None]
    k = np.einsum('rmk,rmk->rm', T, u_a)
    c1 = np.sqrt(1.0 / (1.0 - k * k))
    c2 = -c1 * k
    u_j = c1[..., None] * u_a + c2[..., None] * T
    u_j = u_j / np.linalg.norm(u_j, axis=-1, keepdims=True)
    return Pe + cdj[None, :, None] * u_j


def effective_rows(T, i0, i1):
    """r_0, r_1, r_0_joint, r_1_joint for control points i0..i1 against all vortices: (R, N, 3) each."""
    n = T.n
    R = i1 - i0
    rows = np.arange(i0, i1)
    # default: other aircraft -> plain nodes and plain joints in absolute earth coordinates (scene.py:469-475, 501-504)
    r0 = T.PC_abs[rows, None, :] - T.P0_abs[None, :, :]
    r1 = T.PC_abs[rows, None, :] - T.P1_abs[None, :, :]
    r0j = T.PC_abs[rows, None, :] - T.P0J_abs[None, :, :]
    r1j = T.PC_abs[rows, None, :] - T.P1J_abs[None, :, :]
    same_ac = T.aircraft[rows, None] == T.aircraft[None, :]
    # same aircraft: body-relative differences (airplane.py:672-675), plain nodes + plain joints (airplane.py:530-531, 557-558)
    r0_b = T.PC[rows, None, :] - T.P0[None, :, :]
    r1_b = T.PC[rows, None, :] - T.P1[None, :, :]
    r0j_b = T.PC[rows, None, :] - T.P0J[None, :, :]
    r1j_b = T.PC[rows, None, :] - T.P1J[None, :, :]
    # rows without Reid corrections: joints collapse onto the nodes for every vortex of the aircraft (airplane.py:665-669)
    noreid = ~T.reid[rows]
    r0j_b[noreid] = r0_b[noreid]
    r1j_b[noreid] = r1_b[noreid]
    m3 = same_ac[..., None]
    r0 = np.where(m3, r0_b, r0)
    r1 = np.where(m3, r1_b, r1)
    r0j = np.where(m3, r0j_b, r0j)
    r1j = np.where(m3, r1j_b, r1j)

    # Reid rows: blended nodes and jointed corners for the vortices of the same wing (airplane.py:580-647)
    for wb in np.unique(T.wing_begin[rows]):
        sel = np.nonzero((T.wing_begin[rows] == wb) & T.reid[rows])[0]
        if sel.size == 0:
            continue
        ii = rows[sel]
        we = int(T.wing_end[ii[0]])
        ws = slice(int(wb), we)
        PC = T.PC[ii]                        # (R', 3)
        span = T.SPC[ii]
        deriv = T.PCD[ii]
        sig = T.SIGMA[ii]
        # blend P0 (airplane.py:587-590)
        ds0 = T.SP0[ws][None, :] - span[:, None]
        straight = PC[:, None, :] + deriv[:, None, :] * ds0[..., None]
        blend0 = np.exp(-sig[:, None] * ds0 * ds0)[..., None]
        P0e = straight * blend0 + T.P0[ws][None, :, :] * (1.0 - blend0)
        # blend P1 (airplane.py:593-596)
        ds1 = T.SP1[ws][None, :] - span[:, None]
        straight = PC[:, None, :] + deriv[:, None, :] * ds1[..., None]
        blend1 = np.exp(-sig[:, None] * ds1 * ds1)[..., None]
        P1e = straight * blend1 + T.P1[ws][None, :, :] * (1.0 - blend1)
        # blend u_a_unswept (airplane.py:599-603)
        ds = T.SPC[ws][None, :] - span[:, None]
        blend = np.exp(-sig[:, None] * ds * ds)[..., None]
        u_a = T.UAU[ii][:, None, :] * blend + T.UAU[ws][None, :, :] * (1.0 - blend)
        u_a = u_a / np.linalg.norm(u_a, axis=2, keepdims=True)
        P0je = _joint_points(P0e, u_a, T.CDJ0[ws])
        P1je = _joint_points(P1e, u_a, T.CDJ1[ws])
        r0[sel, ws] = PC[:, None, :] - P0e
        r1[sel, ws] = PC[:, None, :] - P1e
        r0j[sel, ws] = PC[:, None, :] - P0je
        r1j[sel, ws] = PC[:, None, :] - P1je
    return r0, r1, r0j, r1j


# ------------------------------------------------------------------------------------------
# section model (airfoil_db linear type; two bracketing airfoils blended in span, wing_segment.py:1062-1124)
# ------------------------------------------------------------------------------------------
def _af(T, ids, field):
    return T.af[ids, field]


def _blend(T, fa, fb):
    return (1.0 - T.WB) * fa + T.WB * fb


# poly_fit airfoils (airfoil_db "poly_fit" type; third party, semantics defined in machupx_b200/airfoil.py and
# include/lls.h; PARITY UNPINNED by the reference).  Plain restatement: decompose the coefficient index into the exponent
# tuple (last variable fastest, reference machupX/poly_fits.py:433-480) and sum the monomials.
def _poly_record(T, a, which):
    flat = T.af.ravel()
    blk = T.af[a]
    V = int(blk[L.AF_POLY_NDOF])
    kinds = [int(blk[L.AF_POLY_DOF0 + v]) for v in range(V)]
    off = int(blk[L.AF_POLY_OFF])
    limits = flat[off:off + 2 * V].reshape(V, 2)
    pos = off + 2 * V
    for _ in range(which):
        pos += 1 + V + int(flat[pos])
    ncoef = int(flat[pos])
    degrees = [int(x) for x in flat[pos + 1:pos + 1 + V]]
    coefs = flat[pos + 1 + V:pos + 1 + V + ncoef]
    return kinds, limits, degrees, coefs


def _poly_one(T, a, mask, alpha, Re, M, which, dkind):
    kinds, limits, degrees, coefs = _poly_record(T, a, which)
    xs = []
    for k in kinds:
        xs.append({0: alpha, 1: Re, 2: M, 3: T.FLAP_DEFL, 4: T.FLAP_CF}[k][mask])
    if dkind is not None and dkind not in kinds:
        return np.zeros(mask.sum())
    out = np.zeros(mask.sum())
    for j, c in enumerate(coefs):
        rest, exps = j, []
        for d in reversed(degrees):
            exps.append(rest % (d + 1))
            rest //= (d + 1)
        exps = exps[::-1]
        term = np.full(mask.sum(), c)
        for k, x, e in zip(kinds, xs, exps):
            if dkind is not None and k == dkind:
                term = term * (e * x ** (e - 1) if e > 0 else 0.0)
            else:
                term = term * x ** e
        out += term
    return out


# database airfoils (airfoil_db "database" type; third party).  Restated from its behaviour: barycentric interpolation on the
# Delaunay triangulation of the data points in variables scaled to [0, 1] (scipy griddata, method "linear"), CLa / CLRe / CLM as
# central differences of that interpolant with steps 0.001 / 1000 / 0.05, aL0 as its root in alpha.  PINNED by the reference's
# golden of test/scene_tests/test_solve_forces.py:453-494 (tests/test_database_airfoils.py).  The triangulation itself is an
# INPUT here (the pool record built by machupx_b200/airfoil.py with scipy, layout in include/lls.h).
DB_EPS = 1e-12
DB_STEP = {0: 0.001, 1: 1000.0, 2: 0.05}


def _db_record(T, a):
    flat = T.af.ravel()
    blk = T.af[a]
    V = int(blk[L.AF_POLY_NDOF])
    kinds = [int(blk[L.AF_POLY_DOF0 + v]) for v in range(V)]
    off = int(blk[L.AF_POLY_OFF])
    P, S = int(flat[off]), int(flat[off + 1])
    pos = off + 2
    lim = flat[pos:pos + 2 * V].reshape(V, 2)
    pos += 2 * V
    vals = flat[pos:pos + 3 * P].reshape(P, 3)
    pos += 3 * P
    w = V + 1 + V * V + V
    simp = flat[pos:pos + S * w].reshape(S, w)
    return kinds, lim, vals, simp[:, :V + 1].astype(np.int64), simp[:, V + 1:V + 1 + V * V].reshape(S, V, V), simp[:, V + 1 + V * V:]


def _db_eval(rec, X, which, found=None):
    """Interpolated coefficient `which` at the raw points X (n, V); points outside the table extrapolate from the simplex whose
    smallest barycentric coordinate is largest (and are reported through `found`, a list that receives one bool per point)."""
    kinds, lim, vals, verts, Tinv, r = rec
    xn = (X - lim[:, 0]) / lim[:, 1]
    out = np.zeros(len(xn))
    for q in range(len(xn)):
        c = np.einsum("sij,sj->si", Tinv, xn[q] - r)
        c = np.concatenate((c, 1.0 - c.sum(axis=1, keepdims=True)), axis=1)
        cmin = np.where(np.isnan(c).any(axis=1), -np.inf, c.min(axis=1))
        inside = np.nonzero(cmin >= -DB_EPS)[0]
        s = inside[0] if inside.size else int(np.argmax(cmin))
        if found is not None:
            found.append(bool(inside.size))
        out[q] = np.dot(c[s], vals[verts[s], which])
    return out


def _db_points(T, rec, mask, alpha, Re, M):
    cols = {0: alpha, 1: Re, 2: M, 3: T.FLAP_DEFL, 4: T.FLAP_CF}
    return np.stack([np.broadcast_to(cols[k], (T.n,))[mask] for k in rec[0]], axis=1).astype(float)


def _db_one(T, a, mask, alpha, Re, M, which, dkind):
    rec = _db_record(T, a)
    X = _db_points(T, rec, mask, alpha, Re, M)
    if dkind is None:           # CL, CD, Cm themselves: a point outside the table is what LLS_STATUS_SECTION_BOUNDS reports
        found = []
        out = _db_eval(rec, X, which, found)
        if not all(found):
            T.section_bounds = True
        return out
    if dkind not in rec[0]:
        return np.zeros(mask.sum())
    v = rec[0].index(dkind)
    lo, hi = rec[1][v, 0], rec[1][v, 0] + rec[1][v, 1]
    x = X[:, v]
    up = np.minimum(x + DB_STEP[dkind], np.maximum(hi, x)) - x           # samples stay inside the table's limits
    dn = np.maximum(x - DB_STEP[dkind], np.minimum(lo, x)) - x
    Xu, Xd = X.copy(), X.copy()
    Xu[:, v] += up
    Xd[:, v] += dn
    return (_db_eval(rec, Xu, which) - _db_eval(rec, Xd, which)) / (up - dn)


def _db_aL0(T, a, mask, Re, M):
    rec = _db_record(T, a)
    m = int(mask.sum())
    if 0 not in rec[0]:
        return np.zeros(m)
    def CL(al):
        alpha = np.zeros(T.n)
        alpha[mask] = al
        return _db_eval(rec, _db_points(T, rec, mask, alpha, Re, M), 0)
    a0, a1 = np.zeros(m), np.full(m, 0.01)
    f0, f1 = CL(a0), CL(a1)
    for _ in range(50):                                                   # secant, exact on one linear piece
        active = np.abs(a1 - a0) > 1e-10
        if not active.any():
            break
        with np.errstate(divide="ignore", invalid="ignore"):
            a2 = np.where(active, a1 - f1 * (a0 - a1) / (f0 - f1), a1)
        f2 = CL(a2)
        a0, f0, a1, f1 = a1, f1, a2, f2
    return a1


def _per_airfoil(T, ids, alpha, Re, M, linear_fn, which, dkind=None):
    """Evaluates one bracketing airfoil per control point: linear_fn(ids) for linear airfoils, the polynomial or the table
    otherwise."""
    out = linear_fn(ids)
    for a in np.unique(ids):
        kind = T.af[a, L.AF_TYPE]
        if kind > 0.5:
            mask = ids == a
            out = np.array(out, dtype=float)
            one = _db_one if kind > 1.5 else _poly_one
            out[mask] = one(T, a, mask, alpha, Re, M, which, dkind)
    return out


def _zeros(T):
    return np.zeros(T.n)


def section_CL(T, alpha, Re=None, M=None):
    Re = _zeros(T) if Re is None else Re
    M = _zeros(T) if M is None else M

    def lin(ids):
        cl = _af(T, ids, L.AF_CLA) * (alpha - _af(T, ids, L.AF_AL0) + T.FLAP_DA)
        m = _af(T, ids, L.AF_CLMAX)
        return np.clip(cl, -m, m)
    return _blend(T, _per_airfoil(T, T.afa, alpha, Re, M, lin, 0), _per_airfoil(T, T.afb, alpha, Re, M, lin, 0))


def _section_CL_d(T, alpha, Re, M, dkind):
    lin = (lambda ids: _af(T, ids, L.AF_CLA)) if dkind == 0 else (lambda ids: np.zeros(T.n))
    return _blend(T, _per_airfoil(T, T.afa, alpha, Re, M, lin, 0, dkind), _per_airfoil(T, T.afb, alpha, Re, M, lin, 0, dkind))


def section_CLa(T, alpha=None, Re=None, M=None):
    z = _zeros(T)
    return _section_CL_d(T, z if alpha is None else alpha, z if Re is None else Re, z if M is None else M, 0)


def section_CLRe(T, alpha, Re, M):
    return _section_CL_d(T, alpha, Re, M, 1)


def section_CLM(T, alpha, Re, M):
    return _section_CL_d(T, alpha, Re, M, 2)


def section_aL0(T, Re=None, M=None):
    Re = _zeros(T) if Re is None else Re
    M = _zeros(T) if M is None else M

    def one(ids):
        out = _af(T, ids, L.AF_AL0) - T.FLAP_DA
        for a in np.unique(ids):
            if T.af[a, L.AF_TYPE] > 1.5:            # database airfoil: root of the table in alpha
                mask = ids == a
                out = np.array(out, dtype=float)
                out[mask] = _db_aL0(T, a, mask, Re, M)
            elif T.af[a, L.AF_TYPE] > 0.5:          # root of the CL fit in alpha by Newton's method from alpha = 0
                mask = ids == a
                a0 = np.zeros(T.n)
                for _ in range(30):
                    f = _poly_one(T, a, mask, a0, Re, M, 0, None)
                    df = _poly_one(T, a, mask, a0, Re, M, 0, 0)
                    step = f / df
                    a0[mask] -= step
                    if not (np.abs(step) > 1e-15).any():
                        break
                out = np.array(out, dtype=float)
                out[mask] = a0[mask]
        return out
    return _blend(T, one(T.afa), one(T.afb))


def section_CD(T, alpha, Re=None, M=None):
    Re = _zeros(T) if Re is None else Re
    M = _zeros(T) if M is None else M

    def lin(ids):
        m = _af(T, ids, L.AF_CLMAX)
        cl = np.clip(_af(T, ids, L.AF_CLA) * (alpha - _af(T, ids, L.AF_AL0)), -m, m)
        return _af(T, ids, L.AF_CD0) + _af(T, ids, L.AF_CD1) * cl + _af(T, ids, L.AF_CD2) * cl * cl + T.FLAP_DCD
    return _blend(T, _per_airfoil(T, T.afa, alpha, Re, M, lin, 1), _per_airfoil(T, T.afb, alpha, Re, M, lin, 1))


def section_Cm(T, alpha, Re=None, M=None):
    Re = _zeros(T) if Re is None else Re
    M = _zeros(T) if M is None else M

    def lin(ids):
        return _af(T, ids, L.AF_CMA) * (alpha - _af(T, ids, L.AF_AL0)) + _af(T, ids, L.AF_CML0) + T.FLAP_DCM
    return _blend(T, _per_airfoil(T, T.afa, alpha, Re, M, lin, 2), _per_airfoil(T, T.afb, alpha, Re, M, lin, 2))


# ------------------------------------------------------------------------------------------
# flow state (scene.py:557-611, 636-666)
# ------------------------------------------------------------------------------------------
class Flow:
    pass


def flow_state(T):
    F = Flow()
    ac = T.ac[T.aircraft]
    v_trans = ac[:, L.AC_VTRANS:L.AC_VTRANS + 3]
    w = ac[:, L.AC_W:L.AC_W + 3]
    cgr = ac[:, L.AC_CGR:L.AC_CGR + 3]
    F.v_inf = v_trans + T.VWIND                                               # scene.py:571
    F.v_inf_and_rot = F.v_inf - np.cross(w, T.RCG)                            # scene.py:569, 572
    if T.match_pro:                                                           # scene.py:575-577
        j0 = F.v_inf.copy()
        j1 = F.v_inf.copy()
    else:                                                                     # scene.py:579-582
        j0 = F.v_inf - np.cross(w, T.P0J - cgr)
        j1 = F.v_inf - np.cross(w, T.P1J - cgr)
    F.V_inf = np.linalg.norm(F.v_inf, axis=1)
    F.u_inf = F.v_inf / F.V_inf[:, None]
    F.V_inf_and_rot = np.linalg.norm(F.v_inf_and_rot, axis=1)
    F.u_trailing_0 = j0 / np.linalg.norm(j0, axis=-1, keepdims=True)          # scene.py:591
    F.u_trailing_1 = j1 / np.linalg.norm(j1, axis=-1, keepdims=True)          # scene.py:592
    if T.constrain_sheet:                                                     # scene.py:595-611
        zf = ac[:, L.AC_ZF:L.AC_ZF + 3]                                       # body z axis of the owning aircraft, earth axes
        for name in ("u_trailing_0", "u_trailing_1"):
            u = getattr(F, name)
            u = u - zf * np.einsum('ij,ij->i', zf, u)[:, None]                # (I - z z^T) u
            setattr(F, name, u / np.linalg.norm(u, axis=-1, keepdims=True))
    if T.in_plane:                                                            # scene.py:637-643
        dot = np.einsum('ij,ij->i', T.US, F.v_inf)
        F.v_inf_in_plane = F.v_inf - T.US * dot[:, None]
        F.V_inf_in_plane = np.linalg.norm(F.v_inf_in_plane, axis=1)
        dot = np.einsum('ij,ij->i', T.US, F.v_inf_and_rot)
        v_ir_ip = F.v_inf_and_rot - T.US * dot[:, None]
        F.V_inf_and_rot_in_plane = np.linalg.norm(v_ir_ip, axis=1)
        F.Re = F.V_inf_and_rot_in_plane * T.CBAR / T.NU
        F.M = F.V_inf_in_plane / T.SOS
    else:                                                                     # scene.py:645-646
        F.Re = F.V_inf * T.CBAR / T.NU
        F.M = F.V_inf / T.SOS
    v_n = np.einsum('ij,ij->i', F.v_inf_and_rot, T.UN)                        # scene.py:649-651
    v_a = np.einsum('ij,ij->i', F.v_inf_and_rot, T.UA)
    F.alpha_inf = np.arctan2(v_n, v_a)
    F.CLa = section_CLa(T, F.alpha_inf, F.Re, F.M)                            # scene.py:659-661
    F.CL = section_CL(T, F.alpha_inf, F.Re, F.M)
    F.aL0 = section_aL0(T, F.Re, F.M)
    if T.swept:                                                               # scene.py:665-666, 797
        F.CL = F.CL + F.CLa * (F.aL0 - F.aL0 * T.CSWI)
    return F


# ------------------------------------------------------------------------------------------
# influence matrix (scene.py:522-541, 621-634)
# ------------------------------------------------------------------------------------------
def influence_rows(T, F, i0, i1, want_denoms=False):
    """V_ji[i0:i1, :, :] (R, N, 3)."""
    r0, r1, r0j, r1j = effective_rows(T, i0, i1)
    rows = np.arange(i0, i1)
    with np.errstate(divide='ignore', invalid='ignore'):
        r0m = np.linalg.norm(r0, axis=-1)
        r1m = np.linalg.norm(r1, axis=-1)
        r0jm = np.linalg.norm(r0j, axis=-1)
        r1jm = np.linalg.norm(r1j, axis=-1)
        r0r1 = r0m * r1m
        r0r0j = r0m * r0jm
        r1r1j = r1m * r1jm
        # bound (scene.py:525-528)
        numer = (r0m + r1m)[..., None] * np.cross(r0, r1)
        denom = r0r1 * (r0r1 + np.einsum('ijk,ijk->ij', r0, r1))
        Vb = numer / denom[..., None]
        Vb[np.arange(i1 - i0), rows] = 0.0
        # joints (scene.py:531-538)
        numer = (r0jm + r0m)[..., None] * np.cross(r0j, r0)
        denom = r0r0j * (r0r0j + np.einsum('ijk,ijk->ij', r0j, r0))
        Vj0 = numer / denom[..., None]
        numer = (r1jm + r1m)[..., None] * np.cross(r1, r1j)
        denom = r1r1j * (r1r1j + np.einsum('ijk,ijk->ij', r1, r1j))
        Vj1 = numer / denom[..., None]
        # trailing legs (scene.py:621-630)
        d0 = r0jm * (r0jm - np.einsum('ijk,ijk->ij', F.u_trailing_0[None], r0j))
        V0 = np.where(d0[..., None] > 1e-13, np.nan_to_num(-np.cross(F.u_trailing_0[None], r0j) / d0[..., None]), 0.0)
        d1 = r1jm * (r1jm - np.einsum('ijk,ijk->ij', F.u_trailing_1[None], r1j))
        V1 = np.where(d1[..., None] > 1e-13, np.nan_to_num(np.cross(F.u_trailing_1[None], r1j) / d1[..., None]), 0.0)
        V = FOUR_PI_INV * (V0 + (Vb + Vj0 + Vj1) + V1)                         # scene.py:541, 634
    if want_denoms == "conditioning":
        # |dV/V| amplification of the two trailing legs: r (r - u.r) cancels when the control point lies
        # almost exactly downstream of a joint.  Used by the parity tests to scale their tolerance.
        with np.errstate(divide='ignore', invalid='ignore'):
            k0 = np.where(d0 > 1e-13, r0jm * r0jm / d0, 0.0)
            k1 = np.where(d1 > 1e-13, r1jm * r1jm / d1, 0.0)
        amp = FOUR_PI_INV * (np.abs(V0).max(axis=-1) * k0 + np.abs(V1).max(axis=-1) * k1)
        return V, amp
    if want_denoms:
        return V, np.minimum(np.abs(d0).min(), np.abs(d1).min())
    return V


def influence_with_conditioning(T, F, block=256):
    """(V, amp): amp[i, j] bounds how strongly one ulp of input perturbation shows in |V[i, j, :]|."""
    n = T.n
    V = np.empty((n, n, 3))
    amp = np.empty((n, n))
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        V[i0:i1], amp[i0:i1] = influence_rows(T, F, i0, i1, want_denoms="conditioning")
    return V, amp


def influence(T, F, block=256, impingement_threshold=None):
    """Full V_ji (N, N, 3).  Returns (V, impinged) when a threshold is given (scene.py:622-623)."""
    n = T.n
    V = np.empty((n, n, 3))
    dmin = np.inf
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        V[i0:i1], d = influence_rows(T, F, i0, i1, want_denoms=True)
        dmin = min(dmin, d)
    if impingement_threshold is None:
        return V
    return V, bool(dmin < impingement_threshold)


# ------------------------------------------------------------------------------------------
# linear solve (scene.py:800-829)
# ------------------------------------------------------------------------------------------
def linear_system(T, F, V):
    vxdl = np.cross(F.v_inf_and_rot, T.DL)                                    # scene.py:811
    V_dot_un = np.einsum('ijk,ik->ij', V, T.UN)                               # scene.py:815
    if T.in_plane:                                                            # scene.py:817-818
        A = -(F.V_inf_in_plane * F.CLa * T.DS)[:, None] * V_dot_un
        b = F.V_inf_in_plane * F.V_inf_in_plane * T.DS * F.CL
    else:                                                                     # scene.py:820-821
        A = -(F.V_inf * F.CLa * T.DS)[:, None] * V_dot_un
        b = F.V_inf * F.V_inf * T.DS * F.CL
    A[np.diag_indices(T.n)] += 2.0 * np.linalg.norm(vxdl, axis=1)             # scene.py:824
    return A, b


def solve_linear(T, F, V):
    A, b = linear_system(T, F, V)
    return np.linalg.solve(A, b)                                              # scene.py:827


# ------------------------------------------------------------------------------------------
# nonlinear solve (scene.py:705-797, 832-917)
# ------------------------------------------------------------------------------------------
class Local:
    pass


def residual(T, F, V, gamma):
    S = Local()
    S.v_i = F.v_inf_and_rot + np.einsum('ijk,j->ik', V, gamma)                # scene.py:728
    S.w_i = np.cross(S.v_i, T.DL)                                             # scene.py:715-717
    S.w_i_mag = np.linalg.norm(S.w_i, axis=1)
    L_vortex = 2.0 * S.w_i_mag * gamma
    if T.in_plane:                                                            # scene.py:737-742
        S.v_i_in_plane = S.v_i - T.US * np.einsum('ij,ij->i', T.US, S.v_i)[:, None]
        S.V_i_in_plane_2 = np.einsum('ij,ij->i', S.v_i_in_plane, S.v_i_in_plane)
        S.V_i_in_plane = np.sqrt(S.V_i_in_plane_2)
        S.Re = S.V_i_in_plane * T.CBAR / T.NU
        S.M = S.V_i_in_plane / T.SOS
    else:                                                                     # scene.py:744-748
        S.V_i_2 = np.einsum('ij,ij->i', S.v_i, S.v_i)
        S.V_i = np.sqrt(S.V_i_2)
        S.Re = S.V_i * T.CBAR / T.NU
        S.M = S.V_i / T.SOS
    S.v_a = np.einsum('ij,ij->i', S.v_i, T.UA)                                # scene.py:751-753
    S.v_n = np.einsum('ij,ij->i', S.v_i, T.UN)
    S.alpha = np.arctan2(S.v_n, S.v_a)
    S.CLa = section_CLa(T, S.alpha, S.Re, S.M)                                # scene.py:765-768
    S.CL = section_CL(T, S.alpha, S.Re, S.M)
    S.CLRe = section_CLRe(T, S.alpha, S.Re, S.M)
    S.CLM = section_CLM(T, S.alpha, S.Re, S.M)
    if T.match_pro:                                                           # scene.py:774-775
        L_section = F.V_inf * F.V_inf * S.CL * T.DS
    else:
        if T.swept:                                                           # scene.py:778-779
            S.CL = S.CL + S.CLa * (F.aL0 - F.aL0 * T.CSWI)
        if T.total_velocity:                                                  # scene.py:782-791
            L_section = (S.V_i_in_plane_2 if T.in_plane else S.V_i_2) * S.CL * T.DS
        else:
            Vref = F.V_inf_in_plane if T.in_plane else F.V_inf
            L_section = Vref * Vref * S.CL * T.DS
    S.R = L_vortex - L_section                                                # scene.py:723
    return S


def jacobian(T, F, V, gamma, S):
    if T.in_plane:                                                            # scene.py:846-849
        # NB np.matmul(self._P_in_plane (N,3,3), V_ji[..., None] (N,N,3,1)) broadcasts the projector over the
        # VORTEX index j (scene.py:847): the reference projects V_ji[i,j] with P_j, not P_i.  Kept as is.
        Vp = V - T.US[None, :, :] * np.einsum('ijk,jk->ij', V, T.US)[..., None]
        v_iji = np.einsum('ik,ijk->ij', S.v_i_in_plane, Vp)                   # scene.py:862
        Vmag, V2 = S.V_i_in_plane, S.V_i_in_plane_2
    else:
        Vp = V
        v_iji = np.einsum('ik,ijk->ij', S.v_i, Vp)                            # scene.py:864
        Vmag, V2 = S.V_i, S.V_i_2
    # NB np.cross(V_ji, self._dl) in the reference broadcasts dl over the VORTEX index j (scene.py:867), not the
    # control point i the derivation calls for; the iteration counts of the reference depend on it, so it is kept.
    J = (2.0 * gamma / S.w_i_mag)[:, None] * np.einsum('ik,ijk->ij', S.w_i, np.cross(Vp, T.DL[None, :, :]))
    if T.total_velocity:
        J -= (2.0 * T.DS * S.CL)[:, None] * v_iji                             # scene.py:870
    # NB the reference broadcasts c_bar over the vortex index j here (scene.py:873); kept as is
    CL_gamma_Re = S.CLRe[:, None] * T.CBAR[None, :] / (T.NU * Vmag)[:, None] * v_iji
    CL_gamma_M = S.CLM[:, None] / (T.SOS * Vmag)[:, None] * v_iji            # scene.py:874
    CL_gamma_alpha = S.CLa[:, None] * (S.v_a[:, None] * np.einsum('ijk,ik->ij', Vp, T.UN)
                                       - S.v_n[:, None] * np.einsum('ijk,ik->ij', Vp, T.UA)) \
        / (S.v_n * S.v_n + S.v_a * S.v_a)[:, None]                            # scene.py:879
    if T.total_velocity:                                                      # scene.py:881-890
        scale = V2 * T.DS
    else:
        Vref = F.V_inf_in_plane if T.in_plane else F.V_inf
        scale = Vref * Vref * T.DS
    J -= scale[:, None] * (CL_gamma_alpha + CL_gamma_Re + CL_gamma_M)
    J[np.diag_indices(T.n)] += 2.0 * S.w_i_mag                                # scene.py:893
    return J


def solve_nonlinear(T, F, V, gamma0, convergence=1e-10, relaxation=1.0, max_iterations=100):
    """Returns (gamma, iterations, final_error, converged, error_history)."""
    gamma = np.array(gamma0, dtype=float)
    iteration = 0
    error = 100.0
    hist = []
    while error > convergence:                                                # scene.py:853
        iteration += 1
        S = residual(T, F, V, gamma)
        error = np.linalg.norm(S.R)                                           # scene.py:858
        hist.append(error)
        J = jacobian(T, F, V, gamma, S)
        dGamma = np.linalg.solve(J, -S.R)                                     # scene.py:896
        gamma = gamma + relaxation * dGamma                                   # scene.py:899
        if iteration >= max_iterations:                                       # scene.py:905-908
            S = residual(T, F, V, gamma)
            return gamma, iteration, float(np.linalg.norm(S.R)), False, hist
    S = residual(T, F, V, gamma)                                              # scene.py:911-913
    return gamma, iteration, float(np.linalg.norm(S.R)), True, hist


# ------------------------------------------------------------------------------------------
# force integration (scene.py:927-1117, numeric part)
# ------------------------------------------------------------------------------------------
def integrate(T, F, V, gamma):
    """Returns (seg_fm (n_segments, 12) = [F_inv, M_inv, F_visc, M_visc] earth axes, per-CP dict)."""
    gamma = np.array(gamma, dtype=float)
    if T.match_pro:                                                           # scene.py:938-939
        gamma = gamma * (F.V_inf_and_rot / F.V_inf) ** 2
    D = {}
    if T.total_velocity or T.match_pro:                                       # scene.py:942-949
        v_i = F.v_inf_and_rot + np.einsum('ijk,j->ik', V, gamma)
        V_i_2 = np.einsum('ij,ij->i', v_i, v_i)
        V_i = np.sqrt(V_i_2)
        u_i = v_i / V_i[:, None]
        if T.in_plane:
            v_ip = v_i - T.US * np.einsum('ij,ij->i', T.US, v_i)[:, None]
            V_ip_2 = np.einsum('ij,ij->i', v_ip, v_ip)
    else:
        raise NotImplementedError("integrate() needs v_i from a previous residual when use_total_velocity is off")
    dF_inv = (T.RHO * gamma)[:, None] * np.cross(v_i, T.DL)                   # scene.py:952
    v_a = np.einsum('ij,ij->i', v_i, T.UA)                                    # scene.py:955-957
    v_n = np.einsum('ij,ij->i', v_i, T.UN)
    alpha = np.arctan2(v_n, v_a)
    if T.swept:                                                               # scene.py:958-970
        Re_unswept = V_i * T.CBAR * T.CSWI / T.NU
        alpha_unswept = np.arctan2(np.einsum('ij,ij->i', v_i, T.UNU), np.einsum('ij,ij->i', v_i, T.UAU))
    else:
        Re_unswept = V_i * T.CBAR / T.NU
        alpha_unswept = alpha
    if T.in_plane:                                                            # scene.py:975-981
        V_ip = np.sqrt(V_ip_2)
        Re = V_ip * T.CBAR / T.NU
        M = V_ip / T.SOS
    else:
        Re = V_i * T.CBAR / T.NU
        M = V_i / T.SOS
    if T.total_velocity:                                                      # scene.py:984-991
        redim_full = 0.5 * T.RHO * V_i_2 * T.DS
        redim_ip = 0.5 * T.RHO * V_ip_2 * T.DS if T.in_plane else None
    else:
        redim_full = 0.5 * T.RHO * F.V_inf * F.V_inf * T.DS
        redim_ip = 0.5 * T.RHO * F.V_inf_in_plane * F.V_inf_in_plane * T.DS if T.in_plane else None
    CD = section_CD(T, alpha_unswept, Re_unswept, M)                          # scene.py:1014-1017
    Cm = section_Cm(T, alpha, Re, M)                                          # scene.py:1020
    if T.swept:
        Cm = Cm * T.CSWI                                                      # scene.py:1026
    if T.in_plane:                                                            # scene.py:1029-1032
        dM_section = (redim_ip * T.CBAR * Cm)[:, None] * T.US
    else:
        dM_section = (redim_full * T.CBAR * Cm)[:, None] * T.US
    dM_inv = np.cross(T.RCG, dF_inv) + dM_section                             # scene.py:1035-1036
    dD = redim_full * CD                                                      # scene.py:1039-1043
    dF_visc = dD[:, None] * (u_i if (T.total_velocity or T.match_pro) else F.u_inf)
    dM_visc = np.cross(T.RCG, dF_visc)                                        # scene.py:1046
    seg = np.zeros((T.n_segments, 12))
    for s in range(T.n_segments):                                             # scene.py:1108-1117 (sums only)
        m = T.segment == s
        seg[s, 0:3] = np.sum(dF_inv[m], axis=0)
        seg[s, 3:6] = np.sum(dM_inv[m], axis=0)
        seg[s, 6:9] = np.sum(dF_visc[m], axis=0)
        seg[s, 9:12] = np.sum(dM_visc[m], axis=0)
    D.update(gamma=gamma, v_i=v_i, alpha=alpha, CD=CD, Cm=Cm, Re=Re, M=M, dF_inv=dF_inv, dM_inv=dM_inv,
             dF_visc=dF_visc, dM_visc=dM_visc, V_i_2=V_i_2, redim_full=redim_full, redim_in_plane=redim_ip)
    return seg, D


def solve(T, solver_type="nonlinear", convergence=1e-10, relaxation=1.0, max_iterations=100, block=256):
    """Whole path for one flight state; mirrors Scene.solve_forces (scene.py:1462-1500) without the dict."""
    F = flow_state(T)
    V = influence(T, F, block=block)
    gamma_lin = solve_linear(T, F, V)
    out = {"flow": F, "V": V, "gamma_linear": gamma_lin, "iterations": 0, "error": 0.0, "converged": True, "history": []}
    gamma = gamma_lin
    if solver_type == "nonlinear":
        gamma, it, err, ok, hist = solve_nonlinear(T, F, V, gamma_lin, convergence, relaxation, max_iterations)
        out.update(iterations=it, error=err, converged=ok, history=hist)
    seg, D = integrate(T, F, V, gamma)
    out.update(gamma=gamma, seg_fm=seg, dist=D, section_bounds=T.section_bounds)
    return out
