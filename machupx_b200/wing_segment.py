"""Host-side geometry of one wing segment (one side of one lifting surface).

This is O(N) set-up work, not the hot path, but every table the device consumes derives from it, so the
numerical procedures follow the reference's machupX/wing_segment.py step for step (cited below) — same
clustering, same scipy quadrature of the sweep/dihedral distributions, same second-order ``numpy.gradient``
and linear ``interp1d`` of the section unit vectors — while the code is organised differently: span-wise
distributions are small callable objects instead of closures, and the quarter-chord integration caches whole
break-point intervals instead of re-integrating them for every span station.

Public attribute and method names match the reference (``N``, ``b``, ``side``, ``nodes``, ``control_points``,
``cp_span_locs``, ``u_a_cp`` ..., ``get_chord``, ``get_sweep``, ``get_cp_CL`` ...), because Scene-level code and
user scripts read them.  CAD export (stl/stp/dxf/vtk outlines) is outside the solve path and not provided.
"""
import math
import warnings

import numpy as np
import scipy.integrate as integ
import scipy.interpolate as interp

from .units import read_value


# ---------------------------------------------------------------------------------------------------
# span-wise distributions
# ---------------------------------------------------------------------------------------------------
class SpanConstant:
    """f(s) = value (optionally sign-flipped).  Reference: wing_segment.py:447-476."""

    def __init__(self, value, negate=False):
        self.value = value
        self.negate = negate

    def __call__(self, span):
        if isinstance(span, float):
            return -self.value if self.negate else self.value
        out = np.full(np.shape(span), self.value)
        return -out if self.negate else out


class SpanTable:
    """Piecewise-linear f(s) from a two-column table.  Reference: wing_segment.py:479-516."""

    def __init__(self, table, angular=False, negate=False):
        if isinstance(table[0], np.void):   # structured rows (integers in the JSON / csv)
            table = np.array([[row[0], row[1]] for row in table], dtype=float)
        self.table = np.array(table, dtype=float, copy=True)
        self.angular = angular
        self.negate = negate

    def __call__(self, span):
        scalar = isinstance(span, float)
        s = np.asarray(span, dtype=float)[np.newaxis] if scalar else span
        y = np.radians(self.table[:, 1]) if self.angular else self.table[:, 1]
        out = np.interp(s, self.table[:, 0], y)
        if self.negate:
            out = -out
        return out.item() if scalar else out


class SpanElliptic:
    """c(s) = c_root sqrt(1 - s^2).  Reference: wing_segment.py:521-528."""

    def __init__(self, root_chord):
        self.root_chord = root_chord

    def __call__(self, span):
        return self.root_chord * np.sqrt(1 - span * span)


class SpanSigned:
    """User callable with the side-dependent sign convention (wing_segment.py:345-353, 400-408)."""

    def __init__(self, fn, negate):
        self.fn = fn
        self.negate = negate

    def __call__(self, span):
        return -self.fn(span) if self.negate else self.fn(span)


def _distribution(data, angular=False, negate=False):
    """(callable, stored parameter) for a constant or tabulated input."""
    if isinstance(data, float):
        value = math.radians(data) if angular else data
        return SpanConstant(value, negate), value
    fn = SpanTable(data, angular, negate)
    return fn, fn.table


# Quadratures of the sweep / dihedral integrands, shared by every wing segment of the process whose sweep and dihedral inputs are
# the same numbers on the same side (the 64 identical aircraft of a swarm ask for the same 10^4 integrals 64 times).  The memo
# returns the very floats scipy produced the first time, so the geometry stays bit-identical; callables are never memoised.
_QUAD_MEMO = {}
_QUAD_MEMO_LIMIT = 100000      # ~30 MB; cleared when full (a 64 000-point swarm of identical aircraft needs 2 000 entries)


def _memo_key(data, negate):
    if isinstance(data, float):
        return ("c", data, bool(negate))
    if isinstance(data, np.ndarray):
        return ("t", data.shape, data.tobytes(), bool(negate))
    return None


def _as_span_array(span):
    return np.asarray(span, dtype=float)[np.newaxis] if isinstance(span, float) else np.asarray(span)


# ---------------------------------------------------------------------------------------------------
class WingSegment:
    def __init__(self, name, input_dict, side, unit_sys, airfoil_dict, origin=(0.0, 0.0, 0.0)):
        self.name = name
        self._input_dict = input_dict
        self._unit_sys = unit_sys
        self.side = side
        self._origin = np.asarray(origin, dtype=float)
        self._getter_data = {}
        self.parent_has_mirror = False

        self.ID = input_dict.get("ID")
        if self.ID == 0 and name != "origin":
            raise IOError("Wing segment ID for {0} may not be 0.".format(name))
        self.has_mirror = False if name == "origin" else input_dict.get("side", "both") == "both"
        if self.ID == 0:
            return

        self._read_layout()
        self._read_airfoils(airfoil_dict)
        self._build_distributions()
        self._place_lifting_line()
        self._build_section_frames()
        self._setup_control_surface(input_dict.get("control_surface", None))
        self._setup_cp_data()
        self._setup_node_data()

    # -- grid and connection ------------------------------------------------------------------------
    def _read_layout(self):
        """Grid options, connection offsets and the span-wise node / control-point stations
        (wing_segment.py:87-229)."""
        d = self._input_dict
        self.is_main = d.get("is_main", False)
        self._shear_dihedral = d.get("shear_dihedral", False)
        grid = d.get("grid", {})
        self.N = grid.get("N", 40)
        spacing = grid.get("distribution", "cosine_cluster")
        self.reid_corr = grid.get("reid_corrections", True)
        self.delta_joint = grid.get("joint_length", 0.15)
        self.blend_dist = grid.get("blending_distance", 1.0)

        connect = d.get("connect_to", {})
        self._connected_to_ID = connect.get("ID", 0)
        self._connected_to_loc = connect.get("location", "tip")
        self._delta_origin = np.array([connect.get("dx", 0.0), connect.get("dy", 0.0), connect.get("dz", 0.0)], dtype=float)
        self.y_offset = connect.get("y_offset", 0.0)
        self._delta_origin[1] += -self.y_offset if self.side == "left" else self.y_offset

        if spacing == "cosine_cluster":
            nodes, cps = self._cosine_stations(grid)
        elif spacing == "linear":
            nodes = np.linspace(0.0, 1.0, self.N + 1)
            cps = np.linspace(1 / (2 * self.N), 1.0 - 1 / (2 * self.N), self.N)
        elif isinstance(spacing, list):
            if len(spacing) != self.N * 2 + 1:
                raise IOError("User specified distribution must have length of 2*N+1. Got length {0}; needed length {1}."
                              .format(len(spacing), self.N * 2 + 1))
            if spacing[0] != 0.0 or spacing[-1] != 1.0:
                raise IOError("User specified distribution must begin at 0 and end at 1.")
            if not all(spacing[i] < spacing[i + 1] for i in range(len(spacing) - 1)):
                raise IOError("User specified distribution must be monotonically increasing.")
            nodes = np.array(spacing[0::2])
            cps = np.array(spacing[1::2])
        else:
            raise IOError("Distribution type {0} not recognized for wing segment {1}.".format(spacing, self.name))

        # indices always run from the left to the right (positive circulation = positive lift)
        if self.side == "left":
            nodes, cps = nodes[::-1], cps[::-1]
        self.node_span_locs = nodes
        self.cp_span_locs = cps

    def _cosine_stations(self, grid):
        """Cosine clustering between the root, the tip and any interior cluster points (wing_segment.py:124-196)."""
        breaks = []
        if grid.get("flap_edge_cluster", True):
            flap = self._input_dict.get("control_surface", None)
            if flap is not None:
                breaks += [flap.get("root_span", 0.0), flap.get("tip_span", 1.0)]
        breaks += list(grid.get("cluster_points", []))
        breaks = sorted(x for x in breaks if x != 0.0 and x != 1.0)
        breaks = [0.0] + breaks + [1.0]
        counts = [int(round(self.N * (hi - lo))) for lo, hi in zip(breaks[:-1], breaks[1:])]
        counts[0] -= int(sum(counts) - self.N)     # the root section absorbs the rounding

        node_locs, cp_locs = [0.0], []
        for lo, hi, n in zip(breaks[:-1], breaks[1:], counts):
            if n == 0:
                warnings.warn("Not enough control points for {0} to distribute between {1} and {2} percent span. Properties "
                              "of this section will not factor into results. If undesired, increase number of control "
                              "points or alter clustering.".format(self.name, int(lo * 100), int(hi * 100)))
                continue
            for theta in list(np.linspace(0.0, np.pi, n + 1))[1:]:
                node_locs.append(lo + 0.5 * (1 - np.cos(theta)) * (hi - lo))
            for theta in np.linspace(np.pi / n, np.pi, n) - np.pi / (2 * n):
                cp_locs.append(lo + 0.5 * (1 - np.cos(theta)) * (hi - lo))
        return np.array(node_locs), np.array(cp_locs)

    def is_continuation(self):
        """True if this segment starts exactly at the tip of another one (wing_segment.py:232-236)."""
        return self._connected_to_ID != 0 and self._connected_to_loc == "tip" and np.linalg.norm(self._delta_origin) < 1e-12

    # -- airfoils -------------------------------------------------------------------------------------
    def _read_airfoils(self, airfoil_dict):
        """One airfoil, or a span-wise list of them with the control-point range each pair brackets
        (wing_segment.py:559-624)."""
        default = list(airfoil_dict.keys())[0]
        spec = read_value("airfoil", self._input_dict, self._unit_sys, default)
        self._airfoils, spans = [], []
        if isinstance(spec, str):
            if spec not in airfoil_dict:
                raise IOError("'{0}' must be specified in 'airfoils'.".format(spec))
            self._airfoils.append(airfoil_dict[spec])
            self._airfoil_slices = [slice(0, self.N)]
        elif isinstance(spec, np.ndarray):
            for row in spec:
                af_name = row[1].item()
                if af_name not in airfoil_dict:
                    raise IOError("'{0}' must be specified in 'airfoils'.".format(af_name))
                self._airfoils.append(airfoil_dict[af_name])
                spans.append(float(row[0]))
            self._airfoil_slices = []
            if self.side == "right":
                start = 0
                for s in spans[1:]:
                    stop = int(np.sum(self.cp_span_locs < s))
                    self._airfoil_slices.append(slice(start, stop))
                    start = stop
            else:
                stop = self.N
                for s in spans[1:]:
                    start = int(np.sum(self.cp_span_locs > s))
                    self._airfoil_slices.append(slice(start, stop))
                    stop = start
        else:
            raise IOError("Airfoil definition must a be a string or an array.")
        self._num_airfoils = len(self._airfoils)
        self._airfoil_spans = np.asarray(spans)

    # -- span-wise distributions -----------------------------------------------------------------------
    def _build_distributions(self):
        """twist / dihedral / sweep / chord getters and the quarter-chord description (wing_segment.py:239-432)."""
        d, us = self._input_dict, self._unit_sys
        self.b = read_value("semispan", d, us, "not_given")
        qc_pts = read_value("quarter_chord_locs", d, us, "not_given")
        dihedral = read_value("dihedral", d, us, "not_given")
        sweep = read_value("sweep", d, us, "not_given")
        have_pts = isinstance(qc_pts, np.ndarray)
        if isinstance(self.b, str) and not have_pts:
            raise IOError("Either 'semispan' or 'quarter_chord_locs' must be specified.")
        if not isinstance(self.b, str) and have_pts:
            raise IOError("'semispan' and 'quarter_chord_locs' may not both be specified at once.")
        if not isinstance(dihedral, str) and have_pts:
            raise IOError("'dihedral' and 'quarter_chord_locs' may not both be specified at once.")
        if not isinstance(sweep, str) and have_pts:
            raise IOError("'sweep' and 'quarter_chord_locs' may not both be specified at once.")

        twist = read_value("twist", d, us, 0.0)
        if callable(twist):
            self.get_twist = twist
        else:
            self.get_twist, self._getter_data["twist"] = _distribution(twist, angular=True)

        if have_pts:
            self._qc_data_type = "points"
            pts = np.zeros((qc_pts.shape[0] + 1, 4))
            pts[1:, 1:] = qc_pts
            # arc length in the y-z plane defines the span coordinate (wing_segment.py:265-283)
            pts[1:, 0] = np.cumsum(np.linalg.norm(np.diff(pts[:, 2:], axis=0), axis=1))
            self.b = float(pts[-1, 0])
            pts[:, 0] /= self.b
            self._qc_loc_data = pts
            self.get_dihedral = self._dihedral_from_points
            self.get_sweep = self._sweep_from_points
        else:
            self._qc_data_type = "standard"
            self._discont = []
            if isinstance(dihedral, str):
                dihedral = 0.0
            if isinstance(sweep, str):
                sweep = 0.0
            if callable(dihedral):
                self.get_dihedral = SpanSigned(dihedral, negate=(self.side != "left"))
            else:
                self.get_dihedral, self._getter_data["dihedral"] = _distribution(dihedral, angular=True, negate=(self.side == "right"))
                self._note_breaks(self._getter_data["dihedral"])
            if callable(sweep):
                self.get_sweep = SpanSigned(sweep, negate=(self.side == "left"))
            else:
                self.get_sweep, self._getter_data["sweep"] = _distribution(sweep, angular=True, negate=(self.side == "left"))
                self._note_breaks(self._getter_data["sweep"])
            for end in (0.0, 1.0):
                if end not in self._discont:
                    self._discont.append(end)
            self._discont = sorted(self._discont)
            self._interval_cache = {}
            ks = None if callable(sweep) else _memo_key(self._getter_data["sweep"], self.side == "left")
            kd = None if callable(dihedral) else _memo_key(self._getter_data["dihedral"], self.side == "right")
            self._quad_key = None if ks is None or kd is None else (ks, kd)

        chord = read_value("chord", d, us, 1.0)
        if isinstance(chord, tuple):
            self._root_chord = chord[1]
            self.get_chord = SpanElliptic(chord[1])
        elif callable(chord):
            self.get_chord = chord
        else:
            self.get_chord, self._getter_data["chord"] = _distribution(chord)

    def _note_breaks(self, table):
        if isinstance(table, np.ndarray):
            for s in table[:, 0]:
                if s.item() not in self._discont:
                    self._discont.append(s.item())

    def _chord_slope_points(self, s):
        if s < 0.005:
            return self._get_quarter_chord_loc(s), self._get_quarter_chord_loc(s + 0.01)
        if s > 0.995:
            return self._get_quarter_chord_loc(s - 0.01), self._get_quarter_chord_loc(s)
        return self._get_quarter_chord_loc(s - 0.005), self._get_quarter_chord_loc(s + 0.005)

    def _dihedral_from_points(self, span):
        """Finite-difference dihedral of a quarter-chord polyline (wing_segment.py:310-341)."""
        scalar = isinstance(span, float)
        s_arr = _as_span_array(span)
        out = np.zeros_like(s_arr)
        for i, s in enumerate(s_arr):
            p0, p1 = self._chord_slope_points(float(s))
            out[i] = np.arctan((p1[2] - p0[2]) / (p1[1] - p0[1]))
        return out.item() if scalar else out

    def _sweep_from_points(self, span):
        """Finite-difference sweep of a quarter-chord polyline; the denominator is the reference's own expression
        (wing_segment.py:389: the square root is applied to the y difference before squaring)."""
        scalar = isinstance(span, float)
        s_arr = _as_span_array(span)
        out = np.zeros_like(s_arr)
        for i, s in enumerate(s_arr):
            p0, p1 = self._chord_slope_points(float(s))
            out[i] = -np.arctan((p1[0] - p0[0]) / (np.sqrt((p1[1] - p0[1])) ** 2 + (p1[2] - p0[2]) ** 2))
        return out.item() if scalar else out

    # -- quarter chord and lifting line ------------------------------------------------------------------
    def _interval(self, lo, hi):
        """Integrals of (tan sweep, -cos dihedral, -sin dihedral) over [lo, hi] of the span fraction, by adaptive
        quadrature exactly as the reference does (wing_segment.py:953-959); whole intervals between break points are
        cached because every station beyond them needs the same numbers."""
        key = (lo, hi)
        hit = self._interval_cache.get(key)
        if hit is None:
            hit = self._quad3(lo, hi)
            self._interval_cache[key] = hit
        return hit

    def _quad3(self, lo, hi):
        """The three integrals over [lo, hi] (process-wide memo, see _QUAD_MEMO)."""
        key = None if self._quad_key is None else (self._quad_key, lo, hi)
        hit = _QUAD_MEMO.get(key) if key is not None else None
        if hit is None:
            hit = np.array([
                integ.quad(lambda s: np.tan(self.get_sweep(s)), lo, hi)[0],
                integ.quad(lambda s: -np.cos(self.get_dihedral(s)), lo, hi)[0],
                integ.quad(lambda s: -np.sin(self.get_dihedral(s)), lo, hi)[0]])
            if key is not None:
                if len(_QUAD_MEMO) >= _QUAD_MEMO_LIMIT:
                    _QUAD_MEMO.clear()
                _QUAD_MEMO[key] = hit
        return hit

    def _get_quarter_chord_loc(self, span):
        scalar = isinstance(span, float)
        s_arr = _as_span_array(span)
        offset = np.zeros((s_arr.shape[0], 3))
        if self._qc_data_type == "standard":
            brk = self._discont
            for i, s in enumerate(s_arr):
                s = float(s)
                for j in range(1, len(brk)):
                    if s > brk[j]:
                        offset[i] += self._interval(brk[j - 1], brk[j]) * self.b
                    else:
                        # each station has its own upper limit: not in the per-segment cache, but shared between equal segments
                        offset[i] += self._quad3(brk[j - 1], s) * self.b
                        break
            loc = self.get_root_loc() + offset if self.side == "left" else self.get_root_loc() - offset
        else:
            tab = self._qc_loc_data
            offset[:, 0] = np.interp(s_arr, tab[:, 0], tab[:, 1])
            y = np.interp(s_arr, tab[:, 0], tab[:, 2])
            offset[:, 1] = -y if self.side == "left" else y
            offset[:, 2] = np.interp(s_arr, tab[:, 0], tab[:, 3])
            loc = self.get_root_loc() + offset
        return loc.flatten() if scalar else loc

    def get_root_loc(self):
        return self._origin if self.ID == 0 else self._origin + self._delta_origin

    def get_tip_loc(self):
        return self._origin if self.ID == 0 else self._get_quarter_chord_loc(1.0)

    def _unswept_trig(self, span):
        s = _as_span_array(span)
        tw, di = self.get_twist(s), self.get_dihedral(s)
        return s, np.cos(tw), np.sin(tw), np.cos(di), np.sin(di)

    def _get_unswept_axial_vec(self, span):
        _, ct, st, cd, sd = self._unswept_trig(span)                           # wing_segment.py:989-1004
        return np.asarray([-ct, -st * sd, st * cd]).T

    def _get_unswept_normal_vec(self, span):
        _, ct, st, cd, sd = self._unswept_trig(span)                           # wing_segment.py:1007-1022
        return np.asarray([-st, ct * sd, -ct * cd]).T

    def _get_unswept_span_vec(self, span):
        s = _as_span_array(span)                                               # wing_segment.py:1025-1037
        di = self.get_dihedral(s)
        return np.asarray([np.zeros(s.size), np.cos(di), np.sin(di)]).T

    def _get_ll_loc(self, span):
        """Lifting line = quarter chord + ll_offset * chord along the unswept axial direction (wing_segment.py:1040-1053)."""
        s = _as_span_array(span)
        loc = self._get_quarter_chord_loc(s)
        loc += (self._get_ll_offset(s) * self.get_chord(s))[:, np.newaxis] * self._get_unswept_axial_vec(s)
        return loc

    def _place_lifting_line(self):
        """Optional Kuchemann shift of the locus of aerodynamic centres, then nodes and control points
        (wing_segment.py:713-801)."""
        offset = read_value("ll_offset", self._input_dict, self._unit_sys, 0)
        if isinstance(offset, str) and offset == "kuchemann":
            offset = self._kuchemann_offset()
        if callable(offset):
            self._get_ll_offset = offset
        else:
            self._get_ll_offset, self._getter_data["ll_offset"] = _distribution(offset)
        self.control_points = self._get_ll_loc(self.cp_span_locs)
        self.nodes = self._get_ll_loc(self.node_span_locs)

    def _kuchemann_offset(self):
        msg = ("Kuchemann's equations for the locus of aerodynamic centers cannot be used in the case of non-constant "
               "sweep. Reverting to no offset.")
        sweep_data = self._getter_data.get("sweep", None)
        if sweep_data is not None:
            if not isinstance(sweep_data, float):
                warnings.warn(msg)
                return 0.0
        elif self._qc_data_type == "points":
            probe = self.get_sweep(np.linspace(0.0, 1.0, 10))
            if not np.allclose(np.full(10, probe[0]), probe, rtol=1e-10, atol=1e-3):
                warnings.warn(msg)
                return 0.0
        else:
            warnings.warn(msg)
            return 0.0

        CLa_root = self._airfoils[0].get_CLa(alpha=0.0)                        # wing_segment.py:751-766
        area = integ.quad(lambda s: self.get_chord(s), 0, 1)[0]
        R_A = 2.0 * self.b / area
        sweep = abs(self.get_sweep(0.0))
        sweep_eff = sweep / ((1 + (CLa_root * math.cos(sweep) / (math.pi * R_A)) ** 2) ** 0.25)
        tan_k = math.tan(sweep_eff)
        try:
            sweep_div = tan_k / sweep_eff
        except ZeroDivisionError:
            sweep_div = 1.0
        expo = math.pi / (4.0 * (math.pi + 2.0 * abs(sweep_eff)))
        K = (1 + (CLa_root * math.cos(sweep_eff) / (math.pi * R_A)) ** 2) ** expo

        locs = np.copy(self.node_span_locs)[::-1] if self.side == "left" else np.copy(self.node_span_locs)
        z = locs * self.b                                                      # wing_segment.py:769-789
        c = self.get_chord(locs)
        with np.errstate(divide="ignore", invalid="ignore"):
            from_centre = np.where(c != 0.0, z / c, 0.0)
            from_tip = np.where(c != 0.0, (self.b - z) / c, 0.0)
        two_pi = 2.0 * math.pi
        lam_c = np.sqrt(1 + (two_pi * sweep_div * from_centre) ** 2) - two_pi * sweep_div * from_centre
        lam_t = np.sqrt(1 + (two_pi * sweep_div * from_tip) ** 2) - two_pi * sweep_div * from_tip
        lam = lam_c - lam_t
        shift = -(0.25 * (1.0 - 1.0 / K * (1.0 + 2.0 * lam * sweep_eff / math.pi)))
        return np.concatenate((locs[:, np.newaxis], shift[:, np.newaxis]), axis=1)

    # -- section frames --------------------------------------------------------------------------------
    def _build_section_frames(self):
        """Swept unit vectors at the nodes from the lifting-line tangent (wing_segment.py:531-556)."""
        line = self._get_ll_loc(self.node_span_locs)
        arc = np.zeros(self.N + 1)
        arc[1:] = np.cumsum(np.linalg.norm(np.diff(line, axis=0), axis=1))
        tangent = np.gradient(line, arc, edge_order=2, axis=0)
        self._u_s_dist = tangent / np.linalg.norm(tangent, axis=1, keepdims=True)
        self._get_span_vec = interp.interp1d(self.node_span_locs, self._u_s_dist, axis=0)

        ua0 = self._get_unswept_axial_vec(self.node_span_locs)
        k = np.einsum("ij,ij->i", self._u_s_dist, ua0)
        c1 = np.sqrt(1 / (1 - k * k))
        c2 = -c1 * k
        ua = c1[:, np.newaxis] * ua0 + c2[:, np.newaxis] * self._u_s_dist
        self._u_a_dist = ua / np.linalg.norm(ua, axis=1, keepdims=True)
        self._get_axial_vec = interp.interp1d(self.node_span_locs, self._u_a_dist, axis=0)

        self._u_n_dist = np.cross(self._u_a_dist, self._u_s_dist)
        self._get_normal_vec = interp.interp1d(self.node_span_locs, self._u_n_dist, axis=0)

    # -- control surface ---------------------------------------------------------------------------------
    def _setup_control_surface(self, control_dict):
        """Flap extent, chord fraction per control point and control mixing (wing_segment.py:627-659)."""
        self._delta_flap = np.zeros(self.N)      # positive down
        self._cp_c_f = np.zeros(self.N)
        self._has_control_surface = control_dict is not None
        if not self._has_control_surface:
            return
        self._cntrl_root_span = control_dict.get("root_span", 0.0)
        self._cntrl_tip_span = control_dict.get("tip_span", 1.0)
        self._saturation_angle = np.radians(control_dict.get("saturation_angle", np.inf))
        self._cp_in_cntrl_surf = (self.cp_span_locs >= self._cntrl_root_span) & (self.cp_span_locs <= self._cntrl_tip_span)
        frac = read_value("chord_fraction", control_dict, self._unit_sys, 0.25)
        if not isinstance(frac, float):
            if frac[0, 0] != self._cntrl_root_span or frac[-1, 0] != self._cntrl_tip_span:
                raise IOError("Endpoints of flap chord distribution must match specified root and tip span locations.")
        self.get_c_f, self._getter_data["flap_chord_fraction"] = _distribution(frac)
        self._cp_c_f[self._cp_in_cntrl_surf] = self.get_c_f(self.cp_span_locs[self._cp_in_cntrl_surf])
        self._control_mixing = control_dict.get("control_mixing", {})

    def apply_control(self, control_state, control_symmetry):
        """Flap deflection [rad] at every control point from the aircraft-level control state
        (wing_segment.py:1327-1374)."""
        if not self._has_control_surface:
            return
        total = np.zeros(self.N)
        for key in self._control_mixing:
            deflection = read_value(key, control_state, self._unit_sys, 0.0)
            if isinstance(deflection, np.ndarray):
                if deflection[0, 0] != self._cntrl_root_span or deflection[-1, 0] != self._cntrl_tip_span:
                    raise IOError("Endpoints of flap deflection distribution must match specified root and tip span locations.")
                spread = np.zeros(self.N)
                spread[self._cp_in_cntrl_surf] = np.interp(self.cp_span_locs[self._cp_in_cntrl_surf], deflection[:, 0], deflection[:, 1])
                deflection = spread
            elif callable(deflection):
                deflection = deflection(self.cp_span_locs)
            contribution = deflection * self._control_mixing.get(key, 0.0) * self._cp_in_cntrl_surf
            if self.side == "right" or control_symmetry[key]:
                total += contribution
            else:
                total -= contribution
        total = np.radians(total)
        total = np.where(total > self._saturation_angle, self._saturation_angle, total)
        total = np.where(total < -self._saturation_angle, -self._saturation_angle, total)
        self._delta_flap = total

    # -- cached per-point data ------------------------------------------------------------------------------
    def _setup_cp_data(self):
        s = self.cp_span_locs                                                  # wing_segment.py:679-705
        self.u_a_cp = self._get_axial_vec(s)
        self.u_n_cp = self._get_normal_vec(s)
        self.u_s_cp = self._get_span_vec(s)
        self.u_a_cp_unswept = self._get_unswept_axial_vec(s)
        self.u_n_cp_unswept = self._get_unswept_normal_vec(s)
        self.u_s_cp_unswept = self._get_unswept_span_vec(s)
        self.c_bar_cp = self._get_cp_avg_chord_lengths()
        self.dihedral_cp = self.get_dihedral(s)
        self.sweep_cp = self.get_sweep(s)
        self.twist_cp = self.get_twist(s)
        self.dS = abs(self.node_span_locs[1:] - self.node_span_locs[:-1]) * self.b * self.c_bar_cp
        camber = np.array([a.get_max_camber() for a in self._airfoils], dtype=float)
        thick = np.array([a.get_max_thickness() for a in self._airfoils], dtype=float)
        if self._num_airfoils == 1:
            self.max_camber_cp = np.ones(self.N) * camber[0]
            self.max_thickness_cp = np.ones(self.N) * thick[0]
        else:
            self.max_camber_cp = np.interp(s, self._airfoil_spans, camber)
            self.max_thickness_cp = np.interp(s, self._airfoil_spans, thick)

    def _setup_node_data(self):
        self.u_a_node = self._get_axial_vec(self.node_span_locs)               # wing_segment.py:708-710
        self.c_node = self.get_chord(self.node_span_locs)

    def _get_cp_avg_chord_lengths(self):
        c = self.get_chord(self.node_span_locs)                                # wing_segment.py:1056-1059
        return (c[1:] + c[:-1]) / 2

    # -- section coefficients on the host (API parity; the solver evaluates them on the device) ----------------
    def airfoil_blend(self):
        """(index of the inboard bracketing airfoil, weight toward the outboard one) per control point
        (wing_segment.py:1067-1070); single-airfoil segments return zeros."""
        if self._num_airfoils == 1:
            return np.zeros(self.N, dtype=int), np.zeros(self.N)
        j = np.searchsorted(self._airfoil_spans, self.cp_span_locs) - 1
        j = np.where(j < 0, 0, j)
        d = (self.cp_span_locs - self._airfoil_spans[j]) / (self._airfoil_spans[j + 1] - self._airfoil_spans[j])
        return j, d

    def _get_control_point_coef(self, alpha, Rey, Mach, coef_func):
        """Coefficient ``coef_func`` at every control point (wing_segment.py:1075-1127)."""
        flap = dict(trailing_flap_deflection=self._delta_flap, trailing_flap_fraction=self._cp_c_f)
        if self._num_airfoils == 1:
            return getattr(self._airfoils[0], coef_func)(alpha=alpha, Rey=Rey, Mach=Mach, **flap)
        vals = np.zeros((self.N, self._num_airfoils))
        for j, sl in enumerate(self._airfoil_slices):
            kw = dict(alpha=alpha[sl], Rey=Rey[sl], Mach=Mach[sl], trailing_flap_deflection=self._delta_flap[sl],
                      trailing_flap_fraction=self._cp_c_f[sl])
            vals[sl, j] = getattr(self._airfoils[j], coef_func)(**kw)
            vals[sl, j + 1] = getattr(self._airfoils[j + 1], coef_func)(**kw)
        j, d = self.airfoil_blend()
        i = np.arange(self.N)
        return (1 - d) * vals[i, j] + d * vals[i, j + 1]

    def get_cp_CLa(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_CLa")

    def get_cp_aL0(self, Rey, Mach):
        return self._get_control_point_coef(np.zeros_like(Rey), Rey, Mach, "get_aL0")

    def get_cp_CLRe(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_CLRe")

    def get_cp_CLM(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_CLM")

    def get_cp_CL(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_CL")

    def get_cp_CD(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_CD")

    def get_cp_Cm(self, alpha, Rey, Mach):
        return self._get_control_point_coef(alpha, Rey, Mach, "get_Cm")
