"""Host-side aircraft: flight state, controls, wing segments and the O(N) geometry tables of the solver.

API mirror of the reference's machupX/airplane.py (``Airplane(name, input, unit_system, scene, init_state,
init_control_state, v_wind)``, ``set_state``, ``get_state``, ``set_aerodynamic_state``, ``get_aerodynamic_state``,
``set_control_state``, ``add_wing_segment``, ``get_MAC`` and the per-control-point attribute names).

The decisive difference: the reference materialises the effective lifting line seen from every control point
(``P0_eff``, ``P1_eff``, ``P0_joint_eff``, ``P1_joint_eff``, ``r_0`` ..., eleven N x N (x 3) arrays built by a
Python loop, airplane.py:569-686).  Here none of them exist.  ``_calculate_geometry`` stops at the per-point
rows; the blended, jointed geometry is recomputed per (control point, vortex) pair in registers by the device
influence kernel (csrc/influence.cu) from exactly these rows.
"""
import json
import math

import numpy as np

from .airfoil import Airfoil
from .rotations import body_to_earth, earth_to_body, euler_to_quat, quat_conjugate
from .units import read_value, require_file
from .wing_segment import WingSegment


class Airplane:
    def __init__(self, name, airplane_input, unit_system, scene, init_state={}, init_control_state={}, v_wind=[0.0, 0.0, 0.0]):
        self.name = name
        self._unit_sys = unit_system
        self._scene = scene
        self.wing_segments = {}
        self._airfoil_database = {}
        self.N = 0

        self._load_params(airplane_input)
        self.set_state(**init_state, v_wind=v_wind)
        self._create_airfoil_database()
        self._origin_segment = WingSegment("origin", {"ID": 0, "is_main": False}, "both", unit_system, self._airfoil_database)
        self._load_wing_segments()
        self._check_reference_params()
        self._initialize_controls(init_control_state)

    # -- input ------------------------------------------------------------------------------------------
    def _load_params(self, airplane_input):
        if isinstance(airplane_input, str):                                    # airplane.py:84-95
            require_file(airplane_input, ".json")
            with open(airplane_input) as handle:
                self._input_dict = json.load(handle)
        elif isinstance(airplane_input, dict):
            self._input_dict = airplane_input
        else:
            raise IOError("{0} is not an allowed airplane input type. Must be path or dictionary.".format(type(airplane_input)))
        ref = self._input_dict.get("reference", {})
        us = self._unit_sys
        self.CG = read_value("CG", self._input_dict, us, [0.0, 0.0, 0.0])       # airplane.py:98-102
        self.W = read_value("weight", self._input_dict, us, None)
        self.S_w = read_value("area", ref, us, -1)
        self.l_ref_lon = read_value("longitudinal_length", ref, us, -1)
        self.l_ref_lat = read_value("lateral_length", ref, us, -1)

    def _create_airfoil_database(self):
        airfoils = self._input_dict.get("airfoils", {"default": {}})           # airplane.py:906-922
        if isinstance(airfoils, str):
            require_file(airfoils, ".json")
            with open(airfoils, "r") as handle:
                airfoils = json.load(handle)
        elif not isinstance(airfoils, dict):
            raise IOError("'airfoils' must be a string or dict.")
        for key, spec in airfoils.items():
            self._airfoil_database[key] = Airfoil(key, spec)

    # -- flight state -------------------------------------------------------------------------------------
    def set_state(self, **kwargs):
        """position / velocity (+alpha, beta) / orientation / angular_rates (+frame); airplane.py:105-213.
        The velocity is stored in earth-fixed axes."""
        v_wind = kwargs.get("v_wind", [0.0, 0.0, 0.0])
        us = self._unit_sys
        self.state_type = read_value("type", kwargs, us, "none")
        if self.state_type == "rigid_body":
            raise IOError("Use of 'rigid_body' state type is no longer supported and the 'state_type' key in general is "
                          "depreciated. See documentation for details.")
        self.p_bar = read_value("position", kwargs, us, [0.0, 0.0, 0.0])

        q = read_value("orientation", kwargs, us, [1.0, 0.0, 0.0, 0.0])
        if q.shape[0] == 3:
            self.q = euler_to_quat(np.radians(q))
        elif q.shape[0] == 4:
            self.q = q / np.linalg.norm(q)
        else:
            raise IOError("{0} is not an allowable orientation definition.".format(q))

        v_value = read_value("velocity", kwargs, us, None)
        alpha = beta = None
        if isinstance(v_value, float):
            alpha = read_value("alpha", kwargs, us, 0.0)
            beta = read_value("beta", kwargs, us, 0.0)
            self.v = np.array([9.181994, 0.0, 0.0])   # placeholder direction; overwritten by the next call
            self.set_aerodynamic_state(alpha=alpha, beta=beta, velocity=v_value, v_wind=v_wind)
        elif isinstance(v_value, np.ndarray) and v_value.shape == (3,):
            if "alpha" in kwargs or "beta" in kwargs:
                raise IOError("Alpha and beta are not allowed when the freestream velocity is a vector.")
            self.v = body_to_earth(self.q, v_value)
        else:
            raise IOError("{0} is not an allowable velocity definition.".format(v_value))

        w_raw = read_value("angular_rates", kwargs, us, [0.0, 0.0, 0.0])
        self.angular_rate_frame = kwargs.get("angular_rate_frame", "body")
        if self.angular_rate_frame == "body":
            self.w = w_raw
        elif self.angular_rate_frame in ("stab", "wind"):
            if alpha is not None:
                a_rad, b_rad = math.radians(alpha), math.radians(beta)
            else:
                a_rad = math.atan2(v_value[2], v_value[0])
                b_rad = math.asin(v_value[1] / math.sqrt(v_value[0] ** 2 + v_value[1] ** 2 + v_value[2] ** 2))
            if self.angular_rate_frame == "stab":
                self.q_to_stab = quat_conjugate(euler_to_quat([0.0, a_rad, 0.0]))
                self.w = body_to_earth(self.q_to_stab, w_raw)
            else:
                self.q_to_wind = quat_conjugate(euler_to_quat([0.0, a_rad, b_rad]))
                self.w = body_to_earth(self.q_to_wind, w_raw)
        else:
            raise IOError("{0} is not an allowable angular rate frame.".format(self.angular_rate_frame))

    def get_state(self):
        return np.copy(self.v), np.copy(self.w), np.copy(self.p_bar), np.copy(self.q)

    def get_aerodynamic_state(self, v_wind=[0.0, 0.0, 0.0]):
        """(alpha [deg], beta [deg], V) of the air-relative velocity; airplane.py:237-265."""
        v = earth_to_body(self.q, self.v - v_wind)
        V = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        return math.degrees(math.atan2(v[2], v[0])), math.degrees(math.asin(v[1] / V)), V

    def set_aerodynamic_state(self, **kwargs):
        """Rescales / redirects the velocity to the requested alpha, beta, V without touching the orientation
        (Phillips, Mechanics of Flight eqs. 7.1.10-12; airplane.py:268-316)."""
        v_wind = kwargs.get("v_wind", [0.0, 0.0, 0.0])
        now = self.get_aerodynamic_state(v_wind=v_wind)
        alpha = kwargs.get("alpha", now[0])
        beta = kwargs.get("beta", now[1])
        speed = kwargs.get("velocity", now[2])
        C_a = math.cos(math.radians(alpha))
        S_a = math.sin(math.radians(alpha))
        flank = math.atan(math.tan(math.radians(beta)) / C_a)
        C_B, S_B = math.cos(flank), math.sin(flank)
        scale = 1 / math.sqrt(1 - S_a * S_a * S_B * S_B)
        v_body = np.zeros(3)
        v_body[0] = speed * C_a * C_B * scale
        v_body[1] = speed * C_a * S_B * scale
        v_body[2] = speed * S_a * C_B * scale
        self.v = v_wind + body_to_earth(self.q, v_body)

    # -- controls ---------------------------------------------------------------------------------------------
    def _initialize_controls(self, init_control_state):
        self._control_symmetry = {}                                            # airplane.py:319-333
        self.control_names = []
        self.current_control_state = {}
        for key, spec in self._input_dict.get("controls", {}).items():
            self.control_names.append(key)
            self._control_symmetry[key] = spec.get("is_symmetric", True)
            self.current_control_state[key] = 0.0
        self.set_control_state(control_state=init_control_state)

    def set_control_state(self, control_state={}):
        """Deflections not given revert to zero (airplane.py:925-944)."""
        self._controls_rev = getattr(self, "_controls_rev", 0) + 1   # lets Scene know the flap rows on the device are stale
        for key in self.current_control_state:
            self.current_control_state[key] = control_state.get(key, 0.0)
        for segment in self.wing_segments.values():
            segment.apply_control(control_state, self._control_symmetry)

    # -- wing segments --------------------------------------------------------------------------------------------
    def _find_parent(self, parent_ID, side):
        """The segment a new half attaches to: the origin, or the existing segment with that ID — its same-side half
        when it is mirrored (wing_segment.py:838-885)."""
        if parent_ID == 0:
            return self._origin_segment
        for segment in self.wing_segments.values():
            if segment.ID == parent_ID and not (segment.has_mirror and side not in segment.name):
                return segment
        return None

    def _attach(self, name, input_dict, side):
        connect = input_dict.get("connect_to", {})
        parent = self._find_parent(connect.get("ID", 0), side)
        if parent is None:
            raise RuntimeError("Could not attach wing segment {0}. Check ID of parent is valid.".format(name))
        if connect.get("location", "tip") == "root":
            anchor = np.array(parent.get_root_loc(), dtype=float)
            if parent.ID != 0:
                anchor[1] += parent.y_offset if parent.side == "left" else -parent.y_offset
        else:
            anchor = parent.get_tip_loc()
        segment = WingSegment(name, input_dict, side, self._unit_sys, self._airfoil_database, anchor)
        segment.parent_has_mirror = parent.has_mirror
        self.wing_segments[name] = segment
        self.N += segment.N

    def add_wing_segment(self, wing_segment_name, input_dict, recalculate_geometry=True):
        if wing_segment_name in self.wing_segments.keys():
            raise IOError("Wing segment {0} already exists in this airplane.".format(wing_segment_name))
        side = input_dict.get("side", "both")
        if side not in ("left", "right", "both"):
            raise IOError("{0} is not a proper side designation.".format(side))
        if side in ("left", "both"):
            self._attach(wing_segment_name + "_left", input_dict, "left")
        if side in ("right", "both"):
            self._attach(wing_segment_name + "_right", input_dict, "right")
        if recalculate_geometry:
            self._calculate_geometry()

    def _load_wing_segments(self):
        """Parents before children (airplane.py:405-431)."""
        wings = self._input_dict.get("wings", {})
        order = []
        for key, spec in wings.items():
            for pos, earlier in enumerate(order):
                if wings[earlier].get("connect_to", {}).get("ID", 0) == spec["ID"]:
                    order.insert(pos, key)
                    break
            else:
                order.append(key)
        for key in order:
            self.add_wing_segment(key, wings[key], recalculate_geometry=False)
        self._calculate_geometry()

    def _get_wing_segment(self, wing_segment_name):
        return self.wing_segments.get(wing_segment_name, False)

    def get_num_cps(self):
        return self.N

    # -- grouping of segments into lifting lines -----------------------------------------------------------------
    def _sort_segments_into_wings(self):
        """Which segments share one lifting line ("wing"), the scope of the effective-lifting-line blending
        (airplane.py:709-816).  A line is seeded by every segment that does not continue another one's tip, plus
        mirrored segments sitting on top of an unmirrored parent (T-tail); the left half of a mirrored segment
        without y-offset rides with its right half."""
        segs = self.wing_segments
        seeds = []
        for seg_name, seg in segs.items():
            seg.wing_ID = -1
            if seg.side == "left" and seg.has_mirror and abs(seg.y_offset) < 1e-12:
                continue
            if not seg.is_continuation() or (seg.has_mirror and not seg.parent_has_mirror):
                seeds.append(seg_name)

        self._segments_in_wings = []
        for line, seed_name in enumerate(seeds):
            seed = segs[seed_name]
            seed.wing_ID = line
            members = [seed]
            joined_at_middle = abs(seed.y_offset) < 1e-12
            if seed.has_mirror and joined_at_middle:
                twin = segs[seed_name.replace("right", "left")]
                twin.wing_ID = line
                members.append(twin)
            grew = True
            while grew:
                grew = False
                for seg_name, seg in segs.items():
                    if seg_name in seeds or seg.wing_ID != -1 or not seg.is_continuation():
                        continue
                    if seg._connected_to_ID == seed.ID:
                        if not seed.has_mirror and seg.has_mirror:
                            continue
                        if joined_at_middle or seg.side == seed.side:
                            seg.wing_ID = line
                            members.append(seg)
                            grew = True
                    else:
                        for other in members:
                            if seg._connected_to_ID != other.ID:
                                continue
                            if not other.has_mirror and seg.has_mirror:
                                continue
                            if seg.side == other.side:
                                seg.wing_ID = line
                                members.append(seg)
                                grew = True
            self._segments_in_wings.append(members)
        self._num_wings = len(seeds)
        self._sort_segments_left_to_right()

    def _sort_segments_left_to_right(self):
        """Left halves outermost-first, then right halves innermost-first, by tip distance from the x axis
        (airplane.py:819-869)."""
        def tip_radius(seg):
            tip = seg.get_tip_loc()
            return math.sqrt(tip[1] * tip[1] + tip[2] * tip[2])

        for w, members in enumerate(self._segments_in_wings):
            ordered = []
            while True:   # left: repeatedly take the farthest remaining (radius must exceed 0)
                best, best_r = None, 0.0
                for seg in members:
                    if "_left" in seg.name and seg not in ordered:
                        r = tip_radius(seg)
                        if best_r < r:
                            best, best_r = seg, r
                if best is None:
                    break
                ordered.append(best)
            while True:   # right: repeatedly take the nearest remaining
                best, best_r = None, np.inf
                for seg in members:
                    if "_right" in seg.name and seg not in ordered:
                        r = tip_radius(seg)
                        if best_r > r:
                            best, best_r = seg, r
                if best is None:
                    break
                ordered.append(best)
            self._segments_in_wings[w] = ordered
        self.segments = [seg for members in self._segments_in_wings for seg in members]

    # -- per-control-point rows ---------------------------------------------------------------------------------------
    def _calculate_geometry(self):
        """Concatenates the segments' data in scene order (wing -> segment -> control point), airplane.py:434-567, 689.
        No N x N arrays are formed (see the module docstring)."""
        self._sort_segments_into_wings()
        N = self.N
        self.c_bar, self.b_seg, self.dS = np.zeros(N), np.zeros(N), np.zeros(N)
        self.P0_chord, self.P1_chord = np.zeros(N), np.zeros(N)
        self.max_camber, self.max_thickness = np.zeros(N), np.zeros(N)
        self.PC, self.P0, self.P1 = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.PC_span_locs, self.P0_span_locs, self.P1_span_locs = np.zeros(N), np.zeros(N), np.zeros(N)
        self.u_a, self.u_n, self.u_s = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.u_a_unswept, self.u_n_unswept, self.u_s_unswept = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.P0_u_a, self.P1_u_a = np.zeros((N, 3)), np.zeros((N, 3))
        self.reid_corr = np.zeros(N)
        self.sigma_blend = np.zeros(N)
        self.delta_joint = np.zeros(N)

        self._wing_N, self.wing_slices = [], []
        k = 0
        for members in self._segments_in_wings:
            wing_start = k
            from_left_tip = 0.0
            for seg in members:
                sl = slice(k, k + seg.N)
                self.c_bar[sl] = seg.c_bar_cp
                self.b_seg[sl] = seg.b
                self.PC[sl, :] = seg.control_points
                self.dS[sl] = seg.dS
                self.P0_chord[sl] = seg.c_node[:-1]
                self.P1_chord[sl] = seg.c_node[1:]
                self.max_camber[sl] = seg.max_camber_cp
                self.max_thickness[sl] = seg.max_thickness_cp
                self.reid_corr[sl] = seg.reid_corr
                self.delta_joint[sl] = seg.delta_joint
                # dimensional blending scale (airplane.py:513): the span of the segment enters because several
                # segments may share one lifting line
                self.sigma_blend[sl] = (2.0 / (seg.b * seg.blend_dist * np.cos(seg.sweep_cp))) ** 2
                if seg.side == "left":                                         # airplane.py:518-525
                    self.PC_span_locs[sl] = from_left_tip + (1.0 - seg.cp_span_locs) * seg.b
                    node_spans = from_left_tip + (1.0 - seg.node_span_locs) * seg.b
                else:
                    self.PC_span_locs[sl] = from_left_tip + seg.cp_span_locs * seg.b
                    node_spans = from_left_tip + seg.node_span_locs * seg.b
                self.P0_span_locs[sl] = node_spans[:-1]
                self.P1_span_locs[sl] = node_spans[1:]
                self.P0[sl, :] = seg.nodes[:-1]
                self.P1[sl, :] = seg.nodes[1:]
                self.u_a[sl, :], self.u_n[sl, :], self.u_s[sl, :] = seg.u_a_cp, seg.u_n_cp, seg.u_s_cp
                self.u_a_unswept[sl, :] = seg.u_a_cp_unswept
                self.u_n_unswept[sl, :] = seg.u_n_cp_unswept
                self.u_s_unswept[sl, :] = seg.u_s_cp_unswept
                self.P0_u_a[sl, :] = seg.u_a_node[:-1]
                self.P1_u_a[sl, :] = seg.u_a_node[1:]
                k += seg.N
                from_left_tip += seg.b
            self._wing_N.append(k - wing_start)
            self.wing_slices.append(slice(wing_start, k))

        # joints of the plain lifting line, seen by other wings and other aircraft (airplane.py:555-556)
        self.P0_joint = self.P0 + (self.P0_chord * self.delta_joint * self.reid_corr)[:, np.newaxis] * self.P0_u_a
        self.P1_joint = self.P1 + (self.P1_chord * self.delta_joint * self.reid_corr)[:, np.newaxis] * self.P1_u_a
        self.PC_CG = self.PC - self.CG[np.newaxis, :]                          # airplane.py:561
        self.PC_deriv = self.u_s / np.linalg.norm(self.u_s[:, 1:], axis=1, keepdims=True)   # airplane.py:564
        self.section_sweep = -np.arctan(self.PC_deriv[:, 0])                   # airplane.py:567
        self.dl = self.P1 - self.P0                                            # airplane.py:689

    # -- reference quantities ------------------------------------------------------------------------------------------
    def _check_reference_params(self):
        if self.S_w == -1:                                                     # airplane.py:872-894
            self.S_w = 0.0
            for seg in self.wing_segments.values():
                if seg.is_main:
                    self.S_w += np.sum(seg.dS)
        if self.l_ref_lat == -1:
            self.l_ref_lat = 0.0
            for seg in self.wing_segments.values():
                if seg.is_main and seg.side == "right":
                    self.l_ref_lat += seg.b * 2.0
        try:
            if self.l_ref_lon == -1:
                self.l_ref_lon = self.S_w / self.l_ref_lat
        except Exception:
            raise IOError("No wing was specified as main for {0} so reference parameters cannot be determined.".format(self.name))

    def get_MAC(self):
        """Mean aerodynamic chord and the x location of its quarter chord (airplane.py:947-988)."""
        mac = loc = area = 0.0
        for seg in self.wing_segments.values():
            if seg.is_main:
                weight = seg.dS * np.cos(seg.dihedral_cp)
                mac += np.sum(weight * seg.c_bar_cp)
                loc += np.sum(weight * seg.control_points[:, 0])
                area += np.sum(seg.dS)
        return {"length": mac / area, "C_point": loc / area}

    # -- outside the solve path -----------------------------------------------------------------------------------------
    def _no_cad(self, *args, **kwargs):
        raise NotImplementedError("CAD export (stl/vtk/stp/dxf) is outside the lifting-line solve path this package implements.")

    export_stl = export_vtk = export_stp = export_dxf = _no_cad
