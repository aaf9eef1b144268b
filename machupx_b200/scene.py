"""``Scene`` — the drop-in boundary of the lifting-line solve path.

MachUpX has no plugin or FFI layer: its hot path is a set of private numpy methods on ``machupX.Scene``
(reference scene.py:431-1408).  This class keeps the reference's public API and JSON input format
(scene.py:37, 268, 304, 1411, 1535, 1574, 1822-2437, 2440-2986, 2989-3268, 3271, 3408, 3830, 3887) and replaces
exactly those private methods with calls into liblls (include/lls.h) through ``DeviceScene``:

    reference (numpy)                                   here (device)
    ------------------------------------------------    -----------------------------------------------
    _perform_geometry_and_atmos_calcs  scene.py:431     O(N) table upload (tables.build_tables); the N x N
                                                        geometry and V_ji_const are never materialised
    _calc_invariant_flow_properties    scene.py:557     lls_flow_state + lls_build_influence
    _solve_linear                      scene.py:800     lls_solve_linear
    _solve_nonlinear (+ residual, section lift)  :832   lls_solve_nonlinear
    _integrate_forces_and_moments      scene.py:927     lls_integrate + forces.assemble_fm (O(#segments) host)

There is no CPU fallback: without a CUDA device (or without liblls.so) ``solve_forces`` raises.  Inputs the device
section model cannot evaluate (callable airfoils, the scipy_fsolve solver) raise NotImplementedError before anything
is launched; "linear", "poly_fit" and "database" airfoils are evaluated on the device.
"""
import copy
import json
import math
import time
import warnings

import numpy as np
import scipy.interpolate as sinterp

from . import layout as L
from .airfoil import DatabaseBoundsError
from .airplane import Airplane
from .atmosphere import StandardAtmosphere
from .exceptions import MaxIterationError, SolverNotConvergedError
from .forces import AircraftFrame, assemble_fm
from .rotations import body_to_earth, earth_to_body, euler_to_quat, quat_conjugate, quat_product, quat_to_euler
from .tables import build_aircraft_state, build_tables, flap_rows
from .units import read_value, require_file

_IMPINGEMENT_TEXT = ("MachUpX detected a trailing vortex impinging upon a control point. This can lead to greatly exaggerated "
                     "induced velocities at the control point. See \"Common Issues\" in the documentation for more information. "
                     "This warning can be suppressed by reducing \"impingement_threshold\" in the solver parameters.")

_SMALL_PIVOT_TEXT = ("The unpivoted LU factorisation of the lifting-line system met a pivot more than six orders of magnitude below the "
                     "largest one: the matrix is far from the diagonally dominant form the solver relies on (numpy.linalg.solve in "
                     "the reference pivots), so the circulation and the forces may be inaccurate.")
_NAN_TEXT = ("The lifting-line solve produced NaN (a zero or NaN pivot in the LU factorisation, or a NaN residual); the reference "
             "propagates NaN silently in this situation (docs/source/common_issues.md).")

_BOUNDS_TEXT = "A poly_fit airfoil was evaluated outside the limits of its fit, or a database airfoil outside its table."

_FRAME_KEYS = {"body": ("Cx", "Cy", "Cz", "Cl", "Cm", "Cn"),
               "stab": ("Cx_s", "Cy_s", "Cz_s", "Cl_s", "Cm_s", "Cn_s"),
               "wind": ("CL", "CD", "CS", "Cl_w", "Cm_w", "Cn_w")}

# per-control-point results other code reads after a solve (scene.py:3134-3225) -> (out-table field, width)
_LAZY_OUT = {"_gamma": (L.O_GAMMA, 1), "_v_i": (L.O_VI, 3), "_alpha": (L.O_ALPHA, 1), "_CL": (L.O_CL, 1), "_CD": (L.O_CD, 1),
             "_Cm": (L.O_CM, 1), "_Re": (L.O_RE, 1), "_M": (L.O_MACH, 1), "_dF_inv": (L.O_DFINV, 3), "_dM_inv": (L.O_DMINV, 3),
             "_dF_visc": (L.O_DFVISC, 3), "_dM_visc": (L.O_DMVISC, 3), "_aL0": (L.O_AL0, 1),
             "_redim_in_plane": (L.O_REDIM_IP, 1), "_redim_full": (L.O_REDIM_FULL, 1),
             "_u_trailing_0": (L.O_UTRAIL0, 3), "_u_trailing_1": (L.O_UTRAIL1, 3), "_V_i_2": (L.O_VI2, 1)}


class Scene:
    def __init__(self, scene_input={}):
        self._airplanes = {}
        self._airplane_objects = []
        self._airplane_slices = []
        self._N = 0
        self._num_aircraft = 0
        self._solved = False
        self._device = None
        self._device_revs = None
        self._tables = None
        self._out_cache = None
        self._FM = {}
        self._last_iterations = 0
        self._last_error = 0.0
        self._load_params(scene_input)
        self.set_err_state()

    # ------------------------------------------------------------------------------------------------------
    # input
    # ------------------------------------------------------------------------------------------------------
    def _load_params(self, scene_input):
        if isinstance(scene_input, str):                                       # scene.py:60-72
            require_file(scene_input, ".json")
            with open(scene_input) as handle:
                self._input_dict = json.load(handle)
        elif isinstance(scene_input, dict):
            self._input_dict = copy.deepcopy(scene_input)
        else:
            raise IOError("Input to Scene class initializer must be a file path or Python dictionary, not type {0}."
                          .format(type(scene_input)))

        sp = self._input_dict.get("solver", {})                               # scene.py:75-87
        self._solver_type = sp.get("type", "nonlinear")
        self._solver_convergence = sp.get("convergence", 1e-10)
        self._solver_relaxation = sp.get("relaxation", 1.0)
        self._max_solver_iterations = sp.get("max_iterations", 100)
        self._use_swept_sections = sp.get("use_swept_sections", True)
        self._use_total_velocity = sp.get("use_total_velocity", True)
        self._use_in_plane = sp.get("use_in_plane", True)
        self._match_machup_pro = sp.get("match_machup_pro", False)
        self._impingement_threshold = sp.get("impingement_threshold", 1e-10)
        self._constrain_vortex_sheet = sp.get("constrain_vortex_sheet", False)
        if self._solver_type == "scipy_fsolve":
            raise NotImplementedError("solver type 'scipy_fsolve' (scene.py:671-702) runs the residual on the host; this "
                                      "package has no CPU solve path. Use 'nonlinear' or 'linear'.")

        self._unit_sys = self._input_dict.get("units", "English")
        scene_dict = self._input_dict.get("scene", {})
        atmos = scene_dict.get("atmosphere", {})
        self._std_atmos = StandardAtmosphere(unit_sys=self._unit_sys)
        self._get_density = self._initialize_density_getter(**atmos)
        self._get_wind = self._initialize_wind_getter(**atmos)
        self._get_viscosity = self._initialize_viscosity_getter(**atmos)
        self._get_sos = self._initialize_sos_getter(**atmos)

        # the scene-wide tables are built once, after the last aircraft of the input (the reference rebuilds its N x N arrays after
        # every single one, scene.py:108-111 -> 299-301; with 64 aircraft that is 64 table builds for one scene)
        aircraft = list(scene_dict.get("aircraft", {}).items())
        for k, (key, spec) in enumerate(aircraft):
            self.add_aircraft(key, spec["file"], state=spec.get("state", {}), control_state=spec.get("control_state", {}),
                              _refresh=(k == len(aircraft) - 1))

    # -- atmosphere getters (scene.py:114-265): constant, "standard", altitude profile or scattered field ------------
    def _initialize_density_getter(self, **kwargs):
        rho = read_value("rho", kwargs, self._unit_sys, self._std_atmos.rho(0.0))
        if isinstance(rho, float):
            self._constant_rho = rho
            return lambda position: self._constant_rho
        if isinstance(rho, str):
            if rho not in ["standard"]:
                raise IOError("{0} is not an allowable profile name.".format(rho))
            return lambda position: self._std_atmos.rho(-np.transpose(position)[2])
        if isinstance(rho, np.ndarray):
            self._density_data = rho
            if rho.shape[1] == 2:
                return lambda position: np.interp(-np.transpose(position)[2], self._density_data[:, 0], self._density_data[:, 1])
            if rho.shape[1] == 4:
                self._density_field_interpolator = sinterp.LinearNDInterpolator(rho[:, :3], rho[:, 3])

                def field(position):
                    # scipy answers a single point with an array of shape (1,); the reference passes that on, so that every
                    # coefficient of its result dictionary becomes a one-element array.  A single point gets a scalar here.
                    out = self._density_field_interpolator(position)
                    return out.item() if np.ndim(position) == 1 else out
                return field
        raise IOError("Density improperly specified as {0}.".format(rho))

    def _initialize_wind_getter(self, **kwargs):
        V_wind = read_value("V_wind", kwargs, self._unit_sys, [0.0, 0.0, 0.0])
        if not isinstance(V_wind, np.ndarray):
            raise IOError("Wind velocity improperly specified as {0}".format(V_wind))
        if V_wind.shape == (3,):
            self._constant_wind = V_wind
            return lambda position: self._constant_wind * np.ones(np.shape(position))
        self._wind_data = V_wind
        if V_wind.shape[1] == 6:
            xyz = V_wind[:, :3]
            self._wind_field_x_interpolator = sinterp.LinearNDInterpolator(xyz, V_wind[:, 3], fill_value=0.0)
            self._wind_field_y_interpolator = sinterp.LinearNDInterpolator(xyz, V_wind[:, 4], fill_value=0.0)
            self._wind_field_z_interpolator = sinterp.LinearNDInterpolator(xyz, V_wind[:, 5], fill_value=0.0)

            def field(position):
                comps = [f(position) for f in (self._wind_field_x_interpolator, self._wind_field_y_interpolator,
                                               self._wind_field_z_interpolator)]
                if len(np.shape(position)) == 1:
                    return np.array([c.item() for c in comps])
                return np.array(comps).T
            return field
        if V_wind.shape[1] == 4:
            def profile(position):
                h = -np.transpose(position)[2]
                comps = [np.interp(h, self._wind_data[:, 0], self._wind_data[:, c]) for c in (1, 2, 3)]
                if len(np.shape(position)) == 1:
                    return np.array([c.item() for c in comps])
                return np.array(comps).T
            return profile
        raise IOError("Wind array has the wrong number of columns.")

    def _initialize_viscosity_getter(self, **kwargs):
        nu = read_value("viscosity", kwargs, self._unit_sys, self._std_atmos.nu(0.0))
        if isinstance(nu, float):
            self._constant_nu = nu
            return lambda position: self._constant_nu * np.ones(np.shape(position)[:-1])
        if isinstance(nu, str):
            if nu not in ["standard"]:
                raise IOError("{0} is not an allowable profile name.".format(nu))
            return lambda position: self._std_atmos.nu(-np.transpose(position)[2])
        raise IOError("Viscosity improperly specified as {0}.".format(nu))

    def _initialize_sos_getter(self, **kwargs):
        a = read_value("speed_of_sound", kwargs, self._unit_sys, self._std_atmos.a(0.0))
        if isinstance(a, float):
            self._constant_a = a
            return lambda position: self._constant_a * np.ones(np.shape(position)[:-1])
        if isinstance(a, str):
            if a not in ["standard"]:
                raise IOError("{0} is not an allowable profile name.".format(a))
            return lambda position: self._std_atmos.a(-np.transpose(position)[2])
        raise IOError("Speed of sound improperly specified as {0}.".format(a))

    # ------------------------------------------------------------------------------------------------------
    # aircraft management
    # ------------------------------------------------------------------------------------------------------
    def add_aircraft(self, airplane_name, airplane_input, state={}, control_state={}, _refresh=True):
        position = np.array(state.get("position", [0.0, 0.0, 0.0]))           # scene.py:287-301
        v_wind = self._get_wind(position)
        # Re-adding a name REPLACES that aircraft (the reference's own tests do it, test_airfoil_interpolation.py:139-142).  The
        # reference then counts the replaced aircraft's control points twice (scene.py:296-297) and carries zero rows; here the
        # old aircraft is taken out of the counts first, so the scene stays solvable.
        previous = self._airplanes.pop(airplane_name, None)
        if previous is not None:
            self._N -= previous.N
            self._num_aircraft -= 1
            self._device = self._device_revs = self._batch_device = self._part_device = None
        self._airplanes[airplane_name] = Airplane(airplane_name, airplane_input, self._unit_sys, self, init_state=state,
                                                  init_control_state=control_state, v_wind=v_wind)
        self._N += self._airplanes[airplane_name].N
        self._num_aircraft += 1
        if _refresh:
            self._store_aircraft_properties()
            self._perform_geometry_and_atmos_calcs()

    def remove_aircraft(self, airplane_name):
        try:
            gone = self._airplanes.pop(airplane_name)
        except KeyError:
            raise RuntimeError("The scene has no aircraft named {0}.".format(airplane_name))
        self._N -= gone.get_num_cps()
        self._num_aircraft -= 1
        self._device = None
        self._device_revs = None
        self._batch_device = None
        self._part_device = None
        self._tables = None
        if self._num_aircraft != 0:
            self._store_aircraft_properties()
            self._perform_geometry_and_atmos_calcs()

    def _store_aircraft_properties(self):
        """State-independent per-point data in scene order (scene.py:397-428)."""
        self._airplane_objects = list(self._airplanes.values())
        self._airplane_slices = []
        k = 0
        self._c_bar, self._dS, self._section_sweep = np.zeros(self._N), np.zeros(self._N), np.zeros(self._N)
        for ap in self._airplane_objects:
            sl = slice(k, k + ap.N)
            self._airplane_slices.append(sl)
            self._c_bar[sl], self._dS[sl], self._section_sweep[sl] = ap.c_bar, ap.dS, ap.section_sweep
            k += ap.N
        if self._use_swept_sections:
            cos_sweep = np.cos(self._section_sweep)
            self._c_bar *= cos_sweep
            self._C_sweep_inv = 1.0 / cos_sweep
        self._solved = False

    def _perform_geometry_and_atmos_calcs(self):
        """Pose-dependent O(N) rows (earth-frame points and unit vectors, atmosphere at the control points) and their
        upload.  The reference builds eleven N x N arrays and V_ji_const here (scene.py:431-549); on the device that
        work is part of the influence kernel, so this is O(N)."""
        N = self._N
        self._PC, self._r_CG, self._dl = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self._u_a, self._u_n, self._u_s = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        for ap, sl in zip(self._airplane_objects, self._airplane_slices):
            self._r_CG[sl] = body_to_earth(ap.q, ap.PC_CG)
            self._PC[sl] = ap.p_bar + body_to_earth(ap.q, ap.PC)
            self._dl[sl] = body_to_earth(ap.q, ap.dl)
            ua, un, us = (ap.u_a, ap.u_n, ap.u_s) if self._use_swept_sections else (ap.u_a_unswept, ap.u_n_unswept, ap.u_s_unswept)
            self._u_a[sl], self._u_n[sl], self._u_s[sl] = body_to_earth(ap.q, ua), body_to_earth(ap.q, un), body_to_earth(ap.q, us)
        self._rho = self._get_density(self._PC)                                # scene.py:544-547
        self._nu = self._get_viscosity(self._PC)
        self._a = self._get_sos(self._PC)
        self._v_wind = np.zeros((N, 3))
        self._v_wind[:, :] = self._get_wind(self._PC)
        self._tables = build_tables(self)
        self._pose_key = self._current_pose_key()
        self._controls_key = self._current_controls_key()
        self._tables_rev = getattr(self, "_tables_rev", 0) + 1
        self._tables_full_rev = getattr(self, "_tables_full_rev", 0) + 1
        self._solved = False

    def _current_pose_key(self):
        return tuple((tuple(np.asarray(ap.q, dtype=float)), tuple(np.asarray(ap.p_bar, dtype=float))) for ap in self._airplane_objects)

    def _current_controls_key(self):
        return tuple(ap._controls_rev for ap in self._airplane_objects)

    def get_max_bound_vortex_length(self):
        return np.max(np.linalg.norm(self._dl, axis=1)).item()

    # ------------------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------------------
    def _refresh_tables(self):
        """Brings the HOST tables in line with the airplane objects and returns the table revision.  Drivers (derivatives,
        trim) change the airplane objects directly, as in the reference, so pose and control state are compared here rather
        than trusted to setters.  Every consumer of the tables (single-scene, batch and partitioned device objects) keys its
        uploads on the revision, so a control change can never leave a stale flap row on a device."""
        if self._current_pose_key() != self._pose_key:
            self._perform_geometry_and_atmos_calcs()
        elif self._current_controls_key() != self._controls_key:
            rows = flap_rows(self)
            for f, values in rows.items():
                self._tables["tab"][f, :self._N] = values
            self._controls_key = self._current_controls_key()
            self._tables_rev += 1
        return self._tables_rev

    def _sync_device(self):
        """Brings the single-scene device tables in line with the host objects and uploads the flight state."""
        from .device import DeviceScene
        self._refresh_tables()
        rev = (self._tables_rev, self._tables_full_rev)
        if self._device is None or self._device.n != self._N or self._device.n_segments != self._tables["n_segments"] \
                or self._device.n_airfoils != len(self._tables["airfoils"]) or self._device.n_aircraft != self._num_aircraft \
                or self._device.flags != self._tables["flags"]:
            self._device = DeviceScene(self._tables)
            self._device_revs = rev
        self._device_revs = self._upload_if_stale(self._device, self._device_revs)
        self._device.upload_state(build_aircraft_state(self))

    def _upload_if_stale(self, dev, have):
        """Uploads to ``dev`` (anything with upload_tables / upload_rows) whatever changed in the host tables since revision
        pair ``have`` = (revision, full revision); returns the pair now on the device."""
        want = (self._tables_rev, self._tables_full_rev)
        if have is None or have[1] != want[1]:
            dev.upload_tables(self._tables)
        elif have[0] != want[0]:
            dev.upload_rows(self._tables["tab"], L.FLAP_FIELDS)      # only control deflections changed: five table rows
        return want

    def __getattr__(self, name):
        spec = _LAZY_OUT.get(name)
        if spec is None:
            raise AttributeError(name)
        if self._out_cache is None:
            if self.__dict__.get("_device") is None:
                raise AttributeError("{0} is only available after solve_forces()".format(name))
            self._out_cache = self._device.out()
        f, width = spec
        return self._out_cache[f] if width == 1 else np.ascontiguousarray(self._out_cache[f:f + width].T)

    # ------------------------------------------------------------------------------------------------------
    # the hot entry point
    # ------------------------------------------------------------------------------------------------------
    def _get_frames(self, **kwargs):
        return kwargs.get("body_frame", True), kwargs.get("stab_frame", False), kwargs.get("wind_frame", True)

    def solve_forces(self, **kwargs):
        """Forces and moments on every aircraft at the current state (scene.py:1411-1532).  kwargs: filename,
        non_dimensional, dimensional, report_by_segment, body_frame, stab_frame, wind_frame, initial_guess, verbose."""
        if self._num_aircraft == 0:
            raise RuntimeError("There are no aircraft in this scene. No calculations can be performed.")
        if kwargs.get("partitioned", False):
            return self._solve_forces_partitioned(**kwargs)
        initial_guess = kwargs.get("initial_guess", "linear")
        verbose = kwargs.get("verbose", False)
        self._FM = {}
        self._out_cache = None
        t0 = time.time()
        self._sync_device()
        dev = self._device
        linear_time = nonlinear_time = 0.0
        status = 0
        try:
            # scene.py:557-668 (called from _solve_linear, or directly when the previous gamma seeds Newton)
            dev.flow_state()
            dev.build_influence(self._impingement_threshold)
            if self._solver_type == "linear" or (self._solver_type == "nonlinear" and initial_guess == "linear"):
                if verbose: print("Running linear solver...")
                dev.solve_linear()
                linear_time = time.time() - t0
            if self._solver_type == "nonlinear":
                if verbose: print("Running nonlinear solver...")
                t1 = time.time()
                try:
                    its, err, status, hist = dev.solve_nonlinear(self._solver_convergence, self._solver_relaxation,
                                                                 self._max_solver_iterations)
                except KeyboardInterrupt:
                    # scene.py:1486-1489: Ctrl-C during the Newton iteration moves on to the integration with the circulation
                    # reached so far.  The device loop is one library call, so the interrupt is delivered when it returns:
                    # what "so far" means here is the finished (or iteration-capped) solve.
                    print("")
                    print("!!! Nonlinear solver interrupted by Ctrl+C event. Moving on to force and moment integration...")
                    its = err = hist = None
                nonlinear_time = time.time() - t1
                self._last_iterations, self._last_error, self._error_history = its, err, hist
                if hist is not None:
                    if verbose:
                        print("{0:<20}{1:<20}".format("Iteration", "Error"))
                        print("".join(["-"] * 40))
                        for k, e in enumerate(hist):
                            print("{0:<20}{1:<20}".format(k + 1, e))
                    self._report_status(status, err)
                    if verbose: print("Nonlinear solver successfully converged. Final error: {0}".format(err))
            else:
                status = dev.status()
                self._report_status(status, 0.0)
        except Exception as e:
            self._handle_error(e)

        t2 = time.time()
        try:
            seg_fm = dev.integrate()
            self._seg_fm = seg_fm
            if dev.last_status & L.STATUS_SECTION_BOUNDS and not status & L.STATUS_SECTION_BOUNDS:
                # CD / Cm of the integration are evaluated at (alpha, Re, Mach) combinations the Newton loop never visited
                self._handle_error(DatabaseBoundsError(_BOUNDS_TEXT))
            body, stab, wind = self._get_frames(**kwargs)
            self._FM = assemble_fm(seg_fm, self._frames(), non_dimensional=kwargs.get("non_dimensional", True),
                                   dimensional=kwargs.get("dimensional", True), report_by_segment=kwargs.get("report_by_segment", False),
                                   body_frame=body, stab_frame=stab, wind_frame=wind)
        except Exception as e:
            self._handle_error(e)
        integrate_time = time.time() - t2

        if verbose:                                                            # scene.py:1506-1521
            if linear_time > 0.0:
                print("Time to compute circulation distribution using linear equations: {0} s".format(linear_time))
            if nonlinear_time > 0.0:
                print("Time to compute nonlinear improvement to circulation distribution: {0} s".format(nonlinear_time))
            total = linear_time + nonlinear_time + integrate_time
            print("Time to integrate forces: {0} s".format(integrate_time))
            print("Total time: {0} s".format(total))
            if total > 0.0:
                print("Solution rate: {0} Hz".format(1 / total))

        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(self._FM, handle, indent=4)
        self._solved = True
        return self._FM

    def _solve_forces_partitioned(self, **kwargs):
        """``solve_forces(partitioned=True)``: the rows of V_ji / A / J are dealt over the ranks of the default
        ``torch.distributed`` group (one process per GPU, NCCL), so scenes the reference cannot hold
        at all (N >= ~9k) and that exceed one GPU (N > ~72k) still solve.  Every rank must call it with the same scene;
        every rank gets the same ``FM``.  ``_gamma`` is replicated; the other per-point arrays are only filled for the rows a
        rank owns."""
        from .partitioned import PartitionedDeviceScene
        self._refresh_tables()
        dev = getattr(self, "_part_device", None)
        if dev is None or dev.n != self._N or dev.flags != self._tables["flags"] or dev.n_segments != self._tables["n_segments"]:
            dev = self._part_device = PartitionedDeviceScene(self._tables)
            self._part_revs = (self._tables_rev, self._tables_full_rev)
        self._part_revs = self._upload_if_stale(dev, self._part_revs)
        dev.upload_state(build_aircraft_state(self))
        self._FM = {}
        self._out_cache = None
        seg, its, err, status, hist = dev.solve(self._solver_type, self._solver_convergence, self._solver_relaxation,
                                                self._max_solver_iterations, self._impingement_threshold)
        self._last_iterations, self._last_error, self._error_history, self._seg_fm = its, err, hist, seg
        self._device = None
        self._out_cache = dev.d_out[:, :self._N].cpu().numpy()
        try:
            self._report_status(status, err)
        except Exception as e:
            self._handle_error(e)
        body, stab, wind = self._get_frames(**kwargs)
        self._FM = assemble_fm(seg, self._frames(), non_dimensional=kwargs.get("non_dimensional", True),
                               dimensional=kwargs.get("dimensional", True), report_by_segment=kwargs.get("report_by_segment", False),
                               body_frame=body, stab_frame=stab, wind_frame=wind)
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(self._FM, handle, indent=4)
        self._solved = True
        return self._FM

    def solve_forces_batch(self, states, aircraft=None, max_batch=4096):
        """Solves a list of flight states of ONE aircraft of the scene in a single batched device pass (SURVEY.md section 8e
        row 1; BASELINE.json configs[2]): the reference's idiom is a Python loop of set_aircraft_state + solve_forces.

        ``states`` are state dictionaries as accepted by ``set_aircraft_state``; they may change velocity / alpha / beta /
        angular rates but must keep the position and orientation the scene currently has (the geometry tables are shared
        by all cases).  Returns {"totals": {aircraft_name: {key: ndarray(ncases)}}, "iterations", "final_error",
        "converged", "status"}.  The aircraft states of the scene are left untouched."""
        from .device import BatchDeviceScene
        from .forces import batch_totals
        if self._num_aircraft == 0:
            raise RuntimeError("There are no aircraft in this scene. No calculations can be performed.")
        name = self._only_aircraft(aircraft)
        self._refresh_tables()
        return self._solve_batch_rows(self._batch_rows(name, states), max_batch)

    def _batch_rows(self, name, states):
        """(ncases, n_aircraft, AC_STRIDE) state blocks of a list of state dictionaries for airplane ``name``.  Plain
        {"position", "velocity": float, "alpha", "beta"} states (an alpha / beta / airspeed sweep, BASELINE.json configs[2]) are
        converted in one vectorised pass with the formulas of Airplane.set_state / set_aerodynamic_state; anything else goes
        through the airplane object one state at a time."""
        ap = self._airplanes[name]
        ia = [a.name for a in self._airplane_objects].index(name)
        pose = (np.array(ap.p_bar, dtype=float), np.array(ap.q, dtype=float))
        plain = self._plain_sweep_columns(states)
        if plain is not None:
            pos, speed, alpha, beta = plain
            if not (pos == pose[0][np.newaxis, :]).all() or not np.allclose(pose[1], [1.0, 0.0, 0.0, 0.0], rtol=0, atol=1e-15):
                # set_state without an "orientation" key resets it to level flight: only then is the pose kept
                raise ValueError("solve_forces_batch: every state must keep the scene's current position and orientation")
            alpha, beta = np.radians(alpha), np.radians(beta)
            C_a, S_a = np.cos(alpha), np.sin(alpha)                       # Airplane.set_aerodynamic_state
            flank = np.arctan(np.tan(beta) / C_a)
            C_B, S_B = np.cos(flank), np.sin(flank)
            scale = 1.0 / np.sqrt(1.0 - S_a * S_a * S_B * S_B)
            v_body = np.stack([speed * C_a * C_B * scale, speed * C_a * S_B * scale, speed * S_a * C_B * scale], axis=1)
            wind = np.asarray(self._get_wind(pose[0]), dtype=float)
            saved = ap.get_state()
            try:
                ap.w = np.zeros(3)                                        # no "angular_rates" key: zero
                rows = np.repeat(build_aircraft_state(self)[np.newaxis], len(states), axis=0)
            finally:
                ap.v, ap.w, ap.p_bar, ap.q = saved
            rows[:, ia, L.AC_VTRANS:L.AC_VTRANS + 3] = -(wind[np.newaxis, :] + v_body)     # level flight: earth axes = body axes
            return rows
        saved = ap.get_state()
        saved_frame = ap.angular_rate_frame
        rows = []
        try:
            for st in states:
                position = np.array(st.get("position", [0.0, 0.0, 0.0]))
                ap.set_state(**st, v_wind=self._get_wind(position))
                if not (np.allclose(ap.p_bar, pose[0], rtol=0, atol=0) and np.allclose(ap.q, pose[1], rtol=0, atol=1e-15)):
                    raise ValueError("solve_forces_batch: every state must keep the scene's current position and orientation")
                rows.append(build_aircraft_state(self))
        finally:
            ap.v, ap.w, ap.p_bar, ap.q = saved
            ap.angular_rate_frame = saved_frame
        return np.array(rows)

    @staticmethod
    def _plain_sweep_columns(states):
        """(position (n, 3), velocity, alpha, beta) of a list of PLAIN sweep states -- dictionaries with no keys but "position"
        (three numbers), "velocity" (a float), "alpha" and "beta" (numbers) -- or None if any state is anything else.  Column-wise:
        the conversion of each column to a float array is the type check (a vector velocity, a unit suffix, a string make the
        column ragged or non-numeric), which costs a millisecond for the 4,096 states of a sweep where a per-state loop of
        isinstance tests costs twenty."""
        n = len(states)
        if n == 0:
            return None
        allowed = {"position", "velocity", "alpha", "beta"}
        try:
            if not all(allowed.issuperset(st) for st in states):
                return None
            zero3, cols = [0.0, 0.0, 0.0], []
            for key, default in (("velocity", None), ("alpha", 0.0), ("beta", 0.0)):
                col = [st[key] for st in states] if default is None else [st.get(key, default) for st in states]
                arr = np.asarray(col)
                if arr.shape != (n,) or arr.dtype.kind not in "fi" or (key == "velocity" and arr.dtype.kind != "f"):
                    return None
                cols.append(arr.astype(float))
            pos = np.asarray([st.get("position", zero3) for st in states])
            if pos.shape != (n, 3) or pos.dtype.kind not in "fi":
                return None
        except (KeyError, ValueError, TypeError):
            return None
        return pos.astype(float), cols[0], cols[1], cols[2]

    def _solve_batch_rows(self, ac_states, max_batch=4096, case_tables=None):
        """Device part of a batch: ``ac_states`` (ncases, n_aircraft, AC_STRIDE) per-aircraft state blocks as build_aircraft_state
        makes them, all for the geometry the scene currently has -- or, with ``case_tables`` (ncases, NF, ld), every case with
        its own copy of the double table (control deflections / pose differ between the cases)."""
        from .device import BatchDeviceScene
        from .forces import batch_totals
        n_total = len(ac_states)
        info = []
        seg0 = 0
        for a in self._airplane_objects:
            info.append({"name": a.name, "q": a.q, "wind": self._get_wind(a.p_bar), "rho": self._get_density(a.p_bar), "S_w": a.S_w,
                         "l_ref_lat": a.l_ref_lat, "l_ref_lon": a.l_ref_lon, "seg_range": (seg0, seg0 + len(a.segments))})
            seg0 += len(a.segments)
        seg_fm = np.zeros((n_total, self._tables["n_segments"], 12))
        its, err, status = np.zeros(n_total, dtype=int), np.zeros(n_total), np.zeros(n_total, dtype=int)
        for lo in range(0, n_total, max_batch):
            hi = min(lo + max_batch, n_total)
            nc = hi - lo
            dev = getattr(self, "_batch_device", None)
            if dev is None or dev.ncases != nc or dev.n != self._N or dev.flags != self._tables["flags"] \
                    or dev.n_segments != self._tables["n_segments"] or dev.n_aircraft != self._num_aircraft:
                dev = self._batch_device = BatchDeviceScene(self._tables, nc)
                self._batch_revs = (self._tables_rev, self._tables_full_rev)
            if case_tables is None:
                self._batch_revs = self._upload_if_stale(dev, self._batch_revs)
                dev.use_shared_tables()
            else:
                dev.upload_case_tables(case_tables[lo:hi])
            dev.upload_states(ac_states[lo:hi])
            seg_fm[lo:hi], its[lo:hi], err[lo:hi], status[lo:hi] = dev.solve(
                self._solver_type, self._solver_convergence, self._solver_relaxation, self._max_solver_iterations,
                self._impingement_threshold)
        converged = (status & L.STATUS_NOT_CONVERGED) == 0
        try:
            # one report for the whole batch: the union of the cases' status bits, the largest final error of the failed cases
            self._report_status(int(np.bitwise_or.reduce(status)) if n_total else 0,
                                float(err[~converged].max()) if not converged.all() else 0.0)
        except Exception as e:
            self._handle_error(e)
        return {"totals": batch_totals(seg_fm, ac_states, info) if case_tables is None else None, "iterations": its, "final_error": err, "converged": converged,
                "status": status, "seg_fm": seg_fm}

    def _report_status(self, status, err):
        """Device status bits -> the reference's warnings and exceptions (scene.py:622-623, 905-908; wing_segment.py:1087), routed
        through the error policy by the caller.  Bounds errors come first: the reference raises them from inside the residual."""
        status = int(status)
        if status & L.STATUS_IMPINGEMENT:
            warnings.warn(_IMPINGEMENT_TEXT)
        if status & L.STATUS_SMALL_PIVOT:
            warnings.warn(_SMALL_PIVOT_TEXT)
        if status & L.STATUS_NAN:
            warnings.warn(_NAN_TEXT)
        if status & L.STATUS_SECTION_BOUNDS:
            self._handle_error(DatabaseBoundsError(_BOUNDS_TEXT))
        if status & L.STATUS_NOT_CONVERGED:
            raise SolverNotConvergedError(self._solver_type, err)

    def _frames(self):
        return [AircraftFrame(ap.name, ap.q, ap.v, self._get_wind(ap.p_bar), self._get_density(ap.p_bar), ap.S_w,
                              ap.l_ref_lat, ap.l_ref_lon, [seg.name for seg in ap.segments]) for ap in self._airplane_objects]

    # ------------------------------------------------------------------------------------------------------
    # state setters
    # ------------------------------------------------------------------------------------------------------
    def _only_aircraft(self, aircraft):
        if aircraft is None:
            if self._num_aircraft == 1:
                return list(self._airplanes.keys())[0]
            raise IOError("Aircraft name must be specified if there is more than one aircraft in the scene.")
        return aircraft

    def set_aircraft_state(self, state={}, aircraft=None):
        aircraft = self._only_aircraft(aircraft)                               # scene.py:1535-1571
        position = np.array(state.get("position", [0.0, 0.0, 0.0]))
        ap = self._airplanes[aircraft]
        old_p, old_q = ap.p_bar, ap.q
        ap.set_state(**state, v_wind=self._get_wind(position))
        if not np.allclose(old_p, position) or not np.allclose(old_q, ap.q):
            self._perform_geometry_and_atmos_calcs()
        self._solved = False

    def set_aircraft_control_state(self, control_state={}, aircraft=None):
        aircraft = self._only_aircraft(aircraft)                               # scene.py:1574-1598
        self._airplanes[aircraft].set_control_state(control_state)
        self._solved = False

    def get_aircraft_reference_geometry(self, aircraft=None):
        ap = self._airplanes[self._only_aircraft(aircraft)]
        return ap.S_w, ap.l_ref_lon, ap.l_ref_lat

    def _get_aircraft(self, **kwargs):
        aircraft = kwargs.get("aircraft", None)                                # scene.py:3769-3789
        if aircraft is None:
            return list(self._airplanes.keys())
        if isinstance(aircraft, list):
            return copy.copy(aircraft)
        if isinstance(aircraft, str):
            return [aircraft]
        raise IOError("{0} is not an allowable aircraft name specification.".format(aircraft))

    # ------------------------------------------------------------------------------------------------------
    # error policy
    # ------------------------------------------------------------------------------------------------------
    def set_err_state(self, **kwargs):
        self._err_state = {"not_converged": kwargs.get("not_converged", "raise"),
                           "database_bounds": kwargs.get("database_bounds", "raise")}
        for ap in self._airplanes.values():
            for airfoil in ap._airfoil_database.values():
                airfoil.set_err_state(**kwargs)

    def _handle_error(self, error):
        if isinstance(error, SolverNotConvergedError):                         # scene.py:3864-3884
            key = "not_converged"
        elif isinstance(error, DatabaseBoundsError):
            key = "database_bounds"
        else:
            raise error
        instruction = self._err_state[key]
        if instruction == "raise":
            raise error
        if instruction == "warn":
            warnings.warn(str(error))
        elif instruction != "ignore":
            raise RuntimeError("MachUpX got an incorrect error handling instruction. '{0}' is invalid.".format(instruction))

    # ------------------------------------------------------------------------------------------------------
    # finite-difference drivers (central differences around the current state; each stencil point is one solve)
    # ------------------------------------------------------------------------------------------------------
    def _frame_keys(self, **kwargs):
        body, stab, wind = self._get_frames(**kwargs)
        keys = []
        if body: keys += _FRAME_KEYS["body"]
        if stab: keys += _FRAME_KEYS["stab"]
        if wind: keys += _FRAME_KEYS["wind"]
        return keys, wind

    def derivatives(self, **kwargs):
        derivs = {}                                                            # scene.py:1822-1873
        for name in self._get_aircraft(**kwargs):
            sub = dict(kwargs)
            sub["aircraft"] = name
            sub.pop("filename", None)
            derivs[name] = {"stability": self.stability_derivatives(**sub)[name],
                            "damping": self.damping_derivatives(**sub)[name],
                            "control": self.control_derivatives(**sub)[name]}
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(derivs, handle, indent=4)
        return derivs

    def _solve_nd(self, **kwargs):
        kw = {k: v for k, v in kwargs.items() if k not in ("dimensional", "filename")}
        self.solve_forces(dimensional=False, **kw)
        return self._FM

    def _stencil_totals(self, name, setters, opts=None):
        """Totals of a finite-difference stencil as ONE batched device pass.  Every callable of
        ``setters`` perturbs the airplane ``name`` starting from its current state: velocity, angular rates, control deflections,
        position or orientation.  Cases that only change the flight state share the geometry tables; as soon as one case changes
        control deflections or pose, every case gets its own copy of the double table (lls_set_case_tables).  Returns one
        {key: float} dictionary of aircraft totals per case, laid out like ``_FM[name]["total"]`` for the frames / dimensional
        switches in ``opts``.  The airplane and the host tables are left as they were."""
        ap = self._airplanes[name]
        self._refresh_tables()
        base_revs = (self._tables_rev, self._tables_full_rev)
        base_tab = self._tables["tab"]
        saved = tuple(np.copy(x) for x in ap.get_state())
        saved_controls = copy.deepcopy(ap.current_control_state)
        saved_frame = ap.angular_rate_frame

        def restore():
            ap.v, ap.w, ap.p_bar, ap.q = (np.copy(x) for x in saved)
            ap.angular_rate_frame = saved_frame
            if ap.current_control_state != saved_controls:
                ap.set_control_state(saved_controls)

        rows, tabs, frames = [], [], []
        if not setters:          # e.g. the control stencil of an aircraft without controls: nothing to solve
            return []
        try:
            for perturb in setters:
                restore()
                perturb(ap)
                self._refresh_tables()
                rows.append(build_aircraft_state(self))
                changed = (self._tables_rev, self._tables_full_rev) != base_revs or self._tables["tab"] is not base_tab
                tabs.append(self._tables["tab"].copy() if changed else None)
                frames.append(self._frames())
        finally:
            restore()
            self._refresh_tables()
        if any(t is not None for t in tabs):
            here = self._tables["tab"]
            tabs = np.array([here if t is None else t for t in tabs])
        else:
            tabs = None
        res = self._solve_batch_rows(np.array(rows), case_tables=tabs)
        kwargs = dict(opts or {})
        body, stab, wind = self._get_frames(**kwargs)
        out = []
        for i in range(len(setters)):
            fm = assemble_fm(res["seg_fm"][i], frames[i], non_dimensional=kwargs.get("non_dimensional", True),
                             dimensional=kwargs.get("dimensional", True), body_frame=body, stab_frame=stab, wind_frame=wind)
            out.append(fm[name]["total"])
        return out

    def _batched_ok(self, kwargs):
        """``batched=True``: the driver's finite-difference stencil runs as one batched device pass instead of one solve per
        stencil point."""
        return bool(kwargs.get("batched", False))

    def stability_derivatives(self, **kwargs):
        derivs = {}                                                            # scene.py:1876-1994
        keys, wind = self._frame_keys(**kwargs)
        dtheta = kwargs.get("dtheta", 0.5)
        batched = self._batched_ok(kwargs)
        kwargs = {k: v for k, v in kwargs.items() if k != "batched"}
        for name in self._get_aircraft(**kwargs):
            ap = self._airplanes[name]
            alpha_0, beta_0, _ = ap.get_aerodynamic_state()
            if batched:
                a_fwd, a_bwd, b_fwd, b_bwd = self._stencil_totals(name, [
                    lambda a: a.set_aerodynamic_state(alpha=alpha_0 + dtheta),
                    lambda a: a.set_aerodynamic_state(alpha=alpha_0 - dtheta),
                    lambda a: a.set_aerodynamic_state(alpha=alpha_0, beta=beta_0 + dtheta),
                    lambda a: a.set_aerodynamic_state(alpha=alpha_0, beta=beta_0 - dtheta)], dict(kwargs, dimensional=False))
                scale = 1 / (2 * np.radians(dtheta))
                d = {}
                for k in keys:
                    d[k + ",a"] = (a_fwd[k] - a_bwd[k]) * scale
                    d[k + ",b"] = (b_fwd[k] - b_bwd[k]) * scale
                if wind:
                    d["%_static_margin"] = -d["Cm_w,a"] / d["CL,a"] * 100.0
                derivs[name] = d
                continue
            ap.set_aerodynamic_state(alpha=alpha_0 + dtheta)
            a_fwd = self._solve_nd(**kwargs)[name]["total"]
            ap.set_aerodynamic_state(alpha=alpha_0 - dtheta)
            a_bwd = self._solve_nd(**kwargs)[name]["total"]
            ap.set_aerodynamic_state(alpha=alpha_0, beta=beta_0 + dtheta)
            b_fwd = self._solve_nd(**kwargs)[name]["total"]
            ap.set_aerodynamic_state(beta=beta_0 - dtheta)
            b_bwd = self._solve_nd(**kwargs)[name]["total"]
            scale = 1 / (2 * np.radians(dtheta))
            d = {}
            for k in keys:
                d[k + ",a"] = (a_fwd[k] - a_bwd[k]) * scale
                d[k + ",b"] = (b_fwd[k] - b_bwd[k]) * scale
            if wind:
                d["%_static_margin"] = -d["Cm_w,a"] / d["CL,a"] * 100.0
            derivs[name] = d
            ap.set_aerodynamic_state(alpha=alpha_0, beta=beta_0)
            self._solved = False
        return derivs

    def damping_derivatives(self, **kwargs):
        derivs = {}                                                            # scene.py:1997-2177
        keys, _ = self._frame_keys(**kwargs)
        step = kwargs.get("dtheta_dot", 0.005)
        batched = self._batched_ok(kwargs)
        kwargs = {k: v for k, v in kwargs.items() if k != "batched"}
        for name in self._get_aircraft(**kwargs):
            ap = self._airplanes[name]
            _, _, vel_0 = ap.get_aerodynamic_state()
            omega_0 = ap.w
            perts = [np.array([step, 0.0, 0.0]), np.array([0.0, step, 0.0]), np.array([0.0, 0.0, step])]
            if ap.angular_rate_frame == "stab":
                perts = [body_to_earth(ap.q_to_stab, p) for p in perts]
            elif ap.angular_rate_frame == "wind":
                perts = [body_to_earth(ap.q_to_wind, p) for p in perts]
            pairs = []
            if batched:
                def rate(dw):
                    return lambda a: setattr(a, "w", omega_0 + dw)
                tot = self._stencil_totals(name, [rate(sgn * p) for p in perts for sgn in (1.0, -1.0)], dict(kwargs, dimensional=False))
                pairs = [(tot[2 * i], tot[2 * i + 1]) for i in range(3)]
            else:
                for p in perts:
                    ap.w = omega_0 + p
                    fwd = self._solve_nd(**kwargs)[name]["total"]
                    ap.w = omega_0 - p
                    bwd = self._solve_nd(**kwargs)[name]["total"]
                    pairs.append((fwd, bwd))
                ap.w = omega_0
                self._solved = False
            _, c, b = self.get_aircraft_reference_geometry(aircraft=name)
            lat, lon = 2 * vel_0 / b, 2 * vel_0 / c
            dx_inv = 1 / (2 * step)
            d = {}
            for (fwd, bwd), tag, nd in zip(pairs, ("pbar", "qbar", "rbar"), (lat, lon, lat)):
                for k in keys:
                    d[k + "," + tag] = (fwd[k] - bwd[k]) * dx_inv * nd
            derivs[name] = d
        return derivs

    def control_derivatives(self, **kwargs):
        derivs = {}                                                            # scene.py:2180-2273
        keys, _ = self._frame_keys(**kwargs)
        dtheta = kwargs.get("dtheta", 0.5)
        batched = self._batched_ok(kwargs)
        kwargs = {k: v for k, v in kwargs.items() if k != "batched"}
        for name in self._get_aircraft(**kwargs):
            ap = self._airplanes[name]
            current = copy.deepcopy(ap.current_control_state)
            pert = copy.deepcopy(current)
            d = {}
            if batched:      # 2 x n_controls cases, each with its own flap rows, in one device pass
                def deflect(control, delta):
                    def perturb(a):
                        moved = copy.deepcopy(current)
                        moved[control] = current.get(control, 0.0) + delta
                        a.set_control_state(control_state=moved)
                    return perturb
                tot = self._stencil_totals(name, [deflect(c, sgn * dtheta) for c in ap.control_names for sgn in (1.0, -1.0)],
                                           dict(kwargs, dimensional=False))
                diff = 2 * np.radians(dtheta)
                for i, control in enumerate(ap.control_names):
                    for k in keys:
                        d[k + ",d" + control] = (tot[2 * i][k] - tot[2 * i + 1][k]) / diff
                derivs[name] = d
                continue
            for control in ap.control_names:
                value = current.get(control, 0.0)
                pert[control] = value + dtheta
                ap.set_control_state(control_state=pert)
                fwd = self._solve_nd(**kwargs)[name]["total"]
                pert[control] = value - dtheta
                ap.set_control_state(control_state=pert)
                bwd = self._solve_nd(**kwargs)[name]["total"]
                pert[control] = value
                ap.set_control_state(control_state=pert)
                self._solved = False
                diff = 2 * np.radians(dtheta)
                for k in keys:
                    d[k + ",d" + control] = (fwd[k] - bwd[k]) / diff
            derivs[name] = d
        return derivs

    def state_derivatives(self, **kwargs):
        derivs = {}                                                            # scene.py:2276-2369
        dx, dV = kwargs.get("dx", 0.5), kwargs.get("dV", 0.5)
        de, dw = kwargs.get("de", 0.001), kwargs.get("dw", 0.01)
        plan = [("velocity", "u", 0, dV), ("velocity", "v", 1, dV), ("velocity", "w", 2, dV),
                ("position", "x_f", 0, dx), ("position", "y_f", 1, dx), ("position", "z_f", 2, dx),
                ("angular_rates", "p", 0, dw), ("angular_rates", "q", 1, dw), ("angular_rates", "r", 2, dw),
                ("orientation", "qx", 0, de), ("orientation", "qy", 1, de), ("orientation", "qz", 2, de)]
        batched = self._batched_ok(kwargs)
        kwargs = {k: v for k, v in kwargs.items() if k != "batched"}
        for name in self._get_aircraft(**kwargs):
            ap = self._airplanes[name]
            v0, w0, p0, q0 = ap.get_state()
            orig = {"position": np.asarray(p0, dtype=float), "velocity": earth_to_body(q0, v0 - self._get_wind(p0)),
                    "orientation": q0, "angular_rates": np.asarray(w0, dtype=float)}
            d = {}
            # The reference refreshes the scene geometry only for the position and orientation perturbations (scene.py:2383,
            # 2390): its angular-rate perturbations, which come right after the backward z_f one, are solved with the geometry and
            # the atmosphere at the control points of THAT position (1.5e-5 relative in a standard atmosphere).  This package always
            # solves the state it is given, so it is given that position for those six solves and the numbers come out the same.
            stale = copy.deepcopy(orig)
            stale["position"][2] += dx
            stale["position"][2] -= 2.0 * dx
            base = {"angular_rates": stale}
            if batched:      # the 24 stencil points (12 of them move or rotate the aircraft: per-case tables) in one device pass
                setters = []
                for variable, tag, index, step in plan:
                    setters += self._state_stencil(variable, index, step, base.get(variable, orig))
                tot = self._stencil_totals(name, setters, {k: v for k, v in kwargs.items() if k != "filename"})
                for i, (variable, tag, index, step) in enumerate(plan):
                    f, b = tot[2 * i], tot[2 * i + 1]
                    d.update({"d{0},d{1}".format(k, tag): (f[k] - b[k]) * 0.5 / step for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz")})
                derivs[name] = d
                continue
            for variable, tag, index, step in plan:
                d.update(self._determine_state_derivs(variable, tag, index, step, base.get(variable, orig), name, **kwargs))
            derivs[name] = d
            ap.set_state(**orig)
            self._solved = False
        return derivs

    @staticmethod
    def _state_stencil(variable, index, perturbation, orig_state):
        """The forward and backward perturbation of one state variable as setters for ``_stencil_totals``: exactly the states
        ``_determine_state_derivs`` solves one after the other (scene.py:2372-2437)."""
        def setter(pert):
            return lambda a: a.set_state(**pert)
        fwd, bwd = copy.deepcopy(orig_state), copy.deepcopy(orig_state)
        if variable in ("position", "velocity", "angular_rates"):
            fwd[variable][index] += perturbation
            bwd[variable][index] += perturbation
            bwd[variable][index] -= 2.0 * perturbation
        else:
            dq = np.array([1.0, 0.0, 0.0, 0.0])
            dq[index + 1] = 0.5 * perturbation
            dq = dq / np.linalg.norm(dq)
            q0, v0 = orig_state["orientation"], orig_state["velocity"]
            fwd["orientation"], fwd["velocity"] = quat_product(q0, dq), earth_to_body(dq, v0)
            bwd["orientation"], bwd["velocity"] = quat_product(q0, quat_conjugate(dq)), body_to_earth(dq, v0)
        return [setter(fwd), setter(bwd)]

    def _determine_state_derivs(self, variable, tag, index, perturbation, orig_state, aircraft_name, **kwargs):
        ap = self._airplanes[aircraft_name]                                    # scene.py:2372-2437
        kw = {k: v for k, v in kwargs.items() if k != "filename"}
        pert = copy.deepcopy(orig_state)
        if variable in ("position", "velocity", "angular_rates"):
            pert[variable][index] += perturbation
            ap.set_state(**pert)
            fwd = copy.deepcopy(self.solve_forces(**kw))
            pert[variable][index] -= 2.0 * perturbation
            ap.set_state(**pert)
            bwd = copy.deepcopy(self.solve_forces(**kw))
        else:   # rotate by a small quaternion, keeping the earth-fixed velocity
            dq = np.array([1.0, 0.0, 0.0, 0.0])
            dq[index + 1] = 0.5 * perturbation
            dq = dq / np.linalg.norm(dq)
            q0, v0 = pert["orientation"], pert["velocity"]
            pert["orientation"], pert["velocity"] = quat_product(q0, dq), earth_to_body(dq, v0)
            ap.set_state(**pert)
            fwd = copy.deepcopy(self.solve_forces(**kw))
            pert["orientation"], pert["velocity"] = quat_product(q0, quat_conjugate(dq)), body_to_earth(dq, v0)
            ap.set_state(**pert)
            bwd = copy.deepcopy(self.solve_forces(**kw))
        scale = 0.5 / perturbation
        f, b = fwd[aircraft_name]["total"], bwd[aircraft_name]["total"]
        return {"d{0},d{1}".format(k, tag): (f[k] - b[k]) * scale for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz")}

    # ------------------------------------------------------------------------------------------------------
    # trim and friends
    # ------------------------------------------------------------------------------------------------------
    def _get_aircraft_q_inf(self, aircraft_name):
        ap = self._airplanes[aircraft_name]                                    # scene.py:2654-2660
        rho = self._get_density(ap.p_bar)
        V = np.linalg.norm(ap.v - self._get_wind(ap.p_bar))
        return 0.5 * rho * V * V

    def _get_aircraft_pitch_trim_residuals(self, aircraft_name, CL_des, Cm_des):
        tot = self.solve_forces(dimensional=False)[aircraft_name]["total"]     # scene.py:2643-2651
        return np.array([tot["CL"] - CL_des, tot["Cm"] - Cm_des])

    def _trim_target(self, kwargs):
        if len(self._airplanes) == 1:
            name = list(self._airplanes.keys())[0]
        else:
            name = kwargs.get("aircraft")
        ap = self._airplanes[name]
        CW = ap.W / (self._get_aircraft_q_inf(name) * ap.S_w)
        control = kwargs.get("pitch_control", "elevator")
        if control not in ap.current_control_state:
            raise IOError("{0} has no control named {1}. Cannot be trimmed in pitch.".format(name, control))
        return name, ap, kwargs.get("CL", CW), kwargs.get("Cm", 0.0), control

    def pitch_trim(self, **kwargs):
        """Angle of attack and pitch-control deflection for CL = CW (or given) and Cm = 0, by Newton iteration on a
        finite-difference 2 x 2 Jacobian (scene.py:2440-2640; the Jacobian mixes Cm_w with the Cm residual, and an
        unused stability_derivatives() evaluation per iteration is part of the reference's observable state changes)."""
        verbose = kwargs.get("verbose", False)
        name, ap, CL_des, Cm_des, control = self._trim_target(kwargs)
        v_wind = self._get_wind(ap.p_bar)
        alpha_original, _, _ = ap.get_aerodynamic_state(v_wind=v_wind)
        controls_original = copy.copy(ap.current_control_state)
        flap = copy.copy(controls_original[control])
        max_iter, relax = kwargs.get("max_iterations", 100), kwargs.get("relaxation", 1.0)
        if verbose:
            print("\nTrimming...\nTrimming {0} using {1}.".format(name, control))
            print("{0:<20}{1:<20}{2:<25}{3:<25}".format("Alpha", control, "Lift Residual", "Moment Residual"))
        R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
        controls = copy.copy(controls_original)
        alpha = copy.copy(alpha_original)
        if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(alpha, flap, R[0], R[1]))
        opts = dict(dimensional=False, wind_frame=True, body_frame=False, stab_frame=False)
        J = np.zeros((2, 2))
        it = 0
        alpha_new, flap_new = alpha, flap
        batched = self._batched_ok(kwargs)
        while (abs(R) > 1e-10).any():
            step = 0.5
            here = copy.deepcopy(ap.current_control_state)
            pert = copy.deepcopy(here)
            value = here.get(control, 0.0)
            if batched:      # the four Jacobian solves of this Newton step in one device pass
                alpha_0, beta_0, _ = ap.get_aerodynamic_state()

                def deflect(delta):
                    def perturb(a):
                        moved = copy.deepcopy(here)
                        moved[control] = value + delta
                        a.set_control_state(control_state=moved)
                    return perturb
                c_f, c_b, a_f, a_b = self._stencil_totals(name, [deflect(step), deflect(-step),
                                                                 lambda a: a.set_aerodynamic_state(alpha=alpha_0 + step),
                                                                 lambda a: a.set_aerodynamic_state(alpha=alpha_0 - step)], opts)
                diff = 2 * np.radians(step)
                J[0, 1], J[1, 1] = (c_f["CL"] - c_b["CL"]) / diff, (c_f["Cm_w"] - c_b["Cm_w"]) / diff
                J[0, 0], J[1, 0] = (a_f["CL"] - a_b["CL"]) / diff, (a_f["Cm_w"] - a_b["Cm_w"]) / diff
                delta = np.linalg.solve(J, -R)
                alpha_new = alpha + np.degrees(delta[0]) * relax
                ap.set_aerodynamic_state(alpha=alpha_new)
                flap_new = flap + np.degrees(delta[1]) * relax
                controls[control] = flap_new
                ap.set_control_state(controls)
                alpha, flap = alpha_new, flap_new
                R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
                if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(alpha, flap, R[0], R[1]))
                it += 1
                if it == max_iter:
                    raise MaxIterationError("pitch_trim", np.max(np.abs(R)))
                continue
            pert[control] = value + step
            ap.set_control_state(control_state=pert)
            fwd = self.solve_forces(**opts)[name]["total"]
            pert[control] = value - step
            ap.set_control_state(control_state=pert)
            bwd = self.solve_forces(**opts)[name]["total"]
            pert[control] = value
            ap.set_control_state(control_state=pert)
            self._solved = False
            diff = 2 * np.radians(step)
            J[0, 1] = (fwd["CL"] - bwd["CL"]) / diff
            J[1, 1] = (fwd["Cm_w"] - bwd["Cm_w"]) / diff

            alpha_0, beta_0, _ = ap.get_aerodynamic_state()
            ap.set_aerodynamic_state(alpha=alpha_0 + step)
            fwd = self.solve_forces(**opts)[name]["total"]
            ap.set_aerodynamic_state(alpha=alpha_0 - step)
            bwd = self.solve_forces(**opts)[name]["total"]
            J[0, 0] = (fwd["CL"] - bwd["CL"]) / diff
            J[1, 0] = (fwd["Cm_w"] - bwd["Cm_w"]) / diff
            ap.set_aerodynamic_state(alpha=alpha_0, beta=beta_0)
            self._solved = False

            delta = np.linalg.solve(J, -R)
            alpha_new = alpha + np.degrees(delta[0]) * relax
            ap.set_aerodynamic_state(alpha=alpha_new)
            flap_new = flap + np.degrees(delta[1]) * relax
            controls[control] = flap_new
            ap.set_control_state(controls)
            alpha, flap = alpha_new, flap_new
            R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
            if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(alpha, flap, R[0], R[1]))
            it += 1
            if it == max_iter:
                raise MaxIterationError("pitch_trim", np.max(np.abs(R)))
        trim = {name: {"alpha": alpha_new, control: flap_new}}
        if kwargs.get("set_trim_state", True):
            ap.set_aerodynamic_state(alpha=alpha_new)
            self.set_aircraft_control_state({control: flap_new}, aircraft=name)
        else:
            ap.set_aerodynamic_state(alpha=alpha_original)
            self.set_aircraft_control_state(controls_original, aircraft=name)
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(trim, handle, indent=4)
        return trim

    def pitch_trim_using_orientation(self, **kwargs):
        """Trim by elevation angle and pitch control at constant earth-fixed velocity (scene.py:2663-2888)."""
        verbose = kwargs.get("verbose", False)
        name, ap, CL_des, Cm_des, control = self._trim_target(kwargs)
        v_orig, w_orig, p_orig, q_orig = ap.get_state()
        phi, theta_orig, psi = quat_to_euler(q_orig)
        v_wind = self._get_wind(ap.p_bar)
        controls_original = copy.copy(ap.current_control_state)
        flap = copy.copy(controls_original[control])
        curr_state = {"position": list(p_orig), "velocity": list(earth_to_body(q_orig, v_orig - v_wind)),
                      "orientation": list(q_orig), "angular_rates": list(w_orig)}
        max_iter, relax = kwargs.get("max_iterations", 100), kwargs.get("relaxation", 1.0)
        if verbose:
            print("Trimming {0} using {1}.".format(name, control))
            print("{0:<20}{1:<20}{2:<25}{3:<25}".format("Elevation", control, "Lift Residual", "Moment Residual"))
        R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
        theta = copy.copy(theta_orig)
        if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(math.degrees(theta), flap, R[0], R[1]))

        def pose(th):
            q = euler_to_quat([phi, th, psi])
            return q, earth_to_body(q, v_orig - v_wind)

        it = 0
        J = np.zeros((2, 2))
        pert = copy.copy(controls_original)
        batched = self._batched_ok(kwargs)
        while (abs(R) > 1e-10).any():
            step = 0.001
            if batched:      # two control and two elevation-angle perturbations (the latter rotate the aircraft) in one device pass
                def deflect(delta, base=copy.copy(pert), value=flap):
                    def perturb(a):
                        moved = copy.copy(base)
                        moved[control] = value + delta
                        a.set_control_state(control_state=moved)
                    return perturb

                def elevate(th):
                    q_t, v_t = pose(th)
                    return lambda a: a.set_state(position=p_orig, velocity=v_t, orientation=q_t, angular_rates=w_orig)
                c_f, c_b, t_f, t_b = self._stencil_totals(name, [deflect(step), deflect(-step), elevate(theta + step),
                                                                 elevate(theta - step)], dict(dimensional=False))
                J[0, 1], J[1, 1] = (c_f["CL"] - c_b["CL"]) / (2.0 * step), (c_f["Cm"] - c_b["Cm"]) / (2.0 * step)
                J[0, 0], J[1, 0] = (t_f["CL"] - t_b["CL"]) / (2.0 * step), (t_f["Cm"] - t_b["Cm"]) / (2.0 * step)
                delta = np.linalg.solve(J, -R)
                theta += delta[0] * relax
                flap += delta[1] * relax
                q_n, v_n = pose(theta)
                curr_state = {"position": list(p_orig), "velocity": list(v_n), "orientation": list(q_n), "angular_rates": list(w_orig)}
                ap.set_state(**curr_state)
                pert[control] = flap
                ap.set_control_state(pert)
                R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
                if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(math.degrees(theta), flap, R[0], R[1]))
                it += 1
                if it == max_iter:
                    raise MaxIterationError("pitch_trim_using_orientation", np.max(np.abs(R)))
                continue
            pert[control] = flap + step
            ap.set_control_state(control_state=pert)
            fwd = self.solve_forces(dimensional=False)[name]["total"]
            pert[control] = flap - step
            ap.set_control_state(control_state=pert)
            bwd = self.solve_forces(dimensional=False)[name]["total"]
            pert[control] = flap
            ap.set_control_state(control_state=pert)
            self._solved = False
            J[0, 1] = (fwd["CL"] - bwd["CL"]) / (2.0 * step)
            J[1, 1] = (fwd["Cm"] - bwd["Cm"]) / (2.0 * step)

            q_f, v_f = pose(theta + step)
            ap.set_state(position=p_orig, velocity=v_f, orientation=q_f, angular_rates=w_orig)
            fwd = self.solve_forces(dimensional=False)[name]["total"]
            q_b, v_b = pose(theta - step)
            ap.set_state(position=p_orig, velocity=v_b, orientation=q_b, angular_rates=w_orig)
            bwd = self.solve_forces(dimensional=False)[name]["total"]
            J[0, 0] = (fwd["CL"] - bwd["CL"]) / (2.0 * step)
            J[1, 0] = (fwd["Cm"] - bwd["Cm"]) / (2.0 * step)

            delta = np.linalg.solve(J, -R)
            theta += delta[0] * relax
            flap += delta[1] * relax
            q_n, v_n = pose(theta)
            curr_state = {"position": list(p_orig), "velocity": list(v_n), "orientation": list(q_n), "angular_rates": list(w_orig)}
            ap.set_state(**curr_state)
            pert[control] = flap
            ap.set_control_state(pert)
            R = self._get_aircraft_pitch_trim_residuals(name, CL_des, Cm_des)
            if verbose: print("{0:<20}{1:<20}{2:<25}{3:<25}".format(math.degrees(theta), flap, R[0], R[1]))
            it += 1
            if it == max_iter:
                raise MaxIterationError("pitch_trim_using_orientation", np.max(np.abs(R)))
        if not kwargs.get("set_trim_state", True):
            ap.set_state(position=p_orig, velocity=earth_to_body(q_orig, v_orig - v_wind), orientation=q_orig, angular_rates=w_orig)
            self.set_aircraft_control_state(controls_original, aircraft=name)
            self._solved = False
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(curr_state, handle, indent=4)
                json.dump(pert, handle, indent=4)
        return curr_state, pert

    def aero_center(self, **kwargs):
        """Aerodynamic centre from first and second alpha-derivatives of Cx, Cz, Cm (Mechanics of Flight eqs. 4.8.29-31;
        scene.py:2891-2986)."""
        out = {}
        verbose = kwargs.get("verbose", False)
        for name in self._get_aircraft(**kwargs):
            if verbose: print("Calculating the aerodynamic center for {0}...".format(name))
            ap = self._airplanes[name]
            v_wind = self._get_wind(ap.p_bar)
            h = 0.5
            if self._batched_ok(kwargs):      # the three angles of attack of the second-difference stencil in one device pass
                a0, B0, V0 = ap.get_aerodynamic_state(v_wind=v_wind)
                mid, lo, hi = self._stencil_totals(name, [
                    lambda a: None,
                    lambda a: a.set_aerodynamic_state(alpha=a0 - h, beta=B0, velocity=V0, v_wind=v_wind),
                    lambda a: a.set_aerodynamic_state(alpha=a0 + h, beta=B0, velocity=V0, v_wind=v_wind)], dict(dimensional=False))
            else:
                mid = self.solve_forces(dimensional=False)[name]["total"]
                a0, B0, V0 = ap.get_aerodynamic_state(v_wind=v_wind)
                ap.set_aerodynamic_state(alpha=a0 - h, beta=B0, velocity=V0, v_wind=v_wind)
                lo = self.solve_forces(dimensional=False)[name]["total"]
                ap.set_aerodynamic_state(alpha=a0 + h, beta=B0, velocity=V0, v_wind=v_wind)
                hi = self.solve_forces(dimensional=False)[name]["total"]
                ap.set_aerodynamic_state(alpha=a0, beta=B0, velocity=V0, v_wind=v_wind)
                self._solved = False
            CA_a = (-hi["Cx"] + lo["Cx"]) / (2.0 * h)
            CN_a = (-hi["Cz"] + lo["Cz"]) / (2.0 * h)
            Cm_a = (hi["Cm"] - lo["Cm"]) / (2.0 * h)
            CA_a2 = (-hi["Cx"] + 2.0 * mid["Cx"] - lo["Cx"]) / (h * h)
            CN_a2 = (-hi["Cz"] + 2.0 * mid["Cz"] - lo["Cz"]) / (h * h)
            Cm_a2 = (hi["Cm"] - 2.0 * mid["Cm"] + lo["Cm"]) / (h * h)
            denom = CN_a * CA_a2 - CA_a * CN_a2
            x_ac = (CA_a * Cm_a2 - Cm_a * CA_a2) / denom
            z_ac = (CN_a * Cm_a2 - Cm_a * CN_a2) / denom
            Cm_ac = mid["Cm"] - x_ac * mid["Cz"] + z_ac * mid["Cx"]
            l_ref = ap.l_ref_lon
            out[name] = {"aero_center": [-x_ac * l_ref + ap.CG[0], 0.0, -z_ac * l_ref + ap.CG[2]], "Cm_ac": Cm_ac}
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(out, handle, indent=4)
        return out

    def target_CL(self, **kwargs):
        """Angle of attack [deg] giving the requested CL (one aircraft, constant wind; scene.py:3887-4006)."""
        names = list(self._airplanes.keys())
        if len(names) != 1:
            raise IOError("target_CL() may not be used when there is more than one aircraft in the scene.")
        if not hasattr(self, "_constant_wind"):
            raise IOError("target_CL() may not be used when the wind is not constant.")
        name = names[0]
        ap = self._airplanes[name]
        verbose = kwargs.get("verbose", False)
        CL_target = kwargs.get("CL")
        controls = kwargs.get("control_state", {})
        if verbose:
            print("\nSetting angle of attack for CL={0}...".format(CL_target))
            print("{0:<25}{1:<25}".format("Alpha", "CL"))
        alpha_original, _, _ = ap.get_aerodynamic_state(v_wind=self._get_wind(ap.p_bar))
        controls_original = copy.copy(ap.current_control_state)
        max_iter, relax = kwargs.get("max_iterations", 100), kwargs.get("relaxation", 1.0)
        alpha = 0.0
        ap.set_aerodynamic_state(alpha=alpha)
        ap.set_control_state(controls)
        CL = self.solve_forces(dimensional=False)[name]["total"]["CL"]
        res = abs(CL - CL_target)
        if verbose: print("{0:<25}{1:<25}".format(alpha, CL))
        it = 0
        batched = self._batched_ok(kwargs)
        while res > 1e-10:
            if batched:      # both finite-difference solves of this secant step in one device pass
                f, b = self._stencil_totals(name, [lambda a: a.set_aerodynamic_state(alpha=alpha + 0.005),
                                                   lambda a: a.set_aerodynamic_state(alpha=alpha - 0.005)], dict(dimensional=False))
                CL_fwd, CL_bwd = f["CL"], b["CL"]
            else:
                ap.set_aerodynamic_state(alpha=alpha + 0.005)
                CL_fwd = self.solve_forces(dimensional=False)[name]["total"]["CL"]
                ap.set_aerodynamic_state(alpha=alpha - 0.005)
                CL_bwd = self.solve_forces(dimensional=False)[name]["total"]["CL"]
            alpha += (CL_target - CL) / ((CL_fwd - CL_bwd) / 0.01) * relax
            ap.set_aerodynamic_state(alpha=alpha)
            CL = self.solve_forces(dimensional=False)[name]["total"]["CL"]
            res = abs(CL - CL_target)
            if verbose: print("{0:<25}{1:<25}".format(alpha, CL))
            it += 1
            if it == max_iter:
                raise MaxIterationError("target_CL", res)
        if kwargs.get("set_state", True):
            ap.set_aerodynamic_state(alpha=alpha)
            self.set_aircraft_control_state(control_state=controls, aircraft=name)
        else:
            ap.set_aerodynamic_state(alpha=alpha_original)
            self.set_aircraft_control_state(controls_original, aircraft=name)
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump({"CL": CL_target, "alpha": alpha}, handle, indent=4)
        return alpha

    def MAC(self, **kwargs):
        out = {name: self._airplanes[name].get_MAC() for name in self._get_aircraft(**kwargs)}
        filename = kwargs.get("filename", None)
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(out, handle)
        return out

    # ------------------------------------------------------------------------------------------------------
    # span-wise distributions (scene.py:2989-3268)
    # ------------------------------------------------------------------------------------------------------
    _DIST_COLUMNS = ("span_frac", "cpx", "cpy", "cpz", "chord", "swept_chord", "twist", "dihedral", "sweep", "aero_sweep", "area",
                     "alpha", "delta_flap", "u", "v", "w", "Re", "M", "q", "section_CL", "section_Cm", "section_parasitic_CD",
                     "section_aL0", "Fx", "Fy", "Fz", "Mx", "My", "Mz", "circ", "CD_i")
    _DIST_HEADER = ("aircraft", "segment", "span_fraction", "control_x", "control_y", "control_z", "chord", "swept_chord", "twist",
                    "dihedral", "sweep", "aero_sweep", "area", "alpha", "flap_deflection", "u", "v", "w", "Re", "M", "q", "CL", "Cm",
                    "parasitic_CD", "alpha_L0", "Fx", "Fy", "Fz", "Mx", "My", "Mz", "Circ", "CD_i")

    def distributions(self, **kwargs):
        if not self._solved:
            self.solve_forces(**kwargs)
        filename = kwargs.get("filename", None)
        if filename is not None and ".csv" not in filename:
            raise IOError("Export file for Scene.distributions() must be .csv.")
        if kwargs.get("make_plots", []):
            raise NotImplementedError("plotting is outside the solve path this package implements")
        radians = kwargs.get("radians", True)
        ang = (lambda x: x) if radians else np.degrees
        swept = self._use_swept_sections
        dist, rows = {}, []
        k = 0
        for ap in self._airplane_objects:
            dist[ap.name] = {}
            dF_inv_b = earth_to_body(ap.q, self._dF_inv)
            dF_b = dF_inv_b + earth_to_body(ap.q, self._dF_visc)
            dM_b = earth_to_body(ap.q, self._dM_inv + self._dM_visc)
            v_inf = -ap.v + self._get_wind(ap.p_bar)
            V_inf = np.linalg.norm(v_inf)
            u_inf = earth_to_body(ap.q, v_inf / V_inf)
            CD_i = 2.0 * np.einsum("j,ij->i", u_inf, dF_inv_b) / (self._get_density(ap.p_bar) * V_inf ** 2 * self._dS)
            for seg in ap.segments:
                sl = slice(k, k + seg.N)
                v_body = earth_to_body(ap.q, self._v_i[sl, :])
                aL0 = self._aL0[sl] * self._C_sweep_inv[sl] if swept else self._aL0[sl]
                redim = self._redim_in_plane if self._use_in_plane else self._redim_full
                d = {
                    "span_frac": seg.cp_span_locs, "cpx": self._PC[sl, 0], "cpy": self._PC[sl, 1], "cpz": self._PC[sl, 2],
                    "chord": self._c_bar[sl] * self._C_sweep_inv[sl] if swept else self._c_bar[sl], "swept_chord": self._c_bar[sl],
                    "area": self._dS[sl], "twist": ang(seg.twist_cp), "dihedral": ang(seg.dihedral_cp), "sweep": ang(seg.sweep_cp),
                    "aero_sweep": ang(self._section_sweep[sl]), "section_aL0": ang(aL0), "alpha": ang(self._alpha[sl]),
                    "delta_flap": ang(seg._delta_flap), "section_CL": self._CL[sl], "section_Cm": self._Cm[sl],
                    "section_parasitic_CD": self._CD[sl], "Fx": dF_b[sl, 0], "Fy": dF_b[sl, 1], "Fz": dF_b[sl, 2],
                    "Mx": dM_b[sl, 0], "My": dM_b[sl, 1], "Mz": dM_b[sl, 2], "circ": self._gamma[sl], "CD_i": CD_i[sl],
                    "u": v_body[:, 0], "v": v_body[:, 1], "w": v_body[:, 2], "Re": self._Re[sl], "M": self._M[sl],
                    "q": redim[sl] / self._dS[sl]}
                dist[ap.name][seg.name] = {key: list(val) for key, val in d.items()}
                if filename is not None:
                    for i in range(seg.N):
                        rows.append((ap.name, seg.name) + tuple(float(d[c][i]) for c in self._DIST_COLUMNS))
                k += seg.N
        if filename is not None:
            with open(filename, "w") as handle:
                handle.write(",".join(self._DIST_HEADER) + "\n")
                fmt = "%-20s,%-20s," + ",".join(["%20.12e"] * len(self._DIST_COLUMNS))
                for row in rows:
                    handle.write(fmt % row + "\n")
        return dist

    # ------------------------------------------------------------------------------------------------------
    # outside the lifting-line solve path
    # ------------------------------------------------------------------------------------------------------
    def _out_of_scope(self, *args, **kwargs):
        raise NotImplementedError("plotting / CAD / simulator-model export is outside the lifting-line solve path this package implements")

    display_wireframe = display_planform = export_stl = export_vtk = export_stp = export_dxf = _out_of_scope

    def out_gamma(self, filename="gamma_dist.txt"):
        """Writes the circulation distribution and the local velocities of the last solve to a text file in the reference's
        layout (scene.py:3792-3827: ``i y gamma`` per control point, a header line, then ``y v_i V_i``); the two matplotlib figures
        the reference opens afterwards are outside the solve path and are not produced."""
        if not self._solved:
            self.solve_forces()
        y_locs = self._PC[:, 1]
        v_i = np.asarray(self._v_i)
        V_i = np.linalg.norm(v_i, axis=-1)
        gamma = np.asarray(self._gamma)
        with open(filename, "w") as handle:
            for i in range(self._N):
                print(i, y_locs[i], gamma[i], file=handle)
            print("i  y  v_i  V_i", file=handle)
            for i in range(self._N):
                print(y_locs[i], v_i[i, :], V_i[i], file=handle)

    # ------------------------------------------------------------------------------------------------------
    # linearised model for the Pylot simulator: 59 solves of the hot path (scene.py:3521-3747)
    # ------------------------------------------------------------------------------------------------------
    def export_pylot_model(self, **kwargs):
        """Linearised coefficient model of the single aircraft of the scene (reference solve, all derivatives, a 21-point
        drag polar in alpha and one in beta).  ``batched=True`` runs the derivative stencils and the two 21-state sweeps as
        batched device passes.  Sets the aircraft to zero aerodynamic angles and zero control deflections, like the reference.
        Returns the model dictionary (the reference only writes the file)."""
        names = list(self._airplanes.keys())
        if len(names) != 1:
            raise IOError("export_pylot_model() may not be used when there is more than one aircraft in the scene.")
        name = names[0]
        ap = self._airplanes[name]
        batched = self._batched_ok(kwargs)
        model = copy.deepcopy(ap._input_dict)
        model.pop("wings")
        model.pop("airfoils")
        model["units"] = self._unit_sys
        model["CG"] = np.asarray(ap.CG).tolist()
        model["weight"] = float(ap.W)
        model["reference"] = {"area": float(ap.S_w), "longitudinal_length": float(ap.l_ref_lon), "lateral_length": float(ap.l_ref_lat)}
        if "inertia" not in model:
            model["inertia"] = kwargs.get("inertia", {k: "PLEASE SPECIFY" for k in ("Ixx", "Iyy", "Izz", "Ixy", "Ixz", "Iyz")})
        if "angular_momentum" not in model:
            model["angular_momentum"] = list(kwargs.get("angular_momentum", [0.0, 0.0, 0.0]))
        controller = kwargs.get("controller_type", None)
        for value in model.get("controls", {}).values():
            if controller in ("keyboard", "joystick", None):
                if "max_deflection" not in value and controller is None:
                    value["max_deflection"] = "PLEASE SPECIFY"
                value["input_axis"] = value.get("input_axis", "PLEASE SPECIFY")
            if controller in ("time_sequence", None):
                value["column_index"] = value.get("column_index", "PLEASE SPECIFY")
        model["aero_model"] = {"type": "linearized_coefficients"}
        for key in ("stall_angle_of_attack", "stall_sideslip_angle"):
            if key in kwargs:
                model["aero_model"][key] = kwargs[key]

        V_ref = kwargs.get("velocity", 100)
        self.set_aircraft_state(state={"velocity": V_ref})
        self.set_aircraft_control_state()
        FM_ref = self.solve_forces(dimensional=False)
        derivs = self.derivatives(batched=batched)[name]
        co = model["coefficients"] = {}
        co["CL0"] = float(FM_ref[name]["total"]["CL"])
        co["Cm0"] = float(FM_ref[name]["total"]["Cm"])
        for key in ("CL,a", "Cm,a", "CS,b", "Cl,b", "Cn,b"):
            co[key] = float(derivs["stability"][key])
        for coef, rate in (("CS", "p"), ("Cl", "p"), ("Cn", "p"), ("CL", "q"), ("CD", "q"), ("Cm", "q"), ("CS", "r"), ("Cl", "r"), ("Cn", "r")):
            co["{0},{1}_bar".format(coef, rate)] = float(derivs["damping"]["{0},{1}bar".format(coef, rate)])
        unknown = 0.0 if kwargs.get("set_accel_derivs", False) else "PLEASE SPECIFY"
        for key in ("CL,a_hat", "CD,a_hat", "Cm,a_hat", "CS,b_hat", "Cl,b_hat", "Cn,b_hat"):
            co[key] = unknown
        for control in ap.control_names:
            co[control] = {k: float(derivs["control"]["{0},d{1}".format(k, control)]) for k in ("CL", "CD", "CS", "Cl", "Cm", "Cn")}

        def sweep(states, keys):
            if batched:
                tot = self.solve_forces_batch([dict(st, position=[float(x) for x in ap.p_bar], velocity=float(V_ref)) for st in states])["totals"][name]
                return [np.array(tot[k], dtype=float) for k in keys]
            vals = [[] for _ in keys]
            for st in states:
                self.set_aircraft_state(state=dict(st, velocity=V_ref))
                tot = self.solve_forces(dimensional=False, body_frame=False)[name]["total"]
                for v, k in zip(vals, keys):
                    v.append(tot[k])
            return [np.array(v) for v in vals]

        alphas = np.linspace(-10, 10, 21)
        CL, CD = sweep([{"alpha": float(a)} for a in alphas], ("CL", "CD"))
        fit = np.polyfit(CL, CD, 2)
        co["CD0"], co["CD1"], co["CD2"] = float(fit[2]), float(fit[1]), float(fit[0])
        line = np.polyfit(alphas, CL, 1)
        a_L0 = -line[1] / line[0]
        betas = np.linspace(-10, 10, 21)
        CS, CD = sweep([{"alpha": float(a_L0), "beta": float(b)} for b in betas], ("CS", "CD"))
        co["CD3"] = float(np.polyfit(CS, CD, 2)[0])
        if batched:      # the sequential sweep leaves the aircraft at the last state it solved
            self.set_aircraft_state(state={"velocity": V_ref, "alpha": float(a_L0), "beta": float(betas[-1])})
        model["engines"] = model.get("engines", {"placeholder_engine": {}})
        model["landing_gear"] = model.get("landing_gear", {"placeholder_landing_gear": {}})
        filename = kwargs.get("filename", name + "_linearized.json")
        if filename is not None:
            with open(filename, "w") as handle:
                json.dump(model, handle, indent=4)
        return model
