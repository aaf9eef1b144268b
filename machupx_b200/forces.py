"""Host-side assembly of the force/moment dictionary from the per-segment device sums.

The device (lls_integrate) returns, per wing segment, the earth-axes sums of the inviscid and viscous
force and moment elements.  This module rotates them into the body / stability / wind frames,
non-dimensionalises them and lays them out exactly like the reference's ``Scene._FM``
(scene.py:1048-1406): ``FM[aircraft][inviscid|viscous][key][segment|"total"]`` and
``FM[aircraft]["total"][key]``.  O(#segments) work; no control-point loops here.
"""
import numpy as np

from .rotations import earth_to_body

BODY_F, BODY_M = ("Fx", "Fy", "Fz"), ("Mx", "My", "Mz")
BODY_CF, BODY_CM = ("Cx", "Cy", "Cz"), ("Cl", "Cm", "Cn")
STAB_F, STAB_M = ("Fx_s", "Fy_s", "Fz_s"), ("Mx_s", "My_s", "Mz_s")
STAB_CF, STAB_CM = ("Cx_s", "Cy_s", "Cz_s"), ("Cl_s", "Cm_s", "Cn_s")
WIND_F, WIND_M = ("FD", "FS", "FL"), ("Mx_w", "My_w", "Mz_w")
WIND_CF, WIND_CM = ("CD", "CS", "CL"), ("Cl_w", "Cm_w", "Cn_w")


class AircraftFrame:
    """Reference quantities of one aircraft needed to express and scale its loads."""

    def __init__(self, name, q, v_earth, wind_at_origin, rho_at_origin, S_w, l_ref_lat, l_ref_lon, segment_names):
        self.name = name
        self.q = np.asarray(q, dtype=float)
        self.segment_names = list(segment_names)
        v_inf = -np.asarray(v_earth, dtype=float) + np.asarray(wind_at_origin, dtype=float)      # scene.py:1078
        self.V_inf = float(np.linalg.norm(v_inf))
        u_inf = earth_to_body(self.q, (v_inf / self.V_inf).flatten())                             # scene.py:1080
        u_lift = np.cross(u_inf, [0.0, 1.0, 0.0])                                                 # scene.py:1084-1085
        u_lift = u_lift / np.linalg.norm(u_lift)
        u_x_stab = np.cross(u_lift, [0.0, 1.0, 0.0])                                              # scene.py:1087-1089
        u_x_stab = u_x_stab / np.linalg.norm(u_x_stab)
        self.to_stab = np.array([u_x_stab, [0.0, 1.0, 0.0], -u_lift])
        u_side = np.cross(u_lift, u_inf)                                                          # scene.py:1091-1093
        u_side = u_side / np.linalg.norm(u_side)
        self.to_wind = np.array([u_inf, u_side, u_lift])
        # numpy scalars, as in the reference: a zero reference length (a main wing that exists on the left side only,
        # airplane.py:883-892) gives infinite / NaN moment coefficients there, not an exception
        with np.errstate(divide="ignore", invalid="ignore"):
            self.nd_force = np.float64(2.0) / (np.float64(rho_at_origin) * self.V_inf * self.V_inf * np.float64(S_w))   # scene.py:1097-1099
            self.nd_lat = self.nd_force / np.float64(l_ref_lat)
            self.nd_lon = self.nd_force / np.float64(l_ref_lon)


def frames_from_scene(scene):
    """AircraftFrame list for a Scene-like object (reads only attributes the reference's Scene also has)."""
    frames = []
    for ap in scene._airplane_objects:
        frames.append(AircraftFrame(ap.name, ap.q, ap.v, scene._get_wind(ap.p_bar), scene._get_density(ap.p_bar), ap.S_w,
                                    ap.l_ref_lat, ap.l_ref_lon, [seg.name for seg in ap.segments]))
    return frames


def _store(dst, keys, vec, scale, where):
    for k, x, s in zip(keys, vec, scale):
        dst[k][where] = float(x) * s


def assemble_fm(seg_fm, frames, non_dimensional=True, dimensional=True, report_by_segment=False,
                body_frame=True, stab_frame=False, wind_frame=True):
    """seg_fm: (n_segments, 12) rows [F_inv, M_inv, F_visc, M_visc] in earth axes, scene order."""
    seg_fm = np.asarray(seg_fm, dtype=float)
    FM = {}
    row = 0
    for fr in frames:
        coef_keys, dim_keys = [], []
        if body_frame:
            coef_keys += BODY_CF + BODY_CM
            dim_keys += BODY_F + BODY_M
        if stab_frame:
            coef_keys += STAB_CF + STAB_CM
            dim_keys += STAB_F + STAB_M
        if wind_frame:   # the reference declares CL, CD, CS first (scene.py:1003) but fills by frame axis order
            coef_keys += ("CL", "CD", "CS") + WIND_CM
            dim_keys += ("FL", "FD", "FS") + WIND_M
        entry = {"inviscid": {}, "viscous": {}, "total": {}}
        for part in ("inviscid", "viscous"):
            if non_dimensional:
                entry[part].update({k: {} for k in coef_keys})
            if dimensional:
                entry[part].update({k: {} for k in dim_keys})

        nd_f = (fr.nd_force,) * 3
        nd_m = (fr.nd_lat, fr.nd_lon, fr.nd_lat)
        one = (1.0, 1.0, 1.0)
        tot = {p: {f: np.zeros(6) for f in "bsw"} for p in ("inviscid", "viscous")}

        def emit(part, frame_id, F, M, where):
            cf, cm, df, dm = {"b": (BODY_CF, BODY_CM, BODY_F, BODY_M), "s": (STAB_CF, STAB_CM, STAB_F, STAB_M),
                              "w": (WIND_CF, WIND_CM, WIND_F, WIND_M)}[frame_id]
            if non_dimensional:
                _store(entry[part], cf, F, nd_f, where)
                _store(entry[part], cm, M, nd_m, where)
            if dimensional:
                _store(entry[part], df, F, one, where)
                _store(entry[part], dm, M, one, where)

        for seg_name in fr.segment_names:
            s = seg_fm[row]
            row += 1
            for part, F_e, M_e in (("inviscid", s[0:3], s[3:6]), ("viscous", s[6:9], s[9:12])):
                F_b = earth_to_body(fr.q, F_e)                                                    # scene.py:1108-1117
                M_b = earth_to_body(fr.q, M_e)
                by_frame = {}
                if body_frame:
                    by_frame["b"] = (F_b, M_b)
                if stab_frame:
                    by_frame["s"] = (fr.to_stab @ F_b, fr.to_stab @ M_b)                          # scene.py:1125-1129
                if wind_frame:
                    by_frame["w"] = (fr.to_wind @ F_b, fr.to_wind @ M_b)                          # scene.py:1120-1124
                for fid, (F, M) in by_frame.items():
                    if report_by_segment:
                        emit(part, fid, F, M, seg_name)
                    tot[part][fid][:3] += F                                                       # scene.py:1238-1252
                    tot[part][fid][3:] += M

        frames_on = [f for f, on in (("b", body_frame), ("s", stab_frame), ("w", wind_frame)) if on]
        for fid in frames_on:
            for part in ("inviscid", "viscous"):
                emit(part, fid, tot[part][fid][:3], tot[part][fid][3:], "total")
        # airplane totals, keyed directly (scene.py:1273-1280 etc.)
        for want, scales in ((non_dimensional, (nd_f, nd_m, 0)), (dimensional, (one, one, 1))):
            if not want:
                continue
            for fid in frames_on:
                both = tot["viscous"][fid] + tot["inviscid"][fid]
                cf, cm, df, dm = {"b": (BODY_CF, BODY_CM, BODY_F, BODY_M), "s": (STAB_CF, STAB_CM, STAB_F, STAB_M),
                                  "w": (WIND_CF, WIND_CM, WIND_F, WIND_M)}[fid]
                fk, mk = (cf, cm) if scales[2] == 0 else (df, dm)
                for k, x, sc in zip(fk, both[:3], scales[0]):
                    entry["total"][k] = float(x) * sc
                for k, x, sc in zip(mk, both[3:], scales[1]):
                    entry["total"][k] = float(x) * sc
        FM[fr.name] = entry
    return FM


def batch_totals(seg_fm, ac_states, aircraft):
    """Aircraft totals of a batch of flight states without building per-case dictionaries.

    seg_fm: (ncases, n_segments, 12) earth-axes segment sums from the device; ac_states: (ncases, n_aircraft, AC_STRIDE)
    state blocks the cases were solved with; aircraft: list of dicts with keys name, q, wind, rho, S_w, l_ref_lat,
    l_ref_lon, seg_range (start, stop).  Returns {name: {key: ndarray(ncases)}} with the body- and wind-frame
    coefficients and dimensional loads of ``Scene._FM[name]["total"]`` (scene.py:1078-1129, 1238-1406)."""
    from . import layout as L
    from .rotations import rotation_matrix
    seg_fm = np.asarray(seg_fm, dtype=float)
    out = {}
    for ia, ac in enumerate(aircraft):
        r0, r1 = ac["seg_range"]
        tot = seg_fm[:, r0:r1, :].sum(axis=1)                      # (ncases, 12)
        F_e = tot[:, 0:3] + tot[:, 6:9]
        M_e = tot[:, 3:6] + tot[:, 9:12]
        C = rotation_matrix(ac["q"])                                # v_body = C v_earth
        F_b, M_b = F_e @ C.T, M_e @ C.T
        v_inf = ac_states[:, ia, L.AC_VTRANS:L.AC_VTRANS + 3] + np.asarray(ac["wind"], dtype=float)[np.newaxis, :]
        V = np.linalg.norm(v_inf, axis=1)
        u_inf = (v_inf / V[:, np.newaxis]) @ C.T
        y = np.array([0.0, 1.0, 0.0])
        u_lift = np.cross(u_inf, y)
        u_lift /= np.linalg.norm(u_lift, axis=1, keepdims=True)
        u_side = np.cross(u_lift, u_inf)
        u_side /= np.linalg.norm(u_side, axis=1, keepdims=True)
        dot = lambda a, b: np.einsum("ij,ij->i", a, b)
        nd = 2.0 / (float(ac["rho"]) * V * V * ac["S_w"])
        nd_lat, nd_lon = nd / ac["l_ref_lat"], nd / ac["l_ref_lon"]
        d = {"Fx": F_b[:, 0], "Fy": F_b[:, 1], "Fz": F_b[:, 2], "Mx": M_b[:, 0], "My": M_b[:, 1], "Mz": M_b[:, 2],
             "FD": dot(F_b, u_inf), "FS": dot(F_b, u_side), "FL": dot(F_b, u_lift),
             "Mx_w": dot(M_b, u_inf), "My_w": dot(M_b, u_side), "Mz_w": dot(M_b, u_lift)}
        d.update({"Cx": d["Fx"] * nd, "Cy": d["Fy"] * nd, "Cz": d["Fz"] * nd,
                  "Cl": d["Mx"] * nd_lat, "Cm": d["My"] * nd_lon, "Cn": d["Mz"] * nd_lat,
                  "CD": d["FD"] * nd, "CS": d["FS"] * nd, "CL": d["FL"] * nd,
                  "Cl_w": d["Mx_w"] * nd_lat, "Cm_w": d["My_w"] * nd_lon, "Cn_w": d["Mz_w"] * nd_lat})
        out[ac["name"]] = d
    return out
