"""poly_fit stand-ins of linear airfoils: with no flap deflection and below CL_max the airfoil_db "linear" model is itself a
polynomial in alpha (CL degree 1, CD degree 2 through CD(CL), Cm degree 1), so the reference's goldens pin the poly_fit code path."""
import copy


def linear_as_poly(spec, extra_dofs=()):
    aL0, CLa = spec.get("aL0", 0.0), spec.get("CLa", 6.283185307179586)
    CmL0, Cma = spec.get("CmL0", 0.0), spec.get("Cma", 0.0)
    CD0, CD1, CD2 = spec.get("CD0", 0.0), spec.get("CD1", 0.0), spec.get("CD2", 0.0)
    c0, c1 = -CLa * aL0, CLa
    cl = [c0, c1]
    cd = [CD0 + CD1 * c0 + CD2 * c0 * c0, CD1 * c1 + 2.0 * CD2 * c0 * c1, CD2 * c1 * c1]
    cm = [CmL0 - Cma * aL0, Cma]
    dofs, limits = ["alpha"], [[-1.0, 1.0]]
    pad = 1
    for d, lim in extra_dofs:           # declared but with degree 0: must not change anything
        dofs.append(d)
        limits.append(list(lim))
    def widen(coefs):
        return [float(x) for x in coefs]
    return {"type": "poly_fit", "degrees_of_freedom": dofs, "limits": limits,
            "fit_degrees": {"CL": [1] + [0] * (len(dofs) - 1), "CD": [2] + [0] * (len(dofs) - 1), "Cm": [1] + [0] * (len(dofs) - 1)},
            "fit_coefs": {"CL": widen(cl), "CD": widen(cd), "Cm": widen(cm)}, "geometry": spec.get("geometry", {})}


def with_poly_airfoils(scene_input, extra_dofs=()):
    d = copy.deepcopy(scene_input)
    for ac in d["scene"]["aircraft"].values():
        af = ac["file"]["airfoils"]
        for name in list(af):
            af[name] = linear_as_poly(af[name], extra_dofs)
    return d
