"""Table layout shared by the host package, the oracle and liblls (mirrors include/lls.h).

The integer constants here MUST match the enums in ``include/lls.h``;
``tests/test_abi.py`` parses the header and checks every one of them.
"""

# per-control-point double fields (enum lls_field)
F_PC, F_P0, F_P1, F_P0J, F_P1J, F_PCD, F_UAU, F_UNU, F_UA, F_UN, F_US, F_DL, F_RCG = (
    0, 3, 6, 9, 12, 15, 18, 21, 24, 27, 30, 33, 36)
F_SPC, F_SP0, F_SP1, F_SIGMA, F_CDJ0, F_CDJ1, F_CBAR, F_DS, F_CSWI = 39, 40, 41, 42, 43, 44, 45, 46, 47
F_RHO, F_NU, F_SOS, F_VWIND = 48, 49, 50, 51
F_WB, F_FLAP_DA, F_FLAP_DCM, F_FLAP_DCD, F_FLAP_DEFL, F_FLAP_CF = 54, 55, 56, 57, 58, 59
NF = 60
FLAP_FIELDS = (F_FLAP_DA, F_FLAP_DCM, F_FLAP_DCD, F_FLAP_DEFL, F_FLAP_CF)   # the rows that change with the control state only

# per-control-point int32 fields (enum lls_ifield)
I_WING_BEGIN, I_WING_END, I_AIRCRAFT, I_REID, I_SEGMENT, I_AFA, I_AFB = 0, 1, 2, 3, 4, 5, 6
NI = 7

# airfoil block (enum lls_airfoil_field)
AF_TYPE, AF_AL0, AF_CLA, AF_CML0, AF_CMA, AF_CD0, AF_CD1, AF_CD2, AF_CLMAX = 0, 1, 2, 3, 4, 5, 6, 7, 8
AF_POLY_OFF, AF_POLY_NDOF, AF_POLY_DOF0 = 1, 2, 3   # poly_fit (AF_TYPE == 1) and database (AF_TYPE == 2) airfoils reuse the block
AF_STRIDE = 16
POLY_DOF_KINDS = ("alpha", "Rey", "Mach", "trailing_flap_deflection", "trailing_flap_fraction")

# per-aircraft state block (enum lls_aircraft_field)
AC_VTRANS, AC_W, AC_POS, AC_CGR, AC_ZF = 0, 3, 6, 9, 12
AC_STRIDE = 16

# per-control-point output fields (enum lls_out_field)
O_GAMMA, O_VI, O_ALPHA, O_CL, O_CD, O_CM, O_RE, O_MACH = 0, 1, 4, 5, 6, 7, 8, 9
O_DFINV, O_DMINV, O_DFVISC, O_DMVISC = 10, 13, 16, 19
O_AL0, O_REDIM_IP, O_REDIM_FULL, O_UTRAIL0, O_UTRAIL1, O_R, O_VI2 = 22, 23, 24, 25, 28, 31, 32
NO = 33

# solver flags
FLAG_SWEPT_SECTIONS, FLAG_TOTAL_VELOCITY, FLAG_IN_PLANE, FLAG_MATCH_MACHUP_PRO, FLAG_CONSTRAIN_VORTEX_SHEET = 1, 2, 4, 8, 16
FLAG_DATABASE_AIRFOILS = 32   # some airfoil block is a table look-up (AF_TYPE == 2)
FLAG_DEFAULT = FLAG_SWEPT_SECTIONS | FLAG_TOTAL_VELOCITY | FLAG_IN_PLANE

# status bits
STATUS_NOT_CONVERGED, STATUS_IMPINGEMENT, STATUS_NAN, STATUS_SMALL_PIVOT, STATUS_SECTION_BOUNDS, STATUS_COMM = 1, 2, 4, 8, 16, 32
STATUS_BITS = 6

LD_ALIGN = 64  # leading dimension is N rounded up to this many elements


def pad_ac_state(ac_state):
    """(n_aircraft, AC_STRIDE) view of a state block; blocks written before the body-z-axis columns existed (12 doubles per
    aircraft, as in the committed golden fixtures) are widened with zeros (the columns are read only with
    FLAG_CONSTRAIN_VORTEX_SHEET)."""
    import numpy as np
    a = np.asarray(ac_state, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 12 if a.size % AC_STRIDE else AC_STRIDE)
    if a.shape[-1] == AC_STRIDE:
        return a
    assert a.shape[-1] == 12, a.shape
    out = np.zeros(a.shape[:-1] + (AC_STRIDE,))
    out[..., :12] = a
    return out


def padded_ld(n):
    return ((int(n) + LD_ALIGN - 1) // LD_ALIGN) * LD_ALIGN
