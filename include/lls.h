/*
 * liblls — C ABI of the B200-native lifting-line solve path (sm_100a CUDA).
 *
 * MachUpX (the reference) has no FFI: its hot path is a set of private numpy methods on
 * machupX.Scene.  This header is the boundary a maintainer would bind (ctypes, see
 * INTEGRATION.md) to replace exactly those methods.  Each entry point cites the reference
 * method it replaces (file:line under the reference tree).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, int return codes (LLS_OK == 0), no exceptions.
 *   - Every `d_*` pointer is DEVICE memory owned by the caller (the Python host passes
 *     torch.Tensor.data_ptr()); the library allocates no large buffers itself.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - All floating point data is IEEE double.  Tables are field-major ("SoA"):
 *       tab[f * ld + i]   f = field index below, i = control-point / horseshoe index, ld >= n.
 *   - Vectors are stored in earth-fixed AXES but relative to the owning aircraft's origin
 *     (rotated by the aircraft quaternion, not translated); aircraft origins are in the
 *     per-aircraft state block, so same-aircraft differences never lose digits to a large
 *     position offset.
 */
#ifndef LLS_H
#define LLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return / status codes ------------------------------------------------------------ */
#define LLS_OK 0
#define LLS_ERR_CUDA 1
#define LLS_ERR_ARG 2
#define LLS_ERR_UNSUPPORTED 3
#define LLS_ERR_STATE 4

/* status bits written by the solvers (reference: scene.py:622-623, 905-908; common_issues.md:12) */
#define LLS_STATUS_NOT_CONVERGED 1
#define LLS_STATUS_IMPINGEMENT 2
#define LLS_STATUS_NAN 4
#define LLS_STATUS_SMALL_PIVOT 8
#define LLS_STATUS_BITS 6 /* number of status bits defined below */
#define LLS_STATUS_COMM 32 /* row-partitioned solve: a rank waited 30 s for a peer's data (a peer died or fell out of step) */
#define LLS_STATUS_SECTION_BOUNDS 16 /* a poly_fit airfoil was evaluated outside the limits of its fit, or a database airfoil outside its table (airfoil_db: DatabaseBoundsError) */

/* ---- solver flags (reference: scene.py:82-87) ------------------------------------------ */
#define LLS_FLAG_SWEPT_SECTIONS 1
#define LLS_FLAG_TOTAL_VELOCITY 2
#define LLS_FLAG_IN_PLANE 4
#define LLS_FLAG_MATCH_MACHUP_PRO 8
#define LLS_FLAG_CONSTRAIN_VORTEX_SHEET 16 /* trailing directions projected off the body z axis and renormalised (scene.py:595-611) */
#define LLS_FLAG_DATABASE_AIRFOILS 32 /* at least one airfoil block is a table look-up (LLS_AF_TYPE == 2): the section-model kernels with the table walk are launched */
#define LLS_FLAG_DEFAULT (LLS_FLAG_SWEPT_SECTIONS | LLS_FLAG_TOTAL_VELOCITY | LLS_FLAG_IN_PLANE)

/* ---- per-control-point double table fields --------------------------------------------- */
enum lls_field {
    LLS_F_PC = 0,      /* 3: control point (airplane.py:503, scene.py:445-447)                */
    LLS_F_P0 = 3,      /* 3: inbound node (airplane.py:528)                                   */
    LLS_F_P1 = 6,      /* 3: outbound node (airplane.py:529)                                  */
    LLS_F_P0J = 9,     /* 3: inbound plain joint (airplane.py:555)                            */
    LLS_F_P1J = 12,    /* 3: outbound plain joint (airplane.py:556)                           */
    LLS_F_PCD = 15,    /* 3: d(PC)/d(span) direction, PC_deriv (airplane.py:564)              */
    LLS_F_UAU = 18,    /* 3: unswept axial unit vector (airplane.py:537)                      */
    LLS_F_UNU = 21,    /* 3: unswept normal unit vector (airplane.py:538)                     */
    LLS_F_UA = 24,     /* 3: section axial vector used by the solver (scene.py:451-458)       */
    LLS_F_UN = 27,     /* 3: section normal vector                                            */
    LLS_F_US = 30,     /* 3: section spanwise vector                                          */
    LLS_F_DL = 33,     /* 3: bound segment P1-P0 (airplane.py:689)                            */
    LLS_F_RCG = 36,    /* 3: control point relative to CG (airplane.py:561)                   */
    LLS_F_SPC = 39,    /* control-point span location from the wing's left tip (airplane.py:519) */
    LLS_F_SP0 = 40,    /* inbound node span location (airplane.py:524)                        */
    LLS_F_SP1 = 41,    /* outbound node span location (airplane.py:525)                       */
    LLS_F_SIGMA = 42,  /* blending sigma (airplane.py:513)                                    */
    LLS_F_CDJ0 = 43,   /* P0 chord * joint length (airplane.py:635)                           */
    LLS_F_CDJ1 = 44,   /* P1 chord * joint length (airplane.py:647)                           */
    LLS_F_CBAR = 45,   /* mean chord, already * cos(sweep) if swept sections (scene.py:423-425) */
    LLS_F_DS = 46,     /* planform area element (wing_segment.py:691)                         */
    LLS_F_CSWI = 47,   /* 1/cos(section sweep) (scene.py:426); 1 if swept sections are off    */
    LLS_F_RHO = 48,    /* density at the control point (scene.py:544)                         */
    LLS_F_NU = 49,     /* kinematic viscosity (scene.py:545)                                  */
    LLS_F_SOS = 50,    /* speed of sound (scene.py:546)                                       */
    LLS_F_VWIND = 51,  /* 3: wind at the control point (scene.py:547)                         */
    LLS_F_WB = 54,     /* spanwise airfoil blend weight d (wing_segment.py:1070)              */
    LLS_F_FLAP_DA = 55,  /* flap: delta * eps_ideal * eta_hinge * eta_defl   [rad]            */
    LLS_F_FLAP_DCM = 56, /* flap: delta * (sin 2theta_f - 2 sin theta_f)/4                    */
    LLS_F_FLAP_DCD = 57, /* flap: 0.002 * |delta| * 180/pi                                    */
    LLS_F_FLAP_DEFL = 58, /* raw flap deflection [rad] (wing_segment.py:1370-1374)            */
    LLS_F_FLAP_CF = 59,   /* flap chord fraction (wing_segment.py:656)                        */
    LLS_NF = 60
};

/* ---- per-control-point int32 table fields ---------------------------------------------- */
enum lls_ifield {
    LLS_I_WING_BEGIN = 0, /* first scene index of the wing (lifting line) holding i (airplane.py:548) */
    LLS_I_WING_END = 1,   /* one past the last                                                 */
    LLS_I_AIRCRAFT = 2,   /* aircraft index (scene.py:405-420)                                 */
    LLS_I_REID = 3,       /* 1 if general-NLL (Reid) corrections are on for i (airplane.py:511) */
    LLS_I_SEGMENT = 4,    /* scene-wide segment index, for the force sums (scene.py:1102-1117) */
    LLS_I_AFA = 5,        /* airfoil id of the inboard bracketing airfoil                      */
    LLS_I_AFB = 6,        /* airfoil id of the outboard bracketing airfoil (== AFA if single)  */
    LLS_NI = 7
};

/* ---- airfoil parameter block (one per airfoil id), doubles ------------------------------ */
enum lls_airfoil_field {
    LLS_AF_TYPE = 0, /* 0 = linear (airfoil_db "linear") ; 1 = poly_fit ; 2 = database (needs LLS_FLAG_DATABASE_AIRFOILS) */
    LLS_AF_AL0 = 1,
    LLS_AF_CLA = 2,
    LLS_AF_CML0 = 3,
    LLS_AF_CMA = 4,
    LLS_AF_CD0 = 5,
    LLS_AF_CD1 = 6,
    LLS_AF_CD2 = 7,
    LLS_AF_CLMAX = 8,
    /* poly_fit and database airfoils (LLS_AF_TYPE == 1, 2) reuse the block: */
    LLS_AF_POLY_OFF = 1,  /* offset (in doubles, from the start of the airfoil buffer) of this airfoil's record in the pool      */
    LLS_AF_POLY_NDOF = 2, /* number of independent variables V (<= 5)                                                       */
    LLS_AF_POLY_DOF0 = 3, /* V kind codes: 0 alpha, 1 Rey, 2 Mach, 3 trailing_flap_deflection [rad], 4 trailing_flap_fraction */
    LLS_AF_STRIDE = 16
};
/* Pool record of a poly_fit airfoil (doubles), after the n_airfoils parameter blocks of the airfoil buffer:
 *   limits: V x (lower, upper); then for CL, CD, Cm in this order: ncoef, V degrees, ncoef coefficients
 * with the coefficient index running like airfoil_db's fits: exponent tuple row-major, LAST variable fastest
 * (reference machupX/poly_fits.py:398-431 compose_j).
 *
 * Pool record of a database airfoil (airfoil_db "database": CL, CD, Cm tabulated over V of the five variables, interpolated
 * barycentrically on the Delaunay triangulation of the P data points in variables scaled to [0, 1]; for V = 1 the S = P - 1
 * simplices are the intervals between consecutive points).  The caller triangulates (scipy.spatial.Delaunay on the scaled
 * points is what airfoil_db's griddata call does); the library only walks the table:
 *   P, S; limits: V x (lower, span = upper - lower); values: P x (CL, CD, Cm);
 *   simplices: S x [V + 1 vertex ids (as doubles) | V x V row-major map Tinv | V origin r (= the last vertex, scaled)],
 * barycentric coordinates of a scaled point x: c[0..V) = Tinv (x - r), c[V] = 1 - sum c; a flat simplex carries NaN in Tinv.
 * CLa / CLRe / CLM: central differences of the CL table with steps 0.001 rad / 1000 / 0.05, samples kept inside the limits;
 * aL0: root of the CL table in alpha (secant from 0 and 0.01).  A point outside the table sets LLS_STATUS_SECTION_BOUNDS and
 * extrapolates from the closest simplex.  Pinned by the reference's golden, tests/test_database_airfoils.py. */

/* ---- per-aircraft state block, doubles --------------------------------------------------- */
enum lls_aircraft_field {
    LLS_AC_VTRANS = 0, /* 3: -v (earth frame), freestream due to translation (scene.py:565)   */
    LLS_AC_W = 3,      /* 3: angular rate rotated to earth axes, R^-1(q) w (scene.py:569)     */
    LLS_AC_POS = 6,    /* 3: earth-fixed position p_bar (scene.py:442)                        */
    LLS_AC_CGR = 9,    /* 3: CG rotated to earth axes (scene.py:579)                          */
    LLS_AC_ZF = 12,    /* 3: body z axis in earth axes (scene.py:601), read only with LLS_FLAG_CONSTRAIN_VORTEX_SHEET; [15] is padding */
    LLS_AC_STRIDE = 16
};

/* ---- per-control-point solution table (outputs), field-major with the same ld ------------ */
enum lls_out_field {
    LLS_O_GAMMA = 0,   /* circulation (scene.py:827, 899)                                     */
    LLS_O_VI = 1,      /* 3: local velocity v_i (scene.py:728)                                */
    LLS_O_ALPHA = 4,   /* section angle of attack (scene.py:753, 957)                         */
    LLS_O_CL = 5,      /* section CL incl. sweep correction: of the last residual evaluation (scene.py:766, 797), or, after a linear solve only, the flow-state estimate at alpha_inf (scene.py:659-666) */
    LLS_O_CD = 6,      /* section CD (scene.py:1015)                                          */
    LLS_O_CM = 7,      /* section Cm / cos(sweep) (scene.py:1020, 1026)                       */
    LLS_O_RE = 8,      /* Reynolds number (scene.py:977)                                      */
    LLS_O_MACH = 9,    /* Mach number (scene.py:978)                                          */
    LLS_O_DFINV = 10,  /* 3: inviscid force element (scene.py:952)                            */
    LLS_O_DMINV = 13,  /* 3: inviscid moment element (scene.py:1036)                          */
    LLS_O_DFVISC = 16, /* 3: viscous force element (scene.py:1041)                            */
    LLS_O_DMVISC = 19, /* 3: viscous moment element (scene.py:1046)                           */
    LLS_O_AL0 = 22,    /* zero-lift angle of the linear pre-pass (scene.py:661)               */
    LLS_O_REDIM_IP = 23, /* 0.5 rho V_inplane^2 dS (scene.py:987)                             */
    LLS_O_REDIM_FULL = 24, /* 0.5 rho V^2 dS (scene.py:985)                                   */
    LLS_O_UTRAIL0 = 25, /* 3: trailing direction at the inbound joint (scene.py:591)          */
    LLS_O_UTRAIL1 = 28, /* 3: trailing direction at the outbound joint (scene.py:592)         */
    LLS_O_R = 31,      /* last residual vector (scene.py:723)                                 */
    LLS_O_VI2 = 32,    /* V_i^2 (scene.py:944)                                                */
    LLS_NO = 33
};

typedef struct lls_scene lls_scene; /* opaque */

/* Sizes (bytes) of the caller-provided device buffers for a scene of n control points. */
typedef struct lls_sizes {
    size_t ld;           /* padded leading dimension (elements) used by every table/matrix    */
    size_t tab_bytes;    /* LLS_NF  * ld doubles                                              */
    size_t itab_bytes;   /* LLS_NI  * ld int32                                                */
    size_t out_bytes;    /* LLS_NO  * ld doubles                                              */
    size_t vji_bytes;    /* n * 3 * ld doubles : V[i][c][j]  (scene.py:634 `_V_ji`, component-planar rows) */
    size_t mat_bytes;    /* n * ld doubles     : A / J, overwritten by its LU factors         */
    size_t work_bytes;   /* O(n) scratch + LU block inverses                                  */
} lls_sizes;

int lls_query_sizes(int n, int n_segments, lls_sizes* out);

/* Handle life cycle.  flags = LLS_FLAG_* (reference: scene.py:75-87 solver block). */
int lls_create(lls_scene** out, int n, int n_aircraft, int n_segments, int n_airfoils, int flags);
int lls_destroy(lls_scene* h);

/* Bind caller-owned device buffers (sizes from lls_query_sizes). */
int lls_bind(lls_scene* h, const double* d_tab, const int32_t* d_itab, const double* d_airfoils,
             double* d_out, double* d_vji, double* d_mat, void* d_work, size_t work_bytes);

/* Batch of `ncases` independent flight states of ONE geometry (alpha / beta / rate sweeps, finite-difference stencils of
 * the derivative and trim drivers: scene.py:1876-2437, 2440-2640).  The reference runs them as a serial Python loop of
 * set_aircraft_state + solve_forces; here one set of launches solves all cases (case index = blockIdx.z).
 * d_tab / d_itab / d_airfoils are shared; d_out, d_vji, d_mat and d_work hold ncases consecutive copies of the
 * single-scene buffers (sizes from lls_query_sizes; work_bytes_per_case a multiple of 256, >= work_bytes).
 * d_ctl: 2 * ncases + 4 int32 (active mask, Newton iteration counts, number of active cases).
 * After this call lls_set_state expects ncases * n_aircraft * LLS_AC_STRIDE doubles, lls_flow_state,
 * lls_build_influence, lls_solve_linear and lls_integrate (ncases * n_segments * 12 doubles) act on every case, and
 * lls_solve_nonlinear_batch replaces lls_solve_nonlinear.  lls_bind returns the handle to a single scene. */
int lls_bind_batch(lls_scene* h, int ncases, const double* d_tab, const int32_t* d_itab, const double* d_airfoils,
                   double* d_out, double* d_vji, double* d_mat, void* d_work, size_t work_bytes_per_case, int32_t* d_ctl);

/* Per-case copies of the double table: the cases of the bound batch read d_tab + case * case_stride_doubles (LLS_NF * ld
 * doubles each) instead of one shared table, so a batch can also carry cases that differ in control deflections (the five flap
 * rows), position or orientation (every earth-axes row): the stencils of control_derivatives (scene.py:2180), state_derivatives
 * (:2276), pitch_trim (:2440).  The int table and the airfoils stay shared.  case_stride_doubles = 0 returns to one shared table. */
int lls_set_case_tables(lls_scene* h, const double* d_tab, size_t case_stride_doubles);

/* Newton iteration of every case of the bound batch in lock step; a case leaves the iteration when its residual norm
 * falls below `convergence` exactly as the reference's `while` test does (scene.py:853), so per-case iteration counts
 * equal those of separate solves.  Host outputs, ncases entries each. */
int lls_solve_nonlinear_batch(lls_scene* h, double convergence, double relaxation, int max_iterations, void* stream,
                              int* iterations, double* final_error, int* status);

/* Status bits of every case of the bound batch (host array, ncases entries). */
int lls_get_status_batch(lls_scene* h, void* stream, int* status);

/* ---- one large scene partitioned by rows over the GPUs of one box ------------------------
 * The reference holds V_ji (24 N^2 B) and J (8 N^2 B) in one address space and cannot run N >= ~9k at all; here the rows
 * of V_ji / A / J are dealt to the ranks in 64-row blocks, round-robin (owner of block b = b mod world), and stored
 * consecutively on their owner.  Per-point tables, gamma, the right-hand side and x stay replicated (O(N)).
 * The library issues no collective itself: the host (torch.distributed / NCCL) broadcasts the per-step buffer of the LU,
 * all-reduces the residual norm and the segment sums, and broadcasts each 64-value block of x in the back substitution.
 *   lls_query_sizes_partitioned : buffer sizes per rank (ld = n rounded up to 64 * world)
 *   lls_set_partition           : after lls_create, before lls_bind
 *   lls_flow_state / lls_build_influence / lls_integrate : act on the rows this rank owns
 *   lls_dist_assemble (newton = 0: A, b of scene.py:811-824; 1: Jacobian of scene.py:862-893), lls_dist_residual (local sum of
 *   R_i^2), lls_dist_update_gamma (relaxation < 0: gamma = x), lls_dist_vectors (device pointers of rhs, x, gamma)
 *   LU block step k:  owner(k) calls lls_dist_lu_owner -> step buffer [diag image | y_k | U tiles], host broadcasts
 *   lls_dist_lu_step_doubles(k) doubles of it, every rank calls lls_dist_lu_update; then for k = last..0 the owner calls
 *   lls_dist_back and the host broadcasts x[64 k .. 64 k + 63]. */
int lls_query_sizes_partitioned(int n, int n_segments, int world, lls_sizes* out);
int lls_set_partition(lls_scene* h, int world, int rank);
int lls_dist_assemble(lls_scene* h, int newton, void* stream);
int lls_dist_residual(lls_scene* h, void* stream, double* local_sum);
int lls_dist_update_gamma(lls_scene* h, double relaxation, void* stream);
int lls_dist_vectors(lls_scene* h, double** d_rhs, double** d_x, double** d_gamma);
int lls_dist_lu_begin(lls_scene* h, size_t* max_step_doubles, void* stream);
int lls_dist_lu_step_doubles(lls_scene* h, int k, size_t* doubles);
int lls_dist_lu_owner(lls_scene* h, int k, int with_rhs, double* d_step, void* stream);
int lls_dist_lu_update(lls_scene* h, int k, int with_rhs, const double* d_step, void* stream);
int lls_dist_back(lls_scene* h, int k, void* stream);

/* The same factorisation with the exchange inside the kernels, over NVLink peer memory (no collective call and no host code
 * between two block steps).  Every rank owns a "mailbox" (panel slots, arrival flags, progress counters,
 * the replicated solution) in device memory that the LIBRARY allocates (cudaMalloc: the one large buffer it owns, because only a
 * whole cudaMalloc allocation can be exported through CUDA IPC) and that every other rank maps into its process:
 *   lls_dist_comm_create  : after lls_bind; allocates the mailbox (`slots` panel slots, 2..32) and returns its 64-byte
 *                           cudaIpcMemHandle_t (zeros when world == 1)
 *   lls_dist_comm_connect : all_handles = world x 64 bytes in rank order (the host gathers them, e.g. torch.distributed
 *                           all_gather); opens the peers' mailboxes.  A no-op when world == 1.
 *   lls_dist_lu_solve     : factorises the matrix the ranks hold with the right-hand side carried along and back-substitutes:
 *                           x (lls_dist_vectors) is replicated on return.  Owner kernels store their panel straight into every
 *                           rank's slot and release an arrival flag; the next panel is pushed one step ahead (look-ahead).
 *                           A peer that never arrives raises LLS_STATUS_COMM after 30 s instead of hanging.
 *   lls_dist_comm_stats   : out4 = {ms the panel kernels waited for panels, ms the owner kernels waited for free slots,
 *                           global block steps so far, back substitutions so far}; resets the two timers.
 *   lls_dist_comm_destroy : closes the peers' mailboxes and frees the own one (also done by lls_destroy).
 *   lls_dist_comm_bytes / lls_dist_comm_attach : the alternative to create + connect when the HOST owns symmetric memory (e.g.
 *                           torch.distributed._symmetric_memory): d_mailboxes = world device pointers, the mailbox of every rank as
 *                           mapped into this process (lls_dist_comm_bytes bytes each, zero-filled), d_multicast = the multicast
 *                           (NVLS) mapping of all of them or NULL.  With a multicast mapping every panel store of an owner kernel is
 *                           ONE store replicated by the NVSwitch into all mailboxes: the owner's link carries a panel once instead of
 *                           world - 1 times.  lls_dist_comm_multicast returns 1 if such a mapping is attached. */
int lls_dist_comm_create(lls_scene* h, int slots, unsigned char* ipc_handle_out64);
int lls_dist_comm_bytes(lls_scene* h, int slots, size_t* bytes);
int lls_dist_comm_attach(lls_scene* h, int slots, void* const* d_mailboxes, void* d_multicast);
int lls_dist_comm_multicast(lls_scene* h);
int lls_dist_comm_connect(lls_scene* h, const unsigned char* all_handles);
int lls_dist_lu_solve(lls_scene* h, void* stream);
int lls_dist_comm_stats(lls_scene* h, double* out4, void* stream);
int lls_dist_comm_destroy(lls_scene* h);

/* The launch schedule of lls_dist_lu_solve as data (pure host code: callable without a GPU; tests/test_dist_schedule_cpu.py checks
 * with it that every tile of every rank receives every panel exactly once, in order, and that every panel is factored / solved once).
 *   lls_dist_debug_schedule : the launches of one factorisation of `rank` in order, 9 ints each {k, passes, thin, base, rem, lb0, nloc,
 *                             own, grid}: tiles with block rows / columns >= base = k + passes take panels k .. k + passes - 1 and the
 *                             roles factor panel `base`; pair_min_rem >= 0: threshold of the rank-128 pairs, -1 / -2: the default with /
 *                             without a multicast mapping.  Returns the number of launches (out9 may be NULL), -1 on bad arguments.
 *   lls_dist_debug_role     : block index of a launch -> role (0 DIAG, 1 UROW, 2 LCOL, 3 OTHER, 4 none) and its tile. */
int lls_dist_debug_schedule(int nblk, int world, int rank, int pair_min_rem, int* out9, int max_launches);
int lls_dist_debug_role(int block, int own, int rem, int base, int lb0, int nloc, int thin, int* local_block_row, int* block_col);

/* Per-aircraft flight state, n_aircraft * LLS_AC_STRIDE doubles on the device
 * (reference: Airplane.set_state airplane.py:105-213 -> scene.py:562-582). */
int lls_set_state(lls_scene* h, const double* d_ac_state);

/* K-state: freestream at control points and joints, trailing directions, in-plane speeds,
 * alpha_inf, linear-pass CLa/CL/aL0 with sweep correction.
 * Replaces the O(N) part of Scene._calc_invariant_flow_properties (scene.py:557-592, 636-666). */
int lls_flow_state(lls_scene* h, void* stream);

/* K0+K1: horseshoe influence build V_ji = (trail0 + bound + joint0 + joint1 + trail1)/4pi with the
 * effective-lifting-line (blended, jointed) geometry recomputed per pair in registers.
 * Replaces Airplane._calculate_geometry effective-LAC loop (airplane.py:569-689),
 * Scene._perform_geometry_and_atmos_calcs (scene.py:469-541) and the trailing-leg part of
 * Scene._calc_invariant_flow_properties (scene.py:621-634).
 * impingement_threshold: scene.py:86.  Sets LLS_STATUS_IMPINGEMENT in the scene status. */
int lls_build_influence(lls_scene* h, double impingement_threshold, void* stream);

/* K2+K3: assemble A, b and solve A gamma = b.  Replaces Scene._solve_linear (scene.py:800-829)
 * except its call to _calc_invariant_flow_properties (= lls_flow_state + lls_build_influence). */
int lls_solve_linear(lls_scene* h, void* stream);

/* K4-K8 + K3: Newton iteration on the lifting-line residual with on-device Jacobian.
 * Replaces Scene._solve_nonlinear + _lifting_line_residual + _calc_v_i + _get_section_lift
 * (scene.py:705-797, 832-917).  Host outputs: Newton update count, final ||R||_2, status bits.
 * err_hist (host, may be NULL) receives the residual norm tested at each iteration (<= max_iter). */
int lls_solve_nonlinear(lls_scene* h, double convergence, double relaxation, int max_iterations,
                        void* stream, int* iterations, double* final_error, int* status, double* err_hist);

/* K9+K10: force/moment elements and per-segment sums (earth axes).
 * d_seg_fm: n_segments * 12 doubles [F_inv(3) M_inv(3) F_visc(3) M_visc(3)].
 * Replaces the numeric part of Scene._integrate_forces_and_moments (scene.py:941-1046, 1108-1117). */
int lls_integrate(lls_scene* h, double* d_seg_fm, void* stream);

/* Status bits accumulated since the last lls_flow_state (device flag read back; synchronises). */
int lls_get_status(lls_scene* h, void* stream, int* status);
/* Byte offset of the device status word inside the caller's work buffer (so the host can fetch it together with the force sums,
 * in the same synchronisation, instead of paying a second round trip through lls_get_status). */
int lls_status_offset(lls_scene* h, size_t* offset_bytes);

/* Dense FP64 LU (no pivoting; guarded: a zero / NaN pivot raises LLS_STATUS_NAN, and a pivot more than six orders of
 * magnitude below the largest pivot of the factorisation raises LLS_STATUS_SMALL_PIVOT in the scene's status word, where
 * LAPACK's partial pivoting would have started to swap rows) on a row-major n x n matrix with
 * leading dimension ld, in place, then triangular solves for one right-hand side.
 * Stands in for numpy.linalg.solve -> LAPACK dgesv at scene.py:827 and scene.py:896. */
int lls_lu_factor(lls_scene* h, double* d_mat, void* stream);
int lls_lu_solve(lls_scene* h, const double* d_mat, double* d_rhs_inout, void* stream);

/* Per-phase device timings (ms, CUDA events) of the last lls_solve_nonlinear / lls_solve_linear:
 * [0] influence build, [1] assembly (A or J, summed), [2] LU factor (summed), [3] triangular solves,
 * [4] matvec+residual (summed), [5] integrate.  Only filled when timing is enabled. */
int lls_set_timing(lls_scene* h, int enabled);
int lls_get_timings(lls_scene* h, double* ms6);

/* FP64 roofline probes: out4 = {DFMA TFLOP/s, DMMA m8n8k4 TFLOP/s, DMMA m16n8k16 TFLOP/s, copy GB/s}. */
int lls_probe_peaks(double* out4, void* stream);

/* Last error message (thread-local static buffer). */
const char* lls_last_error(void);

/* Number of CUDA kernels this library has launched since it was loaded (bench.py reports the difference
 * over its timed region as gpu_launches). */
unsigned long long lls_launch_count(void);

/* ABI version, bumped on any signature or table-layout change. */
int lls_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LLS_H */
