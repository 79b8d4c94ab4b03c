/* fjsph_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C ABI of the CPU oracle: a dependency-free restatement of FJSPH's WCSPH time-step path
 * (reference: /root/reference/src/{IO,Neighbours,Shifting,Geometry,Resid,Aero,Newmark_Beta,
 * Runge_Kutta,Integration}.cpp, Kernel.h, Var.h, shapes/inlet.cpp) and of the particle tracker downstream of it
 * (IPT.cpp, Containment.cpp:896-1079; ipt_oracle.inc).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (fjsph_b200/) never links, imports or calls it.
 *
 * PARITY: the reference ships no tests or golden vectors for this path, and its build as shipped cannot run here
 * (Eigen, nanoflann, TECIO, NetCDF, HDF5 are un-vendored).  What can be done is done: FJSPH's OWN time-step translation
 * units (Neighbours, Shifting, Resid, Geometry, Containment, Newmark_Beta, Runge_Kutta, Integration, shapes/inlet .cpp)
 * compile unmodified against stand-in headers for the two header-only libraries (oracle/shim/, oracle/Makefile.ref ->
 * oracle/_ref/liborc_ref{3d,3d_dsph,2d}.so, same orc_* ABI through oracle/ref_harness.cpp).  This restatement is pinned
 * against them: live in tests/test_oracle_vs_reference.py and through the committed vectors of tests/golden/
 * (tests/test_golden_reference.py) -- every stage and full NB / RK4 steps, walls, aero models, inlets, mesh containment,
 * 2D: same flags, counts and sub-iterations, FP64 fields <= 1e-11.  STILL UNPINNED: the arithmetic INSIDE Eigen and
 * nanoflann (ColPivHouseholderQR, computeDirect, 4x4 determinant, dot/norm summation order, KD-tree result order), which
 * both this file and the stand-ins restate from the published algorithms; and the one path where the reference is
 * undefined (erasing escaped particles by FindCell's duplicated index list), where the contract below stands alone.
 */
#ifndef FJSPH_ORACLE_H
#define FJSPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Particle types, VarDefs.h:92-102 */
enum { ORC_BOUND = 0, ORC_PISTON, ORC_BUFFER, ORC_BACK, ORC_PIPE, ORC_FREE, ORC_OUTLET, ORC_LOST };

/* All settings the path reads (Var.h INTEG_SETT/FLUID/AERO/SIM) followed by the constants
 * Set_Values derives (IO.cpp:26-128). Vectors always have 3 slots; 2D builds ignore [2]. */
typedef struct OrcParams
{
    /* switches */
    int32_t dim;          /* SIMDIM the caller expects; checked against the compiled DIM */
    int32_t ale;          /* 1 = behaviour of the -DALE binary, 0 = the delta-SPH binary */
    int32_t pressure_rel; /* 0 Cole, 1 isothermal (Var.h:203-236) */
    int32_t solver_type;  /* 0 Newmark-Beta, 1 Runge-Kutta */
    int32_t acase;        /* aero_force enum: 0 none, 1 Gissler */
    int32_t asource;      /* aero_source enum: 0 constVel, 1 meshInfl */
    int32_t use_lam;
    int32_t use_TAB_def;
    int32_t max_subits;
    int32_t n_stable, n_stable_limit, n_unstable, n_unstable_limit;
    int32_t reserved0;
    /* inputs */
    double particle_step, H_fac, rho_rest, press_pipe, press_back, rho_max, rho_min, rho_var, rho_max_iter;
    double visc_alpha, speed_sound, mu, sig, gam, dsph_delta;
    double grav[3];
    double v_inf[3];
    double p_ref, rho_g, mu_g, temp_g, R_g, gamma_g, lam_cutoff, i_interp_fac;
    double tab_Cf, tab_Ck, tab_Cd, tab_Cb;
    double cfl, cfl_step, cfl_max, cfl_min, subits_factor, min_residual;
    double delta_t, delta_t_max, delta_t_min, max_shift_vel;
    double current_time, last_frame_time, frame_time_interval;
    /* derived by orc_set_values */
    double B, rho_pipe, dx, sim_mass, bnd_mass, H, H_sq, sr, dsph_cont, nu, W_correc, W_dx, nb_beta, nb_gamma;
    double aero_L, A_sphere, A_plate, td, omega, tmax, Cdef, ycoef, n_full, i_n_full, interp_fac, sos;
} OrcParams;

/* Per-step diagnostics = the columns of the reference's step table (Integration.cpp:250-265). */
typedef struct OrcStepStats
{
    double dt, cfl_ratio, rms_error, maxRho_pc, maxf, maxAf, maxShift, safe_dt, npd, logbase;
    int32_t iterations, n_add, n_del, total_points;
} OrcStepStats;

typedef struct Orc Orc;

void orc_default_params(OrcParams* p, int dim);   /* defaults of Var.h */
void orc_set_values(OrcParams* p);                /* IO.cpp:26-128 */
int orc_compiled_dim(void);

Orc* orc_create(const OrcParams* p);
void orc_destroy(Orc* o);
void orc_get_params(Orc* o, OrcParams* out);
void orc_set_params(Orc* o, const OrcParams* in);

/* Blocks ("limits", Var.h:779-859). Boundary blocks must be added before fluid blocks and cover
 * [0,bound_points) then [bound_points,total) contiguously.  back/buffer may be NULL/0. */
int orc_add_block(Orc* o, int is_fluid, int64_t first, int64_t second, int bound_solver, int no_slip,
                  int block_type, int fixed_vel_or_dynamic, int ntimes, const double* times,
                  const double* vels /* max(1,ntimes) x 3 */, const double* insert_norm, double insconst,
                  const double* delete_norm, double delconst, const double* aero_norm, double aeroconst,
                  int nback, const int64_t* back, int nbuf, const int64_t* buffer /* nback x nbuf */);
void orc_clear_blocks(Orc* o);
int orc_get_block_range(Orc* o, int block, int64_t* first, int64_t* second);

/* State: sets N particles on both time levels (pn = pnp1, Init.cpp:496). Missing optional arrays
 * (NULL) are zero / defaults of the SPHPart constructor (Var.h:502-545). */
int orc_set_particles(Orc* o, int64_t n, int64_t bound_points, const double* xi, const double* v,
                      const double* rho, const double* p, const double* m, const int32_t* b,
                      const int64_t* part_id);
int64_t orc_count(Orc* o);
int64_t orc_bound_points(Orc* o);
/* Generic field access. level 0 = pn, 1 = pnp1. Returns number of doubles/ints per particle, <0 if unknown.
 * Float fields: xi v acc Af aVisc cellV gradRho norm bNorm vPert (dim each) L (dim*dim) Rrho rho p m curve
 * norm_curve woccl pDist deltaD cellP cellRho colourG colour lam lam_nb kernsum y.
 * Int fields (as int64): part_id cellID b surf surfzone internal. */
int orc_get_f64(Orc* o, int level, const char* name, double* out);
int orc_set_f64(Orc* o, int level, const char* name, const double* in);
int orc_get_i64(Orc* o, int level, const char* name, int64_t* out);
int orc_set_i64(Orc* o, int level, const char* name, const int64_t* in);

/* Stage entry points (operate on pnp1 with the current neighbour list unless noted). */
void orc_update_neighbours(Orc* o);                       /* Neighbours.cpp:7-31 on pnp1 */
int64_t orc_neighbour_total(Orc* o);
void orc_get_neighbours(Orc* o, int64_t* offsets /* n+1 */, int64_t* idx, double* d2); /* ascending j */
void orc_neighbour_counts(Orc* o, int64_t* counts /* n */);                            /* restatement builds only */
double orc_prestep(Orc* o);                               /* Shifting.cpp:12-123; returns npd */
void orc_aero_velocity(Orc* o);                           /* Resid.cpp:471-612 (constVel) */
void orc_set_mesh(Orc* o, int64_t n_verts, const double* verts, int64_t n_faces, const int64_t* face_ptr,
                  const int64_t* face_vtx, const int32_t* leftright, int64_t n_cells, const int64_t* cell_ptr,
                  const int64_t* cell_faces, const double* cCentre, const double* cVel, const double* cP,
                  const double* cRho);                     /* MESH, Var.h:396-451 (aero source meshInfl) */
int orc_first_cell_errors(Orc* o);                        /* FirstCell failures (the reference exits) */
void orc_detect_surface(Orc* o);                          /* Geometry.cpp:14-280 */
void orc_dissipation(Orc* o);                             /* Shifting.cpp:126-186 */
void orc_particle_shift(Orc* o);                          /* Shifting.cpp:189-290 */
void orc_forces(Orc* o, double npd);                      /* Resid.cpp:426-469 */
void orc_nb_iter(Orc* o, double npd);                     /* Newmark_Beta.cpp:54-301 */
double orc_find_timestep(Orc* o);                         /* Integration.cpp:370-443 */
double orc_integrate_no_update(Orc* o, OrcStepStats* s);  /* Integration.cpp:27-107 */
double orc_integrate(Orc* o, OrcStepStats* s);            /* Integration.cpp:233-303 */

/* The implicit particle tracker downstream of the delete planes (ipt_oracle.inc).  IPT_SETT (Var.h:313-337) and what
 * IPT::Integrate reads from SIM / MESH beside it. */
typedef struct OrcIptSettings
{
    int32_t eq_order;      /* ipt_eq_order: 1 BFD1, 2 BFD2 */
    int32_t max_subits;    /* svar.integrator.max_subits */
    int32_t record;        /* streak_out == 1 || cells_out == 1: time_record keeps every step */
    int32_t reserved0;
    int64_t max_steps;     /* bound on the cell-to-cell steps of one particle (the reference has none) */
    double relax, n_relax; /* Var.h:323-324 */
    double max_x;          /* svar.ipt.max_x, grid scale applied (IO.cpp:29) */
    double max_length;     /* cells.maxlength (CDFIO.cpp:867-898, 1117-1183) */
    double diam, area;     /* ipt_diam, ipt_area (IO.cpp:126-127) */
    double grav[3], mu_g, rho_rest;
} OrcIptSettings;
/* what IPTPart(SPHPart const&, time, diam, area) copies from the erased particle (Var.h:737-763) */
typedef struct OrcIptStart
{
    int64_t part_id, cellID;
    double t;
    double xi[3], v[3];
    double mass;
    double cellV[3], cellRho;
} OrcIptStart;
/* an IPTPart as the tracker's outputs see it (Write_Point, IPT.cpp:180-190, plus the ids) */
typedef struct OrcIptPoint
{
    int64_t part_id, cellID, faceID; /* faceID -1: c_no_face */
    int32_t going, failed;           /* failed 2: stopped by max_steps */
    double t, dt, acc;
    double xi[3], v[3], cellV[3], cellRho;
} OrcIptPoint;
/* IPT::Integrate (IPT.cpp:871-1107) for n particles on the mesh of orc_set_mesh.  last[n]: pnp1 as Integrate leaves it;
 * records[n][record_cap] / n_records[n]: the time_record Terminate_Particle hands to iptdata; n_steps[n]: cell-to-cell
 * steps taken (restatement builds only; the reference build leaves 0).  Any output pointer may be NULL. */
int orc_ipt_integrate(Orc* o, const OrcIptSettings* s, int64_t n, const OrcIptStart* in, OrcIptPoint* last, int32_t* n_steps,
                      OrcIptPoint* records, int64_t record_cap, int32_t* n_records, int64_t* n_success, int64_t* n_failed);

/* Small-matrix restatements, exposed for unit tests. a is row-major dim x dim. */
int orc_qr_inverse(const double* a, double* inv);          /* returns isInvertible */
double orc_min_eigenvalue(const double* a);                /* SelfAdjointEigenSolver::computeDirect */
double orc_kernel(double r, double H, double Wc);
double orc_get_n_full(double dx, double H);

#ifdef __cplusplus
}
#endif
#endif
