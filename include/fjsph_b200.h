/* fjsph_b200.h — C ABI of the B200-native WCSPH time-step engine.
 *
 * Drop-in boundary for FJSPH's time-step path.  FJSPH has no plugin/FFI layer; the path
 * sits behind ordinary C++ functions called from main's frame loop (reference src/FJSPH.cpp:278-280).
 * Each entry point below names the reference function it replaces.  The binding a FJSPH maintainer
 * would add is shown in INTEGRATION.md.
 *
 * Conventions: opaque handle; every function returns 0 on success and a non-zero FjsphStatus on error
 * (text via fjsph_last_error()); never calls exit().  Single-threaded caller, one simulation per handle.
 * All arrays are caller-owned HOST memory, row-major, FP64 / int32 / int64, copied in/out explicitly.
 * Vectors are [n][3], L is [n][3][3], in the 2D build (dim = 2) too: third components 0.  Particle order is the reference's:
 * boundary blocks first, then fluid blocks (Init.cpp:298-475); the engine re-sorts internally and
 * returns everything in the caller's order.
 */
#ifndef FJSPH_B200_H
#define FJSPH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum FjsphStatus
{
    FJSPH_OK = 0,
    FJSPH_ERR_INVALID = 1,     /* bad argument / unsupported compile-time option of the reference */
    FJSPH_ERR_CUDA = 2,        /* CUDA runtime failure (no device, out of memory, launch error) */
    FJSPH_ERR_STATE = 3,       /* call order (e.g. stage before upload / neighbour build) */
    FJSPH_ERR_CAPACITY = 4,    /* particle / neighbour / cell-table capacity exceeded */
    FJSPH_ERR_IO = 5           /* settings file problems */
} FjsphStatus;

/* partType, VarDefs.h:92-102 */
enum { FJSPH_BOUND = 0, FJSPH_PISTON, FJSPH_BUFFER, FJSPH_BACK, FJSPH_PIPE, FJSPH_FREE, FJSPH_OUTLET, FJSPH_LOST };
/* bound_solve_type, VarDefs.h:142-147 */
enum { FJSPH_DBC = 0, FJSPH_PRESSURE_G = 1, FJSPH_GHOST = 2 };
/* shape_type inletZone, VarDefs.h:118-129 */
enum { FJSPH_INLET_ZONE = 6 };

/* Every setting the path reads (Var.h INTEG_SETT / FLUID / AERO / SIM), then the constants that
 * Set_Values derives from them (IO.cpp:26-128).  fjsph_set_values() fills the second half. */
typedef struct FjsphParams
{
    /* switches */
    int32_t dim;          /* SIMDIM (VarDefs.h:13-41): 3 or 2.  In 2D every view below keeps its [n][3] / [n][3][3] shape with
                             the third components 0 (L: third row and column of the identity); a 2D engine takes 2D aero meshes
                             (faces = edges of two vertices, every z = 0) */
    int32_t ale;          /* 1 = the -DALE binary (shifting, surfzone-gated ST), 0 = the delta-SPH binary */
    int32_t pressure_rel; /* 0 Cole, 1 isothermal (Var.h:203-236, IO.cpp:397) */
    int32_t solver_type;  /* 0 Newmark-Beta, 1 Runge-Kutta (IO.cpp:393) */
    int32_t acase;        /* aero_force: 0 none, 1 Gissler (IO.cpp:432) */
    int32_t asource;      /* aero_source: 0 constVel, 1 meshInfl */
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
    /* derived by fjsph_set_values (IO.cpp:26-128, Var.h:244-266, Geometry.cpp:282-308) */
    double B, rho_pipe, dx, sim_mass, bnd_mass, H, H_sq, sr, dsph_cont, nu, W_correc, W_dx, nb_beta, nb_gamma;
    double aero_L, A_sphere, A_plate, td, omega, tmax, Cdef, ycoef, n_full, i_n_full, interp_fac, sos;
} FjsphParams;

/* One block of the reference's LIMITS vector (bound_block, Var.h:779-859). */
typedef struct FjsphBlock
{
    int64_t first, second;        /* index range [first, second) in the caller's particle order */
    int32_t is_fluid;             /* 0 boundary block, 1 fluid block */
    int32_t bound_solver;         /* FJSPH_DBC / FJSPH_PRESSURE_G / FJSPH_GHOST */
    int32_t no_slip;
    int32_t block_type;           /* FJSPH_INLET_ZONE for inlets */
    int32_t fixed_vel_or_dynamic;
    int32_t n_times;              /* 0: static velocity vels[0..2] */
    const double* times;          /* [n_times] */
    const double* vels;           /* [max(1,n_times)][3] */
    double insert_norm[3], insconst;
    double delete_norm[3], delconst;
    double aero_norm[3], aeroconst;
    int32_t n_back, n_buf;
    const int64_t* back;          /* [n_back] */
    const int64_t* buffer;        /* [n_back][n_buf] */
} FjsphBlock;

/* Host view of one time level: the field list of SPHPart (Var.h:499-642).  NULL pointers are skipped
 * (upload: keep the device value / default; download: do not fetch). */
typedef struct FjsphStateView
{
    int64_t n;
    int64_t* part_id;
    int64_t* cellID;
    int32_t *b, *surf, *surfzone, *internal;
    double *xi, *v, *acc, *Af, *aVisc, *cellV, *gradRho, *norm, *bNorm, *vPert; /* [n][3] (dim 2: z = 0) */
    double* L;                                                                   /* [n][3][3] */
    double *Rrho, *rho, *p, *m, *curve, *norm_curve, *woccl, *pDist, *deltaD, *cellP, *cellRho, *colourG, *colour,
        *lam, *lam_nb, *kernsum, *y; /* [n] */
} FjsphStateView;

/* Columns of the reference's per-step table (Integration.cpp:250-265) plus counters. */
typedef struct FjsphStepStats
{
    double dt, cfl_ratio, rms_error, maxRho_pc, maxf, maxAf, maxShift, safe_dt, npd, logbase;
    int32_t iterations, n_add, n_del, total_points;
    int32_t force_evals, neighbour_builds, kernel_launches;
    int32_t skin_builds; /* cell-list sweeps behind those neighbour builds (the rest were filtered from the skin list) */
} FjsphStepStats;

/* What the reference hands to its particle tracker when a particle passes its block's delete plane: the fields
 * IPTPart(SPHPart const&, time, diam, area) copies (Var.h:733-763; Integration.cpp:151-169).  Diameter and area are IPT
 * settings of the host (ipt_diam, ipt_area). */
typedef struct FjsphDeleted
{
    int64_t part_id, cellID;
    double t;        /* integrator.current_time at the hand-off */
    double xi[3], v[3];
    double mass;
    double cellV[3], cellRho;
} FjsphDeleted;

/* IPT_SETT (Var.h:313-337) and what IPT::Integrate reads from SIM / MESH beside it (IPT.cpp:871-1107). */
typedef struct FjsphIptSettings
{
    int32_t eq_order;      /* ipt_eq_order: 1 BFD1, 2 BFD2 (IO.cpp:448, 668-672) */
    int32_t max_subits;    /* svar.integrator.max_subits */
    int32_t record;        /* streak_out == 1 || cells_out == 1: keep the state after every cell-to-cell step */
    int32_t reserved0;
    int64_t max_steps;     /* bound on the steps of one particle; the reference loops `while (pnp1.going != 0)` unbounded */
    double relax, n_relax; /* Var.h:323-324 (n_relax is a real there) */
    double max_x;          /* svar.ipt.max_x with the grid scale applied (IO.cpp:29) */
    double max_length;     /* cells.maxlength: fjsph_mesh_max_length */
    double diam, area;     /* ipt_diam, ipt_area (IO.cpp:126-127) */
    double grav[3], mu_g, rho_rest;
} FjsphIptSettings;

/* An IPTPart (Var.h:645-813) as the tracker's outputs see it: the columns of Write_Point (IPT.cpp:180-190) and the ids. */
typedef struct FjsphIptPoint
{
    int64_t part_id, cellID, faceID; /* cellID < 0: the boundary marker it left through; faceID -1: c_no_face */
    int32_t going, failed;           /* failed 1: as the reference fails a particle; 2: stopped by max_steps */
    double t, dt, acc;
    double xi[3], v[3], cellV[3], cellRho;
} FjsphIptPoint;

/* The reference's MESH (Var.h:396-451) as plain arrays: vertices, faces as vertex lists (CSR), leftright (owner cell,
 * neighbour cell or boundary marker: -1 inner wall, -2 outer boundary), cell -> faces (CSR), cell centres and the
 * cell-averaged CFD solution.  What TAU::Read_* (CDFIO.cpp:1103-1356) or FOAM::Read_FOAM (FOAMIO.cpp:538-955) fill. */
typedef struct FjsphMesh
{
    int64_t n_verts;
    const double* verts;        /* [n_verts][3] */
    int64_t n_faces;
    const int64_t* face_ptr;    /* [n_faces+1] */
    const int64_t* face_vtx;
    const int32_t* leftright;   /* [n_faces][2] */
    int64_t n_cells;
    const int64_t* cell_ptr;    /* [n_cells+1] */
    const int64_t* cell_faces;
    const double* cCentre;      /* [n_cells][3] */
    const double* cVel;         /* [n_cells][3] */
    const double* cP;           /* [n_cells] */
    const double* cRho;         /* [n_cells] */
} FjsphMesh;

typedef struct FjsphEngine FjsphEngine;

const char* fjsph_last_error(void);
const char* fjsph_version(void);

/* Host-side restatement of the settings path: Var.h defaults; GetInput's `key : value` para parser
 * (IO.cpp:305-456, the keys of SURVEY Appendix B); Set_Values (IO.cpp:26-128). */
int fjsph_default_params(FjsphParams* p, int dim);
int fjsph_read_para(const char* path, FjsphParams* p, char* fluid_file, char* bound_file, int name_cap);
int fjsph_set_values(FjsphParams* p);

/* Lifetime.  device = CUDA ordinal.  capacity = max particles this rank will ever hold (>= n). */
int fjsph_create(const FjsphParams* p, int device, int64_t capacity, FjsphEngine** out);
int fjsph_destroy(FjsphEngine* e);
int fjsph_get_params(FjsphEngine* e, FjsphParams* out);
int fjsph_set_params(FjsphEngine* e, const FjsphParams* in);

/* LIMITS (Init.cpp:298-475).  Optional: without it one pressure_G wall block [0,bound_points) and one
 * fluid block [bound_points,n) are assumed. */
int fjsph_set_blocks(FjsphEngine* e, int32_t n_blocks, const FjsphBlock* blocks);

/* pn / pnp1 (FJSPH.cpp:49-58).  fjsph_upload_state sets BOTH time levels from `s` (pn = pnp1,
 * Init.cpp:496); fjsph_upload_level overwrites the given fields of one level (0 = pn, 1 = pnp1). */
int fjsph_upload_state(FjsphEngine* e, const FjsphStateView* s, int64_t bound_points);
int fjsph_upload_level(FjsphEngine* e, int level, const FjsphStateView* s);
int fjsph_download_state(FjsphEngine* e, int level, FjsphStateView* s);
int64_t fjsph_count(FjsphEngine* e);

/* Aero mesh for aero source meshInfl (asource = 1): replaces the MESH argument and the cell-centre KD-tree
 * (Vec_Tree CELL_TREE, FJSPH.cpp:148-149) of get_aero_velocity / FindCell / FirstCell / Check_Pipe_Outlet
 * (Resid.h:43-46, Containment.h:20-31). */
int fjsph_upload_mesh(FjsphEngine* e, const FjsphMesh* m);

/* Stage entry points; each replaces the reference function named on the right and acts on pnp1. */
int fjsph_build_neighbours(FjsphEngine* e);                     /* update_neighbours      Neighbours.h:9 */
int fjsph_neighbour_counts(FjsphEngine* e, int64_t* counts);    /* outlist[i].size() incl. self */
int fjsph_get_neighbours(FjsphEngine* e, const int64_t* offsets, int64_t* idx); /* CSR, ascending j, self incl. */
int fjsph_prestep(FjsphEngine* e, double* npd);                 /* dSPH_PreStep           Shifting.h:10 */
int fjsph_aero_velocity(FjsphEngine* e);                        /* get_aero_velocity      Resid.h:43-46 */
int fjsph_detect_surface(FjsphEngine* e);                       /* Detect_Surface         Geometry.h:98-101 */
int fjsph_dissipation(FjsphEngine* e);                          /* dissipation_terms      Shifting.h:13-15 */
int fjsph_shift(FjsphEngine* e);                                /* particle_shift         Shifting.h:18-20 */
int fjsph_forces(FjsphEngine* e, double npd);                   /* get_acc_and_Rrho       Resid.h:39-41 */
int fjsph_nb_iter(FjsphEngine* e, double npd, double* errsum);  /* Newmark_Beta::Do_NB_Iter + the sum of
                                                                   Check_Error, Newmark_Beta.h:11-26 */
int fjsph_find_timestep(FjsphEngine* e, double* dt);            /* Integrator::find_timestep */
int fjsph_integrate_no_update(FjsphEngine* e, FjsphStepStats* s); /* Integration.h:25-28 */
int fjsph_step(FjsphEngine* e, FjsphStepStats* s);              /* Integrator::integrate  Integration.h:20-23 */

/* IPT hand-off: the particles erased at a delete plane (Integration.cpp:127-169) since the last call, in the reference's
 * order (ascending index within a step, steps in sequence) -- what update_data turns into IPTPart objects before it
 * erases them.  Copies up to `capacity` records into `out` and drops them from the engine's queue; with out == NULL it
 * only reports how many are waiting.  Under slab decomposition every rank holds the particles erased on it. */
int fjsph_take_deleted(FjsphEngine* e, FjsphDeleted* out, int64_t capacity, int64_t* n_out);

/* Implicit particle tracking of the particles handed over at a delete plane: IPT::Integrate (IPT.cpp:871-1107) with
 * FindFace / CheckCellFace (Containment.cpp:896-1079), Cross_Plane, MollerTrumbore and RayNormalIntersection
 * (Geometry.cpp:399-478, 579-744), one device thread per particle, on the mesh of fjsph_upload_mesh -- what update_data
 * does with to_del when `using_ipt` is set and the aero source is a mesh (Integration.cpp:151-169).
 *   fjsph_ipt_default_settings  IPT_SETT's defaults, ipt_diam / ipt_area from the simulation mass (IO.cpp:126-127), gravity,
 *                               gas viscosity, rest density and max_subits from `p`
 *   fjsph_read_para_ipt         the para keys of IO.cpp:447-453; max_x is multiplied by `scale` (IO.cpp:29); *using_ipt is
 *                               cleared when max_x < the SPH conversion coordinate (IO.cpp:674-679)
 *   fjsph_mesh_max_length       cells.maxlength as the TAU readers leave it: 5 x the longest edge of a triangle / longer
 *                               diagonal of any other face (CDFIO.cpp:1117-1183,1214); 4 x the longest edge in 2D (CDFIO.cpp:867-898,931).
 *                               FOAM::Read_FOAM never sets it (0: every particle would fail its first step)
 *   fjsph_ipt_integrate         last[n]: pnp1 as Integrate leaves it; records[n][record_cap], n_records[n]: the time_record
 *                               Terminate_Particle hands to iptdata (n_records counts them all, also those beyond
 *                               record_cap, which are dropped); n_steps[n]: cell-to-cell steps taken.  Any output may be NULL.
 * The surface-impact tallies of Terminate_Particle (IPT.cpp:746-798) follow from last[i].faceID and the host's marker table. */
int fjsph_ipt_default_settings(const FjsphParams* p, FjsphIptSettings* s);
int fjsph_read_para_ipt(const char* path, double scale, int32_t* using_ipt, FjsphIptSettings* s);
int fjsph_mesh_max_length(const FjsphMesh* m, int32_t dim, double* max_length);
int fjsph_ipt_integrate(FjsphEngine* e, const FjsphIptSettings* s, int64_t n, const FjsphDeleted* in, FjsphIptPoint* last,
                        int32_t* n_steps, FjsphIptPoint* records, int64_t record_cap, int32_t* n_records, int64_t* n_success,
                        int64_t* n_failed);

/* Convenience for hosts that keep particles on the host between steps (the end-to-end path):
 * upload -> n_steps x fjsph_step -> download, one call.  When `in` holds the particle set the engine already has (same
 * count and blocks), the cell order and the superset neighbour list are kept and the upload is split: positions first,
 * the other fields beside the first update_neighbours of the step.  `in` must stay untouched until the call returns;
 * `out` may alias `in`. */
int fjsph_step_host(FjsphEngine* e, const FjsphStateView* in, int64_t bound_points, int32_t n_steps,
                    FjsphStateView* out, FjsphStepStats* last);

/* Instrumentation: CUDA-event time (ms) and launch count per kernel family since the last reset. */
int fjsph_timers_reset(FjsphEngine* e);
int fjsph_timers_enable(FjsphEngine* e, int on);
int fjsph_timers_get(FjsphEngine* e, int32_t cap, char* names /* cap x 32 */, double* ms, int64_t* launches,
                     int64_t* calls /* timed scopes, may be NULL */, int32_t* n_out);
int64_t fjsph_launch_count(FjsphEngine* e);

/* Run on the caller's CUDA stream (a cudaStream_t, e.g. the host framework's current stream) instead of the
 * engine's own; NULL goes back to a private stream.  The reference is single-threaded host code with no notion
 * of streams (FJSPH.cpp:262-283); this is what lets a host time or order the engine with its own events. */
int fjsph_set_stream(FjsphEngine* e, void* cuda_stream);

/* Neighbour-build policy.  update_neighbours (Neighbours.cpp:7-31) rebuilds the KD-tree and searches it at every
 * call; the engine instead keeps a superset ("skin") list of every j within 2H + skin and filters the exact
 * list { j : d2 < 4H^2 } from it with the bit-exact distance test at every call, falling back to a cell-list
 * sweep when some PAIR may have closed in by more than skin since the superset was built: some particle has moved
 * more than skin/2 AND the displacements do not all lie within skin/2 of a common drift (a jet moving as a whole
 * displaces every particle but no pair, and keeps its superset list).  The neighbour sets are identical either way.  skin_over_dx = 0 sweeps the cell list at every call; default 0.4. */
int fjsph_set_skin(FjsphEngine* e, double skin_over_dx);

/* Slab decomposition.  The reference is one shared-memory process (OpenMP only, FJSPH.cpp:62); large
 * cases are split into x-slabs, one engine per GPU.  Rank r owns the particles with x in [x_lo, x_hi) (the end
 * ranks pass -/+1e300).  The engine selects, packs and unpacks migrating and ghost particles on the device; the
 * transport belongs to the host, which supplies one callback used for every exchange:
 *   op FJSPH_COMM_SUM / MAX : all-reduce of the host array a (na/8 doubles) in place;
 *   op FJSPH_COMM_SENDRECV_DEV / _HOST : send a (na bytes) to rank-1 and b (nb bytes) to rank+1, receive c (nc bytes)
 *      from rank-1 and d (nd bytes) from rank+1; device or host pointers respectively; zero sizes at the domain ends.
 *   op FJSPH_COMM_SENDRECV_DEV_ASYNC : the same on device pointers, but ordered on the engine's COMM stream
 *      (fjsph_slab_comm_stream) instead of its main stream: the forward halo exchanges run there while the main
 *      stream sweeps the interior particles.  The callback must not block the host on it.
 * Blocks under slabs: every rank passes the SAME block list to fjsph_set_blocks with ranges over its own particles; an
 * inlet block carries its back / buffer tables (local indices) only on the rank that holds its buffer region, which must
 * lie inside one slab.  Insertions, delete planes and the erasures of the aero-mesh lookup then work as on one GPU, with
 * globally unique particle ids; the particle ORDER is this rank's own.  A replicated aero mesh is uploaded on every rank.
 * The callback returns 0 on success.  fjsph_b200/slab.py implements it with NCCL send/recv over NVLink
 * (torch.distributed); upload the rank's own particles with fjsph_upload_state first, then call fjsph_set_slab. */
enum { FJSPH_COMM_SUM = 0, FJSPH_COMM_MAX = 1, FJSPH_COMM_SENDRECV_DEV = 2, FJSPH_COMM_SENDRECV_HOST = 3,
       FJSPH_COMM_SENDRECV_DEV_ASYNC = 4,
       /* all-reduce of the DEVICE array a (na/8 doubles) in place, ordered on the engine's MAIN stream (fjsph_get_stream):
          used for the per-sub-iteration residual, npd and the time-step maxima when fjsph_slab_device_reductions is on --
          the scalar then crosses PCIe once, after the all-reduce, instead of host -> device -> NCCL -> host */
       FJSPH_COMM_SUM_DEV = 5, FJSPH_COMM_MAX_DEV = 6 };
typedef int (*FjsphCommFn)(void* user, int32_t op, void* a, int64_t na, void* b, int64_t nb, void* c, int64_t nc,
                           void* d, int64_t nd);
int fjsph_set_slab(FjsphEngine* e, int32_t rank, int32_t world, double x_lo, double x_hi, FjsphCommFn fn, void* user);
/* the cudaStream_t FJSPH_COMM_SENDRECV_DEV_ASYNC exchanges must be ordered on (valid after fjsph_set_slab) */
int fjsph_slab_comm_stream(FjsphEngine* e, void** stream);
/* the engine's main stream (its own, or the one given to fjsph_set_stream): FJSPH_COMM_*_DEV reductions are ordered on it */
int fjsph_get_stream(FjsphEngine* e, void** stream);
/* tell the engine that the callback implements FJSPH_COMM_SUM_DEV / MAX_DEV (off by default: host-array reductions) */
int fjsph_slab_device_reductions(FjsphEngine* e, int32_t on);
/* forward exchanges that ran beside an interior sweep since fjsph_set_slab */
int fjsph_slab_overlapped(FjsphEngine* e, int64_t* n);
/* owned / ghost particle counts, halo exchanges, re-decompositions and bytes sent since fjsph_set_slab */
int fjsph_slab_stats(FjsphEngine* e, int64_t* n_owned, int64_t* n_ghost, int64_t* exchanges, int64_t* redecomps,
                     int64_t* bytes_sent);
/* Slab mode: overwrite the given fields of BOTH time levels of the owned particles (s->n == fjsph_count, in the
 * order fjsph_download_state returns them), keeping the decomposition; the end-to-end path of a host that holds
 * the particles between steps.  Positions may have changed: the neighbour lists are rebuilt by the next step. */
int fjsph_upload_owned(FjsphEngine* e, const FjsphStateView* s);
int fjsph_set_owned(FjsphEngine* e, int64_t n_owned);

/* OpenFOAM case ingestion (host only): FOAM::Read_FOAM for ASCII and binary cases (reference src/FOAMIO.cpp:21-342,346-955) -- constant/
 * polyMesh/{boundary,points,faces,owner,neighbour} and <solution_dir>/{p or p_rgh,U} -> the MESH arrays fjsph_upload_mesh
 * takes: faces fanned into triangles, boundary markers -1 (wall patches) / -2 (other patches), the reference's cell
 * centres.  The reference never fills cRho: it is set to rho_fill.  solution_dir NULL or "" = mesh only
 * (zero velocity and pressure).  The view's pointers live as long as the FjsphFoamMesh. */
typedef struct FjsphFoamMesh FjsphFoamMesh;
int fjsph_foam_read(const char* foam_dir, const char* solution_dir, int buoyant, double rho_fill, FjsphFoamMesh** out);
/* TAU case ingestion (host only, no NetCDF library): TAU::Read_tau_mesh_FACE + TAU::Read_SOLUTION (reference src/CDFIO.cpp:
 * 1228-1356, 655-822; FJSPH.cpp:76-78) on the face-based mesh file FJSPH's Cell2Face writes and a TAU solution file, both
 * NetCDF-3 classic (CDF-1 / CDF-2).  Faces = triangles then quadrilaterals (kept four-cornered), right cell < 0 = the file's
 * boundary marker, cell values = Kahan-summed means of the point data over the cell's vertices; coordinates times `scale`
 * ("Grid scale").  solution_file NULL or "" = mesh only.  Same handle type as fjsph_foam_read (view / free below). */
int fjsph_tau_read(const char* mesh_file, const char* solution_file, double scale, FjsphFoamMesh** out);
/* TAU 2D case ingestion (the reference's -DSIMDIM=2 build): TAU::Read_tau_mesh_EDGE + TAU::Read_SOLUTION (reference src/CDFIO.cpp:
 * 992-1097, 828-990, 655-822; FJSPH.cpp:85-91) on the edge-based mesh FJSPH's Cell2Edge writes.  Faces = edges; the mesh's
 * plane is named by the coordinate variable the file lacks; `vertices_in_use` maps mesh points to solution points;
 * offset_axis (the para's "2D offset vector": 1 = x, 2 = y, 3 = z) picks the two velocity components.  The view is in the
 * ABI's shape with z = 0, for fjsph_upload_mesh on a 2D engine. */
int fjsph_tau_read_edge(const char* mesh_file, const char* solution_file, double scale, int32_t offset_axis, FjsphFoamMesh** out);
int fjsph_foam_view(const FjsphFoamMesh* m, FjsphMesh* view);
void fjsph_foam_free(FjsphFoamMesh* m);

/* Checkpoint / resume: a raw-binary mirror of the reference's <prefix>_particles.h5 restart data (H5IO.cpp:395-538,
 * 915-962; HDF5 is not available here): position, velocity, acceleration, pressure, density, density gradient, mass,
 * boundary condition, particle ID, cell ID, cell velocity / density / pressure under the reference's dataset names, the
 * LIMITS blocks with their inlet back / buffer tables, every setting (FjsphParams incl. current time, previous frame
 * time and the CFL controller state) and the particle index to add.  fjsph_read_restart restores pn = pnp1 from it, as
 * Read_HDF5 does; the next step rebuilds the neighbour lists and the frozen terms.  Layout: csrc/restart.cu. */
int fjsph_write_restart(FjsphEngine* e, const char* path, int32_t frame);
int fjsph_read_restart(FjsphEngine* e, const char* path, int32_t* frame);

/* ------------------------------------------------------------------------------------------------------------
 * Case front end (host only): GetInput + Init_Particles (reference src/IO.cpp:305-723, src/Init.cpp:270-496,
 * src/shapes/{shapes,line,square,circle,cylinder,arc,inlet,coordinates}.cpp).  Reads a FJSPH para file and the fluid /
 * boundary block files (bmap) it names, generates every block's particles with the reference's perturbation stream
 * (std::default_random_engine through uniform_real_distribution(0, eps dx)), removes intersecting particles
 * (Check_Intersection) and lays the particles out in the reference's order: boundary blocks first, then fluid blocks,
 * inlet blocks as PIPE layers | BACK row | BUFFER rows with their back / buffer tables.  dim is the SIMDIM of the
 * build the deck was written for (shape names differ: Line/Plane, Square/Cube, Circle/Sphere); xi and v of
 * fjsph_case_state are [n][dim].  Arc / Arch blocks (shapes/arc.cpp) and JSON block files (the reference's JSON keys, blocks
 * in key order) are read as well; what arc.cpp answers with exit() is an error return here.  The aero source follows
 * IO.cpp:465-533: a TAU mesh, else an OpenFOAM case (asource 1; file names the working directory does not hold are looked for
 * beside the para file), else in 3D a VLM definition (asource 2, which fjsph_create refuses: the vortex lattice is out of
 * scope), else the constant free stream. */
typedef struct FjsphCase FjsphCase;
int fjsph_case_read(const char* para_path, int dim, FjsphCase** out);
void fjsph_case_free(FjsphCase* c);
int64_t fjsph_case_count(const FjsphCase* c);
int64_t fjsph_case_bound_points(const FjsphCase* c);
int32_t fjsph_case_num_blocks(const FjsphCase* c);
int32_t fjsph_case_dim(const FjsphCase* c);
int32_t fjsph_case_offset_axis(const FjsphCase* c); /* "2D offset vector" (IO.cpp:358,686-705): 1 = x, 2 = y, 3 = z; 0 in 3D */
int fjsph_case_params(const FjsphCase* c, FjsphParams* out);               /* after Set_Values */
/* run control of the frame loop (FJSPH.cpp:262-330): "SPH frame count", "SPH maximum particle count" (-1 when the deck
 * does not set them), "Output files prefix", "SPH restart prefix" */
/* "OpenFOAM input directory" / "OpenFOAM solution directory" / "OpenFOAM buoyant (0/1)" of the deck ("" when it names
 * no mesh; naming one sets params.asource = meshInfl, IO.cpp:464-499) */
int fjsph_case_foam(const FjsphCase* c, char* foam_dir, char* solution_dir, int32_t* buoyant, int32_t cap);
/* "Primary grid face filename" / "Restart-data prefix" / "Grid scale" of the deck ("" when it names no TAU mesh; naming one
 * sets params.asource = meshInfl, needs "Boundary mapping filename" (IO.cpp:494-533), and turns gravity by the angle of
 * attack of the para or the boundary map as TAU::Read_BMAP does, CDFIO.cpp:234-315): for fjsph_tau_read */
int fjsph_case_tau(const FjsphCase* c, char* mesh_file, char* solution_file, double* scale, int32_t cap);
int fjsph_case_io(const FjsphCase* c, int32_t* max_frames, int64_t* max_points, char* output_prefix, char* restart_prefix,
                  int32_t cap);
int fjsph_case_block(const FjsphCase* c, int32_t i, FjsphBlock* out, char* name, int32_t name_cap); /* LIMITS[i]; the
                                                                              pointers live as long as the case */
int fjsph_case_state(const FjsphCase* c, FjsphStateView* s);               /* fills xi, v, rho, p, m, b, part_id */

#ifdef __cplusplus
}
#endif
#endif
