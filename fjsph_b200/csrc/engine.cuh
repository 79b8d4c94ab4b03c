// engine.cuh — device data layout and shared helpers of the B200 WCSPH engine (sm_100a only).
//
// Layout (DESIGN.md "Data layout in HBM"): particles live in cell-sorted order; every quantity a pair
// sweep GATHERS from neighbour j is packed in a 32-byte record (one DRAM/L2 sector, two LDG.128):
//   P0 = {x, y, z, V=m/rho}   P1 = {vx, vy, vz, rho}   P2 = {vPert.xyz, p/rho^2}
//   P3 = {gradRho.xyz, lam}   P4 = {norm.xyz, surf}     (norm = Detect_Surface normals, surf as 0.0/1.0)
// V and p/rho^2 are the per-particle quotients the reference recomputes per PAIR (m_j/rho_j in every
// loop, p_j/rho_j^2 in BasePos, Kernel.h:153-157); m_j is recovered as rho_j*V_j.  Quantities only ever
// read for particle i itself are packed the same way (one coalesced 32-byte access per thread):
//   ACC = {acc.xyz, Rrho}  AF = {Af.xyz, deltaD}  AV = {aVisc.xyz, curve}  CV = {cellV.xyz, cellP}
//   NP = {normal.xyz, lam_nb}  BN = {bNorm.xyz, y}  TH = {p, m, woccl, cellRho}
//   SC = {colourG, colour, kernsum, pDist}          L0..L8 = the 3x3 renormalisation matrix
// surf_i mirrors the surf flag of P4 (0 / 1) as an int: the bulk ("lean") shifting sweep reads nothing else of P4, and a
// 4-byte coalesced load costs an eighth of the L1 wavefronts of a strided 8-byte one.
// Field list = SPHPart, reference src/Var.h:499-642.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"

#define FJ_IDX_MASK 0x07FFFFFFu /* 2^27 - 1 particles per engine: a skin run packs first index + length in 32 bits */
#define FJ_RUN_EMPTY 0xFFFFFFFFu
#define FJ_ROW_WARPS 8 /* warps (= rows) of a CTA of the list builds: 4 rows along v x 2 rows along w */

// X(type, name): every per-particle array of one time level
#define FJ_LEVEL_FIELDS(X)                                                                                     \
    X(double4, P0) X(double4, P1) X(double4, P2) X(double4, P3) X(double4, P4)                                 \
    X(double4, ACC) X(double4, AF) X(double4, AV) X(double4, CV) X(double4, NP) X(double4, BN)                 \
    X(double4, TH) X(double4, SC)                                                                              \
    X(double, L0) X(double, L1) X(double, L2) X(double, L3) X(double, L4) X(double, L5) X(double, L6)          \
    X(double, L7) X(double, L8)                                                                                \
    X(long long, part_id) X(int, cellID) X(int, b) X(int, surfzone) X(int, internal) X(int, surf_i)

struct Level
{
#define X(T, n) T* n = nullptr;
    FJ_LEVEL_FIELDS(X)
#undef X
};

// constants the kernels read (subset of FjsphParams, pre-combined where the reference recomputes them)
struct DevConst
{
    double H, H_sq, iH, sr, W_correc, W_dx, iW_dx, gk_fac /* 5*Wc/H^2 */;
    double rho_rest, rho_min, rho_max, B, gam, c2, press_back, Bgam;
    double visc_alpha, nu, dsph_cont, sig, dx, particle_step;
    double gx, gy, gz;
    double vinf_x, vinf_y, vinf_z;
    double lam_cutoff, interp_fac, i_n_full, aero_L, A_sphere, A_plate, mu_g, sos2, gamma_g, ycoef, tab_Cb;
    double max_shift_vel, bnd_mass, sim_mass, c_sound, rho_g;
    int ale, pressure_rel, acase, asource, use_lam, use_TAB_def;
    /* per-pair constants kept in the constant bank so that the sweeps' loops neither rebuild nor hold them in registers */
    double mhalf_iH /* -1 / 2H */, tiny2 /* (1e-12 H)^2 */, eps_f /* 0.001 H^2 */, eps_d /* 0.0001 H^2 */, q_st /* 0.75 / H */;
    int dim; /* 2: SIMDIM=2, every z component exactly 0 (records keep their three components) */
};

// Cell grid in ROW coordinates (u, v, w): u = the row axis (the longest extent of the particles' bounding box, component
// ax0 of a position), v and w the transverse axes (components ax1, ax2).  A ROW is a pencil of cells one particle spacing
// wide along v and w; its particles are contiguous in memory and sorted along u (key = row bits << bx | u-cell).
struct Grid
{
    double ox, oy, oz, inv_cell; // origin in (u, v, w); inv_cell: 1 / cell edge along u (>= 2H + skin: reach 1)
    double inv_cy, inv_cz;       // 1 / cell edge along v and w (~ one particle spacing)
    double pw2, r_skin2;         // (cell edge along v, w)^2 and (2H + skin)^2, to cull rows outside the disc
    int ry, rz;                  // rows to visit either side along v and w
    int nx, ny, nz;              // cells per axis (u, v, w)
    int bx, by, bz;              // key bits per axis; the u bits are the low bx bits of a key
    int ax0, ax1, ax2;           // position component (0 = x, 1 = y, 2 = z) of u, v, w
    unsigned int n_keys;         // 2^(bx+by+bz) keys per class; 2^(by+bz) rows per class
};

// Work mapping of the row sweeps: CTA (group g, chunk c) holds WARPS warps (4, 8 or 16); warp w walks particles
// [32 c, 32 c + 32) of row  row0 + g * WARPS + w.  Rows that are adjacent in v and w have adjacent ids (the low
// row-id bits are v0, w0, v1, w1), so the warps of a CTA sweep the same stretch of a 2 x 2, 4 x 2 or 4 x 4 block of
// neighbouring rows and share what they pull into L1.  warp_start[row] + c numbers the work warps; the run lists are
// laid out by that number, whatever the CTA shape.
struct RowMap
{
    const unsigned* __restrict__ cell_start;
    const unsigned* __restrict__ warp_start;
    unsigned row0, n_rows; // rows [row0, row0 + n_rows) of the table (n_rows a multiple of 16)
    int bx;
    int n_chunks;
};
#ifdef __CUDACC__
// particle and work-warp number of this thread; false for lanes past the end of the row (the whole warp when the row
// holds no chunk c).  i is a valid particle index whenever the WARP has work (clamped to the row's last particle).
template <int WARPS = FJ_ROW_WARPS>
__device__ __forceinline__ bool fj_row_thread(const RowMap& M, int& i, int& W, bool& warp_has_work)
{
    const unsigned chunk = blockIdx.x % unsigned(M.n_chunks), group = blockIdx.x / unsigned(M.n_chunks);
    const unsigned row = M.row0 + group * WARPS + (threadIdx.x >> 5);
    const unsigned s = M.cell_start[size_t(row) << M.bx], e = M.cell_start[size_t(row + 1u) << M.bx];
    const unsigned first = s + chunk * 32u;
    warp_has_work = first < e;
    W = int(M.warp_start[row] + chunk);
    const unsigned ii = first + (threadIdx.x & 31u);
    i = int(warp_has_work ? (ii < e ? ii : e - 1u) : 0u);
    return ii < e;
}
#endif

// X(name): the double4 record arrays of a level, in halo-mask bit order
#define FJ_D4_FIELDS(X) X(P0) X(P1) X(P2) X(P3) X(P4) X(ACC) X(AF) X(AV) X(CV) X(NP) X(BN) X(TH) X(SC)
enum
{
    FJ_HX_P0 = 1 << 0, FJ_HX_P1 = 1 << 1, FJ_HX_P2 = 1 << 2, FJ_HX_P3 = 1 << 3, FJ_HX_P4 = 1 << 4,
    FJ_HX_TH = 1 << 11, FJ_HX_SURFZONE = 1 << 13, FJ_HX_B = 1 << 14
};

// 1-D slab decomposition along x: this rank owns x in [x_lo, x_hi); ghosts are the neighbours'
// particles within 2H + skin of the faces.  Transport is the host's (FjsphCommFn): the engine only packs
// and unpacks on the device.
struct Slab
{
    bool on = false;
    int rank = 0, world = 1;
    double x_lo = -1e300, x_hi = 1e300;
    FjsphCommFn fn = nullptr;
    void* user = nullptr;
    int64_t n_send[2] = {0, 0}, n_recv[2] = {0, 0}; // side 0 = lower-x neighbour, 1 = upper-x neighbour
    int* send_idx[2] = {nullptr, nullptr};          // caller indices of the owned particles sent as ghosts
    int64_t send_cap = 0;
    char* sbuf[2] = {nullptr, nullptr};
    char* rbuf[2] = {nullptr, nullptr};
    size_t buf_bytes = 0;
    unsigned *flag[3] = {nullptr, nullptr, nullptr}, *scan[3] = {nullptr, nullptr, nullptr}; // classify scratch
    int* list[3] = {nullptr, nullptr, nullptr};
    unsigned* scan_tmp = nullptr;
    double n_fluid_global = 0.0, n_total_global = 0.0;
    long long exchanges = 0, redecomps = 0, bytes_sent = 0;
    // Forward exchanges overlapped with interior work: pack, NCCL send/recv and unpack run on comm_stream while the
    // main stream sweeps the INTERIOR owned particles (slots [0, n_interior): farther than 2H + skin from every face
    // with a neighbour rank, so none of their neighbours is a ghost); the EDGE particles [n_interior, n_owned) are
    // swept after the main stream has waited for ev_done.  Every other kernel family waits first (KScope).
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    bool dev_reduce = false;  // the callback implements FJSPH_COMM_SUM_DEV / MAX_DEV (fjsph_slab_device_reductions)
    bool overlap = true;      // FJSPH_SLAB_OVERLAP=0 in the environment: exchanges complete before the next launch
    bool pending = false;     // an exchange is in flight on comm_stream
    bool hold = false;        // set around the interior launches: KScope does not wait
    int64_t n_interior = 0;   // multiple of 32 (whole list warps)
    long long overlapped = 0; // exchanges that ran beside an interior sweep
    // owned particles to drop at the next re-decomposition (escaped from the aero mesh): flags by CALLER index
    const unsigned* del_by_caller = nullptr;
};

// aero mesh on the device (mesh.cu): per-face vertex coordinates, boundary markers, cell -> faces, cell centres and
// solution, and the host-built bins over the cell centres
struct DeviceMesh
{
    bool loaded = false;
    int dim = 3;
    int n_cells = 0, n_faces = 0;
    double4* fx = nullptr;
    int* fmark = nullptr;
    int* fown = nullptr;     // leftright.first (the tracker steps to the cell on the other side of a face, ipt.cu)
    double4* fq = nullptr;   // {face[3].xyz, vertex count}: the fourth corner RayNormalIntersection reads (ipt.cu)
    int *cell_ptr = nullptr, *cell_faces = nullptr;
    double4 *cc = nullptr, *cvp = nullptr;
    double ox = 0, oy = 0, oz = 0, hx = 0, hy = 0, hz = 0, bin = 1;
    int bx = 1, by = 1, bz = 1;
    int *bin_start = nullptr, *bin_cells = nullptr;
    int* counters = nullptr;
};

struct HostBlock
{
    int64_t first, second;
    int is_fluid, bound_solver, no_slip, block_type, fixed_vel_or_dynamic;
    std::vector<double> times;
    std::vector<double> vels; // [max(1,nt)][3]
    double insert_norm[3], insconst, delete_norm[3], delconst, aero_norm[3], aeroconst;
    std::vector<int64_t> back;
    std::vector<std::vector<int64_t>> buffer;
    int *d_back = nullptr, *d_buffer = nullptr; // device copies of back / buffer (caller indices)
};

struct Timer
{
    std::string name;
    double ms = 0.0;
    long long launches = 0;
    long long calls = 0; // timed scopes (one force evaluation = one call)
};
struct PendingTiming
{
    int id;
    cudaEvent_t a, b;
};

struct FjsphEngine
{
    FjsphParams P;
    DevConst C;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int64_t cap = 0;          // particle capacity
    int64_t n = 0;            // particles held (owned + ghosts)
    int64_t n_owned = 0;
    int64_t bound_points = 0;
    int64_t next_part_id = 0;
    int n_sm = 148;

    Level lv[3];              // 0 = pn, 1 = pnp1, 2 = scratch for permutes
    int* oidx = nullptr;      // slot -> caller index
    int* oidx_tmp = nullptr;
    int* slot_of = nullptr;   // caller index -> slot (rebuilt after each sort)
    int* blk = nullptr;       // slot -> block id
    int* blk_tmp = nullptr;

    // neighbour structures
    Grid grid;
    unsigned int* key = nullptr;        // [cap]
    unsigned int* rank_in_cell = nullptr;
    int* perm = nullptr;                // [cap] new slot -> old slot
    int* perm2 = nullptr;
    unsigned int* cell_count = nullptr; // [key_cap]
    unsigned int* cell_start = nullptr; // [key_cap+1]
    unsigned int* scan_tmp = nullptr;
    size_t key_cap = 0;
    unsigned int *mtab_y = nullptr, *mtab_z = nullptr; // key bits of a row's v and w cell coordinates
    int mtab_cap = 0;
    int row_axis = -1;                  // FJSPH_B200_ROW_AXIS: 0 | 1 | 2 pins the row axis (default: the longest extent)
    double row_width_cells = 0.0;       // FJSPH_B200_ROW_WIDTH: row width along v, w in particle spacings (default 1)
    bool list_stats = false;            // FJSPH_B200_LIST_STATS=1: print the fill of the lockstep walk after every build
    int max_key_bits = 25;              // FJSPH_B200_MAX_KEY_BITS: rows widen until the cell table fits 2^bits keys
    // Row-run neighbour lists (neighbours.cu).  A particle's neighbours inside one row are a window of consecutive
    // indices (rows are sorted along u), so a list is one RUN per neighbouring row: {first index, 32-bit membership
    // mask}.  Slot k of work warp W, lane l is element (W * cap + k) * 32 + l: the 32 lanes of a warp -- 32 consecutive
    // particles of one row -- walk the same neighbouring row at the same time and gather consecutive records.
    // Slots that are empty for every lane of a warp are squeezed out (erows / srows = slots in use per warp).
    uint2* erun = nullptr;              // exact list (the reference's OUTL): {first, mask}, mask bit o <-> index first + o
    int* erows = nullptr;               // [n_warp]
    int ecap = 0;
    int* ncount = nullptr;              // [cap] neighbours excluding self
    double4* x0 = nullptr;              // positions the exact list was built on: r = |x0_j - x0_i| is the reference's
                                        // sqrt(jj.second), frozen through the sub-iterations (Resid.cpp:289)
    bool x_moved = true;                // level-1 positions differ from x0 (sweeps then take r from x0)
    // skin list (superset with d < 2H + skin at its build time): {first | (length - 1) << 27}, FJ_RUN_EMPTY = none
    unsigned int* srun = nullptr;
    int* srows = nullptr;
    int scap = 0;
    size_t run_warps_cap = 0;           // work warps the run arrays hold
    unsigned* warp_start = nullptr;     // [rows + 1] exclusive scan of the 32-particle chunks per owned row
    unsigned* row_warps = nullptr;
    unsigned* row_scan_tmp = nullptr;
    size_t row_cap = 0;
    unsigned n_warp = 0;                // work warps of the owned rows
    int n_chunks = 1;                   // ceil(longest owned row / 32)
    int2* row_off = nullptr;            // device table of the neighbouring-row offsets (dv, dw) inside the disc
    int n_row_off = 0;
    double4* xref = nullptr;            // positions at the skin build
    bool skin_valid = false;
    int64_t skin_n = 0;
    double skin = 0.0;                  // skin width (m); 0 = rebuild the cell list at every update_neighbours
    long long skin_builds = 0;
    int* near_inlet = nullptr;          // [cap] Boundary_Ghost flag, valid within one sub-iteration
    bool list_valid = false;
    // fjsph_step_host overlaps the host -> device copy with the first neighbour build: positions go up first on the
    // engine's stream, everything else on upload_stream; upload_pending tells fj_integrate_no_update to run the build
    // ahead of find_timestep and to make the engine's stream wait for ev_upload before anything reads the other fields.
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t ev_upload_x = nullptr, ev_upload_b = nullptr, ev_upload = nullptr;
    bool upload_pending = false;
    int upload_parts = 3;      /* FJSPH_B200_UPLOAD_PARTS=2 keeps the two-part upload */
    bool upload_early = false; /* three-part upload: the prestep may run as soon as x, rho, m, b are in (ev_upload_b) */
    // fused surface / shifting sweep in two launches (lean bulk + near-surface rest) when few warps are near a surface
    // (sweeps.cu, k_surf23_shift CLASS; FJSPH_B200_SPLIT_SURFACE=0 keeps the single launch)
    bool split_surface_sweep = true;
    int sweep_warps = 4;               // FJSPH_B200_SWEEP_WARPS: warps (rows) per CTA of the pair sweeps, 4 (2 x 2 rows) | 8 (4 x 2)
    double split_surface_below = 0.35; // fraction of near-surface warps below which the two launches pay

    // reductions / scalars
    double* red = nullptr;              // device scratch for block partials
    size_t red_cap = 0;
    double* red_out = nullptr;          // device [16]
    double* h_red = nullptr;            // pinned [16]
    int* d_flag = nullptr;              // device error / overflow flags [4]
    int* h_flag = nullptr;              // pinned

    // RK4 accumulators (Runge_Kutta.cpp:363-389): sum of (v+vPert), acc, Rrho over stages
    double4* rk_sum_v = nullptr;        // {sum (v+vPert).xyz, sum Rrho}
    double4* rk_sum_a = nullptr;        // {sum acc.xyz, unused}

    // staging for upload / download
    void* stage = nullptr;
    size_t stage_bytes = 0;

    std::vector<HostBlock> blocks;
    int n_bound_blocks = 0;
    bool inlet_tables_dirty = true;
    unsigned* scan_particles = nullptr; // scan scratch sized for the particle count (delete planes)
    Slab slab;
    DeviceMesh mesh;
    long long mesh_deleted = 0;
    std::vector<FjsphDeleted> deleted;  // particles erased at a delete plane since the last fjsph_take_deleted (IPT hand-off)

    // Integrator members, Integration.h:53-68
    double safe_dt = 0.0, maxf = 0.0, maxAf = 0.0, maxRho_pc = 0.0, maxRhoi = 0.0, maxdrho = 0.0, minST = 0.0,
           maxU = 0.0, maxShift = 0.0;
    unsigned iteration = 0;
    double npd = 1.0;

    // instrumentation
    bool timers_on = false;
    std::vector<Timer> timers;
    std::vector<PendingTiming> pending;   // event pairs recorded on the stream, resolved lazily (no sync per kernel)
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long launches = 0;
    int force_evals = 0, nb_builds = 0;
};

void fj_set_error(const char* fmt, ...);
int fj_cuda_fail(cudaError_t err, const char* what, const char* file, int line);
#define FJ_CUDA(call)                                                   \
    do                                                                  \
    {                                                                   \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess)                                          \
            return fj_cuda_fail(_e, #call, __FILE__, __LINE__);         \
    } while (0)

// RAII-free kernel timing scope: records CUDA events around a kernel family when timers are enabled.
struct KScope
{
    FjsphEngine* e;
    int id;
    KScope(FjsphEngine* e_, const char* name, int launches = 1);
    ~KScope();
};

static inline int fj_blocks(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// stage implementations (each returns FjsphStatus)
int fj_build_neighbours(FjsphEngine* e);
RowMap fj_row_map(const FjsphEngine* e, int first_class, int n_classes); /* rows of the slab classes [first, first + n) */
unsigned fj_row_grid(const RowMap& M, int warps = FJ_ROW_WARPS);         /* CTAs of a row sweep over M */
int fj_owned_classes(const FjsphEngine* e);                              /* 1, or 2 (interior, edge) under slabs */
int fj_neighbours_to_csr(FjsphEngine* e, const long long* d_offsets, long long* d_idx, int* d_bad);
int fj_prestep(FjsphEngine* e, double* npd);
int fj_aero_velocity(FjsphEngine* e);
int fj_surface_and_dissipation(FjsphEngine* e, bool do_surface, bool do_dissipation, bool fuse_shift = false);
int fj_shift(FjsphEngine* e);
int fj_check_pipe_outlet(FjsphEngine* e);
int fj_forces(FjsphEngine* e, int level_idx, double npd);
int fj_walls(FjsphEngine* e, int level_idx, bool nb_comparator);
int fj_nb_iter(FjsphEngine* e, double npd, double* errsum);
int fj_find_timestep(FjsphEngine* e, double* dt);
int fj_integrate_no_update(FjsphEngine* e, FjsphStepStats* s);
int fj_step(FjsphEngine* e, FjsphStepStats* s);
int fj_copy_level(FjsphEngine* e, int dst, int src, int which = 0); /* which: 0 all fields | 1 the prestep's outputs | 2 the rest */
int fj_permute_levels(FjsphEngine* e);
int fj_reduce_sum(FjsphEngine* e, int nblocks, int ncomp, double* out_host);
void fj_refresh_constants(FjsphEngine* e);
void fj_timers_flush(FjsphEngine* e);
// slab decomposition (halo.cu); all are no-ops / identities on a single rank
int fj_halo_exchange(FjsphEngine* e, int level, unsigned mask); /* begin; completed by the next fj_halo_wait */
int fj_halo_wait(FjsphEngine* e);
int fj_upload_wait(FjsphEngine* e); /* engine stream waits for a split upload in flight (abi.cu) */                               /* main stream waits for the exchange in flight */
/* true while an exchange is in flight that an interior/edge split sweep may run beside */
static inline bool fj_halo_overlappable(const FjsphEngine* e)
{
    return e->slab.on && e->slab.pending && e->slab.n_interior > 0 && e->slab.n_interior < e->n_owned;
}
int fj_allreduce(FjsphEngine* e, int op, double* v, int n);
/* in place on a DEVICE array, ordered on e->stream; returns FJSPH_OK with *done = false when the transport cannot */
int fj_allreduce_dev(FjsphEngine* e, int op, double* d_v, int n, bool* done);
int fj_redecompose(FjsphEngine* e);
double fj_fluid_count(FjsphEngine* e);
double fj_total_count(FjsphEngine* e);
// inlet buffer regions and end-of-step bookkeeping (inlet.cu)
bool fj_has_inlets(FjsphEngine* e);
int fj_inlet_motion(FjsphEngine* e, double dt, bool nb_solver, int* n_partials);
int fj_update_data(FjsphEngine* e, int* n_add, int* n_del);
int fj_delete_flagged(FjsphEngine* e, unsigned* d_del_by_caller, bool both_levels, int* n_del, bool hand_off = false);
/* slab re-decomposition: the inlet tables (caller indices) follow their particles into the compacted order */
int fj_inlet_tables_remap(FjsphEngine* e, const unsigned* d_stay_flag, const unsigned* d_stay_scan);
// aero-mesh containment (mesh.cu)
int fj_aero_velocity_mesh(FjsphEngine* e);
int fj_pipe_outlet_mesh(FjsphEngine* e);
void fj_free_mesh(FjsphEngine* e);
