/* ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * The orc_* C ABI of fjsph_oracle.h on top of FJSPH's OWN sources, compiled unmodified from where they lie under
 * /root/reference/src: the time step (Neighbours, Shifting, Resid, Geometry, Containment, Newmark_Beta, Runge_Kutta,
 * Integration .cpp) and the callers either side of it (IO.cpp: GetInput / Set_Values; Init.cpp: Init_Particles;
 * shapes/{shapes,square,circle,cylinder,line,coordinates,inlet}.cpp: the bmap reader and the generators; FOAMIO.cpp: the
 * OpenFOAM reader).  Recipe: oracle/Makefile.ref, output: oracle/_ref/liborc_ref*.so.  The two header-only libraries
 * the reference does not vendor (Eigen, nanoflann) are replaced by the stand-ins of oracle/shim/, so:
 *   - what this library pins is FJSPH's arithmetic and control flow (pair loops, surface logic, boundary treatment,
 *     aero coupling, containment, integrators, inlet bookkeeping) -- the oracle restatement is checked against it in
 *     tests/test_oracle_vs_reference.py and through the fixtures of tests/golden/;
 *   - Eigen's / nanoflann's own arithmetic (QR inverse, direct eigenvalues, 4x4 determinant, search order) is the
 *     shim's restatement of the published algorithms and stays unpinned.
 * Nothing in fjsph_b200/ links or loads this.  Functions of the reference that the path links but never runs here
 * (the VLM, the Tecplot-binary and h5part writers) are stubs that abort.
 *
 * OpenMP: the reference's loops keep their pragmas; the harness pins one thread so that reductions are deterministic
 * and the npd data race cannot occur.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <chrono>
#include <set>

#include "Var.h"
#define private public /* Integrator keeps find_timestep and the step maxima private (Integration.h:30-75) */
#include "Integration.h"
#undef private
#include "Containment.h"
#include "Geometry.h"
#include "Kernel.h"
#include "Neighbours.h"
#include "Newmark_Beta.h"
#include "Resid.h"
#include "Runge_Kutta.h"
#include "Shifting.h"
#include "BinaryIO.h"
#include "CDFIO.h"
#include "FOAMIO.h"
#include "H5IO.h"
#include "IO.h"
#include "IPT.h"
#include "Init.h"
#include "shapes/arc.h"
#include "shapes/inlet.h"

#include <filesystem>
#include <unistd.h>

#include "fjsph_oracle.h"

/* ---- out-of-path symbols the objects reference ------------------------------------------------ */
static void not_on_path(const char* what)
{
    std::fprintf(stderr, "ref_harness: %s is outside the time-step path and is not compiled\n", what);
    std::abort();
}
/* IPT.cpp is compiled (IPT::Integrate runs in orc_ipt_integrate); its Tecplot-binary writers live in BinaryIO.cpp */
namespace IPT
{
namespace BINARY
{
void Init_IPT_Files(SIM&) { not_on_path("IPT::BINARY::Init_IPT_Files"); }
void Write_Point(SIM const&, IPTPart const&) { not_on_path("IPT::BINARY::Write_Point"); }
void Write_State(SIM const&, IPTState const&, string const&, void*&) { not_on_path("IPT::BINARY::Write_State"); }
void Write_Cells(SIM&, MESH const&, IPTState const&, int32_t const&, vector<StateVecD> const&, vector<int32_t> const&,
                 vector<vector<int32_t>> const&, vector<int32_t> const&, vector<int32_t> const&)
{
    not_on_path("IPT::BINARY::Write_Cells");
}
} // namespace BINARY
} // namespace IPT
void flush_file(void*) { not_on_path("flush_file"); }
#if SIMDIM == 3
StateVecD VLM::getVelocity(StateVecD const&) const
{
    not_on_path("VLM::getVelocity");
    return StateVecD::Zero();
}
/* Set_Values sets the lattice up when a deck names the VLM aero source (IO.cpp:101-108): nothing to set up here, so that
   GetInput + Init_Particles of such a deck (Examples/VC10, Examples/VLM) can still be compared; asking it for a velocity aborts */
void VLM::Init(std::string) {}
void VLM::GetGamma(StateVecD) {}
void VLM::write_VLM_Panels(std::string&) {}
void VLM::Plot_Streamlines(std::string&) {}
#endif
/* TECIO / HDF5 writers (BinaryIO.cpp, H5IO.cpp need the absent libraries) */
void Write_Binary_Timestep(SIM const&, real const&, SPHState const&, bound_block const&, char const*, int32_t const&,
                           void* const&)
{
    not_on_path("Write_Binary_Timestep");
}
void Init_Binary_PLT(SIM&, std::string const&, std::string const&, std::string const&, void*&) { not_on_path("Init_Binary_PLT"); }
void close_file(void*) { not_on_path("close_file"); }
void open_h5part_files(SIM&, string const&) { not_on_path("open_h5part_files"); }
void close_h5part_files(SIM&) { not_on_path("close_h5part_files"); }
void write_h5part_data(SIM&, SPHState const&) { not_on_path("write_h5part_data"); }

/* ---- the handle ----------------------------------------------------------------------------- */
struct Orc
{
    OrcParams P;
    SIM svar;
    LIMITS limits;
    OUTL outlist;
    MESH cells;
    SURFS surf_marks;
    SPHState pn, pnp1;
    vector<IPTState> iptdata;
    Sim_Tree* sph_tree = nullptr;
    Vec_Tree* cell_tree = nullptr;
    Integrator* integ = nullptr;
    ~Orc()
    {
        delete sph_tree;
        delete cell_tree;
        delete integ;
    }
};

static StateVecD vec_from(const double* p)
{
    StateVecD v = StateVecD::Zero();
    if (p)
        for (int d = 0; d < SIMDIM; ++d) v[d] = p[d];
    return v;
}

/* OrcParams (inputs and the derived constants) -> SIM.  Field by field: Var.h INTEG_SETT / FLUID / AERO / SIM. */
static void params_to_sim(const OrcParams& P, SIM& s)
{
    s.integrator.solver_type = uint(P.solver_type);
    s.integrator.max_subits = uint(P.max_subits);
    s.integrator.n_stable = uint(P.n_stable);
    s.integrator.n_stable_limit = uint(P.n_stable_limit);
    s.integrator.n_unstable = uint(P.n_unstable);
    s.integrator.n_unstable_limit = uint(P.n_unstable_limit);
    s.integrator.subits_factor = P.subits_factor;
    s.integrator.cfl = P.cfl;
    s.integrator.cfl_step = P.cfl_step;
    s.integrator.cfl_max = P.cfl_max;
    s.integrator.cfl_min = P.cfl_min;
    s.integrator.current_time = P.current_time;
    s.integrator.last_frame_time = P.last_frame_time;
    s.integrator.frame_time_interval = P.frame_time_interval;
    s.integrator.min_residual = P.min_residual;
    s.integrator.delta_t = P.delta_t;
    s.integrator.delta_t_max = P.delta_t_max;
    s.integrator.delta_t_min = P.delta_t_min;
    s.integrator.nb_beta = P.nb_beta;
    s.integrator.nb_gamma = P.nb_gamma;
    s.integrator.max_shift_vel = P.max_shift_vel;

    FLUID& f = s.fluid;
    f.pressure_rel = uint(P.pressure_rel);
    f.H = P.H;
    f.H_sq = P.H_sq;
    f.sr = P.sr;
    f.H_fac = P.H_fac;
    f.W_dx = P.W_dx;
    f.rho_rest = P.rho_rest;
    f.rho_pipe = P.rho_pipe;
    f.press_pipe = P.press_pipe;
    f.press_back = P.press_back;
    f.rho_max = P.rho_max;
    f.rho_min = P.rho_min;
    f.rho_var = P.rho_var;
    f.rho_max_iter = P.rho_max_iter;
    f.sim_mass = P.sim_mass;
    f.bnd_mass = P.bnd_mass;
    f.W_correc = P.W_correc;
    f.visc_alpha = P.visc_alpha;
    f.speed_sound = P.speed_sound;
    f.mu = P.mu;
    f.nu = P.nu;
    f.sig = P.sig;
    f.gam = P.gam;
    f.B = P.B;
    f.dsph_delta = P.dsph_delta;
    f.dsph_cont = P.dsph_cont;
    f.dsph_mom = 2.0 * (SIMDIM + 2.0); /* IO.cpp:80 */

    AERO& a = s.air;
    a.v_inf = vec_from(P.v_inf);
    a.L = P.aero_L;
    a.td = P.td;
    a.omega = P.omega;
    a.tmax = P.tmax;
    a.ycoef = P.ycoef;
    a.Cf = P.tab_Cf;
    a.Ck = P.tab_Ck;
    a.Cd = P.tab_Cd;
    a.Cb = P.tab_Cb;
    a.Cdef = P.Cdef;
    a.A_sphere = P.A_sphere;
    a.A_plate = P.A_plate;
    a.p_ref = P.p_ref;
    a.rho_g = P.rho_g;
    a.mu_g = P.mu_g;
    a.temp_g = P.temp_g;
    a.R_g = P.R_g;
    a.gamma = P.gamma_g;
    a.sos = P.sos;
    a.i_sos_sq = 1.0 / (P.sos * P.sos); /* IO.cpp:56 */
    a.lam_cutoff = P.lam_cutoff;
    a.interp_fac = P.interp_fac;
    a.i_interp_fac = P.i_interp_fac;
    a.n_full = P.n_full;
    a.i_n_full = P.i_n_full;
    a.acase = P.acase;
    a.use_lam = P.use_lam;
    a.use_TAB_def = P.use_TAB_def;

    s.particle_step = P.particle_step;
    s.dx = P.dx;
    s.grav = vec_from(P.grav);
    s.Asource = P.asource;
    s.ipt.using_ipt = 0;
    s.numThreads = 1;
}

/* SIM -> OrcParams, every field (after the reference's own GetInput + Set_Values) */
static void sim_to_params_full(const SIM& s, OrcParams& P)
{
    std::memset(&P, 0, sizeof(P));
    P.dim = SIMDIM;
#ifdef ALE
    P.ale = 1;
#endif
    P.pressure_rel = int(s.fluid.pressure_rel);
    P.solver_type = int(s.integrator.solver_type);
    P.acase = s.air.acase;
    P.asource = s.Asource;
    P.use_lam = s.air.use_lam;
    P.use_TAB_def = s.air.use_TAB_def;
    P.max_subits = int(s.integrator.max_subits);
    P.n_stable = int(s.integrator.n_stable);
    P.n_stable_limit = int(s.integrator.n_stable_limit);
    P.n_unstable = int(s.integrator.n_unstable);
    P.n_unstable_limit = int(s.integrator.n_unstable_limit);
    P.particle_step = s.particle_step;
    P.H_fac = s.fluid.H_fac;
    P.rho_rest = s.fluid.rho_rest;
    P.press_pipe = s.fluid.press_pipe;
    P.press_back = s.fluid.press_back;
    P.rho_max = s.fluid.rho_max;
    P.rho_min = s.fluid.rho_min;
    P.rho_var = s.fluid.rho_var;
    P.rho_max_iter = s.fluid.rho_max_iter;
    P.visc_alpha = s.fluid.visc_alpha;
    P.speed_sound = s.fluid.speed_sound;
    P.mu = s.fluid.mu;
    P.sig = s.fluid.sig;
    P.gam = s.fluid.gam;
    P.dsph_delta = s.fluid.dsph_delta;
    for (int d = 0; d < SIMDIM; ++d)
    {
        P.grav[d] = s.grav[d];
        P.v_inf[d] = s.air.v_inf[d];
    }
    P.p_ref = s.air.p_ref;
    P.rho_g = s.air.rho_g;
    P.mu_g = s.air.mu_g;
    P.temp_g = s.air.temp_g;
    P.R_g = s.air.R_g;
    P.gamma_g = s.air.gamma;
    P.lam_cutoff = s.air.lam_cutoff;
    P.i_interp_fac = s.air.i_interp_fac;
    P.tab_Cf = s.air.Cf;
    P.tab_Ck = s.air.Ck;
    P.tab_Cd = s.air.Cd;
    P.tab_Cb = s.air.Cb;
    P.cfl = s.integrator.cfl;
    P.cfl_step = s.integrator.cfl_step;
    P.cfl_max = s.integrator.cfl_max;
    P.cfl_min = s.integrator.cfl_min;
    P.subits_factor = s.integrator.subits_factor;
    P.min_residual = s.integrator.min_residual;
    P.delta_t = s.integrator.delta_t;
    P.delta_t_max = s.integrator.delta_t_max;
    P.delta_t_min = s.integrator.delta_t_min;
    P.max_shift_vel = s.integrator.max_shift_vel;
    P.current_time = s.integrator.current_time;
    P.last_frame_time = s.integrator.last_frame_time;
    P.frame_time_interval = s.integrator.frame_time_interval;
    P.B = s.fluid.B;
    P.rho_pipe = s.fluid.rho_pipe;
    P.dx = s.dx;
    P.sim_mass = s.fluid.sim_mass;
    P.bnd_mass = s.fluid.bnd_mass;
    P.H = s.fluid.H;
    P.H_sq = s.fluid.H_sq;
    P.sr = s.fluid.sr;
    P.dsph_cont = s.fluid.dsph_cont;
    P.nu = s.fluid.nu;
    P.W_correc = s.fluid.W_correc;
    P.W_dx = s.fluid.W_dx;
    P.nb_beta = s.integrator.nb_beta;
    P.nb_gamma = s.integrator.nb_gamma;
    P.aero_L = s.air.L;
    P.A_sphere = s.air.A_sphere;
    P.A_plate = s.air.A_plate;
    P.td = s.air.td;
    P.omega = s.air.omega;
    P.tmax = s.air.tmax;
    P.Cdef = s.air.Cdef;
    P.ycoef = s.air.ycoef;
    P.n_full = s.air.n_full;
    P.i_n_full = s.air.i_n_full;
    P.interp_fac = s.air.interp_fac;
    P.sos = s.air.sos;
}

static void sim_to_params(const SIM& s, OrcParams& P)
{
    P.n_stable = int(s.integrator.n_stable);
    P.n_unstable = int(s.integrator.n_unstable);
    P.cfl = s.integrator.cfl;
    P.current_time = s.integrator.current_time;
    P.delta_t = s.integrator.delta_t;
}

extern "C" {

int orc_compiled_dim(void) { return SIMDIM; }

/* Set_Values (IO.cpp:26-128) cannot be compiled (IO.cpp needs TECIO/HDF5 headers); this walks the same assignments
 * through the reference's OWN inline pieces -- FLUID::get_density (Var.h:220-236), Kernel (Kernel.h:37-45),
 * AERO::GetYcoef (Var.h:244-266), get_n_full (Geometry.cpp:282-308) -- so the derived constants of the restatement can
 * be checked against them. */
void orc_ref_set_values(OrcParams* p)
{
    SIM s;
    params_to_sim(*p, s);
    s.fluid.B = s.fluid.rho_rest * pow(s.fluid.speed_sound, 2) / s.fluid.gam;
    s.fluid.rho_pipe = s.fluid.get_density(s.fluid.press_pipe);
    if (s.fluid.rho_max == 1500 && s.fluid.rho_min == 500)
    {
        s.fluid.rho_max = s.fluid.rho_rest * (1.0 + s.fluid.rho_var * 0.01);
        s.fluid.rho_min = s.fluid.rho_rest * (1.0 - s.fluid.rho_var * 0.01);
    }
    s.dx = s.particle_step * pow(s.fluid.rho_pipe / s.fluid.rho_rest, 1.0 / SIMDIM);
    s.fluid.sim_mass = s.fluid.rho_rest * pow(s.particle_step, SIMDIM);
    s.fluid.bnd_mass = s.fluid.sim_mass;
    s.air.sos = sqrt(s.air.temp_g * s.air.R_g * s.air.gamma);
    s.fluid.H = s.fluid.H_fac * s.particle_step;
    s.fluid.H_sq = s.fluid.H * s.fluid.H;
    s.fluid.sr = 4 * s.fluid.H_sq;
    s.fluid.dsph_cont = 2.0 * s.fluid.dsph_delta * s.fluid.H * s.fluid.speed_sound;
    s.fluid.nu = s.fluid.mu / s.fluid.rho_rest;
#if SIMDIM == 2
    s.fluid.W_correc = 7.0 / (4.0 * M_PI * s.fluid.H * s.fluid.H);
#else
    s.fluid.W_correc = (21 / (16 * M_PI * s.fluid.H * s.fluid.H * s.fluid.H));
#endif
    s.fluid.W_dx = Kernel(s.particle_step, s.fluid.H, s.fluid.W_correc);
    s.air.GetYcoef(s.fluid, s.particle_step);
    s.air.n_full = get_n_full(s.particle_step, s.fluid.H);
    s.air.i_n_full = 1.0 / s.air.n_full;
    s.air.interp_fac = 1.0 / s.air.i_interp_fac;
#if SIMDIM == 3
    s.air.A_plate = s.particle_step * s.particle_step;
#else
    s.air.A_plate = s.particle_step;
#endif
    p->B = s.fluid.B;
    p->rho_pipe = s.fluid.rho_pipe;
    p->rho_max = s.fluid.rho_max;
    p->rho_min = s.fluid.rho_min;
    p->dx = s.dx;
    p->sim_mass = s.fluid.sim_mass;
    p->bnd_mass = s.fluid.bnd_mass;
    p->H = s.fluid.H;
    p->H_sq = s.fluid.H_sq;
    p->sr = s.fluid.sr;
    p->dsph_cont = s.fluid.dsph_cont;
    p->nu = s.fluid.nu;
    p->W_correc = s.fluid.W_correc;
    p->W_dx = s.fluid.W_dx;
    p->nb_beta = 0.25;
    p->nb_gamma = 0.5;
    p->aero_L = s.air.L;
    p->A_sphere = s.air.A_sphere;
    p->A_plate = s.air.A_plate;
    p->td = s.air.td;
    p->omega = s.air.omega;
    p->tmax = s.air.tmax;
    p->Cdef = s.air.Cdef;
    p->ycoef = s.air.ycoef;
    p->n_full = s.air.n_full;
    p->i_n_full = s.air.i_n_full;
    p->interp_fac = s.air.interp_fac;
    p->sos = s.air.sos;
}

/* the reference's EOS and kernel, for the unit tests */
double orc_ref_pressure(const OrcParams* p, double rho)
{
    SIM s;
    params_to_sim(*p, s);
    return s.fluid.get_pressure(rho);
}
double orc_ref_density(const OrcParams* p, double press)
{
    SIM s;
    params_to_sim(*p, s);
    return s.fluid.get_density(press);
}
double orc_kernel(double r, double H, double Wc) { return Kernel(r, H, Wc); }
double orc_get_n_full(double dx, double H) { return get_n_full(dx, H); }

Orc* orc_create(const OrcParams* p)
{
    if (p->dim != SIMDIM)
    {
        std::fprintf(stderr, "orc_create(ref): params.dim=%d but library compiled with SIMDIM=%d\n", p->dim, SIMDIM);
        return nullptr;
    }
#ifdef ALE
    const int ale = 1;
#else
    const int ale = 0;
#endif
    if (p->ale != ale)
    {
        std::fprintf(stderr, "orc_create(ref): params.ale=%d but this is the %s binary\n", p->ale, ale ? "-DALE" : "delta-SPH");
        return nullptr;
    }
#ifdef ORC_REF_THREADS
    /* timing flavour (bench.py's reference arm): every thread OMP_NUM_THREADS grants, as FJSPH's main() does
       (FJSPH.cpp:62); results are then as reproducible as the reference's own (reduction order, the npd race) */
    omp_set_num_threads(std::getenv("OMP_NUM_THREADS") ? std::max(atoi(std::getenv("OMP_NUM_THREADS")), 1) : omp_get_num_procs());
#else
    omp_set_num_threads(1);
#endif
    Orc* o = new Orc();
    o->P = *p;
    params_to_sim(o->P, o->svar);
    return o;
}
void orc_destroy(Orc* o) { delete o; }
void orc_get_params(Orc* o, OrcParams* out)
{
    sim_to_params(o->svar, o->P);
    *out = o->P;
}
void orc_set_params(Orc* o, const OrcParams* in)
{
    o->P = *in;
    params_to_sim(o->P, o->svar);
}

int orc_add_block(Orc* o, int is_fluid, int64_t first, int64_t second, int bound_solver, int no_slip, int block_type,
                  int fixed_vel_or_dynamic, int ntimes, const double* times, const double* vels, const double* insert_norm,
                  double insconst, const double* delete_norm, double delconst, const double* aero_norm, double aeroconst,
                  int nback, const int64_t* back, int nbuf, const int64_t* buffer)
{
    size_t const i0 = size_t(first), i1 = size_t(second);
    bound_block B(i0, i1);
    B.bound_solver = bound_solver;
    B.no_slip = no_slip;
    B.block_type = block_type;
    B.fixed_vel_or_dynamic = fixed_vel_or_dynamic;
    B.nTimes = size_t(ntimes);
    for (int t = 0; t < ntimes; ++t) B.times.push_back(times[t]);
    int const nv = std::max(1, ntimes);
    for (int t = 0; t < nv; ++t) B.vels.push_back(vels ? vec_from(vels + 3 * t) : StateVecD::Zero());
    if (insert_norm)
        B.insert_norm = vec_from(insert_norm);
    if (delete_norm)
        B.delete_norm = vec_from(delete_norm);
    if (aero_norm)
        B.aero_norm = vec_from(aero_norm);
    B.insconst = insconst;
    B.delconst = delconst;
    B.aeroconst = aeroconst;
    for (int i = 0; i < nback; ++i)
    {
        B.back.push_back(size_t(back[i]));
        std::vector<size_t> buf;
        for (int j = 0; j < nbuf; ++j) buf.push_back(size_t(buffer[size_t(i) * nbuf + j]));
        B.buffer.push_back(buf);
    }
    if (is_fluid)
        o->svar.n_fluid_blocks++;
    else
    {
        if (o->svar.n_fluid_blocks != 0)
            return -1;
        o->svar.n_bound_blocks++;
    }
    o->limits.push_back(B);
    return int(o->limits.size()) - 1;
}
void orc_clear_blocks(Orc* o)
{
    o->limits.clear();
    o->svar.n_bound_blocks = o->svar.n_fluid_blocks = 0;
}
/* ---- the callers either side of the path: FJSPH's own front end ------------------------------------------------ */

/* GetInput (IO.cpp:305-723, with its Set_Values) + Init_Particles (Init.cpp:270-496) on a para deck, run from the deck's
 * directory as FJSPH is.  The reference exits the process on a bad deck. */
Orc* orc_ref_read_case(const char* para_path)
{
    namespace fs = std::filesystem;
    omp_set_num_threads(1);
    Orc* o = new Orc();
    fs::path const para = fs::absolute(para_path);
    fs::path const cwd = fs::current_path();
    fs::current_path(para.parent_path());
    std::string name = para.filename().string();
    char prog[] = "FJSPH";
    char* argv[2] = {prog, name.data()};
    GetInput(2, argv, o->svar);
    o->pn.reserve(o->svar.max_points); /* FJSPH.cpp:98-99 */
    o->pnp1.reserve(o->svar.max_points);
    Init_Particles(o->svar, o->pn, o->pnp1, o->limits);
    if (o->svar.Asource != meshInfl)
    { /* FJSPH.cpp:113-126 */
        for (size_t ii = 0; ii < o->pnp1.size(); ++ii)
        {
            o->pn[ii].cellRho = o->pnp1[ii].cellRho = o->svar.air.rho_g;
            o->pn[ii].cellP = o->pnp1[ii].cellP = o->svar.air.p_ref;
            o->pn[ii].cellV = o->pnp1[ii].cellV = o->svar.air.v_inf;
        }
    }
    fs::current_path(cwd);
    sim_to_params_full(o->svar, o->P);
    o->svar.numThreads = 1;
    o->sph_tree = new Sim_Tree(SIMDIM, o->pnp1, 20);
    if (o->cells.cCentre.size() == 0)
        o->cells.cCentre.emplace_back(StateVecD::Zero());
    o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10);
    o->cell_tree->index->buildIndex();
    o->integ = new Integrator(int(o->svar.integrator.solver_type));
    return o;
}
/* IPT_SETT as GetInput + Set_Values leave it (IO.cpp:29,126-127,447-453,666-680): using_ipt, ipt_eq_order, streak_out,
 * cells_out, part_out, then max_x, max_x_sph, ipt_diam, ipt_area, relax, n_relax */
void orc_ref_ipt_settings(Orc* o, int32_t* ints /* 5 */, double* reals /* 6 */)
{
    IPT_SETT const& I = o->svar.ipt;
    ints[0] = I.using_ipt;
    ints[1] = I.ipt_eq_order;
    ints[2] = int32_t(I.streak_out);
    ints[3] = int32_t(I.cells_out);
    ints[4] = int32_t(I.part_out);
    reals[0] = I.max_x;
    reals[1] = I.max_x_sph;
    reals[2] = I.ipt_diam;
    reals[3] = I.ipt_area;
    reals[4] = I.relax;
    reals[5] = I.n_relax;
}
int orc_ref_num_blocks(Orc* o) { return int(o->limits.size()); }
int orc_ref_num_bound_blocks(Orc* o) { return int(o->svar.n_bound_blocks); }
/* sizes first (ints: nTimes, n_back, n_buf, bound_solver, no_slip, block_type, fixed_vel_or_dynamic, particle_order),
 * then the arrays into caller buffers of those sizes (any may be NULL) */
int orc_ref_block_info(Orc* o, int block, int64_t* range, int32_t* ints, double* norms /* 3 x 3 */, double* consts /* 3 */,
                       char* name, int name_cap)
{
    if (block < 0 || size_t(block) >= o->limits.size())
        return -1;
    bound_block const& B = o->limits[size_t(block)];
    range[0] = int64_t(B.index.first);
    range[1] = int64_t(B.index.second);
    ints[0] = int32_t(B.nTimes);
    ints[1] = int32_t(B.back.size());
    ints[2] = int32_t(B.buffer.empty() ? 0 : B.buffer[0].size());
    ints[3] = B.bound_solver;
    ints[4] = B.no_slip;
    ints[5] = B.block_type;
    ints[6] = B.fixed_vel_or_dynamic;
    ints[7] = B.particle_order;
    for (int d = 0; d < 3; ++d)
    {
        norms[d] = d < SIMDIM ? B.insert_norm[d] : 0.0;
        norms[3 + d] = d < SIMDIM ? B.delete_norm[d] : 0.0;
        norms[6 + d] = d < SIMDIM ? B.aero_norm[d] : 0.0;
    }
    consts[0] = B.insconst;
    consts[1] = B.delconst;
    consts[2] = B.aeroconst;
    if (name && name_cap > 0)
    {
        std::strncpy(name, B.name.c_str(), size_t(name_cap) - 1);
        name[name_cap - 1] = 0;
    }
    return 0;
}
int orc_ref_block_arrays(Orc* o, int block, double* times, double* vels /* max(1,nTimes) x 3 */, int64_t* back, int64_t* buffer)
{
    if (block < 0 || size_t(block) >= o->limits.size())
        return -1;
    bound_block const& B = o->limits[size_t(block)];
    if (times)
        for (size_t t = 0; t < B.times.size(); ++t) times[t] = B.times[t];
    if (vels)
        for (size_t t = 0; t < B.vels.size(); ++t)
            for (int d = 0; d < 3; ++d) vels[3 * t + d] = d < SIMDIM ? B.vels[t][d] : 0.0;
    if (back)
        for (size_t i = 0; i < B.back.size(); ++i) back[i] = int64_t(B.back[i]);
    if (buffer)
        for (size_t i = 0; i < B.buffer.size(); ++i)
            for (size_t j = 0; j < B.buffer[i].size(); ++j) buffer[i * B.buffer[i].size() + j] = int64_t(B.buffer[i][j]);
    return int(B.vels.size());
}

#if SIMDIM == 3
/* FOAM::Read_FOAM (FOAMIO.cpp:538-955) on an ASCII case directory; the mesh goes into the handle (and can be read back
 * with orc_ref_mesh_sizes / orc_ref_mesh_arrays) */
int orc_ref_read_foam(Orc* o, const char* foam_dir, const char* foam_sol, int buoyant)
{
    o->svar.io.foam_dir = foam_dir;
    o->svar.io.foam_sol = foam_sol;
    o->svar.io.foam_is_binary = false;
    o->svar.io.foam_buoyant_sim = buoyant != 0;
    o->svar.io.mesh_source = OpenFOAM;
    o->svar.Asource = meshInfl;
    delete o->cell_tree;
    o->cell_tree = nullptr;
    o->cells = MESH();
    FOAM::Read_FOAM(o->svar, o->cells);
    o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10);
    o->cell_tree->index->buildIndex();
    return 0;
}
#endif
/* TAU::Read_BMAP (CDFIO.cpp:234-315) on a boundary map: what it does to gravity, given the para's angle of attack (degrees) */
int orc_ref_read_bmap(Orc* o, const char* bmap_file, double alpha_deg, const double* grav_in, double* grav_out)
{
    SIM& s = o->svar; /* not copyable (it owns streams): run in place and put gravity and the angle back */
    const StateVecD grav_saved = s.grav;
    const real alpha_saved = s.io.angle_alpha;
    s.io.tau_bmap = bmap_file;
    s.io.angle_alpha = alpha_deg;
    for (int d = 0; d < SIMDIM; ++d) s.grav[d] = grav_in[d];
    TAU::Read_BMAP(s);
    for (int d = 0; d < SIMDIM; ++d) grav_out[d] = s.grav[d];
    s.grav = grav_saved;
    s.io.angle_alpha = alpha_saved;
    return 0;
}
#if SIMDIM == 3
/* TAU::Read_tau_mesh_FACE + TAU::Read_SOLUTION (CDFIO.cpp:1228-1356,655-822; FJSPH.cpp:76-78) on a face-based NetCDF mesh
 * and a solution file, through the stand-in netcdf.h of shim/; the mesh goes into the handle */
int orc_ref_read_tau(Orc* o, const char* mesh_file, const char* sol_file, double scale)
{
    o->svar.io.tau_mesh = mesh_file;
    o->svar.io.tau_sol = sol_file;
    o->svar.scale = scale;
    o->svar.io.mesh_source = TAU_CDF;
    o->svar.Asource = meshInfl;
    delete o->cell_tree;
    o->cell_tree = nullptr;
    o->cells = MESH();
    o->cells.maxlength = 0.0;
    o->cells.minlength = 1e300;
    TAU::Read_tau_mesh_FACE(o->svar, o->cells);
    vector<uint> empty;
    TAU::Read_SOLUTION(o->svar, o->svar.io.offset_axis, o->cells, empty);
    o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10);
    o->cell_tree->index->buildIndex();
    return 0;
}
#endif
#if SIMDIM == 2
/* TAU::Read_tau_mesh_EDGE + TAU::Read_SOLUTION (CDFIO.cpp:992-1097,655-822; FJSPH.cpp:85-91) on an edge-based NetCDF mesh
 * and a solution file, through the stand-in netcdf.h of shim/; `scale` carries the grid scale, offset_axis the para's
 * "2D offset vector".  The mesh goes into the handle. */
int orc_ref_read_tau_edge(Orc* o, const char* mesh_file, const char* sol_file, double scale, int offset_axis)
{
    o->svar.io.tau_mesh = mesh_file;
    o->svar.io.tau_sol = sol_file;
    o->svar.scale = scale;
    o->svar.io.offset_axis = unsigned(offset_axis);
    o->svar.Asource = meshInfl;
    delete o->cell_tree;
    o->cell_tree = nullptr;
    o->cells = MESH();
    o->cells.maxlength = 0.0;
    o->cells.minlength = 1e300;
    vector<uint> used;
    TAU::Read_tau_mesh_EDGE(o->svar, o->cells, used);
    TAU::Read_SOLUTION(o->svar, o->svar.io.offset_axis, o->cells, used);
    o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10);
    o->cell_tree->index->buildIndex();
    return 0;
}
#endif
/* cells.maxlength as the TAU readers leave it (CDFIO.cpp:867-898, 1117-1183): the particle tracker's bound on one step */
double orc_ref_mesh_max_length(Orc* o) { return o->cells.maxlength; }
void orc_ref_mesh_sizes(Orc* o, int64_t* out /* verts, faces, cells, face_vtx total, cell_faces total */)
{
    MESH const& M = o->cells;
    out[0] = int64_t(M.verts.size());
    out[1] = int64_t(M.faces.size());
    out[2] = int64_t(M.cFaces.size());
    int64_t t = 0;
    for (auto const& f : M.faces) t += int64_t(f.size());
    out[3] = t;
    t = 0;
    for (auto const& c : M.cFaces) t += int64_t(c.size());
    out[4] = t;
}
void orc_ref_mesh_arrays(Orc* o, double* verts, int64_t* face_ptr, int64_t* face_vtx, int32_t* leftright, int64_t* cell_ptr,
                         int64_t* cell_faces, double* cCentre, double* cVel, double* cP, double* cRho)
{
    MESH const& M = o->cells;
    for (size_t i = 0; i < M.verts.size(); ++i)
        for (int d = 0; d < SIMDIM; ++d) verts[i * SIMDIM + d] = M.verts[i][d];
    int64_t k = 0;
    face_ptr[0] = 0;
    for (size_t f = 0; f < M.faces.size(); ++f)
    {
        for (size_t v : M.faces[f]) face_vtx[k++] = int64_t(v);
        face_ptr[f + 1] = k;
        leftright[2 * f] = M.leftright[f].first;
        leftright[2 * f + 1] = M.leftright[f].second;
    }
    k = 0;
    cell_ptr[0] = 0;
    for (size_t c = 0; c < M.cFaces.size(); ++c)
    {
        for (size_t f : M.cFaces[c]) cell_faces[k++] = int64_t(f);
        cell_ptr[c + 1] = k;
        for (int d = 0; d < SIMDIM; ++d)
        {
            cCentre[c * SIMDIM + d] = M.cCentre[c][d];
            cVel[c * SIMDIM + d] = M.cVel[c][d];
        }
        cP[c] = M.cP[c];
        cRho[c] = c < M.cRho.size() ? M.cRho[c] : 0.0;
    }
}

int orc_get_block_range(Orc* o, int block, int64_t* first, int64_t* second)
{
    if (block < 0 || size_t(block) >= o->limits.size())
        return -1;
    *first = int64_t(o->limits[block].index.first);
    *second = int64_t(o->limits[block].index.second);
    return 0;
}

int orc_set_particles(Orc* o, int64_t n, int64_t bound_points, const double* xi, const double* v, const double* rho,
                      const double* p, const double* m, const int32_t* b, const int64_t* part_id)
{
    SIM& svar = o->svar;
    delete o->sph_tree;
    o->sph_tree = nullptr;
    o->pn.clear();
    o->pnp1.clear();
    /* FJSPH.cpp:98-99: the vectors never reallocate under the tree's reference */
    o->pn.reserve(size_t(n) * 4 + 1024);
    o->pnp1.reserve(size_t(n) * 4 + 1024);
    size_t maxid = 0;
    for (size_t i = 0; i < size_t(n); ++i)
    {
        const size_t id = part_id ? size_t(part_id[i]) : i;
        SPHPart part(vec_from(xi + i * SIMDIM), v ? vec_from(v + i * SIMDIM) : StateVecD::Zero(), rho[i], m[i], p[i], b[i],
                     uint(id));
        /* FJSPH.cpp:115-126 (Asource != meshInfl) */
        part.cellRho = svar.air.rho_g;
        part.cellP = svar.air.p_ref;
        part.cellV = svar.air.v_inf;
        o->pnp1.push_back(part);
        maxid = std::max(maxid, id);
    }
    o->pn = o->pnp1;
    svar.bound_points = size_t(bound_points);
    svar.total_points = size_t(n);
    svar.fluid_points = size_t(n - bound_points);
    svar.part_id = maxid + 1;
    svar.delete_count = 0;
    svar.internal_count = 0;
    if (o->limits.empty())
    {
        double z[3] = {0, 0, 0};
        if (bound_points > 0)
            orc_add_block(o, 0, 0, bound_points, pressure_G, 0, 0, 0, 0, nullptr, z, nullptr, default_val, nullptr,
                          default_val, nullptr, default_val, 0, nullptr, 0, nullptr);
        orc_add_block(o, 1, bound_points, n, 0, 0, 0, 0, 0, nullptr, z, nullptr, default_val, nullptr, default_val,
                      nullptr, default_val, 0, nullptr, 0, nullptr);
    }
    o->sph_tree = new Sim_Tree(SIMDIM, o->pnp1, 20); /* FJSPH.cpp:145 */
    if (!o->cell_tree)
    {
        if (o->cells.cCentre.size() == 0)
            o->cells.cCentre.emplace_back(StateVecD::Zero()); /* FJSPH.cpp:137-139 */
        o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10);
        o->cell_tree->index->buildIndex();
    }
    delete o->integ;
    o->integ = new Integrator(int(svar.integrator.solver_type));
    o->outlist.clear();
    return 0;
}
int64_t orc_count(Orc* o) { return int64_t(o->pnp1.size()); }
int64_t orc_bound_points(Orc* o) { return int64_t(o->svar.bound_points); }

static StateVecD* find_vec(SPHPart& p, std::string const& n)
{
    if (n == "xi") return &p.xi;
    if (n == "v") return &p.v;
    if (n == "acc") return &p.acc;
    if (n == "Af") return &p.Af;
    if (n == "aVisc") return &p.aVisc;
    if (n == "cellV") return &p.cellV;
    if (n == "gradRho") return &p.gradRho;
    if (n == "norm") return &p.norm;
    if (n == "bNorm") return &p.bNorm;
    if (n == "vPert") return &p.vPert;
    return nullptr;
}
static real* find_scalar(SPHPart& p, std::string const& n)
{
    if (n == "Rrho") return &p.Rrho;
    if (n == "rho") return &p.rho;
    if (n == "p") return &p.p;
    if (n == "m") return &p.m;
    if (n == "curve") return &p.curve;
    if (n == "norm_curve") return &p.norm_curve;
    if (n == "woccl") return &p.woccl;
    if (n == "pDist") return &p.pDist;
    if (n == "deltaD") return &p.deltaD;
    if (n == "cellP") return &p.cellP;
    if (n == "cellRho") return &p.cellRho;
    if (n == "colourG") return &p.colourG;
    if (n == "colour") return &p.colour;
    if (n == "lam") return &p.lam;
    if (n == "lam_nb") return &p.lam_nb;
    if (n == "kernsum") return &p.kernsum;
    if (n == "y") return &p.y;
    return nullptr;
}

static int access_f64(Orc* o, int level, const char* name, double* out, const double* in)
{
    SPHState& S = level == 0 ? o->pn : o->pnp1;
    const std::string n(name);
    if (S.empty())
        return 0;
    if (n == "L")
    {
        for (size_t i = 0; i < S.size(); ++i)
            for (int r = 0; r < SIMDIM; ++r)
                for (int c = 0; c < SIMDIM; ++c)
                {
                    if (out)
                        out[i * SIMDIM * SIMDIM + r * SIMDIM + c] = S[i].L(r, c);
                    else
                        S[i].L(r, c) = in[i * SIMDIM * SIMDIM + r * SIMDIM + c];
                }
        return SIMDIM * SIMDIM;
    }
    if (find_vec(S[0], n))
    {
        for (size_t i = 0; i < S.size(); ++i)
        {
            StateVecD& v = *find_vec(S[i], n);
            for (int d = 0; d < SIMDIM; ++d)
            {
                if (out)
                    out[i * SIMDIM + d] = v[d];
                else
                    v[d] = in[i * SIMDIM + d];
            }
        }
        return SIMDIM;
    }
    if (find_scalar(S[0], n))
    {
        for (size_t i = 0; i < S.size(); ++i)
        {
            real& s = *find_scalar(S[i], n);
            if (out)
                out[i] = s;
            else
                s = in[i];
        }
        return 1;
    }
    return -1;
}
int orc_get_f64(Orc* o, int level, const char* name, double* out) { return access_f64(o, level, name, out, nullptr); }
int orc_set_f64(Orc* o, int level, const char* name, const double* in) { return access_f64(o, level, name, nullptr, in); }

static int access_i64(Orc* o, int level, const char* name, int64_t* out, const int64_t* in)
{
    SPHState& S = level == 0 ? o->pn : o->pnp1;
    const std::string n(name);
    for (size_t i = 0; i < S.size(); ++i)
    {
        SPHPart& p = S[i];
#define FJ_INT_FIELD(F, T)                 \
    if (n == #F)                           \
    {                                      \
        if (out)                           \
            out[i] = int64_t(p.F);         \
        else                               \
            p.F = T(in[i]);                \
        continue;                          \
    }
        FJ_INT_FIELD(part_id, size_t)
        FJ_INT_FIELD(cellID, long)
        FJ_INT_FIELD(b, uint)
        FJ_INT_FIELD(surf, uint)
        FJ_INT_FIELD(surfzone, uint)
        FJ_INT_FIELD(internal, uint)
        FJ_INT_FIELD(ipt_n_failed, uint)
#undef FJ_INT_FIELD
        return -1;
    }
    return 1;
}
int orc_get_i64(Orc* o, int level, const char* name, int64_t* out) { return access_i64(o, level, name, out, nullptr); }
int orc_set_i64(Orc* o, int level, const char* name, const int64_t* in) { return access_i64(o, level, name, nullptr, in); }

/* ---- stages: the reference's functions, called as integrate_no_update calls them (Integration.cpp:27-107) ---- */
void orc_update_neighbours(Orc* o) { o->outlist = update_neighbours(o->svar.fluid, *o->sph_tree, o->pnp1); }
int64_t orc_neighbour_total(Orc* o)
{
    int64_t t = 0;
    for (auto const& l : o->outlist) t += int64_t(l.size());
    return t;
}
void orc_get_neighbours(Orc* o, int64_t* offsets, int64_t* idx, double* d2)
{
    int64_t k = 0;
    offsets[0] = 0;
    for (size_t i = 0; i < o->outlist.size(); ++i)
    {
        /* ascending j, as the oracle reports them (the shim's scan already returns them so) */
        std::vector<neighbour_index> l = o->outlist[i];
        std::sort(l.begin(), l.end(), [](neighbour_index const& a, neighbour_index const& b) { return a.first < b.first; });
        for (auto const& e : l)
        {
            idx[k] = int64_t(e.first);
            d2[k] = e.second;
            ++k;
        }
        offsets[i + 1] = k;
    }
}
double orc_prestep(Orc* o)
{
    real npd = 1.0;
    dSPH_PreStep(o->svar.fluid, o->svar.total_points, o->pnp1, o->outlist, npd);
    return npd;
}
void orc_aero_velocity(Orc* o)
{
    size_t const start = o->svar.bound_points;
    size_t end = o->svar.total_points;
    real npd = 1.0;
    get_aero_velocity(*o->sph_tree, *o->cell_tree, o->svar, o->cells, start, end, o->outlist, o->limits, o->pn, o->pnp1, npd);
}
void orc_set_mesh(Orc* o, int64_t n_verts, const double* verts, int64_t n_faces, const int64_t* face_ptr,
                  const int64_t* face_vtx, const int32_t* leftright, int64_t n_cells, const int64_t* cell_ptr,
                  const int64_t* cell_faces, const double* cCentre, const double* cVel, const double* cP, const double* cRho)
{
    delete o->cell_tree;
    o->cell_tree = nullptr;
    MESH& M = o->cells;
    M = MESH();
    M.nPnts = size_t(n_verts);
    M.nElem = size_t(n_cells);
    M.nFace = size_t(n_faces);
    M.verts.resize(size_t(n_verts));
    for (int64_t i = 0; i < n_verts; ++i) M.verts[size_t(i)] = vec_from(verts + i * SIMDIM);
    M.faces.resize(size_t(n_faces));
    M.leftright.resize(size_t(n_faces));
    for (int64_t f = 0; f < n_faces; ++f)
    {
        M.faces[size_t(f)].assign(face_vtx + face_ptr[f], face_vtx + face_ptr[f + 1]);
        M.leftright[size_t(f)] = std::make_pair(int(leftright[2 * f]), int(leftright[2 * f + 1]));
    }
    M.cFaces.resize(size_t(n_cells));
    M.cCentre.resize(size_t(n_cells));
    M.cVel.resize(size_t(n_cells));
    M.cP.assign(cP, cP + n_cells);
    M.cRho.assign(cRho, cRho + n_cells);
    for (int64_t c = 0; c < n_cells; ++c)
    {
        M.cFaces[size_t(c)].assign(cell_faces + cell_ptr[c], cell_faces + cell_ptr[c + 1]);
        M.cCentre[size_t(c)] = vec_from(cCentre + c * SIMDIM);
        M.cVel[size_t(c)] = vec_from(cVel + c * SIMDIM);
    }
    o->cell_tree = new Vec_Tree(SIMDIM, o->cells.cCentre, 10); /* FJSPH.cpp:148-149 */
    o->cell_tree->index->buildIndex();
}
int orc_first_cell_errors(Orc*) { return 0; } /* the reference exits instead */

/* IPT::Integrate (IPT.cpp:871-1107) on the particles update_data would hand over (Integration.cpp:151-169) */
static void ipt_point_out(IPTPart const& p, OrcIptPoint& q)
{
    q.part_id = int64_t(p.part_id);
    q.cellID = p.cellID;
    q.faceID = int64_t(int(p.faceID)); /* uint c_no_face -> -1 */
    q.going = int32_t(p.going);
    q.failed = int32_t(p.failed);
    q.t = p.t;
    q.dt = p.dt;
    q.acc = p.acc;
    q.cellRho = p.cellRho;
    for (int d = 0; d < 3; ++d)
    {
        q.xi[d] = d < SIMDIM ? p.xi[d] : 0.0;
        q.v[d] = d < SIMDIM ? p.v[d] : 0.0;
        q.cellV[d] = d < SIMDIM ? p.cellV[d] : 0.0;
    }
}
int orc_ipt_integrate(Orc* o, const OrcIptSettings* S, int64_t n, const OrcIptStart* in, OrcIptPoint* last, int32_t* n_steps,
                      OrcIptPoint* records, int64_t record_cap, int32_t* n_records, int64_t* n_success, int64_t* n_failed)
{
    SIM& svar = o->svar;
    svar.ipt.using_ipt = 1;
    svar.ipt.ipt_eq_order = S->eq_order;
    svar.ipt.relax = S->relax;
    svar.ipt.n_relax = S->n_relax;
    svar.ipt.max_x = S->max_x;
    svar.ipt.ipt_diam = S->diam;
    svar.ipt.ipt_area = S->area;
    svar.ipt.streak_out = S->record ? 1 : 0;
    svar.ipt.cells_out = 0;
    svar.ipt.part_out = 0;
    svar.ipt.ipt_n_success = 0;
    svar.ipt.ipt_n_failed = 0;
    svar.integrator.max_subits = uint(S->max_subits);
    svar.grav = vec_from(S->grav);
    svar.air.mu_g = S->mu_g;
    svar.fluid.rho_rest = S->rho_rest;
    if (S->max_length >= 0.0) /* negative: keep what the reference's own mesh reader left */
        o->cells.maxlength = S->max_length;
    for (int64_t i = 0; i < n; ++i)
    {
        SPHPart sp;
        sp.part_id = size_t(in[i].part_id);
        sp.cellID = long(in[i].cellID);
        sp.cellV = vec_from(in[i].cellV);
        sp.cellRho = in[i].cellRho;
        sp.v = vec_from(in[i].v);
        sp.xi = vec_from(in[i].xi);
        sp.m = in[i].mass;
        IPTPart nm1(sp, in[i].t, svar.ipt.ipt_diam, svar.ipt.ipt_area); /* Integration.cpp:156-165 */
        IPTPart pn = nm1, np1 = nm1;
        o->iptdata.clear();
        IPT::Integrate(svar, o->cells, size_t(i), nm1, pn, np1, o->surf_marks, o->iptdata);
        if (last)
            ipt_point_out(np1, last[i]);
        if (n_steps)
            n_steps[i] = 0;
        IPTState const empty;
        IPTState const& rec = o->iptdata.empty() ? empty : o->iptdata.back();
        if (n_records)
            n_records[i] = int32_t(rec.size());
        if (records)
            for (size_t k = 0; k < rec.size() && int64_t(k) < record_cap; ++k) ipt_point_out(rec[k], records[i * record_cap + int64_t(k)]);
    }
    o->iptdata.clear();
    if (n_success)
        *n_success = int64_t(svar.ipt.ipt_n_success);
    if (n_failed)
        *n_failed = int64_t(svar.ipt.ipt_n_failed);
    return 0;
}
void orc_detect_surface(Orc* o)
{
    Detect_Surface(o->svar, o->svar.bound_points, o->svar.total_points, o->outlist, o->cells, o->pnp1);
}
void orc_dissipation(Orc* o)
{
    dissipation_terms(o->svar.fluid, o->svar.bound_points, o->svar.total_points, o->outlist, o->pnp1);
}
void orc_particle_shift(Orc* o)
{
#ifdef ALE
    particle_shift(o->svar, o->svar.bound_points, o->svar.total_points, o->outlist, o->pnp1);
#else
    (void)o;
#endif
}
void orc_forces(Orc* o, double npd) { get_acc_and_Rrho(o->svar, o->cells, o->outlist, npd, o->pnp1); }
void orc_nb_iter(Orc* o, double npd)
{
    size_t const start = o->svar.bound_points;
    size_t end = o->svar.total_points;
    Newmark_Beta::Do_NB_Iter(*o->cell_tree, o->svar, start, end, npd, o->cells, o->limits, o->outlist, o->pn, o->pnp1);
}
double orc_find_timestep(Orc* o)
{
    return o->integ->find_timestep(o->svar, o->cells, o->pnp1, o->svar.bound_points, o->svar.total_points);
}
static void fill_stats(Orc* o, OrcStepStats* s, double e)
{
    if (!s)
        return;
    s->dt = o->svar.integrator.delta_t;
    s->rms_error = e;
    s->safe_dt = o->integ->safe_dt;
    s->cfl_ratio = o->svar.integrator.delta_t / o->integ->safe_dt;
    s->maxRho_pc = o->integ->maxRho_pc;
    s->maxf = o->integ->maxf;
    s->maxAf = o->integ->maxAf;
#ifdef ALE
    s->maxShift = o->integ->maxShift;
#endif
    s->iterations = int(o->integ->iteration);
    s->total_points = int(o->pnp1.size());
}
double orc_integrate_no_update(Orc* o, OrcStepStats* s)
{
    if (s)
        std::memset(s, 0, sizeof(*s));
    o->integ->solver_method = int(o->svar.integrator.solver_type);
    real const e = o->integ->integrate_no_update(*o->sph_tree, *o->cell_tree, o->svar, o->cells, o->limits, o->outlist,
                                                 o->pn, o->pnp1);
    fill_stats(o, s, e);
    return e;
}
double orc_integrate(Orc* o, OrcStepStats* s)
{
    if (s)
        std::memset(s, 0, sizeof(*s));
    o->integ->solver_method = int(o->svar.integrator.solver_type);
    size_t const total0 = o->svar.total_points, del0 = o->svar.delete_count;
    real const e = o->integ->integrate(*o->sph_tree, *o->cell_tree, o->svar, o->cells, o->surf_marks, o->limits, o->outlist,
                                       o->pn, o->pnp1, o->iptdata);
    fill_stats(o, s, e);
    if (s)
    {
        s->n_del = int(o->svar.delete_count - del0);
        s->n_add = int(o->svar.total_points + size_t(s->n_del) - total0);
    }
    return e;
}

} /* extern "C" */
