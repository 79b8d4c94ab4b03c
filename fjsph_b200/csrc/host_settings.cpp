// host_settings.cpp — host-side settings path in front of the device engine.
//
// Mirrors, for the keys the time-step path consumes:
//   struct defaults                         reference src/Var.h:63-92,160-200,298-309,349-354
//   GetInput's `key : value` para parser    reference src/IO.cpp:305-456, src/IOFunctions.h:32-131
//   post-parse name -> enum mapping         reference src/IO.cpp:586-652
//   Set_Values (derived constants)          reference src/IO.cpp:26-128
//   AERO::GetYcoef                          reference src/Var.h:244-266
//   get_n_full                              reference src/Geometry.cpp:282-308
// Table-driven instead of the reference's one Get_Number call per key per line.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/fjsph_b200.h"

void fj_set_error(const char* fmt, ...);

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace
{
enum Kind
{
    K_F64,
    K_I32,
    K_VEC3
};
struct KeySpec
{
    const char* key;
    Kind kind;
    size_t offset;
};
#define KF(k, f) {k, K_F64, offsetof(FjsphParams, f)}
#define KI(k, f) {k, K_I32, offsetof(FjsphParams, f)}
#define KV(k, f) {k, K_VEC3, offsetof(FjsphParams, f)}
const KeySpec kKeys[] = {
    KF("SPH frame time interval", frame_time_interval),
    KF("Reference dispersed density", rho_rest),
    KF("Sutherland reference viscosity", mu_g),
    KF("Reference dispersed viscosity", mu),
    KF("Reference surface tension", sig),
    KI("SPH equation of state (0=Cole/1=Isothermal)", pressure_rel),
    KF("SPH solver minimum residual", min_residual),
    KF("SPH maximum timestep", delta_t_max),
    KF("SPH minimum timestep", delta_t_min),
    KF("SPH maximum CFL", cfl_max),
    KF("SPH minimum CFL", cfl_min),
    KF("SPH CFL condition", cfl),
    KF("SPH unstable CFL step", cfl_step),
    KI("SPH unstable CFL count limit", n_unstable_limit),
    KI("SPH stable CFL count limit", n_stable_limit),
    KF("SPH stable CFL count iteration factor", subits_factor),
    KF("SPH maximum shifting velocity", max_shift_vel),
    KF("SPH background pressure", press_back),
    KF("SPH starting pressure", press_pipe),
    KF("SPH maximum absolute density variation (%)", rho_var),
    KF("SPH density variation to reduce timestep (%)", rho_max_iter),
    KF("SPH maximum density", rho_max),
    KF("SPH minimum density", rho_min),
    KF("SPH delta coefficient", dsph_delta),
    KF("SPH artificial viscosity factor", visc_alpha),
    KF("SPH speed of sound", speed_sound),
    KI("SPH Newmark-Beta iteration limit", max_subits),
    KV("SPH gravity vector", grav),
    KF("SPH initial spacing", particle_step),
    KF("SPH smoothing length factor", H_fac),
    KI("SPH use TAB deformation (0/1)", use_TAB_def),
    KI("SPH interpolation factor (0=ncount/1=lambda)", use_lam),
    KF("SPH aerodynamic cutoff value", lam_cutoff),
    KF("SPH aerodynamic interpolation factor", i_interp_fac),
    KV("SPH freestream velocity", v_inf),
    KF("Reference pressure", p_ref),
    KF("Reference density", rho_g),
    KF("Reference temperature", temp_g),
    KF("Gas constant gamma", gamma_g),
};

std::string trim(const std::string& s)
{
    const char* ws = " \n\r\t\f\v";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos)
        return "";
    const size_t b = s.find_last_not_of(ws);
    return s.substr(a, b - a + 1);
}

double wendland(double r, double H, double Wc)
{
    const double t = 1.0 - 0.5 * r / H;
    const double t2 = t * t;
    return (t2 * t2) * (2.0 * r / H + 1.0) * Wc;
}

// number of lattice points strictly inside the support sphere of radius 2H (Geometry.cpp:282-308);
// the lattice is accumulated with x += dx exactly as the reference does, so the <= 6 axis points at
// distance 2H fall in or out by rounding the same way.
double lattice_support_count(double dx, double H, int dim)
{
    const double sr = 4.0 * H * H;
    const double lim = 2.0 * (H + dx);
    long count = 0;
    for (double x = -lim; x <= lim; x += dx)
        for (double y = -lim; y <= lim; y += dx)
        {
            if (dim == 3)
            {
                for (double z = -lim; z <= lim; z += dx)
                {
                    double d2 = (0.0 - x) * (0.0 - x);
                    d2 += (0.0 - y) * (0.0 - y);
                    d2 += (0.0 - z) * (0.0 - z);
                    count += d2 < sr;
                }
            }
            else
            {
                double d2 = (0.0 - x) * (0.0 - x);
                d2 += (0.0 - y) * (0.0 - y);
                count += d2 < sr;
            }
        }
    return double(count);
}
} // namespace

extern "C" int fjsph_default_params(FjsphParams* p, int dim)
{
    if (!p || (dim != 2 && dim != 3))
    {
        fj_set_error("default_params: dim must be 2 or 3");
        return FJSPH_ERR_INVALID;
    }
    std::memset(p, 0, sizeof(*p));
    p->dim = dim;
    p->ale = 1;
    p->use_lam = 1;
    p->max_subits = 20;
    p->n_stable_limit = 10;
    p->n_unstable_limit = 3;
    p->particle_step = -1.0;
    p->H_fac = 2.0;
    p->rho_rest = 1000.0;
    p->rho_max = 1500.0;
    p->rho_min = 500.0;
    p->rho_var = 50.0;
    p->rho_max_iter = 1.0;
    p->visc_alpha = 0.1;
    p->speed_sound = 300.0;
    p->mu = 8.94e-4;
    p->sig = 0.0708;
    p->gam = 7.0;
    p->dsph_delta = 0.1;
    p->grav[dim - 1] = -9.81;
    p->p_ref = 101353.0;
    p->rho_g = 1.29251;
    p->mu_g = 1.716e-5;
    p->temp_g = 298.0;
    p->R_g = 287.0;
    p->gamma_g = 1.403;
    p->lam_cutoff = 0.75;
    p->i_interp_fac = 0.5;
    p->tab_Cf = 1.0 / 3.0;
    p->tab_Ck = 8.0;
    p->tab_Cd = 5.0;
    p->tab_Cb = 0.5;
    p->cfl = 1.0;
    p->cfl_step = 0.05;
    p->cfl_max = 2.0;
    p->cfl_min = 0.1;
    p->subits_factor = 0.333;
    p->min_residual = -7.0;
    p->delta_t = 2e-10;
    p->delta_t_max = 1.0;
    p->max_shift_vel = 9999999;
    p->frame_time_interval = -1.0;
    return FJSPH_OK;
}

extern "C" int fjsph_set_values(FjsphParams* p)
{
    if (!p)
        return FJSPH_ERR_INVALID;
    FjsphParams& P = *p;
    const int D = P.dim;
    if (D != 2 && D != 3)
    {
        fj_set_error("set_values: dim must be 2 or 3");
        return FJSPH_ERR_INVALID;
    }
    if (!(P.particle_step > 0.0))
    {
        fj_set_error("ERROR: SPH initial spacing has not been defined."); /* IO.cpp:604-608 */
        return FJSPH_ERR_INVALID;
    }
    if (P.i_interp_fac < 0 || P.i_interp_fac > 1.0)
    {
        fj_set_error("aerodynamic interpolation factor must lie in [0,1]");
        return FJSPH_ERR_INVALID;
    }
    P.B = P.rho_rest * std::pow(P.speed_sound, 2) / P.gam;
    if (P.pressure_rel == 0)
        P.rho_pipe = P.rho_rest * std::pow(((P.press_pipe - P.press_back) / P.B) + 1.0, 1.0 / P.gam);
    else
        P.rho_pipe = (P.press_pipe - P.press_back) / (P.speed_sound * P.speed_sound) + P.rho_rest;
    if (P.rho_max == 1500 && P.rho_min == 500)
    {
        P.rho_max = P.rho_rest * (1.0 + P.rho_var * 0.01);
        P.rho_min = P.rho_rest * (1.0 - P.rho_var * 0.01);
    }
    P.dx = P.particle_step * std::pow(P.rho_pipe / P.rho_rest, 1.0 / D);
    P.nb_beta = 0.25;
    P.nb_gamma = 0.5;
    P.sim_mass = P.rho_rest * std::pow(P.particle_step, D);
    P.bnd_mass = P.sim_mass;
    P.sos = std::sqrt(P.temp_g * P.R_g * P.gamma_g);
    P.delta_t = (P.delta_t_min > 0) ? P.delta_t_min : 2E-010;
    P.H = P.H_fac * P.particle_step;
    P.H_sq = P.H * P.H;
    P.sr = 4 * P.H_sq;
    P.dsph_cont = 2.0 * P.dsph_delta * P.H * P.speed_sound;
    P.nu = P.mu / P.rho_rest;
    P.W_correc = (D == 2) ? 7.0 / (4.0 * M_PI * P.H * P.H) : (21 / (16 * M_PI * P.H * P.H * P.H));
    P.W_dx = wendland(P.particle_step, P.H, P.W_correc);

    const double diam = P.particle_step;
    if (D == 3)
    {
        P.aero_L = diam * std::cbrt(3.0 / (4.0 * M_PI));
        P.A_sphere = M_PI * P.aero_L * P.aero_L;
    }
    else
    {
        P.aero_L = diam / std::sqrt(M_PI);
        P.A_sphere = 2 * P.aero_L;
    }
    P.td = (2.0 * P.rho_rest * std::pow(P.aero_L, D - 1)) / (P.tab_Cd * P.mu);
    P.omega = std::sqrt((P.tab_Ck * P.sig) / (P.rho_rest * std::pow(P.aero_L, D)) - 1.0 / std::pow(P.td, 2.0));
    P.tmax = -2.0 * (std::atan(std::sqrt(std::pow(P.td * P.omega, 2.0) + 1) + P.td * P.omega) - M_PI) / P.omega;
    P.Cdef = 1.0 - std::exp(-P.tmax / P.td) *
                       (std::cos(P.omega * P.tmax) + 1 / (P.omega * P.td) * std::sin(P.omega * P.tmax));
    P.ycoef = 0.5 * P.Cdef * (P.tab_Cf / (P.tab_Ck * P.tab_Cb)) * (P.rho_g * P.aero_L) / P.sig;
    P.n_full = lattice_support_count(P.particle_step, P.H, D);
    P.i_n_full = 1.0 / P.n_full;
    P.interp_fac = 1.0 / P.i_interp_fac;
    P.A_plate = (D == 3) ? P.particle_step * P.particle_step : P.particle_step;
    return FJSPH_OK;
}

// Reads a FJSPH para file.  Text after '#' is cut (IO.cpp:344-346); the key is the text left of the
// first ':' (left-trimmed), the value the trimmed text right of it.  Unknown keys are ignored like the
// reference does.  Returns the block file names through fluid_file / bound_file when non-NULL.
extern "C" int fjsph_read_para(const char* path, FjsphParams* p, char* fluid_file, char* bound_file, int name_cap)
{
    if (!path || !p)
        return FJSPH_ERR_INVALID;
    std::ifstream fin(path);
    if (!fin.is_open())
    {
        fj_set_error("could not open SPH parameter file \"%s\"", path);
        return FJSPH_ERR_IO;
    }
    std::string solver_name, aero_case, line;
    while (std::getline(fin, line))
    {
        const size_t hash = line.find('#');
        if (hash != std::string::npos)
            line = line.substr(0, hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos)
            continue;
        const std::string key = trim(line.substr(0, colon));
        const std::string val = trim(line.substr(colon + 1));
        if (key == "SPH integration solver")
            solver_name = val;
        else if (key == "SPH aerodynamic case")
        {
            aero_case = val;
        }
        else if (key == "Input fluid definition filename" && fluid_file)
            std::snprintf(fluid_file, size_t(name_cap), "%s", val.c_str());
        else if (key == "Input boundary definition filename" && bound_file)
            std::snprintf(bound_file, size_t(name_cap), "%s", val.c_str());
        for (const KeySpec& k : kKeys)
        {
            if (key != k.key)
                continue;
            char* dst = (char*)p + k.offset;
            std::istringstream iss(val);
            if (k.kind == K_F64)
            {
                double v;
                if (iss >> v)
                    *(double*)dst = v;
            }
            else if (k.kind == K_I32)
            {
                int v;
                if (iss >> v)
                    *(int32_t*)dst = v;
            }
            else
            {
                /* comma separated components, exactly as IOFunctions.h:133-218 reads them: SIMDIM tokens; a token that
                   is missing leaves the previous token's text in place (getline on an exhausted stream does not touch
                   its string), so "0,0" in a 3D deck reads as (0, 0, 0); text that is not a number reads as 0 */
                std::string item;
                for (int d = 0; d < (p->dim == 2 ? 2 : 3); ++d)
                {
                    std::getline(iss, item, ',');
                    std::istringstream is2(item);
                    double v = 0.0;
                    if (!(is2 >> v))
                        v = 0.0;
                    ((double*)dst)[d] = v;
                }
            }
        }
    }
    if (!solver_name.empty())
    {
        if (solver_name == "Newmark-Beta")
            p->solver_type = 0;
        else if (solver_name == "Runge-Kutta")
            p->solver_type = 1;
        else
        {
            fj_set_error("ERROR: Unrecognised solver name \"%s\". Choose Newmark-Beta or Runge-Kutta.", solver_name.c_str());
            return FJSPH_ERR_INVALID;
        }
    }
    /* AERO::aero_case starts empty and GetInput rejects anything but the four names (IO.cpp:606-627): a deck must
       name its aerodynamic case, "(none)" included */
    {
        if (aero_case == "(none)")
            p->acase = 0;
        else if (aero_case == "Gissler")
            p->acase = 1;
        else if (aero_case == "Induced_pressure")
            p->acase = 2;
        else if (aero_case == "Skin_friction")
            p->acase = 3;
        else
        {
            fj_set_error("Aerodynamic coupling model is not defined or correct.");
            return FJSPH_ERR_INVALID;
        }
    }
    return FJSPH_OK;
}
