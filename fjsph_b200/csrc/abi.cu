// abi.cu — the extern "C" boundary declared in include/fjsph_b200.h: lifetime, host<->device state
// transfer (AoS-of-vec3 host arrays <-> 32-byte device records), blocks, stage entry points, timers.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "engine.cuh"

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";

void fj_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fj_cuda_fail(cudaError_t err, const char* what, const char* file, int line)
{
    fj_set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(err), cudaGetErrorString(err), file, line, what);
    return FJSPH_ERR_CUDA;
}

// ------------------------------------------------------------------ timers
KScope::KScope(FjsphEngine* e_, const char* name, int launches) : e(e_), id(-1)
{
    e->launches += launches;
    if (e->slab.pending && !e->slab.hold)
        fj_halo_wait(e); /* every kernel family but the interior launches of the split sweeps sees complete ghosts */
    if (!e->timers_on)
        return;
    for (size_t k = 0; k < e->timers.size(); ++k)
        if (e->timers[k].name == name)
            id = int(k);
    if (id < 0)
    {
        Timer t;
        t.name = name;
        e->timers.push_back(t);
        id = int(e->timers.size()) - 1;
    }
    e->timers[id].launches += launches;
    e->timers[id].calls += 1;
    PendingTiming pt;
    pt.id = id;
    for (cudaEvent_t* ev : {&pt.a, &pt.b})
    {
        if (e->event_pool.empty())
            cudaEventCreate(ev);
        else
        {
            *ev = e->event_pool.back();
            e->event_pool.pop_back();
        }
    }
    cudaEventRecord(pt.a, e->stream);
    e->pending.push_back(pt);
}
KScope::~KScope()
{
    if (id < 0)
        return;
    cudaEventRecord(e->pending.back().b, e->stream);
    if (e->pending.size() >= 4096)
        fj_timers_flush(e);
}
// resolve the recorded event pairs into per-family milliseconds (one synchronisation for all of them)
void fj_timers_flush(FjsphEngine* e)
{
    if (e->pending.empty())
        return;
    cudaEventSynchronize(e->pending.back().b);
    for (const PendingTiming& pt : e->pending)
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pt.a, pt.b) == cudaSuccess && pt.id < int(e->timers.size()))
            e->timers[pt.id].ms += ms;
        e->event_pool.push_back(pt.a);
        e->event_pool.push_back(pt.b);
    }
    e->pending.clear();
}

// ------------------------------------------------------------------ constants
void fj_refresh_constants(FjsphEngine* e)
{
    const FjsphParams& P = e->P;
    DevConst& C = e->C;
    C.H = P.H;
    C.H_sq = P.H_sq;
    C.iH = 1.0 / P.H;
    C.sr = P.sr;
    C.W_correc = P.W_correc;
    C.W_dx = P.W_dx;
    C.iW_dx = 1.0 / P.W_dx;
    C.gk_fac = 5.0 * P.W_correc / (P.H * P.H);
    C.mhalf_iH = -0.5 * C.iH;
    {
        const double tiny = 1e-12 * P.H;
        C.tiny2 = tiny * tiny;
    }
    C.eps_f = 0.001 * C.H_sq;
    C.eps_d = 0.0001 * C.H_sq;
    C.q_st = 0.75 * C.iH;
    C.rho_rest = P.rho_rest;
    C.rho_min = P.rho_min;
    C.rho_max = P.rho_max;
    C.B = P.B;
    C.gam = P.gam;
    C.c2 = P.speed_sound * P.speed_sound;
    C.press_back = P.press_back;
    C.Bgam = P.B * P.gam;
    C.visc_alpha = P.visc_alpha;
    C.nu = P.nu;
    C.dsph_cont = P.dsph_cont;
    C.sig = P.sig;
    C.dx = P.dx;
    C.particle_step = P.particle_step;
    C.gx = P.grav[0];
    C.gy = P.grav[1];
    C.gz = P.grav[2];
    C.vinf_x = P.v_inf[0];
    C.vinf_y = P.v_inf[1];
    C.vinf_z = P.v_inf[2];
    C.lam_cutoff = P.lam_cutoff;
    C.interp_fac = P.interp_fac;
    C.i_n_full = P.i_n_full;
    C.aero_L = P.aero_L;
    C.A_sphere = P.A_sphere;
    C.A_plate = P.A_plate;
    C.mu_g = P.mu_g;
    C.sos2 = P.sos * P.sos;
    C.gamma_g = P.gamma_g;
    C.ycoef = P.ycoef;
    C.tab_Cb = P.tab_Cb;
    C.max_shift_vel = P.max_shift_vel;
    C.bnd_mass = P.bnd_mass;
    C.sim_mass = P.sim_mass;
    C.c_sound = P.speed_sound;
    C.ale = P.ale;
    C.pressure_rel = P.pressure_rel;
    C.acase = P.acase;
    C.rho_g = P.rho_g;
    C.asource = P.asource;
    C.use_lam = P.use_lam;
    C.use_TAB_def = P.use_TAB_def;
    C.dim = P.dim;
}

static int validate_params(const FjsphParams& P)
{
    if (P.dim != 2 && P.dim != 3)
    {
        fj_set_error("dim must be 2 or 3 (got %d)", P.dim);
        return FJSPH_ERR_INVALID;
    }
    if (P.dim == 2 && (P.grav[2] != 0.0 || P.v_inf[2] != 0.0))
    {
        fj_set_error("SIMDIM=2: the third component of gravity and of the free stream must be 0");
        return FJSPH_ERR_INVALID;
    }
    if (!(P.H > 0.0) || !(P.sr > 0.0) || !(P.W_correc > 0.0))
    {
        fj_set_error("derived constants missing: call fjsph_set_values() before fjsph_create()");
        return FJSPH_ERR_INVALID;
    }
    if (P.acase < 0 || P.acase > 3)
    {
        fj_set_error("aerodynamic case %d unknown (0 none, 1 Gissler, 2 Induced_pressure, 3 Skin_friction)", P.acase);
        return FJSPH_ERR_INVALID;
    }
    if (P.asource != 0 && P.asource != 1)
    {
        fj_set_error("aero source %d is not supported (0 constVel, 1 meshInfl; VLM is out of scope)", P.asource);
        return FJSPH_ERR_INVALID;
    }
    if (P.solver_type != 0 && P.solver_type != 1)
    {
        fj_set_error("solver_type %d unknown (0 Newmark-Beta, 1 Runge-Kutta)", P.solver_type);
        return FJSPH_ERR_INVALID;
    }
    if (P.pressure_rel != 0 && P.pressure_rel != 1)
    {
        fj_set_error("pressure_rel %d unknown (0 Cole, 1 isothermal)", P.pressure_rel);
        return FJSPH_ERR_INVALID;
    }
    return FJSPH_OK;
}

// ------------------------------------------------------------------ pack / unpack kernels
struct StageView
{
    long long* part_id;
    long long* cellID;
    int *b, *surf, *surfzone, *internal;
    double *xi, *v, *acc, *Af, *aVisc, *cellV, *gradRho, *norm, *bNorm, *vPert, *L;
    double *Rrho, *rho, *p, *m, *curve, *norm_curve, *woccl, *pDist, *deltaD, *cellP, *cellRho, *colourG, *colour,
        *lam, *lam_nb, *kernsum, *y;
};

namespace
{
constexpr int TPB = 256;

__global__ void k_init_level(Level S, const int* __restrict__ oidx, DevConst C, double cellRho, double cellP, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double4 z = make_double4(0, 0, 0, 0);
    S.P0[i] = make_double4(0, 0, 0, 1.0);
    S.P1[i] = make_double4(0, 0, 0, 1.0);
    S.P2[i] = z;
    S.P3[i] = z;
    S.P4[i] = z;
    S.ACC[i] = z;
    S.AF[i] = z;
    S.AV[i] = z;
    S.CV[i] = make_double4(C.vinf_x, C.vinf_y, C.vinf_z, cellP); /* FJSPH.cpp:115-126 */
    S.NP[i] = z;
    S.BN[i] = z;
    S.TH[i] = make_double4(0.0, 1.0, 0.0, cellRho);
    S.SC[i] = z;
    S.L0[i] = S.L1[i] = S.L2[i] = S.L3[i] = S.L4[i] = S.L5[i] = S.L6[i] = S.L7[i] = S.L8[i] = 0.0;
    S.part_id[i] = oidx[i]; /* the caller's index of the particle in slot i */
    S.cellID[i] = -3; /* c_no_cell, VarDefs.h:197 */
    S.b[i] = 0;
    S.surfzone[i] = 0;
    S.internal[i] = 0;
    S.surf_i[i] = 0;
}

__global__ void k_identity_index(int* oidx, int* slot_of, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        oidx[i] = i;
        slot_of[i] = i;
    }
}

__global__ void k_assign_blk(const int* __restrict__ oidx, const long long* __restrict__ ranges, int n_blocks,
                             int* __restrict__ blk, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const long long c = oidx[i];
    int out = n_blocks - 1;
    for (int k = 0; k < n_blocks; ++k)
        if (c >= ranges[2 * k] && c < ranges[2 * k + 1])
        {
            out = k;
            break;
        }
    blk[i] = out;
}

// host-order staged arrays -> device records (only non-null fields are overwritten)
__global__ void k_pack(Level S, StageView h, const int* __restrict__ slot_of, int n)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n)
        return;
    const int i = slot_of[c];
    double4 p0 = S.P0[i], p1 = S.P1[i], p2 = S.P2[i], th = S.TH[i];
    double rho = p1.w, p = th.x, m = th.y;
    if (h.xi)
    {
        p0.x = h.xi[3 * c];
        p0.y = h.xi[3 * c + 1];
        p0.z = h.xi[3 * c + 2];
    }
    if (h.v)
    {
        p1.x = h.v[3 * c];
        p1.y = h.v[3 * c + 1];
        p1.z = h.v[3 * c + 2];
    }
    if (h.vPert)
    {
        p2.x = h.vPert[3 * c];
        p2.y = h.vPert[3 * c + 1];
        p2.z = h.vPert[3 * c + 2];
    }
    if (h.rho)
        rho = h.rho[c];
    if (h.p)
        p = h.p[c];
    if (h.m)
        m = h.m[c];
    p0.w = m / rho;
    p1.w = rho;
    p2.w = p / (rho * rho);
    th.x = p;
    th.y = m;
    if (h.woccl)
        th.z = h.woccl[c];
    if (h.cellRho)
        th.w = h.cellRho[c];
    S.P0[i] = p0;
    S.P1[i] = p1;
    S.P2[i] = p2;
    S.TH[i] = th;
    if (h.gradRho || h.lam)
    {
        double4 r = S.P3[i];
        if (h.gradRho)
        {
            r.x = h.gradRho[3 * c];
            r.y = h.gradRho[3 * c + 1];
            r.z = h.gradRho[3 * c + 2];
        }
        if (h.lam)
            r.w = h.lam[c];
        S.P3[i] = r;
    }
    if (h.norm || h.surf || h.lam_nb)
    {
        double4 r = S.P4[i], q = S.NP[i];
        if (h.norm)
        {
            r.x = q.x = h.norm[3 * c];
            r.y = q.y = h.norm[3 * c + 1];
            r.z = q.z = h.norm[3 * c + 2];
        }
        if (h.surf)
            r.w = double(h.surf[c]);
        if (h.lam_nb)
            q.w = h.lam_nb[c];
        S.P4[i] = r;
        S.surf_i[i] = (r.w != 0.0) ? 1 : 0;
        S.NP[i] = q;
    }
    if (h.acc || h.Rrho)
    {
        double4 r = S.ACC[i];
        if (h.acc)
        {
            r.x = h.acc[3 * c];
            r.y = h.acc[3 * c + 1];
            r.z = h.acc[3 * c + 2];
        }
        if (h.Rrho)
            r.w = h.Rrho[c];
        S.ACC[i] = r;
    }
    if (h.Af || h.deltaD)
    {
        double4 r = S.AF[i];
        if (h.Af)
        {
            r.x = h.Af[3 * c];
            r.y = h.Af[3 * c + 1];
            r.z = h.Af[3 * c + 2];
        }
        if (h.deltaD)
            r.w = h.deltaD[c];
        S.AF[i] = r;
    }
    if (h.aVisc || h.curve)
    {
        double4 r = S.AV[i];
        if (h.aVisc)
        {
            r.x = h.aVisc[3 * c];
            r.y = h.aVisc[3 * c + 1];
            r.z = h.aVisc[3 * c + 2];
        }
        if (h.curve)
            r.w = h.curve[c];
        S.AV[i] = r;
    }
    if (h.cellV || h.cellP)
    {
        double4 r = S.CV[i];
        if (h.cellV)
        {
            r.x = h.cellV[3 * c];
            r.y = h.cellV[3 * c + 1];
            r.z = h.cellV[3 * c + 2];
        }
        if (h.cellP)
            r.w = h.cellP[c];
        S.CV[i] = r;
    }
    if (h.bNorm || h.y)
    {
        double4 r = S.BN[i];
        if (h.bNorm)
        {
            r.x = h.bNorm[3 * c];
            r.y = h.bNorm[3 * c + 1];
            r.z = h.bNorm[3 * c + 2];
        }
        if (h.y)
            r.w = h.y[c];
        S.BN[i] = r;
    }
    if (h.colourG || h.colour || h.kernsum || h.pDist)
    {
        double4 r = S.SC[i];
        if (h.colourG)
            r.x = h.colourG[c];
        if (h.colour)
            r.y = h.colour[c];
        if (h.kernsum)
            r.z = h.kernsum[c];
        if (h.pDist)
            r.w = h.pDist[c];
        S.SC[i] = r;
    }
    if (h.L)
    {
        S.L0[i] = h.L[9 * c];
        S.L1[i] = h.L[9 * c + 1];
        S.L2[i] = h.L[9 * c + 2];
        S.L3[i] = h.L[9 * c + 3];
        S.L4[i] = h.L[9 * c + 4];
        S.L5[i] = h.L[9 * c + 5];
        S.L6[i] = h.L[9 * c + 6];
        S.L7[i] = h.L[9 * c + 7];
        S.L8[i] = h.L[9 * c + 8];
    }
    if (h.part_id)
        S.part_id[i] = h.part_id[c];
    if (h.cellID)
        S.cellID[i] = int(h.cellID[c]);
    if (h.b)
        S.b[i] = h.b[c];
    if (h.surfzone)
        S.surfzone[i] = h.surfzone[c];
    if (h.internal)
        S.internal[i] = h.internal[c];
}

__global__ void k_unpack(Level S, StageView h, const int* __restrict__ slot_of, double dx, int n)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n)
        return;
    const int i = slot_of[c];
    const double4 p0 = S.P0[i], p1 = S.P1[i], p2 = S.P2[i], th = S.TH[i];
#define V3(dst, r)            \
    if (h.dst)                \
    {                         \
        h.dst[3 * c] = r.x;   \
        h.dst[3 * c + 1] = r.y; \
        h.dst[3 * c + 2] = r.z; \
    }
    V3(xi, p0)
    V3(v, p1)
    V3(vPert, p2)
    if (h.rho)
        h.rho[c] = p1.w;
    if (h.p)
        h.p[c] = th.x;
    if (h.m)
        h.m[c] = th.y;
    if (h.woccl)
        h.woccl[c] = th.z;
    if (h.cellRho)
        h.cellRho[c] = th.w;
    const double4 p3 = S.P3[i], p4 = S.P4[i], np = S.NP[i], acc = S.ACC[i], af = S.AF[i], av = S.AV[i], cv = S.CV[i],
                  bn = S.BN[i], sc = S.SC[i];
    V3(gradRho, p3)
    if (h.lam)
        h.lam[c] = p3.w;
    V3(norm, np)
    if (h.lam_nb)
        h.lam_nb[c] = np.w;
    if (h.surf)
        h.surf[c] = (p4.w != 0.0) ? 1 : 0;
    V3(acc, acc)
    if (h.Rrho)
        h.Rrho[c] = acc.w;
    V3(Af, af)
    if (h.deltaD)
        h.deltaD[c] = af.w;
    V3(aVisc, av)
    if (h.curve)
        h.curve[c] = av.w;
    if (h.norm_curve)
        h.norm_curve[c] = dx * av.w; /* Geometry.cpp:259 */
    V3(cellV, cv)
    if (h.cellP)
        h.cellP[c] = cv.w;
    V3(bNorm, bn)
    if (h.y)
        h.y[c] = bn.w;
    if (h.colourG)
        h.colourG[c] = sc.x;
    if (h.colour)
        h.colour[c] = sc.y;
    if (h.kernsum)
        h.kernsum[c] = sc.z;
    if (h.pDist)
        h.pDist[c] = sc.w;
#undef V3
    if (h.L)
    {
        h.L[9 * c] = S.L0[i];
        h.L[9 * c + 1] = S.L1[i];
        h.L[9 * c + 2] = S.L2[i];
        h.L[9 * c + 3] = S.L3[i];
        h.L[9 * c + 4] = S.L4[i];
        h.L[9 * c + 5] = S.L5[i];
        h.L[9 * c + 6] = S.L6[i];
        h.L[9 * c + 7] = S.L7[i];
        h.L[9 * c + 8] = S.L8[i];
    }
    if (h.part_id)
        h.part_id[c] = S.part_id[i];
    if (h.cellID)
        h.cellID[c] = S.cellID[i];
    if (h.b)
        h.b[c] = S.b[i];
    if (h.surfzone)
        h.surfzone[c] = S.surfzone[i];
    if (h.internal)
        h.internal[c] = S.internal[i] & 0xFF; /* the upper bytes count FindCell's failed queries (ipt_n_failed) */
}

__global__ void k_counts_out(const int* __restrict__ ncount, const int* __restrict__ slot_of, long long* out, int n)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n)
        out[c] = ncount[slot_of[c]] + 1;
}

struct FieldSpec
{
    size_t host_off;  // offset of the pointer inside FjsphStateView
    size_t stage_off; // offset of the pointer inside StageView
    int bytes;        // bytes per particle
};
#define FS(name, bytes) {offsetof(FjsphStateView, name), offsetof(StageView, name), bytes}
const FieldSpec kFields[] = {
    FS(part_id, 8), FS(cellID, 8),   FS(b, 4),       FS(surf, 4),       FS(surfzone, 4), FS(internal, 4), FS(xi, 24),
    FS(v, 24),      FS(acc, 24),     FS(Af, 24),     FS(aVisc, 24),     FS(cellV, 24),   FS(gradRho, 24), FS(norm, 24),
    FS(bNorm, 24),  FS(vPert, 24),   FS(L, 72),      FS(Rrho, 8),       FS(rho, 8),      FS(p, 8),        FS(m, 8),
    FS(curve, 8),   FS(norm_curve, 8), FS(woccl, 8), FS(pDist, 8),      FS(deltaD, 8),   FS(cellP, 8),    FS(cellRho, 8),
    FS(colourG, 8), FS(colour, 8),   FS(lam, 8),     FS(lam_nb, 8),     FS(kernsum, 8),  FS(y, 8),
};
constexpr size_t kStageBytesPerParticle = 8 * 2 + 4 * 4 + 24 * 10 + 72 + 8 * 17;

// lays the non-null host fields out in the staging buffer; returns the device-side view
int stage_layout(FjsphEngine* e, const FjsphStateView* s, StageView* dv, std::vector<std::pair<void*, void*>>* copies,
                 std::vector<size_t>* sizes, size_t* stage_off = nullptr)
{
    std::memset(dv, 0, sizeof(*dv));
    const size_t n = size_t(s->n);
    size_t off = stage_off ? *stage_off : 0;
    for (const FieldSpec& f : kFields)
    {
        void* hp = *(void* const*)((const char*)s + f.host_off);
        if (!hp)
            continue;
        const size_t bytes = n * size_t(f.bytes);
        if (off + bytes > e->stage_bytes)
        {
            fj_set_error("staging buffer too small");
            return FJSPH_ERR_CAPACITY;
        }
        void* dp = (char*)e->stage + off;
        *(void**)((char*)dv + f.stage_off) = dp;
        copies->push_back(std::make_pair(hp, dp));
        sizes->push_back(bytes);
        off += (bytes + 255) & ~size_t(255);
    }
    if (stage_off)
        *stage_off = off;
    return FJSPH_OK;
}

// host arrays -> the device staging buffer (one H2D copy per field), then k_pack scatters them into the slots of a level.
// both_levels: the same staged data go into pn and pnp1 (one trip over PCIe, two scatters on the device).
int upload_fields(FjsphEngine* e, int level, const FjsphStateView* s, bool both_levels = false, size_t* stage_off = nullptr)
{
    StageView dv;
    std::vector<std::pair<void*, void*>> copies;
    std::vector<size_t> sizes;
    int st = stage_layout(e, s, &dv, &copies, &sizes, stage_off);
    if (st)
        return st;
    for (size_t k = 0; k < copies.size(); ++k)
        FJ_CUDA(cudaMemcpyAsync(copies[k].second, copies[k].first, sizes[k], cudaMemcpyHostToDevice, e->stream));
    const int n = int(s->n);
    e->launches += both_levels ? 2 : 1;
    k_pack<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[level], dv, e->slot_of, n);
    if (both_levels)
        k_pack<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1 - level], dv, e->slot_of, n);
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

int rebuild_blk(FjsphEngine* e)
{
    std::vector<long long> ranges;
    for (const HostBlock& B : e->blocks)
    {
        ranges.push_back(B.first);
        ranges.push_back(B.second);
    }
    long long* d = nullptr;
    FJ_CUDA(cudaMalloc(&d, ranges.size() * sizeof(long long)));
    FJ_CUDA(cudaMemcpyAsync(d, ranges.data(), ranges.size() * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    const int n = int(e->n);
    e->launches++;
    k_assign_blk<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->oidx, d, int(e->blocks.size()), e->blk, n);
    FJ_CUDA(cudaGetLastError());
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    cudaFree(d);
    return FJSPH_OK;
}

void default_blocks(FjsphEngine* e)
{
    e->blocks.clear();
    e->n_bound_blocks = 0;
    auto mk = [](int64_t a, int64_t b, int fluid) {
        HostBlock B;
        B.first = a;
        B.second = b;
        B.is_fluid = fluid;
        B.bound_solver = FJSPH_PRESSURE_G;
        B.no_slip = 0;
        B.block_type = 0;
        B.fixed_vel_or_dynamic = 0;
        B.vels.assign(3, 0.0);
        for (int d = 0; d < 3; ++d) B.insert_norm[d] = B.delete_norm[d] = B.aero_norm[d] = 9999999.0;
        B.insconst = B.delconst = B.aeroconst = 9999999.0;
        return B;
    };
    if (e->bound_points > 0)
    {
        e->blocks.push_back(mk(0, e->bound_points, 0));
        e->n_bound_blocks = 1;
    }
    e->blocks.push_back(mk(e->bound_points, e->n, 1));
}

} // namespace

// ================================================================== C ABI
extern "C" {

const char* fjsph_last_error(void) { return g_err; }
const char* fjsph_version(void) { return "fjsph_b200 0.1 (sm_100a)"; }

int fjsph_create(const FjsphParams* p, int device, int64_t capacity, FjsphEngine** out)
{
    if (!p || !out || capacity <= 0)
    {
        fj_set_error("fjsph_create: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    if (capacity > int64_t(FJ_IDX_MASK))
    {
        fj_set_error("capacity %lld exceeds the 2^27-1 particles one engine can index", (long long)capacity);
        return FJSPH_ERR_CAPACITY;
    }
    int st = validate_params(*p);
    if (st)
        return st;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
    {
        fj_set_error("no CUDA device available (%s): the engine has no CPU fallback",
                     ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
        return FJSPH_ERR_CUDA;
    }
    FJ_CUDA(cudaSetDevice(device));
    FjsphEngine* e = new FjsphEngine();
    e->P = *p;
    e->device = device;
    e->cap = capacity;
    fj_refresh_constants(e);
    cudaDeviceProp prop;
    FJ_CUDA(cudaGetDeviceProperties(&prop, device));
    e->n_sm = prop.multiProcessorCount;
    FJ_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    FJ_CUDA(cudaEventCreate(&e->ev0));
    FJ_CUDA(cudaEventCreate(&e->ev1));
    const size_t cap = size_t(capacity);
    for (int l = 0; l < 3; ++l)
    {
#define X(T, f) FJ_CUDA(cudaMalloc(&e->lv[l].f, cap * sizeof(T)));
        FJ_LEVEL_FIELDS(X)
#undef X
    }
    FJ_CUDA(cudaMalloc(&e->oidx, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->oidx_tmp, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->slot_of, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->blk, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->blk_tmp, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->key, cap * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->rank_in_cell, cap * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->perm, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->perm2, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->ncount, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->xref, cap * sizeof(double4)));
    FJ_CUDA(cudaMalloc(&e->x0, cap * sizeof(double4)));
    e->skin = 0.4 * e->P.particle_step; /* default skin: 0.4 dx = 5 % of the support radius at H_fac = 2 */
    if (const char* v = std::getenv("FJSPH_B200_ROW_AXIS")) /* neighbours.cu, rebuild_skin */
        e->row_axis = std::atoi(v);
    if (const char* v = std::getenv("FJSPH_B200_ROW_WIDTH"))
        e->row_width_cells = std::atof(v);
    if (const char* v = std::getenv("FJSPH_B200_MAX_KEY_BITS"))
        e->max_key_bits = std::min(29, std::max(6, std::atoi(v)));
    if (const char* v = std::getenv("FJSPH_B200_LIST_STATS"))
        e->list_stats = std::atoi(v) != 0;
    if (const char* v = std::getenv("FJSPH_B200_SWEEP_WARPS"))
        e->sweep_warps = (std::atoi(v) == 8) ? 8 : 4;
    if (const char* parts = std::getenv("FJSPH_B200_UPLOAD_PARTS")) /* 2: fjsph_step_host uploads x | the rest (no early prestep) */
        e->upload_parts = std::atoi(parts) == 2 ? 2 : 3;
    if (const char* split = std::getenv("FJSPH_B200_SPLIT_SURFACE"))
        e->split_surface_sweep = std::string(split) != "0";
    if (const char* v = std::getenv("FJSPH_B200_SPLIT_BELOW")) /* near-surface warp fraction below which the sweep splits */
        e->split_surface_below = std::atof(v);
    FJ_CUDA(cudaMalloc(&e->near_inlet, cap * sizeof(int)));
    FJ_CUDA(cudaMemset(e->near_inlet, 0, cap * sizeof(int)));
    FJ_CUDA(cudaMalloc(&e->rk_sum_v, cap * sizeof(double4)));
    FJ_CUDA(cudaMalloc(&e->rk_sum_a, cap * sizeof(double4)));
    const size_t red_n = std::max<size_t>((cap + 127) / 128 + 16, 2048 * 8);
    FJ_CUDA(cudaMalloc(&e->red, red_n * sizeof(double)));
    e->red_cap = red_n;
    FJ_CUDA(cudaMalloc(&e->red_out, 16 * sizeof(double)));
    FJ_CUDA(cudaMallocHost(&e->h_red, 16 * sizeof(double)));
    FJ_CUDA(cudaMalloc(&e->d_flag, 4 * sizeof(int)));
    FJ_CUDA(cudaMallocHost(&e->h_flag, 4 * sizeof(int)));
    std::memset(e->h_flag, 0, 4 * sizeof(int));
    e->stage_bytes = cap * kStageBytesPerParticle + 64 * 256;
    FJ_CUDA(cudaMalloc(&e->stage, e->stage_bytes));
    *out = e;
    return FJSPH_OK;
}

int fjsph_destroy(FjsphEngine* e)
{
    if (!e)
        return FJSPH_OK;
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    cudaStreamSynchronize(e->stream);
    for (int l = 0; l < 3; ++l)
    {
#define X(T, f) cudaFree(e->lv[l].f);
        FJ_LEVEL_FIELDS(X)
#undef X
    }
    for (HostBlock& B : e->blocks)
    {
        if (B.d_back)
            cudaFree(B.d_back);
        if (B.d_buffer)
            cudaFree(B.d_buffer);
    }
    if (e->scan_particles)
        cudaFree(e->scan_particles);
    fj_free_mesh(e);
    void* ptrs[] = {e->oidx,       e->oidx_tmp, e->slot_of,  e->blk,      e->blk_tmp,   e->key,     e->rank_in_cell,
                    e->perm,       e->perm2,    e->ncount,   e->near_inlet, e->rk_sum_v, e->rk_sum_a, e->red,
                    e->red_out,    e->d_flag,   e->stage,    e->cell_count, e->cell_start, e->scan_tmp, e->mtab_y,
                    e->erun,       e->erows,    e->srun,     e->srows,      e->xref,       e->x0,       e->warp_start,
                    e->row_warps,  e->row_scan_tmp, e->row_off};
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    if (e->h_red)
        cudaFreeHost(e->h_red);
    if (e->h_flag)
        cudaFreeHost(e->h_flag);
    cudaEventDestroy(e->ev0);
    cudaEventDestroy(e->ev1);
    fj_timers_flush(e);
    for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
    if (e->upload_stream)
    {
        cudaStreamSynchronize(e->upload_stream);
        cudaStreamDestroy(e->upload_stream);
        cudaEventDestroy(e->ev_upload_x);
        cudaEventDestroy(e->ev_upload_b);
        cudaEventDestroy(e->ev_upload);
    }
    if (e->slab.comm_stream)
    {
        cudaStreamSynchronize(e->slab.comm_stream);
        cudaStreamDestroy(e->slab.comm_stream);
        cudaEventDestroy(e->slab.ev_ready);
        cudaEventDestroy(e->slab.ev_done);
    }
    if (e->own_stream)
        cudaStreamDestroy(e->stream);
    delete e;
    return FJSPH_OK;
}

int fjsph_take_deleted(FjsphEngine* e, FjsphDeleted* out, int64_t capacity, int64_t* n_out)
{
    if (!e || !n_out || (out && capacity < 0))
    {
        fj_set_error("take_deleted: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    if (!out)
    {
        *n_out = int64_t(e->deleted.size());
        return FJSPH_OK;
    }
    const size_t n = std::min<size_t>(e->deleted.size(), size_t(capacity));
    if (n)
        std::memcpy(out, e->deleted.data(), n * sizeof(FjsphDeleted));
    e->deleted.erase(e->deleted.begin(), e->deleted.begin() + std::ptrdiff_t(n));
    *n_out = int64_t(n);
    return FJSPH_OK;
}

int fjsph_get_params(FjsphEngine* e, FjsphParams* out)
{
    *out = e->P;
    return FJSPH_OK;
}
int fjsph_set_params(FjsphEngine* e, const FjsphParams* in)
{
    int st = validate_params(*in);
    if (st)
        return st;
    if (in->sr != e->P.sr || in->particle_step != e->P.particle_step)
    {
        /* the support radius changed: keep skin/dx, drop both lists */
        e->skin = (e->P.particle_step > 0.0 ? e->skin / e->P.particle_step : 0.4) * in->particle_step;
        e->skin_valid = false;
        e->list_valid = false;
    }
    e->P = *in;
    fj_refresh_constants(e);
    return FJSPH_OK;
}

int fjsph_set_blocks(FjsphEngine* e, int32_t n_blocks, const FjsphBlock* blocks)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    std::vector<HostBlock> out;
    int n_bound = 0;
    bool seen_fluid = false;
    for (int k = 0; k < n_blocks; ++k)
    {
        const FjsphBlock& b = blocks[k];
        HostBlock B;
        B.first = b.first;
        B.second = b.second;
        B.is_fluid = b.is_fluid;
        B.bound_solver = b.bound_solver;
        B.no_slip = b.no_slip;
        B.block_type = b.block_type;
        B.fixed_vel_or_dynamic = b.fixed_vel_or_dynamic;
        for (int t = 0; t < b.n_times; ++t) B.times.push_back(b.times[t]);
        const int nv = std::max(1, b.n_times);
        for (int t = 0; t < 3 * nv; ++t) B.vels.push_back(b.vels ? b.vels[t] : 0.0);
        for (int d = 0; d < 3; ++d)
        {
            B.insert_norm[d] = b.insert_norm[d];
            B.delete_norm[d] = b.delete_norm[d];
            B.aero_norm[d] = b.aero_norm[d];
        }
        B.insconst = b.insconst;
        B.delconst = b.delconst;
        B.aeroconst = b.aeroconst;
        for (int i = 0; i < b.n_back; ++i)
        {
            B.back.push_back(b.back[i]);
            std::vector<int64_t> buf;
            for (int j = 0; j < b.n_buf; ++j) buf.push_back(b.buffer[size_t(i) * b.n_buf + j]);
            B.buffer.push_back(buf);
        }
        if (b.is_fluid)
            seen_fluid = true;
        else
        {
            if (seen_fluid)
            {
                fj_set_error("set_blocks: boundary blocks must precede fluid blocks (Init.cpp:298-475)");
                return FJSPH_ERR_INVALID;
            }
            n_bound++;
        }
        out.push_back(B);
    }
    if (out.empty() || !seen_fluid)
    {
        fj_set_error("set_blocks: at least one fluid block is required");
        return FJSPH_ERR_INVALID;
    }
    for (HostBlock& B : e->blocks)
    {
        if (B.d_back)
            cudaFree(B.d_back);
        if (B.d_buffer)
            cudaFree(B.d_buffer);
    }
    e->blocks = out;
    e->inlet_tables_dirty = true;
    e->n_bound_blocks = n_bound;
    if (e->n > 0)
        return rebuild_blk(e);
    return FJSPH_OK;
}

int fjsph_upload_state(FjsphEngine* e, const FjsphStateView* s, int64_t bound_points)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (!s || s->n <= 0 || !s->xi || !s->rho || !s->p || !s->m || !s->b)
    {
        fj_set_error("upload_state: xi, rho, p, m and b are required");
        return FJSPH_ERR_INVALID;
    }
    if (s->n > e->cap)
    {
        fj_set_error("upload_state: %lld particles exceed the capacity %lld given to fjsph_create", (long long)s->n,
                     (long long)e->cap);
        return FJSPH_ERR_CAPACITY;
    }
    if (bound_points < 0 || bound_points > s->n)
    {
        fj_set_error("upload_state: bound_points out of range");
        return FJSPH_ERR_INVALID;
    }
    const bool keep_blocks = !e->blocks.empty() && e->blocks.back().second == s->n && e->n == s->n &&
                             e->bound_points == bound_points;
    /* A host that round-trips the SAME particle set every step (fjsph_step_host) keeps the engine's cell order and its
       superset neighbour list: the uploaded values land in the slots the particles already occupy (slot_of), and the
       next neighbour build decides from the displacements against the list's reference positions whether the list
       still holds -- exactly as for device-resident state.  Anything else (another count, other blocks, slab mode, no
       list yet) resets the slots to the caller's order and drops the list. */
    const bool keep_order = keep_blocks && e->skin_valid && e->skin_n == s->n && e->n_owned == s->n && !e->slab.on;
    e->n = s->n;
    e->n_owned = s->n;
    e->bound_points = bound_points;
    e->list_valid = false;
    /* FJSPH's id counter only grows (Init.cpp:283, inlet.cpp:610): a host that hands its ids back after erasures and
       insertions must not see them reused */
    {
        int64_t next = s->n;
        if (s->part_id)
            for (int64_t k = 0; k < s->n; ++k) next = std::max(next, s->part_id[k] + 1);
        e->next_part_id = (keep_blocks && s->part_id) ? std::max(e->next_part_id, next) : next;
    }
    e->inlet_tables_dirty = true;
    const int n = int(e->n);
    e->launches += 2;
    if (!keep_order)
    {
        e->skin_valid = false;
        k_identity_index<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->oidx, e->slot_of, n);
    }
    k_init_level<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1], e->oidx, e->C, e->P.rho_g, e->P.p_ref, n);
    FJ_CUDA(cudaGetLastError());
    int st = upload_fields(e, 1, s);
    if (st)
        return st;
    st = fj_copy_level(e, 0, 1); /* pnp1 = pn, Init.cpp:496 */
    if (st)
        return st;
    if (!keep_blocks)
        default_blocks(e);
    st = rebuild_blk(e);
    if (st)
        return st;
    e->maxShift = 0.0;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}

// fjsph_upload_state for a host that round-trips the SAME particle set (fjsph_step_host): positions cross PCIe first, on
// the engine's stream, and the neighbour build of the step that follows needs nothing else; the other fields follow on a
// second stream beside it.  Returns with that second half in flight (e->upload_pending).  Falls back to the plain
// upload whenever the cell order cannot be kept.
static int upload_state_split(FjsphEngine* e, const FjsphStateView* s, int64_t bound_points)
{
    const bool keep = s && s->xi && s->rho && s->p && s->m && s->b && !e->blocks.empty() && e->blocks.back().second == s->n &&
                      e->n == s->n && e->bound_points == bound_points && e->skin_valid && e->skin_n == s->n &&
                      e->n_owned == s->n && !e->slab.on;
    if (!keep)
        return fjsph_upload_state(e, s, bound_points);
    cudaSetDevice(e->device);
    if (!e->upload_stream)
    {
        FJ_CUDA(cudaStreamCreateWithFlags(&e->upload_stream, cudaStreamNonBlocking));
        FJ_CUDA(cudaEventCreateWithFlags(&e->ev_upload_x, cudaEventDisableTiming));
        FJ_CUDA(cudaEventCreateWithFlags(&e->ev_upload_b, cudaEventDisableTiming));
        FJ_CUDA(cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming));
    }
    /* Three parts when the host hands over none of the fields dSPH_PreStep writes (gradRho, lam, norm, surf, lam_nb, the
       colour terms, pDist, L): x | rho, m, b -- all the prestep reads -- | the rest.  The step then runs the neighbour build
       beside part two and the prestep beside part three (fj_integrate_no_update). */
    const bool early = e->upload_parts == 3 && !s->gradRho && !s->lam && !s->norm && !s->surf && !s->lam_nb && !s->colourG && !s->colour && !s->kernsum &&
                       !s->pDist && !s->L;
    e->list_valid = false;
    {
        int64_t next = s->n;
        if (s->part_id)
            for (int64_t k = 0; k < s->n; ++k) next = std::max(next, s->part_id[k] + 1);
        e->next_part_id = s->part_id ? std::max(e->next_part_id, next) : next; /* ids handed back: the counter only grows */
    }
    e->inlet_tables_dirty = true;
    e->maxShift = 0.0;
    const int n = int(e->n);
    e->launches += 1;
    k_init_level<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1], e->oidx, e->C, e->P.rho_g, e->P.p_ref, n);
    FJ_CUDA(cudaGetLastError());
    if (early)
    {
        /* pn = pnp1 (Init.cpp:496) for the fields the prestep will overwrite in pnp1: their upload-time values are the
           defaults just set, so pn takes them now; the other fields follow when the upload is complete */
        int stc = fj_copy_level(e, 0, 1, 1);
        if (stc)
            return stc;
    }
    /* first part: positions */
    FjsphStateView xs;
    std::memset(&xs, 0, sizeof(xs));
    xs.n = s->n;
    xs.xi = s->xi;
    size_t stage_off = 0;
    int st = upload_fields(e, 1, &xs, false, &stage_off);
    if (st)
        return st;
    FJ_CUDA(cudaEventRecord(e->ev_upload_x, e->stream));
    /* the other fields on the upload stream behind the positions: what the prestep reads first (three-part upload), then
       the rest; in the two-part form pn = pnp1 (Init.cpp:496) follows on the same stream */
    FjsphStateView rest = *s;
    rest.xi = nullptr;
    FJ_CUDA(cudaStreamWaitEvent(e->upload_stream, e->ev_upload_x, 0));
    cudaStream_t main_stream = e->stream;
    const bool timers = e->timers_on;
    e->timers_on = false; /* the lazily resolved event timers belong to the engine's stream */
    e->stream = e->upload_stream;
    if (early)
    {
        FjsphStateView second;
        std::memset(&second, 0, sizeof(second));
        second.n = s->n;
        second.rho = s->rho;
        second.m = s->m;
        second.b = s->b;
        rest.rho = nullptr;
        rest.m = nullptr;
        rest.b = nullptr;
        st = upload_fields(e, 1, &second, false, &stage_off);
        if (!st && cudaEventRecord(e->ev_upload_b, e->upload_stream) != cudaSuccess)
            st = FJSPH_ERR_CUDA;
    }
    if (!st)
        st = upload_fields(e, 1, &rest, false, &stage_off);
    if (!st && !early)
        st = fj_copy_level(e, 0, 1);
    e->stream = main_stream;
    e->timers_on = timers;
    if (st)
    {
        cudaStreamSynchronize(e->upload_stream);
        return st;
    }
    FJ_CUDA(cudaEventRecord(e->ev_upload, e->upload_stream));
    e->upload_pending = true;
    e->upload_early = early;
    return FJSPH_OK;
}

int fjsph_upload_level(FjsphEngine* e, int level, const FjsphStateView* s)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (level < 0 || level > 1 || !s || s->n != e->n)
    {
        fj_set_error("upload_level: bad level or particle count (have %lld)", (long long)e->n);
        return FJSPH_ERR_INVALID;
    }
    if (s->xi)
        e->list_valid = false;
    int st = upload_fields(e, level, s);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}

int fjsph_upload_owned(FjsphEngine* e, const FjsphStateView* s)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (!s || s->n != e->n_owned)
    {
        fj_set_error("upload_owned: view holds %lld particles, this rank owns %lld", s ? (long long)s->n : -1LL,
                     (long long)e->n_owned);
        return FJSPH_ERR_INVALID;
    }
    int st = upload_fields(e, 1, s, true);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream)); /* the host arrays and the staging buffer are free again */
    e->list_valid = false;
    return fj_halo_exchange(e, 1, FJ_HX_P0 | FJ_HX_P1 | FJ_HX_P2 | FJ_HX_TH);
}

int fjsph_download_state(FjsphEngine* e, int level, FjsphStateView* s)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (level < 0 || level > 1 || !s)
    {
        fj_set_error("download_state: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    if (s->n < e->n_owned)
    {
        fj_set_error("download_state: view holds %lld particles, engine has %lld", (long long)s->n,
                     (long long)e->n_owned);
        return FJSPH_ERR_INVALID;
    }
    FjsphStateView view = *s;
    view.n = e->n_owned;
    StageView dv;
    std::vector<std::pair<void*, void*>> copies;
    std::vector<size_t> sizes;
    int st = stage_layout(e, &view, &dv, &copies, &sizes);
    if (st)
        return st;
    const int n = int(e->n_owned);
    e->launches++;
    k_unpack<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[level], dv, e->slot_of, e->P.dx, n);
    FJ_CUDA(cudaGetLastError());
    for (size_t k = 0; k < copies.size(); ++k)
        FJ_CUDA(cudaMemcpyAsync(copies[k].first, copies[k].second, sizes[k], cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    s->n = e->n_owned;
    return FJSPH_OK;
}

int64_t fjsph_count(FjsphEngine* e) { return e ? e->n_owned : -1; }

int fjsph_build_neighbours(FjsphEngine* e)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    return fj_build_neighbours(e);
}

int fjsph_neighbour_counts(FjsphEngine* e, int64_t* counts)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (!e->list_valid)
    {
        fj_set_error("neighbour_counts: list not built");
        return FJSPH_ERR_STATE;
    }
    const int n = int(e->n_owned);
    long long* d = (long long*)e->stage;
    k_counts_out<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->ncount, e->slot_of, d, n);
    FJ_CUDA(cudaGetLastError());
    FJ_CUDA(cudaMemcpyAsync(counts, d, size_t(n) * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}

// Debug / parity view of the neighbour list in the reference's shape: CSR over caller indices, self
// included, ascending j.  offsets must be the exclusive prefix sum of fjsph_neighbour_counts.
int fjsph_get_neighbours(FjsphEngine* e, const int64_t* offsets, int64_t* idx)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    if (!e->list_valid)
    {
        fj_set_error("get_neighbours: list not built");
        return FJSPH_ERR_STATE;
    }
    const size_t n = size_t(e->n_owned);
    const size_t total = size_t(offsets[n]);
    long long *d_off = nullptr, *d_idx = nullptr;
    int* d_bad = nullptr;
    FJ_CUDA(cudaMalloc(&d_off, (n + 1) * sizeof(long long)));
    FJ_CUDA(cudaMalloc(&d_idx, std::max<size_t>(total, 1) * sizeof(long long)));
    FJ_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    FJ_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), e->stream));
    FJ_CUDA(cudaMemcpyAsync(d_off, offsets, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    int st = fj_neighbours_to_csr(e, d_off, d_idx, d_bad);
    int bad = 0;
    if (!st)
    {
        cudaMemcpyAsync(idx, d_idx, total * sizeof(long long), cudaMemcpyDeviceToHost, e->stream);
        cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
        if (cudaStreamSynchronize(e->stream) != cudaSuccess)
            st = fj_cuda_fail(cudaGetLastError(), "get_neighbours", __FILE__, __LINE__);
    }
    cudaFree(d_off);
    cudaFree(d_idx);
    cudaFree(d_bad);
    if (st)
        return st;
    if (bad)
    {
        fj_set_error("get_neighbours: offsets do not match neighbour_counts");
        return FJSPH_ERR_INVALID;
    }
    for (size_t c = 0; c < n; ++c) std::sort(idx + offsets[c], idx + offsets[c + 1]);
    return FJSPH_OK;
}

int fjsph_prestep(FjsphEngine* e, double* npd)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    return fj_prestep(e, npd);
}
int fjsph_aero_velocity(FjsphEngine* e)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_aero_velocity(e);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}
int fjsph_detect_surface(FjsphEngine* e)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_surface_and_dissipation(e, true, false);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}
int fjsph_dissipation(FjsphEngine* e)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_surface_and_dissipation(e, false, true);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}
int fjsph_shift(FjsphEngine* e)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_shift(e);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}
int fjsph_forces(FjsphEngine* e, double npd)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_forces(e, 1, npd);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}
int fjsph_nb_iter(FjsphEngine* e, double npd, double* errsum)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    return fj_nb_iter(e, npd, errsum);
}
int fjsph_find_timestep(FjsphEngine* e, double* dt)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    return fj_find_timestep(e, dt);
}
int fjsph_integrate_no_update(FjsphEngine* e, FjsphStepStats* s)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    return fj_integrate_no_update(e, s);
}
int fjsph_step(FjsphEngine* e, FjsphStepStats* s)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    int st = fj_step(e, s);
    if (st)
        return st;
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    return FJSPH_OK;
}

int fjsph_step_host(FjsphEngine* e, const FjsphStateView* in, int64_t bound_points, int32_t n_steps,
                    FjsphStateView* out, FjsphStepStats* last)
{
    int st = n_steps > 0 ? upload_state_split(e, in, bound_points) : fjsph_upload_state(e, in, bound_points);
    if (st)
        return st;
    for (int k = 0; k < n_steps; ++k)
    {
        st = fjsph_step(e, last);
        if (st)
            break;
    }
    if (e->upload_pending)
    { /* a step that failed before it consumed the upload: the host arrays must be free again on return */
        cudaStreamSynchronize(e->upload_stream);
        e->upload_pending = false;
    }
    e->upload_early = false;
    if (st)
        return st;
    return fjsph_download_state(e, 1, out);
}

int fjsph_timers_reset(FjsphEngine* e)
{
    fj_timers_flush(e);
    e->timers.clear();
    return FJSPH_OK;
}
int fjsph_timers_enable(FjsphEngine* e, int on)
{
    fj_timers_flush(e);
    e->timers_on = on != 0;
    return FJSPH_OK;
}
int fjsph_timers_get(FjsphEngine* e, int32_t cap, char* names, double* ms, int64_t* launches, int64_t* calls,
                     int32_t* n_out)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    fj_timers_flush(e);
    const int n = std::min<int>(cap, int(e->timers.size()));
    for (int k = 0; k < n; ++k)
    {
        std::snprintf(names + size_t(k) * 32, 32, "%s", e->timers[k].name.c_str());
        ms[k] = e->timers[k].ms;
        launches[k] = e->timers[k].launches;
        if (calls)
            calls[k] = e->timers[k].calls;
    }
    *n_out = n;
    return FJSPH_OK;
}
int64_t fjsph_launch_count(FjsphEngine* e) { return e->launches; }

int fjsph_set_stream(FjsphEngine* e, void* cuda_stream)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e); /* slab mode: an exchange may still be in flight on the comm stream */
    fj_timers_flush(e);
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (e->own_stream && e->stream)
        cudaStreamDestroy(e->stream);
    if (cuda_stream)
    {
        e->stream = (cudaStream_t)cuda_stream;
        e->own_stream = false;
    }
    else
    {
        FJ_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->own_stream = true;
    }
    return FJSPH_OK;
}

int fjsph_set_skin(FjsphEngine* e, double skin_over_dx)
{
    if (!(skin_over_dx >= 0.0) || skin_over_dx > 2.0)
    {
        fj_set_error("set_skin: skin/dx must lie in [0, 2]");
        return FJSPH_ERR_INVALID;
    }
    e->skin = skin_over_dx * e->P.particle_step;
    e->skin_valid = false;
    e->list_valid = false;
    return FJSPH_OK;
}

int fjsph_set_owned(FjsphEngine* e, int64_t n_owned)
{
    if (n_owned < 0 || n_owned > e->n)
    {
        fj_set_error("set_owned: out of range");
        return FJSPH_ERR_INVALID;
    }
    e->n_owned = n_owned;
    return FJSPH_OK;
}

} // extern "C"
