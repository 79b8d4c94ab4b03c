// sweeps.cu — the FP64 pair sweeps of the WCSPH step.  One thread per particle; a warp is 32 consecutive particles of
// one ROW (neighbours.cu) and walks the run lists row by row, so at every step its lanes gather 32 (nearly) consecutive
// neighbour records: coalesced 256-bit loads instead of 32 scattered sectors.  The warps of a CTA sit in adjacent rows
// over the same stretch of the row axis and share those records through L1.
//
// Replaces, with results within 1e-10 (normwise) of the CPU oracle:
//   dSPH_PreStep        reference src/Shifting.cpp:12-123
//   get_aero_velocity   reference src/Resid.cpp:569-611 (constVel)
//   Detect_Surface      reference src/Geometry.cpp:14-280   (loops 1 / 2 / 3)
//   dissipation_terms   reference src/Shifting.cpp:126-186  (fused with Detect_Surface loop 1)
//   particle_shift      reference src/Shifting.cpp:189-290
//   get_acc_and_Rrho    reference src/Resid.cpp:243-469, Kernel.h, Aero.h:10-98
//   Get_Boundary_Pressure / Boundary_Ghost / Set_No_Slip / Boundary_DBC   reference src/Resid.cpp:21-186
//
// Sweep fusion: prestep | surface loop 1 + dissipation | surface loops 2+3 | shifting | force.
#include <cmath>
#include <type_traits>

#include "engine.cuh"
#include "small_matrix.cuh"

namespace
{

// CTA shapes of the pair sweeps: WARPS rows per CTA; resident CTAs per SM chosen so that the heaviest instantiation
// keeps its registers (8 warps x 1 CTA: up to 255; 4 warps x 3 CTAs: up to 168; the lean bulk sweep twice that)
constexpr int min_blocks(int warps, bool lean) { return warps == 8 ? (lean ? 2 : 1) : (lean ? 4 : 3); }
#define FJ_PI 3.14159265358979323846
#define FJ_FULL 0xffffffffu

struct RunView
{
    const uint2* __restrict__ erun;
    const int* __restrict__ erows;
    const int* __restrict__ ncount;
    const double4* __restrict__ x0;
    int ecap;
};

// The reference carries d^2 in the neighbour list (OUTL = vector<vector<pair<idx, dist2>>>, Var.h:889-890) and
// every pair loop reads r = sqrt(jj.second) from it, so r stays FROZEN at its list-build value while
// Rji = xj - xi follows the positions through the Newmark-Beta sub-iterations / RK stages.  The run lists store no
// per-pair value: a sweep that runs on the positions the list was built on (FROZEN = false: prestep, surface,
// shifting, the first force evaluation after a build) takes r from the positions it gathers anyway; once the
// particles have moved (FROZEN = true) it gathers the build-time positions x0 as well and takes r from those.
__device__ __forceinline__ uint2 ld_desc(const uint2* p)
{
    uint2 v;
    asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// 32-byte record gather with ONE 256-bit load (LDG.E.256).  Only for arrays no thread writes during the kernel.
__device__ __forceinline__ double4 gather(const double4* __restrict__ base, unsigned j)
{
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(base + j));
    return v;
}
// same, through the coherent path: for arrays some thread of the SAME kernel updates (its own record only)
__device__ __forceinline__ double4 gather_rw(const double4* base, unsigned j)
{
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(base + j) : "memory");
    return v;
}

// Branch-free, two neighbours at a time.  load(j) gathers the records of neighbour j unconditionally -- a lane with
// nobody to visit at a step is handed its OWN index (self), a valid address whose pair terms vanish or are masked by
// the `take` flags body(recA, takeA, recB, takeB) receives -- so the loop body is one basic block and the compiler
// interleaves the two independent FP64 dependency chains of the two pairs: the pair algebra is a long serial chain
// (d^2 -> rsqrt -> r -> kernel gradient -> accumulators) and with ~3 warps per scheduler a single chain leaves the FP64
// pipe idle two cycles out of three (ncu, profiles/r2b_*).
//
// L1 staging: on entering a slot every lane queues `cp.async.ca` copies (LDGSTS: global -> L1 -> a scratch word of shared
// memory nobody reads) of the first and the last record of its window in the NEXT slot, and of its descriptor four slots
// on.  The copies cost no register and no scoreboard entry, nobody ever waits for them, and they pull the stretch of the
// next row the warp is about to walk (the lanes' windows are shifted copies of each other: firsts and lasts cover it) into
// L1, so the gathers of a fresh row and of its leading edge are L1 hits instead of one L2 round trip per step.
// (`prefetch.global.L1` changed nothing on this part; plain loads into registers consumed a slot later stalled on the
// register moves the compiler put behind them -- profiles/r2c_*.)
__device__ __forceinline__ void stage_l1(const void* gptr)
{
    __shared__ unsigned scratch[16 * 32];
    const unsigned saddr = unsigned(__cvta_generic_to_shared(&scratch[threadIdx.x & 511u]));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void stage_l1_drain() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// What a warp stages of one record array for the slot after the current one.  FJ_STAGE_MODE 1 (default): every lane the
// first and the last record of its own window.  2: first / last are the WARP's bounds and lane l touches the l-th 128-byte
// line of the stretch between them (one copy per line instead of two per lane).  0: nothing.  Measured (profiles/
// r2k_sweep_variants.txt): 1 beats 0 by 7 % and 2 by 10 % on the whole step.
#ifndef FJ_STAGE_MODE
#define FJ_STAGE_MODE 1
#endif
template <class T>
__device__ __forceinline__ void stage_span(const T* base, unsigned first, unsigned last)
{
#if FJ_STAGE_MODE == 1
    stage_l1(base + first);
    stage_l1(base + last);
#elif FJ_STAGE_MODE == 3 /* the lanes' first records only (the trailing few records of the stretch are demand-loaded) */
    stage_l1(base + first);
#elif FJ_STAGE_MODE == 4 /* the lanes' last records only */
    stage_l1(base + last);
#elif FJ_STAGE_MODE == 2
    const unsigned long long a0 = (unsigned long long)(base + first) & ~127ull;
    const unsigned long long p = a0 + (threadIdx.x & 31u) * 128ull;
    if (p < (unsigned long long)(base + last + 1u))
        stage_l1((const void*)p);
#endif
}

// CLAMP: a lane with nobody to visit at a step of a slot it has a window in re-reads the LAST record of that window (a line
// its neighbouring lanes read at the same step) instead of its own record; the body must then mask the pair by `take`
// (own-index zeros no longer do it).
template <bool CLAMP = false, class Load, class Body, class StageFn>
__device__ __forceinline__ void for_neighbours2(const RunView& L, int W, bool active, unsigned self, Load&& load,
                                                Body&& body, StageFn&& stage)
{
    const int nrow = L.erows[W];
    const uint2* __restrict__ dp = L.erun + (size_t(W) * size_t(L.ecap)) * 32u + (threadIdx.x & 31u);
    int k = 0, T = 0, o = 0;
    uint2 d = make_uint2(0u, 0u), dn = make_uint2(0u, 0u), dnn = make_uint2(0u, 0u);
    for (int a = 2; a < 6 && a < nrow; ++a) stage_l1(dp + size_t(a) * 32u);
    if (nrow > 0)
        dn = ld_desc(dp);
    if (nrow > 1)
        dnn = ld_desc(dp + 32u);
    /* next slot with somebody in it: false at the end of the list (warp-uniform).  Descriptors run two slots ahead in
       registers (dnn was issued a slot ago: the move below does not wait) and six slots ahead in L1. */
    auto next_slot = [&]() -> bool {
        while (k < nrow)
        {
            d = dn;
            dn = dnn;
            if (!active)
                d.y = 0u;
            ++k;
            if (k + 5 < nrow)
                stage_l1(dp + size_t(k + 5) * 32u);
            if (k + 1 < nrow)
                dnn = ld_desc(dp + size_t(k + 1) * 32u);
            T = __reduce_max_sync(FJ_FULL, 32 - __clz(int(d.y)));
#if FJ_STAGE_MODE == 2
            if (k < nrow)
            {
                const bool any = active && dn.y != 0u;
                const unsigned lo = __reduce_min_sync(FJ_FULL, any ? dn.x : 0xFFFFFFFFu);
                const unsigned hi = __reduce_max_sync(FJ_FULL, any ? dn.x + unsigned(31 - __clz(int(dn.y))) : 0u);
                if (lo <= hi)
                    stage(lo, hi);
            }
#else
            if (k < nrow && active && dn.y != 0u) /* the records of the slot after this one */
                stage(dn.x, dn.x + unsigned(31 - __clz(int(dn.y))));
#endif
            if (T > 0)
                return true;
        }
        return false;
    };
    /* the lane's neighbour at the next step, or itself when it has none there; false past the end of the list */
    auto advance = [&](unsigned& j, bool& take) -> bool {
        if (++o >= T)
        {
            if (!next_slot())
            {
                j = self;
                take = false;
                return false;
            }
            o = 0;
        }
        take = ((d.y >> o) & 1u) != 0u;
        if (CLAMP)
            j = d.y != 0u ? d.x + min(unsigned(o), unsigned(31 - __clz(int(d.y)))) : self;
        else
            j = take ? d.x + unsigned(o) : self;
        return true;
    };
    o = -1;
    T = 0;
    for (;;)
    {
        unsigned ja, jb;
        bool ta, tb;
        if (!advance(ja, ta))
            break;
        const bool more = advance(jb, tb);
        const auto ra = load(ja);
        const auto rb = load(jb);
        body(ra, ta, rb, tb);
        if (!more)
            break;
    }
    stage_l1_drain();
}

// plain form for the wall treatments (few particles, short bodies): body(j)
template <class Body>
__device__ __forceinline__ void for_neighbours_simple(const RunView& L, int W, bool active, Body&& body)
{
    const int nrow = L.erows[W];
    const uint2* __restrict__ dp = L.erun + (size_t(W) * size_t(L.ecap)) * 32u + (threadIdx.x & 31u);
    if (!active)
        return;
    for (int k = 0; k < nrow; ++k)
    {
        const uint2 d = dp[size_t(k) * 32u];
        for (unsigned m = d.y, j = d.x; m; m >>= 1, ++j)
            if (m & 1u)
                body(j);
    }
}

// 1/x for positive normal x without the IEEE-division slow path: MUFU.RCP64H seed + two Newton steps
// (relative error ~1 ulp; the parity bar is 1e-10).
__device__ __forceinline__ double fj_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
// 1/sqrt(x) for positive normal x: MUFU.RSQ64H seed + two Newton steps
__device__ __forceinline__ double fj_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}

// 1/sqrt(x) for positive normal x with ONE third-order step on the MUFU.RSQ64H seed: y = y0 (1 + e/2 + 3 e^2 / 8),
// e = 1 - x y0^2 (error ~ seed^3: full double precision) -- four dependent FP64 operations instead of the six of two
// Newton steps, on the critical path of every pair
#ifndef FJ_RSQRT3_POLISH
#define FJ_RSQRT3_POLISH 0
#endif
__device__ __forceinline__ double fj_rsqrt3(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double t = x * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    double y = fma(ye, p, y0);
    if (FJ_RSQRT3_POLISH)
    {
        const double e2 = fma(-(x * y), y, 1.0);
        y = fma(0.5 * y, e2, y);
    }
    return y;
}

// Geometry of one pair, branch-free and with a short dependency chain: Rji (current positions), rr = the list's d^2
// (current separation, or the build-time one under FROZEN), 1 / r, t = 1 - r / 2H formed straight from rr and 1 / r, and
// the Wendland C2 gradient factor gk = 5 Wc / H^2 t^3, 0 when r / H < 1e-12 (Kernel.h:37-61).  W = t^4 (2 r / H + 1) Wc
// = t^4 (5 - 4 t) Wc.  A lane that has nobody to visit at a step is handed its own index: rr = 0, gk = 0 exactly, 1 / r
// and t not finite -- every use of those is a select on the `take` flag, never a product.
struct PairGeo
{
    double rx, ry, rz, rr, r2c, ir, t, gk;
};
// RAW: gk = t^3 without the constant 5 Wc / H^2 -- for a sweep whose pair sums are all linear in gk and which applies the
// constant once per particle after the walk (one FP64 multiplication less per pair).
template <bool FROZEN, bool RAW = false>
__device__ __forceinline__ PairGeo pair_geo(const DevConst& C, const double4& pi, const double4& x0i, const double4& pj,
                                            const double4& x0j)
{
    PairGeo g;
    g.rx = pj.x - pi.x;
    g.ry = pj.y - pi.y;
    g.rz = pj.z - pi.z;
    g.r2c = fma(g.rz, g.rz, fma(g.ry, g.ry, g.rx * g.rx));
    g.rr = g.r2c;
    if (FROZEN)
    {
        const double ex = x0j.x - x0i.x, ey = x0j.y - x0i.y, ez = x0j.z - x0i.z;
        g.rr = fma(ez, ez, fma(ey, ey, ex * ex));
    }
    g.ir = fj_rsqrt3(g.rr);
    g.t = fma(g.rr * C.mhalf_iH, g.ir, 1.0);
    g.gk = (g.rr < C.tiny2) ? 0.0 : (RAW ? (g.t * g.t) * g.t : (C.gk_fac * g.t) * (g.t * g.t));
    return g;
}
__device__ __forceinline__ double wend_W_t(const DevConst& C, double t)
{
    const double t2 = t * t;
    return (t2 * t2) * (fma(-4.0, t, 5.0) * C.W_correc);
}
// 1 / x with ONE Newton step on the MUFU.RCP64H seed (~1e-12 relative): for the 1 / (r^2 + eps h^2) of the viscous and
// density-diffusion terms, which are small corrections of the sums they enter
#ifndef FJ_RCP1_STEPS
#define FJ_RCP1_STEPS 1
#endif
__device__ __forceinline__ double fj_rcp1(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int k = 0; k < FJ_RCP1_STEPS; ++k)
    {
        const double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}

// Wendland C2 (Kernel.h:37-61).  t = 1 - q/2.  W = t^4 (2q+1) Wc ; GradK(R, r) = R * gk with
// gk = 5 Wc/H^2 * t^3, and 0 when r/H < 1e-12.
__device__ __forceinline__ double wend_t(const DevConst& C, double r) { return 1.0 - 0.5 * r * C.iH; }
__device__ __forceinline__ double wend_W(const DevConst& C, double r, double t)
{
    const double t2 = t * t;
    return (t2 * t2) * (2.0 * r * C.iH + 1.0) * C.W_correc;
}
__device__ __forceinline__ double wend_gk(const DevConst& C, double r, double t)
{
    return (r * C.iH < 1e-12) ? 0.0 : C.gk_fac * (t * t * t);
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FJ_FULL, v, o);
    return v;
}

// ================================================================= dSPH_PreStep
struct RecPre
{
    double4 p, x0;
    double rho;
    int b;
};

// npd_partial[W]: one partial per work warp (summed in a fixed order by fj_reduce_sum)
template <bool FROZEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, min_blocks(WARPS, true))
    k_prestep(Level S, RunView lv, RowMap M, DevConst C, double* __restrict__ npd_partial)
{
    int i, W;
    bool work;
    const bool active = fj_row_thread<WARPS>(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    double npd_ = 0.0;
    {
        const double4 pi = S.P0[i];
        const double4 x0i = FROZEN ? lv.x0[i] : pi;
        const double rho_i = S.P1[i].w;
        double l00 = 0, l01 = 0, l02 = 0, l11 = 0, l12 = 0, l22 = 0;
        double n00 = 0, n01 = 0, n02 = 0, n11 = 0, n12 = 0, n22 = 0;
        double g0 = 0, g1 = 0, g2 = 0, m0 = 0, m1 = 0, m2 = 0;
        /* Every sum below is linear in the kernel's constants: the pair loop takes gk = t^3 and W = t^4 (5 - 4 t), and
           5 Wc / H^2 and Wc are applied once after the walk (two FP64 multiplications less per pair; the kernel sum and the
           npd sum are the same sum, kept once). */
        double ksum = 0.0; /* sum_j W' over fluid neighbours: kernsum = Wc (1 + ksum), Shifting.cpp:39-45; npd = Wc ksum */
        double colour = 0.0;
        /* Lmat_nb and the prestep normal sum over FLUID neighbours only (b > PISTON).  Most warps have no other kind in
           reach, so the loop sums over ALL neighbours (l**, and m = -sum a) and a second set of accumulators takes the
           non-fluid ones under a warp-uniform vote -- skipped wherever there is no wall nearby; the fluid-only sums are the
           differences (n** = l** - w**, m -= mw) formed after the walk. */
        double w00 = 0, w01 = 0, w02 = 0, w11 = 0, w12 = 0, w22 = 0, mw0 = 0, mw1 = 0, mw2 = 0;
        struct PreOut
        {
            double ax, ay, az, rx, ry, rz;
        };
        auto pair = [&](const RecPre& q, const bool take) -> PreOut {
            const double4 pj = q.p;
            const PairGeo g = pair_geo<FROZEN, true>(C, pi, x0i, pj, q.x0);
            const double vg = pj.w * g.gk; /* V_j * t^3 ; Grad = GradK(-Rji) = -Rji*gk; 0 for a lane's own index */
            const double ax = vg * g.rx, ay = vg * g.ry, az = vg * g.rz;
            /* Lmat -= V Rji (x) Grad  ==  += V gk Rji (x) Rji */
            l00 = fma(ax, g.rx, l00);
            l01 = fma(ax, g.ry, l01);
            l02 = fma(ax, g.rz, l02);
            l11 = fma(ay, g.ry, l11);
            l12 = fma(ay, g.rz, l12);
            l22 = fma(az, g.rz, l22);
            const double dr = q.rho - rho_i;
            g0 = fma(-dr, ax, g0);
            g1 = fma(-dr, ay, g1);
            g2 = fma(-dr, az, g2);
            m0 -= ax; /* a lane's own index contributes exact zeros */
            m1 -= ay;
            m2 -= az;
            /* the kernel sums take fluid neighbours only (b > PISTON): a select, not a branch */
            const bool fl = take && q.b > FJSPH_PISTON;
            const double t2 = g.t * g.t;
            const double W_ = fl ? (t2 * t2) * fma(-4.0, g.t, 5.0) : 0.0;
            ksum += W_;
            colour = fma(pj.w, W_, colour);
            return PreOut{ax, ay, az, g.rx, g.ry, g.rz};
        };
        auto pair_wall = [&](const PreOut& o, const bool nf) {
            const double fx = nf ? o.ax : 0.0, fy = nf ? o.ay : 0.0, fz = nf ? o.az : 0.0;
            w00 = fma(fx, o.rx, w00);
            w01 = fma(fx, o.ry, w01);
            w02 = fma(fx, o.rz, w02);
            w11 = fma(fy, o.ry, w11);
            w12 = fma(fy, o.rz, w12);
            w22 = fma(fz, o.rz, w22);
            mw0 -= fx;
            mw1 -= fy;
            mw2 -= fz;
        };
        for_neighbours2(
            lv, W, active, unsigned(i),
            [&](const unsigned j) {
                RecPre q;
                q.p = gather(S.P0, j);
                q.rho = __ldg(&S.P1[j].w);
                q.b = __ldg(&S.b[j]);
                if (FROZEN)
                    q.x0 = gather(lv.x0, j);
                return q;
            },
            [&](const RecPre& qa, const bool ta, const RecPre& qb, const bool tb) {
                const PreOut oa = pair(qa, ta);
                const PreOut ob = pair(qb, tb);
                const bool nfa = ta && !(qa.b > FJSPH_PISTON), nfb = tb && !(qb.b > FJSPH_PISTON);
                if (__any_sync(FJ_FULL, nfa || nfb))
                {
                    pair_wall(oa, nfa);
                    pair_wall(ob, nfb);
                }
            },
            [&](const unsigned first, const unsigned last) {
                stage_span(S.P0, first, last);
                stage_span(S.P1, first, last);
                stage_span(S.b, first, last);
                if (FROZEN)
                {
                    stage_span(lv.x0, first, last);
                }
            });
        npd_ = C.W_correc * ksum;
        if (active)
        {
        n00 = l00 - w00;
        n01 = l01 - w01;
        n02 = l02 - w02;
        n11 = l11 - w11;
        n12 = l12 - w12;
        n22 = l22 - w22;
        m0 -= mw0;
        m1 -= mw1;
        m2 -= mw2;
        const double kernsum = fma(C.W_correc, ksum, C.W_correc);
        colour *= C.W_correc;
        l00 *= C.gk_fac;
        l01 *= C.gk_fac;
        l02 *= C.gk_fac;
        l11 *= C.gk_fac;
        l12 *= C.gk_fac;
        l22 *= C.gk_fac;
        n00 *= C.gk_fac;
        n01 *= C.gk_fac;
        n02 *= C.gk_fac;
        n11 *= C.gk_fac;
        n12 *= C.gk_fac;
        n22 *= C.gk_fac;
        g0 *= C.gk_fac;
        g1 *= C.gk_fac;
        g2 *= C.gk_fac;
        m0 *= C.gk_fac;
        m1 *= C.gk_fac;
        m2 *= C.gk_fac;
        double Lm[3][3] = {{l00, l01, l02}, {l01, l11, l12}, {l02, l12, l22}};
        double Li[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        if (C.dim == 2)
        {
            /* SIMDIM = 2: L is 2 x 2 (every z term above is an exact zero); the third row and column stay the identity's */
            const double L2[2][2] = {{l00, l01}, {l01, l11}};
            double t2[2][2];
            if (fj_qr_inverse2(L2, t2))
            {
                Li[0][0] = t2[0][0];
                Li[0][1] = t2[0][1];
                Li[1][0] = t2[1][0];
                Li[1][1] = t2[1][1];
            }
        }
        else
        {
            double tmp[3][3];
            if (fj_qr_inverse3(Lm, tmp))
            {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Li[a][c] = tmp[a][c];
            }
        }
        const double gr0 = Li[0][0] * g0 + Li[0][1] * g1 + Li[0][2] * g2;
        const double gr1 = Li[1][0] * g0 + Li[1][1] * g1 + Li[1][2] * g2;
        const double gr2 = Li[2][0] * g0 + Li[2][1] * g1 + Li[2][2] * g2;
        double lam = 1.0, lam_nb = 1.0;
        if (S.b[i] != FJSPH_BOUND)
        {
            lam = (C.dim == 2) ? fj_min_eig2(l00, l01, l11) : fj_min_eig3(l00, l01, l11, l02, l12, l22);
            lam_nb = (C.dim == 2) ? fj_min_eig2(n00, n01, n11) : fj_min_eig3(n00, n01, n11, n02, n12, n22);
        }
        S.L0[i] = Li[0][0];
        S.L1[i] = Li[0][1];
        S.L2[i] = Li[0][2];
        S.L3[i] = Li[1][0];
        S.L4[i] = Li[1][1];
        S.L5[i] = Li[1][2];
        S.L6[i] = Li[2][0];
        S.L7[i] = Li[2][1];
        S.L8[i] = Li[2][2];
        S.P3[i] = make_double4(gr0, gr1, gr2, lam);
        S.NP[i] = make_double4(Li[0][0] * m0 + Li[0][1] * m1 + Li[0][2] * m2,
                               Li[1][0] * m0 + Li[1][1] * m1 + Li[1][2] * m2,
                               Li[2][0] * m0 + Li[2][1] * m1 + Li[2][2] * m2, lam_nb);
        const double colourG = (lam > 0.7) ? 2.0 : 2.0 * fmax(1.0, 1.0 / (2.0 * colour));
        double4 sc = S.SC[i];
        sc.x = colourG;
        sc.y = colour;
        sc.z = kernsum;
        S.SC[i] = sc;
        }
    }
    const double tot = warp_sum(npd_);
    if ((threadIdx.x & 31u) == 0u)
        npd_partial[W] = tot;
}

// ================================================================= get_aero_velocity (constVel)
__global__ void k_aero_velocity(Level S, const int* __restrict__ ncount, const int* __restrict__ blk,
                                int n_bound_blocks, DevConst C, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] < n_bound_blocks)
        return;
    bool cond;
    if (C.use_lam)
        cond = S.NP[i].w < C.lam_cutoff;
    else
        cond = double(ncount[i] + 1) * C.i_n_full < C.lam_cutoff;
    double4 cv = S.CV[i];
    if (cond && S.b[i] == FJSPH_FREE)
    {
        cv.x = C.vinf_x;
        cv.y = C.vinf_y;
        cv.z = C.vinf_z;
        S.cellID[i] = 1;
    }
    else
    {
        cv.x = cv.y = cv.z = 0.0;
        S.cellID[i] = -3;
    }
    S.CV[i] = cv;
}

// ================================================================= Detect_Surface loop 1 + dissipation_terms
struct RecS1
{
    double4 p, g, v, x0;
    int b;
};

template <bool SURF, bool DISS, bool FROZEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, min_blocks(WARPS, false))
    k_surf1_diss(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int n_bound_blocks, DevConst C,
                 int* __restrict__ near_warps = nullptr)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread<WARPS>(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    const bool active = in_row && blk[i] >= n_bound_blocks;
    if (!__any_sync(FJ_FULL, active))
        return;
    const double4 x0i = FROZEN ? lv.x0[i] : make_double4(0, 0, 0, 0);
    const double4 pi = S.P0[i];
    const double4 vi = S.P1[i];
    const double4 gi = S.P3[i]; /* gradRho_i, lam_i */
    const double4 npi = S.NP[i]; /* prestep normal, lam_nb */
    const int b_i = S.b[i];
    const double rho_i = vi.w, lam_i = gi.w, lam_nb = npi.w;
    const double h = 1.33 * C.particle_step;
    const double sqrt2h = sqrt(2.0) * h;

    // --- surface flag set-up (Geometry.cpp:31-103)
    int surf = 0;
    bool need_test = false;
    double nhx = 0, nhy = 0, nhz = 0, Tx = 0, Ty = 0, Tz = 0;
    if (SURF)
    {
        if (b_i < FJSPH_PIPE)
            surf = 0;
        else if (lam_nb < 0.2)
            surf = 1;
        else if (lam_nb < 0.75)
        {
            surf = 1;
            need_test = true;
            const double nn = npi.x * npi.x + npi.y * npi.y + npi.z * npi.z;
            const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 1.0;
            nhx = npi.x * inv;
            nhy = npi.y * inv;
            nhz = npi.z * inv;
            Tx = pi.x + h * nhx;
            Ty = pi.y + h * nhy;
            Tz = pi.z + h * nhz;
        }
    }
    const bool hi_lam = lam_i > 0.7;
    const double lam_ref = hi_lam ? lam_i : 0.0;
    /* cbar = (sqrt(B gam / rho_i) + sqrt(B gam / rho_j)) / 2, Kernel.h:217-244 */
    /* the pair loop works in units of the sums' constants (gk = t^3, cbar in units of sqrt(B gam) / 2, ...): they are applied
       once after the walk -- six FP64 instructions less per pair */
    const double rs_i = DISS ? 1.0 / sqrt(rho_i) : 0.0; /* c_i / sqrt(B gam) */
    const double eps_d = C.eps_d; /* 0.0001 H^2 -- Q3: 0.0001 here, 0.001 in the force loop */
    double nx = 0, ny = 0, nz = 0;
    double avx = 0, avy = 0, avz = 0, Rrhod = 0;
    /* particle_shift limits the shifting velocity by max_j |v_j - v_i| / 2 (Shifting.cpp:253-256), the only thing the fused
       surface + shifting sweep would gather the velocity record for.  This sweep holds v_j - v_i already (dissipation), so
       when both halves run (the fused pass) it takes the maximum and leaves it in AV.w -- the curvature's slot, which the
       sweep that follows reads first and then overwrites with the curvature. */
    constexpr bool STASH = SURF && DISS;
    double maxU2 = 0.0;
    const double cos_pi4 = 0.70710678118654757;

    /* main part of one pair, branch-free */
    auto pair = [&](const RecS1& q, const bool take, PairGeo& g) {
        const double4 pj = q.p;
        const double4 gj = q.g;
        g = pair_geo<FROZEN, true>(C, pi, x0i, pj, q.x0);
        const double vg = pj.w * g.gk; /* V_j t^3; 0 for a lane's own index */
        if (SURF)
        {
            /* normal from the eigenvalue gradient, GradK(Rij = xi - xj) = -Rji*gk (Geometry.cpp:109-137) */
            const double sn = -vg * (gj.w - lam_ref);
            nx = fma(sn, g.rx, nx);
            ny = fma(sn, g.ry, ny);
            nz = fma(sn, g.rz, nz);
        }
        if (DISS)
        {
            const double4 vj = q.v;
            const double rho_j = vj.w;
            const double idist2 = fj_rcp1(g.rr + eps_d);
            const double w = vg * (g.r2c * idist2); /* V_j (Rji . gradK) idist2 */
            const double drho = rho_j - rho_i;
            const bool fl = q.b > FJSPH_PISTON;
            const double ux = vj.x - vi.x, uy = vj.y - vi.y, uz = vj.z - vi.z;
            const double vdotr = ux * g.rx + uy * g.ry + uz * g.rz;
            if (STASH)
                maxU2 = fmax(maxU2, fma(uz, uz, fma(uy, uy, ux * ux)));
            /* ArtVisc = 0 when Vji.Rji > 0, and for a non-fluid neighbour (selects, no branches).
               m_j gk alpha cbar mu / rhobar with m_j = rho_j V_j, mu = H Vji.Rji idist2, cbar = sqrt(B gam) (rho_i^-1/2 +
               rho_j^-1/2) / 2, rhobar = (rho_i + rho_j) / 2: the constants 5 Wc / H^2 * alpha * sqrt(B gam) * H after the walk */
            double f = (rho_j * vg) * (rs_i + fj_rsqrt3(rho_j)) * (vdotr * idist2) * fj_rcp1(rho_i + rho_j);
            f = (vdotr > 0.0 || !fl) ? 0.0 : f;
            avx = fma(f, g.rx, avx);
            avy = fma(f, g.ry, avy);
            avz = fma(f, g.rz, avz);
            const double gdot = (gi.x + gj.x) * g.rx + (gi.y + gj.y) * g.ry + (gi.z + gj.z) * g.rz;
            Rrhod = fma(fma(0.5, fl ? gdot : 0.0, drho), w, Rrhod);
        }
    };
    /* Detect_Surface's cone test (Geometry.cpp:60-103): only particles with 0.2 <= lam_nb < 0.75 */
    auto cone = [&](const RecS1& q, const bool take, const PairGeo& g) {
        if (!take)
            return;
        const double4 pj = q.p;
        const double r = g.rr * g.ir;
        if (r >= sqrt2h)
        {
            const double ex = pj.x - Tx, ey = pj.y - Ty, ez = pj.z - Tz;
            if (sqrt(ex * ex + ey * ey + ez * ez) < h)
                surf = 0;
        }
        else if (C.dim == 2)
        {
            /* SIMDIM = 2, Geometry.cpp:76-84: |nhat . x_jT| + |tau . x_jT| < h with tau = (n_y, -n_x) normalised */
            const double ex = pj.x - Tx, ey = pj.y - Ty;
            if (fabs(nhx * ex + nhy * ey) + fabs(nhy * ex - nhx * ey) < h)
                surf = 0;
        }
        else
        {
            /* acos(nhat . (Rji/r)) < pi/4 ; acos is NaN outside [-1,1] so those never trigger */
            const double c = (nhx * g.rx + nhy * g.ry + nhz * g.rz) * g.ir;
            if (c > cos_pi4 && c <= 1.0)
                surf = 0;
        }
    };
    for_neighbours2(
        lv, W, active, unsigned(i),
        [&](const unsigned j) {
            RecS1 q;
            q.p = gather(S.P0, j);
            q.g = gather(S.P3, j);
            if (DISS)
            {
                q.v = gather(S.P1, j);
                q.b = __ldg(&S.b[j]);
            }
            if (FROZEN)
                q.x0 = gather(lv.x0, j);
            return q;
        },
        [&](const RecS1& qa, const bool ta, const RecS1& qb, const bool tb) {
            PairGeo ga, gb;
            pair(qa, ta, ga);
            pair(qb, tb, gb);
            if (SURF && need_test)
            {
                cone(qa, ta, ga);
                cone(qb, tb, gb);
            }
        },
        [&](const unsigned first, const unsigned last) {
            stage_span(S.P0, first, last);
            stage_span(S.P3, first, last);
            if (DISS)
            {
                stage_span(S.P1, first, last);
                stage_span(S.b, first, last);
            }
            if (FROZEN)
            {
                stage_span(lv.x0, first, last);
            }
        });
    if (SURF)
    {
        nx *= C.gk_fac;
        ny *= C.gk_fac;
        nz *= C.gk_fac;
        const double tx = S.L0[i] * nx + S.L1[i] * ny + S.L2[i] * nz;
        const double ty = S.L3[i] * nx + S.L4[i] * ny + S.L5[i] * nz;
        const double tz = S.L6[i] * nx + S.L7[i] * ny + S.L8[i] * nz;
        const double nn = sqrt(tx * tx + ty * ty + tz * tz);
        double4 out = make_double4(0.0, 0.0, 0.0, double(surf));
        if (nn > 0.1 * lam_i / C.H)
        {
            const double inv = 1.0 / nn;
            out.x = tx * inv;
            out.y = ty * inv;
            out.z = tz * inv;
        }
        if (active)
        {
            S.P4[i] = out;
            S.surf_i[i] = surf;
        }
        if (near_warps)
        {
            /* how many warps hold a particle the lean surface / shifting sweep cannot take (k_surf23_shift, CLASS) */
            const bool near = active && ((out.x != 0.0 || out.y != 0.0 || out.z != 0.0) ||
                                         ((b_i == FJSPH_FREE) && (C.acase == 1) && (lam_nb < C.lam_cutoff)));
            const unsigned m = __ballot_sync(FJ_FULL, near);
            if (m && (threadIdx.x & 31u) == 0u)
                atomicAdd(near_warps, 1);
        }
        if (active && b_i < FJSPH_PIPE)
        {
            double4 th = S.TH[i];
            th.z = 1.0; /* woccl = 1, Geometry.cpp:33-37 */
            S.TH[i] = th;
        }
    }
    if (DISS && active)
    {
        const double k_av = C.gk_fac * C.visc_alpha * sqrt(C.Bgam) * C.H;
        double4 av = S.AV[i];
        av.x = k_av * avx;
        av.y = k_av * avy;
        av.z = k_av * avz;
        if (STASH)
            av.w = maxU2;
        S.AV[i] = av;
        double4 af = S.AF[i];
        af.w = (C.dsph_cont * C.gk_fac) * Rrhod;
        S.AF[i] = af;
    }
}

// ================================================================= Detect_Surface loops 2 and 3 + particle_shift
// One sweep: everything particle_shift reads from neighbour j (positions, velocities, the Detect_Surface
// normals) is final after loop 1, and what it reads from i itself (surfzone, the kept normal) is produced by
// this very loop, so loops 2+3 of Detect_Surface (Geometry.cpp:145-277) and particle_shift
// (Shifting.cpp:189-290) share one pass over the list.  The stage entry points run the halves separately.
struct RecS2
{
    double4 n, p, v, x0;
    int b, s; /* s: the surf flag as the lean class reads it (an int: no conversion, no FP64 compare per pair) */
};

#ifndef FJ_FUSE_SHIFT
#define FJ_FUSE_SHIFT 1
#endif
// CLASS splits the fused sweep by what a particle needs (0: everybody, the stage entry points):
//   1 "lean": particles WITHOUT a loop-1 normal and without an occlusion to compute -- the bulk of a fluid.  Their
//      curvature is 0, n_i . n_j is 0 for every j, and of neighbour j's P4 record only the surf flag matters (one
//      8-byte load instead of the 32-byte gather); no L matrix, no occlusion, far fewer registers, more resident warps;
//   2 the others (near a free surface), with the full body.
// The two launches cover disjoint particles; a warp that holds both kinds runs in both.
template <bool SURF23, bool SHIFT, int CLASS, bool FROZEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, min_blocks(WARPS, CLASS == 1))
    k_surf23_shift(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int n_bound_blocks, DevConst C)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread<WARPS>(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    bool active = in_row && blk[i] >= n_bound_blocks;
    const double4 pi = S.P0[i];
    const double4 x0i = FROZEN ? lv.x0[i] : make_double4(0, 0, 0, 0);
    const double4 vi = S.P1[i];
    const double4 ni = S.P4[i]; /* Detect_Surface loop-1 normal (unit or zero), surf */
    const int b_i = S.b[i];
    const double lam_nb = S.NP[i].w;
    const bool ni_nz_real = (ni.x * ni.x + ni.y * ni.y + ni.z * ni.z) > 0.0;
    if (CLASS != 0)
    {
        const bool near = ni_nz_real || ((b_i == FJSPH_FREE) && (C.acase == 1) && (lam_nb < C.lam_cutoff));
        if (near != (CLASS == 2))
            active = false;
    }
    if (!__any_sync(FJ_FULL, active))
        return;
    const bool ni_nz = (CLASS == 1) ? false : ni_nz_real;
    bool has_fluid = false; /* lean: min_j n_i . n_j over fluid neighbours is 0 if there is one */

    // ---- Detect_Surface loop 2 set-up
    const bool occl = (b_i == FJSPH_FREE) && (C.acase == 1);
    double vdx = 0, vdy = 0, vdz = 0;
    if (SURF23 && S.cellID[i] != -1 && occl)
    {
        if (C.asource == 1)
        {
            const double4 cv = S.CV[i];
            vdx = cv.x - vi.x;
            vdy = cv.y - vi.y;
            vdz = cv.z - vi.z;
        }
        else
        {
            vdx = C.vinf_x - vi.x;
            vdy = C.vinf_y - vi.y;
            vdz = C.vinf_z - vi.z;
        }
    }
    /* -1/|Vdiff|: inf for Vdiff = 0, so frac = 0 * inf = NaN never passes the '>' (Geometry.cpp:237-246) */
    const double mivdn = -1.0 / sqrt(vdx * vdx + vdy * vdy + vdz * vdz);
    double L0 = 0, L1 = 0, L2 = 0, L3 = 0, L4 = 0, L5 = 0, L6 = 0, L7 = 0, L8 = 0;
    if (SURF23 && ni_nz)
    {
        L0 = S.L0[i];
        L1 = S.L1[i];
        L2 = S.L2[i];
        L3 = S.L3[i];
        L4 = S.L4[i];
        L5 = S.L5[i];
        L6 = S.L6[i];
        L7 = S.L7[i];
        L8 = S.L8[i];
    }
    double curve = 0.0, woccl_ = 0.0;
    int zone = (ni.w != 0.0) ? 1 : 0; /* the list of the reference includes self */
    /* woccl is overwritten with 1 when lam_nb >= lam_cutoff, so the occlusion max is only needed below it */
    const bool need_occl = (CLASS != 1) && SURF23 && occl && (lam_nb < C.lam_cutoff);

    // ---- particle_shift set-up (Shifting.cpp:205-212)
    const bool do_shift = SHIFT && !(lam_nb < 0.55 || b_i == FJSPH_BUFFER);
    /* with the stage entry point surfzone is already known; fused, it is this loop's `zone` */
    const bool known_bulk = SHIFT && !SURF23 && (S.surfzone[i] == 0) && (lam_nb > 0.55);
    /* fused pass: max_j |v_j - v_i|^2 was left in AV.w by the sweep before (k_surf1_diss), so the velocity record of j is
       not gathered here at all */
    constexpr bool STASHED = SURF23 && SHIFT;
    double dux = 0, duy = 0, duz = 0, maxU2 = (STASHED && do_shift) ? S.AV[i].w : 0.0;
    const double kq_fac = C.W_correc * C.iW_dx;
    /* max_j acos(c_j) over c_j in [-1,1] == acos(min_j c_j); NaNs (|c|>1) are skipped by the reference's '>' */
    double min_c = 2.0;
    const bool need_pos = (SURF23 && (ni_nz || need_occl)) || do_shift;

    /* one pair, branch-free: what a particle does not need is computed and dropped by a select */
    auto pair = [&](const RecS2& q, const bool take) {
        const double4 nj = q.n;
        if (CLASS == 1)
            zone |= (take && q.s != 0) ? 1 : 0;
        else
            zone |= (take && nj.w != 0.0) ? 1 : 0;
        const double4 pj = q.p;
        const PairGeo g = pair_geo<FROZEN, true>(C, pi, x0i, pj, q.x0); /* gk = t^3: 5 Wc / H^2 once after the walk */
        if (SURF23)
        {
            if (CLASS != 1)
            {
                /* curvature (Geometry.cpp:187-213): L is zero unless particle i has a loop-1 normal */
                const double vg = pj.w * g.gk;
                const double dx_ = nj.x - ni.x, dy_ = nj.y - ni.y, dz_ = nj.z - ni.z;
                const double lx = L0 * dx_ + L1 * dy_ + L2 * dz_;
                const double ly = L3 * dx_ + L4 * dy_ + L5 * dz_;
                const double lz = L6 * dx_ + L7 * dy_ + L8 * dz_;
                const double term = vg * (lx * g.rx + ly * g.ry + lz * g.rz);
                curve += (ni_nz && (nj.x * nj.x + nj.y * nj.y + nj.z * nj.z) > 0.0) ? term : 0.0;
                const double frac = (g.rx * vdx + g.ry * vdy + g.rz * vdz) * mivdn * g.ir;
                woccl_ = (need_occl && take && frac > woccl_) ? frac : woccl_;
            }
        }
        if (SHIFT)
        {
            const double4 vj = q.v;
            const double t2 = g.t * g.t;
            const double kq = ((t2 * t2) * fma(-4.0, g.t, 5.0)) * kq_fac; /* W(r) / W(dx) */
            const double kq2 = kq * kq;
            const double f = take ? fma(0.2, kq2 * kq2, 1.0) * g.gk * pj.w : 0.0;
            dux = fma(f, g.rx, dux);
            duy = fma(f, g.ry, duy);
            duz = fma(f, g.rz, duz);
            const bool fl = take && q.b > FJSPH_PISTON;
            if (CLASS == 1)
                has_fluid = has_fluid || fl;
            else
            {
                /* n_i and n_j are unit vectors or zero already (loop 1), normalized() is the identity */
                const double c = ni.x * nj.x + ni.y * nj.y + ni.z * nj.z;
                min_c = (fl && !known_bulk && c >= -1.0 && c <= 1.0) ? fmin(min_c, c) : min_c;
            }
            if (!STASHED)
            {
                const double ux = vj.x - vi.x, uy = vj.y - vi.y, uz = vj.z - vi.z;
                maxU2 = fmax(maxU2, fma(uz, uz, fma(uy, uy, ux * ux)));
            }
        }
    };
    for_neighbours2(
        lv, W, active, unsigned(i),
        [&](const unsigned j) {
            RecS2 q;
            if (CLASS == 1)
                q.s = __ldg(&S.surf_i[j]); /* the surf flag is all a lean particle reads of n_j */
            else
                q.n = gather(S.P4, j);
            q.p = gather(S.P0, j);
            if (SHIFT)
            {
                if (!STASHED)
                    q.v = gather(S.P1, j);
                q.b = __ldg(&S.b[j]);
            }
            if (FROZEN)
                q.x0 = gather(lv.x0, j);
            return q;
        },
        [&](const RecS2& qa, const bool ta, const RecS2& qb, const bool tb) {
            pair(qa, ta);
            pair(qb, tb);
        },
        [&](const unsigned first, const unsigned last) {
            if (CLASS == 1)
            {
                stage_span(S.surf_i, first, last);
            }
            else
            {
                stage_span(S.P4, first, last);
            }
            stage_span(S.P0, first, last);
            if (SHIFT)
            {
                if (!STASHED)
                    stage_span(S.P1, first, last);
                stage_span(S.b, first, last);
            }
            if (FROZEN)
            {
                stage_span(lv.x0, first, last);
            }
        });
    if (CLASS == 1 && has_fluid)
        min_c = 0.0;
    if (!active)
        return; /* the walk is over: nothing below votes */

    if (SURF23)
    {
        double4 th = S.TH[i];
        th.z = (lam_nb < C.lam_cutoff) ? fmax(0.0, fmin(woccl_, 1.0)) : 1.0;
        S.TH[i] = th;
        double4 np = S.NP[i];
        np.x = ni.x;
        np.y = ni.y;
        np.z = ni.z;
        S.NP[i] = np; /* pi.norm = norms[ii] */
        double4 av = S.AV[i];
        av.w = C.gk_fac * curve;
        S.AV[i] = av;
        double4 sc = S.SC[i];
        sc.w = S.P3[i].w; /* pDist = lam */
        S.SC[i] = sc;
        if (C.ale)
            S.surfzone[i] = zone;
    }
    if (SHIFT)
    {
        double4 out = S.P2[i];
        if (!do_shift)
        {
            out.x = out.y = out.z = 0.0;
            S.P2[i] = out;
            return;
        }
        const int surfzone = SURF23 ? zone : S.surfzone[i];
        const bool bulk = (surfzone == 0) && (lam_nb > 0.55);
        const double vnorm = sqrt(vi.x * vi.x + vi.y * vi.y + vi.z * vi.z);
        const double sc = (-2.0 * C.H * C.gk_fac) * vnorm;
        dux *= sc;
        duy *= sc;
        duz *= sc;
        const double dn = sqrt(dux * dux + duy * duy + duz * duz);
        const double lim = fmin(dn, fmin(sqrt(maxU2) / 2.0, C.max_shift_vel));
        if (dn > 0.0)
        {
            const double s2 = lim / dn;
            dux *= s2;
            duy *= s2;
            duz *= s2;
        }
        else
        {
            dux *= lim;
            duy *= lim;
            duz *= lim;
        }
        if (bulk)
        {
            out.x = dux;
            out.y = duy;
            out.z = duz;
        }
        else
        {
            const double woccl = (min_c <= 1.0) ? acos(min_c) : 0.0;
            if (woccl < FJ_PI / 12.0)
            {
                /* (I - n n^T) deltaU with n = -nhat */
                const double nd = ni.x * dux + ni.y * duy + ni.z * duz;
                out.x = dux - ni.x * nd;
                out.y = duy - ni.y * nd;
                out.z = duz - ni.z * nd;
            }
            else
            {
                out.x = out.y = out.z = 0.0;
            }
        }
        S.P2[i] = out;
    }
}

// ================================================================= get_acc_and_Rrho
__device__ __forceinline__ double get_cd(double Re)
{
    return (1.0 + 0.197 * pow(Re, 0.63) + 2.6e-04 * pow(Re, 1.38)) * (24.0 / (Re + 0.00001));
}

// CalcAeroAcc, Aero.h:204-263: Gissler (Aero.h:37-98), induced pressure (Aero.h:106-202), skin friction (Aero.h:224-257).
// Vd = cellV - v; np = {surface normal, lam_nb}; th = {p, m, woccl, cellRho}; norm_curve = dx * curvature.
__device__ void calc_aero_acc(const DevConst& C, double dx_, double dy_, double dz_, const double4& np, const double4& th,
                              double cellP, double norm_curve, double nneigh, double (&out)[3])
{
    const double vd2 = dx_ * dx_ + dy_ * dy_ + dz_ * dz_;
    const double vd = sqrt(vd2);
    const double lam_nb = np.w, mass = th.y, cellRho = th.w;
    out[0] = out[1] = out[2] = 0.0;
    if (C.acase == 1)
    {
        const double Re = 2.0 * cellRho * vd * C.aero_L / C.mu_g;
        double frac2;
        if (C.use_lam)
            frac2 = fmin(C.interp_fac * lam_nb, 1.0);
        else
            frac2 = fmin(C.interp_fac * nneigh * C.i_n_full, 1.0);
        const double frac1 = 1.0 - frac2;
        const double Cds = get_cd(Re);
        double Cdl, Adrop;
        if (C.use_TAB_def)
        {
            double ymax = vd2 * C.ycoef;
            if (ymax > 1.0)
                ymax = 1.0;
            Cdl = Cds * (1 + 2.632 * ymax);
            const double rr_ = C.aero_L + C.tab_Cb * C.aero_L * ymax;
            /* Aero.h:57-70: a disc in 3D, a chord in 2D */
            Adrop = (C.dim == 2) ? C.A_sphere + 2.0 * (C.tab_Cb * C.aero_L * ymax) : FJ_PI * rr_ * rr_;
        }
        else
        {
            Cdl = Cds;
            Adrop = C.A_sphere;
        }
        const double Cdi = frac1 * Cdl + frac2;
        const double Ai = (1.0 - th.z) * (frac1 * Adrop + frac2 * C.A_plate);
        const double f = 0.5 * vd / C.sos2 * C.gamma_g * cellP * Cdi * Ai / mass;
        out[0] = f * dx_;
        out[1] = f * dy_;
        out[2] = f * dz_;
        return;
    }
    /* unit normal and unit Vdiff; Eigen's normalized() leaves the zero vector unchanged */
    double nx = np.x, ny = np.y, nz = np.z;
    {
        const double nn = nx * nx + ny * ny + nz * nz;
        if (nn > 0.0)
        {
            const double inv = 1.0 / sqrt(nn);
            nx *= inv;
            ny *= inv;
            nz *= inv;
        }
    }
    const double Vnorm = dx_ * nx + dy_ * ny + dz_ * nz;
    if (C.acase == 2)
    {
        const double ivd = vd2 > 0.0 ? 1.0 / vd : 1.0;
        const double theta = fabs(acos(-(nx * dx_ + ny * dy_ + nz * dz_) * ivd));
        double Cp_s, Cp_p, Cp_b, Cp_tot;
        if (theta < 2.4455)
        {
            const double st = sin(theta);
            Cp_s = 1.0 - 2.25 * (st * st);
        }
        else
            Cp_s = 0.075;
        if (theta < 1.570797)
            Cp_p = cos(theta);
        else if (theta < 1.9918)
            Cp_p = -pow(cos(6.0 * theta + 0.5 * FJ_PI), 1.5);
        else if (theta < 2.0838)
            Cp_p = 5.5836 * theta - 11.5601;
        else
            Cp_p = 0.075;
        if (theta < 0.7854)
            Cp_b = 1.0;
        else if (theta < 1.570797)
            Cp_b = 0.5 * (cos(4.0 * theta - FJ_PI) + 1.0);
        else
            Cp_b = 0.0;
        const double fac1 = 0.25, ifac1 = 4.0;
        if (norm_curve < -fac1)
            Cp_tot = Cp_b;
        else if (norm_curve < 0.0)
        {
            const double frac = (norm_curve + fac1) * ifac1;
            Cp_tot = frac * Cp_b + (1.0 - frac) * Cp_p;
        }
        else if (norm_curve < fac1)
        {
            const double frac = norm_curve * ifac1;
            Cp_tot = frac * Cp_p + (1.0 - frac) * Cp_s;
        }
        else
            Cp_tot = Cp_s;
        const double q = C.gamma_g * cellP / C.sos2; /* compressible dynamic pressure factor */
        const double Plocali = 0.5 * vd2 * q * Cp_tot;
        const double Cdi = get_cd(cellRho * vd * C.aero_L / C.mu_g);
        const double fd = 0.5 * vd * q * (FJ_PI * C.aero_L * C.aero_L * 0.25) * Cdi / mass;
        const double fk = -Plocali * C.A_plate / mass;
        const double px = dx_ - Vnorm * nx, py = dy_ - Vnorm * ny, pz = dz_ - Vnorm * nz;
        const double vpar = sqrt(px * px + py * py + pz * pz);
        const double Cf = 0.027 / pow(cellRho * vpar * C.aero_L / C.mu_g + 1e-6, 1.0 / 7.0);
        const double fs = 0.5 * vpar * q * Cf * C.A_plate / mass;
        double frac1;
        if (C.use_lam)
            frac1 = fmin(C.interp_fac * lam_nb, 1.0);
        else
            frac1 = fmin(C.interp_fac * nneigh * C.i_n_full, 1.0);
        out[0] = frac1 * (fk * nx + fs * px) + (1.0 - frac1) * (fd * dx_);
        out[1] = frac1 * (fk * ny + fs * py) + (1.0 - frac1) * (fd * dy_);
        out[2] = frac1 * (fk * nz + fs * pz) + (1.0 - frac1) * (fd * dz_);
        return;
    }
    if (C.acase == 3 && Vnorm > 0.001)
    {
        const double Re = C.rho_g * vd * C.aero_L / C.mu_g;
        const double fp = 0.5 * C.rho_g * Vnorm * Vnorm * C.A_plate / mass;
        const double an = fabs(Vnorm);
        const double px = dx_ - an * nx, py = dy_ - an * ny, pz = dz_ - an * nz;
        const double vpar = sqrt(px * px + py * py + pz * pz);
        const double Cf = 0.027 / pow(Re, 1.0 / 7.0);
        const double fs = 0.5 * C.rho_g * vpar * Cf * C.A_plate / mass;
        const double frac2 = fmin(1.5 * lam_nb, 1.0), frac1 = 1.0 - frac2;
        const double fd = 0.5 * C.rho_g * vd * (FJ_PI * C.aero_L * C.aero_L / 4) * get_cd(Re) / mass;
        out[0] = frac2 * (fp * nx + fs * px) + frac1 * (fd * dx_);
        out[1] = frac2 * (fp * ny + fs * py) + frac1 * (fd * dy_);
        out[2] = frac2 * (fp * nz + fs * pz) + frac1 * (fd * dz_);
    }
}

struct RecF
{
    double4 p, v, q, x0;
    int b;
};

// Pair algebra (V_j = m_j / rho_j, G = V_j gradK, Pi = p_i / rho_i^2, u = v_j - v_i, w = vPert):
//   pressure    BasePos, Kernel.h:153-157          acc_     -= m_j (Pi + Pj) gradK         = rho_j (Pi + Pj) G
//   viscosity   Kernel.h:262-269                   visc_    += m_j nu (rho_i + rho_j)/(rho_i rho_j) (Rji.gradK) idist2 u
//                                                            = (nu + nu rho_j / rho_i) (V_j gk |Rji|^2 idist2) u
//   ALE moment. Kernel.h:179-185                   acc_ale_ += V_j [(v_j (w_j.gK) + v_i (w_i.gK)) - v_i ((w_j - w_i).gK)]
//                                                            = u (w_j.G) + 2 v_i (w_i.G)
//   continuity  Kernel.h:187-196                   Rrho_    -= V_j ((u + w_j - w_i).gK)     = -(u.G + w_j.G) + w_i.G
//                                                  Rrhoc_   += V_j (rho_j w_j.gK + rho_i w_i.gK) = rho_j (w_j.G) + rho_i (w_i.G)
// The w_i.G terms are linear in G, so sum_j G is accumulated once and w_i applied after the loop.
// FP64 instructions per pair are what the sweep costs (profiles/r2k_sweep_variants.txt: 16 warps per SM run no faster than 12,
// fewer instructions do), so: pressure, viscosity and ALE momentum share ONE accumulator (they are only ever summed; the
// viscous and the ALE factor of u are added first: a += u (vf + w_j.G)), u.G is accumulated by its own FMAs, every pair sum
// is taken with gk = t^3 and the constant 5 Wc / H^2 applied once after the walk, G = (V_j t^3) Rji, and what is needed after
// the walk only (vPert_i, Af, the mass) is re-read there instead of held in registers through it (168 -> 152 registers).
// experiment knobs of the force sweep, both measured and left off (profiles/r2k_sweep_variants.txt)
#ifndef FJ_FORCE_CLAMP
#define FJ_FORCE_CLAMP 0
#endif
#ifndef FJ_FORCE_MINB
#define FJ_FORCE_MINB min_blocks(WARPS, false)
#endif
template <bool ALE, bool FROZEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, FJ_FORCE_MINB)
    k_force(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int n_bound_blocks, DevConst C, double npdm2)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread<WARPS>(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    const bool active = in_row && blk[i] >= n_bound_blocks;
    if (!__any_sync(FJ_FULL, active))
        return;
    const double4 x0i = FROZEN ? lv.x0[i] : make_double4(0, 0, 0, 0);
    const double4 pi = S.P0[i];
    const double4 vi = S.P1[i];
    const double Pi = S.P2[i].w; /* p_i/rho_i^2; vPert_i is re-read after the walk */
    const double rho_i = vi.w, irho_i = 1.0 / rho_i;
    const int b_i = S.b[i];
    const bool do_st = ALE ? (S.surfzone[i] == 1) : true;
    const double st_bound_fac = 1.0 + 0.5 * cos(0.5 * FJ_PI * 7.0 / 9.0);
    const double eps_f = C.eps_f; /* 0.001 H^2 */
    const double nu_irho_i = C.nu * irho_i;
    const double q_st = C.q_st; /* cos(3 pi/4 r/H) = cospi(0.75 r/H) */

    /* pair sums, all in units of 5 Wc / H^2 (applied after the walk) */
    double ax = 0, ay = 0, az = 0;       /* acc_ (pressure) + visc_ + the pair part of acc_ale_ */
    double sx = 0, sy = 0, sz = 0;       /* surf_t_ (no kernel gradient in it: true units) */
    double sgx = 0, sgy = 0, sgz = 0;    /* sum_j G */
    double Ug = 0.0, Pg = 0.0, Rrhoc_ = 0.0; /* sum_j u.G, sum_j w_j.G, sum_j rho_j w_j.G */

    /* aero term, Resid.cpp:267-277 (Q1: evaluated whenever cellID != -1): written to AF now, added to acc after the walk */
    const bool aero_on = S.cellID[i] != -1;
    if (aero_on)
    {
        double4 af = S.AF[i]; /* Af, deltaD */
        af.x = af.y = af.z = 0.0; /* CalcAeroAcc returns zero for NoAero */
        if (C.acase != 0)
        {
            const double4 cv = S.CV[i];
            const double4 np = S.NP[i]; /* surface normal, lam_nb */
            const double4 th = S.TH[i]; /* p, m, woccl, cellRho */
            double a3[3];
            calc_aero_acc(C, cv.x - vi.x, cv.y - vi.y, cv.z - vi.z, np, th, cv.w, S.AV[i].w * C.dx,
                          double(lv.ncount[i] + 1), a3);
            af.x = a3[0];
            af.y = a3[1];
            af.z = a3[2];
        }
        if (active)
            S.AF[i] = af;
    }

    /* one pair, branch-free (a lane's own index as j gives exact zeros: Rji = 0, gradK = 0) */
    auto pair_main = [&](const RecF& q, PairGeo& g, const bool take) {
        const double4 pj = q.p;
        const double4 vj = q.v;
        const double4 qj = q.q;
        g = pair_geo<FROZEN, true>(C, pi, x0i, pj, q.x0);
        if (FJ_FORCE_CLAMP && !take)
            g.gk = 0.0; /* every term below carries gk */
        const double ux = vj.x - vi.x, uy = vj.y - vi.y, uz = vj.z - vi.z;
        const double idist2 = fj_rcp1(g.rr + eps_f);
        const double rho_j = vj.w;
        const double s = pj.w * g.gk;                              /* V_j t^3 */
        const double Gx = s * g.rx, Gy = s * g.ry, Gz = s * g.rz; /* V_j gradK / (5 Wc / H^2) */
        const double pf = rho_j * (Pi + qj.w);
        ax = fma(-pf, Gx, ax);
        ay = fma(-pf, Gy, ay);
        az = fma(-pf, Gz, az);
        const double vf = fma(nu_irho_i, rho_j, C.nu) * ((s * g.r2c) * idist2);
        Ug = fma(uz, Gz, fma(uy, Gy, fma(ux, Gx, Ug)));
        if (ALE)
        {
            const double pjg = qj.x * Gx + qj.y * Gy + qj.z * Gz; /* V_j (vPert_j . gradK) */
            const double wv = vf + pjg;                           /* viscosity and ALE momentum both multiply u */
            ax = fma(ux, wv, ax);
            ay = fma(uy, wv, ay);
            az = fma(uz, wv, az);
            sgx += Gx;
            sgy += Gy;
            sgz += Gz;
            Pg += pjg;
            Rrhoc_ = fma(rho_j, pjg, Rrhoc_);
        }
        else
        {
            ax = fma(ux, vf, ax);
            ay = fma(uy, vf, ay);
            az = fma(uz, vf, az);
        }
    };
    /* pairwise surface tension, Kernel.h:101-113 (surface-zone particles only under ALE) */
    auto pair_st = [&](const RecF& q, const PairGeo& g, const bool take) {
        const double fac = (b_i == FJSPH_BOUND || q.b == FJSPH_BOUND) ? st_bound_fac : 1.0;
        const double sf = take ? -npdm2 * fac * cospi(q_st * (g.rr * g.ir)) * g.ir : 0.0;
        sx = fma(sf, g.rx, sx);
        sy = fma(sf, g.ry, sy);
        sz = fma(sf, g.rz, sz);
    };
    for_neighbours2<FJ_FORCE_CLAMP != 0>(
        lv, W, active, unsigned(i),
        [&](const unsigned j) {
            RecF q;
            q.p = gather(S.P0, j);
            q.v = gather(S.P1, j);
            q.q = gather(S.P2, j);
            if (do_st)
                q.b = __ldg(&S.b[j]);
            if (FROZEN)
                q.x0 = gather(lv.x0, j);
            return q;
        },
        [&](const RecF& qa, const bool ta, const RecF& qb, const bool tb) {
            PairGeo ga, gb;
            pair_main(qa, ga, ta);
            pair_main(qb, gb, tb);
            if (do_st)
            {
                pair_st(qa, ga, ta);
                pair_st(qb, gb, tb);
            }
        },
        [&](const unsigned first, const unsigned last) {
            stage_span(S.P0, first, last);
            stage_span(S.P1, first, last);
            stage_span(S.P2, first, last);
            if (do_st)
            {
                stage_span(S.b, first, last);
            }
            if (FROZEN)
            {
                stage_span(lv.x0, first, last);
            }
        });
    double Rrho_ = -Ug; /* Rrho_ -= V_j ((u + w_j - w_i) . gK) */
    if (ALE)
    {
        const double4 wi = gather_rw(S.P2, unsigned(i));
        const double pig = wi.x * sgx + wi.y * sgy + wi.z * sgz; /* vPert_i . sum_j G */
        ax += 2.0 * vi.x * pig;
        ay += 2.0 * vi.y * pig;
        az += 2.0 * vi.z * pig;
        Rrho_ += pig - Pg;
        Rrhoc_ += rho_i * pig;
    }
    /* the constant of the kernel gradient, once per particle */
    ax *= C.gk_fac;
    ay *= C.gk_fac;
    az *= C.gk_fac;
    const double Rrho_pairs = (Rrho_ * rho_i + Rrhoc_) * C.gk_fac;
    const double4 af = gather_rw(S.AF, unsigned(i));
    if (aero_on)
    {
        ax += af.x;
        ay += af.y;
        az += af.z;
    }
    if ((S.internal[i] & 0xFF) == 1)
    {
        /* NormalBoundaryRepulsion, Kernel.h:64-75,272-277 */
        const double4 bn = S.BN[i];
        const double beta = 4.0 * C.c_sound * C.c_sound;
        const double q = bn.w * C.iH;
        double kern = 0.0;
        if (q < 2.0 / 3.0)
            kern = beta * 2.0 / 3.0;
        else if (q < 1.0)
            kern = beta * (2 * q - 3.0 / 2.0 * q * q);
        else if (q < 2.0)
            kern = 0.5 * beta * ((2 - q) * (2 - q));
        const double f = C.bnd_mass / (C.bnd_mass + C.sim_mass) * kern;
        ax += f * bn.x;
        ay += f * bn.y;
        az += f * bn.z;
    }
    if (!active)
        return; /* the walk is over: nothing below votes */
    const double4 av = S.AV[i];
    const double im = 1.0 / gather_rw(S.TH, unsigned(i)).y;
    double4 acc;
    acc.x = ax + av.x + sx * im + C.gx;
    acc.y = ay + av.y + sy * im + C.gy;
    acc.z = az + av.z + sz * im + C.gz;
    acc.w = Rrho_pairs + af.w;
    S.ACC[i] = acc;
}

// ================================================================= walls (Resid.cpp:21-186)
__global__ void k_wall_velocity(Level S, const int* __restrict__ blk, int block, double vx, double vy, double vz, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] != block)
        return;
    double4 v = S.P1[i];
    v.x = vx;
    v.y = vy;
    v.z = vz;
    S.P1[i] = v;
}

// r of a pair for the wall treatments (they run with the particles moved: the build-time positions give the frozen r)
__device__ __forceinline__ double wall_pair_r(const double4* __restrict__ x0, const double4& x0i, unsigned j)
{
    const double4 q = x0[j];
    const double ex = q.x - x0i.x, ey = q.y - x0i.y, ez = q.z - x0i.z;
    return sqrt(fma(ez, ez, fma(ey, ey, ex * ex)));
}

// Set_No_Slip: v_i = 2 v_i - sum(v_j W)/sum(W) over fluid neighbours
__global__ void __launch_bounds__(FJ_ROW_WARPS * 32)
    k_wall_no_slip(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int block, DevConst C)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread(M, i, W, work);
    if (!work || !in_row || blk[i] != block)
        return;
    const double4 x0i = lv.x0[i];
    double sxx = 0, syy = 0, szz = 0, ks = 0;
    for_neighbours_simple(lv, W, true, [&](const unsigned j) {
        if (!(S.b[j] > FJSPH_PISTON))
            return;
        const double r = wall_pair_r(lv.x0, x0i, j);
        const double4 vj = gather_rw(S.P1, j);
        const double W_ = wend_W(C, r, wend_t(C, r));
        ks += W_;
        sxx += vj.x * W_;
        syy += vj.y * W_;
        szz += vj.z * W_;
    });
    if (ks > 0.0)
    {
        double4 v = S.P1[i];
        v.x = 2.0 * v.x - sxx / ks;
        v.y = 2.0 * v.y - syy / ks;
        v.z = 2.0 * v.z - szz / ks;
        S.P1[i] = v;
    }
}

__device__ __forceinline__ double eos_pressure(const DevConst& C, double rho)
{
    if (C.pressure_rel == 0)
        return C.B * (pow(rho / C.rho_rest, C.gam) - 1.0) + C.press_back;
    return C.c2 * (rho - C.rho_rest) + C.press_back;
}
__device__ __forceinline__ double eos_density(const DevConst& C, double p)
{
    if (C.pressure_rel == 0)
        return C.rho_rest * pow(((p - C.press_back) / C.B) + 1.0, 1.0 / C.gam);
    return (p - C.press_back) / C.c2 + C.rho_rest;
}
// keeps the derived gather quantities V = m/rho and p/rho^2 consistent with (rho, p)
__device__ __forceinline__ void store_thermo(Level& S, int i, double rho, double p)
{
    double4 th = S.TH[i];
    th.x = p;
    S.TH[i] = th;
    double4 a = S.P0[i];
    a.w = th.y / rho;
    S.P0[i] = a;
    double4 v = S.P1[i];
    v.w = rho;
    S.P1[i] = v;
    double4 q = S.P2[i];
    q.w = p / (rho * rho);
    S.P2[i] = q;
}

// Get_Boundary_Pressure (Adami et al. 2012): walls read fluid records only, so in-place update is race free
__global__ void __launch_bounds__(FJ_ROW_WARPS * 32)
    k_wall_pressure(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int block, DevConst C)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread(M, i, W, work);
    if (!work || !in_row || blk[i] != block)
        return;
    const double4 pi = S.P0[i];
    const double4 x0i = lv.x0[i];
    const double4 acc = S.ACC[i];
    double ks = 0, pk = 0, akx = 0, aky = 0, akz = 0;
    int near_surface = 0;
    for_neighbours_simple(lv, W, true, [&](const unsigned j) {
        if (!(S.b[j] > FJSPH_PISTON))
            return;
        const double r = wall_pair_r(lv.x0, x0i, j);
        const double4 pj = gather_rw(S.P0, j);
        const double rho_j = S.P1[j].w;
        const double p_j = S.TH[j].x;
        const double rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
        const double kern = pj.w * wend_W(C, r, wend_t(C, r));
        ks += kern;
        pk += p_j * kern;
        const double kr = kern * rho_j;
        akx += kr * rx;
        aky += kr * ry;
        akz += kr * rz;
        if (S.surfzone[j])
            near_surface = 1;
    });
    double p = 0.0;
    if (ks > 0.0)
    {
        p = (pk + ((C.gx - acc.x) * akx + (C.gy - acc.y) * aky + (C.gz - acc.z) * akz)) / ks;
        if (near_surface)
            p = fmax(0.0, p);
    }
    store_thermo(S, i, eos_density(C, p), p);
}

// Boundary_Ghost: Rrho_i = -rho_i sum V_j Vji.gradK over all neighbours; near_inlet = no PIPE/FREE neighbour.
__global__ void __launch_bounds__(FJ_ROW_WARPS * 32)
    k_wall_ghost(Level S, RunView lv, RowMap M, const int* __restrict__ blk, int block, DevConst C,
                 int* __restrict__ near_inlet_out)
{
    int i, W;
    bool work;
    const bool in_row = fj_row_thread(M, i, W, work);
    if (!work || !in_row || blk[i] != block)
        return;
    const double4 pi = S.P0[i];
    const double4 x0i = lv.x0[i];
    const double4 vi = S.P1[i];
    double Rrhoi = 0.0;
    int near_inlet = 1;
    for_neighbours_simple(lv, W, true, [&](const unsigned j) {
        const int bj = S.b[j];
        if (bj == FJSPH_PIPE || bj == FJSPH_FREE)
            near_inlet = 0;
        const double r = wall_pair_r(lv.x0, x0i, j);
        const double4 pj = gather_rw(S.P0, j);
        const double4 vj = gather_rw(S.P1, j);
        const double rx = pj.x - pi.x, ry = pj.y - pi.y, rz = pj.z - pi.z;
        const double gk = wend_gk(C, r, wend_t(C, r));
        Rrhoi -= pj.w * gk * ((vj.x - vi.x) * rx + (vj.y - vi.y) * ry + (vj.z - vi.z) * rz);
    });
    double4 acc = S.ACC[i];
    acc.w = Rrhoi * vi.w;
    S.ACC[i] = acc;
    near_inlet_out[i] = near_inlet;
}

// Boundary_DBC: the reference accumulates out of bounds (Resid.cpp:84,107); its defined effect is acc = 0
__global__ void k_wall_dbc(Level S, const int* __restrict__ blk, int block, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] != block)
        return;
    double4 acc = S.ACC[i];
    acc.x = acc.y = acc.z = 0.0;
    S.ACC[i] = acc;
}

// Check_Pipe_Outlet (Containment.cpp:822-847), constVel part: PIPE -> FREE past the block's aero plane
__global__ void k_pipe_outlet(Level S, const int* __restrict__ blk, int block, double nx, double ny, double nz,
                              double aeroconst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] != block)
        return;
    if (S.b[i] == FJSPH_PIPE)
    {
        const double4 p = S.P0[i];
        if (p.x * nx + p.y * ny + p.z * nz > aeroconst)
            S.b[i] = FJSPH_FREE;
    }
}

RunView run_view(FjsphEngine* e)
{
    RunView v;
    v.erun = e->erun;
    v.erows = e->erows;
    v.ncount = e->ncount;
    v.x0 = e->x0;
    v.ecap = e->ecap;
    return v;
}

// Launches a sweep over the owned rows.  Slab mode, with a forward exchange in flight on the comm stream: the rows of
// the INTERIOR class first (none of their neighbours is a ghost, so they run beside the exchange), then -- once the
// main stream has waited for the ghosts -- the rows of the EDGE class.  launch(M) must queue the kernel for the rows
// of M.  Call with e->slab.hold set while the KScope of the sweep is constructed (SplitScope below).
template <class F>
int launch_split(FjsphEngine* e, F&& launch)
{
    if (fj_halo_overlappable(e))
    {
        launch(fj_row_map(e, 0, 1));
        int st = fj_halo_wait(e);
        if (st)
            return st;
        launch(fj_row_map(e, 1, 1));
        e->slab.overlapped++;
    }
    else
    {
        int st = fj_halo_wait(e);
        if (st)
            return st;
        launch(fj_row_map(e, 0, fj_owned_classes(e)));
    }
    return FJSPH_OK;
}
// KScope that does not wait for the exchange in flight when the sweep is going to split
struct SplitScope
{
    bool split;
    KScope ks;
    static bool arm(FjsphEngine* e)
    {
        const bool s = fj_halo_overlappable(e);
        e->slab.hold = s;
        return s;
    }
    SplitScope(FjsphEngine* e, const char* name) : split(arm(e)), ks(e, name, split ? 2 : 1) { e->slab.hold = false; }
};

int need_list(FjsphEngine* e, const char* who)
{
    if (!e->list_valid)
    {
        fj_set_error("%s: neighbour list not built (call fjsph_build_neighbours first)", who);
        return FJSPH_ERR_STATE;
    }
    return FJSPH_OK;
}

// picks the instantiation of a sweep: FROZEN once the level-1 positions differ from the ones the list was built on,
// WARPS = rows per CTA (e->sweep_warps).  f(FROZEN, WARPS) receives both as integral constants.
template <class F>
void by_shape(const FjsphEngine* e, bool frozen, F&& f)
{
    using std::integral_constant;
    if (e->sweep_warps == 4)
    {
        if (frozen)
            f(integral_constant<bool, true>{}, integral_constant<int, 4>{});
        else
            f(integral_constant<bool, false>{}, integral_constant<int, 4>{});
    }
    else
    {
        if (frozen)
            f(integral_constant<bool, true>{}, integral_constant<int, 8>{});
        else
            f(integral_constant<bool, false>{}, integral_constant<int, 8>{});
    }
}

template <bool SURF, bool DISS>
int sweep_surf1(FjsphEngine* e, int* near_warps)
{
    const RunView lv = run_view(e);
    return launch_split(e, [&](const RowMap& M) {
        by_shape(e, e->x_moved, [&](auto fr, auto wp) {
            constexpr bool FR = decltype(fr)::value;
            constexpr int WP = decltype(wp)::value;
            k_surf1_diss<SURF, DISS, FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(e->lv[1], lv, M, e->blk,
                                                                                           e->n_bound_blocks, e->C, near_warps);
        });
    });
}
template <bool SURF23, bool SHIFT, int CLASS>
int sweep_surf23(FjsphEngine* e)
{
    const RunView lv = run_view(e);
    return launch_split(e, [&](const RowMap& M) {
        by_shape(e, e->x_moved, [&](auto fr, auto wp) {
            constexpr bool FR = decltype(fr)::value;
            constexpr int WP = decltype(wp)::value;
            k_surf23_shift<SURF23, SHIFT, CLASS, FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(
                e->lv[1], lv, M, e->blk, e->n_bound_blocks, e->C);
        });
    });
}

} // namespace

// ------------------------------------------------------------------ host wrappers
int fj_prestep(FjsphEngine* e, double* npd)
{
    int st = need_list(e, "prestep");
    if (st)
        return st;
    const RowMap M = fj_row_map(e, 0, fj_owned_classes(e));
    {
        KScope ks(e, "prestep", 1);
        by_shape(e, e->x_moved, [&](auto fr, auto wp) {
            constexpr bool FR = decltype(fr)::value;
            constexpr int WP = decltype(wp)::value;
            k_prestep<FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(e->lv[1], run_view(e), M, e->C, e->red);
        });
    }
    FJ_CUDA(cudaGetLastError());
    double sum = 0.0;
    st = fj_reduce_sum(e, int(e->n_warp), 1, &sum);
    if (st)
        return st;
    /* npd = npd_ / end (Shifting.cpp:121); Q4: the race-free sum */
    e->npd = sum / fj_total_count(e);
    if (npd)
        *npd = e->npd;
    return FJSPH_OK;
}

int fj_aero_velocity(FjsphEngine* e)
{
    int st = need_list(e, "aero_velocity");
    if (st)
        return st;
    if (e->P.asource == 1)
        return fj_aero_velocity_mesh(e); /* FindCell; may erase particles and redo the list and the prestep */
    const int n = int(e->n_owned);
    KScope ks(e, "aero_vel", 1);
    k_aero_velocity<<<fj_blocks(n, 256), 256, 0, e->stream>>>(e->lv[1], e->ncount, e->blk, e->n_bound_blocks, e->C, n);
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

int fj_surface_and_dissipation(FjsphEngine* e, bool do_surface, bool do_dissipation, bool fuse_shift)
{
    int st = need_list(e, "detect_surface/dissipation");
    if (st)
        return st;
    if (do_surface && do_dissipation)
    {
        /* count the warps holding near-surface particles for the split decision below; the count of THIS pass is read
           at the next one (it has crossed PCIe by then: every pass ends in synchronising readbacks) */
        FJ_CUDA(cudaMemsetAsync(e->d_flag + 2, 0, sizeof(int), e->stream));
        {
            SplitScope ks(e, "surf1+diss");
            int* nw = e->d_flag + 2;
            st = sweep_surf1<true, true>(e, nw);
        }
        FJ_CUDA(cudaMemcpyAsync(e->h_flag + 2, e->d_flag + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    }
    else if (do_surface)
    {
        SplitScope ks(e, "surf1");
        st = sweep_surf1<true, false>(e, nullptr);
    }
    else if (do_dissipation)
    {
        SplitScope ks(e, "diss");
        st = sweep_surf1<false, true>(e, nullptr);
    }
    if (st)
        return st;
    if (do_surface)
    {
        st = fj_halo_exchange(e, 1, FJ_HX_P4); /* loop-1 normals and surf flags of the ghosts */
        if (st)
            return st;
    }
    if (do_surface && do_dissipation && fuse_shift && e->P.ale && FJ_FUSE_SHIFT) /* the fused sweep reads what surf1 + diss left in AV.w */
    {
        SplitScope ks(e, "surf2+3+shift");
        /* Two launches (lean bulk, then the near-surface rest) pay when few warps hold near-surface particles: with the
           lean body at ~0.6 of the full one, below ~0.4 of the warps (block: 0.17; a 1 M droplet: 0.5, where one fused
           launch is faster).  The fraction is the one counted at the previous pass. */
        const double near_frac = double(e->h_flag[2]) / double(std::max(1u, e->n_warp));
        if (e->split_surface_sweep && near_frac < e->split_surface_below)
        {
            st = sweep_surf23<true, true, 1>(e);
            if (!st)
                st = sweep_surf23<true, true, 2>(e);
        }
        else
            st = sweep_surf23<true, true, 0>(e);
    }
    else if (do_surface)
    {
        {
            SplitScope ks(e, "surf2+3");
            st = sweep_surf23<true, false, 0>(e);
        }
        if (st)
            return st;
        if (fuse_shift && e->P.ale)
        {
            FJ_CUDA(cudaGetLastError());
            return fj_shift(e);
        }
    }
    if (st)
        return st;
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

int fj_shift(FjsphEngine* e)
{
    int st = need_list(e, "shift");
    if (st)
        return st;
    if (!e->P.ale)
        return FJSPH_OK;
    const RowMap M = fj_row_map(e, 0, fj_owned_classes(e));
    KScope ks(e, "shift", 1);
    by_shape(e, e->x_moved, [&](auto fr, auto wp) {
        constexpr bool FR = decltype(fr)::value;
        constexpr int WP = decltype(wp)::value;
        k_surf23_shift<false, true, 0, FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(e->lv[1], run_view(e), M, e->blk,
                                                                                             e->n_bound_blocks, e->C);
    });
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

int fj_check_pipe_outlet(FjsphEngine* e)
{
    if (e->P.asource == 1)
        return fj_pipe_outlet_mesh(e);
    const int n = int(e->n_owned);
    for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
    {
        const HostBlock& B = e->blocks[bl];
        if (B.aeroconst == 9999999.0)
            continue; /* plane undefined: x.dot(default) > default never holds for sane coordinates */
        KScope ks(e, "pipe_outlet", 1);
        k_pipe_outlet<<<fj_blocks(n, 256), 256, 0, e->stream>>>(e->lv[1], e->blk, int(bl), B.aero_norm[0],
                                                                B.aero_norm[1], B.aero_norm[2], B.aeroconst, n);
    }
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

int fj_forces(FjsphEngine* e, int level_idx, double npd)
{
    int st = need_list(e, "forces");
    if (st)
        return st;
    /* Resid.cpp:437-448 */
    const double lam = (6.0 / 81.0 * std::pow((2.0 * e->P.H), 3.0) / std::pow(FJ_PI, 4.0) *
                        (9.0 / 4.0 * std::pow(FJ_PI, 3.0) - 6.0 * FJ_PI - 4.0));
    const double npdm2 = (0.5 * e->P.sig / lam) / (npd * npd);
    /* the list was built on level 1's positions: any other level, or level 1 after a move, takes r from x0 */
    const bool frozen = e->x_moved || level_idx != 1;
    {
        SplitScope ks(e, "force");
        RunView lv = run_view(e);
        st = launch_split(e, [&](const RowMap& M) {
            by_shape(e, frozen, [&](auto fr, auto wp) {
                constexpr bool FR = decltype(fr)::value;
                constexpr int WP = decltype(wp)::value;
                if (e->P.ale)
                    k_force<true, FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(e->lv[level_idx], lv, M, e->blk,
                                                                                        e->n_bound_blocks, e->C, npdm2);
                else
                    k_force<false, FR, WP><<<fj_row_grid(M, WP), WP * 32, 0, e->stream>>>(e->lv[level_idx], lv, M, e->blk,
                                                                                         e->n_bound_blocks, e->C, npdm2);
            });
        });
        if (st)
            return st;
    }
    e->force_evals++;
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

// Wall treatment at the head of Do_NB_Iter (Newmark_Beta.cpp:69-132) and of the RK stages
// (Runge_Kutta.cpp:36-135).  nb_comparator selects the schedule comparison (Q10).
int fj_walls(FjsphEngine* e, int level_idx, bool nb_comparator)
{
    if (e->n_bound_blocks == 0)
        return FJSPH_OK;
    int st = need_list(e, "walls");
    if (st)
        return st;
    const int n = int(e->n_owned);
    const int nb = fj_blocks(n, 256);
    RunView lv = run_view(e);
    const RowMap M = fj_row_map(e, 0, fj_owned_classes(e));
    const unsigned grid = fj_row_grid(M);
    Level& S = e->lv[level_idx];
    for (int bl = 0; bl < e->n_bound_blocks; ++bl)
    {
        const HostBlock& B = e->blocks[bl];
        double vel[3] = {B.vels[0], B.vels[1], B.vels[2]};
        if (!B.times.empty())
        {
            vel[0] = vel[1] = vel[2] = 0.0;
            for (size_t t = 0; t < B.times.size(); ++t)
            {
                const bool take = nb_comparator ? (e->P.current_time > B.times[t]) : (B.times[t] > e->P.current_time);
                if (take)
                    for (int d = 0; d < 3; ++d) vel[d] = B.vels[3 * t + d];
            }
        }
        KScope ks(e, "walls", 3);
        k_wall_velocity<<<nb, 256, 0, e->stream>>>(S, e->blk, bl, vel[0], vel[1], vel[2], n);
        if (B.no_slip)
            k_wall_no_slip<<<grid, FJ_ROW_WARPS * 32, 0, e->stream>>>(S, lv, M, e->blk, bl, e->C);
        switch (B.bound_solver)
        {
        case FJSPH_DBC: k_wall_dbc<<<nb, 256, 0, e->stream>>>(S, e->blk, bl, n); break;
        case FJSPH_PRESSURE_G: k_wall_pressure<<<grid, FJ_ROW_WARPS * 32, 0, e->stream>>>(S, lv, M, e->blk, bl, e->C); break;
        case FJSPH_GHOST: k_wall_ghost<<<grid, FJ_ROW_WARPS * 32, 0, e->stream>>>(S, lv, M, e->blk, bl, e->C, e->near_inlet); break;
        default: break;
        }
    }
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}
