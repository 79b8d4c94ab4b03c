/* fjsph_oracle.cpp — TEST INFRASTRUCTURE ONLY (see fjsph_oracle.h).
 *
 * CPU restatement of FJSPH's WCSPH time step on plain arrays.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/src).  Build with -DDIM=2 or -DDIM=3.
 * Parity build: -O2 -ffp-contract=off, no OpenMP (deterministic summation in ascending-j order).
 * Timing build: -O3 -ffast-math -funroll-loops -fopenmp -march=native (the reference's makefile:16).
 *
 * PARITY UNPINNED (no reference tests / golden vectors exist and the reference cannot be built here).
 * Third-party arithmetic restated from the published algorithms:
 *   nanoflann metric_L2_Simple radiusSearch  -> d2 = sum_d (q_d - p_d)^2 accumulated d=0..DIM-1, d2 < radius
 *   Eigen ColPivHouseholderQR::{isInvertible,inverse}, SelfAdjointEigenSolver::computeDirect (3.4)
 * Known reference UB that cannot be reproduced and is restated as its defined part:
 *   Boundary_DBC writes RV_[jj.first] out of bounds (Resid.cpp:84,107) -> wall acc = 0;
 *   RK fixed-velocity inlet indexes limits[jj] (Runge_Kutta.cpp:211,433) -> restated with limits[block];
 *   FindCell with cellID == -3 (Containment.cpp:592-600) -> not contained.
 */
#include "fjsph_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>

#ifndef DIM
#define DIM 3
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef double real;
static const real MEPS = std::numeric_limits<real>::epsilon();
static const real default_val = 9999999.0; /* VarDefs.h:194 */
static const long c_no_cell = -3;         /* VarDefs.h:197 */

/* ------------------------------------------------------------------ small vector helpers */
struct Vec
{
    real a[DIM];
    real& operator[](int i) { return a[i]; }
    real const& operator[](int i) const { return a[i]; }
};
static inline Vec vzero()
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = 0.0;
    return r;
}
static inline Vec operator+(Vec const& x, Vec const& y)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = x[d] + y[d];
    return r;
}
static inline Vec operator-(Vec const& x, Vec const& y)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = x[d] - y[d];
    return r;
}
static inline Vec operator-(Vec const& x)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = -x[d];
    return r;
}
static inline Vec operator*(real s, Vec const& x)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = s * x[d];
    return r;
}
static inline Vec operator*(Vec const& x, real s)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = x[d] * s;
    return r;
}
static inline Vec operator/(Vec const& x, real s)
{
    Vec r;
    for (int d = 0; d < DIM; ++d) r[d] = x[d] / s;
    return r;
}
static inline Vec& operator+=(Vec& x, Vec const& y)
{
    for (int d = 0; d < DIM; ++d) x[d] += y[d];
    return x;
}
static inline Vec& operator-=(Vec& x, Vec const& y)
{
    for (int d = 0; d < DIM; ++d) x[d] -= y[d];
    return x;
}
static inline real dot(Vec const& x, Vec const& y)
{
    real s = 0.0;
    for (int d = 0; d < DIM; ++d) s += x[d] * y[d];
    return s;
}
static inline real sqnorm(Vec const& x) { return dot(x, x); }
static inline real norm(Vec const& x) { return std::sqrt(dot(x, x)); }
/* Eigen's normalized(): x / norm when squaredNorm > 0, else x unchanged (i.e. zero stays zero). */
static inline Vec normalized(Vec const& x)
{
    real z = sqnorm(x);
    if (z > 0.0)
        return x / std::sqrt(z);
    return x;
}

#if DIM == 3
static inline Vec cross3(Vec const& a, Vec const& b)
{
    Vec r;
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
    return r;
}
#endif

struct Mat
{
    real a[DIM][DIM];
};
static inline Mat mzero()
{
    Mat m;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) m.a[i][j] = 0.0;
    return m;
}
static inline Mat midentity()
{
    Mat m = mzero();
    for (int i = 0; i < DIM; ++i) m.a[i][i] = 1.0;
    return m;
}
static inline Vec mul(Mat const& m, Vec const& x)
{
    Vec r;
    for (int i = 0; i < DIM; ++i)
    {
        real s = 0.0;
        for (int j = 0; j < DIM; ++j) s += m.a[i][j] * x[j];
        r[i] = s;
    }
    return r;
}

/* ------------------------------------------------------------------ Eigen restatements */

/* ColPivHouseholderQR (Eigen 3.4 algorithm): Householder QR with column pivoting on the largest
 * remaining column norm; rank = #{ |R_kk| > eps*n*max|R_kk| }; inverse = solve(I).
 * Column norms are recomputed instead of down-dated (differs from Eigen only in rounding).
 * Returns isInvertible(); inv is filled only when invertible. */
static int qr_inverse(Mat const& A, Mat& inv)
{
    const int n = DIM;
    real qr[DIM][DIM];
    real hco[DIM];
    int perm[DIM];
    for (int i = 0; i < n; ++i)
    {
        perm[i] = i;
        for (int j = 0; j < n; ++j) qr[i][j] = A.a[i][j];
    }
    real maxcol = 0.0;
    for (int j = 0; j < n; ++j)
    {
        real s = 0.0;
        for (int i = 0; i < n; ++i) s += qr[i][j] * qr[i][j];
        maxcol = std::max(maxcol, std::sqrt(s));
    }
    real const thr_helper = (maxcol * MEPS / real(n)) * (maxcol * MEPS / real(n));
    int nonzero_pivots = n;
    real maxpivot = 0.0;
    for (int k = 0; k < n; ++k)
    {
        /* pivot column */
        int big = k;
        real bigsq = -1.0;
        for (int j = k; j < n; ++j)
        {
            real s = 0.0;
            for (int i = k; i < n; ++i) s += qr[i][j] * qr[i][j];
            if (s > bigsq)
            {
                bigsq = s;
                big = j;
            }
        }
        if (nonzero_pivots == n && bigsq < thr_helper * real(n - k))
            nonzero_pivots = k;
        if (big != k)
        {
            for (int i = 0; i < n; ++i) std::swap(qr[i][k], qr[i][big]);
            std::swap(perm[k], perm[big]);
        }
        /* Householder vector for column k, rows k..n-1 (Eigen makeHouseholderInPlace) */
        real c0 = qr[k][k];
        real tailsq = 0.0;
        for (int i = k + 1; i < n; ++i) tailsq += qr[i][k] * qr[i][k];
        real beta, tau;
        if (tailsq <= std::numeric_limits<real>::min())
        {
            tau = 0.0;
            beta = c0;
            for (int i = k + 1; i < n; ++i) qr[i][k] = 0.0;
        }
        else
        {
            beta = std::sqrt(c0 * c0 + tailsq);
            if (c0 >= 0.0)
                beta = -beta;
            for (int i = k + 1; i < n; ++i) qr[i][k] /= (c0 - beta);
            tau = (beta - c0) / beta;
        }
        qr[k][k] = beta;
        hco[k] = tau;
        if (std::fabs(beta) > maxpivot)
            maxpivot = std::fabs(beta);
        /* apply H = I - tau v v^T (v = [1, essential]) to the trailing columns */
        for (int j = k + 1; j < n; ++j)
        {
            real s = qr[k][j];
            for (int i = k + 1; i < n; ++i) s += qr[i][k] * qr[i][j];
            s *= tau;
            qr[k][j] -= s;
            for (int i = k + 1; i < n; ++i) qr[i][j] -= s * qr[i][k];
        }
    }
    int rank = 0;
    real const premult = maxpivot * (MEPS * real(n));
    for (int i = 0; i < nonzero_pivots; ++i)
        if (std::fabs(qr[i][i]) > premult)
            rank++;
    if (rank != n)
        return 0;

    /* inverse: for each unit vector e_c solve A x = e_c : c = Q^T e_c ; R y = c ; x[perm] = y */
    for (int c = 0; c < n; ++c)
    {
        real rhs[DIM];
        for (int i = 0; i < n; ++i) rhs[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; ++k)
        {
            real s = rhs[k];
            for (int i = k + 1; i < n; ++i) s += qr[i][k] * rhs[i];
            s *= hco[k];
            rhs[k] -= s;
            for (int i = k + 1; i < n; ++i) rhs[i] -= s * qr[i][k];
        }
        for (int i = n - 1; i >= 0; --i)
        {
            real s = rhs[i];
            for (int j = i + 1; j < n; ++j) s -= qr[i][j] * rhs[j];
            rhs[i] = s / qr[i][i];
        }
        for (int i = 0; i < n; ++i) inv.a[perm[i]][c] = rhs[i];
    }
    return 1;
}

/* SelfAdjointEigenSolver<DIMxDIM>::computeDirect(...).eigenvalues().minCoeff()
 * (Eigen 3.4 direct_selfadjoint_eigenvalues; lower triangle only; shift by trace/n, scale by max|a|). */
static real min_eigenvalue(Mat const& A)
{
#if DIM == 3
    real const shift = (A.a[0][0] + A.a[1][1] + A.a[2][2]) / 3.0;
    real m00 = A.a[0][0] - shift, m11 = A.a[1][1] - shift, m22 = A.a[2][2] - shift;
    real m10 = A.a[1][0], m20 = A.a[2][0], m21 = A.a[2][1];
    real scale = std::max(
        std::max(std::fabs(m00), std::max(std::fabs(m11), std::fabs(m22))),
        std::max(std::fabs(m10), std::max(std::fabs(m20), std::fabs(m21)))
    );
    if (scale > 0.0)
    {
        m00 /= scale;
        m11 /= scale;
        m22 /= scale;
        m10 /= scale;
        m20 /= scale;
        m21 /= scale;
    }
    real const s_inv3 = 1.0 / 3.0;
    real const s_sqrt3 = std::sqrt(3.0);
    real const c0 = m00 * m11 * m22 + 2.0 * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 -
                    m22 * m10 * m10;
    real const c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
    real const c2 = m00 + m11 + m22;
    real const c2_over_3 = c2 * s_inv3;
    real a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
    a_over_3 = std::max(a_over_3, 0.0);
    real const half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
    real q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
    q = std::max(q, 0.0);
    real const rho = std::sqrt(a_over_3);
    real const theta = std::atan2(std::sqrt(q), half_b) * s_inv3;
    real const cos_theta = std::cos(theta);
    real const sin_theta = std::sin(theta);
    real const r0 = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    real const r1 = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    real const r2 = c2_over_3 + 2.0 * rho * cos_theta;
    real const rmin = std::min(r0, std::min(r1, r2));
    return rmin * scale + shift;
#else
    real const shift = (A.a[0][0] + A.a[1][1]) / 2.0;
    real m00 = A.a[0][0] - shift, m11 = A.a[1][1] - shift, m10 = A.a[1][0];
    real scale = std::max(std::fabs(m00), std::max(std::fabs(m11), std::fabs(m10)));
    if (scale > 0.0)
    {
        m00 /= scale;
        m11 /= scale;
        m10 /= scale;
    }
    real const t0 = 0.5 * std::sqrt((m00 - m11) * (m00 - m11) + 4.0 * m10 * m10);
    real const t1 = 0.5 * (m00 + m11);
    return (t1 - t0) * scale + shift;
#endif
}

/* ------------------------------------------------------------------ Kernel.h */
/* Kernel.h:37-45 (Wendland C2). pow(x,4) restated as (x*x)*(x*x). */
static inline real Kernel(real dist, real H, real Wc)
{
    real const t = 1 - 0.5 * dist / H;
    real const t2 = t * t;
    return (t2 * t2) * (2 * dist / H + 1) * Wc;
}
/* Kernel.h:48-61. pow(x,3) restated as x*x*x. */
static inline Vec GradK(Vec const& Rij, real dist, real H, real Wc)
{
    if (dist / H < 1e-12)
        return vzero();
    real const t = 1 - 0.5 * dist / H;
    real const t3 = t * t * t;
    Vec g;
    for (int d = 0; d < DIM; ++d) g[d] = 5.0 * (Rij[d] / (H * H)) * t3 * Wc;
    return g;
}
/* Kernel.h:64-75 */
static inline real BoundaryKernel(real dist, real H, real beta)
{
    real const q = dist / H;
    if (q < 2.0 / 3.0)
        return beta * 2.0 / 3.0;
    else if (2.0 / 3.0 <= q && q < 1.0)
        return beta * (2 * q - 3.0 / 2.0 * q * q);
    else if (1 <= q && q < 2)
        return 0.5 * beta * ((2 - q) * (2 - q));
    return 0;
}

/* ------------------------------------------------------------------ state */
struct State
{
    size_t n = 0;
    std::vector<long> part_id, cellID;
    std::vector<int> b, surf, surfzone, internal, ipt_n_failed;
    std::vector<Vec> xi, v, acc, Af, aVisc, cellV, gradRho, norm, bNorm, vPert;
    std::vector<Mat> L;
    std::vector<real> Rrho, rho, p, m, curve, norm_curve, woccl, pDist, deltaD, cellP, cellRho, colourG,
        colour, lam, lam_nb, kernsum, y;

    void resize(size_t n_)
    {
        n = n_;
        part_id.resize(n, 0);
        cellID.resize(n, c_no_cell);
        b.resize(n, 0);
        surf.resize(n, 0);
        surfzone.resize(n, 0);
        internal.resize(n, 0);
        ipt_n_failed.resize(n, 0);
        for (auto* f : vecs()) f->resize(n, vzero());
        L.resize(n, mzero());
        for (auto* f : scalars()) f->resize(n, 0.0);
    }
    std::vector<std::vector<Vec>*> vecs()
    {
        return {&xi, &v, &acc, &Af, &aVisc, &cellV, &gradRho, &norm, &bNorm, &vPert};
    }
    std::vector<std::vector<real>*> scalars()
    {
        return {&Rrho, &rho,     &p,       &m,      &curve, &norm_curve, &woccl, &pDist,   &deltaD,
                &cellP, &cellRho, &colourG, &colour, &lam,   &lam_nb,     &kernsum, &y};
    }
    void erase(size_t i)
    {
        part_id.erase(part_id.begin() + i);
        cellID.erase(cellID.begin() + i);
        b.erase(b.begin() + i);
        surf.erase(surf.begin() + i);
        surfzone.erase(surfzone.begin() + i);
        internal.erase(internal.begin() + i);
        ipt_n_failed.erase(ipt_n_failed.begin() + i);
        for (auto* f : vecs()) f->erase(f->begin() + i);
        L.erase(L.begin() + i);
        for (auto* f : scalars()) f->erase(f->begin() + i);
        n--;
    }
    /* SPHPart(X, pj, bound, p_id) constructor, Var.h:548-590, inserted at position pos */
    void insert_from(size_t pos, Vec const& X, size_t src, int bound, long pid)
    {
        real rho_ = rho[src], p_ = p[src], m_ = m[src], cP = cellP[src], cR = cellRho[src];
        Vec v_ = v[src];
        part_id.insert(part_id.begin() + pos, pid);
        cellID.insert(cellID.begin() + pos, c_no_cell);
        b.insert(b.begin() + pos, bound);
        surf.insert(surf.begin() + pos, 0);
        surfzone.insert(surfzone.begin() + pos, 0);
        internal.insert(internal.begin() + pos, 0);
        ipt_n_failed.insert(ipt_n_failed.begin() + pos, 0);
        for (auto* f : vecs()) f->insert(f->begin() + pos, vzero());
        L.insert(L.begin() + pos, mzero());
        for (auto* f : scalars()) f->insert(f->begin() + pos, 0.0);
        n++;
        xi[pos] = X;
        v[pos] = v_;
        rho[pos] = rho_;
        p[pos] = p_;
        m[pos] = m_;
        cellP[pos] = cP;
        cellRho[pos] = cR;
    }
};

struct Block /* bound_block, Var.h:779-859 */
{
    long first = 0, second = 0;
    int is_fluid = 0, bound_solver = 0, no_slip = 0, block_type = 0, fixed_vel_or_dynamic = 0;
    size_t nTimes = 0;
    std::vector<real> times;
    std::vector<Vec> vels;
    Vec insert_norm, delete_norm, aero_norm;
    real insconst = default_val, delconst = default_val, aeroconst = default_val;
    std::vector<long> back;
    std::vector<std::vector<long>> buffer;
};

enum
{
    DBC = 0,
    pressure_G,
    ghost
}; /* VarDefs.h:142-147 */
enum
{
    inletZone = 6
}; /* VarDefs.h:118-129 */
enum
{
    NoAero = 0,
    Gissler
};
enum
{
    constVel = 0,
    meshInfl
};

/* MESH, Var.h:396-451 (faces as vertex lists, leftright, cell -> faces, cell centres and solution) */
struct Mesh
{
    std::vector<Vec> verts;
    std::vector<std::vector<long>> faces;
    std::vector<std::pair<int, int>> leftright;
    std::vector<std::vector<long>> cFaces;
    std::vector<Vec> cCentre, cVel;
    std::vector<real> cP, cRho;
    size_t size() const { return cCentre.size(); }
};

struct Orc
{
    OrcParams P;
    State pn, pnp1;
    Mesh cells;
    std::vector<size_t> pipe_outlet_del;
    int first_cell_errors = 0; /* FirstCell's exit(-1) (Containment.cpp:563-571), reported instead of exiting */
    std::vector<Block> limits;
    size_t n_bound_blocks = 0, n_fluid_blocks = 0;
    size_t bound_points = 0, total_points = 0, fluid_points = 0;
    long next_part_id = 0;
    size_t max_points = 9999999;
    size_t delete_count = 0;
    /* neighbour list (CSR, ascending j; the reference's order is KD-tree traversal order) */
    std::vector<long> nb_off;
    std::vector<long> nb_idx;
    std::vector<real> nb_d2;
    /* Integrator members, Integration.h:53-68 */
    real safe_dt = 0.0;
    real maxf = MEPS, maxAf = MEPS, maxRho_pc = MEPS, maxRhoi = MEPS, maxdrho = MEPS, minST = 9999999.0,
         maxU = MEPS, maxShift = MEPS;
    size_t start_index = 0, end_index = 0;
    unsigned iteration = 0;
    int last_nadd = 0, last_ndel = 0;
};

/* ------------------------------------------------------------------ EOS, Var.h:203-236 */
static inline real get_pressure(OrcParams const& P, real rho)
{
    if (P.pressure_rel == 0)
        return P.B * (std::pow(rho / P.rho_rest, P.gam) - 1) + P.press_back;
    return P.speed_sound * P.speed_sound * (rho - P.rho_rest) + P.press_back;
}
static inline real get_density(OrcParams const& P, real press)
{
    if (P.pressure_rel == 0)
        return P.rho_rest * std::pow(((press - P.press_back) / P.B) + 1.0, 1.0 / P.gam);
    return (press - P.press_back) / (P.speed_sound * P.speed_sound) + P.rho_rest;
}

/* ------------------------------------------------------------------ Set_Values, IO.cpp:26-128 */
/* Geometry.cpp:282-308: lattice -2(H+dx)..2(H+dx) accumulated with x += dx; count d2 < 4H^2 from 0. */
static real get_n_full(real dx, real H)
{
    real const search_radius = 4.0 * H * H;
    size_t count = 0;
    for (real x = -2.0 * (H + dx); x <= 2.0 * (H + dx); x += dx)
        for (real y = -2.0 * (H + dx); y <= 2.0 * (H + dx); y += dx)
        {
#if DIM == 3
            for (real z = -2.0 * (H + dx); z <= 2.0 * (H + dx); z += dx)
            {
                real d2 = (0.0 - x) * (0.0 - x);
                d2 += (0.0 - y) * (0.0 - y);
                d2 += (0.0 - z) * (0.0 - z);
                if (d2 < search_radius)
                    count++;
            }
#else
            real d2 = (0.0 - x) * (0.0 - x);
            d2 += (0.0 - y) * (0.0 - y);
            if (d2 < search_radius)
                count++;
#endif
        }
    return real(count);
}

extern "C" void orc_default_params(OrcParams* p, int dim)
{
    std::memset(p, 0, sizeof(*p));
    p->dim = dim;
    p->ale = 1;
    p->pressure_rel = 0;
    p->solver_type = 0;
    p->acase = NoAero;
    p->asource = constVel;
    p->use_lam = 1;
    p->use_TAB_def = 0;
    p->max_subits = 20;
    p->n_stable = 0;
    p->n_stable_limit = 10;
    p->n_unstable = 0;
    p->n_unstable_limit = 3;
    p->particle_step = -1.0;
    p->H_fac = 2.0;
    p->rho_rest = 1000.0;
    p->press_pipe = 0.0;
    p->press_back = 0.0;
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
    p->grav[0] = p->grav[1] = p->grav[2] = 0.0;
    p->grav[dim - 1] = -9.81; /* Var.h:349-354 */
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
    p->delta_t_min = 0.0;
    p->max_shift_vel = 9999999;
    p->current_time = 0.0;
    p->last_frame_time = 0.0;
    p->frame_time_interval = 1.0;
}

extern "C" void orc_set_values(OrcParams* p)
{
    OrcParams& P = *p;
    P.B = P.rho_rest * std::pow(P.speed_sound, 2) / P.gam; /* IO.cpp:32-33 */
    P.rho_pipe = get_density(P, P.press_pipe);             /* :36 */
    if (P.rho_max == 1500 && P.rho_min == 500)             /* :39-43 */
    {
        P.rho_max = P.rho_rest * (1.0 + P.rho_var * 0.01);
        P.rho_min = P.rho_rest * (1.0 - P.rho_var * 0.01);
    }
    P.dx = P.particle_step * std::pow(P.rho_pipe / P.rho_rest, 1.0 / DIM); /* :45 */
    P.nb_beta = 0.25;                                                    /* :47-48 */
    P.nb_gamma = 0.5;
    P.sim_mass = P.rho_rest * std::pow(P.particle_step, DIM); /* :51-53 */
    P.bnd_mass = P.sim_mass;
    P.sos = std::sqrt(P.temp_g * P.R_g * P.gamma_g); /* :55 */
    if (P.delta_t_min > 0)                           /* :68-71 */
        P.delta_t = P.delta_t_min;
    else
        P.delta_t = 2E-010;
    P.H = P.H_fac * P.particle_step; /* :73-75 */
    P.H_sq = P.H * P.H;
    P.sr = 4 * P.H_sq;
    P.dsph_cont = 2.0 * P.dsph_delta * P.H * P.speed_sound; /* :77 */
    P.nu = P.mu / P.rho_rest;                               /* :81 */
#if DIM == 2
    P.W_correc = 7.0 / (4.0 * M_PI * P.H * P.H); /* :87 */
#else
    P.W_correc = (21 / (16 * M_PI * P.H * P.H * P.H)); /* :94 */
#endif
    P.W_dx = Kernel(P.particle_step, P.H, P.W_correc); /* :98 */

    /* AERO::GetYcoef(fvar, diam = particle_step), Var.h:244-266 */
    real const diam = P.particle_step;
#if DIM == 3
    P.aero_L = diam * std::cbrt(3.0 / (4.0 * M_PI));
    P.A_sphere = M_PI * P.aero_L * P.aero_L;
    P.A_plate = diam * diam;
#else
    P.aero_L = diam / std::sqrt(M_PI);
    P.A_sphere = 2 * P.aero_L;
    P.A_plate = diam;
#endif
    P.td = (2.0 * P.rho_rest * std::pow(P.aero_L, DIM - 1)) / (P.tab_Cd * P.mu);
    P.omega = std::sqrt((P.tab_Ck * P.sig) / (P.rho_rest * std::pow(P.aero_L, DIM)) - 1.0 / std::pow(P.td, 2.0));
    P.tmax = -2.0 * (std::atan(std::sqrt(std::pow(P.td * P.omega, 2.0) + 1) + P.td * P.omega) - M_PI) / P.omega;
    P.Cdef = 1.0 - std::exp(-P.tmax / P.td) *
                       (std::cos(P.omega * P.tmax) + 1 / (P.omega * P.td) * std::sin(P.omega * P.tmax));
    P.ycoef = 0.5 * P.Cdef * (P.tab_Cf / (P.tab_Ck * P.tab_Cb)) * (P.rho_g * P.aero_L) / P.sig;

    P.n_full = get_n_full(P.particle_step, P.H); /* IO.cpp:112-115 */
    P.i_n_full = 1.0 / P.n_full;
    P.interp_fac = 1.0 / P.i_interp_fac;
#if DIM == 3
    P.A_plate = P.particle_step * P.particle_step; /* :118-122 */
#else
    P.A_plate = P.particle_step;
#endif
}

/* ------------------------------------------------------------------ neighbours */
/* Neighbours.cpp:7-47 + nanoflann radiusSearch semantics: {j : sum_d (x_i,d - x_j,d)^2 < sr}, self
 * included, payload d2.  Implemented as a cell list (edge slightly above 2H) + exact test; results
 * sorted by ascending j. */
static void update_neighbours(Orc& o, State const& S)
{
    size_t const n = S.n;
    real const sr = o.P.sr;
    o.nb_off.assign(n + 1, 0);
    o.nb_idx.clear();
    o.nb_d2.clear();
    if (n == 0)
        return;
    real const cell = std::sqrt(sr) * (1.0 + 1e-6);
    real lo[DIM], hi[DIM];
    for (int d = 0; d < DIM; ++d)
    {
        lo[d] = S.xi[0][d];
        hi[d] = S.xi[0][d];
    }
    for (size_t i = 1; i < n; ++i)
        for (int d = 0; d < DIM; ++d)
        {
            lo[d] = std::min(lo[d], S.xi[i][d]);
            hi[d] = std::max(hi[d], S.xi[i][d]);
        }
    long nc[3] = {1, 1, 1};
    for (int d = 0; d < DIM; ++d) nc[d] = long(std::floor((hi[d] - lo[d]) / cell)) + 1;
    /* hashed cell list to stay robust for sparse domains */
    size_t const ncell_dense = size_t(nc[0]) * size_t(nc[1]) * size_t(nc[2]);
    bool const dense = ncell_dense <= 8 * n + 1024;
    auto cellof = [&](size_t i, long c[3]) {
        c[0] = c[1] = c[2] = 0;
        for (int d = 0; d < DIM; ++d)
        {
            long k = long(std::floor((S.xi[i][d] - lo[d]) / cell));
            if (k < 0)
                k = 0;
            if (k >= nc[d])
                k = nc[d] - 1;
            c[d] = k;
        }
    };
    auto cid = [&](long const c[3]) { return (size_t(c[2]) * size_t(nc[1]) + size_t(c[1])) * size_t(nc[0]) + size_t(c[0]); };
    std::vector<size_t> key(n);
    for (size_t i = 0; i < n; ++i)
    {
        long c[3];
        cellof(i, c);
        key[i] = cid(c);
    }
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b_) {
        return key[a] < key[b_] || (key[a] == key[b_] && a < b_);
    });
    std::vector<size_t> cstart;
    std::map<size_t, std::pair<size_t, size_t>> sparse;
    if (dense)
    {
        cstart.assign(ncell_dense + 1, 0);
        for (size_t i = 0; i < n; ++i) cstart[key[i] + 1]++;
        for (size_t c = 0; c < ncell_dense; ++c) cstart[c + 1] += cstart[c];
    }
    else
    {
        size_t s = 0;
        while (s < n)
        {
            size_t e = s;
            while (e < n && key[order[e]] == key[order[s]]) e++;
            sparse[key[order[s]]] = std::make_pair(s, e);
            s = e;
        }
    }
    std::vector<std::vector<std::pair<long, real>>> lists(n);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (long ii = 0; ii < long(n); ++ii)
    {
        size_t const i = size_t(ii);
        long c[3];
        cellof(i, c);
        std::vector<std::pair<long, real>>& out = lists[i];
        out.reserve(DIM == 3 ? 280 : 56);
        long const zlo = (DIM == 3) ? -1 : 0, zhi = (DIM == 3) ? 1 : 0;
        for (long dz = zlo; dz <= zhi; ++dz)
            for (long dy = -1; dy <= 1; ++dy)
                for (long dx_ = -1; dx_ <= 1; ++dx_)
                {
                    long cc[3] = {c[0] + dx_, c[1] + dy, c[2] + dz};
                    if (cc[0] < 0 || cc[1] < 0 || cc[2] < 0 || cc[0] >= nc[0] || cc[1] >= nc[1] || cc[2] >= nc[2])
                        continue;
                    size_t const k = cid(cc);
                    size_t s, e;
                    if (dense)
                    {
                        s = cstart[k];
                        e = cstart[k + 1];
                    }
                    else
                    {
                        auto it = sparse.find(k);
                        if (it == sparse.end())
                            continue;
                        s = it->second.first;
                        e = it->second.second;
                    }
                    for (size_t q = s; q < e; ++q)
                    {
                        size_t const j = order[q];
                        /* nanoflann L2_Simple: result += (a[d] - b[d])^2, d ascending */
                        real d2 = 0.0;
                        for (int d = 0; d < DIM; ++d)
                        {
                            real const diff = S.xi[i][d] - S.xi[j][d];
                            d2 += diff * diff;
                        }
                        if (d2 < sr)
                            out.emplace_back(long(j), d2);
                    }
                }
        std::sort(out.begin(), out.end());
    }
    for (size_t i = 0; i < n; ++i) o.nb_off[i + 1] = o.nb_off[i] + long(lists[i].size());
    o.nb_idx.resize(size_t(o.nb_off[n]));
    o.nb_d2.resize(size_t(o.nb_off[n]));
    for (size_t i = 0; i < n; ++i)
    {
        size_t k = size_t(o.nb_off[i]);
        for (auto const& e : lists[i])
        {
            o.nb_idx[k] = e.first;
            o.nb_d2[k] = e.second;
            k++;
        }
    }
}

#define NB_LOOP(ii, k) for (long k = o.nb_off[ii]; k < o.nb_off[(ii) + 1]; ++k)

/* ------------------------------------------------------------------ Shifting.cpp:12-123 */
static void dSPH_PreStep(Orc& o, size_t end, State& S, real& npd)
{
    OrcParams const& P = o.P;
    real npd_ = 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : npd_)
#endif
    for (long ii_ = 0; ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        Mat Lmat_ = mzero(), Lmat_nb_ = mzero();
        Vec gradRho_ = vzero(), norm_ = vzero();
        real kernsum_ = 0.0, colour_ = 0.0;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (ii == jj)
            {
                kernsum_ += P.W_correc;
                continue;
            }
            Vec const Rji = S.xi[jj] - S.xi[ii];
            real const r = std::sqrt(o.nb_d2[k]);
            real const volj = S.m[jj] / S.rho[jj];
            real const kern = Kernel(r, P.H, P.W_correc);
            Vec const Grad = GradK(-Rji, r, P.H, P.W_correc);
            for (int a = 0; a < DIM; ++a)
                for (int c = 0; c < DIM; ++c) Lmat_.a[a][c] -= (volj * Rji[a]) * Grad[c];
            gradRho_ += (volj * (S.rho[jj] - S.rho[ii])) * Grad;
            if (S.b[jj] > ORC_PISTON)
            {
                for (int a = 0; a < DIM; ++a)
                    for (int c = 0; c < DIM; ++c) Lmat_nb_.a[a][c] -= (volj * Rji[a]) * Grad[c];
                norm_ += volj * Grad;
                kernsum_ += kern;
                colour_ += volj * kern;
                npd_ += kern;
            }
        }
        Mat Linv = midentity();
        Mat tmp;
        if (qr_inverse(Lmat_, tmp))
            Linv = tmp;
        gradRho_ = mul(Linv, gradRho_);
        if (S.b[ii] == ORC_BOUND)
        {
            S.lam[ii] = 1.0;
            S.lam_nb[ii] = 1.0;
        }
        else
        {
            S.lam[ii] = min_eigenvalue(Lmat_);
            S.lam_nb[ii] = min_eigenvalue(Lmat_nb_);
        }
        S.L[ii] = Linv;
        S.gradRho[ii] = gradRho_;
        S.norm[ii] = mul(Linv, norm_);
        if (S.lam[ii] > 0.7)
            S.colourG[ii] = 2;
        else
            S.colourG[ii] = 2.0 * std::max(1.0, 1.0 / (2.0 * colour_));
        S.kernsum[ii] = kernsum_;
        S.colour[ii] = colour_;
    }
    /* Q4: the race-free value of npd_ / end (PAIRWISE is the default ST model, VarDefs.h:18-20) */
    npd = npd_ / real(end);
}

/* ------------------------------------------------------------------ Kernel.h:200-244 */
static inline Vec ArtVisc(OrcParams const& P, real rhoi, real rhoj, Vec const& Rji, Vec const& Vji, real idist2,
                          Vec const& gradK)
{
    real const vdotr = dot(Vji, Rji);
    if (vdotr > 0.0)
        return vzero();
    real const muij = P.H * vdotr * idist2;
    real const rhoij = 0.5 * (rhoi + rhoj);
    real const cbar = 0.5 * (std::sqrt((P.B * P.gam) / rhoi) + std::sqrt((P.B * P.gam) / rhoj));
    return gradK * P.visc_alpha * cbar * muij / rhoij;
}

/* ------------------------------------------------------------------ Shifting.cpp:126-186 */
static void dissipation_terms(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
#ifdef _OPENMP
#ifndef ORC_SERIAL_DISSIPATION /* the reference has no pragma here (serial) */
#pragma omp parallel for schedule(static)
#endif
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        Vec artViscI = vzero();
        real Rrhod = 0.0;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (ii == jj)
                continue;
            Vec const Rji = S.xi[jj] - S.xi[ii];
            Vec const Vji = S.v[jj] - S.v[ii];
            real const rr = o.nb_d2[k];
            real const r = std::sqrt(rr);
            real const idist2 = 1.0 / (rr + 0.0001 * P.H_sq);
            real const volj = S.m[jj] / S.rho[jj];
            Vec const gradK = GradK(Rji, r, P.H, P.W_correc);
            if (S.b[jj] > ORC_PISTON)
                artViscI += S.m[jj] * ArtVisc(P, S.rho[ii], S.rho[jj], Rji, Vji, idist2, gradK);
            if (S.b[jj] > ORC_PISTON) /* Kernel.h:200-206 */
                Rrhod += volj * ((S.rho[jj] - S.rho[ii]) + 0.5 * dot(S.gradRho[ii] + S.gradRho[jj], Rji)) *
                         dot(Rji, gradK) * idist2;
            else /* Kernel.h:208-214 */
                Rrhod += volj * (S.rho[jj] - S.rho[ii]) * dot(Rji, gradK) * idist2;
        }
        S.aVisc[ii] = artViscI;
        S.deltaD[ii] = P.dsph_cont * Rrhod;
    }
}

/* ------------------------------------------------------------------ Shifting.cpp:189-290 (ALE) */
static void particle_shift(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        if (S.lam_nb[ii] < 0.55 || S.b[ii] == ORC_BUFFER)
        {
            S.vPert[ii] = vzero();
            continue;
        }
        Vec deltaU = vzero();
        real maxUij = 0.0;
        real woccl = 0.0;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (ii == jj)
                continue;
            Vec const Rji = S.xi[jj] - S.xi[ii];
            real const r = std::sqrt(o.nb_d2[k]);
            real const volj = S.m[jj] / S.rho[jj];
            real const kern = Kernel(r, P.H, P.W_correc);
            Vec const gradK = GradK(Rji, r, P.H, P.W_correc);
            real const kq = kern / P.W_dx;
            real const kq2 = kq * kq;
            deltaU += (1.0 + 0.2 * (kq2 * kq2)) * gradK * volj;
            if (S.b[jj] > ORC_PISTON)
            {
                real theta = std::acos(dot(normalized(S.norm[ii]), normalized(S.norm[jj])));
                if (theta > woccl)
                    woccl = theta;
            }
            real const du = norm(S.v[jj] - S.v[ii]);
            if (du > maxUij)
                maxUij = du;
        }
        deltaU = deltaU * (-2.0 * P.H * norm(S.v[ii]));
        deltaU = std::min(norm(deltaU), std::min(maxUij / 2.0, P.max_shift_vel)) * normalized(deltaU);
        Vec const nrm = -normalized(S.norm[ii]);
        if (S.surfzone[ii] == 0 && S.lam_nb[ii] > 0.55)
        {
            S.vPert[ii] = deltaU;
        }
        else
        {
            if (woccl < M_PI / 12.0)
            {
                /* (I - n n^T) deltaU */
                Vec out;
                for (int a = 0; a < DIM; ++a)
                {
                    real s = 0.0;
                    for (int c = 0; c < DIM; ++c) s += (((a == c) ? 1.0 : 0.0) - nrm[a] * nrm[c]) * deltaU[c];
                    out[a] = s;
                }
                S.vPert[ii] = out;
            }
            else
                S.vPert[ii] = vzero();
        }
    }
}

/* ------------------------------------------------------------------ Geometry.cpp:14-280 */
static void Detect_Surface(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
    std::vector<Vec> norms(end, vzero());
    real const h = 1.33 * P.particle_step;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        Vec nrm = vzero();
        if (S.b[ii] < ORC_PIPE)
        {
            S.surf[ii] = 0;
            S.woccl[ii] = 1;
        }
        else if (S.lam_nb[ii] < 0.2)
        {
            S.surf[ii] = 1;
        }
        else if (S.lam_nb[ii] < 0.75)
        {
            Vec const nhat = normalized(S.norm[ii]);
            Vec const pointT = S.xi[ii] + h * nhat;
            int surf = 1;
            NB_LOOP(ii, k)
            {
                size_t const jj = size_t(o.nb_idx[k]);
                if (jj == ii)
                    continue;
                Vec const x_jT = S.xi[jj] - pointT;
                real const r = std::sqrt(o.nb_d2[k]);
                if (r >= std::sqrt(2.0) * h)
                {
                    if (norm(x_jT) < h)
                    {
                        surf = 0;
                        break;
                    }
                }
                else
                {
#if DIM == 2
                    Vec tau;
                    tau[0] = S.norm[ii][1];
                    tau[1] = -S.norm[ii][0];
                    if ((std::fabs(dot(nhat, x_jT)) + std::fabs(dot(normalized(tau), x_jT))) < h)
                    {
                        surf = 0;
                        break;
                    }
#else
                    Vec const Rij = S.xi[ii] - S.xi[jj];
                    if (std::acos(dot(nhat, (-Rij) / r)) < M_PI / 4.0)
                    {
                        surf = 0;
                        break;
                    }
#endif
                }
            }
            S.surf[ii] = surf;
        }
        else
        {
            S.surf[ii] = 0;
        }

        if (S.lam[ii] > 0.7)
        {
            NB_LOOP(ii, k)
            {
                size_t const jj = size_t(o.nb_idx[k]);
                if (jj == ii)
                    continue;
                Vec const Rij = S.xi[ii] - S.xi[jj];
                real const r = std::sqrt(o.nb_d2[k]);
                real const volj = S.m[jj] / S.rho[jj];
                nrm += (volj * (S.lam[jj] - S.lam[ii])) * GradK(Rij, r, P.H, P.W_correc);
            }
        }
        else
        {
            NB_LOOP(ii, k)
            {
                size_t const jj = size_t(o.nb_idx[k]);
                if (jj == ii)
                    continue;
                Vec const Rij = S.xi[ii] - S.xi[jj];
                real const r = std::sqrt(o.nb_d2[k]);
                real const volj = S.m[jj] / S.rho[jj];
                nrm += (volj * S.lam[jj]) * GradK(Rij, r, P.H, P.W_correc);
            }
        }
        nrm = mul(S.L[ii], nrm);
        if (norm(nrm) > 0.1 * S.lam[ii] / P.H)
            norms[ii] = normalized(nrm);
    }

#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        real woccl_ = 0.0;
        Vec Vdiff = vzero();
        if (S.cellID[ii] != -1 && S.b[ii] == ORC_FREE && P.acase == Gissler)
        {
            if (P.asource == meshInfl)
                Vdiff = S.cellV[ii] - S.v[ii];
            else
            {
                Vec vinf;
                for (int d = 0; d < DIM; ++d) vinf[d] = P.v_inf[d];
                Vdiff = vinf - S.v[ii];
            }
        }
        real curve = 0.0;
        bool const ni_nz = norm(norms[ii]) > 0;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (jj == ii)
                continue;
            Vec const Rij = S.xi[jj] - S.xi[ii]; /* sic: named Rij, is x_j - x_i (Geometry.cpp:179) */
            real const r = std::sqrt(o.nb_d2[k]);
            real const volj = S.m[jj] / S.rho[jj];
            if (ni_nz && norm(norms[jj]) > 0)
                curve += volj * dot(mul(S.L[ii], norms[jj] - norms[ii]), GradK(Rij, r, P.H, P.W_correc));
            if (S.b[ii] == ORC_FREE && P.acase == Gissler)
            {
                real const frac = -dot(Rij, Vdiff) / (norm(Vdiff) * r);
                if (frac > woccl_)
                    woccl_ = frac;
            }
        }
        if (S.lam_nb[ii] < P.lam_cutoff)
            S.woccl[ii] = std::max(0.0, std::min(woccl_, 1.0));
        else
            S.woccl[ii] = 1.0;
        S.norm[ii] = norms[ii];
        S.curve[ii] = curve;
        S.norm_curve[ii] = P.dx * curve;
        S.pDist[ii] = S.lam[ii];
    }

    if (P.ale) /* #if defined(ALE) || defined(TIC), Geometry.cpp:263-277 */
    {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (long ii_ = long(start); ii_ < long(end); ++ii_)
        {
            size_t const ii = size_t(ii_);
            S.surfzone[ii] = 0;
            NB_LOOP(ii, k)
            {
                if (S.surf[size_t(o.nb_idx[k])] == 1)
                {
                    S.surfzone[ii] = 1;
                    break;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ Aero.h:10-98 */
static inline real GetCd(real Re)
{
    return (1.0 + 0.197 * std::pow(Re, 0.63) + 2.6e-04 * std::pow(Re, 1.38)) * (24.0 / (Re + 0.00001));
}
static Vec gissler_force(OrcParams const& P, Vec const& Vdiff, real rho, real press, real mass, real lam,
                         real nneigh, real woccl)
{
    real const Re = 2.0 * rho * norm(Vdiff) * P.aero_L / P.mu_g;
    real frac2;
    if (P.use_lam)
        frac2 = std::min(P.interp_fac * lam, 1.0);
    else
        frac2 = std::min(P.interp_fac * nneigh * P.i_n_full, 1.0);
    real const frac1 = (1.0 - frac2);
    real const Cds = GetCd(Re);
    real Cdl, Adrop;
    if (P.use_TAB_def)
    {
        real ymax = sqnorm(Vdiff) * P.ycoef;
        if (ymax > 1.0)
            ymax = 1.0;
        Cdl = Cds * (1 + 2.632 * ymax);
#if DIM == 3
        Adrop = M_PI * std::pow((P.aero_L + P.tab_Cb * P.aero_L * ymax), 2);
#else
        Adrop = P.A_sphere + 2 * (P.tab_Cb * P.aero_L * ymax);
#endif
    }
    else
    {
        Cdl = Cds;
        Adrop = P.A_sphere;
    }
    real const Cdi = frac1 * Cdl + frac2;
    real const Aunocc = (frac1 * Adrop + frac2 * P.A_plate);
    real const Ai = (1.0 - woccl) * Aunocc;
    return 0.5 * norm(Vdiff) * Vdiff / (P.sos * P.sos) * P.gamma_g * press * Cdi * Ai / mass;
}
/* Aero.h:106-202 */
static Vec induced_pressure(OrcParams const& P, State const& S, size_t ii, Vec const& Vdiff, Vec const& nrm, real lam,
                            real nneigh)
{
    Vec const nh = normalized(nrm);
    real const theta = std::fabs(std::acos(-dot(nh, normalized(Vdiff))));
    real Cp_s, Cp_p, Cp_b, Cp_tot;
    if (theta < 2.4455)
        Cp_s = 1.0 - (2.25) * std::pow(std::sin(theta), 2.0);
    else
        Cp_s = 0.075;
    if (theta < 1.570797)
        Cp_p = std::cos(theta);
    else if (theta < 1.9918)
        Cp_p = -std::pow(std::cos(6.0 * theta + 0.5 * M_PI), 1.5);
    else if (theta < 2.0838)
        Cp_p = 5.5836 * theta - 11.5601;
    else
        Cp_p = 0.075;
    if (theta < 0.7854)
        Cp_b = 1.0;
    else if (theta < 1.570797)
        Cp_b = 0.5 * (std::cos(4.0 * theta - M_PI) + 1.0);
    else
        Cp_b = 0.0;
    real const normCurve = S.norm_curve[ii];
    real const fac1 = 0.25, ifac1 = 1 / fac1;
    if (normCurve < -fac1)
        Cp_tot = Cp_b;
    else if (normCurve < 0.0)
    {
        real const frac = (normCurve + fac1) * ifac1;
        Cp_tot = frac * Cp_b + (1.0 - frac) * Cp_p;
    }
    else if (normCurve < fac1)
    {
        real const frac = (normCurve)*ifac1;
        Cp_tot = frac * Cp_p + (1.0 - frac) * Cp_s;
    }
    else
        Cp_tot = Cp_s;
    real const sos2 = P.sos * P.sos;
    real const Plocali = 0.5 * sqnorm(Vdiff) / sos2 * P.gamma_g * S.cellP[ii] * Cp_tot;
    real const Re = S.cellRho[ii] * norm(Vdiff) * P.aero_L / P.mu_g;
    real const Cdi = GetCd(Re);
    Vec const acc_drop = 0.5 * Vdiff * norm(Vdiff) / sos2 * P.gamma_g * S.cellP[ii] *
                         (M_PI * P.aero_L * P.aero_L * 0.25) * Cdi / S.m[ii];
    Vec const acc_kern = -Plocali * P.A_plate * nh / S.m[ii];
    real const Vnorm = dot(Vdiff, nh);
    Vec const Vpar = Vdiff - Vnorm * nh;
    real const Re_par = S.cellRho[ii] * norm(Vpar) * P.aero_L / P.mu_g;
    real const Cf = 0.027 / std::pow(Re_par + 1e-6, 1.0 / 7.0);
    Vec const acc_skin = 0.5 * norm(Vpar) * Vpar / sos2 * P.gamma_g * S.cellP[ii] * Cf * P.A_plate / S.m[ii];
    real frac1;
    if (P.use_lam)
        frac1 = std::min(P.interp_fac * lam, 1.0);
    else
        frac1 = std::min(P.interp_fac * nneigh * P.i_n_full, 1.0);
    return (frac1 * (acc_kern + acc_skin) + (1.0 - frac1) * acc_drop);
}
/* Aero.h:224-257 */
static Vec skin_friction(OrcParams const& P, State const& S, size_t ii, Vec const& Vdiff, Vec const& nrm, real lam)
{
    Vec const nh = normalized(nrm);
    real const Vnorm = dot(Vdiff, nh);
    if (!(Vnorm > 0.001))
        return vzero();
    real const Re = P.rho_g * norm(Vdiff) * P.aero_L / P.mu_g;
    Vec const acc_press = 0.5 * P.rho_g * Vnorm * Vnorm * P.A_plate * nh / S.m[ii];
    Vec const Vpar = Vdiff - std::fabs(Vnorm) * nh;
    real const Cf = 0.027 / std::pow(Re, 1.0 / 7.0);
    Vec const acc_skin = 0.5 * P.rho_g * norm(Vpar) * Cf * P.A_plate * Vpar / S.m[ii];
    real const frac2 = std::min(1.5 * lam, 1.0);
    real const frac1 = (1.0 - frac2);
    real const Cdi = GetCd(Re);
    Vec const acc_drop = 0.5 * P.rho_g * norm(Vdiff) * Vdiff * (M_PI * P.aero_L * P.aero_L / 4) * Cdi / S.m[ii];
    return frac2 * (acc_press + acc_skin) + frac1 * acc_drop;
}
/* Aero.h:204-263 */
static Vec CalcAeroAcc(OrcParams const& P, State const& S, size_t ii, Vec const& Vdiff, real lam, real nneigh)
{
    if (P.acase == Gissler)
        return gissler_force(P, Vdiff, S.cellRho[ii], S.cellP[ii], S.m[ii], lam, nneigh, S.woccl[ii]);
    if (P.acase == 2) /* InducedPressure */
        return induced_pressure(P, S, ii, Vdiff, S.norm[ii], lam, nneigh);
    if (P.acase == 3) /* SkinFric */
        return skin_friction(P, S, ii, Vdiff, S.norm[ii], lam);
    return vzero();
}

/* ------------------------------------------------------------------ Resid.cpp:21-186 boundaries */
static void Get_Boundary_Pressure(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
    std::vector<real> pressure(end - start, 0.0);
    Vec grav;
    for (int d = 0; d < DIM; ++d) grav[d] = P.grav[d];
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        int isNearSurface = 0;
        real kernsum = 0.0, pkern = 0.0;
        Vec acckern = vzero();
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (S.b[jj] > ORC_PISTON)
            {
                Vec const Rji = S.xi[ii] - S.xi[jj];
                real const r = std::sqrt(o.nb_d2[k]);
                real const volj = S.m[jj] / S.rho[jj];
                real const kern = volj * Kernel(r, P.H, P.W_correc);
                kernsum += kern;
                pkern += S.p[jj] * kern;
                acckern += (kern * S.rho[jj]) * Rji;
                if (S.surfzone[jj])
                    isNearSurface = 1;
            }
        }
        if (kernsum > 0.0)
        {
            real const p = (pkern + dot(grav - S.acc[ii], acckern)) / kernsum;
            if (isNearSurface)
                pressure[ii - start] = std::max(0.0, p);
            else
                pressure[ii - start] = p;
        }
    }
    for (size_t ii = start; ii < end; ++ii)
    {
        S.p[ii] = pressure[ii - start];
        S.rho[ii] = get_density(P, pressure[ii - start]);
    }
}
/* Resid.cpp:78-117: the accumulation target RV_[jj.first] is out of bounds in the reference (UB); the
 * defined part is acc = 0 on the wall block. */
static void Boundary_DBC(Orc&, size_t start, size_t end, State& S)
{
    for (size_t ii = start; ii < end; ++ii) S.acc[ii] = vzero();
}
static void Boundary_Ghost(Orc& o, size_t start, size_t end, State& S, std::vector<int>& near_inlet)
{
    OrcParams const& P = o.P;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        real Rrhoi = 0.0;
        near_inlet[ii] = 1;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (ii == jj)
                continue;
            if (S.b[jj] == ORC_PIPE || S.b[jj] == ORC_FREE)
                near_inlet[ii] = 0;
            Vec const Rji = S.xi[jj] - S.xi[ii];
            Vec const Vji = S.v[jj] - S.v[ii];
            real const r = std::sqrt(o.nb_d2[k]);
            real const volj = S.m[jj] / S.rho[jj];
            Vec const gradK = GradK(Rji, r, P.H, P.W_correc);
            Rrhoi -= volj * dot(Vji, gradK);
        }
        S.Rrho[ii] = Rrhoi * S.rho[ii];
    }
}
static void Set_No_Slip(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
    for (size_t ii = start; ii < end; ++ii)
    {
        Vec velsum = vzero();
        real kernsum = 0.0;
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (ii == jj || S.b[jj] <= ORC_PISTON)
                continue;
            real const r = std::sqrt(o.nb_d2[k]);
            real const kern = Kernel(r, P.H, P.W_correc);
            kernsum += kern;
            velsum += S.v[jj] * kern;
        }
        if (kernsum > 0.0)
            S.v[ii] = 2.0 * S.v[ii] - velsum / kernsum;
    }
}

/* ------------------------------------------------------------------ Resid.cpp:243-469 */
static inline real pairwise_ST_fac(int bA, int bB) /* Kernel.h:88-96 */
{
    if (bA == ORC_BOUND || bB == ORC_BOUND)
    {
        real const contang = 0.5 * M_PI * 7.0 / 9.0;
        return (1.0 + 0.5 * std::cos(contang));
    }
    return 1.0;
}

static void get_acc_and_Rrho(Orc& o, real npd, State& S)
{
    OrcParams const& P = o.P;
    size_t const start = o.bound_points, end = o.total_points;
    real const lam = (6.0 / 81.0 * std::pow((2.0 * P.H), 3.0) / std::pow(M_PI, 4.0) *
                      (9.0 / 4.0 * std::pow(M_PI, 3.0) - 6.0 * M_PI - 4.0)); /* Resid.cpp:437-439 */
    real const npdm2 = (0.5 * P.sig / lam) / (npd * npd);
    real const pi3o4 = 3.0 * M_PI / 4.0;
    Vec grav;
    for (int d = 0; d < DIM; ++d) grav[d] = P.grav[d];
    std::vector<Vec> acc_out(end - start), af_out(end - start);
    std::vector<real> rrho_out(end - start);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long ii_ = long(start); ii_ < long(end); ++ii_)
    {
        size_t const ii = size_t(ii_);
        Vec acc_ = vzero(), acc_ale_ = vzero(), visc_ = vzero(), surf_t_ = vzero();
        real Rrhoc_ = 0.0, Rrho_ = 0.0;
        Vec af = S.Af[ii]; /* acc_aero_i is only assigned inside the branch (Resid.cpp:267-277) */
        if (S.cellID[ii] != -1)
        {
            Vec const V_diff = S.cellV[ii] - S.v[ii];
            Vec const aero =
                CalcAeroAcc(P, S, ii, V_diff, S.lam_nb[ii], real(o.nb_off[ii + 1] - o.nb_off[ii]));
            acc_ += aero;
            af = aero;
        }
        real const pi_rho2 = S.p[ii] / (S.rho[ii] * S.rho[ii]);
        NB_LOOP(ii, k)
        {
            size_t const jj = size_t(o.nb_idx[k]);
            if (S.part_id[ii] == S.part_id[jj]) /* Q2 */
                continue;
            Vec const Rji = S.xi[jj] - S.xi[ii];
            Vec const Vji = S.v[jj] - S.v[ii];
            real const rr = o.nb_d2[k];
            real const r = std::sqrt(rr);
            real const idist2 = 1.0 / (rr + 0.001 * P.H_sq);
            real const volj = S.m[jj] / S.rho[jj];
            Vec const gradK = GradK(Rji, r, P.H, P.W_correc);
            /* BasePos, Kernel.h:153-157 */
            Vec const contrib = S.m[jj] * gradK * (pi_rho2 + S.p[jj] / (S.rho[jj] * S.rho[jj]));
            /* Viscosity, Kernel.h:262-269 */
            Vec const visc =
                (P.nu * (S.rho[ii] + S.rho[jj]) / (S.rho[ii] * S.rho[jj]) * dot(Rji, gradK) * idist2) * Vji;
            visc_ += S.m[jj] * visc;
            /* SurfTenContrib, Resid.cpp:212-240, Kernel.h:101-113 */
            if (!P.ale || S.surfzone[ii] == 1)
            {
                real const fac = pairwise_ST_fac(S.b[ii], S.b[jj]);
                surf_t_ += (-npdm2 * fac * std::cos(pi3o4 * r / P.H)) * (Rji / r);
            }
            acc_ -= contrib;
            if (P.ale)
            {
                /* ALEMomentum, Kernel.h:179-185 */
                Vec const& vi = S.v[ii];
                Vec const& vj = S.v[jj];
                Vec const& pi_ = S.vPert[ii];
                Vec const& pj_ = S.vPert[jj];
                real const pjg = dot(pj_, gradK), pig = dot(pi_, gradK);
                real const dpg = dot(pj_ - pi_, gradK);
                Vec ale;
                for (int a = 0; a < DIM; ++a) ale[a] = ((vj[a] * pjg + vi[a] * pig) - vi[a] * dpg) * volj;
                acc_ale_ += ale;
                /* ALEContinuity / ALECont2ndterm, Kernel.h:187-196 */
                Rrho_ -= dot((vj + pj_) - (vi + pi_), gradK) * volj;
                Rrhoc_ += dot(S.rho[jj] * pj_ + S.rho[ii] * pi_, gradK) * volj;
            }
            else
            {
                Rrho_ -= volj * dot(Vji, gradK);
            }
        }
        if (S.internal[ii] == 1)
        {
            /* NormalBoundaryRepulsion, Kernel.h:272-277 */
            real const beta = 4 * P.speed_sound * P.speed_sound;
            real const kern = BoundaryKernel(S.y[ii], P.H, beta);
            acc_ += (P.bnd_mass / (P.bnd_mass + P.sim_mass) * kern) * S.bNorm[ii];
        }
        /* frozen-dissipation variant, Resid.cpp:401-411 */
        if (P.ale)
        {
            acc_out[ii - start] = (acc_ + acc_ale_ + S.aVisc[ii] + visc_ + surf_t_ / S.m[ii] + grav);
            rrho_out[ii - start] = Rrho_ * S.rho[ii] + Rrhoc_ + S.deltaD[ii];
        }
        else
        {
            acc_out[ii - start] = (acc_ + S.aVisc[ii] + visc_ + surf_t_ / S.m[ii] + grav);
            rrho_out[ii - start] = Rrho_ * S.rho[ii] + S.deltaD[ii];
        }
        af_out[ii - start] = af;
    }
    /* the reference writes pnp1[ii].acc in place while other threads read only xi,v,rho,p,m,vPert of
     * neighbours, so deferred assignment is equivalent */
    for (size_t ii = start; ii < end; ++ii)
    {
        S.acc[ii] = acc_out[ii - start];
        S.Af[ii] = af_out[ii - start];
        S.Rrho[ii] = rrho_out[ii - start];
    }
}

/* ------------------------------------------------------------------ Resid.cpp:471-612 (constVel) */
#if DIM == 3
/* Geometry.cpp:485-575.  Only the SIGN of the 4x4 determinants |p 1; a 1; b 1; c 1| is used; expanding along the
 * last column gives det4 = -det3[(a-p); (b-p); (c-p)], which is how it is evaluated here (Eigen's cofactor
 * formula is not available; the two agree in sign except within rounding of a degenerate configuration).
 * Q6: only face[0..2], the plane of the first three vertices and the edges (last,0),(0,1),(1,2) are tested. */
static inline real det4_sign_arg(Vec const& p, Vec const& a, Vec const& b, Vec const& c)
{
    Vec const u = a - p, v = b - p, w = c - p;
    real const det3 = u[0] * (v[1] * w[2] - v[2] * w[1]) - u[1] * (v[0] * w[2] - v[2] * w[0]) +
                      u[2] * (v[0] * w[1] - v[1] * w[0]);
    return -det3;
}
static int Crossings3D(Mesh const& M, std::vector<long> const& face, Vec const& testp, Vec const& rayp)
{
    Vec const &f0 = M.verts[face[0]], &f1 = M.verts[face[1]], &f2 = M.verts[face[2]];
    int const flag1 = det4_sign_arg(testp, f0, f1, f2) < 0.0;
    int const flag2 = det4_sign_arg(rayp, f0, f1, f2) < 0.0;
    if (flag1 != flag2)
    {
        Vec vtx0 = M.verts[face.back()], vtx1 = f0;
        /* rows: testp, vtx0, vtx1, rayp */
        int const flag3 = det4_sign_arg(testp, vtx0, vtx1, rayp) < 0.0;
        for (size_t ii = 1; ii < 3; ++ii)
        {
            vtx0 = vtx1;
            vtx1 = M.verts[face[ii]];
            int const flag4 = det4_sign_arg(testp, vtx0, vtx1, rayp) < 0.0;
            if (flag4 != flag3)
                return 0;
        }
        return 1;
    }
    return 0;
}
/* the containment test of one face (ray along +x from the point) and the boundary test (segment point -> cell centre) */
static inline int face_contains_ray(Mesh const& M, std::vector<long> const& face, Vec const& testp)
{
    Vec rayp = testp;
    rayp[0] += 1e+5; /* Containment.cpp:388-390 */
    return Crossings3D(M, face, testp, rayp);
}
static inline int face_cut_by_segment(Mesh const& M, std::vector<long> const& face, Vec const& testp, Vec const& rayp)
{
    return Crossings3D(M, face, testp, rayp);
}
#else
/* Crossings2D, Geometry.cpp:354-399: does the +x ray from the point cross this edge (Haines' crossings test, one edge) */
static int Crossings2D(Mesh const& M, std::vector<long> const& edge, Vec const& point)
{
    real const tx = point[0], ty = point[1];
    Vec const &vtx0 = M.verts[edge[0]], &vtx1 = M.verts[edge[1]];
    int const yflag0 = (vtx0[1] >= ty), yflag1 = (vtx1[1] >= ty);
    int inside_flag = 0;
    if (yflag0 != yflag1)
        if (((vtx1[1] - ty) * (vtx1[0] - vtx0[0]) >= (vtx1[0] - tx) * (vtx1[1] - vtx0[1])) == yflag1)
            inside_flag = !inside_flag;
    return inside_flag;
}
/* get_line_intersection, Geometry.cpp:312-341: segment p1 -> cellC against the edge; a denominator below MEPSILON --
 * collinear, or any NEGATIVE one -- is "no intersection", as in the reference */
static int get_line_intersection(Mesh const& M, std::vector<long> const& edge, Vec const& p1, Vec const& cellC)
{
    Vec const &e1 = M.verts[edge[0]], &e2 = M.verts[edge[1]];
    Vec const s_ = cellC - p1, r = e2 - e1;
    real const denom = (-r[0] * s_[1] + s_[0] * r[1]);
    if (denom < MEPS)
        return 0;
    real const u = (-s_[1] * (p1[0] - e1[0]) + s_[0] * (p1[1] - e1[1])) / denom;
    real const t = (r[0] * (p1[1] - e1[1]) - r[1] * (p1[0] - e1[0])) / denom;
    return (u > 0 && u < 1 && t > 0 && t < 1) ? 1 : 0;
}
static inline int face_contains_ray(Mesh const& M, std::vector<long> const& face, Vec const& testp)
{
    return Crossings2D(M, face, testp);
}
static inline int face_cut_by_segment(Mesh const& M, std::vector<long> const& face, Vec const& testp, Vec const& rayp)
{
    return get_line_intersection(M, face, testp, rayp);
}
#endif
/* Containment.cpp:385-420; Q7: a negative (or out of range) cell is "not contained" */
static unsigned CheckCell(Mesh const& M, long cell, Vec const& testp)
{
    if (cell < 0 || size_t(cell) >= M.size())
        return 0;
    unsigned line_flag = 0, inside_flag = 0;
    for (long f : M.cFaces[size_t(cell)])
        if (face_contains_ray(M, M.faces[size_t(f)], testp))
        {
            inside_flag = !inside_flag;
            if (line_flag)
                break; /* convex assumption */
            line_flag = 1;
        }
    return inside_flag;
}
/* nanoflann KNNResultSet on the cell centres: the k nearest in ascending distance; ties by cell index */
static std::vector<long> nearest_cells(Mesh const& M, Vec const& p, size_t k)
{
    std::vector<std::pair<real, long>> d(M.size());
    for (size_t c = 0; c < M.size(); ++c)
    {
        real d2 = 0.0;
        for (int a = 0; a < DIM; ++a) d2 += (p[a] - M.cCentre[c][a]) * (p[a] - M.cCentre[c][a]);
        d[c] = std::make_pair(d2, long(c));
    }
    k = std::min(k, d.size());
    std::partial_sort(d.begin(), d.begin() + long(k), d.end());
    std::vector<long> out(k);
    for (size_t i = 0; i < k; ++i) out[i] = d[i].second;
    return out;
}
static void take_cell(Mesh const& M, State& S, size_t ii, long cell)
{
    S.cellID[ii] = cell;
    S.cellV[ii] = M.cVel[size_t(cell)];
    S.cellP[ii] = M.cP[size_t(cell)];
    S.cellRho[ii] = M.cRho[size_t(cell)];
}
/* FindCell, Containment.cpp:579-820.  Returns the particles to delete (each once: the reference can list a
 * particle several times, which would make it erase the wrong ones). */
static std::vector<size_t> FindCell(Orc& o, State& pnp1)
{
    Mesh const& M = o.cells;
    std::vector<size_t> toDelete;
    for (size_t ii = o.bound_points; ii < o.total_points; ++ii)
    {
        if (pnp1.b[ii] != ORC_FREE || pnp1.lam_nb[ii] > o.P.lam_cutoff)
        {
            pnp1.cellID[ii] = c_no_cell;
            continue;
        }
        Vec const testp = pnp1.xi[ii];
        if (CheckCell(M, pnp1.cellID[ii], testp))
        {
            take_cell(M, pnp1, ii, pnp1.cellID[ii]);
            pnp1.ipt_n_failed[ii] = 0;
            continue;
        }
        bool inside = false;
        std::vector<long> ret;
        for (size_t num_results : {size_t(5), size_t(500)})
        {
            ret = nearest_cells(M, testp, num_results);
            for (long cell : ret)
                if (CheckCell(M, cell, testp))
                {
                    take_cell(M, pnp1, ii, cell);
                    pnp1.ipt_n_failed[ii] = 0;
                    pnp1.internal[ii] = 0;
                    inside = true;
                    break;
                }
            if (inside)
                break;
        }
        if (inside)
            continue;
        /* across a boundary?  ray from the point to each of the 500 nearest cell centres */
        unsigned cross = 0;
        bool del = false;
        for (long index : ret)
        {
            Vec const rayp = M.cCentre[size_t(index)];
            for (long findex : M.cFaces[size_t(index)])
                if (M.leftright[size_t(findex)].second < 0)
                    if (face_cut_by_segment(M, M.faces[size_t(findex)], testp, rayp))
                    {
                        cross = !cross;
                        if (M.leftright[size_t(findex)].second == -1)
                        {
                            pnp1.internal[ii] = 1;
                            break;
                        }
                        else if (M.leftright[size_t(findex)].second == -2)
                        {
                            del = true;
                            break;
                        }
                    }
        }
        if (cross == 0)
        {
            if (pnp1.ipt_n_failed[ii] > 10)
                del = true;
            else
                pnp1.ipt_n_failed[ii]++;
        }
        if (del)
            toDelete.push_back(ii);
    }
    return toDelete;
}
/* FirstCell, Containment.cpp:425-573: PIPE -> FREE transition, no previous cell */
static void FirstCell(Orc& o, State& S, size_t ii, unsigned& to_del)
{
    Mesh const& M = o.cells;
    Vec const testp = S.xi[ii];
    std::vector<long> const ret = nearest_cells(M, testp, DIM == 3 ? 150 : 20); /* Containment.cpp:431-436 */
    for (long cell : ret)
        if (CheckCell(M, cell, testp))
        {
            take_cell(M, S, ii, cell);
            return;
        }
    unsigned cross = 0;
    for (long index : ret)
    {
        Vec const rayp = M.cCentre[size_t(index)];
        for (long findex : M.cFaces[size_t(index)])
            if (M.leftright[size_t(findex)].second < 0)
            {
                std::vector<long> const& face = M.faces[size_t(findex)];
                if (face_cut_by_segment(M, face, testp, rayp))
                {
                    cross = !cross;
                    if (M.leftright[size_t(findex)].second == -1)
                    {
#if DIM == 3
                        Vec const r1 = M.verts[face[1]] - M.verts[face[0]], r2 = M.verts[face[2]] - M.verts[face[0]];
                        Vec nrm = normalized(cross3(r1, r2));
#else
                        Vec const r1 = M.verts[face[1]] - M.verts[face[0]]; /* Containment.cpp:540-544 */
                        Vec nrm;
                        nrm[0] = -r1[1];
                        nrm[1] = r1[0];
                        nrm = normalized(nrm);
#endif
                        S.v[ii] = S.v[ii] - (2 * dot(S.v[ii], nrm)) * nrm;
                        real const plane = dot(nrm, M.verts[face[1]]);
                        real const dist = (plane - dot(S.xi[ii], nrm)) / dot(nrm, nrm);
                        S.xi[ii] = S.xi[ii] + dist * nrm;
                    }
                    else if (M.leftright[size_t(findex)].second == -2)
                        to_del = 1;
                }
            }
    }
    if (cross == 0)
        o.first_cell_errors++; /* the reference prints and calls exit(-1) here */
}

static void dSPH_PreStep(Orc& o, size_t end, State& S, real& npd);
static void update_neighbours(Orc& o, State const& S);

/* Resid.cpp:471-523: meshInfl.  Escaped particles are erased from both time levels, then the neighbour list and
 * the prestep are redone. */
static void get_aero_velocity_mesh(Orc& o, State& pn, State& pnp1, real& npd)
{
    std::vector<size_t> toDelete = FindCell(o, pnp1);
    if (toDelete.empty())
        return;
    std::sort(toDelete.begin(), toDelete.end());
    size_t const nDel = toDelete.size();
    for (auto it = toDelete.rbegin(); it != toDelete.rend(); ++it)
    {
        pnp1.erase(*it);
        pn.erase(*it);
    }
    /* block ranges: the reference only shifts back/buffer (by nDel, whatever their block) and leaves
     * limits[].index stale; the contract keeps the ranges consistent with the erased particles */
    for (Block& B : o.limits)
    {
        long before_first = 0, before_second = 0;
        for (size_t d : toDelete)
        {
            before_first += long(d) < B.first;
            before_second += long(d) < B.second;
        }
        B.first -= before_first;
        B.second -= before_second;
        for (long& x : B.back) x -= long(nDel);
        for (auto& buf : B.buffer)
            for (long& x : buf) x -= long(nDel);
    }
    o.delete_count += nDel;
    o.fluid_points -= nDel;
    o.total_points -= nDel;
    o.end_index -= nDel;
    update_neighbours(o, pnp1);
    dSPH_PreStep(o, o.total_points, pnp1, npd);
}

static void get_aero_velocity(Orc& o, size_t start, size_t end, State& S)
{
    OrcParams const& P = o.P;
    if (P.asource != constVel)
        return; /* meshInfl goes through get_aero_velocity_mesh (needs both time levels) */
    for (size_t ii = start; ii < end; ++ii)
    {
        bool cond;
        if (P.use_lam)
            cond = S.lam_nb[ii] < P.lam_cutoff;
        else
            cond = real(o.nb_off[ii + 1] - o.nb_off[ii]) * P.i_n_full < P.lam_cutoff;
        if (cond && S.b[ii] == ORC_FREE)
        {
            for (int d = 0; d < DIM; ++d) S.cellV[ii][d] = P.v_inf[d];
            S.cellID[ii] = 1;
        }
        else
        {
            S.cellV[ii] = vzero();
            S.cellID[ii] = c_no_cell;
        }
    }
}

/* Containment.cpp:822-890, constVel part: PIPE -> FREE past the aero plane. */
static void Check_Pipe_Outlet(Orc& o, State& S)
{
    for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
    {
        Block const& B = o.limits[block];
        for (long ii = B.first; ii < B.second; ++ii)
            if (S.b[ii] == ORC_PIPE)
                if (dot(S.xi[ii], B.aero_norm) > B.aeroconst)
                {
                    S.b[ii] = ORC_FREE;
                    if (S.lam_nb[ii] < o.P.lam_cutoff && o.P.asource == meshInfl)
                    {
                        unsigned to_del = 0;
                        FirstCell(o, S, size_t(ii), to_del);
                        if (to_del)
                            o.pipe_outlet_del.push_back(size_t(ii));
                    }
                }
    }
    /* Containment.cpp:849-890: erase the particles that left through an outer boundary (pnp1 only) */
    if (!o.pipe_outlet_del.empty())
    {
        std::vector<size_t>& del = o.pipe_outlet_del;
        std::sort(del.begin(), del.end());
        for (auto it = del.rbegin(); it != del.rend(); ++it)
        {
            S.erase(*it);
            if (&S == &o.pnp1)
                o.pn.erase(*it); /* the reference erases pnp1 only and leaves pn misaligned; the contract keeps
                                    the two time levels index-aligned */
            o.total_points--;
            o.fluid_points--;
            o.end_index--;
            o.delete_count++;
        }
        for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
        {
            Block& B = o.limits[block];
            long before_first = 0, before_second = 0;
            for (size_t d : del)
            {
                before_first += long(d) < B.first;
                before_second += long(d) < B.second;
            }
            B.first -= before_first;
            B.second -= before_second;
            for (long& x : B.back) x -= before_second;
            for (auto& buf : B.buffer)
                for (long& x : buf) x -= before_second;
        }
        del.clear();
        update_neighbours(o, S); /* the reference carries on with a stale, index-based list; contract: rebuilt */
    }
}

/* ------------------------------------------------------------------ wall treatment shared by NB and RK */
static Vec block_velocity(Orc& o, Block const& B, bool nb_comparator)
{
    if (B.nTimes != 0)
    {
        Vec vel = vzero();
        for (size_t t = 0; t < B.nTimes; ++t)
        {
            /* Q10: NB uses current_time > times[t] (Newmark_Beta.cpp:77), RK times[t] > current_time
             * (Runge_Kutta.cpp:45) */
            bool const take = nb_comparator ? (o.P.current_time > B.times[t]) : (B.times[t] > o.P.current_time);
            if (take)
                vel = B.vels[t];
        }
        return vel;
    }
    return B.vels[0];
}

/* ------------------------------------------------------------------ Newmark_Beta.cpp:54-301 */
static void Do_NB_Iter(Orc& o, real npd, State const& pn, State& pnp1)
{
    OrcParams const& P = o.P;
    std::vector<int> near_inlet(o.bound_points, 0);
    real const gamma_t1 = P.nb_gamma, gamma_t2 = 1 - gamma_t1;
    real const beta_t1 = P.nb_beta, beta_t2 = 0.5 * (1 - 2 * beta_t1);

    for (size_t block = 0; block < o.n_bound_blocks; ++block)
    {
        Block const& B = o.limits[block];
        Vec const vel = block_velocity(o, B, true);
        for (long jj = B.first; jj < B.second; ++jj) pnp1.v[jj] = vel;
        if (B.no_slip)
            Set_No_Slip(o, B.first, B.second, pnp1);
        switch (B.bound_solver)
        {
        case DBC: Boundary_DBC(o, B.first, B.second, pnp1); break;
        case pressure_G: Get_Boundary_Pressure(o, B.first, B.second, pnp1); break;
        case ghost: Boundary_Ghost(o, B.first, B.second, pnp1, near_inlet); break;
        default: break;
        }
    }

    get_acc_and_Rrho(o, npd, pnp1);

    real const dt = P.delta_t;
    real const dt2 = dt * dt;
    auto clamp_rho = [&](real lo, real val) { return std::max(lo, std::min(P.rho_max, val)); };

    for (size_t block = 0; block < o.n_bound_blocks; ++block)
    {
        Block const& B = o.limits[block];
        if (B.bound_solver == DBC)
        {
            for (long ii = B.first; ii < B.second; ++ii)
            {
                real const rho =
                    clamp_rho(P.rho_min, pn.rho[ii] + dt * (gamma_t1 * pnp1.Rrho[ii] + gamma_t2 * pn.Rrho[ii]));
                pnp1.rho[ii] = rho;
                pnp1.p[ii] = get_pressure(P, rho);
            }
        }
        else if (B.bound_solver == ghost)
        {
            for (long ii = B.first; ii < B.second; ++ii)
            {
                real const lo = near_inlet[ii] ? P.rho_rest : P.rho_min;
                real const rho = clamp_rho(lo, pn.rho[ii] + dt * (gamma_t1 * pnp1.Rrho[ii] + gamma_t2 * pn.Rrho[ii]));
                pnp1.rho[ii] = rho;
                pnp1.p[ii] = get_pressure(P, rho);
                if (near_inlet[ii])
                    pnp1.Rrho[ii] = std::fmax(0.0, pnp1.Rrho[ii]);
            }
        }
    }

    for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
    {
        Block const& B = o.limits[block];
        for (long ii = B.first; ii < B.second; ++ii)
        {
            if (pnp1.b[ii] > ORC_BUFFER && pnp1.b[ii] != ORC_OUTLET)
            {
                if (P.ale)
                    pnp1.xi[ii] = pn.xi[ii] + dt * (pn.v[ii] + pnp1.vPert[ii]) +
                                  dt2 * (beta_t1 * pnp1.acc[ii] + beta_t2 * pn.acc[ii]);
                else
                    pnp1.xi[ii] = pn.xi[ii] + dt * pn.v[ii] + dt2 * (beta_t2 * pn.acc[ii] + beta_t1 * pnp1.acc[ii]);
                pnp1.v[ii] = pn.v[ii] + dt * (gamma_t1 * pnp1.acc[ii] + gamma_t2 * pn.acc[ii]);
                real const rho =
                    clamp_rho(P.rho_min, pn.rho[ii] + dt * (gamma_t1 * pnp1.Rrho[ii] + gamma_t2 * pn.Rrho[ii]));
                pnp1.rho[ii] = rho;
                pnp1.p[ii] = get_pressure(P, rho);
            }
            else if (pnp1.b[ii] == ORC_OUTLET)
            {
                pnp1.xi[ii] = pn.xi[ii] + dt * pnp1.v[ii];
            }
        }
        if (B.block_type == inletZone)
        {
            if (B.fixed_vel_or_dynamic == 1)
            {
                Vec const unorm = normalized(B.insert_norm);
                for (size_t ii = 0; ii < B.back.size(); ++ii)
                {
                    long const backID = B.back[ii];
                    Vec const xi = pnp1.xi[backID];
                    for (size_t jj = 0; jj < B.buffer[ii].size(); ++jj)
                    {
                        long const buffID = B.buffer[ii][jj];
                        pnp1.xi[buffID] = xi - (P.dx * (jj + 1.0)) * unorm;
                        pnp1.v[buffID] = pnp1.v[backID];
                        pnp1.rho[buffID] = pnp1.rho[backID];
                        pnp1.p[buffID] = pnp1.p[backID];
                    }
                }
            }
            else
            {
                for (size_t ii = 0; ii < B.back.size(); ++ii)
                    for (size_t jj = 0; jj < B.buffer[ii].size(); ++jj)
                    {
                        long const buffID = B.buffer[ii][jj];
                        real const rho = clamp_rho(
                            P.rho_min, pn.rho[buffID] + dt * (gamma_t1 * pnp1.Rrho[buffID] + gamma_t2 * pn.Rrho[buffID])
                        );
                        pnp1.rho[buffID] = rho;
                        pnp1.p[buffID] = get_pressure(P, rho);
                        pnp1.xi[buffID] = pn.xi[buffID] + dt * pn.v[buffID];
                    }
            }
        }
    }
}

/* Newmark_Beta.cpp:10-52. Returns -1 (restart with dt/2), 0 (continue), 1 (sub-iterations exceeded). */
static int NB_Check_Error(Orc& o, real& rms_error, real& logbase, std::vector<Vec> const& xih, State const& pn,
                          State& pnp1, unsigned& iteration)
{
    size_t const start = o.start_index, end = o.end_index;
    real errsum = 0.0;
    for (size_t ii = start; ii < end; ++ii)
    {
        Vec const r = pnp1.xi[ii] - xih[ii - start];
        errsum += sqnorm(r);
    }
    real const log_error = std::log10(std::sqrt(errsum / real(end - start)));
    if (iteration == 0)
        logbase = log_error;
    rms_error = log_error - logbase;
    if (iteration > unsigned(o.P.max_subits))
    {
        if (rms_error > 0.0)
        {
            pnp1 = pn;
            update_neighbours(o, pnp1);
            o.P.delta_t = 0.5 * o.P.delta_t;
            iteration = 0;
            rms_error = 0.0;
            return -1;
        }
        return 1;
    }
    return 0;
}

/* Newmark_Beta.cpp:303-331 */
static real NB_Solve(Orc& o, real npd, real& logbase, unsigned& iteration, std::vector<Vec>& xih, State const& pn,
                     State& pnp1)
{
    real rms_error = 0.0;
    while (rms_error > o.P.min_residual)
    {
        for (size_t ii = o.start_index; ii < o.end_index; ++ii) xih[ii - o.start_index] = pnp1.xi[ii];
        Do_NB_Iter(o, npd, pn, pnp1);
        int const errstate = NB_Check_Error(o, rms_error, logbase, xih, pn, pnp1, iteration);
        if (errstate == 0)
            iteration++;
        else if (errstate == 1)
            break;
    }
    return rms_error;
}

/* ------------------------------------------------------------------ Runge_Kutta.cpp */
static real Check_RK_Error(Orc& o, real logbase, State const& a, State const& b_)
{
    real errsum = 0.0;
    for (size_t ii = o.start_index; ii < o.end_index; ++ii) errsum += sqnorm(b_.xi[ii] - a.xi[ii]);
    return std::log10(std::sqrt(errsum / real(o.end_index - o.start_index))) - logbase;
}

/* wall part common to intermediate (final=false) and final steps, Runge_Kutta.cpp:36-135 / 244-351 */
static void RK_walls(Orc& o, State const& part_n, State& S, bool final_step, real dt, State const* st_1,
                     State const* st_2, State const* st_3)
{
    OrcParams const& P = o.P;
    auto newrho = [&](long ii) {
        if (!final_step)
            return part_n.rho[ii] + dt * S.Rrho[ii];
        return part_n.rho[ii] +
               (dt / 6.0) * (part_n.Rrho[ii] + 2.0 * st_1->Rrho[ii] + 2.0 * st_2->Rrho[ii] + st_3->Rrho[ii]);
    };
    for (size_t block = 0; block < o.n_bound_blocks; ++block)
    {
        Block const& B = o.limits[block];
        Vec const vel = block_velocity(o, B, false);
        for (long jj = B.first; jj < B.second; ++jj) S.v[jj] = vel;
        if (B.no_slip)
            Set_No_Slip(o, B.first, B.second, S);
        switch (B.bound_solver)
        {
        case DBC:
        {
            Boundary_DBC(o, B.first, B.second, S);
            for (long ii = B.first; ii < B.second; ++ii)
            {
                real const rho = std::max(P.rho_min, std::min(P.rho_max, newrho(ii)));
                S.rho[ii] = rho;
                S.p[ii] = get_pressure(P, rho);
                if (final_step)
                    S.Rrho[ii] = st_3->Rrho[ii];
            }
            break;
        }
        case pressure_G: Get_Boundary_Pressure(o, B.first, B.second, S); break;
        case ghost:
        {
            std::vector<int> near_inlet(o.bound_points, 0);
            Boundary_Ghost(o, B.first, B.second, S, near_inlet);
            for (long ii = B.first; ii < B.second; ++ii)
            {
                real const lo = near_inlet[ii] ? P.rho_rest : P.rho_min;
                real const rho = std::max(lo, std::min(P.rho_max, newrho(ii)));
                S.rho[ii] = rho;
                S.p[ii] = get_pressure(P, rho);
                if (near_inlet[ii])
                    S.Rrho[ii] = std::fmax(0.0, S.Rrho[ii]);
            }
            break;
        }
        default: break;
        }
    }
}

static void RK_inlet_buffers(Orc& o, Block const& B, State const& part_n, State& S, real dt, bool final_step,
                             State const* st_1, State const* st_2, State const* st_3)
{
    OrcParams const& P = o.P;
    if (B.block_type != inletZone)
        return;
    if (B.fixed_vel_or_dynamic == 1)
    {
        Vec const unorm = normalized(B.insert_norm);
        for (size_t ii = 0; ii < B.back.size(); ++ii)
        {
            long const backID = B.back[ii];
            Vec const xi = S.xi[backID];
            for (size_t jj = 0; jj < B.buffer[ii].size(); ++jj)
            {
                long const buffID = B.buffer[ii][jj];
                S.xi[buffID] = xi - (P.dx * (jj + 1.0)) * unorm;
                S.v[buffID] = S.v[backID];
                S.rho[buffID] = S.rho[backID];
                S.p[buffID] = S.p[backID];
            }
        }
    }
    else
    {
        for (size_t ii = 0; ii < B.back.size(); ++ii)
            for (size_t jj = 0; jj < B.buffer[ii].size(); ++jj)
            {
                long const buffID = B.buffer[ii][jj]; /* reference indexes limits[jj] here (UB), see header */
                real val;
                if (!final_step)
                    val = part_n.rho[buffID] + dt * S.Rrho[buffID];
                else /* sic: Rrho terms indexed by ii, Runge_Kutta.cpp:439-441 */
                    val = part_n.rho[buffID] + (dt / 6.0) * (part_n.Rrho[ii] + 2.0 * st_1->Rrho[ii] +
                                                             2.0 * st_2->Rrho[ii] + st_3->Rrho[ii]);
                real const rho = std::max(P.rho_min, std::min(P.rho_max, val));
                S.rho[buffID] = rho;
                S.p[buffID] = get_pressure(P, rho);
                S.xi[buffID] = part_n.xi[buffID] + dt * part_n.v[buffID];
            }
    }
}

/* Runge_Kutta.cpp:28-232 */
static State RK_intermediate(Orc& o, real npd, State const& part_n, State const& part_prev, real dt_inter)
{
    OrcParams const& P = o.P;
    State S = part_prev;
    RK_walls(o, part_n, S, false, dt_inter, nullptr, nullptr, nullptr);
    get_acc_and_Rrho(o, npd, S);
    for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
    {
        Block const& B = o.limits[block];
        for (long ii = B.first; ii < B.second; ++ii)
        {
            if (S.b[ii] > ORC_BUFFER)
            {
                if (P.ale)
                    S.xi[ii] = part_n.xi[ii] + dt_inter * (S.v[ii] + S.vPert[ii]);
                else
                    S.xi[ii] = part_n.xi[ii] + dt_inter * S.v[ii];
                S.v[ii] = part_n.v[ii] + dt_inter * S.acc[ii];
                real const rho = std::max(P.rho_min, std::min(P.rho_max, part_n.rho[ii] + dt_inter * S.Rrho[ii]));
                S.rho[ii] = rho;
                S.p[ii] = get_pressure(P, rho);
            }
        }
        RK_inlet_buffers(o, B, part_n, S, dt_inter, false, nullptr, nullptr, nullptr);
    }
    return S;
}

/* Runge_Kutta.cpp:234-457 */
static State RK_final(Orc& o, real npd, real dt, State const& part_n, State const& st_1, State const& st_2,
                      State const& st_3)
{
    OrcParams const& P = o.P;
    State S = st_3;
    RK_walls(o, part_n, S, true, dt, &st_1, &st_2, &st_3);
    get_acc_and_Rrho(o, npd, S);
    for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
    {
        Block const& B = o.limits[block];
        for (long ii = B.first; ii < B.second; ++ii)
        {
            if (S.b[ii] > ORC_BUFFER && S.b[ii] != ORC_OUTLET)
            {
                if (P.ale)
                    S.xi[ii] = part_n.xi[ii] +
                               (dt / 6.0) * ((part_n.v[ii] + part_n.vPert[ii]) + 2.0 * (st_1.v[ii] + st_1.vPert[ii]) +
                                             2.0 * (st_2.v[ii] + st_2.vPert[ii]) + (st_3.v[ii] + st_3.vPert[ii]));
                else
                    S.xi[ii] = part_n.xi[ii] +
                               (dt / 6.0) * (part_n.v[ii] + 2.0 * st_1.v[ii] + 2.0 * st_2.v[ii] + st_3.v[ii]);
                S.v[ii] = part_n.v[ii] +
                          (dt / 6.0) * (part_n.acc[ii] + 2.0 * st_1.acc[ii] + 2.0 * st_2.acc[ii] + st_3.acc[ii]);
                real const rho = std::max(
                    P.rho_min,
                    std::min(P.rho_max, part_n.rho[ii] + (dt / 6.0) * (part_n.Rrho[ii] + 2.0 * st_1.Rrho[ii] +
                                                                       2.0 * st_2.Rrho[ii] + st_3.Rrho[ii]))
                );
                S.rho[ii] = rho;
                S.p[ii] = get_pressure(P, rho);
            }
            else if (S.b[ii] == ORC_OUTLET)
            {
                S.xi[ii] = part_n.xi[ii] + dt * S.v[ii];
            }
        }
        RK_inlet_buffers(o, B, part_n, S, dt, true, &st_1, &st_2, &st_3);
    }
    return S;
}

/* ------------------------------------------------------------------ Integration.cpp:370-443 */
static real find_timestep(Orc& o, State const& S, size_t start, size_t end)
{
    OrcParams const& P = o.P;
    o.maxf = MEPS;
    o.maxAf = MEPS;
    o.maxRhoi = MEPS;
    o.maxdrho = MEPS;
    o.maxU = MEPS;
    o.minST = 9999999.0;
    for (size_t ii = start; ii < end; ++ii)
    {
        o.maxf = std::max(o.maxf, norm(S.acc[ii]));
        o.maxAf = std::max(o.maxAf, norm(S.Af[ii]));
        o.maxdrho = std::max(o.maxdrho, std::fabs(S.Rrho[ii]));
        o.maxRhoi = std::max(o.maxRhoi, std::fabs(S.rho[ii] - P.rho_rest));
        /* Q5: IEEE semantics, division by sigma*|curve| = 0 gives +inf which min() ignores */
        o.minST =
            std::min(o.minST, std::sqrt(S.rho[ii] * P.dx * P.dx / (2.0 * M_PI * P.sig * std::fabs(S.curve[ii]))));
        o.maxU = std::max(o.maxU, norm(S.v[ii]));
        if (P.ale)
            o.maxShift = std::max(o.maxShift, norm(S.vPert[ii]));
    }
    o.maxRho_pc = 100 * o.maxRhoi / P.rho_rest;
    real f[6];
    f[0] = 0.25 * std::sqrt(P.H / o.maxf);
    f[1] = 2 * P.H / (o.maxU);
    f[2] = 0.125 * P.H_sq * P.rho_rest / P.mu;
    f[3] = 0.067 * o.minST;
    f[4] = 0.5 * std::sqrt(P.H / o.maxdrho);
    f[5] = 1.5 * P.H / P.speed_sound;
    o.safe_dt = 0.75 * *std::min_element(f, f + 6);
    real dt = P.cfl * o.safe_dt;
    if (dt < P.delta_t_min)
        dt = P.delta_t_min;
    else if (dt > P.delta_t_max)
        dt = P.delta_t_max;
    if (dt > P.last_frame_time + P.frame_time_interval - P.current_time)
        dt = P.last_frame_time + P.frame_time_interval - P.current_time + P.delta_t_min;
    return dt;
}

/* ------------------------------------------------------------------ Integration.cpp:27-107 */
static void frozen_terms(Orc& o, State& S, real& npd)
{
    dSPH_PreStep(o, o.total_points, S, npd);
    if (o.P.asource == meshInfl)
        get_aero_velocity_mesh(o, o.pn, S, npd); /* S is pnp1 on this path (Integration.cpp:47,82) */
    get_aero_velocity(o, o.start_index, o.end_index, S);
    Detect_Surface(o, o.start_index, o.end_index, S);
    dissipation_terms(o, o.start_index, o.end_index, S);
    if (o.P.ale)
        particle_shift(o, o.start_index, o.end_index, S);
    Check_Pipe_Outlet(o, S);
}

static real integrate_no_update(Orc& o, OrcStepStats* stats)
{
    o.start_index = o.bound_points;
    o.end_index = o.total_points;
    o.iteration = 0;
    real rms_error = 0.0, logbase = 0.0, npd = 1.0;
    State& pn = o.pn;
    State& pnp1 = o.pnp1;

    o.P.delta_t = find_timestep(o, pnp1, o.start_index, o.end_index);
    update_neighbours(o, pnp1);
    frozen_terms(o, pnp1, npd);

    std::vector<Vec> xih(o.end_index - o.start_index);
    for (size_t ii = o.start_index; ii < o.end_index; ++ii) xih[ii - o.start_index] = pnp1.xi[ii];

    /* solve_prestep, Integration.cpp:306-336 */
    if (o.P.solver_type == 1)
    {
        /* Get_First_RK, Runge_Kutta.cpp:462-476: st_1 (= pnp1) built from part_n (= pn) */
        pnp1 = RK_intermediate(o, npd, pn, pn, 0.5 * o.P.delta_t);
        logbase = Check_RK_Error(o, 0.0, pn, pnp1);
    }
    else
    {
        Do_NB_Iter(o, npd, pn, pnp1);
        (void)NB_Check_Error(o, rms_error, logbase, xih, pn, pnp1, o.iteration);
        o.iteration++;
    }

    update_neighbours(o, pnp1);
    frozen_terms(o, pnp1, npd);

    /* solve_step, Integration.cpp:339-368 */
    if (o.P.solver_type == 1)
    {
        State st_1 = pnp1;
        real const dt_inter = 0.5 * o.P.delta_t, dt_final = o.P.delta_t;
        State st_2 = RK_intermediate(o, npd, pn, st_1, dt_inter);
        State st_3 = RK_intermediate(o, npd, pn, st_2, dt_final);
        pnp1 = RK_final(o, npd, dt_final, pn, st_1, st_2, st_3);
        rms_error = Check_RK_Error(o, logbase, st_3, pnp1);
    }
    else
    {
        rms_error = NB_Solve(o, npd, logbase, o.iteration, xih, pn, pnp1);
    }
    if (stats)
    {
        stats->npd = npd;
        stats->logbase = logbase;
    }
    return rms_error;
}

/* shapes/inlet.cpp:578-640 */
static unsigned update_buffer_region(Orc& o, State& pnp1, size_t& end)
{
    unsigned nAdd = 0;
    for (size_t block_id = o.n_bound_blocks; block_id < o.limits.size(); ++block_id)
    {
        Block& B = o.limits[block_id];
        if (B.block_type != inletZone)
            continue;
        for (size_t ii = 0; ii < B.back.size(); ++ii)
        {
            long const pID = B.back[ii];
            if (dot(pnp1.xi[pID], B.insert_norm) > B.insconst)
            {
                pnp1.b[pID] = ORC_PIPE;
                pnp1.b[B.buffer[ii][0]] = ORC_BACK;
                B.back[ii] = B.buffer[ii][0];
                for (size_t jj = 0; jj + 1 < B.buffer[0].size(); ++jj) B.buffer[ii][jj] = B.buffer[ii][jj + 1];
                if (o.total_points < o.max_points)
                {
                    long const src = B.buffer[ii].back();
                    Vec const xi = pnp1.xi[src] - o.P.dx * B.insert_norm;
                    pnp1.insert_from(size_t(B.second), xi, size_t(src), ORC_BUFFER, o.next_part_id);
                    B.buffer[ii].back() = B.second;
                    B.second++;
                    for (size_t jj = block_id + 1; jj < o.limits.size(); ++jj)
                    {
                        o.limits[jj].first++;
                        o.limits[jj].second++;
                    }
                    o.fluid_points++;
                    o.total_points++;
                    end++;
                    o.next_part_id++;
                    nAdd++;
                }
            }
        }
    }
    return nAdd;
}

/* Integration.cpp:109-226 (IPT hand-off is out of scope) */
static size_t update_data(Orc& o)
{
    State& pn = o.pn;
    State& pnp1 = o.pnp1;
    unsigned const nAdd = update_buffer_region(o, pnp1, o.end_index);
    for (size_t ii = o.start_index; ii < o.end_index && ii < pnp1.n; ++ii)
        if (pnp1.rho[ii] < 0.0001)
            pnp1.rho[ii] = o.P.rho_rest;

    std::set<size_t> to_del;
    std::vector<size_t> n_del_per_block(o.limits.size(), 0);
    for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
    {
        Block const& B = o.limits[block];
        if (B.delconst == default_val)
            continue;
        for (long ii = B.first; ii < B.second; ++ii)
            if (dot(pnp1.xi[ii], B.delete_norm) > B.delconst)
            {
                to_del.insert(size_t(ii));
                n_del_per_block[block]++;
            }
    }
    unsigned nDel = 0;
    if (!to_del.empty())
    {
        for (auto itr = to_del.rbegin(); itr != to_del.rend(); ++itr)
        {
            pnp1.erase(*itr);
            o.total_points--;
            o.fluid_points--;
            nDel++;
            o.delete_count++;
        }
        size_t delshift = 0;
        for (size_t block = o.n_bound_blocks; block < o.limits.size(); ++block)
        {
            Block& B = o.limits[block];
            B.first -= long(delshift);
            delshift += n_del_per_block[block];
            B.second -= long(delshift);
            for (long& back : B.back) back -= long(delshift);
            for (auto& buffer : B.buffer)
                for (long& part : buffer) part -= long(delshift);
        }
    }
    if (nAdd != 0 || nDel != 0)
        update_neighbours(o, pnp1);
    pn = pnp1;
    o.last_nadd = int(nAdd);
    o.last_ndel = int(nDel);
    return pnp1.n;
}

/* Integration.cpp:233-303 */
static real integrate(Orc& o, OrcStepStats* stats)
{
    OrcParams& P = o.P;
    real const step_error = integrate_no_update(o, stats);
    size_t const npts = update_data(o);
    if (stats)
    {
        stats->dt = P.delta_t;
        stats->cfl_ratio = P.delta_t / o.safe_dt;
        stats->rms_error = step_error;
        stats->maxRho_pc = o.maxRho_pc;
        stats->maxf = o.maxf;
        stats->maxAf = o.maxAf;
        stats->maxShift = o.maxShift;
        stats->safe_dt = o.safe_dt;
        stats->iterations = int(o.iteration);
        stats->n_add = o.last_nadd;
        stats->n_del = o.last_ndel;
        stats->total_points = int(npts);
    }
    if (npts == 0)
        return 0;
    P.current_time += P.delta_t;
    if (step_error > P.min_residual || o.maxRho_pc > P.rho_max_iter)
    {
        if (step_error > 0.6 * P.min_residual)
        {
            P.cfl = std::max(P.cfl_min, P.cfl - P.cfl_step);
            P.n_unstable = 0;
        }
        else if (P.n_unstable > P.n_unstable_limit)
        {
            P.cfl = std::max(P.cfl_min, P.cfl - P.cfl_step);
            P.n_unstable = 0;
        }
        else
            P.n_unstable++;
    }
    else
        P.n_unstable = 0;

    if (o.iteration < P.subits_factor * P.max_subits && P.n_unstable == 0)
    {
        if (P.n_stable > P.n_stable_limit)
        {
            P.cfl = std::min(P.cfl_max, P.cfl + P.cfl_step);
            P.n_stable = 0;
        }
        else
            P.n_stable++;
    }
    else
        P.n_stable = 0;
    return step_error;
}

/* ================================================================== C ABI */
extern "C" {

int orc_compiled_dim(void) { return DIM; }

Orc* orc_create(const OrcParams* p)
{
    if (p->dim != DIM)
    {
        std::fprintf(stderr, "orc_create: params.dim=%d but library compiled with DIM=%d\n", p->dim, DIM);
        return nullptr;
    }
    Orc* o = new Orc();
    o->P = *p;
    return o;
}
void orc_destroy(Orc* o) { delete o; }
void orc_get_params(Orc* o, OrcParams* out) { *out = o->P; }
void orc_set_params(Orc* o, const OrcParams* in) { o->P = *in; }

static Vec vec_from(const double* p)
{
    Vec v = vzero();
    if (p)
        for (int d = 0; d < DIM; ++d) v[d] = p[d];
    return v;
}

int orc_add_block(Orc* o, int is_fluid, int64_t first, int64_t second, int bound_solver, int no_slip,
                  int block_type, int fixed_vel_or_dynamic, int ntimes, const double* times, const double* vels,
                  const double* insert_norm, double insconst, const double* delete_norm, double delconst,
                  const double* aero_norm, double aeroconst, int nback, const int64_t* back, int nbuf,
                  const int64_t* buffer)
{
    Block B;
    B.first = long(first);
    B.second = long(second);
    B.is_fluid = is_fluid;
    B.bound_solver = bound_solver;
    B.no_slip = no_slip;
    B.block_type = block_type;
    B.fixed_vel_or_dynamic = fixed_vel_or_dynamic;
    B.nTimes = size_t(ntimes);
    for (int t = 0; t < ntimes; ++t) B.times.push_back(times[t]);
    int const nv = std::max(1, ntimes);
    for (int t = 0; t < nv; ++t) B.vels.push_back(vels ? vec_from(vels + 3 * t) : vzero());
    Vec dv;
    for (int d = 0; d < DIM; ++d) dv[d] = default_val;
    B.insert_norm = insert_norm ? vec_from(insert_norm) : dv;
    B.delete_norm = delete_norm ? vec_from(delete_norm) : dv;
    B.aero_norm = aero_norm ? vec_from(aero_norm) : dv;
    B.insconst = insconst;
    B.delconst = delconst;
    B.aeroconst = aeroconst;
    for (int i = 0; i < nback; ++i)
    {
        B.back.push_back(long(back[i]));
        std::vector<long> buf;
        for (int j = 0; j < nbuf; ++j) buf.push_back(long(buffer[size_t(i) * nbuf + j]));
        B.buffer.push_back(buf);
    }
    if (is_fluid)
        o->n_fluid_blocks++;
    else
    {
        if (o->n_fluid_blocks != 0)
            return -1;
        o->n_bound_blocks++;
    }
    o->limits.push_back(B);
    return int(o->limits.size()) - 1;
}
void orc_clear_blocks(Orc* o)
{
    o->limits.clear();
    o->n_bound_blocks = o->n_fluid_blocks = 0;
}
int orc_get_block_range(Orc* o, int block, int64_t* first, int64_t* second)
{
    if (block < 0 || size_t(block) >= o->limits.size())
        return -1;
    *first = o->limits[block].first;
    *second = o->limits[block].second;
    return 0;
}

int orc_set_particles(Orc* o, int64_t n, int64_t bound_points, const double* xi, const double* v, const double* rho,
                      const double* p, const double* m, const int32_t* b, const int64_t* part_id)
{
    State S;
    S.resize(size_t(n));
    for (size_t i = 0; i < size_t(n); ++i)
    {
        for (int d = 0; d < DIM; ++d)
        {
            S.xi[i][d] = xi[i * DIM + d];
            S.v[i][d] = v ? v[i * DIM + d] : 0.0;
        }
        S.rho[i] = rho[i];
        S.p[i] = p[i];
        S.m[i] = m[i];
        S.b[i] = b[i];
        S.part_id[i] = part_id ? long(part_id[i]) : long(i);
        /* FJSPH.cpp:115-126 (Asource != meshInfl) */
        S.cellRho[i] = o->P.rho_g;
        S.cellP[i] = o->P.p_ref;
        for (int d = 0; d < DIM; ++d) S.cellV[i][d] = o->P.v_inf[d];
    }
    o->pn = S;
    o->pnp1 = S;
    o->bound_points = size_t(bound_points);
    o->total_points = size_t(n);
    o->fluid_points = size_t(n - bound_points);
    long maxid = -1;
    for (size_t i = 0; i < size_t(n); ++i) maxid = std::max(maxid, S.part_id[i]);
    o->next_part_id = maxid + 1;
    o->nb_off.assign(size_t(n) + 1, 0);
    o->nb_idx.clear();
    o->nb_d2.clear();
    if (o->limits.empty())
    {
        /* default: one wall block (if any walls) + one fluid block */
        double z[3] = {0, 0, 0};
        if (bound_points > 0)
            orc_add_block(o, 0, 0, bound_points, pressure_G, 0, 0, 0, 0, nullptr, z, nullptr, default_val, nullptr,
                          default_val, nullptr, default_val, 0, nullptr, 0, nullptr);
        orc_add_block(o, 1, bound_points, n, 0, 0, 0, 0, 0, nullptr, z, nullptr, default_val, nullptr, default_val,
                      nullptr, default_val, 0, nullptr, 0, nullptr);
    }
    return 0;
}
int64_t orc_count(Orc* o) { return int64_t(o->pnp1.n); }
int64_t orc_bound_points(Orc* o) { return int64_t(o->bound_points); }

static std::vector<Vec>* find_vec(State& S, std::string const& n)
{
    if (n == "xi") return &S.xi;
    if (n == "v") return &S.v;
    if (n == "acc") return &S.acc;
    if (n == "Af") return &S.Af;
    if (n == "aVisc") return &S.aVisc;
    if (n == "cellV") return &S.cellV;
    if (n == "gradRho") return &S.gradRho;
    if (n == "norm") return &S.norm;
    if (n == "bNorm") return &S.bNorm;
    if (n == "vPert") return &S.vPert;
    return nullptr;
}
static std::vector<real>* find_scalar(State& S, std::string const& n)
{
    if (n == "Rrho") return &S.Rrho;
    if (n == "rho") return &S.rho;
    if (n == "p") return &S.p;
    if (n == "m") return &S.m;
    if (n == "curve") return &S.curve;
    if (n == "norm_curve") return &S.norm_curve;
    if (n == "woccl") return &S.woccl;
    if (n == "pDist") return &S.pDist;
    if (n == "deltaD") return &S.deltaD;
    if (n == "cellP") return &S.cellP;
    if (n == "cellRho") return &S.cellRho;
    if (n == "colourG") return &S.colourG;
    if (n == "colour") return &S.colour;
    if (n == "lam") return &S.lam;
    if (n == "lam_nb") return &S.lam_nb;
    if (n == "kernsum") return &S.kernsum;
    if (n == "y") return &S.y;
    return nullptr;
}

int orc_get_f64(Orc* o, int level, const char* name, double* out)
{
    State& S = level ? o->pnp1 : o->pn;
    std::string const n(name);
    if (auto* f = find_vec(S, n))
    {
        for (size_t i = 0; i < S.n; ++i)
            for (int d = 0; d < DIM; ++d) out[i * DIM + d] = (*f)[i][d];
        return DIM;
    }
    if (n == "L")
    {
        for (size_t i = 0; i < S.n; ++i)
            for (int a = 0; a < DIM; ++a)
                for (int c = 0; c < DIM; ++c) out[(i * DIM + a) * DIM + c] = S.L[i].a[a][c];
        return DIM * DIM;
    }
    if (auto* f = find_scalar(S, n))
    {
        for (size_t i = 0; i < S.n; ++i) out[i] = (*f)[i];
        return 1;
    }
    return -1;
}
int orc_set_f64(Orc* o, int level, const char* name, const double* in)
{
    State& S = level ? o->pnp1 : o->pn;
    std::string const n(name);
    if (auto* f = find_vec(S, n))
    {
        for (size_t i = 0; i < S.n; ++i)
            for (int d = 0; d < DIM; ++d) (*f)[i][d] = in[i * DIM + d];
        return DIM;
    }
    if (n == "L")
    {
        for (size_t i = 0; i < S.n; ++i)
            for (int a = 0; a < DIM; ++a)
                for (int c = 0; c < DIM; ++c) S.L[i].a[a][c] = in[(i * DIM + a) * DIM + c];
        return DIM * DIM;
    }
    if (auto* f = find_scalar(S, n))
    {
        for (size_t i = 0; i < S.n; ++i) (*f)[i] = in[i];
        return 1;
    }
    return -1;
}
int orc_get_i64(Orc* o, int level, const char* name, int64_t* out)
{
    State& S = level ? o->pnp1 : o->pn;
    std::string const n(name);
    for (size_t i = 0; i < S.n; ++i)
    {
        if (n == "part_id") out[i] = S.part_id[i];
        else if (n == "cellID") out[i] = S.cellID[i];
        else if (n == "b") out[i] = S.b[i];
        else if (n == "surf") out[i] = S.surf[i];
        else if (n == "surfzone") out[i] = S.surfzone[i];
        else if (n == "internal") out[i] = S.internal[i];
        else if (n == "ipt_n_failed") out[i] = S.ipt_n_failed[i];
        else return -1;
    }
    return 1;
}
int orc_set_i64(Orc* o, int level, const char* name, const int64_t* in)
{
    State& S = level ? o->pnp1 : o->pn;
    std::string const n(name);
    for (size_t i = 0; i < S.n; ++i)
    {
        if (n == "part_id") S.part_id[i] = long(in[i]);
        else if (n == "cellID") S.cellID[i] = long(in[i]);
        else if (n == "b") S.b[i] = int(in[i]);
        else if (n == "surf") S.surf[i] = int(in[i]);
        else if (n == "surfzone") S.surfzone[i] = int(in[i]);
        else if (n == "internal") S.internal[i] = int(in[i]);
        else if (n == "ipt_n_failed") S.ipt_n_failed[i] = int(in[i]);
        else return -1;
    }
    return 1;
}

void orc_update_neighbours(Orc* o) { update_neighbours(*o, o->pnp1); }
int64_t orc_neighbour_total(Orc* o) { return o->nb_off.empty() ? 0 : int64_t(o->nb_off.back()); }
void orc_get_neighbours(Orc* o, int64_t* offsets, int64_t* idx, double* d2)
{
    for (size_t i = 0; i < o->nb_off.size(); ++i) offsets[i] = o->nb_off[i];
    for (size_t k = 0; k < o->nb_idx.size(); ++k)
    {
        idx[k] = o->nb_idx[k];
        d2[k] = o->nb_d2[k];
    }
}
/* list lengths (self included, as outlist[i].size()): the full-size parity tests compare these without copying the lists */
void orc_neighbour_counts(Orc* o, int64_t* counts)
{
    for (size_t i = 0; i + 1 < o->nb_off.size(); ++i) counts[i] = int64_t(o->nb_off[i + 1] - o->nb_off[i]);
}
static void set_range(Orc* o)
{
    o->start_index = o->bound_points;
    o->end_index = o->total_points;
}
double orc_prestep(Orc* o)
{
    set_range(o);
    real npd = 1.0;
    dSPH_PreStep(*o, o->total_points, o->pnp1, npd);
    return npd;
}
void orc_aero_velocity(Orc* o)
{
    set_range(o);
    if (o->P.asource == meshInfl)
    {
        real npd = 1.0;
        get_aero_velocity_mesh(*o, o->pn, o->pnp1, npd);
    }
    get_aero_velocity(*o, o->start_index, o->end_index, o->pnp1);
}
/* MESH upload: faces as CSR vertex lists, leftright pairs, cell -> faces CSR, centres and the cell solution */
void orc_set_mesh(Orc* o, int64_t n_verts, const double* verts, int64_t n_faces, const int64_t* face_ptr,
                  const int64_t* face_vtx, const int32_t* leftright, int64_t n_cells, const int64_t* cell_ptr,
                  const int64_t* cell_faces, const double* cCentre, const double* cVel, const double* cP,
                  const double* cRho)
{
    Mesh& M = o->cells;
    M = Mesh();
    M.verts.resize(size_t(n_verts));
    for (int64_t i = 0; i < n_verts; ++i)
        for (int d = 0; d < DIM; ++d) M.verts[size_t(i)][d] = verts[i * DIM + d];
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
        for (int d = 0; d < DIM; ++d)
        {
            M.cCentre[size_t(c)][d] = cCentre[c * DIM + d];
            M.cVel[size_t(c)][d] = cVel[c * DIM + d];
        }
    }
}
int orc_first_cell_errors(Orc* o) { return o->first_cell_errors; }
void orc_detect_surface(Orc* o)
{
    set_range(o);
    Detect_Surface(*o, o->start_index, o->end_index, o->pnp1);
}
void orc_dissipation(Orc* o)
{
    set_range(o);
    dissipation_terms(*o, o->start_index, o->end_index, o->pnp1);
}
void orc_particle_shift(Orc* o)
{
    set_range(o);
    particle_shift(*o, o->start_index, o->end_index, o->pnp1);
}
void orc_forces(Orc* o, double npd)
{
    set_range(o);
    get_acc_and_Rrho(*o, npd, o->pnp1);
}
void orc_nb_iter(Orc* o, double npd)
{
    set_range(o);
    Do_NB_Iter(*o, npd, o->pn, o->pnp1);
}
double orc_find_timestep(Orc* o)
{
    set_range(o);
    return find_timestep(*o, o->pnp1, o->start_index, o->end_index);
}
double orc_integrate_no_update(Orc* o, OrcStepStats* s)
{
    if (s)
        std::memset(s, 0, sizeof(*s));
    real const e = integrate_no_update(*o, s);
    if (s)
    {
        s->dt = o->P.delta_t;
        s->rms_error = e;
        s->safe_dt = o->safe_dt;
        s->cfl_ratio = o->P.delta_t / o->safe_dt;
        s->maxRho_pc = o->maxRho_pc;
        s->maxf = o->maxf;
        s->maxAf = o->maxAf;
        s->maxShift = o->maxShift;
        s->iterations = int(o->iteration);
        s->total_points = int(o->pnp1.n);
    }
    return e;
}
double orc_integrate(Orc* o, OrcStepStats* s)
{
    if (s)
        std::memset(s, 0, sizeof(*s));
    return integrate(*o, s);
}

int orc_qr_inverse(const double* a, double* inv)
{
    Mat A, I;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) A.a[i][j] = a[i * DIM + j];
    int const ok = qr_inverse(A, I);
    if (ok)
        for (int i = 0; i < DIM; ++i)
            for (int j = 0; j < DIM; ++j) inv[i * DIM + j] = I.a[i][j];
    return ok;
}
double orc_min_eigenvalue(const double* a)
{
    Mat A;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) A.a[i][j] = a[i * DIM + j];
    return min_eigenvalue(A);
}
double orc_kernel(double r, double H, double Wc) { return Kernel(r, H, Wc); }
double orc_get_n_full(double dx, double H) { return get_n_full(dx, H); }

} /* extern "C" */

/* the implicit particle tracker downstream of the delete planes (IPT.cpp, Containment.cpp:896-1079) */
#include "ipt_oracle.inc"
