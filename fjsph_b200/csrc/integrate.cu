// integrate.cu — time-step orchestration and the fused streaming kernels around the pair sweeps.
//
// Replaces:
//   Integrator::integrate / integrate_no_update / find_timestep / update_data  reference src/Integration.cpp:27-443
//   Newmark_Beta::Do_NB_Iter (update part) / Check_Error / Newmark_Beta        reference src/Newmark_Beta.cpp:10-331
//   Get_First_RK / do_runge_kutta_{intermediate,final}_step / Runge_Kutta4      reference src/Runge_Kutta.cpp:10-515
// The update kernels fuse: Newmark/RK state update + density clamp + EOS + refresh of the derived gather
// quantities (V, p/rho^2) + the convergence residual sum |x - x_prev|^2 (block partials, reduced in a
// fixed order so the residual is reproducible); one 8-byte readback per sub-iteration.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.cuh"

namespace
{
constexpr int TPB = 256;
#define FJ_PI 3.14159265358979323846
// what a Newmark-Beta / RK update (or a wall treatment) changes on a particle and a neighbour reads
constexpr unsigned FJ_HX_STATE = FJ_HX_P0 | FJ_HX_P1 | FJ_HX_P2 | FJ_HX_TH;

struct BlockTable
{
    int n_bound_blocks;
    unsigned char solver[64];
};

__device__ __forceinline__ double eos_pressure(const DevConst& C, double rho)
{
    if (C.pressure_rel == 0)
        return C.B * (pow(rho / C.rho_rest, C.gam) - 1.0) + C.press_back;
    return C.c2 * (rho - C.rho_rest) + C.press_back;
}

__device__ __forceinline__ double block_sum(double v, double* sm)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int(blockDim.x) >> 5); ++k) r += sm[k];
    __syncthreads();
    return r;
}

// deterministic final reduction of block partials: red[b*ncomp + c] -> out[c]
__global__ void k_reduce_sum(const double* __restrict__ red, int nblocks, int ncomp, double* __restrict__ out)
{
    __shared__ double sm[TPB];
    for (int c = 0; c < ncomp; ++c)
    {
        double s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += red[size_t(b) * ncomp + c];
        sm[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1)
        {
            if (int(threadIdx.x) < o)
                sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0)
            out[c] = sm[0];
        __syncthreads();
    }
}

// ------------------------------------------------------------------ find_timestep reductions
// partial[b*7 + {maxf2, maxAf2, maxdrho, maxRhoi, minST, maxU2, maxShift2}]
__global__ void k_timestep_partials(Level S, const int* __restrict__ blk, int n_bound_blocks, DevConst C, int n,
                                    double* __restrict__ partial)
{
    double v[7] = {0, 0, 0, 0, 1e300, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        if (blk[i] < n_bound_blocks)
            continue;
        const double4 a = S.ACC[i];
        const double4 af = S.AF[i];
        const double4 vv = S.P1[i];
        const double4 q = S.P2[i];
        const double curve = S.AV[i].w;
        v[0] = fmax(v[0], a.x * a.x + a.y * a.y + a.z * a.z);
        v[1] = fmax(v[1], af.x * af.x + af.y * af.y + af.z * af.z);
        v[2] = fmax(v[2], fabs(a.w));
        v[3] = fmax(v[3], fabs(vv.w - C.rho_rest));
        /* Q5: IEEE division, sigma*|curve| == 0 gives +inf which min() ignores */
        v[4] = fmin(v[4], sqrt(vv.w * C.dx * C.dx / (2.0 * FJ_PI * C.sig * fabs(curve))));
        v[5] = fmax(v[5], vv.x * vv.x + vv.y * vv.y + vv.z * vv.z);
        v[6] = fmax(v[6], q.x * q.x + q.y * q.y + q.z * q.z);
    }
    __shared__ double sm[7][TPB / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 7; ++c)
    {
        double x = v[c];
        for (int o = 16; o > 0; o >>= 1)
        {
            const double y = __shfl_xor_sync(0xffffffffu, x, o);
            x = (c == 4) ? fmin(x, y) : fmax(x, y);
        }
        if (lane == 0)
            sm[c][w] = x;
    }
    __syncthreads();
    if (threadIdx.x < 7)
    {
        const int c = threadIdx.x;
        double x = sm[c][0];
        for (int k = 1; k < TPB / 32; ++k) x = (c == 4) ? fmin(x, sm[c][k]) : fmax(x, sm[c][k]);
        partial[blockIdx.x * 7 + c] = x;
    }
}
__global__ void k_timestep_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out,
                                 bool negate_min = false)
{
    // one warp per component, lanes stride the block partials
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= 7)
        return;
    double x = (c == 4) ? 1e300 : 0.0;
    for (int b = lane; b < nblocks; b += 32)
    {
        const double y = partial[b * 7 + c];
        x = (c == 4) ? fmin(x, y) : fmax(x, y);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        const double y = __shfl_xor_sync(0xffffffffu, x, o);
        x = (c == 4) ? fmin(x, y) : fmax(x, y);
    }
    if (lane == 0)
        out[c] = (negate_min && c == 4) ? -x : x; /* min as -max(-x): one MAX all-reduce serves all seven */
}

// keeps V = m/rho and p/rho^2 consistent with (rho, p)
__device__ __forceinline__ void write_state(Level& S, int i, double x, double y, double z, double vx, double vy,
                                            double vz, double rho, const DevConst& C)
{
    const double p = eos_pressure(C, rho);
    double4 th = S.TH[i];
    th.x = p;
    S.TH[i] = th;
    S.P0[i] = make_double4(x, y, z, th.y / rho);
    S.P1[i] = make_double4(vx, vy, vz, rho);
    double4 q = S.P2[i];
    q.w = p / (rho * rho);
    S.P2[i] = q;
}
__device__ __forceinline__ void write_rho(Level& S, int i, double rho, const DevConst& C)
{
    const double p = eos_pressure(C, rho);
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

// ------------------------------------------------------------------ Newmark-Beta update (Newmark_Beta.cpp:137-241)
template <bool ALE>
__global__ void __launch_bounds__(TPB)
    k_nb_update(Level Sn, Level S, const int* __restrict__ blk, BlockTable bt, const int* __restrict__ near_inlet,
                DevConst C, double dt, double beta_t1, double gamma_t1, int n, double* __restrict__ err_partial)
{
    __shared__ double sm[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (i < n)
    {
        const int bl = blk[i];
        const double gamma_t2 = 1 - gamma_t1;
        const double beta_t2 = 0.5 * (1 - 2 * beta_t1);
        if (bl < bt.n_bound_blocks)
        {
            const int solver = bt.solver[bl];
            if (solver == FJSPH_DBC || solver == FJSPH_GHOST)
            {
                const double rho_n = Sn.P1[i].w, Rn = Sn.ACC[i].w;
                double4 acc = S.ACC[i];
                const bool ni = (solver == FJSPH_GHOST) && near_inlet[i];
                const double lo = ni ? C.rho_rest : C.rho_min;
                const double rho = fmax(lo, fmin(C.rho_max, rho_n + dt * (gamma_t1 * acc.w + gamma_t2 * Rn)));
                write_rho(S, i, rho, C);
                if (ni)
                {
                    acc.w = fmax(0.0, acc.w);
                    S.ACC[i] = acc;
                }
            }
        }
        else
        {
            const int b = S.b[i];
            if (b > FJSPH_BUFFER && b != FJSPH_OUTLET)
            {
                const double4 xn = Sn.P0[i], vn = Sn.P1[i], an = Sn.ACC[i];
                const double4 a = S.ACC[i];
                const double4 xo = S.P0[i];
                const double dt2 = dt * dt;
                double x, y, z;
                if (ALE)
                {
                    const double4 q = S.P2[i];
                    x = xn.x + dt * (vn.x + q.x) + dt2 * (beta_t1 * a.x + beta_t2 * an.x);
                    y = xn.y + dt * (vn.y + q.y) + dt2 * (beta_t1 * a.y + beta_t2 * an.y);
                    z = xn.z + dt * (vn.z + q.z) + dt2 * (beta_t1 * a.z + beta_t2 * an.z);
                }
                else
                {
                    x = xn.x + dt * vn.x + dt2 * (beta_t2 * an.x + beta_t1 * a.x);
                    y = xn.y + dt * vn.y + dt2 * (beta_t2 * an.y + beta_t1 * a.y);
                    z = xn.z + dt * vn.z + dt2 * (beta_t2 * an.z + beta_t1 * a.z);
                }
                const double vx = vn.x + dt * (gamma_t1 * a.x + gamma_t2 * an.x);
                const double vy = vn.y + dt * (gamma_t1 * a.y + gamma_t2 * an.y);
                const double vz = vn.z + dt * (gamma_t1 * a.z + gamma_t2 * an.z);
                const double rho = fmax(C.rho_min, fmin(C.rho_max, vn.w + dt * (gamma_t1 * a.w + gamma_t2 * an.w)));
                write_state(S, i, x, y, z, vx, vy, vz, rho, C);
                const double ex = x - xo.x, ey = y - xo.y, ez = z - xo.z;
                err = ex * ex + ey * ey + ez * ez;
            }
            else if (b == FJSPH_OUTLET)
            {
                const double4 xn = Sn.P0[i];
                const double4 v = S.P1[i];
                double4 xo = S.P0[i];
                const double x = xn.x + dt * v.x, y = xn.y + dt * v.y, z = xn.z + dt * v.z;
                const double ex = x - xo.x, ey = y - xo.y, ez = z - xo.z;
                err = ex * ex + ey * ey + ez * ez;
                xo.x = x;
                xo.y = y;
                xo.z = z;
                S.P0[i] = xo;
            }
        }
    }
    const double tot = block_sum(err, sm);
    if (threadIdx.x == 0)
        err_partial[blockIdx.x] = tot;
}

// ------------------------------------------------------------------ Runge-Kutta stage updates
// Intermediate stage (Runge_Kutta.cpp:137-171): S holds the previous stage (forces just evaluated on it).
//   x = x_n + dt_s (v_prev + vPert_prev) ; v = v_n + dt_s acc ; rho = clamp(rho_n + dt_s Rrho)
// err = |x_new - x_ref|^2 with x_ref = Sn.x (Get_First_RK's Check_RK_Error) when err_vs_n, else unused.
template <bool ALE>
__global__ void __launch_bounds__(TPB)
    k_rk_stage(Level Sn, Level S, const int* __restrict__ blk, BlockTable bt, const int* __restrict__ near_inlet,
               DevConst C, double dt_s, int n, double* __restrict__ err_partial, int part)
{
    /* part 0: the DBC / Ghost wall densities, BEFORE the stage's force evaluation (Runge_Kutta.cpp:76-131: the walls are
       advanced with the Rrho the boundary treatment has just given them, and get_acc_and_Rrho sees the result);
       part 1: the fluid, after it (Runge_Kutta.cpp:137-171) */
    __shared__ double sm[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (i < n)
    {
        const int bl = blk[i];
        if (bl < bt.n_bound_blocks)
        {
            const int solver = bt.solver[bl];
            if (part == 0 && (solver == FJSPH_DBC || solver == FJSPH_GHOST))
            {
                double4 acc = S.ACC[i];
                const bool ni = (solver == FJSPH_GHOST) && near_inlet[i];
                const double lo = ni ? C.rho_rest : C.rho_min;
                const double rho = fmax(lo, fmin(C.rho_max, Sn.P1[i].w + dt_s * acc.w));
                write_rho(S, i, rho, C);
                if (ni)
                {
                    acc.w = fmax(0.0, acc.w);
                    S.ACC[i] = acc;
                }
            }
        }
        else if (part == 1 && S.b[i] > FJSPH_BUFFER)
        {
            const double4 xn = Sn.P0[i], vn = Sn.P1[i];
            const double4 v = S.P1[i], a = S.ACC[i];
            double ux = v.x, uy = v.y, uz = v.z;
            if (ALE)
            {
                const double4 q = S.P2[i];
                ux += q.x;
                uy += q.y;
                uz += q.z;
            }
            const double x = xn.x + dt_s * ux, y = xn.y + dt_s * uy, z = xn.z + dt_s * uz;
            const double rho = fmax(C.rho_min, fmin(C.rho_max, vn.w + dt_s * a.w));
            write_state(S, i, x, y, z, vn.x + dt_s * a.x, vn.y + dt_s * a.y, vn.z + dt_s * a.z, rho, C);
            const double ex = x - xn.x, ey = y - xn.y, ez = z - xn.z;
            err = ex * ex + ey * ey + ez * ez;
        }
    }
    if (part == 0)
        return;
    const double tot = block_sum(err, sm);
    if (threadIdx.x == 0)
        err_partial[blockIdx.x] = tot;
}

// accumulate the RK4 weighted sums in the reference's left-to-right order (Runge_Kutta.cpp:363-389):
//   sum = (n) + 2*(st_1) + 2*(st_2) + (st_3).  first = 1 initialises with (pn) + 2*(st_1).
template <bool ALE>
__global__ void k_rk_accumulate(Level Sn, Level S, double4* __restrict__ sum_v, double4* __restrict__ sum_a,
                                double weight, int first, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double4 sv, sa;
    if (first)
    {
        const double4 vn = Sn.P1[i], an = Sn.ACC[i];
        sv = make_double4(vn.x, vn.y, vn.z, an.w);
        if (ALE)
        {
            const double4 qn = Sn.P2[i];
            sv.x += qn.x;
            sv.y += qn.y;
            sv.z += qn.z;
        }
        sa = make_double4(an.x, an.y, an.z, 0.0);
    }
    else
    {
        sv = sum_v[i];
        sa = sum_a[i];
    }
    const double4 v = S.P1[i], a = S.ACC[i];
    double ux = v.x, uy = v.y, uz = v.z;
    if (ALE)
    {
        const double4 q = S.P2[i];
        ux += q.x;
        uy += q.y;
        uz += q.z;
    }
    sv.x += weight * ux;
    sv.y += weight * uy;
    sv.z += weight * uz;
    sv.w += weight * a.w;
    sa.x += weight * a.x;
    sa.y += weight * a.y;
    sa.z += weight * a.z;
    sum_v[i] = sv;
    sum_a[i] = sa;
}

// final stage (Runge_Kutta.cpp:355-395); err = |x_new - x_st3|^2 (Check_RK_Error(st_3, part_np1))
__global__ void __launch_bounds__(TPB)
    k_rk_final(Level Sn, Level S, const int* __restrict__ blk, BlockTable bt, const int* __restrict__ near_inlet,
               const double4* __restrict__ sum_v, const double4* __restrict__ sum_a, DevConst C, double dt, int n,
               double* __restrict__ err_partial, int part)
{
    /* part 0: wall densities before the last force evaluation (Runge_Kutta.cpp:276-349), part 1: the fluid after it */
    __shared__ double sm[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (i < n)
    {
        const int bl = blk[i];
        const double dt6 = dt / 6.0;
        if (bl < bt.n_bound_blocks)
        {
            const int solver = bt.solver[bl];
            if (part == 0 && (solver == FJSPH_DBC || solver == FJSPH_GHOST))
            {
                /* sum_v.w holds Rrho_n + 2 Rrho_1 + 2 Rrho_2 + Rrho_3 for walls as well */
                double4 acc = S.ACC[i];
                const bool ni = (solver == FJSPH_GHOST) && near_inlet[i];
                const double lo = ni ? C.rho_rest : C.rho_min;
                const double rho = fmax(lo, fmin(C.rho_max, Sn.P1[i].w + dt6 * sum_v[i].w));
                write_rho(S, i, rho, C);
                if (ni)
                {
                    acc.w = fmax(0.0, acc.w);
                    S.ACC[i] = acc;
                }
            }
        }
        else if (part == 1)
        {
            const int b = S.b[i];
            if (b > FJSPH_BUFFER && b != FJSPH_OUTLET)
            {
                const double4 xn = Sn.P0[i], vn = Sn.P1[i];
                const double4 sv = sum_v[i], sa = sum_a[i];
                const double4 xo = S.P0[i];
                const double x = xn.x + dt6 * sv.x, y = xn.y + dt6 * sv.y, z = xn.z + dt6 * sv.z;
                const double rho = fmax(C.rho_min, fmin(C.rho_max, vn.w + dt6 * sv.w));
                write_state(S, i, x, y, z, vn.x + dt6 * sa.x, vn.y + dt6 * sa.y, vn.z + dt6 * sa.z, rho, C);
                const double ex = x - xo.x, ey = y - xo.y, ez = z - xo.z;
                err = ex * ex + ey * ey + ez * ez;
            }
            else if (b == FJSPH_OUTLET)
            {
                const double4 xn = Sn.P0[i];
                const double4 v = S.P1[i];
                double4 xo = S.P0[i];
                const double x = xn.x + dt * v.x, y = xn.y + dt * v.y, z = xn.z + dt * v.z;
                const double ex = x - xo.x, ey = y - xo.y, ez = z - xo.z;
                err = ex * ex + ey * ey + ez * ez;
                xo.x = x;
                xo.y = y;
                xo.z = z;
                S.P0[i] = xo;
            }
        }
    }
    if (part == 0)
        return;
    const double tot = block_sum(err, sm);
    if (threadIdx.x == 0)
        err_partial[blockIdx.x] = tot;
}

// which: 0 = every field; 1 = only what dSPH_PreStep writes (P3 = gradRho + lam, NP = its normal + lam_nb, SC = colour
// terms, L); 2 = every field but those
__global__ void k_copy_level(Level in, Level out, int n, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (which != 2)
    {
        out.P3[i] = in.P3[i];
        out.NP[i] = in.NP[i];
        out.SC[i] = in.SC[i];
        out.L0[i] = in.L0[i];
        out.L1[i] = in.L1[i];
        out.L2[i] = in.L2[i];
        out.L3[i] = in.L3[i];
        out.L4[i] = in.L4[i];
        out.L5[i] = in.L5[i];
        out.L6[i] = in.L6[i];
        out.L7[i] = in.L7[i];
        out.L8[i] = in.L8[i];
    }
    if (which != 1)
    {
        out.P0[i] = in.P0[i];
        out.P1[i] = in.P1[i];
        out.P2[i] = in.P2[i];
        out.P4[i] = in.P4[i];
        out.ACC[i] = in.ACC[i];
        out.AF[i] = in.AF[i];
        out.AV[i] = in.AV[i];
        out.CV[i] = in.CV[i];
        out.BN[i] = in.BN[i];
        out.TH[i] = in.TH[i];
        out.part_id[i] = in.part_id[i];
        out.cellID[i] = in.cellID[i];
        out.b[i] = in.b[i];
        out.surfzone[i] = in.surfzone[i];
        out.internal[i] = in.internal[i];
        out.surf_i[i] = in.surf_i[i];
    }
}
static_assert(sizeof(Level) == 28 * sizeof(void*), "k_copy_level lists the fields of a Level by name: keep it in step with FJ_LEVEL_FIELDS");

BlockTable make_block_table(FjsphEngine* e)
{
    BlockTable bt;
    std::memset(&bt, 0, sizeof(bt));
    bt.n_bound_blocks = e->n_bound_blocks;
    for (int b = 0; b < e->n_bound_blocks && b < 64; ++b) bt.solver[b] = (unsigned char)e->blocks[b].bound_solver;
    return bt;
}

int check_supported(FjsphEngine* e)
{
    if (e->n_bound_blocks > 64)
    {
        fj_set_error("more than 64 boundary blocks are not supported");
        return FJSPH_ERR_INVALID;
    }
    return FJSPH_OK;
}

} // namespace

// ------------------------------------------------------------------ shared host helpers
int fj_reduce_sum(FjsphEngine* e, int nblocks, int ncomp, double* out_host)
{
    k_reduce_sum<<<1, TPB, 0, e->stream>>>(e->red, nblocks, ncomp, e->red_out);
    e->launches++;
    FJ_CUDA(cudaGetLastError());
    /* slab decomposition: the sum over all ranks -- on the device, ahead of the one readback, when the transport can */
    bool reduced = false;
    int st = fj_allreduce_dev(e, FJSPH_COMM_SUM, e->red_out, ncomp, &reduced);
    if (st)
        return st;
    FJ_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, ncomp * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < ncomp; ++c) out_host[c] = e->h_red[c];
    return reduced ? FJSPH_OK : fj_allreduce(e, FJSPH_COMM_SUM, out_host, ncomp);
}

// The engine's stream waits for the second half of a split upload (upload_state_split, abi.cu).
int fj_upload_wait(FjsphEngine* e)
{
    if (e->upload_pending)
    {
        FJ_CUDA(cudaStreamWaitEvent(e->stream, e->ev_upload, 0));
        e->upload_pending = false;
    }
    return FJSPH_OK;
}
// ... and for the fields dSPH_PreStep reads (x, rho, m, b), which cross PCIe ahead of the rest (three-part upload)
static int upload_wait_prestep_inputs(FjsphEngine* e)
{
    if (e->upload_pending)
        FJ_CUDA(cudaStreamWaitEvent(e->stream, e->ev_upload_b, 0));
    return FJSPH_OK;
}

int fj_copy_level(FjsphEngine* e, int dst, int src, int which)
{
    const int n = int(e->n);
    KScope ks(e, "copy_level", 1);
    if (dst == 1)
        e->x_moved = true; /* pnp1's positions may differ from the ones the list was built on: pair sweeps take r from x0 */
    k_copy_level<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[src], e->lv[dst], n, which);
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

// Integrator::find_timestep, Integration.cpp:370-443
int fj_find_timestep(FjsphEngine* e, double* dt_out)
{
    const int n = int(e->n_owned);
    const int nb = std::min(fj_blocks(n, TPB), 2048);
    {
        KScope ks(e, "timestep", 2);
        k_timestep_partials<<<nb, TPB, 0, e->stream>>>(e->lv[1], e->blk, e->n_bound_blocks, e->C, n, e->red);
        k_timestep_final<<<1, 7 * 32, 0, e->stream>>>(e->red, nb, e->red_out, e->slab.on && e->slab.dev_reduce);
    }
    FJ_CUDA(cudaGetLastError());
    /* slabs: max over all ranks (the minimum, component 4, travels negated) on the device when the transport can */
    bool reduced = false;
    if (e->slab.on && e->slab.dev_reduce)
    {
        int st = fj_allreduce_dev(e, FJSPH_COMM_MAX, e->red_out, 7, &reduced);
        if (st)
            return st;
    }
    FJ_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (e->slab.on && e->slab.dev_reduce)
        e->h_red[4] = -e->h_red[4];
    if (e->slab.on && !reduced)
    {
        double v[7];
        for (int c = 0; c < 7; ++c) v[c] = (c == 4) ? -e->h_red[c] : e->h_red[c]; /* min as -max(-x) */
        int st = fj_allreduce(e, FJSPH_COMM_MAX, v, 7);
        if (st)
            return st;
        for (int c = 0; c < 7; ++c) e->h_red[c] = (c == 4) ? -v[c] : v[c];
    }
    const FjsphParams& P = e->P;
    const double MEPS = 2.220446049250313e-16;
    e->maxf = std::max(MEPS, std::sqrt(e->h_red[0]));
    e->maxAf = std::max(MEPS, std::sqrt(e->h_red[1]));
    e->maxdrho = std::max(MEPS, e->h_red[2]);
    e->maxRhoi = std::max(MEPS, e->h_red[3]);
    e->minST = std::min(9999999.0, e->h_red[4]);
    e->maxU = std::max(MEPS, std::sqrt(e->h_red[5]));
    if (P.ale)
        e->maxShift = std::max(std::max(e->maxShift, MEPS), std::sqrt(e->h_red[6]));
    e->maxRho_pc = 100 * e->maxRhoi / P.rho_rest;
    double f[6];
    f[0] = 0.25 * std::sqrt(P.H / e->maxf);
    f[1] = 2 * P.H / e->maxU;
    f[2] = 0.125 * P.H_sq * P.rho_rest / P.mu;
    f[3] = 0.067 * e->minST;
    f[4] = 0.5 * std::sqrt(P.H / e->maxdrho);
    f[5] = 1.5 * P.H / P.speed_sound;
    e->safe_dt = 0.75 * *std::min_element(f, f + 6);
    double dt = P.cfl * e->safe_dt;
    if (dt < P.delta_t_min)
        dt = P.delta_t_min;
    else if (dt > P.delta_t_max)
        dt = P.delta_t_max;
    if (dt > P.last_frame_time + P.frame_time_interval - P.current_time)
        dt = P.last_frame_time + P.frame_time_interval - P.current_time + P.delta_t_min;
    if (dt_out)
        *dt_out = dt;
    return FJSPH_OK;
}

// Do_NB_Iter: walls -> forces -> fused update; returns sum |x - x_prev|^2 over the fluid range
int fj_nb_iter(FjsphEngine* e, double npd, double* errsum)
{
    int st = check_supported(e);
    if (st)
        return st;
    st = fj_walls(e, 1, true);
    if (st)
        return st;
    if (e->n_bound_blocks > 0)
    {
        st = fj_halo_exchange(e, 1, FJ_HX_STATE); /* wall rho, p, v seen by the neighbour rank's fluid */
        if (st)
            return st;
    }
    st = fj_forces(e, 1, npd);
    if (st)
        return st;
    const int n = int(e->n_owned);
    const int nb = fj_blocks(n, TPB);
    {
        KScope ks(e, "nb_update", 1);
        if (e->P.ale)
            k_nb_update<true><<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->blk, make_block_table(e),
                                                         e->near_inlet, e->C, e->P.delta_t, e->P.nb_beta,
                                                         e->P.nb_gamma, n, e->red);
        else
            k_nb_update<false><<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->blk, make_block_table(e),
                                                          e->near_inlet, e->C, e->P.delta_t, e->P.nb_beta,
                                                          e->P.nb_gamma, n, e->red);
    }
    FJ_CUDA(cudaGetLastError());
    e->x_moved = true;
    int nparts = nb;
    st = fj_inlet_motion(e, e->P.delta_t, true, &nparts); /* BUFFER particles, Newmark_Beta.cpp:243-297 */
    if (st)
        return st;
    st = fj_halo_exchange(e, 1, FJ_HX_STATE); /* x, v, rho, p of the ghosts for the next sweep */
    if (st)
        return st;
    double s = 0.0;
    st = fj_reduce_sum(e, nparts, 1, &s);
    if (st)
        return st;
    if (errsum)
        *errsum = s;
    return FJSPH_OK;
}

static int frozen_terms(FjsphEngine* e, bool all, bool prestep_done = false)
{
    int st = prestep_done ? FJSPH_OK : fj_prestep(e, nullptr);
    if (st || !all)
        return st;
    st = fj_aero_velocity(e); /* reads and writes particle i only: ahead of the exchange so that the surface sweep's
                                 interior launch follows the exchange directly and runs beside it */
    if (st)
        return st;
    st = fj_halo_exchange(e, 1, FJ_HX_P3); /* gradRho_j, lam_j */
    if (st)
        return st;
    st = fj_surface_and_dissipation(e, true, true, true); /* loops 2+3 fused with particle_shift (ALE) */
    if (st)
        return st;
    st = fj_check_pipe_outlet(e);
    if (st)
        return st;
    return fj_halo_exchange(e, 1, FJ_HX_P2 | FJ_HX_SURFZONE | FJ_HX_B); /* vPert_j; surfzone_j and b_j for the walls */
}

static bool has_density_walls(const FjsphEngine* e)
{
    for (int b = 0; b < e->n_bound_blocks; ++b)
        if (e->blocks[b].bound_solver == FJSPH_DBC || e->blocks[b].bound_solver == FJSPH_GHOST)
            return true;
    return false;
}

static int rk_stage(FjsphEngine* e, double dt_s, double* errsum)
{
    int st = fj_walls(e, 1, false);
    if (st)
        return st;
    const int n = int(e->n_owned);
    const int nb = fj_blocks(n, TPB);
    auto update = [&](int part) {
        KScope ks(e, "rk_update", 1);
        e->x_moved = true;
        if (e->P.ale)
            k_rk_stage<true><<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->blk, make_block_table(e), e->near_inlet,
                                                        e->C, dt_s, n, e->red, part);
        else
            k_rk_stage<false><<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->blk, make_block_table(e),
                                                         e->near_inlet, e->C, dt_s, n, e->red, part);
    };
    if (has_density_walls(e))
        update(0); /* DBC / Ghost wall densities move before the forces see them (Runge_Kutta.cpp:76-131) */
    if (e->n_bound_blocks > 0)
    {
        st = fj_halo_exchange(e, 1, FJ_HX_STATE);
        if (st)
            return st;
    }
    st = fj_forces(e, 1, e->npd);
    if (st)
        return st;
    update(1);
    FJ_CUDA(cudaGetLastError());
    int nparts = nb;
    st = fj_inlet_motion(e, dt_s, false, &nparts); /* Runge_Kutta.cpp:175-228 */
    if (st)
        return st;
    st = fj_halo_exchange(e, 1, FJ_HX_STATE);
    if (st)
        return st;
    if (errsum)
        return fj_reduce_sum(e, nparts, 1, errsum);
    return FJSPH_OK;
}

static int rk_accumulate(FjsphEngine* e, double weight, int first)
{
    const int n = int(e->n_owned);
    KScope ks(e, "rk_update", 1);
    if (e->P.ale)
        k_rk_accumulate<true><<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->rk_sum_v, e->rk_sum_a,
                                                                         weight, first, n);
    else
        k_rk_accumulate<false><<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->rk_sum_v, e->rk_sum_a,
                                                                          weight, first, n);
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

// Integrator::integrate_no_update, Integration.cpp:27-107
int fj_integrate_no_update(FjsphEngine* e, FjsphStepStats* s)
{
    int st = check_supported(e);
    if (st)
        return st;
    if (e->n <= 0)
    {
        fj_set_error("integrate: no particles uploaded");
        return FJSPH_ERR_STATE;
    }
    const long long launches0 = e->launches;
    e->force_evals = 0;
    e->nb_builds = 0;
    const long long skin0 = e->skin_builds;
    e->iteration = 0;
    double rms_error = 0.0, logbase = 0.0;
    e->npd = 1.0;
    /* fluid particles of the whole domain; with slabs it can change at the first neighbour build (migration
     * keeps the global count), so it is read where it is used */
    #define nfluid fj_fluid_count(e)

    bool prestep_done = false;
    if (e->upload_pending)
    {
        /* fjsph_step_host: only the positions are on the device yet.  The list needs nothing else and the time step does
           not depend on the list, so update_neighbours runs first, beside the rest of the upload */
        st = fj_build_neighbours(e);
        if (st)
            return st;
        if (e->upload_early)
        {
            /* three-part upload: x, rho, m and b are here, the rest is still crossing.  dSPH_PreStep reads nothing else and
               neither it nor find_timestep depends on the other, so it runs now, beside the rest of the upload; pn = pnp1
               (Init.cpp:496) follows once everything has landed, for every field but the ones the prestep has just written
               -- those went to pn with their upload-time values before it ran (upload_state_split). */
            st = upload_wait_prestep_inputs(e);
            if (st)
                return st;
            st = fj_prestep(e, nullptr);
            if (st)
                return st;
            prestep_done = true;
        }
        st = fj_upload_wait(e);
        if (st)
            return st;
        if (e->upload_early)
        {
            e->upload_early = false;
            st = fj_copy_level(e, 0, 1, 2);
            if (st)
                return st;
        }
        st = fj_find_timestep(e, &e->P.delta_t);
        if (st)
            return st;
    }
    else
    {
        st = fj_find_timestep(e, &e->P.delta_t);
        if (st)
            return st;
        st = fj_build_neighbours(e);
        if (st)
            return st;
    }

    if (e->P.solver_type == 1)
    {
        /* ---- Runge-Kutta.  Get_First_RK builds st_1 from part_n = part_prev = pn (Runge_Kutta.cpp:462-476),
         * overwriting pnp1, so of the first frozen-term pass only npd (and the neighbour list) survives:
         * the prestep runs, the surface/dissipation/shifting passes are dead work and are skipped. */
        st = frozen_terms(e, false, prestep_done);
        if (st)
            return st;
        st = fj_copy_level(e, 1, 0);
        if (st)
            return st;
        double errsum = 0.0;
        st = rk_stage(e, 0.5 * e->P.delta_t, &errsum);
        if (st)
            return st;
        logbase = std::log10(std::sqrt(errsum / nfluid)); /* Check_RK_Error with logbase 0 */

        st = fj_build_neighbours(e);
        if (st)
            return st;
        st = frozen_terms(e, true);
        if (st)
            return st;
        /* Runge_Kutta4: st_1 = pnp1 (with its fresh frozen terms) */
        st = rk_accumulate(e, 2.0, 1);
        if (st)
            return st;
        st = rk_stage(e, 0.5 * e->P.delta_t, nullptr); /* -> st_2 */
        if (st)
            return st;
        st = rk_accumulate(e, 2.0, 0);
        if (st)
            return st;
        st = rk_stage(e, e->P.delta_t, nullptr); /* -> st_3 */
        if (st)
            return st;
        st = rk_accumulate(e, 1.0, 0);
        if (st)
            return st;
        /* final step on a copy of st_3 */
        st = fj_walls(e, 1, false);
        if (st)
            return st;
        const int n = int(e->n_owned);
        const int nb = fj_blocks(n, TPB);
        auto final_update = [&](int part) {
            KScope ks(e, "rk_update", 1);
            e->x_moved = true;
            k_rk_final<<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->blk, make_block_table(e), e->near_inlet,
                                                  e->rk_sum_v, e->rk_sum_a, e->C, e->P.delta_t, n, e->red, part);
        };
        if (has_density_walls(e))
            final_update(0); /* Runge_Kutta.cpp:276-349: the walls' weighted density step precedes the last forces */
        if (e->n_bound_blocks > 0)
        {
            st = fj_halo_exchange(e, 1, FJ_HX_STATE);
            if (st)
                return st;
        }
        st = fj_forces(e, 1, e->npd);
        if (st)
            return st;
        final_update(1);
        FJ_CUDA(cudaGetLastError());
        int nparts = nb;
        st = fj_inlet_motion(e, e->P.delta_t, false, &nparts); /* Runge_Kutta.cpp:397-452 */
        if (st)
            return st;
        st = fj_halo_exchange(e, 1, FJ_HX_STATE);
        if (st)
            return st;
        st = fj_reduce_sum(e, nparts, 1, &errsum);
        if (st)
            return st;
        rms_error = std::log10(std::sqrt(errsum / nfluid)) - logbase;
    }
    else
    {
        /* ---- Newmark-Beta */
        st = frozen_terms(e, true, prestep_done);
        if (st)
            return st;
        double errsum = 0.0;
        st = fj_nb_iter(e, e->npd, &errsum); /* solve_prestep, Integration.cpp:306-336 */
        if (st)
            return st;
        logbase = std::log10(std::sqrt(errsum / nfluid)); /* Check_Error with iteration == 0 */
        rms_error = 0.0;
        e->iteration++;

        st = fj_build_neighbours(e);
        if (st)
            return st;
        st = frozen_terms(e, true);
        if (st)
            return st;

        /* Newmark_Beta::Newmark_Beta, Newmark_Beta.cpp:303-331 */
        rms_error = 0.0;
        int guard = 0;
        while (rms_error > e->P.min_residual)
        {
            st = fj_nb_iter(e, e->npd, &errsum);
            if (st)
                return st;
            const double log_error = std::log10(std::sqrt(errsum / nfluid));
            if (e->iteration == 0)
                logbase = log_error;
            rms_error = log_error - logbase;
            if (e->iteration > unsigned(e->P.max_subits))
            {
                if (rms_error > 0.0)
                {
                    /* unstable: restore pnp1 = pn, rebuild, halve dt, restart (Newmark_Beta.cpp:32-48) */
                    st = fj_copy_level(e, 1, 0);
                    if (st)
                        return st;
                    st = fj_build_neighbours(e);
                    if (st)
                        return st;
                    e->P.delta_t = 0.5 * e->P.delta_t;
                    e->iteration = 0;
                    rms_error = 0.0;
                    if (++guard > 60)
                    {
                        fj_set_error("Newmark-Beta restart loop did not terminate (dt underflow)");
                        return FJSPH_ERR_STATE;
                    }
                    continue;
                }
                break;
            }
            e->iteration++;
        }
    }
    if (s)
    {
        std::memset(s, 0, sizeof(*s));
        s->dt = e->P.delta_t;
        s->safe_dt = e->safe_dt;
        s->cfl_ratio = e->P.delta_t / e->safe_dt;
        s->rms_error = rms_error;
        s->maxRho_pc = e->maxRho_pc;
        s->maxf = e->maxf;
        s->maxAf = e->maxAf;
        s->maxShift = e->maxShift;
        s->npd = e->npd;
        s->logbase = logbase;
        s->iterations = int(e->iteration);
        s->total_points = int(e->n_owned);
#undef nfluid
        s->force_evals = e->force_evals;
        s->neighbour_builds = e->nb_builds;
        s->skin_builds = int(e->skin_builds - skin0);
        s->kernel_launches = int(e->launches - launches0);
    }
    return FJSPH_OK;
}

// Integrator::integrate, Integration.cpp:233-303 (+ update_data :109-226 without inlets / delete planes)
int fj_step(FjsphEngine* e, FjsphStepStats* s)
{
    FjsphStepStats local;
    FjsphStepStats* ss = s ? s : &local;
    const long long launches0 = e->launches;
    int st = fj_integrate_no_update(e, ss);
    if (st)
        return st;
    int n_add = 0, n_del = 0;
    st = fj_update_data(e, &n_add, &n_del); /* inlet insertion, delete planes; Integration.cpp:109-226 */
    if (st)
        return st;
    st = fj_copy_level(e, 0, 1); /* pn = pnp1 */
    if (st)
        return st;
    ss->n_add = n_add;
    ss->n_del = n_del;
    ss->total_points = int(e->n_owned);
    ss->kernel_launches = int(e->launches - launches0);
    FjsphParams& P = e->P;
    const double step_error = ss->rms_error;
    P.current_time += P.delta_t;
    if (step_error > P.min_residual || e->maxRho_pc > P.rho_max_iter)
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
    if (e->iteration < P.subits_factor * P.max_subits && P.n_unstable == 0)
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
    return FJSPH_OK;
}
