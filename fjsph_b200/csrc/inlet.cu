// inlet.cu — inlet buffer regions and the particle bookkeeping at the end of a step.
//
// Replaces:
//   the inlet part of Newmark_Beta::Do_NB_Iter           reference src/Newmark_Beta.cpp:243-297
//   the inlet part of the Runge-Kutta stage updates      reference src/Runge_Kutta.cpp:175-228,397-452 (fixed-velocity
//                                                        inlets; the dynamic branch indexes out of bounds there)
//   update_buffer_region                                 reference src/shapes/inlet.cpp:578-640
//   Integrator::update_data                              reference src/Integration.cpp:109-226 (IPT hand-off excluded)
// An inlet block ends in one BACK particle per column and n_buf BUFFER particles behind it.  When a BACK particle
// crosses the insertion plane it becomes PIPE, its first buffer particle becomes BACK, the column's buffer list
// shifts and a new BUFFER particle is inserted at the end of the block; particles past a block's delete plane are
// erased.  The per-column decisions are a few hundred plane tests, taken on the host from one small readback exactly
// as the reference takes them; every per-particle operation runs on the device.  The caller's particle order is kept
// the reference's (insert at the block's end, later indices shift up; erase, later indices shift down).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.cuh"
#include "prims.cuh"

namespace
{
constexpr int TPB = 256;

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

__device__ __forceinline__ void write_thermo(Level& S, int s, double rho, double p)
{
    double4 th = S.TH[s];
    th.x = p;
    S.TH[s] = th;
    double4 a = S.P0[s];
    a.w = th.y / rho;
    S.P0[s] = a;
    double4 v = S.P1[s];
    v.w = rho;
    S.P1[s] = v;
    double4 q = S.P2[s];
    q.w = p / (rho * rho);
    S.P2[s] = q;
}

// one thread per (column ii, buffer row jj); tables hold caller indices
// mode 1 (fixed velocity): x = x_back - dx (jj+1) n, v, rho, p copied from the BACK particle
// mode 0 (dynamic, Newmark-Beta): rho = clamp(rho_n + dt (g1 Rrho + g2 Rrho_n)), p = EOS, x = x_n + dt v_n
__global__ void k_inlet_buffers(Level Sn, Level S, const int* __restrict__ slot_of, const int* __restrict__ back,
                                const int* __restrict__ buffer, int n_back, int n_buf, int mode, double nx, double ny,
                                double nz, DevConst C, double dt, double gamma_t1, double* __restrict__ err_partial)
{
    __shared__ double sm[TPB / 32];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (t < n_back * n_buf)
    {
        const int ii = t / n_buf, jj = t % n_buf;
        const int s = slot_of[buffer[ii * n_buf + jj]];
        const double4 xo = S.P0[s];
        double x, y, z;
        if (mode == 1)
        {
            const int sb = slot_of[back[ii]];
            const double4 xb = S.P0[sb], vb = S.P1[sb];
            const double off = C.dx * (jj + 1.0);
            x = xb.x - off * nx;
            y = xb.y - off * ny;
            z = xb.z - off * nz;
            double4 v = S.P1[s];
            v.x = vb.x;
            v.y = vb.y;
            v.z = vb.z;
            S.P1[s] = v;
            write_thermo(S, s, vb.w, S.TH[sb].x);
        }
        else
        {
            const double4 xn = Sn.P0[s], vn = Sn.P1[s];
            const double Rn = Sn.ACC[s].w, R = S.ACC[s].w;
            const double rho = fmax(C.rho_min, fmin(C.rho_max, vn.w + dt * (gamma_t1 * R + (1 - gamma_t1) * Rn)));
            write_thermo(S, s, rho, eos_pressure(C, rho));
            x = xn.x + dt * vn.x;
            y = xn.y + dt * vn.y;
            z = xn.z + dt * vn.z;
        }
        double4 a = S.P0[s];
        a.x = x;
        a.y = y;
        a.z = z;
        S.P0[s] = a;
        const double ex = x - xo.x, ey = y - xo.y, ez = z - xo.z;
        err = ex * ex + ey * ey + ez * ez;
    }
    const double tot = block_sum(err, sm);
    if (threadIdx.x == 0)
        err_partial[blockIdx.x] = tot;
}

__global__ void k_gather_plane(Level S, const int* __restrict__ slot_of, const int* __restrict__ idx, int cnt, double nx,
                               double ny, double nz, double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const double4 p = S.P0[slot_of[idx[k]]];
    /* dot in the reference's order: x nx + y ny + z nz */
    out[k] = p.x * nx + p.y * ny + p.z * nz;
}

__global__ void k_gather_d4(const double4* __restrict__ a, const int* __restrict__ slot_of, const int* __restrict__ idx,
                            int cnt, double4* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt)
        out[k] = a[slot_of[idx[k]]];
}

__global__ void k_set_b(Level S, const int* __restrict__ slot_of, const int* __restrict__ idx, int cnt, int b)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt)
        S.b[slot_of[idx[k]]] = b;
}

// caller indices at or after a block's end move up by the insertions before them
struct ShiftTable
{
    int first[66];
    int shift[66];
    int n;
};
__global__ void k_shift_oidx(int* __restrict__ oidx, int n, ShiftTable T)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n)
        return;
    const int c = oidx[s];
    int sh = 0;
    for (int k = 0; k < T.n; ++k)
        if (c >= T.first[k])
            sh = T.shift[k];
    oidx[s] = c + sh;
}
__global__ void k_fill_slot_of(const int* __restrict__ oidx, int n, int* __restrict__ slot_of)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n)
        slot_of[oidx[s]] = s;
}

// SPHState::insert of the reference (the new record is zero but for v, rho, p, m, cellP, cellRho of `src`)
__global__ void k_insert(Level S, int* __restrict__ oidx, int* __restrict__ blk, const int* __restrict__ slot_of,
                         const int* __restrict__ src_caller, const double* __restrict__ xyz, const int* __restrict__ new_caller,
                         const int* __restrict__ new_blk, const long long* __restrict__ new_pid, int first_slot, int cnt,
                         DevConst C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = first_slot + k;
    const int src = slot_of[src_caller[k]];
    const double4 z = make_double4(0, 0, 0, 0);
    const double4 v = S.P1[src], th = S.TH[src], cv = S.CV[src];
    const double rho = v.w;
    S.P0[s] = make_double4(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], th.y / rho);
    S.P1[s] = v;
    S.P2[s] = make_double4(0, 0, 0, th.x / (rho * rho));
    S.P3[s] = z;
    S.P4[s] = z;
    S.surf_i[s] = 0;
    S.ACC[s] = z;
    S.AF[s] = z;
    S.AV[s] = z;
    S.CV[s] = make_double4(0, 0, 0, cv.w);
    S.NP[s] = z;
    S.BN[s] = z;
    S.TH[s] = make_double4(th.x, th.y, 0.0, th.w);
    S.SC[s] = z;
    S.L0[s] = S.L1[s] = S.L2[s] = S.L3[s] = S.L4[s] = S.L5[s] = S.L6[s] = S.L7[s] = S.L8[s] = 0.0;
    S.part_id[s] = new_pid[k];
    S.cellID[s] = -3;
    S.b[s] = FJSPH_BUFFER;
    S.surfzone[s] = 0;
    S.internal[s] = 0;
    oidx[s] = new_caller[k];
    blk[s] = new_blk[k];
}

// rho < 1e-4 -> rho_rest (Integration.cpp:122-125) and the delete-plane test of every fluid block
struct PlaneTable
{
    double nx[64], ny[64], nz[64], c[64];
    unsigned char on[64];
    int n_bound_blocks;
};
__global__ void k_fix_rho_and_flag_deleted(Level S, const int* __restrict__ blk, const int* __restrict__ oidx, int n,
                                           PlaneTable T, DevConst C, unsigned* __restrict__ del_by_caller)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n)
        return;
    const int bl = blk[s];
    unsigned del = 0;
    if (bl >= T.n_bound_blocks)
    {
        const double rho = S.P1[s].w;
        if (rho < 0.0001)
        {
            const double4 th = S.TH[s];
            write_thermo(S, s, C.rho_rest, th.x);
        }
        const int f = bl - T.n_bound_blocks;
        if (f < 64 && T.on[f])
        {
            const double4 p = S.P0[s];
            del = (p.x * T.nx[f] + p.y * T.ny[f] + p.z * T.nz[f]) > T.c[f];
        }
    }
    del_by_caller[oidx[s]] = del;
}
__global__ void k_survivors(const unsigned* __restrict__ del_by_caller, const unsigned* __restrict__ scan_by_caller,
                            const int* __restrict__ oidx, int n, unsigned* __restrict__ keep_flag, int* __restrict__ new_oidx)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n)
        return;
    const int c = oidx[s];
    keep_flag[s] = !del_by_caller[c];
    new_oidx[s] = c - int(scan_by_caller[c]);
}
__global__ void k_gather_int2(const int* __restrict__ a_in, const int* __restrict__ b_in, const int* __restrict__ list, int cnt,
                              int* __restrict__ a_out, int* __restrict__ b_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt)
    {
        a_out[k] = a_in[list[k]];
        b_out[k] = b_in[list[k]];
    }
}
__global__ void k_map_callers(const unsigned* __restrict__ scan_by_caller, int* __restrict__ idx, int cnt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt)
        idx[k] -= int(scan_by_caller[idx[k]]);
}

// re-decomposition: caller index -> slot -> position among the stayers (the new caller index); bad counts the
// table entries whose particle is leaving the slab
__global__ void k_remap_callers(const int* __restrict__ slot_of, const unsigned* __restrict__ stay_flag,
                                const unsigned* __restrict__ stay_scan, int* __restrict__ idx, int cnt, int* __restrict__ bad)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = slot_of[idx[k]];
    if (!stay_flag[s])
        atomicAdd(bad, 1);
    idx[k] = int(stay_scan[s]);
}
__global__ void k_count_flags(const unsigned* __restrict__ f, int n, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned v = (i < n) ? f[i] : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, v != 0u);
    if ((threadIdx.x & 31) == 0 && m)
        atomicAdd(out, __popc(m));
}

struct InletTables
{
    // device copies of back / buffer (caller indices) of every inlet block, refreshed when they change
    std::vector<int*> d_back, d_buffer;
    std::vector<int> n_back, n_buf;
};

} // namespace

static int upload_tables(FjsphEngine* e)
{
    for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
    {
        HostBlock& B = e->blocks[bl];
        if (B.block_type != FJSPH_INLET_ZONE || B.back.empty())
            continue;
        const size_t nb = B.back.size(), nf = B.buffer[0].size();
        if (!B.d_back)
        {
            FJ_CUDA(cudaMalloc(&B.d_back, nb * sizeof(int)));
            FJ_CUDA(cudaMalloc(&B.d_buffer, nb * nf * sizeof(int)));
        }
        std::vector<int> hb(nb), hf(nb * nf);
        for (size_t i = 0; i < nb; ++i)
        {
            hb[i] = int(B.back[i]);
            for (size_t j = 0; j < nf; ++j) hf[i * nf + j] = int(B.buffer[i][j]);
        }
        FJ_CUDA(cudaMemcpyAsync(B.d_back, hb.data(), nb * sizeof(int), cudaMemcpyHostToDevice, e->stream));
        FJ_CUDA(cudaMemcpyAsync(B.d_buffer, hf.data(), nb * nf * sizeof(int), cudaMemcpyHostToDevice, e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream));
    }
    e->inlet_tables_dirty = false;
    return FJSPH_OK;
}

static int download_tables(FjsphEngine* e, HostBlock& B)
{
    const int nb = int(B.back.size()), nf = int(B.buffer[0].size());
    std::vector<int> hb(nb), hf(size_t(nb) * nf);
    FJ_CUDA(cudaMemcpyAsync(hb.data(), B.d_back, nb * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaMemcpyAsync(hf.data(), B.d_buffer, size_t(nb) * nf * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < nb; ++i)
    {
        B.back[i] = hb[i];
        for (int j = 0; j < nf; ++j) B.buffer[i][j] = hf[size_t(i) * nf + j];
    }
    return FJSPH_OK;
}

int fj_inlet_tables_remap(FjsphEngine* e, const unsigned* d_stay_flag, const unsigned* d_stay_scan)
{
    if (!fj_has_inlets(e))
        return FJSPH_OK;
    if (e->inlet_tables_dirty)
    {
        int st = upload_tables(e);
        if (st)
            return st;
    }
    FJ_CUDA(cudaMemsetAsync(e->d_flag + 1, 0, sizeof(int), e->stream));
    for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
    {
        HostBlock& B = e->blocks[bl];
        if (B.block_type != FJSPH_INLET_ZONE || B.back.empty())
            continue;
        const int nb = int(B.back.size()), nf = int(B.buffer[0].size());
        k_remap_callers<<<fj_blocks(nb, TPB), TPB, 0, e->stream>>>(e->slot_of, d_stay_flag, d_stay_scan, B.d_back, nb, e->d_flag + 1);
        k_remap_callers<<<fj_blocks(nb * nf, TPB), TPB, 0, e->stream>>>(e->slot_of, d_stay_flag, d_stay_scan, B.d_buffer, nb * nf,
                                                                        e->d_flag + 1);
        e->launches += 2;
        int st = download_tables(e, B);
        if (st)
            return st;
    }
    FJ_CUDA(cudaMemcpyAsync(e->h_flag + 1, e->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (e->h_flag[1] != 0)
    {
        fj_set_error("slab %d: %d BACK / BUFFER particle(s) of an inlet block crossed a slab face; an inlet's buffer region "
                     "must lie inside one slab", e->slab.rank, e->h_flag[1]);
        return FJSPH_ERR_STATE;
    }
    return FJSPH_OK;
}

bool fj_has_inlets(FjsphEngine* e)
{
    for (const HostBlock& B : e->blocks)
        if (B.block_type == FJSPH_INLET_ZONE && !B.back.empty())
            return true;
    return false;
}

// Inlet buffer motion after a Newmark-Beta / RK update of `level` (fixed_only: the RK stages).  Adds the buffer
// particles' |x - x_prev|^2 block partials behind the first `nb_partials` entries of e->red; returns the new count.
int fj_inlet_motion(FjsphEngine* e, double dt, bool nb_solver, int* n_partials)
{
    e->x_moved = true;
    if (!fj_has_inlets(e))
        return FJSPH_OK;
    if (e->inlet_tables_dirty)
    {
        int st = upload_tables(e);
        if (st)
            return st;
    }
    for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
    {
        HostBlock& B = e->blocks[bl];
        if (B.block_type != FJSPH_INLET_ZONE || B.back.empty())
            continue;
        const int nb = int(B.back.size()), nf = int(B.buffer[0].size());
        const int mode = B.fixed_vel_or_dynamic == 1 ? 1 : 0;
        if (!nb_solver && mode == 0)
        {
            fj_set_error("Runge-Kutta with a dynamic inlet (fixed_vel_or_dynamic = 0) is not supported: the reference "
                         "indexes limits[] out of bounds there (Runge_Kutta.cpp:211,433)");
            return FJSPH_ERR_INVALID;
        }
        const double nn = std::sqrt(B.insert_norm[0] * B.insert_norm[0] + B.insert_norm[1] * B.insert_norm[1] +
                                    B.insert_norm[2] * B.insert_norm[2]);
        const double inv = nn > 0.0 ? 1.0 / nn : 1.0;
        const int blocks = fj_blocks(int64_t(nb) * nf, TPB);
        KScope ks(e, "inlet_buffers", 1);
        k_inlet_buffers<<<blocks, TPB, 0, e->stream>>>(e->lv[0], e->lv[1], e->slot_of, B.d_back, B.d_buffer, nb, nf, mode,
                                                       B.insert_norm[0] * inv, B.insert_norm[1] * inv, B.insert_norm[2] * inv,
                                                       e->C, dt, e->P.nb_gamma, e->red + *n_partials);
        *n_partials += blocks;
    }
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

// IPT hand-off (Integration.cpp:151-169, Var.h:733-763): a particle past its block's delete plane is "downstream enough
// to convert to IPT" -- the reference builds an IPTPart from it (id, time, position, velocity, mass, cell id, cell velocity and
// density; diameter and area come from the IPT settings) before it erases the SPH particle.  The engine keeps those
// records, in the reference's order (ascending caller index), for the host to collect with fjsph_take_deleted.
__global__ void k_capture_deleted(Level S, const unsigned* __restrict__ del_by_caller, const unsigned* __restrict__ scan_by_caller,
                                  const int* __restrict__ oidx, int n, double t, FjsphDeleted* __restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n)
        return;
    const int c = oidx[s];
    if (!del_by_caller[c])
        return;
    FjsphDeleted d;
    const double4 x = S.P0[s], v = S.P1[s], cv = S.CV[s], th = S.TH[s];
    d.part_id = S.part_id[s];
    d.cellID = S.cellID[s];
    d.t = t;
    d.xi[0] = x.x, d.xi[1] = x.y, d.xi[2] = x.z;
    d.v[0] = v.x, d.v[1] = v.y, d.v[2] = v.z;
    d.mass = th.y;
    d.cellV[0] = cv.x, d.cellV[1] = cv.y, d.cellV[2] = cv.z;
    d.cellRho = th.w;
    out[scan_by_caller[c]] = d;
}

// appends the particles flagged in d_del (by caller index, n_flagged of them, exclusive scan in d_scan) to e->deleted
static int capture_deleted(FjsphEngine* e, const unsigned* d_del, const unsigned* d_scan, int n, int n_flagged)
{
    if (n_flagged <= 0)
        return FJSPH_OK;
    FjsphDeleted* d_out = nullptr;
    FJ_CUDA(cudaMalloc(&d_out, size_t(n_flagged) * sizeof(FjsphDeleted)));
    k_capture_deleted<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1], d_del, d_scan, e->oidx, n, e->P.current_time, d_out);
    e->launches++;
    const size_t at = e->deleted.size();
    e->deleted.resize(at + size_t(n_flagged));
    cudaError_t ce = cudaMemcpyAsync(e->deleted.data() + at, d_out, size_t(n_flagged) * sizeof(FjsphDeleted),
                                     cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess)
        ce = cudaStreamSynchronize(e->stream);
    cudaFree(d_out);
    if (ce != cudaSuccess)
    {
        e->deleted.resize(at);
        return fj_cuda_fail(ce, "capture_deleted", __FILE__, __LINE__);
    }
    return FJSPH_OK;
}

// Erase the particles flagged in d_del (one unsigned per CALLER index, the key scratch array) from pnp1 -- and from pn
// when both_levels -- keeping the reference's order: later indices shift down, block ranges and the inlet tables
// follow (Integration.cpp:171-205, Resid.cpp:483-523).  The neighbour lists become invalid.
int fj_delete_flagged(FjsphEngine* e, unsigned* d_del, bool both_levels, int* n_del_out, bool hand_off)
{
    cudaStream_t st_ = e->stream;
    const int n = int(e->n);
    const int nbb = e->n_bound_blocks;
    unsigned* d_scan = e->rank_in_cell; /* [cap] */
    *n_del_out = 0;
    if (n <= 0)
        return FJSPH_OK;
    if (!e->scan_particles)
        FJ_CUDA(cudaMalloc(&e->scan_particles, (size_t(e->cap) / SCAN_TILE + 2) * sizeof(unsigned)));
    /* exclusive scan in caller order: the arrays hold cap entries, so scan n-1 flags and add the last flag */
    prim_exclusive_scan(st_, d_del, d_scan, unsigned(n - 1), e->scan_particles);
    unsigned h_tail[2] = {0, 0};
    FJ_CUDA(cudaMemcpyAsync(&h_tail[0], d_scan + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    FJ_CUDA(cudaMemcpyAsync(&h_tail[1], d_del + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    FJ_CUDA(cudaStreamSynchronize(st_));
    e->launches += 3;
    const int n_del = int(h_tail[0] + h_tail[1]);
    if (n_del == 0)
        return FJSPH_OK;
    if (hand_off)
    {
        int st = capture_deleted(e, d_del, d_scan, n, n_del);
        if (st)
            return st;
    }
    KScope ks(e, "delete_particles", 8);
    /* survivors in slot order; their new caller index = old - (#deleted before it) */
    unsigned* d_keep = reinterpret_cast<unsigned*>(e->perm);
    unsigned* d_keep_scan = reinterpret_cast<unsigned*>(e->perm2);
    int* d_new_oidx = e->oidx_tmp;
    k_survivors<<<fj_blocks(n, TPB), TPB, 0, st_>>>(d_del, d_scan, e->oidx, n, d_keep, d_new_oidx);
    prim_exclusive_scan(st_, d_keep, d_keep_scan, unsigned(n - 1), e->scan_particles);
    int* d_list = e->near_inlet; /* [cap] scratch: surviving slots */
    k_compact<<<fj_blocks(n, TPB), TPB, 0, st_>>>(d_keep, d_keep_scan, n, d_list);
    const int n_new = n - n_del;
    k_permute_level<<<fj_blocks(n_new, PRIM_TPB), PRIM_TPB, 0, st_>>>(e->lv[1], e->lv[2], d_list, n_new);
    std::swap(e->lv[1], e->lv[2]);
    if (both_levels)
    {
        k_permute_level<<<fj_blocks(n_new, PRIM_TPB), PRIM_TPB, 0, st_>>>(e->lv[0], e->lv[2], d_list, n_new);
        std::swap(e->lv[0], e->lv[2]);
    }
    k_gather_int2<<<fj_blocks(n_new, TPB), TPB, 0, st_>>>(d_new_oidx, e->blk, d_list, n_new, e->oidx, e->blk_tmp);
    std::swap(e->blk, e->blk_tmp);
    /* per-block counts and the inlet tables in the new numbering */
    for (size_t bl = 0; bl < e->blocks.size(); ++bl)
    {
        HostBlock& B = e->blocks[bl];
        unsigned a = 0, b2 = 0;
        const int64_t f = std::min<int64_t>(B.first, n - 1), s2 = std::min<int64_t>(B.second, n - 1);
        FJ_CUDA(cudaMemcpyAsync(&a, d_scan + f, sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
        FJ_CUDA(cudaMemcpyAsync(&b2, d_scan + s2, sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
        FJ_CUDA(cudaStreamSynchronize(st_));
        const unsigned before_first = (B.first >= n) ? unsigned(n_del) : a;
        const unsigned before_second = (B.second >= n) ? unsigned(n_del) : b2;
        if (int(bl) >= nbb && B.block_type == FJSPH_INLET_ZONE && !B.back.empty())
        {
            if (e->inlet_tables_dirty)
            {
                int st = upload_tables(e);
                if (st)
                    return st;
            }
            const int nb = int(B.back.size()), nf = int(B.buffer[0].size());
            k_map_callers<<<fj_blocks(nb, TPB), TPB, 0, st_>>>(d_scan, B.d_back, nb);
            k_map_callers<<<fj_blocks(nb * nf, TPB), TPB, 0, st_>>>(d_scan, B.d_buffer, nb * nf);
            std::vector<int> hb(nb), hf(size_t(nb) * nf);
            FJ_CUDA(cudaMemcpyAsync(hb.data(), B.d_back, nb * sizeof(int), cudaMemcpyDeviceToHost, st_));
            FJ_CUDA(cudaMemcpyAsync(hf.data(), B.d_buffer, size_t(nb) * nf * sizeof(int), cudaMemcpyDeviceToHost, st_));
            FJ_CUDA(cudaStreamSynchronize(st_));
            for (int i = 0; i < nb; ++i)
            {
                B.back[i] = hb[i];
                for (int j = 0; j < nf; ++j) B.buffer[i][j] = hf[size_t(i) * nf + j];
            }
        }
        B.first -= before_first;
        B.second -= before_second;
    }
    e->bound_points = e->n_bound_blocks > 0 ? e->blocks[size_t(e->n_bound_blocks) - 1].second : 0;
    e->n = n_new;
    e->n_owned = n_new;
    k_fill_slot_of<<<fj_blocks(n_new, TPB), TPB, 0, st_>>>(e->oidx, n_new, e->slot_of);
    FJ_CUDA(cudaGetLastError());
    FJ_CUDA(cudaStreamSynchronize(st_));
    e->skin_valid = false;
    e->list_valid = false;
    *n_del_out = n_del;
    return FJSPH_OK;
}

// Integrator::update_data without the final pn = pnp1 (the caller copies the level): insertions, the rho fix,
// deletions, and a neighbour rebuild when the particle set changed.
int fj_update_data(FjsphEngine* e, int* n_add_out, int* n_del_out)
{
    *n_add_out = *n_del_out = 0;
    const int nbb = e->n_bound_blocks;
    Level& S = e->lv[1];
    cudaStream_t st_ = e->stream;
    bool any_delete_plane = false;
    for (size_t bl = size_t(nbb); bl < e->blocks.size(); ++bl) any_delete_plane |= e->blocks[bl].delconst != 9999999.0;
    /* Slab decomposition: every rank carries the same block list (the tables of an inlet block only on the rank that
       holds its buffer region), so the ranks agree on whether this bookkeeping runs and meet at its all-reduces.  The
       reference's particle ORDER has no meaning across ranks: new particles are appended behind the owned ones, the
       ghosts are dropped, erased particles vanish with the forced re-decomposition that follows. */
    const bool slabs = e->slab.on && e->slab.world > 1;
    bool any_inlet_block = false;
    for (size_t bl = size_t(nbb); bl < e->blocks.size(); ++bl) any_inlet_block |= e->blocks[bl].block_type == FJSPH_INLET_ZONE;
    if (!(slabs ? any_inlet_block : fj_has_inlets(e)) && !any_delete_plane)
        return FJSPH_OK;
    if (e->inlet_tables_dirty)
    {
        int st = upload_tables(e);
        if (st)
            return st;
    }

    // ---- update_buffer_region, shapes/inlet.cpp:578-640
    std::vector<int> to_pipe, to_back, src_caller, new_caller, new_blk;
    std::vector<double> new_xyz;
    std::vector<long long> new_pid;
    ShiftTable shifts;
    shifts.n = 0;
    int total_shift = 0;
    for (size_t bl = size_t(nbb); bl < e->blocks.size(); ++bl)
    {
        HostBlock& B = e->blocks[bl];
        B.first += total_shift;
        B.second += total_shift;
        if (B.block_type != FJSPH_INLET_ZONE || B.back.empty())
            continue;
        /* caller indices stored in the tables move with earlier blocks' insertions */
        if (total_shift)
        {
            for (auto& x : B.back) x += total_shift;
            for (auto& col : B.buffer)
                for (auto& x : col) x += total_shift;
            e->inlet_tables_dirty = true;
        }
        const int nb = int(B.back.size()), nf = int(B.buffer[0].size());
        const int pre = total_shift; /* host tables are already in the new numbering, the device still in the old */
        /* plane value of every BACK particle and the position of every column's last buffer particle */
        std::vector<double> h_dot(nb);
        double* d_dot = e->red; /* scratch: nb doubles */
        k_gather_plane<<<fj_blocks(nb, TPB), TPB, 0, st_>>>(S, e->slot_of, B.d_back, nb, B.insert_norm[0], B.insert_norm[1],
                                                            B.insert_norm[2], d_dot);
        e->launches++;
        FJ_CUDA(cudaMemcpyAsync(h_dot.data(), d_dot, nb * sizeof(double), cudaMemcpyDeviceToHost, st_));
        FJ_CUDA(cudaStreamSynchronize(st_));
        int block_add = 0;
        std::vector<int> cols;
        for (int ii = 0; ii < nb; ++ii)
            if (h_dot[ii] > B.insconst)
                cols.push_back(ii);
        if (cols.empty())
            continue;
        /* positions of the source particles (the columns' last buffer particle before the shift) */
        std::vector<double4> h_src(cols.size());
        {
            std::vector<int> h_idx(cols.size());
            for (size_t k = 0; k < cols.size(); ++k) h_idx[k] = int(B.buffer[cols[k]].back()) - pre;
            int* d_idx = reinterpret_cast<int*>(e->stage);
            double4* d_pos = reinterpret_cast<double4*>(reinterpret_cast<char*>(e->stage) + ((cols.size() * 4 + 255) & ~size_t(255)));
            FJ_CUDA(cudaMemcpyAsync(d_idx, h_idx.data(), h_idx.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
            k_gather_d4<<<fj_blocks(cols.size(), TPB), TPB, 0, st_>>>(S.P0, e->slot_of, d_idx, int(cols.size()), d_pos);
            e->launches++;
            FJ_CUDA(cudaMemcpyAsync(h_src.data(), d_pos, h_src.size() * sizeof(double4), cudaMemcpyDeviceToHost, st_));
            FJ_CUDA(cudaStreamSynchronize(st_));
        }
        for (size_t k = 0; k < cols.size(); ++k)
        {
            const int ii = cols[k];
            to_pipe.push_back(int(B.back[ii]) - pre);
            to_back.push_back(int(B.buffer[ii][0]) - pre);
            B.back[ii] = B.buffer[ii][0];
            const int64_t src = B.buffer[ii].back();
            for (int jj = 0; jj + 1 < nf; ++jj) B.buffer[ii][jj] = B.buffer[ii][jj + 1];
            if (e->n + int64_t(new_caller.size()) + 1 > e->cap)
            {
                fj_set_error("inlet insertion: capacity %lld exhausted (reference: total_points < max_points)",
                             (long long)e->cap);
                return FJSPH_ERR_CAPACITY;
            }
            src_caller.push_back(int(src) - pre);
            new_xyz.push_back(h_src[k].x - e->P.dx * B.insert_norm[0]);
            new_xyz.push_back(h_src[k].y - e->P.dx * B.insert_norm[1]);
            new_xyz.push_back(h_src[k].z - e->P.dx * B.insert_norm[2]);
            new_blk.push_back(int(bl));
            if (slabs)
            {
                /* appended behind the owned particles; the ids come from the global counter below */
                const int64_t c = e->n_owned + int64_t(new_caller.size());
                new_caller.push_back(int(c));
                B.buffer[ii].back() = c;
                continue;
            }
            new_caller.push_back(int(B.second));
            new_pid.push_back(e->next_part_id++);
            B.buffer[ii].back() = B.second;
            B.second++;
            block_add++;
        }
        e->inlet_tables_dirty = true;
        if (slabs)
            continue;
        total_shift += block_add;
        if (shifts.n >= 66)
        {
            /* the index-shift table is full: later blocks would be remapped wrongly -- an error, not a silent drop */
            fj_set_error("update_data: more than 66 inlet blocks insert particles in one step (index-shift table full)");
            return FJSPH_ERR_CAPACITY;
        }
        {
            /* callers at or after this block's OLD end (in the pre-insertion numbering) move by total_shift */
            shifts.first[shifts.n] = int(B.second - block_add - (total_shift - block_add));
            shifts.shift[shifts.n] = total_shift;
            shifts.n++;
        }
    }
    const int n_add = int(new_caller.size());
    double n_add_global = n_add;
    if (slabs)
    {
        /* globally unique particle ids: rank r takes [next + sum_{q<r} adds_q, ...) */
        std::vector<double> adds(size_t(e->slab.world), 0.0);
        adds[size_t(e->slab.rank)] = n_add;
        int st = fj_allreduce(e, FJSPH_COMM_SUM, adds.data(), e->slab.world);
        if (st)
            return st;
        long long before = 0, total = 0;
        for (int r = 0; r < e->slab.world; ++r)
        {
            if (r < e->slab.rank)
                before += (long long)adds[size_t(r)];
            total += (long long)adds[size_t(r)];
        }
        for (int k = 0; k < n_add; ++k) new_pid.push_back(e->next_part_id + before + k);
        e->next_part_id += total;
        n_add_global = double(total);
    }
    if (!to_pipe.empty())
    {
        KScope ks(e, "inlet_insert", 5);
        /* slab mode: the new particles take the slots behind the OWNED ones; this rank's ghosts are dropped, which is
           safe because an insertion anywhere makes every rank re-decompose below */
        const int n_old = slabs ? int(e->n_owned) : int(e->n);
        /* flags first (slot_of still in the old numbering) */
        int* d_a = reinterpret_cast<int*>(e->stage);
        const size_t m = to_pipe.size();
        FJ_CUDA(cudaMemcpyAsync(d_a, to_pipe.data(), m * sizeof(int), cudaMemcpyHostToDevice, st_));
        FJ_CUDA(cudaMemcpyAsync(d_a + m, to_back.data(), m * sizeof(int), cudaMemcpyHostToDevice, st_));
        k_set_b<<<fj_blocks(m, TPB), TPB, 0, st_>>>(S, e->slot_of, d_a, int(m), FJSPH_PIPE);
        k_set_b<<<fj_blocks(m, TPB), TPB, 0, st_>>>(S, e->slot_of, d_a + m, int(m), FJSPH_BACK);
        if (n_add > 0)
        {
            /* staging: src_caller | new_caller | new_blk | pid | xyz */
            char* base = reinterpret_cast<char*>(e->stage) + ((2 * m * sizeof(int) + 255) & ~size_t(255));
            int* d_src = reinterpret_cast<int*>(base);
            int* d_newc = d_src + n_add;
            int* d_newb = d_newc + n_add;
            long long* d_pid = reinterpret_cast<long long*>(base + ((3 * size_t(n_add) * sizeof(int) + 255) & ~size_t(255)));
            double* d_xyz = reinterpret_cast<double*>(d_pid + n_add);
            FJ_CUDA(cudaMemcpyAsync(d_src, src_caller.data(), n_add * sizeof(int), cudaMemcpyHostToDevice, st_));
            FJ_CUDA(cudaMemcpyAsync(d_newc, new_caller.data(), n_add * sizeof(int), cudaMemcpyHostToDevice, st_));
            FJ_CUDA(cudaMemcpyAsync(d_newb, new_blk.data(), n_add * sizeof(int), cudaMemcpyHostToDevice, st_));
            FJ_CUDA(cudaMemcpyAsync(d_pid, new_pid.data(), n_add * sizeof(long long), cudaMemcpyHostToDevice, st_));
            FJ_CUDA(cudaMemcpyAsync(d_xyz, new_xyz.data(), 3 * size_t(n_add) * sizeof(double), cudaMemcpyHostToDevice, st_));
            /* copy from the sources through the OLD slot_of, then renumber the callers */
            k_insert<<<fj_blocks(n_add, TPB), TPB, 0, st_>>>(S, e->oidx, e->blk, e->slot_of, d_src, d_xyz, d_newc, d_newb, d_pid,
                                                             n_old, n_add, e->C);
            if (!slabs)
                k_shift_oidx<<<fj_blocks(n_old, TPB), TPB, 0, st_>>>(e->oidx, n_old, shifts);
            e->n_owned += n_add;
            e->n = slabs ? e->n_owned : e->n + n_add;
            k_fill_slot_of<<<fj_blocks(e->n, TPB), TPB, 0, st_>>>(e->oidx, int(e->n), e->slot_of);
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaStreamSynchronize(st_));
    }

    // ---- rho fix + delete planes, Integration.cpp:122-205
    int n_del = 0;
    {
        PlaneTable T;
        std::memset(&T, 0, sizeof(T));
        T.n_bound_blocks = nbb;
        for (size_t bl = size_t(nbb); bl < e->blocks.size() && bl - nbb < 64; ++bl)
        {
            const HostBlock& B = e->blocks[bl];
            const int f = int(bl) - nbb;
            T.on[f] = B.delconst != 9999999.0;
            T.nx[f] = B.delete_norm[0];
            T.ny[f] = B.delete_norm[1];
            T.nz[f] = B.delete_norm[2];
            T.c[f] = B.delconst;
        }
        const int n = slabs ? int(e->n_owned) : int(e->n);
        unsigned* d_del = reinterpret_cast<unsigned*>(e->key);       /* [cap] scratch, by caller index */
        unsigned* d_scan = e->rank_in_cell;                           /* [cap] */
        {
            KScope ks(e, "delete_plane", 4);
            k_fix_rho_and_flag_deleted<<<fj_blocks(n, TPB), TPB, 0, st_>>>(S, e->blk, e->oidx, n, T, e->C, d_del);
        }
        if (any_delete_plane && !slabs)
        {
            int st = fj_delete_flagged(e, d_del, false, &n_del, true); /* pn = pnp1 follows (Integration.cpp:220-223) */
            if (st)
                return st;
        }
        if (slabs)
        {
            double dels = 0.0;
            if (any_delete_plane)
            {
                FJ_CUDA(cudaMemsetAsync(e->d_flag + 1, 0, sizeof(int), st_));
                k_count_flags<<<fj_blocks(n, TPB), TPB, 0, st_>>>(d_del, n, e->d_flag + 1);
                e->launches++;
                FJ_CUDA(cudaMemcpyAsync(e->h_flag + 1, e->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, st_));
                FJ_CUDA(cudaStreamSynchronize(st_));
                dels = double(e->h_flag[1]);
                if (e->h_flag[1] > 0)
                {
                    /* this rank's erased particles, for the IPT hand-off (order: ascending caller index on this rank) */
                    if (!e->scan_particles)
                        FJ_CUDA(cudaMalloc(&e->scan_particles, (size_t(e->cap) / SCAN_TILE + 2) * sizeof(unsigned)));
                    prim_exclusive_scan(st_, d_del, d_scan, unsigned(n), e->scan_particles);
                    int stc = capture_deleted(e, d_del, d_scan, n, e->h_flag[1]);
                    if (stc)
                        return stc;
                }
            }
            int st = fj_allreduce(e, FJSPH_COMM_SUM, &dels, 1);
            if (st)
                return st;
            *n_add_out = int(n_add_global);
            *n_del_out = int(dels);
            if (n_add_global > 0.0 || dels > 0.0)
            {
                /* all ranks re-decompose together: erased particles belong to no class of its compaction, the ghost sets
                   (dropped above) and the global counts are re-made */
                if (dels > 0.0)
                    e->slab.del_by_caller = d_del;
                e->skin_valid = false;
                e->list_valid = false;
                st = fj_copy_level(e, 0, 1); /* pn = pnp1 before the particles move ranks (Integration.cpp:220-223) */
                if (st)
                    return st;
                return fj_build_neighbours(e);
            }
            return FJSPH_OK;
        }
    }
    *n_add_out = n_add;
    *n_del_out = n_del;
    if (n_add || n_del)
    {
        e->skin_valid = false;
        e->list_valid = false;
        return fj_build_neighbours(e); /* Integration.cpp:207-210 */
    }
    return FJSPH_OK;
}
