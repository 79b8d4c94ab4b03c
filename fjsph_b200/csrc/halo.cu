// halo.cu — 1-D slab decomposition of the WCSPH step over the GPUs of one box.
//
// The reference is a single shared-memory process (OpenMP, reference src/FJSPH.cpp:62); this file is what lets
// `Integrator::integrate` (reference src/Integration.cpp:233-303) run as one engine per GPU on x-slabs:
//   * re-decomposition (only when the neighbour superset list is rebuilt): owned particles that left the slab
//     migrate to the neighbour rank with their full SPHPart state of both time levels; then every rank sends its
//     owned particles within 2H + skin of a face to that neighbour as ghosts (appended behind the owned ones);
//   * forward exchange (between dependent sweeps and after every Newmark-Beta / RK update): the SAME ghost set
//     gets the selected 32-byte records refreshed, packed and unpacked on the device;
//   * the step's scalars (npd, residual sums, find_timestep maxima, particle counts, the skin displacement)
//     are all-reduced so every rank takes identical decisions.
// Transport is the host's (FjsphCommFn, include/fjsph_b200.h): NCCL send/recv in fjsph_b200/slab.py.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.cuh"
#include "prims.cuh"

namespace
{
constexpr int TPB = 256;

__host__ __device__ inline size_t align32(size_t x) { return (x + 31) & ~size_t(31); }

// bytes of one whole level + block id for cnt particles (field-major, each array 32-byte aligned)
size_t level_bytes(size_t cnt)
{
    size_t off = 0;
#define X(T, f) off += align32(cnt * sizeof(T));
    FJ_LEVEL_FIELDS(X)
#undef X
    off += align32(cnt * sizeof(int)); // blk
    return off;
}

// del_by_caller (may be null): particles flagged there belong to no class and vanish with the re-decomposition
__global__ void k_classify(const double4* __restrict__ P0, int n_owned, double x_lo, double x_hi,
                           const unsigned* __restrict__ del_by_caller, const int* __restrict__ oidx,
                           unsigned* __restrict__ f_stay, unsigned* __restrict__ f_lo, unsigned* __restrict__ f_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned)
        return;
    const double x = P0[i].x;
    const unsigned keep = (del_by_caller && del_by_caller[oidx[i]]) ? 0u : 1u;
    const unsigned lo = x < x_lo, hi = !(x < x_hi);
    f_lo[i] = keep & lo;
    f_hi[i] = keep & (hi && !lo);
    f_stay[i] = keep & !(lo || hi);
}

__global__ void k_ghost_flags(const double4* __restrict__ P0, int n_owned, double x_ghost_lo, double x_ghost_hi,
                              unsigned* __restrict__ f_lo, unsigned* __restrict__ f_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned)
        return;
    const double x = P0[i].x;
    f_lo[i] = x < x_ghost_lo;
    f_hi[i] = !(x < x_ghost_hi);
}

// whole-level pack / unpack (migration, ghost creation)
__global__ void k_pack_level(Level L, const int* __restrict__ blk, const int* __restrict__ slots, int cnt,
                             char* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = slots[k];
    size_t off = 0;
#define X(T, f)                                   \
    reinterpret_cast<T*>(buf + off)[k] = L.f[s];  \
    off += align32(size_t(cnt) * sizeof(T));
    FJ_LEVEL_FIELDS(X)
#undef X
    reinterpret_cast<int*>(buf + off)[k] = blk[s];
}
__global__ void k_unpack_level(Level L, int* __restrict__ blk, int first, int cnt, const char* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = first + k;
    size_t off = 0;
#define X(T, f)                                         \
    L.f[s] = reinterpret_cast<const T*>(buf + off)[k];  \
    off += align32(size_t(cnt) * sizeof(T));
    FJ_LEVEL_FIELDS(X)
#undef X
    if (blk)
        blk[s] = reinterpret_cast<const int*>(buf + off)[k];
}
__global__ void k_gather_int(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ list, int cnt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt)
        out[k] = in[list[k]];
}
__global__ void k_identity(int* __restrict__ oidx, int* __restrict__ slot_of, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        oidx[i] = i;
        slot_of[i] = i;
    }
}
__global__ void k_count_fluid(const int* __restrict__ blk, int n_bound_blocks, int n, unsigned* __restrict__ out)
{
    // grid-stride count of owned fluid particles (rare: once per re-decomposition)
    unsigned c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += blk[i] >= n_bound_blocks;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(out, c);
}

// forward exchange: selected records of the fixed ghost set
struct D4Table
{
    double4* p[13];
};
__global__ void k_pack_fields(D4Table T, const int* __restrict__ surfzone, const int* __restrict__ b, unsigned mask,
                              const int* __restrict__ send_idx, const int* __restrict__ slot_of, int cnt,
                              char* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = slot_of[send_idx[k]];
    size_t off = 0;
#pragma unroll
    for (int a = 0; a < 13; ++a)
        if (mask & (1u << a))
        {
            reinterpret_cast<double4*>(buf + off)[k] = T.p[a][s];
            off += size_t(cnt) * sizeof(double4);
        }
    if (mask & FJ_HX_SURFZONE)
    {
        reinterpret_cast<int*>(buf + off)[k] = surfzone[s];
        off += align32(size_t(cnt) * sizeof(int));
    }
    if (mask & FJ_HX_B)
        reinterpret_cast<int*>(buf + off)[k] = b[s];
}
__global__ void k_unpack_fields(D4Table T, int* __restrict__ surfzone, int* __restrict__ b, int* __restrict__ surf_i,
                                unsigned mask, int first_caller, const int* __restrict__ slot_of, int cnt,
                                const char* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const int s = slot_of[first_caller + k];
    size_t off = 0;
#pragma unroll
    for (int a = 0; a < 13; ++a)
        if (mask & (1u << a))
        {
            const double4 rec = reinterpret_cast<const double4*>(buf + off)[k];
            T.p[a][s] = rec;
            if (a == 4)
                surf_i[s] = (rec.w != 0.0) ? 1 : 0; /* the int mirror of P4's surf flag (engine.cuh) */
            off += size_t(cnt) * sizeof(double4);
        }
    if (mask & FJ_HX_SURFZONE)
    {
        surfzone[s] = reinterpret_cast<const int*>(buf + off)[k];
        off += align32(size_t(cnt) * sizeof(int));
    }
    if (mask & FJ_HX_B)
        b[s] = reinterpret_cast<const int*>(buf + off)[k];
}

size_t fields_bytes(unsigned mask, size_t cnt)
{
    size_t off = 0;
    for (int a = 0; a < 13; ++a)
        if (mask & (1u << a))
            off += cnt * sizeof(double4);
    if (mask & FJ_HX_SURFZONE)
        off += align32(cnt * sizeof(int));
    if (mask & FJ_HX_B)
        off += align32(cnt * sizeof(int));
    return off;
}

D4Table table_of(Level& L)
{
    D4Table T;
    int a = 0;
#define X(f) T.p[a++] = L.f;
    FJ_D4_FIELDS(X)
#undef X
    return T;
}

int ensure_buffers(FjsphEngine* e, size_t bytes)
{
    Slab& S = e->slab;
    if (bytes <= S.buf_bytes)
        return FJSPH_OK;
    size_t cap = std::max<size_t>(bytes + bytes / 4, size_t(1) << 20);
    for (int s = 0; s < 2; ++s)
    {
        if (S.sbuf[s])
            cudaFree(S.sbuf[s]);
        if (S.rbuf[s])
            cudaFree(S.rbuf[s]);
        S.sbuf[s] = S.rbuf[s] = nullptr;
        FJ_CUDA(cudaMalloc(&S.sbuf[s], cap));
        FJ_CUDA(cudaMalloc(&S.rbuf[s], cap));
    }
    S.buf_bytes = cap;
    return FJSPH_OK;
}

int comm(FjsphEngine* e, int op, void* a, int64_t na, void* b, int64_t nb, void* c, int64_t nc, void* d, int64_t nd)
{
    Slab& S = e->slab;
    const int rc = S.fn(S.user, op, a, na, b, nb, c, nc, d, nd);
    if (rc != 0)
    {
        fj_set_error("slab exchange callback failed (op %d, rc %d)", op, rc);
        return FJSPH_ERR_STATE;
    }
    return FJSPH_OK;
}

// counts to / from the two neighbours through the host path
int exchange_counts(FjsphEngine* e, const int64_t send[2], int64_t recv[2])
{
    Slab& S = e->slab;
    int64_t s_lo = send[0], s_hi = send[1], r_lo = 0, r_hi = 0;
    const bool has_lo = S.rank > 0, has_hi = S.rank < S.world - 1;
    int st = comm(e, FJSPH_COMM_SENDRECV_HOST, &s_lo, has_lo ? 8 : 0, &s_hi, has_hi ? 8 : 0, &r_lo, has_lo ? 8 : 0, &r_hi,
                  has_hi ? 8 : 0);
    if (st)
        return st;
    recv[0] = has_lo ? r_lo : 0;
    recv[1] = has_hi ? r_hi : 0;
    return FJSPH_OK;
}

} // namespace

int fj_allreduce(FjsphEngine* e, int op, double* v, int n)
{
    if (!e->slab.on || e->slab.world == 1)
        return FJSPH_OK;
    return comm(e, op, v, int64_t(n) * 8, nullptr, 0, nullptr, 0, nullptr, 0);
}

int fj_allreduce_dev(FjsphEngine* e, int op, double* d_v, int n, bool* done)
{
    *done = false;
    if (!e->slab.on || e->slab.world == 1)
    {
        *done = true; /* nothing to reduce over */
        return FJSPH_OK;
    }
    if (!e->slab.dev_reduce)
        return FJSPH_OK;
    int st = comm(e, op == FJSPH_COMM_SUM ? FJSPH_COMM_SUM_DEV : FJSPH_COMM_MAX_DEV, d_v, int64_t(n) * 8, nullptr, 0, nullptr, 0,
                  nullptr, 0);
    if (st)
        return st;
    *done = true;
    return FJSPH_OK;
}

double fj_fluid_count(FjsphEngine* e)
{
    return e->slab.on ? e->slab.n_fluid_global : double(e->n_owned - e->bound_points);
}
double fj_total_count(FjsphEngine* e) { return e->slab.on ? e->slab.n_total_global : double(e->n_owned); }

int fj_halo_wait(FjsphEngine* e)
{
    Slab& S = e->slab;
    if (!S.pending)
        return FJSPH_OK;
    S.pending = false;
    FJ_CUDA(cudaStreamWaitEvent(e->stream, S.ev_done, 0));
    return FJSPH_OK;
}

// Forward exchange of the fields in `mask` for the current ghost set.  The pack kernels, the transport (NCCL send/recv,
// FJSPH_COMM_SENDRECV_DEV_ASYNC: ordered on comm_stream) and the unpack kernels are queued on comm_stream behind
// everything the main stream has launched so far; the main stream picks the result up at the next fj_halo_wait --
// issued by KScope before any kernel family, except the interior launches of the split sweeps (sweeps.cu).
int fj_halo_exchange(FjsphEngine* e, int level, unsigned mask)
{
    Slab& S = e->slab;
    if (!S.on || S.world == 1)
        return FJSPH_OK;
    int st = fj_halo_wait(e); /* one exchange in flight at a time: the buffers are shared */
    if (st)
        return st;
    Level& L = e->lv[level];
    const size_t bs[2] = {fields_bytes(mask, size_t(S.n_send[0])), fields_bytes(mask, size_t(S.n_send[1]))};
    const size_t br[2] = {fields_bytes(mask, size_t(S.n_recv[0])), fields_bytes(mask, size_t(S.n_recv[1]))};
    if (std::max(std::max(bs[0], bs[1]), std::max(br[0], br[1])) > S.buf_bytes)
    {
        FJ_CUDA(cudaStreamSynchronize(S.comm_stream)); /* the old buffers may still be in use */
        st = ensure_buffers(e, std::max(std::max(bs[0], bs[1]), std::max(br[0], br[1])));
        if (st)
            return st;
    }
    cudaStream_t cs = S.overlap ? S.comm_stream : e->stream;
    if (S.overlap)
    {
        FJ_CUDA(cudaEventRecord(S.ev_ready, e->stream));
        FJ_CUDA(cudaStreamWaitEvent(cs, S.ev_ready, 0));
    }
    D4Table T = table_of(L);
    e->launches += 4;
    for (int s = 0; s < 2; ++s)
        if (S.n_send[s] > 0)
            k_pack_fields<<<fj_blocks(S.n_send[s], TPB), TPB, 0, cs>>>(T, L.surfzone, L.b, mask, S.send_idx[s], e->slot_of,
                                                                       int(S.n_send[s]), S.sbuf[s]);
    FJ_CUDA(cudaGetLastError());
    st = comm(e, S.overlap ? FJSPH_COMM_SENDRECV_DEV_ASYNC : FJSPH_COMM_SENDRECV_DEV, S.sbuf[0], int64_t(bs[0]), S.sbuf[1],
              int64_t(bs[1]), S.rbuf[0], int64_t(br[0]), S.rbuf[1], int64_t(br[1]));
    if (st)
        return st;
    int first = int(e->n_owned);
    for (int s = 0; s < 2; ++s)
    {
        if (S.n_recv[s] > 0)
            k_unpack_fields<<<fj_blocks(S.n_recv[s], TPB), TPB, 0, cs>>>(T, L.surfzone, L.b, L.surf_i, mask, first, e->slot_of,
                                                                         int(S.n_recv[s]), S.rbuf[s]);
        first += int(S.n_recv[s]);
    }
    FJ_CUDA(cudaGetLastError());
    if (S.overlap)
    {
        FJ_CUDA(cudaEventRecord(S.ev_done, cs));
        S.pending = true;
    }
    S.exchanges++;
    S.bytes_sent += (long long)(bs[0] + bs[1]);
    return FJSPH_OK;
}

// Migration + ghost rebuild.  On return the owned particles occupy slots [0, n_owned) in a fresh caller order
// (identity), the ghosts follow in [n_owned, n), both time levels are consistent, and the send lists for the
// forward exchanges are set.  Called right before the neighbour superset list is rebuilt.
int fj_redecompose(FjsphEngine* e)
{
    Slab& S = e->slab;
    if (!S.on || S.world == 1)
        return FJSPH_OK;
    const bool has_lo = S.rank > 0, has_hi = S.rank < S.world - 1;
    const int n0 = int(e->n_owned);
    cudaStream_t st_ = e->stream;
    int st = fj_halo_wait(e);
    if (st)
        return st;

    // ---- 1. classify the owned particles of pnp1 by slab, compact the three classes in slot order
    {
        KScope ks(e, "slab_classify", 8);
        k_classify<<<fj_blocks(n0, TPB), TPB, 0, st_>>>(e->lv[1].P0, n0, S.x_lo, S.x_hi, S.del_by_caller, e->oidx, S.flag[0],
                                                        S.flag[1], S.flag[2]);
        for (int c = 0; c < 3; ++c)
        {
            prim_exclusive_scan(st_, S.flag[c], S.scan[c], unsigned(n0), S.scan_tmp);
            k_compact<<<fj_blocks(n0, TPB), TPB, 0, st_>>>(S.flag[c], S.scan[c], n0, S.list[c]);
        }
    }
    unsigned h_cnt[3];
    for (int c = 0; c < 3; ++c)
        FJ_CUDA(cudaMemcpyAsync(&h_cnt[c], S.scan[c] + n0, sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    FJ_CUDA(cudaStreamSynchronize(st_));
    S.del_by_caller = nullptr; /* consumed */
    st = fj_inlet_tables_remap(e, S.flag[0], S.scan[0]); /* inlet tables hold caller indices: they follow the stayers */
    if (st)
        return st;
    const int64_t n_stay = h_cnt[0];
    int64_t mig_send[2] = {has_lo ? int64_t(h_cnt[1]) : 0, has_hi ? int64_t(h_cnt[2]) : 0}, mig_recv[2];
    if ((!has_lo && h_cnt[1]) || (!has_hi && h_cnt[2]))
    {
        fj_set_error("slab %d: %u / %u particles left the global domain through a face with no neighbour rank", S.rank,
                     h_cnt[1], h_cnt[2]);
        return FJSPH_ERR_STATE;
    }
    st = exchange_counts(e, mig_send, mig_recv);
    if (st)
        return st;
    const int64_t n_new = n_stay + mig_recv[0] + mig_recv[1];
    if (n_new > e->cap)
    {
        fj_set_error("slab %d: %lld owned particles after migration exceed the capacity %lld", S.rank, (long long)n_new,
                     (long long)e->cap);
        return FJSPH_ERR_CAPACITY;
    }

    // ---- 2. migrate both time levels; stayers are compacted in slot order, arrivals appended (lo, then hi)
    const size_t mb = std::max(std::max(level_bytes(size_t(mig_send[0])), level_bytes(size_t(mig_send[1]))),
                               std::max(level_bytes(size_t(mig_recv[0])), level_bytes(size_t(mig_recv[1]))));
    st = ensure_buffers(e, mb);
    if (st)
        return st;
    for (int l = 0; l < 2; ++l)
    {
        {
            KScope ks(e, "slab_migrate", 5);
            for (int s = 0; s < 2; ++s)
                if (mig_send[s] > 0)
                    k_pack_level<<<fj_blocks(mig_send[s], TPB), TPB, 0, st_>>>(e->lv[l], e->blk, S.list[1 + s],
                                                                               int(mig_send[s]), S.sbuf[s]);
        }
        st = comm(e, FJSPH_COMM_SENDRECV_DEV, S.sbuf[0], int64_t(level_bytes(size_t(mig_send[0]))), S.sbuf[1],
                  int64_t(level_bytes(size_t(mig_send[1]))), S.rbuf[0], int64_t(level_bytes(size_t(mig_recv[0]))), S.rbuf[1],
                  int64_t(level_bytes(size_t(mig_recv[1]))));
        if (st)
            return st;
        S.bytes_sent += (long long)(level_bytes(size_t(mig_send[0])) + level_bytes(size_t(mig_send[1])));
        {
            KScope ks(e, "slab_migrate", 4);
            if (n_stay > 0)
                k_permute_level<<<fj_blocks(n_stay, PRIM_TPB), PRIM_TPB, 0, st_>>>(e->lv[l], e->lv[2], S.list[0], int(n_stay));
            if (l == 1 && n_stay > 0)
                k_gather_int<<<fj_blocks(n_stay, TPB), TPB, 0, st_>>>(e->blk, e->blk_tmp, S.list[0], int(n_stay));
            int first = int(n_stay);
            for (int s = 0; s < 2; ++s)
            {
                if (mig_recv[s] > 0)
                    k_unpack_level<<<fj_blocks(mig_recv[s], TPB), TPB, 0, st_>>>(e->lv[2], l == 1 ? e->blk_tmp : nullptr, first,
                                                                                 int(mig_recv[s]), S.rbuf[s]);
                first += int(mig_recv[s]);
            }
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaStreamSynchronize(st_)); /* buffers are reused for the next level */
        std::swap(e->lv[l], e->lv[2]);
    }
    std::swap(e->blk, e->blk_tmp);
    e->n_owned = n_new;

    // ---- 3. ghosts: owned particles within 2H + skin of a face go to that neighbour
    const int n1 = int(n_new);
    const double wg = std::sqrt(e->P.sr) + e->skin;
    {
        KScope ks(e, "slab_ghosts", 6);
        k_ghost_flags<<<fj_blocks(n1, TPB), TPB, 0, st_>>>(e->lv[1].P0, n1, S.x_lo + wg, S.x_hi - wg, S.flag[1], S.flag[2]);
        for (int c = 1; c < 3; ++c)
        {
            prim_exclusive_scan(st_, S.flag[c], S.scan[c], unsigned(n1), S.scan_tmp);
            k_compact<<<fj_blocks(n1, TPB), TPB, 0, st_>>>(S.flag[c], S.scan[c], n1, S.list[c]);
        }
    }
    for (int c = 1; c < 3; ++c)
        FJ_CUDA(cudaMemcpyAsync(&h_cnt[c], S.scan[c] + n1, sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    FJ_CUDA(cudaStreamSynchronize(st_));
    S.n_send[0] = has_lo ? int64_t(h_cnt[1]) : 0;
    S.n_send[1] = has_hi ? int64_t(h_cnt[2]) : 0;
    st = exchange_counts(e, S.n_send, S.n_recv);
    if (st)
        return st;
    const int64_t n_all = n_new + S.n_recv[0] + S.n_recv[1];
    if (n_all > e->cap)
    {
        fj_set_error("slab %d: %lld owned + %lld ghost particles exceed the capacity %lld", S.rank, (long long)n_new,
                     (long long)(S.n_recv[0] + S.n_recv[1]), (long long)e->cap);
        return FJSPH_ERR_CAPACITY;
    }
    /* the send lists are caller indices; the caller order is reset to the slot order below, so slots it is */
    for (int s = 0; s < 2; ++s)
        if (S.n_send[s] > 0)
            FJ_CUDA(cudaMemcpyAsync(S.send_idx[s], S.list[1 + s], size_t(S.n_send[s]) * sizeof(int), cudaMemcpyDeviceToDevice,
                                    st_));
    const size_t gb = std::max(std::max(level_bytes(size_t(S.n_send[0])), level_bytes(size_t(S.n_send[1]))),
                               std::max(level_bytes(size_t(S.n_recv[0])), level_bytes(size_t(S.n_recv[1]))));
    st = ensure_buffers(e, gb);
    if (st)
        return st;
    for (int l = 1; l >= 0; --l)
    {
        {
            KScope ks(e, "slab_ghosts", 2);
            for (int s = 0; s < 2; ++s)
                if (S.n_send[s] > 0)
                    k_pack_level<<<fj_blocks(S.n_send[s], TPB), TPB, 0, st_>>>(e->lv[l], e->blk, S.send_idx[s],
                                                                               int(S.n_send[s]), S.sbuf[s]);
        }
        st = comm(e, FJSPH_COMM_SENDRECV_DEV, S.sbuf[0], int64_t(level_bytes(size_t(S.n_send[0]))), S.sbuf[1],
                  int64_t(level_bytes(size_t(S.n_send[1]))), S.rbuf[0], int64_t(level_bytes(size_t(S.n_recv[0]))), S.rbuf[1],
                  int64_t(level_bytes(size_t(S.n_recv[1]))));
        if (st)
            return st;
        S.bytes_sent += (long long)(level_bytes(size_t(S.n_send[0])) + level_bytes(size_t(S.n_send[1])));
        {
            KScope ks(e, "slab_ghosts", 2);
            int first = n1;
            for (int s = 0; s < 2; ++s)
            {
                if (S.n_recv[s] > 0)
                    k_unpack_level<<<fj_blocks(S.n_recv[s], TPB), TPB, 0, st_>>>(e->lv[l], l == 1 ? e->blk : nullptr, first,
                                                                                 int(S.n_recv[s]), S.rbuf[s]);
                first += int(S.n_recv[s]);
            }
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaStreamSynchronize(st_)); /* buffers are reused for the other level */
    }
    {
        KScope ks(e, "slab_ghosts", 1);
        k_identity<<<fj_blocks(n_all, TPB), TPB, 0, st_>>>(e->oidx, e->slot_of, int(n_all));
    }
    FJ_CUDA(cudaGetLastError());
    e->n = n_all;

    // ---- 4. global particle counts (npd divides by all points, the residual by the fluid points)
    FJ_CUDA(cudaMemsetAsync(S.flag[0], 0, sizeof(unsigned), st_));
    k_count_fluid<<<std::min(fj_blocks(n1, TPB), 1024), TPB, 0, st_>>>(e->blk, e->n_bound_blocks, n1, S.flag[0]);
    FJ_CUDA(cudaMemcpyAsync(&h_cnt[0], S.flag[0], sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    FJ_CUDA(cudaStreamSynchronize(st_));
    double v[2] = {double(h_cnt[0]), double(n1)};
    st = fj_allreduce(e, FJSPH_COMM_SUM, v, 2);
    if (st)
        return st;
    S.n_fluid_global = v[0];
    S.n_total_global = v[1];
    e->bound_points = int64_t(n1) - int64_t(h_cnt[0]);
    S.redecomps++;
    e->skin_valid = false;
    e->list_valid = false;
    return FJSPH_OK;
}

extern "C" int fjsph_set_slab(FjsphEngine* e, int32_t rank, int32_t world, double x_lo, double x_hi, FjsphCommFn fn,
                              void* user)
{
    cudaSetDevice(e->device);
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn) || !(x_lo < x_hi))
    {
        fj_set_error("set_slab: bad rank / world / bounds / callback");
        return FJSPH_ERR_INVALID;
    }
    if (e->n <= 0)
    {
        fj_set_error("set_slab: upload this rank's particles first");
        return FJSPH_ERR_STATE;
    }
    Slab& S = e->slab;
    S.on = true;
    S.rank = rank;
    S.world = world;
    S.x_lo = x_lo;
    S.x_hi = x_hi;
    S.fn = fn;
    S.user = user;
    S.n_send[0] = S.n_send[1] = S.n_recv[0] = S.n_recv[1] = 0;
    S.pending = false;
    S.n_interior = 0;
    if (const char* ov = getenv("FJSPH_SLAB_OVERLAP"))
        S.overlap = atoi(ov) != 0;
    if (!S.comm_stream)
    {
        FJ_CUDA(cudaStreamCreateWithFlags(&S.comm_stream, cudaStreamNonBlocking));
        FJ_CUDA(cudaEventCreateWithFlags(&S.ev_ready, cudaEventDisableTiming));
        FJ_CUDA(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
    }
    if (!S.flag[0])
    {
        const size_t cap = size_t(e->cap) + 1;
        for (int c = 0; c < 3; ++c)
        {
            FJ_CUDA(cudaMalloc(&S.flag[c], cap * sizeof(unsigned)));
            FJ_CUDA(cudaMalloc(&S.scan[c], cap * sizeof(unsigned)));
            FJ_CUDA(cudaMalloc(&S.list[c], cap * sizeof(int)));
        }
        for (int s = 0; s < 2; ++s) FJ_CUDA(cudaMalloc(&S.send_idx[s], cap * sizeof(int)));
        S.send_cap = int64_t(cap);
        FJ_CUDA(cudaMalloc(&S.scan_tmp, (cap / SCAN_TILE + 2) * sizeof(unsigned)));
    }
    e->n_owned = e->n; /* ghosts are created by the first neighbour build */
    e->skin_valid = false;
    e->list_valid = false;
    double v[2] = {double(e->n_owned - e->bound_points), double(e->n_owned)};
    int st = fj_allreduce(e, FJSPH_COMM_SUM, v, 2);
    if (st)
        return st;
    S.n_fluid_global = v[0];
    S.n_total_global = v[1];
    /* ids of particles an inlet adds later must be unique over all ranks, and the ones a single engine would hand out:
       count from the global total, or from one past the highest id any rank holds if the host supplied ids beyond it
       (fjsph_upload_state keeps next_part_id = max(n, max id + 1): the reference's counter only grows) */
    double ids = double(e->next_part_id);
    st = fj_allreduce(e, FJSPH_COMM_MAX, &ids, 1);
    if (st)
        return st;
    e->next_part_id = std::max((long long)ids, (long long)S.n_total_global);
    return FJSPH_OK;
}

extern "C" int fjsph_get_stream(FjsphEngine* e, void** stream)
{
    *stream = (void*)e->stream;
    return FJSPH_OK;
}
extern "C" int fjsph_slab_device_reductions(FjsphEngine* e, int32_t on)
{
    e->slab.dev_reduce = on != 0;
    return FJSPH_OK;
}

/* the stream device-buffer exchanges with op FJSPH_COMM_SENDRECV_DEV_ASYNC must be ordered on */
extern "C" int fjsph_slab_comm_stream(FjsphEngine* e, void** stream)
{
    if (!e->slab.on || !e->slab.comm_stream)
    {
        fj_set_error("slab_comm_stream: call fjsph_set_slab first");
        return FJSPH_ERR_STATE;
    }
    *stream = (void*)e->slab.comm_stream;
    return FJSPH_OK;
}

/* forward exchanges that ran beside an interior sweep since fjsph_set_slab */
extern "C" int fjsph_slab_overlapped(FjsphEngine* e, int64_t* n)
{
    *n = e->slab.overlapped;
    return FJSPH_OK;
}

extern "C" int fjsph_slab_stats(FjsphEngine* e, int64_t* n_owned, int64_t* n_ghost, int64_t* exchanges, int64_t* redecomps,
                                int64_t* bytes_sent)
{
    if (n_owned)
        *n_owned = e->n_owned;
    if (n_ghost)
        *n_ghost = e->n - e->n_owned;
    if (exchanges)
        *exchanges = e->slab.exchanges;
    if (redecomps)
        *redecomps = e->slab.redecomps;
    if (bytes_sent)
        *bytes_sent = e->slab.bytes_sent;
    return FJSPH_OK;
}
