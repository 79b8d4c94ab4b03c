// comm_nccl.cu — FjsphCommFn on NCCL, natively (include/fjsph_b200_nccl.h).
//
// The engine packs and unpacks ghosts and migrating particles on the device (halo.cu) and asks its host for three
// things: neighbour send/recv of device buffers (on the main stream, or on the comm stream beside the interior sweeps),
// neighbour send/recv of a few host words (counts), and all-reduces of a handful of doubles.  Here every one of them is
// an NCCL call ordered on the stream the engine names; host words travel through a pinned bounce buffer on that stream.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/fjsph_b200_nccl.h"

namespace
{
thread_local char g_err[512] = "";
void set_err(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
#define NCCL_OK(call)                                                              \
    do                                                                             \
    {                                                                              \
        ncclResult_t r_ = (call);                                                  \
        if (r_ != ncclSuccess)                                                     \
        {                                                                          \
            set_err("%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)
#define CUDA_OK(call)                                                              \
    do                                                                             \
    {                                                                              \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess)                                                     \
        {                                                                          \
            set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)
constexpr size_t BOUNCE_BYTES = 1 << 16;
} // namespace

struct FjsphNcclComm
{
    /* two communicators, as many ordering domains: NCCL serialises the operations of ONE communicator in the order the host
       enqueues them, whatever streams they are on, so neighbour exchanges queued on the comm stream beside the sweeps and
       all-reduces queued on the main stream must not share one -- a rank that (legitimately) enqueues the two kinds in the
       other order than its neighbour would deadlock, and an exchange would wait for an all-reduce that waits for a sweep */
    ncclComm_t comm = nullptr; /* collectives */
    ncclComm_t p2p = nullptr;  /* ncclSend / ncclRecv between x-neighbours */
    int rank = 0, world = 1, device = 0;
    cudaStream_t main_stream = nullptr, comm_stream = nullptr, own_stream = nullptr;
    char* d_bounce = nullptr; /* device scratch for host-array ops: [send lo | send hi | recv lo | recv hi] quarters */
    char* h_bounce = nullptr; /* pinned mirror */
    long long calls[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace
{
// send a -> rank-1, b -> rank+1; receive c <- rank-1, d <- rank+1 (device pointers), one NCCL group on `st`
int sendrecv(FjsphNcclComm* c, cudaStream_t st, const void* a, int64_t na, const void* b, int64_t nb, void* rc, int64_t nc,
             void* rd, int64_t nd)
{
    NCCL_OK(ncclGroupStart());
    if (rc && nc > 0)
        NCCL_OK(ncclRecv(rc, size_t(nc), ncclChar, c->rank - 1, c->p2p, st));
    if (rd && nd > 0)
        NCCL_OK(ncclRecv(rd, size_t(nd), ncclChar, c->rank + 1, c->p2p, st));
    if (a && na > 0)
        NCCL_OK(ncclSend(a, size_t(na), ncclChar, c->rank - 1, c->p2p, st));
    if (b && nb > 0)
        NCCL_OK(ncclSend(b, size_t(nb), ncclChar, c->rank + 1, c->p2p, st));
    NCCL_OK(ncclGroupEnd());
    return 0;
}

int callback(void* user, int32_t op, void* a, int64_t na, void* b, int64_t nb, void* rc, int64_t nc, void* rd, int64_t nd)
{
    FjsphNcclComm* c = static_cast<FjsphNcclComm*>(user);
    if (op >= 0 && op < 8)
        c->calls[op]++;
    cudaStream_t st = c->main_stream ? c->main_stream : c->own_stream;
    switch (op)
    {
    case FJSPH_COMM_SUM_DEV:
    case FJSPH_COMM_MAX_DEV:
        NCCL_OK(ncclAllReduce(a, a, size_t(na / 8), ncclDouble, op == FJSPH_COMM_SUM_DEV ? ncclSum : ncclMax, c->comm, st));
        return 0;
    case FJSPH_COMM_SUM:
    case FJSPH_COMM_MAX:
    {
        if (size_t(na) > BOUNCE_BYTES)
        {
            set_err("host all-reduce of %lld bytes exceeds the bounce buffer", (long long)na);
            return 1;
        }
        std::memcpy(c->h_bounce, a, size_t(na));
        CUDA_OK(cudaMemcpyAsync(c->d_bounce, c->h_bounce, size_t(na), cudaMemcpyHostToDevice, st));
        NCCL_OK(ncclAllReduce(c->d_bounce, c->d_bounce, size_t(na / 8), ncclDouble, op == FJSPH_COMM_SUM ? ncclSum : ncclMax,
                              c->comm, st));
        CUDA_OK(cudaMemcpyAsync(c->h_bounce, c->d_bounce, size_t(na), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        std::memcpy(a, c->h_bounce, size_t(na));
        return 0;
    }
    case FJSPH_COMM_SENDRECV_DEV:
        return sendrecv(c, st, a, na, b, nb, rc, nc, rd, nd);
    case FJSPH_COMM_SENDRECV_DEV_ASYNC:
        if (!c->comm_stream)
        {
            set_err("ASYNC exchange requested but the comm stream is unknown (fjsph_nccl_attach)");
            return 1;
        }
        return sendrecv(c, c->comm_stream, a, na, b, nb, rc, nc, rd, nd); /* never blocks the host */
    case FJSPH_COMM_SENDRECV_HOST:
    {
        const size_t q = BOUNCE_BYTES / 4;
        if (size_t(na) > q || size_t(nb) > q || size_t(nc) > q || size_t(nd) > q)
        {
            set_err("host exchange exceeds the bounce buffer");
            return 1;
        }
        if (a && na > 0)
            std::memcpy(c->h_bounce, a, size_t(na));
        if (b && nb > 0)
            std::memcpy(c->h_bounce + q, b, size_t(nb));
        CUDA_OK(cudaMemcpyAsync(c->d_bounce, c->h_bounce, 2 * q, cudaMemcpyHostToDevice, st));
        if (sendrecv(c, st, (a && na > 0) ? c->d_bounce : nullptr, na, (b && nb > 0) ? c->d_bounce + q : nullptr, nb,
                     (rc && nc > 0) ? c->d_bounce + 2 * q : nullptr, nc, (rd && nd > 0) ? c->d_bounce + 3 * q : nullptr, nd))
            return 1;
        CUDA_OK(cudaMemcpyAsync(c->h_bounce + 2 * q, c->d_bounce + 2 * q, 2 * q, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (rc && nc > 0)
            std::memcpy(rc, c->h_bounce + 2 * q, size_t(nc));
        if (rd && nd > 0)
            std::memcpy(rd, c->h_bounce + 3 * q, size_t(nd));
        return 0;
    }
    default:
        set_err("unknown comm op %d", op);
        return 1;
    }
}
} // namespace

extern "C" {

const char* fjsph_nccl_last_error(void) { return g_err; }

int fjsph_nccl_unique_id(char id[FJSPH_NCCL_ID_BYTES])
{
    static_assert(2 * sizeof(ncclUniqueId) <= FJSPH_NCCL_ID_BYTES, "two ncclUniqueIds do not fit FJSPH_NCCL_ID_BYTES");
    ncclUniqueId u[2]; /* one per communicator */
    NCCL_OK(ncclGetUniqueId(&u[0]));
    NCCL_OK(ncclGetUniqueId(&u[1]));
    std::memset(id, 0, FJSPH_NCCL_ID_BYTES);
    std::memcpy(id, u, sizeof(u));
    return 0;
}

int fjsph_nccl_create(const char id[FJSPH_NCCL_ID_BYTES], int32_t rank, int32_t world, int32_t device, FjsphNcclComm** out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world)
    {
        set_err("fjsph_nccl_create: bad arguments");
        return 1;
    }
    CUDA_OK(cudaSetDevice(device));
    FjsphNcclComm* c = new FjsphNcclComm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclUniqueId u[2];
    std::memcpy(u, id, sizeof(u));
    NCCL_OK(ncclCommInitRank(&c->comm, world, u[0], rank));
    NCCL_OK(ncclCommInitRank(&c->p2p, world, u[1], rank));
    CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&c->d_bounce, BOUNCE_BYTES));
    CUDA_OK(cudaMallocHost(&c->h_bounce, BOUNCE_BYTES));
    *out = c;
    return 0;
}

int fjsph_nccl_attach(FjsphNcclComm* c, FjsphEngine* e, double x_lo, double x_hi)
{
    if (!c || !e)
    {
        set_err("fjsph_nccl_attach: bad arguments");
        return 1;
    }
    void* st = nullptr;
    if (fjsph_get_stream(e, &st))
        return 1;
    c->main_stream = static_cast<cudaStream_t>(st);
    if (fjsph_slab_device_reductions(e, 1))
        return 1;
    /* fjsph_set_slab already all-reduces the particle counts through the callback: the main stream is known by now; the
       comm stream exists once it returns */
    if (fjsph_set_slab(e, c->rank, c->world, x_lo, x_hi, callback, c))
    {
        set_err("fjsph_set_slab: %s", fjsph_last_error());
        return 1;
    }
    if (c->world > 1)
    {
        if (fjsph_slab_comm_stream(e, &st))
            return 1;
        c->comm_stream = static_cast<cudaStream_t>(st);
    }
    return 0;
}

int fjsph_nccl_allreduce_host(FjsphNcclComm* c, double* v, int64_t n, int32_t op)
{
    return callback(c, op == FJSPH_COMM_MAX ? FJSPH_COMM_MAX : FJSPH_COMM_SUM, v, n * 8, nullptr, 0, nullptr, 0, nullptr, 0);
}

int fjsph_nccl_barrier(FjsphNcclComm* c)
{
    double z = 0.0;
    return fjsph_nccl_allreduce_host(c, &z, 1, FJSPH_COMM_SUM);
}

int64_t fjsph_nccl_calls(FjsphNcclComm* c, int32_t op) { return (c && op >= 0 && op < 8) ? c->calls[op] : 0; }

int fjsph_nccl_destroy(FjsphNcclComm* c)
{
    if (!c)
        return 0;
    cudaSetDevice(c->device);
    if (c->own_stream)
    {
        cudaStreamSynchronize(c->own_stream);
        cudaStreamDestroy(c->own_stream);
    }
    if (c->d_bounce)
        cudaFree(c->d_bounce);
    if (c->h_bounce)
        cudaFreeHost(c->h_bounce);
    if (c->p2p)
        ncclCommDestroy(c->p2p);
    if (c->comm)
        ncclCommDestroy(c->comm);
    delete c;
    return 0;
}

} // extern "C"
