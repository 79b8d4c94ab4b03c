// prims.cuh — small device primitives shared by the neighbour build and the slab exchange: a block-tiled
// exclusive prefix scan (counting sort of Neighbours.cpp's replacement, compaction of migrating / ghost
// particles) and the gather of one whole time level through an index list.
#pragma once
#include "engine.cuh"

namespace
{
constexpr int PRIM_TPB = 256;
#define TPB_PRIM_GUARD
// ---------------------------------------------------------------- block prefix scan (exclusive)
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = PRIM_TPB * SCAN_ITEMS;

__global__ void k_scan_tiles(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned n,
                             unsigned* __restrict__ tile_sum)
{
    __shared__ unsigned warp_tot[PRIM_TPB / 32];
    const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = s;
    for (int o = 1; o < 32; o <<= 1)
    {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= unsigned(o))
            inc += t;
    }
    if (lane == 31)
        warp_tot[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        unsigned t = (lane < PRIM_TPB / 32) ? warp_tot[lane] : 0u;
        unsigned ti = t;
        for (int o = 1; o < 32; o <<= 1)
        {
            unsigned u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= unsigned(o))
                ti += u;
        }
        if (lane < PRIM_TPB / 32)
            warp_tot[lane] = ti - t; // exclusive warp offsets
        if (lane == PRIM_TPB / 32 - 1)
            tile_sum[blockIdx.x] = ti;
    }
    __syncthreads();
    unsigned run = warp_tot[w] + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        if (base + k < n)
            out[base + k] = run;
        run += v[k];
    }
}

// single block: exclusive scan of tile sums in place, total written to tile_sum[ntiles]
__global__ void k_scan_tile_sums(unsigned* __restrict__ tile_sum, unsigned ntiles)
{
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (unsigned start = 0; start < ntiles; start += blockDim.x)
    {
        const unsigned idx = start + threadIdx.x;
        const unsigned v = (idx < ntiles) ? tile_sum[idx] : 0u;
        unsigned inc = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= unsigned(o))
                inc += t;
        }
        if (lane == 31)
            warp_tot[w] = inc;
        __syncthreads();
        if (w == 0)
        {
            unsigned t = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0u;
            unsigned ti = t;
            for (int o = 1; o < 32; o <<= 1)
            {
                unsigned u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= unsigned(o))
                    ti += u;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const unsigned carry = carry_s;
        if (idx < ntiles)
            tile_sum[idx] = carry + warp_tot[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1)
            carry_s = carry + warp_tot[w] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        tile_sum[ntiles] = carry_s;
}

__global__ void k_scan_add(unsigned* __restrict__ out, unsigned n, const unsigned* __restrict__ tile_sum,
                           unsigned ntiles)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] += tile_sum[i / SCAN_TILE];
    if (i == 0)
        out[n] = tile_sum[ntiles];
}

// ---------------------------------------------------------------- permute one level
__global__ void k_permute_level(Level in, Level out, const int* __restrict__ perm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int s = perm[i];
#define X(T, f) out.f[i] = in.f[s];
    FJ_LEVEL_FIELDS(X)
#undef X
}


// stream compaction: list[scan[i]] = i for flagged i (scan = exclusive scan of flag)
__global__ void k_compact(const unsigned* __restrict__ flag, const unsigned* __restrict__ scan, int n,
                          int* __restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i])
        list[scan[i]] = i;
}


// exclusive scan of in[0..n) into out[0..n], out[n] = total; tmp holds n/SCAN_TILE + 2 words
inline void prim_exclusive_scan(cudaStream_t st, const unsigned* in, unsigned* out, unsigned n, unsigned* tmp)
{
    const unsigned ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<ntiles, PRIM_TPB, 0, st>>>(in, out, n, tmp);
    k_scan_tile_sums<<<1, 1024, 0, st>>>(tmp, ntiles);
    k_scan_add<<<(n + PRIM_TPB - 1) / PRIM_TPB, PRIM_TPB, 0, st>>>(out, n, tmp, ntiles);
}
} // namespace
