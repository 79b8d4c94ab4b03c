// neighbours.cu — uniform cell list + row-run neighbour lists replacing FJSPH's nanoflann KD-tree radius search.
//
// Replaces update_neighbours / find_neighbours / radius_search (reference src/Neighbours.cpp:7-47):
//   list_i = { j : ((xi-xj)^2 + (yi-yj)^2) + (zi-zj)^2 < sr }, strict '<', evaluated WITHOUT fma
// contraction in the order nanoflann's metric_L2_Simple accumulates it, so the sets are bit-exact with
// the CPU oracle.  Self is not stored (callers add the self terms explicitly; outlist[i].size() is
// count+1).
//
// Layout: cells are (2H + skin) x dx x dx bricks along the ROW axis u (the longest extent of the
// bounding box); a ROW is the pencil of cells with the same transverse coordinates, its particles contiguous in memory
// and sorted along u.  The neighbours of particle i inside one neighbouring row are then a WINDOW of consecutive
// indices, so a list is one run {first, 32-bit mask} per neighbouring row instead of one entry per neighbour: ~0.5 KB
// per particle instead of 3.4 KB, and the 32 lanes of a warp (32 consecutive particles of a row) gather 32 consecutive
// records at every step of a sweep.
//
// Pipeline per build (all on e->stream):
//   bounds -> cell key + warp-aggregated histogram -> block prefix scan -> scatter -> per-cell ordering along u
//   (ties by caller index: deterministic) -> permute both time levels into cell order -> work-warp numbering of the rows
//   -> SKIN runs (every j with d < 2H + skin; rebuilt only when a particle has moved more than skin / 2)
//   -> EXACT runs (the reference's OUTL, filtered from the skin runs at every update_neighbours).
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "engine.cuh"
#include "prims.cuh"

namespace
{

constexpr int TPB = 256;

// ---------------------------------------------------------------- bounds (+ displacement since the skin build)
// partial[b*NRED + c]: c = 0-2 min x,y,z | 3-5 max x,y,z | 6 max |x - xref|^2 | 7-9 min (x - xref) | 10-12 max (x - xref)
// Components 7-12 bound the RELATIVE displacement of any two particles: a uniform drift of the whole fluid (a jet)
// moves every particle far but no pair apart, and must not trigger a rebuild of the superset list.
constexpr int NRED = 13;
__device__ __forceinline__ bool red_is_min(int c) { return c < 3 || (c >= 7 && c < 10); }
__device__ __forceinline__ double red_combine(int c, double v, double u)
{
    /* min for the lower bounds; NaN-propagating max (u <= v ? v : u) for the rest: a NaN reads as "moved too far" */
    return red_is_min(c) ? fmin(v, u) : ((u <= v) ? v : u);
}
__device__ __forceinline__ double red_identity(int c) { return red_is_min(c) ? 1e300 : (c == 6 ? 0.0 : -1e300); }

__global__ void k_bounds(const double4* __restrict__ P0, const double4* __restrict__ xref, int n,
                         double* __restrict__ partial)
{
    double acc[NRED];
#pragma unroll
    for (int c = 0; c < NRED; ++c) acc[c] = red_identity(c);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const double4 a = P0[i];
        double v[NRED] = {a.x, a.y, a.z, a.x, a.y, a.z, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (xref)
        {
            const double4 r = xref[i];
            const double dx = a.x - r.x, dy = a.y - r.y, dz = a.z - r.z;
            v[6] = dx * dx + dy * dy + dz * dz;
            v[7] = v[10] = dx;
            v[8] = v[11] = dy;
            v[9] = v[12] = dz;
        }
#pragma unroll
        for (int c = 0; c < NRED; ++c) acc[c] = red_combine(c, acc[c], v[c]);
    }
    __shared__ double sm[NRED][TPB / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < NRED; ++c)
    {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v = red_combine(c, v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0)
            sm[c][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < NRED)
    {
        double v = sm[threadIdx.x][0];
        for (int k = 1; k < TPB / 32; ++k) v = red_combine(threadIdx.x, v, sm[threadIdx.x][k]);
        partial[blockIdx.x * NRED + threadIdx.x] = v;
    }
}

__global__ void k_bounds_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out)
{
    // one warp per component, lanes stride the block partials
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= NRED)
        return;
    double v = red_identity(c);
    for (int b = lane; b < nblocks; b += 32) v = red_combine(c, v, partial[b * NRED + c]);
    for (int o = 16; o > 0; o >>= 1) v = red_combine(c, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0)
        out[c] = v;
}

// ---------------------------------------------------------------- keys + histogram
__device__ __forceinline__ double comp(const double4& a, int ax) { return ax == 0 ? a.x : (ax == 1 ? a.y : a.z); }
__device__ __forceinline__ void cell_of(const Grid& g, const double4& a, int& cx, int& cy, int& cz)
{
    cx = min(max(int(floor((comp(a, g.ax0) - g.ox) * g.inv_cell)), 0), g.nx - 1);
    cy = min(max(int(floor((comp(a, g.ax1) - g.oy) * g.inv_cy)), 0), g.ny - 1);
    cz = min(max(int(floor((comp(a, g.ax2) - g.oz) * g.inv_cz)), 0), g.nz - 1);
}

// Slab mode sorts three classes one behind the other (key offsets 0, n_keys, 2 n_keys): INTERIOR owned particles
// (x in [x_edge_lo, x_edge_hi): farther than 2H + skin from every face with a neighbour rank, so no ghost can be
// their neighbour), EDGE owned particles (the ones sent as ghosts, same predicate as k_ghost_flags) and GHOSTS.
__global__ void k_key_hist(const double4* __restrict__ P0, int n, int n_owned, Grid g, const unsigned* __restrict__ my,
                           const unsigned* __restrict__ mz, unsigned* __restrict__ key, unsigned* __restrict__ rank,
                           unsigned* __restrict__ count, int n_class, double x_edge_lo, double x_edge_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    unsigned k = 0xFFFFFFFFu;
    if (i < n)
    {
        double4 a = P0[i];
        int cx, cy, cz;
        cell_of(g, a, cx, cy, cz);
        k = unsigned(cx) | my[cy] | mz[cz];
        if (n_class == 3)
        {
            if (i >= n_owned)
                k += 2u * g.n_keys; /* ghosts sort behind the owned particles: slots [0, n_owned) stay the owned ones */
            else if (a.x < x_edge_lo || !(a.x < x_edge_hi))
                k += g.n_keys;
        }
        key[i] = k;
    }
    // warp-aggregated atomics: one atomicAdd per distinct key in the warp
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (i < n && int(lane) == leader)
        base = atomicAdd(&count[k], unsigned(__popc(peers)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n)
        rank[i] = base + __popc(peers & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------- scatter + deterministic cell order
__global__ void k_scatter(const unsigned* __restrict__ key, const unsigned* __restrict__ rank,
                          const unsigned* __restrict__ cell_start, int n, int* __restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        perm[cell_start[key[i]] + rank[i]] = i;
}

// One warp per occupied cell: order its members along the row axis u, ties by caller index (rank sort).  The atomics
// above leave an arbitrary order inside a cell; this makes the layout -- and with it every FP64 summation order
// downstream -- reproducible run to run, and the rows sorted along u, which is what makes a particle's neighbours
// inside a row a window of consecutive indices.
__global__ void k_cell_order(const unsigned* __restrict__ cell_start, unsigned n_keys, const int* __restrict__ perm,
                             const int* __restrict__ oidx, const double4* __restrict__ P0, int ax0,
                             int* __restrict__ perm2)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (warp >= n_keys)
        return;
    const unsigned s = cell_start[warp], e = cell_start[warp + 1];
    const unsigned c = e - s;
    if (c == 0)
        return;
    for (unsigned t = lane; t < c; t += 32)
    {
        const int mine = perm[s + t];
        const int mo = oidx[mine];
        const double mu = comp(P0[mine], ax0);
        unsigned r = 0;
        for (unsigned u = 0; u < c; ++u)
        {
            const int other = perm[s + u];
            const double ou = comp(P0[other], ax0);
            r += (ou < mu || (ou == mu && oidx[other] < mo)) ? 1u : 0u;
        }
        perm2[s + r] = mine;
    }
}

__global__ void k_permute_index(const int* __restrict__ oidx_in, const int* __restrict__ blk_in,
                                const int* __restrict__ perm, int n, int* __restrict__ oidx_out,
                                int* __restrict__ blk_out, int* __restrict__ slot_of)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int s = perm[i];
    const int o = oidx_in[s];
    oidx_out[i] = o;
    blk_out[i] = blk_in[s];
    slot_of[o] = i;
}

// ---------------------------------------------------------------- work warps of the rows
// row r spans cell keys [r << bx, (r + 1) << bx); owned rows (the first n_owned_rows of the table) are cut into
// 32-particle chunks, one work warp each.  stats[0] = longest owned row.
__global__ void k_row_warps(const unsigned* __restrict__ cell_start, int bx, unsigned n_rows, unsigned n_owned_rows,
                            unsigned* __restrict__ row_warps, unsigned* __restrict__ stats)
{
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows)
        return;
    const unsigned len = cell_start[size_t(r + 1u) << bx] - cell_start[size_t(r) << bx];
    const bool owned = r < n_owned_rows;
    row_warps[r] = owned ? (len + 31u) >> 5 : 0u;
    if (owned && len > 0)
        atomicMax(stats, len);
}

__global__ void k_snapshot(const double4* __restrict__ P0, double4* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = P0[i];
}

__device__ __forceinline__ double4 ldg256(const double4* __restrict__ base, unsigned j)
{
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(base + j));
    return v;
}

// ---------------------------------------------------------------- skin runs
// For every neighbouring row inside the disc of radius 2H + skin (row_off, ascending w then v) and every class
// segment of it, the lanes scan the u-cells cx-1 .. cx+1 of that row -- one contiguous index range -- and keep the
// window [first, last] of the particles closer than 2H + skin.  The 32 lanes are consecutive particles of one row, so
// their ranges overlap and the scan reads a few distinct records per step.  Windows longer than 32 take several slots.
__global__ void __launch_bounds__(FJ_ROW_WARPS * 32)
    k_build_skin_runs(const double4* __restrict__ P0, RowMap M, Grid g, const unsigned* __restrict__ my,
                      const unsigned* __restrict__ mz, const int2* __restrict__ row_off, int n_off, int n_seg,
                      double sr_skin, int scap, unsigned* __restrict__ srun, int* __restrict__ srows,
                      int* __restrict__ flag)
{
    int i, W;
    bool work;
    const bool valid = fj_row_thread(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    const unsigned lane = threadIdx.x & 31u;
    const double4 a = P0[i];
    int cx, cy, cz;
    cell_of(g, a, cx, cy, cz);
    cy = __shfl_sync(0xffffffffu, cy, 0); /* the warp's row */
    cz = __shfl_sync(0xffffffffu, cz, 0);
    const unsigned cxl = unsigned(max(cx - 1, 0)), cxh = unsigned(min(cx + 1, g.nx - 1));
    unsigned* __restrict__ dst = srun + (size_t(W) * size_t(scap)) * 32u + lane;
    int kk = 0;
    for (int k = 0; k < n_off; ++k)
    {
        const int2 off = row_off[k];
        const int y = cy + off.x, z = cz + off.y;
        if (y < 0 || y >= g.ny || z < 0 || z >= g.nz)
            continue;
        const unsigned rowkey = my[y] | mz[z];
        for (int seg = 0; seg < n_seg; ++seg)
        {
            const unsigned base = rowkey + unsigned(seg) * g.n_keys;
            const unsigned s = M.cell_start[base | cxl];
            const unsigned e = valid ? M.cell_start[(base | cxh) + 1u] : s;
            const int T = __reduce_max_sync(0xffffffffu, int(e - s));
            unsigned first = 0xFFFFFFFFu, last = 0u;
            for (int o = 0; o < T; ++o)
            {
                const unsigned j = s + unsigned(o);
                if (j < e)
                {
                    const double4 q = ldg256(P0, j);
                    const double ddx = a.x - q.x, ddy = a.y - q.y, ddz = a.z - q.z;
                    const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                    if (d2 < sr_skin && int(j) != i)
                    {
                        first = min(first, j);
                        last = max(last, j);
                    }
                }
            }
            const int len = (first != 0xFFFFFFFFu) ? int(last - first) + 1 : 0;
            const int maxlen = __reduce_max_sync(0xffffffffu, len);
            for (int p = 0; p < maxlen; p += 32)
            {
                const int pl = len - p;
                if (kk < scap)
                    dst[size_t(kk) * 32u] = (pl > 0) ? ((first + unsigned(p)) | (unsigned(min(pl, 32) - 1) << 27)) : FJ_RUN_EMPTY;
                kk++;
            }
        }
    }
    if (lane == 0)
    {
        srows[W] = min(kk, scap);
        if (kk > scap)
            atomicMax(flag, kk);
        atomicMax(flag + 1, kk); /* the most slots any warp uses: what the exact runs (a subset, slot for slot) can need */
    }
}

// ---------------------------------------------------------------- exact runs
// exact list from the skin runs: list_i = { j in skin_i : ((xi-xj)^2 + (yi-yj)^2) + (zi-zj)^2 < sr }, as run + mask
__global__ void __launch_bounds__(FJ_ROW_WARPS * 32)
    k_exact_runs(const double4* __restrict__ P0, RowMap M, const unsigned* __restrict__ srun,
                 const int* __restrict__ srows, int scap, double sr, int ecap, uint2* __restrict__ erun,
                 int* __restrict__ erows, int* __restrict__ ncount, unsigned long long* __restrict__ stats)
{
    int i, W;
    bool work;
    const bool valid = fj_row_thread(M, i, W, work);
    if (!work)
        return; /* warp-uniform */
    const unsigned lane = threadIdx.x & 31u;
    const double4 a = P0[i];
    const int nrow = srows[W];
    const unsigned* __restrict__ sp = srun + (size_t(W) * size_t(scap)) * 32u + lane;
    uint2* __restrict__ dst = erun + (size_t(W) * size_t(ecap)) * 32u + lane;
    int kk = 0, cnt = 0, steps = 0;
    auto test = [&](const unsigned j, const double4 q, const bool ok) -> bool {
        // nanoflann metric_L2_Simple order, no fma contraction (bit-exact sets)
        const double ddx = __dsub_rn(a.x, q.x), ddy = __dsub_rn(a.y, q.y), ddz = __dsub_rn(a.z, q.z);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
        return ok && d2 < sr && int(j) != i;
    };
    unsigned d = (nrow > 0) ? sp[0] : FJ_RUN_EMPTY;
    for (int k = 0; k < nrow; ++k)
    {
        const unsigned dn = (k + 1 < nrow) ? sp[size_t(k + 1) * 32u] : FJ_RUN_EMPTY;
        if (!valid)
            d = FJ_RUN_EMPTY;
        const unsigned start = d & FJ_IDX_MASK;
        const int len = (d == FJ_RUN_EMPTY) ? 0 : int(d >> 27) + 1;
        const int T = __reduce_max_sync(0xffffffffu, len);
        unsigned mask = 0u;
        for (int o = 0; o < T; o += 2)
        {
            /* two independent gathers in flight; lanes past their window re-read their own record */
            const bool ok0 = o < len, ok1 = o + 1 < len;
            const unsigned j0 = ok0 ? start + unsigned(o) : unsigned(i), j1 = ok1 ? start + unsigned(o) + 1u : unsigned(i);
            const double4 q0 = ldg256(P0, j0);
            const double4 q1 = ldg256(P0, j1);
            if (test(j0, q0, ok0))
                mask |= 1u << o;
            if (test(j1, q1, ok1))
                mask |= 2u << o;
        }
        if (__ballot_sync(0xffffffffu, mask != 0u))
        {
            const int tz = mask ? __ffs(int(mask)) - 1 : 0;
            dst[size_t(kk) * 32u] = mask ? make_uint2(start + unsigned(tz), mask >> tz) : make_uint2(0u, 0u);
            kk++;
            if (stats) /* lockstep steps a sweep spends on this slot: the longest trimmed window */
                steps += __reduce_max_sync(0xffffffffu, mask ? 32 - __clz(int(mask >> tz)) : 0);
        }
        cnt += __popc(mask);
        d = dn;
    }
    if (lane == 0)
        erows[W] = kk;
    if (valid)
        ncount[i] = cnt;
    if (stats)
    {
        /* [0] lane-steps a sweep walks (32 x steps per warp), [1] pairs, [2] slots, [3] work warps */
        const int pairs = __reduce_add_sync(0xffffffffu, valid ? cnt : 0);
        if (lane == 0)
        {
            atomicAdd(&stats[0], 32ull * unsigned(steps));
            atomicAdd(&stats[1], (unsigned long long)pairs);
            atomicAdd(&stats[2], (unsigned long long)kk);
            atomicAdd(&stats[3], 1ull);
        }
    }
}

// debug / parity view: the runs of every owned particle written out as caller indices (fjsph_get_neighbours)
__global__ void k_runs_to_csr(RowMap M, const uint2* __restrict__ erun, const int* __restrict__ erows, int ecap,
                              const int* __restrict__ oidx, const long long* __restrict__ offsets,
                              long long* __restrict__ idx, int* __restrict__ bad)
{
    int i, W;
    bool work;
    const bool valid = fj_row_thread(M, i, W, work);
    if (!work || !valid)
        return;
    const unsigned lane = threadIdx.x & 31u;
    const uint2* __restrict__ dp = erun + (size_t(W) * size_t(ecap)) * 32u + lane;
    const int nrow = erows[W];
    const int c = oidx[i];
    long long* dst = idx + offsets[c];
    const long long room = offsets[c + 1] - offsets[c];
    long long t = 0;
    for (int k = 0; k < nrow; ++k)
    {
        const uint2 d = dp[size_t(k) * 32u];
        for (unsigned m = d.y, o = 0; m; m >>= 1, ++o)
            if (m & 1u)
            {
                if (t < room)
                    dst[t] = oidx[d.x + o];
                t++;
            }
    }
    if (t < room)
        dst[t] = c;
    t++;
    if (t != room)
        atomicExch(bad, 1);
}

int bits_for(int n)
{
    int b = 0;
    while ((1 << b) < n) b++;
    return b;
}

} // namespace

// ------------------------------------------------------------------ host orchestration
RowMap fj_row_map(const FjsphEngine* e, int first_class, int n_classes)
{
    RowMap M;
    const unsigned n_rowkeys = 1u << (e->grid.by + e->grid.bz);
    M.cell_start = e->cell_start;
    M.warp_start = e->warp_start;
    M.row0 = unsigned(first_class) * n_rowkeys;
    M.n_rows = unsigned(n_classes) * n_rowkeys;
    M.bx = e->grid.bx;
    M.n_chunks = e->n_chunks;
    return M;
}
unsigned fj_row_grid(const RowMap& M, int warps) { return (M.n_rows / unsigned(warps)) * unsigned(M.n_chunks); }
int fj_owned_classes(const FjsphEngine* e) { return (e->slab.on && e->slab.world > 1) ? 2 : 1; }

int fj_permute_levels(FjsphEngine* e)
{
    const int n = int(e->n);
    const int nb = fj_blocks(n, TPB);
    {
        KScope ks(e, "permute", 3);
        k_permute_level<<<nb, TPB, 0, e->stream>>>(e->lv[0], e->lv[2], e->perm, n);
        std::swap(e->lv[0], e->lv[2]);
        k_permute_level<<<nb, TPB, 0, e->stream>>>(e->lv[1], e->lv[2], e->perm, n);
        std::swap(e->lv[1], e->lv[2]);
        k_permute_index<<<nb, TPB, 0, e->stream>>>(e->oidx, e->blk, e->perm, n, e->oidx_tmp, e->blk_tmp, e->slot_of);
        std::swap(e->oidx, e->oidx_tmp);
        std::swap(e->blk, e->blk_tmp);
    }
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}

static int ensure_key_capacity(FjsphEngine* e, size_t n_keys)
{
    if (n_keys <= e->key_cap)
        return FJSPH_OK;
    if (n_keys > (size_t(1) << 29))
    {
        fj_set_error("cell table needs %zu keys (> 2^29): domain too sparse for the dense cell table", n_keys);
        return FJSPH_ERR_CAPACITY;
    }
    if (e->cell_count)
        cudaFree(e->cell_count);
    if (e->cell_start)
        cudaFree(e->cell_start);
    if (e->scan_tmp)
        cudaFree(e->scan_tmp);
    e->cell_count = e->cell_start = e->scan_tmp = nullptr;
    size_t cap = 1;
    while (cap < n_keys) cap <<= 1;
    FJ_CUDA(cudaMalloc(&e->cell_count, cap * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->cell_start, (cap + 1) * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->scan_tmp, (cap / SCAN_TILE + 2) * sizeof(unsigned)));
    e->key_cap = cap;
    return FJSPH_OK;
}

static int ensure_row_capacity(FjsphEngine* e, size_t n_rows)
{
    if (n_rows <= e->row_cap)
        return FJSPH_OK;
    if (e->row_warps)
        cudaFree(e->row_warps);
    if (e->warp_start)
        cudaFree(e->warp_start);
    if (e->row_scan_tmp)
        cudaFree(e->row_scan_tmp);
    e->row_warps = e->warp_start = e->row_scan_tmp = nullptr;
    e->row_cap = 0;
    FJ_CUDA(cudaMalloc(&e->row_warps, (n_rows + 4) * sizeof(unsigned))); /* + the longest-row statistic behind the rows */
    FJ_CUDA(cudaMalloc(&e->warp_start, (n_rows + 1) * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->row_scan_tmp, (n_rows / SCAN_TILE + 2) * sizeof(unsigned)));
    e->row_cap = n_rows;
    return FJSPH_OK;
}

// Run arrays for n_warp work warps.  The skin runs are sized for the worst case (scap slots per warp: every neighbouring
// row of the disc, every class); the exact runs -- a subset of the skin runs, slot for slot -- for the most slots any warp
// was SEEN to use in the skin build (ensure_exact_capacity, after it).  The per-warp partials of the prestep's npd sum live
// in e->red, which grows with the warps.
static int ensure_run_capacity(FjsphEngine* e, size_t n_warp, int scap)
{
    if (n_warp > e->run_warps_cap || scap > e->scap)
    {
        for (void* p : {(void*)e->srun, (void*)e->erun, (void*)e->srows, (void*)e->erows})
            if (p)
                cudaFree(p);
        e->srun = nullptr;
        e->erun = nullptr;
        e->srows = e->erows = nullptr;
        e->run_warps_cap = 0;
        e->scap = e->ecap = 0;
        const size_t nw = std::max(n_warp + n_warp / 16 + 64, e->run_warps_cap);
        const size_t slots = nw * size_t(scap) * 32u;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (slots * 4u + (64u << 20) > free_b)
        {
            fj_set_error("neighbour runs need %.1f GB (%zu work warps x %d row slots) but %.1f GB are free: rows too short "
                         "for this layout (a sheet of particles across the row axis?)",
                         slots * 4.0 / 1e9, nw, scap, free_b / 1e9);
            return FJSPH_ERR_CAPACITY;
        }
        FJ_CUDA(cudaMalloc(&e->srun, slots * sizeof(unsigned)));
        FJ_CUDA(cudaMalloc(&e->srows, nw * sizeof(int)));
        FJ_CUDA(cudaMalloc(&e->erows, nw * sizeof(int)));
        e->run_warps_cap = nw;
        e->scap = scap;
    }
    if (n_warp + 16 > e->red_cap)
    {
        if (e->red)
            cudaFree(e->red);
        e->red = nullptr;
        e->red_cap = 0;
        const size_t want = n_warp + n_warp / 16 + 1024;
        FJ_CUDA(cudaMalloc(&e->red, want * sizeof(double)));
        e->red_cap = want;
    }
    return FJSPH_OK;
}

static int ensure_exact_capacity(FjsphEngine* e, int slots_used)
{
    const int want = std::max(slots_used, 1);
    if (e->erun && want <= e->ecap)
        return FJSPH_OK;
    if (e->erun)
        cudaFree(e->erun);
    e->erun = nullptr;
    e->ecap = 0;
    const int ecap = std::min(e->scap, want + want / 8 + 2); /* a margin, so that a slowly growing fluid does not reallocate at every build */
    const size_t slots = e->run_warps_cap * size_t(ecap) * 32u;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (slots * sizeof(uint2) + (64u << 20) > free_b)
    {
        fj_set_error("neighbour runs need %.1f GB (%zu work warps x %d row slots) but %.1f GB are free", slots * 8.0 / 1e9,
                     e->run_warps_cap, ecap, free_b / 1e9);
        return FJSPH_ERR_CAPACITY;
    }
    FJ_CUDA(cudaMalloc(&e->erun, slots * sizeof(uint2)));
    e->ecap = ecap;
    return FJSPH_OK;
}

// full rebuild: bounds are in h_red[0..5]
static int rebuild_skin(FjsphEngine* e)
{
    const int n = int(e->n);
    const int nb = fj_blocks(n, TPB);
    double lo[3] = {e->h_red[0], e->h_red[1], e->h_red[2]}, hi[3] = {e->h_red[3], e->h_red[4], e->h_red[5]};
    for (int d = 0; d < 3; ++d)
        if (!(std::isfinite(lo[d]) && std::isfinite(hi[d])))
        {
            fj_set_error("build_neighbours: non-finite particle positions");
            return FJSPH_ERR_STATE;
        }
    // Row axis u = the longest extent (long rows fill their 32-lane chunks; FJSPH_B200_ROW_AXIS = 0|1|2 pins it).  Cell
    // edge along u a hair above the skin radius 2H + skin, so +-1 cell always covers it; along v and w about one
    // particle spacing (visiting +-ry, +-rz rows), with the origin half a cell below the lowest particle so that the
    // rows of a lattice-born fluid sit in the middle of their pencils.
    Grid g;
    int ax0 = 0;
    for (int d = 1; d < 3; ++d)
        if (hi[d] - lo[d] > 1.0001 * (hi[ax0] - lo[ax0]))
            ax0 = d;
    /* slab mode cuts the fluid along x into interior | edge | ghost classes, each with rows of its own: rows along x would
       be a handful of particles long in the edge and ghost classes (2H + skin wide), leaving most lanes of their warps
       idle, so the rows run along the longer transverse axis whenever that is long enough to fill warps */
    if (e->slab.on && e->slab.world > 1)
    {
        const int t = (hi[2] - lo[2] > 1.0001 * (hi[1] - lo[1])) ? 2 : 1;
        if (e->P.particle_step > 0.0 && hi[t] - lo[t] >= 64.0 * e->P.particle_step)
            ax0 = t;
    }
    if (e->row_axis >= 0 && e->row_axis < 3)
        ax0 = e->row_axis;
    g.ax0 = ax0;
    g.ax1 = (ax0 + 1) % 3;
    g.ax2 = (ax0 + 2) % 3;
    if (g.ax1 > g.ax2)
        std::swap(g.ax1, g.ax2);
    const double r_skin = std::sqrt(e->P.sr) + e->skin;
    const double cell = r_skin * (1.0 + 1e-7);
    double pw = cell;
    if (e->P.particle_step > 0.0)
        pw = std::min(cell, std::max(e->P.particle_step, cell / 8.0));
    if (e->row_width_cells > 0.0)
        pw = std::min(cell, std::max(e->row_width_cells * e->P.particle_step, cell / 8.0));
    for (;;)
    {
        const bool narrow = pw < cell;
        g.inv_cell = 1.0 / cell;
        g.inv_cy = g.inv_cz = 1.0 / pw;
        g.pw2 = pw * pw;
        g.r_skin2 = r_skin * r_skin;
        g.ry = g.rz = narrow ? int(std::ceil(cell / pw)) : 1;
        g.ox = lo[g.ax0];
        g.oy = lo[g.ax1] - (narrow ? 0.5 * pw : 0.0);
        g.oz = lo[g.ax2] - (narrow ? 0.5 * pw : 0.0);
        g.nx = int(std::floor((hi[g.ax0] - g.ox) * g.inv_cell)) + 1;
        g.ny = int(std::floor((hi[g.ax1] - g.oy) * g.inv_cy)) + 1;
        g.nz = int(std::floor((hi[g.ax2] - g.oz) * g.inv_cz)) + 1;
        g.bx = bits_for(g.nx);
        g.by = std::max(2, bits_for(g.ny)); /* at least 4 x 4 rows: a sweep CTA holds FJ_ROW_WARPS adjacent rows */
        g.bz = std::max(2, bits_for(g.nz));
        /* the key tables (count + start per class) stay below ~1 GB: wider rows for very large cross-sections */
        if (!narrow || g.bx + g.by + g.bz <= e->max_key_bits)
            break;
        pw = std::min(cell, 2.0 * pw);
    }
    if (g.bx + g.by + g.bz > 29)
    {
        fj_set_error("cell grid %d x %d x %d needs more than 2^29 keys", g.nx, g.ny, g.nz);
        return FJSPH_ERR_CAPACITY;
    }
    g.n_keys = 1u << (g.bx + g.by + g.bz);
    /* slab mode: interior | edge | ghost classes, each with its own copy of the cell table */
    const bool three = e->slab.on && e->slab.world > 1;
    const int n_class = three ? 3 : 1;
    const unsigned n_tab = unsigned(n_class) * g.n_keys;
    const double wg = std::sqrt(e->P.sr) + e->skin;
    const double x_edge_lo = (three && e->slab.rank > 0) ? e->slab.x_lo + wg : -1e300;
    const double x_edge_hi = (three && e->slab.rank < e->slab.world - 1) ? e->slab.x_hi - wg : 1e300;
    int st = ensure_key_capacity(e, n_tab);
    if (st)
        return st;
    // key = row bits << bx | u-cell.  Row-id bit order: v0, w0, v1, w1 (the FJ_ROW_WARPS rows of a sweep CTA are the 4 x 2
    // block of adjacent rows under the lowest three), then v2-4, w2-4 (consecutive CTAs stay inside a 32 x 32 block of
    // rows, whose records L2 keeps), then the rest of v and of w.
    {
        const int need = std::max(g.ny, g.nz);
        if (need > e->mtab_cap)
        {
            if (e->mtab_y)
                cudaFree(e->mtab_y);
            int cap = 64;
            while (cap < need) cap <<= 1;
            FJ_CUDA(cudaMalloc(&e->mtab_y, size_t(2) * cap * sizeof(unsigned)));
            e->mtab_z = e->mtab_y + cap;
            e->mtab_cap = cap;
        }
        int pos[2][32];
        int out = g.bx;
        const int bits[2] = {g.by, g.bz};
        auto take = [&](int a, int l0, int l1) {
            for (int l = l0; l < std::min(l1, bits[a]); ++l) pos[a][l] = out++;
        };
        take(0, 0, 1);
        take(1, 0, 1);
        take(0, 1, 2);
        take(1, 1, 2);
        take(0, 2, 5);
        take(1, 2, 5);
        take(0, 5, 32);
        take(1, 5, 32);
        std::vector<unsigned> tab(size_t(2) * e->mtab_cap, 0u);
        const int dims[2] = {g.ny, g.nz};
        for (int a = 0; a < 2; ++a)
            for (int c = 0; c < dims[a]; ++c)
            {
                unsigned v = 0;
                for (int l = 0; l < bits[a]; ++l)
                    if (c & (1 << l))
                        v |= 1u << pos[a][l];
                tab[size_t(a) * e->mtab_cap + c] = v;
            }
        FJ_CUDA(cudaMemcpyAsync(e->mtab_y, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice,
                                e->stream));
        // neighbouring rows inside the disc of radius 2H + skin, ascending w then v: a row whose cross-section lies
        // wholly outside the disc holds nobody
        std::vector<int2> offs;
        for (int dz = -g.rz; dz <= g.rz; ++dz)
            for (int dy = -g.ry; dy <= g.ry; ++dy)
            {
                const int gy = std::max(std::abs(dy) - 1, 0), gz = std::max(std::abs(dz) - 1, 0);
                if (double(gy * gy + gz * gz) * g.pw2 > g.r_skin2)
                    continue;
                offs.push_back(make_int2(dy, dz));
            }
        if (!e->row_off)
            FJ_CUDA(cudaMalloc(&e->row_off, size_t(17 * 17) * sizeof(int2)));
        e->n_row_off = int(offs.size());
        FJ_CUDA(cudaMemcpyAsync(e->row_off, offs.data(), offs.size() * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream)); // tab and offs are stack-lifetime buffers
    }
    e->grid = g;

    // counting sort by cell key
    Level& S = e->lv[1];
    {
        KScope ks(e, "nb_sort", 7);
        FJ_CUDA(cudaMemsetAsync(e->cell_count, 0, size_t(n_tab) * sizeof(unsigned), e->stream));
        k_key_hist<<<nb, TPB, 0, e->stream>>>(S.P0, n, int(e->n_owned), g, e->mtab_y, e->mtab_z, e->key, e->rank_in_cell,
                                              e->cell_count, n_class, x_edge_lo, x_edge_hi);
        prim_exclusive_scan(e->stream, e->cell_count, e->cell_start, n_tab, e->scan_tmp);
        k_scatter<<<nb, TPB, 0, e->stream>>>(e->key, e->rank_in_cell, e->cell_start, n, e->perm2);
        k_cell_order<<<fj_blocks(int64_t(n_tab) * 32, TPB), TPB, 0, e->stream>>>(e->cell_start, n_tab, e->perm2, e->oidx,
                                                                                 S.P0, g.ax0, e->perm);
    }
    FJ_CUDA(cudaGetLastError());

    // move both time levels into cell order
    st = fj_permute_levels(e);
    if (st)
        return st;

    // work warps: 32-particle chunks of the owned rows
    const unsigned n_rowkeys = 1u << (g.by + g.bz);
    const unsigned n_rows = unsigned(n_class) * n_rowkeys, n_owned_rows = unsigned(three ? 2 : 1) * n_rowkeys;
    st = ensure_row_capacity(e, n_rows);
    if (st)
        return st;
    {
        KScope ks(e, "nb_sort", 4);
        unsigned* stats = e->row_warps + n_rows;
        FJ_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned), e->stream));
        k_row_warps<<<fj_blocks(n_rows, TPB), TPB, 0, e->stream>>>(e->cell_start, g.bx, n_rows, n_owned_rows, e->row_warps,
                                                                   stats);
        prim_exclusive_scan(e->stream, e->row_warps, e->warp_start, n_rows, e->row_scan_tmp);
        FJ_CUDA(cudaMemcpyAsync(e->h_flag + 1, e->warp_start + n_rows, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
        FJ_CUDA(cudaMemcpyAsync(e->h_flag + 3, stats, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
        if (three) /* first slot of the edge class = number of interior particles */
            FJ_CUDA(cudaMemcpyAsync(e->h_flag, e->cell_start + g.n_keys, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
    }
    FJ_CUDA(cudaGetLastError());
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    e->n_warp = unsigned(e->h_flag[1]);
    e->n_chunks = std::max(1, (e->h_flag[3] + 31) / 32);
    e->slab.n_interior = three ? int64_t(e->h_flag[0]) : 0;

    // skin runs (retry with more row slots per warp on overflow)
    {
        KScope ks(e, "nb_skin", 1);
        k_snapshot<<<nb, TPB, 0, e->stream>>>(S.P0, e->xref, n);
    }
    int scap = std::max(e->scap, e->n_row_off + (n_class > 1 ? 40 : 4));
    for (int attempt = 0; attempt < 4; ++attempt)
    {
        st = ensure_run_capacity(e, e->n_warp, scap);
        if (st)
            return st;
        FJ_CUDA(cudaMemsetAsync(e->d_flag, 0, 2 * sizeof(int), e->stream));
        const RowMap M = fj_row_map(e, 0, three ? 2 : 1);
        {
            KScope ks(e, "nb_skin", 1);
            k_build_skin_runs<<<fj_row_grid(M), FJ_ROW_WARPS * 32, 0, e->stream>>>(
                S.P0, M, g, e->mtab_y, e->mtab_z, e->row_off, e->n_row_off, n_class, r_skin * r_skin, e->scap, e->srun,
                e->srows, e->d_flag);
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream));
        if (e->h_flag[0] == 0)
        {
            st = ensure_exact_capacity(e, e->h_flag[1]);
            if (st)
                return st;
            e->skin_valid = true;
            e->skin_n = e->n;
            e->skin_builds++;
            return FJSPH_OK;
        }
        scap = e->h_flag[0] + 8;
    }
    fj_set_error("skin run capacity could not be satisfied");
    return FJSPH_ERR_CAPACITY;
}

int fj_build_neighbours(FjsphEngine* e)
{
    const int n = int(e->n);
    if (n <= 0)
    {
        fj_set_error("build_neighbours: no particles uploaded");
        return FJSPH_ERR_STATE;
    }
    e->list_valid = false;
    const int nb = fj_blocks(n, TPB);
    const bool have_skin = e->skin_valid && e->skin_n == e->n && e->skin > 0.0;

    // 1. bounds and the largest displacement since the skin build (one 56-byte readback)
    {
        KScope ks(e, "nb_bounds", 2);
        const int rb = std::min(nb, 1024);
        k_bounds<<<rb, TPB, 0, e->stream>>>(e->lv[1].P0, have_skin ? e->xref : nullptr, n, e->red);
        k_bounds_final<<<1, NRED * 32, 0, e->stream>>>(e->red, rb, e->red_out);
    }
    FJ_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, NRED * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    /* The superset list holds every pair closer than 2H + skin at its build, so it stays valid while no PAIR has
       closed in by more than skin: either nobody moved more than skin/2 (a hair below, for rounding), or -- a fluid
       drifting as a whole -- every displacement lies within skin/2 of the centre of the displacements' bounding box.
       Ownership and ghost sets are re-made with the list, so the same test covers the slab decomposition. */
    const double lim = 0.49 * e->skin;
    double disp[7] = {e->h_red[6], e->h_red[10], e->h_red[11], e->h_red[12], -e->h_red[7], -e->h_red[8], -e->h_red[9]};
    if (!have_skin)
        disp[0] = 1e300;
    if (e->slab.on)
    {
        /* all ranks rebuild together: the ghost sets are only re-made with the superset lists */
        int st = fj_allreduce(e, FJSPH_COMM_MAX, disp, 7);
        if (st)
            return st;
    }
    double half_diag2 = 0.0;
    for (int d = 0; d < 3; ++d)
    {
        const double h = 0.5 * (disp[1 + d] + disp[4 + d]); /* (max - min) / 2 of component d */
        half_diag2 += h * h;
    }
    const bool valid = have_skin && (disp[0] <= lim * lim || half_diag2 <= lim * lim);
    double moved = valid ? 0.0 : 1.0;
    if (moved != 0.0)
    {
        /* the re-sort moves every field of both time levels: a split upload (fjsph_step_host) has to land first */
        int stw = fj_upload_wait(e);
        if (stw)
            return stw;
        if (e->slab.on && e->slab.world > 1)
        {
            int st = fj_redecompose(e); /* migration + new ghost sets; particle counts change */
            if (st)
                return st;
            const int n2 = int(e->n);
            const int rb = std::min(fj_blocks(n2, TPB), 1024);
            k_bounds<<<rb, TPB, 0, e->stream>>>(e->lv[1].P0, nullptr, n2, e->red);
            k_bounds_final<<<1, NRED * 32, 0, e->stream>>>(e->red, rb, e->red_out);
            e->launches += 2;
            FJ_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
            FJ_CUDA(cudaStreamSynchronize(e->stream));
        }
        int st = rebuild_skin(e);
        if (st)
            return st;
    }

    // 2. exact runs, and the positions they were built on (every pair loop takes r from them: frozen pair distances)
    {
        KScope ks(e, "nb_list", 2);
        const RowMap M = fj_row_map(e, 0, fj_owned_classes(e));
        unsigned long long* stats = nullptr;
        if (e->list_stats)
        {
            stats = reinterpret_cast<unsigned long long*>(e->red_out) + 8; /* red_out holds 16 doubles: the upper half */
            FJ_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), e->stream));
        }
        k_exact_runs<<<fj_row_grid(M), FJ_ROW_WARPS * 32, 0, e->stream>>>(e->lv[1].P0, M, e->srun, e->srows, e->scap, e->P.sr,
                                                                         e->ecap, e->erun, e->erows, e->ncount, stats);
        k_snapshot<<<fj_blocks(int(e->n), TPB), TPB, 0, e->stream>>>(e->lv[1].P0, e->x0, int(e->n));
    }
    FJ_CUDA(cudaGetLastError());
    if (e->list_stats)
    {
        /* FJSPH_B200_LIST_STATS=1: how well the lockstep walk is filled (diagnostic, synchronises) */
        unsigned long long h[4];
        FJ_CUDA(cudaMemcpyAsync(h, reinterpret_cast<unsigned long long*>(e->red_out) + 8, sizeof(h), cudaMemcpyDeviceToHost,
                                e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream));
        std::fprintf(stderr,
                     "[fjsph_b200] list: %llu pairs in %llu lane-steps (fill %.3f), %.1f slots and %.1f steps per warp, %llu work "
                     "warps for %lld particles (lane use %.3f), row axis %d, %d x %d x %d cells, %d row offsets\n",
                     h[1], h[0], h[0] ? double(h[1]) / double(h[0]) : 0.0, h[3] ? double(h[2]) / double(h[3]) : 0.0,
                     h[3] ? double(h[0]) / 32.0 / double(h[3]) : 0.0, h[3], (long long)e->n_owned,
                     h[3] ? double(e->n_owned) / (32.0 * double(h[3])) : 0.0, e->grid.ax0, e->grid.nx, e->grid.ny, e->grid.nz,
                     e->n_row_off);
    }
    e->x_moved = false;
    e->list_valid = true;
    e->nb_builds++;
    return FJSPH_OK;
}

// debug / parity view of the exact list (fjsph_get_neighbours): d_offsets = exclusive prefix sum of count + 1 over
// caller indices, d_idx receives the caller indices of the neighbours and of the particle itself (unsorted)
int fj_neighbours_to_csr(FjsphEngine* e, const long long* d_offsets, long long* d_idx, int* d_bad)
{
    const RowMap M = fj_row_map(e, 0, fj_owned_classes(e));
    k_runs_to_csr<<<fj_row_grid(M), FJ_ROW_WARPS * 32, 0, e->stream>>>(M, e->erun, e->erows, e->ecap, e->oidx, d_offsets,
                                                                       d_idx, d_bad);
    FJ_CUDA(cudaGetLastError());
    return FJSPH_OK;
}
