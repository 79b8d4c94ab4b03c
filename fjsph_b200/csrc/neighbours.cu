// neighbours.cu — uniform cell list replacing FJSPH's nanoflann KD-tree radius search.
//
// Replaces update_neighbours / find_neighbours / radius_search (reference src/Neighbours.cpp:7-47):
//   list_i = { j : ((xi-xj)^2 + (yi-yj)^2) + (zi-zj)^2 < sr }, strict '<', evaluated WITHOUT fma
// contraction in the order nanoflann's metric_L2_Simple accumulates it, so the sets are bit-exact with
// the CPU oracle.  Self is not stored (callers add the self terms explicitly; outlist[i].size() is
// count+1).
//
// Pipeline per build (all on e->stream):
//   bounds -> cell key (pencil cells, x fastest, by default; Morton keys for cubic cells: engine.cuh pencil_order)
//   + warp-aggregated histogram -> block prefix scan -> scatter
//   -> per-cell ordering by caller index (determinism) -> permute both time levels into cell order
//   -> neighbour list in a warp-transposed ELL layout (entry (warp w, slot s, lane l) at
//      ((w*nb_cap)+s)*32+l, so a warp reads slot s of its 32 particles with one 128-byte load).
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
__device__ __forceinline__ void cell_of(const Grid& g, double x, double y, double z, int& cx, int& cy, int& cz)
{
    cx = min(max(int(floor((x - g.ox) * g.inv_cell)), 0), g.nx - 1);
    cy = min(max(int(floor((y - g.oy) * g.inv_cy)), 0), g.ny - 1);
    cz = min(max(int(floor((z - g.oz) * g.inv_cz)), 0), g.nz - 1);
}

// Slab mode sorts three classes one behind the other (key offsets 0, n_keys, 2 n_keys): INTERIOR owned particles
// (x in [x_edge_lo, x_edge_hi): farther than 2H + skin from every face with a neighbour rank, so no ghost can be
// their neighbour), EDGE owned particles (the ones sent as ghosts, same predicate as k_ghost_flags) and GHOSTS.
__global__ void k_key_hist(const double4* __restrict__ P0, int n, int n_owned, Grid g, const unsigned* __restrict__ mx,
                           const unsigned* __restrict__ my, const unsigned* __restrict__ mz,
                           unsigned* __restrict__ key, unsigned* __restrict__ rank, unsigned* __restrict__ count,
                           int n_class, double x_edge_lo, double x_edge_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    unsigned k = 0xFFFFFFFFu;
    if (i < n)
    {
        double4 a = P0[i];
        int cx, cy, cz;
        cell_of(g, a.x, a.y, a.z, cx, cy, cz);
        k = mx[cx] | my[cy] | mz[cz];
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

// One warp per occupied cell segment: order its members by caller index (rank sort).  The atomics above
// leave an arbitrary order inside a cell; this makes the layout — and with it every FP64 summation
// order downstream — reproducible run to run.
__global__ void k_cell_order(const unsigned* __restrict__ cell_start, unsigned n_keys, const int* __restrict__ perm,
                             const int* __restrict__ oidx, int* __restrict__ perm2)
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
        unsigned r = 0;
        for (unsigned u = 0; u < c; ++u) r += (oidx[perm[s + u]] < mo) ? 1u : 0u;
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

// ---------------------------------------------------------------- skin list + exact list
// Two-level neighbour build.  The SKIN list holds every j with d < 2H + skin at the time it was built (27-cell
// sweep over the Morton cell list); it stays valid while no particle has moved more than skin/2 since.  The
// EXACT list -- the reference's OUTL: { j : d2 < sr } with r = sqrt(d2), rebuilt at every update_neighbours --
// is filtered from the skin list with the bit-exact nanoflann distance on the CURRENT positions, ~1.3 N_nb
// candidates per particle instead of the 27-cell sweep's ~7 N_nb.  Both use the chunked ELL layout.
__device__ __forceinline__ double4 ldg256(const double4* __restrict__ base, unsigned j)
{
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(base + j));
    return v;
}

// ---------------------------------------------------------------- column-aware order inside a chunk
// The L1 data pipe serves a warp-wide 256-bit gather four lanes at a time: each group of four consecutive lanes costs
// as many wavefronts as lanes of the group collide in the same 32-byte COLUMN of a 128-byte line (address bits 6:5,
// i.e. record index & 3) on different sectors, and one wavefront when the four records sit in four different columns
// -- whatever lines they are in (tools/l1_gather_probe.cu under ncu: 32 random lines cost 8.3 wavefronts with the
// columns arranged so, 16.4 with random columns, 32 in one column).  A lane may visit the four neighbours of a chunk
// in any order, so a neighbour with  index & 3 == (lane + e) & 3  goes to element e of its chunk where that element is
// still free: the four lanes of a group then gather from four different columns at that slot.  The walk
// through memory stays the ascending one (dealing whole lists out by column was measured too: the data-pipe wavefronts
// fall further, but the four column streams of a lane drift apart, the L1 hit rate collapses and L2 -> L1 traffic
// doubles; profiles/r5_*).
// Online form: an accepted neighbour takes element (index - lane) & 3 of the chunk being filled if that element is
// still free, else the lowest free one; the chunk is stored when its four elements are taken.
template <bool COLS>
__device__ __forceinline__ int chunk_slot(int lane, unsigned ent, unsigned freemask)
{
    if (COLS)
    {
        const int want = int((ent - unsigned(lane)) & 3u);
        if ((freemask >> want) & 1u)
            return want;
    }
    return __ffs(int(freemask)) - 1;
}

template <bool COLS>
__global__ void __launch_bounds__(TPB)
    k_build_skin(const double4* __restrict__ P0, const int* __restrict__ b, int n, int n_owned, Grid g,
                 const unsigned* __restrict__ mx, const unsigned* __restrict__ my, const unsigned* __restrict__ mz,
                 const unsigned* __restrict__ cell_start, double sr_skin, int scap, unsigned* __restrict__ slist,
                 int* __restrict__ scount, double4* __restrict__ xref, int* __restrict__ flag, int n_seg)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double4 a = P0[i];
    xref[i] = a;
    if (i >= n_owned)
        return; /* ghosts are neighbours only */
    int cx, cy, cz;
    cell_of(g, a.x, a.y, a.z, cx, cy, cz);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(slist) + (size_t(i >> 5) * size_t(scap >> 2)) * 32u + (i & 31);
    unsigned eb0 = unsigned(i), eb1 = unsigned(i), eb2 = unsigned(i), eb3 = unsigned(i);
    unsigned freemask = 15u; /* free elements of the chunk being filled */
    int cnt = 0;
    /* cells in ascending key order (z, then y, then x), so the list comes out sorted by index */
    for (int dz = -g.rz; dz <= g.rz; ++dz)
    {
        const int z = cz + dz;
        if (z < 0 || z >= g.nz)
            continue;
        const unsigned kz = mz[z];
        const int gz = max(abs(dz) - 1, 0);
        for (int dy = -g.ry; dy <= g.ry; ++dy)
        {
            const int y = cy + dy;
            if (y < 0 || y >= g.ny)
                continue;
            /* narrow cells: a pencil whose cross-section lies wholly outside the disc of radius 2H + skin holds nobody */
            const int gy = max(abs(dy) - 1, 0);
            if (double(gy * gy + gz * gz) * g.pw2 > g.r_skin2)
                continue;
            const unsigned kyz = kz | my[y];
            for (int dx = -1; dx <= 1; ++dx)
            {
                const int x = cx + dx;
                if (x < 0 || x >= g.nx)
                    continue;
                for (int seg = 0; seg < n_seg; ++seg)
                {
                const unsigned k = (kyz | mx[x]) + unsigned(seg) * g.n_keys;
                const unsigned s = cell_start[k], e = cell_start[k + 1];
                for (unsigned j = s; j < e; ++j)
                {
                    const double4 q = P0[j];
                    const double ddx = a.x - q.x, ddy = a.y - q.y, ddz = a.z - q.z;
                    const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                    if (d2 < sr_skin && int(j) != i)
                    {
                        const int bj = b[j];
                        unsigned ent = j;
                        if (bj > FJSPH_PISTON)
                            ent |= FJ_NB_FLUID;
                        if (bj == FJSPH_BOUND)
                            ent |= FJ_NB_BOUND;
                        const int kk = chunk_slot<COLS>(i & 31, ent, freemask);
                        if (kk == 0)
                            eb0 = ent;
                        else if (kk == 1)
                            eb1 = ent;
                        else if (kk == 2)
                            eb2 = ent;
                        else
                            eb3 = ent;
                        freemask &= ~(1u << kk);
                        if (freemask == 0u)
                        {
                            if (cnt < scap)
                                dst[size_t(cnt >> 2) * 32u] = make_uint4(eb0, eb1, eb2, eb3);
                            freemask = 15u;
                        }
                        cnt++;
                    }
                }
                }
            }
        }
    }
    if ((cnt & 3) && cnt < scap)
    {
        /* partial last chunk: its entries move to the front (consumers read `left` elements), the rest hold i itself */
        unsigned v[4] = {eb0, eb1, eb2, eb3}, o[4] = {unsigned(i), unsigned(i), unsigned(i), unsigned(i)};
        int k = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (!((freemask >> m) & 1u))
            {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t == k)
                        o[t] = v[m];
                k++;
            }
        dst[size_t(cnt >> 2) * 32u] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    scount[i] = cnt;
    if (cnt > scap)
        atomicMax(flag, cnt);
}

// exact list from the skin list: list_i = { j in skin_i : ((xi-xj)^2 + (yi-yj)^2) + (zi-zj)^2 < sr }
template <bool COLS>
__global__ void __launch_bounds__(TPB)
    k_exact_from_skin(const double4* __restrict__ P0, int n, const unsigned* __restrict__ slist,
                      const int* __restrict__ scount, int scap, double sr, int nb_cap, unsigned* __restrict__ nlist,
                      double* __restrict__ nr, int* __restrict__ ncount, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double4 a = P0[i];
    const uint4* __restrict__ sp =
        reinterpret_cast<const uint4*>(slist) + (size_t(i >> 5) * size_t(scap >> 2)) * 32u + (i & 31);
    const size_t base = (size_t(i >> 5) * size_t(nb_cap >> 2)) * 32u + (i & 31);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(nlist) + base;
    double4* __restrict__ rdst = reinterpret_cast<double4*>(nr) + base;
    const int sc = scount[i];
    const int nchunk = (sc + 3) >> 2;
    unsigned eb0 = unsigned(i), eb1 = unsigned(i), eb2 = unsigned(i), eb3 = unsigned(i);
    double rb0 = 0.0, rb1 = 0.0, rb2 = 0.0, rb3 = 0.0;
    unsigned freemask = 15u; /* free elements of the chunk being filled */
    int cnt = 0;
    auto test = [&](const unsigned ent, const double4 q, const bool valid) {
        // nanoflann metric_L2_Simple order, no fma contraction (bit-exact sets)
        const double ddx = __dsub_rn(a.x, q.x), ddy = __dsub_rn(a.y, q.y), ddz = __dsub_rn(a.z, q.z);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
        if (valid && d2 < sr)
        {
            const double rv = sqrt(d2); /* r = sqrt(jj.second), as every pair loop takes it */
            const int kk = chunk_slot<COLS>(i & 31, ent, freemask);
            if (kk == 0)
            {
                eb0 = ent;
                rb0 = rv;
            }
            else if (kk == 1)
            {
                eb1 = ent;
                rb1 = rv;
            }
            else if (kk == 2)
            {
                eb2 = ent;
                rb2 = rv;
            }
            else
            {
                eb3 = ent;
                rb3 = rv;
            }
            freemask &= ~(1u << kk);
            if (freemask == 0u)
            {
                if (cnt < nb_cap)
                {
                    dst[size_t(cnt >> 2) * 32u] = make_uint4(eb0, eb1, eb2, eb3);
                    rdst[size_t(cnt >> 2) * 32u] = make_double4(rb0, rb1, rb2, rb3);
                }
                freemask = 15u;
            }
            cnt++;
        }
    };
    if (nchunk > 0)
    {
        uint4 id = sp[0];
        for (int c = 0; c < nchunk; ++c)
        {
            uint4 idn = id;
            if (c + 1 < nchunk)
                idn = sp[size_t(c + 1) * 32u];
            /* four independent gathers in flight (unused slots of the last chunk hold i itself) */
            const double4 q0 = ldg256(P0, id.x & FJ_IDX_MASK);
            const double4 q1 = ldg256(P0, id.y & FJ_IDX_MASK);
            const double4 q2 = ldg256(P0, id.z & FJ_IDX_MASK);
            const double4 q3 = ldg256(P0, id.w & FJ_IDX_MASK);
            const int left = sc - (c << 2);
            test(id.x, q0, true);
            test(id.y, q1, left > 1);
            test(id.z, q2, left > 2);
            test(id.w, q3, left > 3);
            id = idn;
        }
    }
    if ((cnt & 3) && cnt < nb_cap)
    {
        /* partial last chunk: its entries move to the front (consumers read `left` elements); unused slots hold i itself
           (safe to gather) and are never consumed */
        unsigned v[4] = {eb0, eb1, eb2, eb3}, o[4] = {unsigned(i), unsigned(i), unsigned(i), unsigned(i)};
        double vr[4] = {rb0, rb1, rb2, rb3}, orr[4] = {0.0, 0.0, 0.0, 0.0};
        int k = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (!((freemask >> m) & 1u))
            {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t == k)
                    {
                        o[t] = v[m];
                        orr[t] = vr[m];
                    }
                k++;
            }
        dst[size_t(cnt >> 2) * 32u] = make_uint4(o[0], o[1], o[2], o[3]);
        rdst[size_t(cnt >> 2) * 32u] = make_double4(orr[0], orr[1], orr[2], orr[3]);
    }
    ncount[i] = cnt;
    if (cnt > nb_cap)
        atomicMax(flag, cnt);
}

int bits_for(int n)
{
    int b = 0;
    while ((1 << b) < n) b++;
    return b;
}

} // namespace

// ------------------------------------------------------------------ host orchestration
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
        fj_set_error("cell table needs %zu keys (> 2^29): domain too sparse for the dense Morton table", n_keys);
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

static int ensure_list_capacity(FjsphEngine* e, int nb_cap)
{
    const size_t words = size_t((e->cap + 31) / 32) * size_t(nb_cap) * 32u;
    if (nb_cap <= e->nb_cap && words <= e->nlist_words)
        return FJSPH_OK;
    if (e->nlist)
        cudaFree(e->nlist);
    if (e->nr)
        cudaFree(e->nr);
    e->nlist = nullptr;
    e->nr = nullptr;
    FJ_CUDA(cudaMalloc(&e->nlist, words * sizeof(unsigned)));
    FJ_CUDA(cudaMalloc(&e->nr, words * sizeof(double)));
    e->nlist_words = words;
    e->nb_cap = nb_cap;
    return FJSPH_OK;
}

static int ensure_skin_capacity(FjsphEngine* e, int scap)
{
    const size_t words = size_t((e->cap + 31) / 32) * size_t(scap) * 32u;
    if (scap <= e->scap && words <= e->slist_words)
        return FJSPH_OK;
    if (e->slist)
        cudaFree(e->slist);
    e->slist = nullptr;
    FJ_CUDA(cudaMalloc(&e->slist, words * sizeof(unsigned)));
    e->slist_words = words;
    e->scap = scap;
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
    // grid: cell edge along x a hair above the skin radius 2H + skin so +-1 cell always covers it.  Pencil order
    // narrows the cells along y and z to about one particle spacing (and visits +-ry, +-rz of them), with the origin half
    // a cell below the lowest particle so that the rows of a lattice-born fluid sit in the middle of their pencils.
    Grid g;
    const double r_skin = std::sqrt(e->P.sr) + e->skin;
    const double cell = r_skin * (1.0 + 1e-7);
    double pw = cell;
    if (e->pencil_order && e->P.particle_step > 0.0)
        pw = std::min(cell, std::max(e->P.particle_step, cell / 8.0));
    for (;;)
    {
        const bool narrow = pw < cell;
        g.inv_cell = 1.0 / cell;
        g.inv_cy = g.inv_cz = 1.0 / pw;
        g.pw2 = pw * pw;
        g.r_skin2 = r_skin * r_skin;
        g.ry = g.rz = narrow ? int(std::ceil(cell / pw)) : 1;
        g.ox = lo[0];
        g.oy = lo[1] - (narrow ? 0.5 * pw : 0.0);
        g.oz = lo[2] - (narrow ? 0.5 * pw : 0.0);
        g.nx = int(std::floor((hi[0] - g.ox) * g.inv_cell)) + 1;
        g.ny = int(std::floor((hi[1] - g.oy) * g.inv_cy)) + 1;
        g.nz = int(std::floor((hi[2] - g.oz) * g.inv_cz)) + 1;
        g.bx = bits_for(g.nx);
        g.by = bits_for(g.ny);
        g.bz = bits_for(g.nz);
        /* the key tables (count + start per class) stay below ~1 GB: wider pencils for very large cross-sections */
        if (!narrow || g.bx + g.by + g.bz <= 25)
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
    // key tables, one per axis, OR-ed together: lexicographic fields (pencil order) or Morton spreading (bit l of each
    // axis placed round-robin x,y,z among the axes that still have bits)
    {
        const int need = std::max(g.nx, std::max(g.ny, g.nz));
        if (need > e->mtab_cap)
        {
            if (e->mtab_x)
                cudaFree(e->mtab_x);
            int cap = 64;
            while (cap < need) cap <<= 1;
            FJ_CUDA(cudaMalloc(&e->mtab_x, size_t(3) * cap * sizeof(unsigned)));
            e->mtab_y = e->mtab_x + cap;
            e->mtab_z = e->mtab_y + cap;
            e->mtab_cap = cap;
        }
        int pos[3][32];
        int out = 0;
        const int bits[3] = {g.bx, g.by, g.bz};
        if (e->pencil_order)
        {
            /* x in the low bits, then y, then z -- consecutive keys run along x.  With tiles (pencil_tile_y > 0) the
               lowest TX bits of x come first, then the lowest TY bits of y, then the rest of x: memory then runs through
               ~one warp's worth of a pencil (2^TX cells), then through the same stretch of the next pencil, ..., so the
               warps of a block sit in ADJACENT pencils and walk nearly the same neighbour rows one row apart -- what
               one warp pulls into L1 the next one finds there. */
            const int tx = std::min(e->pencil_tile_x, bits[0]), ty = std::min(e->pencil_tile_y, bits[1]);
            for (int l = 0; l < tx; ++l) pos[0][l] = out++;
            for (int l = 0; l < ty; ++l) pos[1][l] = out++;
            for (int l = tx; l < bits[0]; ++l) pos[0][l] = out++;
            for (int l = ty; l < bits[1]; ++l) pos[1][l] = out++;
            for (int l = 0; l < bits[2]; ++l) pos[2][l] = out++;
        }
        else
            for (int l = 0; l < 32; ++l)
                for (int a = 0; a < 3; ++a)
                    if (l < bits[a])
                        pos[a][l] = out++;
        std::vector<unsigned> tab(size_t(3) * e->mtab_cap, 0u);
        const int dims[3] = {g.nx, g.ny, g.nz};
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < dims[a]; ++c)
            {
                unsigned v = 0;
                for (int l = 0; l < bits[a]; ++l)
                    if (c & (1 << l))
                        v |= 1u << pos[a][l];
                tab[size_t(a) * e->mtab_cap + c] = v;
            }
        FJ_CUDA(cudaMemcpyAsync(e->mtab_x, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice,
                                e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream)); // tab is a stack-lifetime buffer
    }
    e->grid = g;

    // counting sort by cell key
    Level& S = e->lv[1];
    {
        KScope ks(e, "nb_sort", 7);
        FJ_CUDA(cudaMemsetAsync(e->cell_count, 0, size_t(n_tab) * sizeof(unsigned), e->stream));
        k_key_hist<<<nb, TPB, 0, e->stream>>>(S.P0, n, int(e->n_owned), g, e->mtab_x, e->mtab_y, e->mtab_z, e->key, e->rank_in_cell,
                                              e->cell_count, n_class, x_edge_lo, x_edge_hi);
        prim_exclusive_scan(e->stream, e->cell_count, e->cell_start, n_tab, e->scan_tmp);
        k_scatter<<<nb, TPB, 0, e->stream>>>(e->key, e->rank_in_cell, e->cell_start, n, e->perm2);
        k_cell_order<<<fj_blocks(int64_t(n_tab) * 32, TPB), TPB, 0, e->stream>>>(e->cell_start, n_tab, e->perm2, e->oidx,
                                                                                 e->perm);
    }
    FJ_CUDA(cudaGetLastError());

    // move both time levels into cell order
    st = fj_permute_levels(e);
    if (st)
        return st;

    // skin list (retry with a larger per-particle capacity on overflow)
    if (e->scap == 0)
    {
        /* expected count N_nb (1 + skin/2H)^3 with N_nb ~ 270, plus headroom */
        const double f = r_skin / std::sqrt(e->P.sr);
        const int want = ((int(300.0 * f * f * f) + 31) / 32) * 32;
        st = ensure_skin_capacity(e, want);
        if (st)
            return st;
    }
    for (int attempt = 0; attempt < 4; ++attempt)
    {
        FJ_CUDA(cudaMemsetAsync(e->d_flag, 0, sizeof(int), e->stream));
        {
            KScope ks(e, "nb_skin", 1);
            if (e->column_order) /* column_order_chunk above */
                k_build_skin<true><<<nb, TPB, 0, e->stream>>>(e->lv[1].P0, e->lv[1].b, n, int(e->n_owned), g, e->mtab_x, e->mtab_y,
                                                              e->mtab_z, e->cell_start, r_skin * r_skin, e->scap, e->slist,
                                                              e->scount, e->xref, e->d_flag, n_class);
            else
                k_build_skin<false><<<nb, TPB, 0, e->stream>>>(e->lv[1].P0, e->lv[1].b, n, int(e->n_owned), g, e->mtab_x, e->mtab_y,
                                                               e->mtab_z, e->cell_start, r_skin * r_skin, e->scap, e->slist,
                                                               e->scount, e->xref, e->d_flag, n_class);
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        if (three) /* first slot of the edge class = number of interior particles */
            FJ_CUDA(cudaMemcpyAsync(e->h_flag + 1, e->cell_start + g.n_keys, sizeof(unsigned), cudaMemcpyDeviceToHost,
                                    e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream));
        e->slab.n_interior = three ? int64_t(e->h_flag[1]) & ~int64_t(31) : 0;
        if (e->h_flag[0] == 0)
        {
            e->skin_valid = true;
            e->skin_n = e->n;
            e->skin_builds++;
            return FJSPH_OK;
        }
        st = ensure_skin_capacity(e, ((e->h_flag[0] + 31) / 32) * 32 + 32);
        if (st)
            return st;
    }
    fj_set_error("skin list capacity could not be satisfied");
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

    // 2. exact list (retry with a larger per-particle capacity on overflow)
    if (e->nb_cap == 0)
    {
        int st = ensure_list_capacity(e, 288);
        if (st)
            return st;
    }
    for (int attempt = 0; attempt < 4; ++attempt)
    {
        FJ_CUDA(cudaMemsetAsync(e->d_flag, 0, sizeof(int), e->stream));
        {
            KScope ks(e, "nb_list", 1);
            if (e->column_order)
                k_exact_from_skin<true><<<fj_blocks(e->n_owned, TPB), TPB, 0, e->stream>>>(
                    e->lv[1].P0, int(e->n_owned), e->slist, e->scount, e->scap, e->P.sr, e->nb_cap, e->nlist, e->nr, e->ncount,
                    e->d_flag);
            else
                k_exact_from_skin<false><<<fj_blocks(e->n_owned, TPB), TPB, 0, e->stream>>>(
                    e->lv[1].P0, int(e->n_owned), e->slist, e->scount, e->scap, e->P.sr, e->nb_cap, e->nlist, e->nr, e->ncount,
                    e->d_flag);
        }
        FJ_CUDA(cudaGetLastError());
        FJ_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        FJ_CUDA(cudaStreamSynchronize(e->stream));
        if (e->h_flag[0] == 0)
        {
            e->list_valid = true;
            e->nb_builds++;
            return FJSPH_OK;
        }
        const int want = ((e->h_flag[0] + 31) / 32) * 32 + 32;
        int st = ensure_list_capacity(e, want);
        if (st)
            return st;
    }
    fj_set_error("neighbour list capacity could not be satisfied");
    return FJSPH_ERR_CAPACITY;
}
