// l1_gather_probe.cu -- how many L1 data-pipe wavefronts does one warp-wide gather cost on sm_100a, as a function of
// the access pattern and the load width?  Each kernel issues ONE kind of load in a loop; run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,\
//       l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum tools/bin/l1_gather_probe
// and divide wavefronts by requests.  Patterns (per lane l, record index within a small L1-resident table):
//   0 coalesced      : l                       (32 sectors in 8 lines for 32-byte records)
//   1 stride 2       : 2 l                     (32 sectors in 16 lines)
//   2 stride 4       : 4 l                     (32 sectors in 32 lines)
//   3 pairs share    : 4 (l / 2)               (16 sectors in 16 lines)
//   4 runs of 2      : 8 (l / 2) + (l & 1)     (32 sectors in 16 lines, two adjacent sectors per line)
//   5 runs of 4      : 16 (l / 4) + (l & 3)    (32 sectors in 8 lines, every 4th line full)
//   6 random         : hashed                  (~30 sectors in ~30 lines)
//   7 run + shift    : l + 5 (l / 8)           (pencil-like: four runs of 8 with gaps)
//   8-11 random lines with the 32-byte column (address bits 6:5) picked per lane: l mod 4 | l / 8 | (l / 4) mod 4 |
//        two columns per half-warp -- separates "wavefront per line" from "wavefront per column conflict"
#include <cstdio>
#include <cuda_runtime.h>

template <int BYTES>
struct Vec;
template <>
struct Vec<32>
{
    double4 v;
};
template <>
struct Vec<16>
{
    double2 v;
};
template <>
struct Vec<8>
{
    double v;
};

__device__ __forceinline__ int pattern(int p, int l, int it)
{
    switch (p)
    {
    case 0: return l;
    case 1: return 2 * l;
    case 2: return 4 * l;
    case 3: return 4 * (l / 2);
    case 4: return 8 * (l / 2) + (l & 1);
    case 5: return 16 * (l / 4) + (l & 3);
    case 6: return int((unsigned(l * 2654435761u + it * 40503u) >> 7) & 1023u);
    case 7: return l + (l / 8) * 5;
    /* random lines, the 32-byte column inside the line chosen by the lane: */
    case 8: return 4 * int((unsigned(l * 2654435761u + it * 40503u) >> 7) & 255u) + (l & 3);         /* l mod 4 */
    case 9: return 4 * int((unsigned(l * 2654435761u + it * 40503u) >> 7) & 255u) + ((l >> 3) & 3);  /* quarter-warp */
    case 10: return 4 * int((unsigned(l * 2654435761u + it * 40503u) >> 7) & 255u) + ((l >> 2) & 3); /* groups of 4 */
    default: return 4 * int((unsigned(l * 2654435761u + it * 40503u) >> 7) & 255u) + ((l >> 4) * 2 + (l & 1)); /* half-warps use 2 columns each */
    }
}

template <int BYTES>
__global__ void probe(const char* __restrict__ table, int p, int iters, double* out)
{
    const int l = threadIdx.x & 31;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it)
    {
        const int rec = (pattern(p, l, it) + (it & 7) * 128) & 4095; /* 32-byte record slots, table = 128 KB */
        const char* a = table + size_t(rec) * 32;
        if (BYTES == 32)
        {
            double4 v;
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(a));
            acc += v.x + v.w;
        }
        else if (BYTES == 16)
        {
            double2 v;
            asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a));
            acc += v.x + v.y;
        }
        else
        {
            double v;
            asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(a));
            acc += v;
        }
    }
    if (acc == 12345.678)
        out[0] = acc;
}

// SoA flavour: 8-byte loads from a packed array of doubles (4 records' worth per sector)
__global__ void probe_soa(const double* __restrict__ table, int p, int iters, double* out)
{
    const int l = threadIdx.x & 31;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it)
    {
        const int rec = (pattern(p, l, it) + (it & 7) * 128) & 4095;
        double v;
        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(table + rec));
        acc += v;
    }
    if (acc == 12345.678)
        out[0] = acc;
}

int main()
{
    char* table;
    double* out;
    cudaMalloc(&table, 1 << 20);
    cudaMemset(table, 0, 1 << 20);
    cudaMalloc(&out, 64);
    const int iters = 4096;
    for (int p = 0; p < 12; ++p)
    {
        probe<32><<<148, 128>>>(table, p, iters, out);
        probe<16><<<148, 128>>>(table, p, iters, out);
        probe<8><<<148, 128>>>(table, p, iters, out);
        probe_soa<<<148, 128>>>((const double*)table, p, iters, out);
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
