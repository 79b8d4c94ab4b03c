// mufu_probe.cu — accuracy of the FP64 reciprocal / reciprocal-square-root sequences the pair sweeps use:
// the MUFU.RSQ64H / MUFU.RCP64H seeds, one third-order step (fj_rsqrt3), one and two Newton steps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_probe tools/mufu_probe.cu && tools/bin/mufu_probe
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ double seed_rsqrt(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ double seed_rcp(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ double rsqrt3(double x)
{
    const double y0 = seed_rsqrt(x), t = x * y0, e = fma(-t, y0, 1.0), p = fma(0.375, e, 0.5), ye = y0 * e;
    return fma(ye, p, y0);
}
__device__ double rsqrt_n2(double x)
{
    double y = seed_rsqrt(x);
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}
__device__ double rcp_n(double x, int steps)
{
    double y = seed_rcp(x);
    for (int k = 0; k < steps; ++k)
    {
        const double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}
__global__ void probe(int n, double lo, double hi, double* out)
{
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const double x = lo * pow(hi / lo, (i + 0.5) / n);
        const double rs = 1.0 / sqrt(x), rc = 1.0 / x;
        const double v[6] = {seed_rsqrt(x), rsqrt3(x), rsqrt_n2(x), seed_rcp(x), rcp_n(x, 1), rcp_n(x, 2)};
        for (int k = 0; k < 6; ++k) m[k] = fmax(m[k], fabs(v[k] / (k < 3 ? rs : rc) - 1.0));
    }
    for (int k = 0; k < 6; ++k) atomicMax(reinterpret_cast<unsigned long long*>(out + k), __double_as_longlong(m[k]));
}
int main()
{
    double* d;
    cudaMalloc(&d, 6 * sizeof(double));
    const char* names[6] = {"rsqrt seed", "rsqrt third-order step", "rsqrt two Newton steps", "rcp seed", "rcp one Newton step",
                            "rcp two Newton steps"};
    const double ranges[3][2] = {{1e-12, 1e-4}, {0.3, 3.0}, {1e2, 1e12}};
    for (auto& r : ranges)
    {
        cudaMemset(d, 0, 6 * sizeof(double));
        probe<<<592, 256>>>(1 << 24, r[0], r[1], d);
        double h[6];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        std::printf("x in [%g, %g]:\n", r[0], r[1]);
        for (int k = 0; k < 6; ++k) std::printf("   %-26s max rel err %.3e\n", names[k], h[k]);
    }
    return 0;
}
