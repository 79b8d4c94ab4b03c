// fp64_peak.cu — FP64 FMA throughput of the device ("the builder must measure it with an FMA microbenchmark").
// 8 independent DFMA chains per thread, 256 threads per block, enough blocks to fill every SM several times over.
// Prints one JSON line: {"fp64_tflops": ..., "sm_count": ..., "clock_mhz_nominal": ..., "ms": ...}.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 8, ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-9 + c;
    for (int it = 0; it < ITERS; ++it)
    {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) /* never true: keeps the chains alive */
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess)
    {
        std::fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int blocks = prop.multiProcessorCount * 32, threads = 256;
    double* out = nullptr;
    cudaMalloc(&out, size_t(blocks) * threads * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 10; ++r)
    {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double flop = 2.0 * double(blocks) * threads * CHAINS * ITERS;
    std::printf("{\"fp64_tflops\": %.3f, \"sm_count\": %d, \"clock_mhz_nominal\": %.0f, \"ms\": %.4f, \"how\": \"%d DFMA chains x %d "
                "iterations per thread, %d x %d threads, best of 10 (CUDA events)\"}\n",
                flop / (best * 1e-3) / 1e12, prop.multiProcessorCount, prop.clockRate / 1000.0, best, CHAINS, ITERS, blocks, threads);
    return 0;
}
