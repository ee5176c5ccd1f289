// Microbenchmark, not product code: what one B200 sustains in FP64 -- DMMA.8x8x4 (mma.sync.m8n8k4.f64) against plain DFMA --
// as a function of the warps per SM.  Decides how the trailing update of the dense core LU (csrc/pe_b200_frontal.cu) is written.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cuda_runtime.h>
#include <cstdio>

template <int ACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0)
{
    double c[ACC][2];
#pragma unroll
    for(int i = 0; i < ACC; ++i) { c[i][0] = c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    for(int it = 0; it < iters; ++it)
    {
#pragma unroll
        for(int i = 0; i < ACC; ++i)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for(int i = 0; i < ACC; ++i) { s += c[i][0] + c[i][1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0)
{
    double c[ACC];
#pragma unroll
    for(int i = 0; i < ACC; ++i) { c[i] = i * 1e-3; }
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    for(int it = 0; it < iters; ++it)
    {
#pragma unroll
        for(int i = 0; i < ACC; ++i) { c[i] = fma(a, c[i], b); }
    }
    double s = 0.0;
#pragma unroll
    for(int i = 0; i < ACC; ++i) { s += c[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int const sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int const iters = 20000;
    std::printf("{\"sms\": %d, \"points\": [", sms);
    bool first = true;
    for(int warps: {4, 8, 16, 32})
    {
        for(int kind = 0; kind < 2; ++kind)
        {
            float ms = 0.f;
            for(int rep = 0; rep < 2; ++rep)
            {
                cudaEventRecord(e0);
                if(kind == 0) { dmma_kernel<16><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999); }
                else { dfma_kernel<16><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999); }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            double const flops = kind == 0 ? (double)sms * warps * iters * 16 * 512.0 : (double)sms * warps * 32 * iters * 16 * 2.0;
            std::printf("%s{\"op\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}", first ? "" : ", ", kind == 0 ? "DMMA.8x8x4" : "DFMA", warps, ms, flops / ms / 1e9);
            first = false;
        }
    }
    std::printf("]}\n");
    return 0;
}
