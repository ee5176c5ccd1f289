// Micro-benchmark: achievable HBM bandwidth for the access pattern of the tree-streaming kernel — every warp reads (and
// optionally writes) rows of ROWB bytes at pseudo-random row indices of a 1.5 GB buffer, ILP independent rows in flight.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ROWB, int ILP, bool WRITE>
__global__ void k(double* buf, uint64_t n_rows, int iters, double* sink)
{
    uint32_t const lane = threadIdx.x & 31;
    uint64_t const warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t s = warp * 0x9E3779B97F4A7C15ull + 12345;
    double acc = 0.0;
    constexpr int DPL = ROWB / 256;  // doubles per lane per row
    for(int it = 0; it < iters; ++it)
    {
        double v[ILP][DPL];
        uint64_t rows[ILP];
#pragma unroll
        for(int i = 0; i < ILP; ++i)
        {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            rows[i] = (s >> 20) % n_rows;
            double const* p = buf + rows[i] * (ROWB / 8) + lane;
#pragma unroll
            for(int d = 0; d < DPL; ++d) { v[i][d] = p[32 * d]; }
        }
#pragma unroll
        for(int i = 0; i < ILP; ++i)
        {
#pragma unroll
            for(int d = 0; d < DPL; ++d) { acc += v[i][d]; }
            if(WRITE && (i & 1))
            {
                double* p = buf + rows[i] * (ROWB / 8) + lane;
#pragma unroll
                for(int d = 0; d < DPL; ++d) { p[32 * d] = acc; }
            }
        }
    }
    if(acc == 123.456) { sink[0] = acc; }
}

template <int ROWB, int ILP, bool WRITE>
void run(double* buf, uint64_t bytes, double* sink, int blocks, int threads)
{
    uint64_t const n_rows = bytes / ROWB;
    int const iters = 400;
    k<ROWB, ILP, WRITE><<<blocks, threads>>>(buf, n_rows, 20, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ROWB, ILP, WRITE><<<blocks, threads>>>(buf, n_rows, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double const warps = (double)blocks * threads / 32;
    double const rd = warps * iters * ILP * ROWB, wr = WRITE ? rd / 2 : 0;
    printf("row %4d B  ilp %2d  %s  warps/SM %4.1f : %7.1f GB/s (read %.0f + write %.0f)\n", ROWB, ILP, WRITE ? "r+w" : "r  ", warps / 148, (rd + wr) / ms / 1e6, rd / ms / 1e6,
           wr / ms / 1e6);
}

int main()
{
    uint64_t const bytes = 1536ull << 20;
    double *buf, *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, bytes);
    for(int wps: {16, 32, 64})
    {
        int const threads = 1024, blocks = 148 * wps / 32;
        run<256, 4, false>(buf, bytes, sink, blocks, threads);
        run<512, 4, false>(buf, bytes, sink, blocks, threads);
        run<512, 8, false>(buf, bytes, sink, blocks, threads);
        run<1024, 4, false>(buf, bytes, sink, blocks, threads);
        run<512, 4, true>(buf, bytes, sink, blocks, threads);
        run<512, 8, true>(buf, bytes, sink, blocks, threads);
    }
    return 0;
}
