// Micro-benchmark for the stream kernel's memory pattern (DESIGN.md §5 "stream kernel"): W warps per SM, each warp owns
// one lane group and walks its node records (8 rows of ROWB bytes per node) forward then backward, like one time step
// of a chain circuit at one stream per lane group:
//   forward  node i: reads rows {0,1,2,3} and {5} of record i (TMA bulk copies into a per-warp shared-memory ring,
//                    mbarrier complete_tx), writes rows {0,1,4,7} (plain stores)
//   backward node i: reads rows {4,5} and {7}, writes row {3}
// with a dependent FP64 chain per node (one reciprocal + a few FMAs) so the per-warp serial latency is realistic.
// Prints the achieved DRAM-side bandwidth (bytes requested / time).  usage: tmastream [nodes] [steps]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(void const* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, void const* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// J doubles per thread and row; U nodes per tile; NS ring stages
template <int J, int U, int NS, int MODE>
__global__ void __launch_bounds__(256, 1) stream_k(double* ws, int n_groups, int nodes, int steps, int warps_per_cta, double* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t ROWB = 256u * J;
    constexpr uint32_t STAGE = U * 5u * ROWB;  // forward tile: 5 rows per node; backward uses 3
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int const group = (int)(blockIdx.x + gridDim.x * warp);
    if(group >= n_groups) { return; }
    unsigned char* const ring = smem + (size_t)warp * (NS * STAGE + 64);
    uint32_t const ring_s = smem_u32(ring);
    uint32_t const bars = ring_s + NS * STAGE;
    if(lane == 0)
    {
        for(int s = 0; s < NS; ++s) { mbar_init(bars + 8u * s, 1u); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    char* const base = reinterpret_cast<char*>(ws) + (size_t)group * nodes * 8u * ROWB;
    int const tiles = nodes / U;
    uint32_t seq = 0;  // running tile counter (stage = seq % NS, parity = (seq / NS) & 1)
    double carry[J];
#pragma unroll
    for(int j = 0; j < J; ++j) { carry[j] = 1.0; }
    for(int st = 0; st < steps; ++st)
    {
        for(int dir = 0; dir < 2; ++dir)
        {
            // all earlier stores of this warp are ordered before the bulk copies that follow
            fence_async();
            __syncwarp();
            auto issue = [&](int t, uint32_t sq)
            {
                uint32_t const stg = sq % NS, bar = bars + 8u * stg, dst = ring_s + stg * STAGE;
                if(MODE == 2)
                {
                    if(dir == 0)
                    {
                        mbar_expect_tx(bar, U * 5u * ROWB);
                        bulk_g2s(dst, base + (size_t)t * U * 5u * ROWB, U * 5u * ROWB, bar);
                    }
                    else
                    {
                        mbar_expect_tx(bar, U * 3u * ROWB);
                        bulk_g2s(dst, base + (size_t)nodes * 5u * ROWB + (size_t)(tiles - 1 - t) * U * 3u * ROWB, U * 3u * ROWB, bar);
                    }
                }
                else if(dir == 0)
                {
                    mbar_expect_tx(bar, U * 5u * ROWB);
#pragma unroll
                    for(int u = 0; u < U; ++u)
                    {
                        char const* rec = base + (size_t)(t * U + u) * 8u * ROWB;
                        bulk_g2s(dst + u * 5u * ROWB, rec, 4u * ROWB, bar);
                        bulk_g2s(dst + u * 5u * ROWB + 4u * ROWB, rec + 5u * ROWB, ROWB, bar);
                    }
                }
                else
                {
                    mbar_expect_tx(bar, U * 3u * ROWB);
#pragma unroll
                    for(int u = 0; u < U; ++u)
                    {
                        char const* rec = base + (size_t)((tiles - 1 - t) * U + (U - 1 - u)) * 8u * ROWB;
                        bulk_g2s(dst + u * 3u * ROWB, rec + 4u * ROWB, 2u * ROWB, bar);
                        bulk_g2s(dst + u * 3u * ROWB + 2u * ROWB, rec + 7u * ROWB, ROWB, bar);
                    }
                }
            };
            int pn = 0;
            if(lane == 0)
            {
                for(; pn < NS && pn < tiles; ++pn) { issue(pn, seq + pn); }
            }
            pn = NS < tiles ? NS : tiles;
            for(int t = 0; t < tiles; ++t)
            {
                uint32_t const sq = seq + t, stg = sq % NS;
                mbar_wait(bars + 8u * stg, (sq / NS) & 1u);
                double const* in = reinterpret_cast<double const*>(ring + stg * STAGE);
                if(dir == 0)
                {
                    double v[U][5][J];
#pragma unroll
                    for(int u = 0; u < U; ++u)
#pragma unroll
                        for(int r = 0; r < 5; ++r)
#pragma unroll
                            for(int j = 0; j < J; ++j) { v[u][r][j] = in[(u * 5 + r) * 32 * J + 32 * j + lane]; }
                    __syncwarp();
                    if(lane == 0 && pn < tiles) { issue(pn, seq + pn); }
                    ++pn;
#pragma unroll
                    for(int u = 0; u < U; ++u)
                    {
                        double* rec = reinterpret_cast<double*>(base + (size_t)(t * U + u) * 8u * ROWB) + lane;
#pragma unroll
                        for(int j = 0; j < J; ++j)
                        {
                            double const g_new = MODE == 1 ? __dmul_rn(2.0, v[u][2][j]) : __ddiv_rn(__dmul_rn(2.0, v[u][2][j]), 1e-8 + 1.0);
                            double const hist = __dsub_rn(__dmul_rn(-__dadd_rn(g_new, v[u][1][j]), v[u][3][j]), v[u][0][j]);
                            double piv = __dadd_rn(__dadd_rn(v[u][4][j], g_new), 1.0);
                            piv = __fma_rn(carry[j], v[u][4][j], piv);
                            double const rc = MODE == 1 ? __dmul_rn(piv, 0.3) : __drcp_rn(piv);
                            carry[j] = __dmul_rn(-v[u][4][j], rc);
                            rec[0 * 32 * J + 32 * j] = __fma_rn(hist, 1e-30, 0.5);
                            rec[1 * 32 * J + 32 * j] = __fma_rn(g_new, 1e-30, 0.5);
                            rec[4 * 32 * J + 32 * j] = __fma_rn(rc, 1e-30, 0.5);
                            rec[7 * 32 * J + 32 * j] = __fma_rn(__fma_rn(carry[j], hist, 1.0), 1e-30, 0.5);
                        }
                    }
                }
                else
                {
                    double v[U][3][J];
#pragma unroll
                    for(int u = 0; u < U; ++u)
#pragma unroll
                        for(int r = 0; r < 3; ++r)
#pragma unroll
                            for(int j = 0; j < J; ++j) { v[u][r][j] = in[(u * 3 + r) * 32 * J + 32 * j + lane]; }
                    __syncwarp();
                    if(lane == 0 && pn < tiles) { issue(pn, seq + pn); }
                    ++pn;
#pragma unroll
                    for(int u = 0; u < U; ++u)
                    {
                        double* rec = reinterpret_cast<double*>(base + (size_t)((tiles - 1 - t) * U + (U - 1 - u)) * 8u * ROWB) + lane;
#pragma unroll
                        for(int j = 0; j < J; ++j)
                        {
                            carry[j] = __dmul_rn(__fma_rn(v[u][1][j], carry[j], v[u][2][j]), v[u][0][j]) * 0.25;
                            rec[3 * 32 * J + 32 * j] = __fma_rn(carry[j], 1e-30, 0.5);
                        }
                    }
                }
            }
            seq += (uint32_t)tiles;
        }
    }
    if(carry[0] == 123.456) { sink[0] = carry[0]; }
}

__global__ void fill_k(double* p, size_t n)
{
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { p[i] = 0.5 + 1e-3 * (double)(i % 977); }
}

template <int J, int U, int NS, int MODE = 0>
void run(double* ws, size_t bytes, double* sink, int lanes, int nodes, int steps)
{
    int const n_groups = (lanes + 32 * J - 1) / (32 * J);
    int const W = (n_groups + 147) / 148;
    size_t const smem = (size_t)W * (NS * U * 5 * 256 * J + 64);
    if((size_t)n_groups * nodes * 8 * 256 * J > bytes || smem > 227 * 1024 || W > 8)
    {
        printf("J %d U %d NS %2d: skipped (W %d smem %zu)\n", J, U, NS, W, smem);
        return;
    }
    cudaFuncSetAttribute(stream_k<J, U, NS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fill_k<<<1184, 256>>>(ws, bytes / 8);
    stream_k<J, U, NS, MODE><<<148, 32 * W, smem>>>(ws, n_groups, nodes, 2, W, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    stream_k<J, U, NS, MODE><<<148, 32 * W, smem>>>(ws, n_groups, nodes, steps, W, sink);
    cudaEventRecord(e1);
    cudaError_t const err = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double const per_node = 13.0 * 256 * J;  // 8 rows read + 5 written... (5 + 3 read, 4 + 1 written)
    double const total = (double)n_groups * nodes * steps * per_node;
    printf("mode %d J %d U %d NS %2d  groups %4d (W %d, ring %3zu KB/SM): %8.3f ms  %7.1f GB/s  %6.2f M lane-steps/s  [%s]\n", MODE, J, U, NS, n_groups, W, smem / 1024, ms, total / ms / 1e6,
           (double)n_groups * 32 * J * steps / ms / 1e3, cudaGetErrorString(err));
    fflush(stdout);
}

int main(int argc, char** argv)
{
    int const nodes = argc > 1 ? atoi(argv[1]) : 1000, steps = argc > 2 ? atoi(argv[2]) : 20, lanes = argc > 3 ? atoi(argv[3]) : 10000;
    size_t const bytes = (size_t)((lanes + 127) / 128) * 128 * nodes * 8 * 8 + (64u << 20);
    double *ws, *sink;
    cudaMalloc(&ws, bytes);
    cudaMalloc(&sink, 8);
    fill_k<<<1184, 256>>>(ws, bytes / 8);
    printf("fill: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    fflush(stdout);
    printf("nodes %d steps %d lanes %d, workspace %.1f MB\n", nodes, steps, lanes, bytes / 1e6);
    fflush(stdout);
    run<1, 4, 8>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 8, 4>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 8, 6>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 4, 8, 1>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 8, 4, 1>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 4, 8, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 8, 4, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<1, 8, 6, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 4, 4>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 4, 8>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 4, 8, 1>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 4, 8, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 5, 8, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<2, 8, 4, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<4, 2, 8>(ws, bytes, sink, lanes, nodes, steps);
    run<4, 2, 8, 2>(ws, bytes, sink, lanes, nodes, steps);
    run<4, 4, 8, 2>(ws, bytes, sink, lanes, nodes, steps);
    return 0;
}
