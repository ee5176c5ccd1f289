// Comparison point, not product code: what the reference's existing GPU path costs per solve on this GPU.
//
// The reference's CUDA solver (include/phy_engine/circuits/solver/cuda_sparse_lu.h:617-633) hands ONE system per call to
// cuSOLVER's sparse QR, cusolverSpDcsrqrsvBatched(..., batchSize = 1, ...), after a symbolic analysis that it caches per
// sparsity pattern.  This program re-issues exactly that call sequence on the MNA matrix of config B (RC ladder, 1000
// sections: 1002 unknowns, trapezoidal companions at dt = 1e-8) and times it with CUDA events, at batch 1 (the reference's
// call) and at larger batches of the same pattern (what the library could do if the reference batched its instances).
//
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o cusolver_point cusolver_point.cu -lcusolver -lcusparse
//   ./cusolver_point [sections] [max batch]
#include <cuda_runtime.h>
#include <cusolverSp.h>
#include <cusparse.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        auto const rc_ = (x);                                                                   \
        if(rc_ != 0)                                                                            \
        {                                                                                       \
            std::fprintf(stderr, "%s failed: %d (line %d)\n", #x, (int)rc_, __LINE__);          \
            return 1;                                                                           \
        }                                                                                       \
    } while(0)

int main(int argc, char** argv)
{
    int const sections = argc > 1 ? std::atoi(argv[1]) : 1000;
    int const max_batch = argc > 2 ? std::atoi(argv[2]) : 1024;
    // unknowns: nodes 0..sections (node 0 = the source's node), one branch current
    int const n_nodes = sections + 1, n = n_nodes + 1, br = n_nodes;
    double const g = 1e-3, gc = 2.0 * 1e-9 / 1e-8;
    std::vector<std::map<int, double>> rows(n);
    for(int k = 0; k < sections; ++k)
    {
        rows[k][k] += g;
        rows[k + 1][k + 1] += g;
        rows[k][k + 1] -= g;
        rows[k + 1][k] -= g;
        rows[k + 1][k + 1] += gc;  // capacitor companion to ground
    }
    rows[0][br] += 1.0;  // VDC branch
    rows[br][0] += 1.0;
    std::vector<int> rp(n + 1, 0), ci;
    std::vector<double> va;
    for(int i = 0; i < n; ++i)
    {
        for(auto const& [j, v]: rows[i])
        {
            ci.push_back(j);
            va.push_back(v);
        }
        rp[i + 1] = (int)ci.size();
    }
    int const nnz = (int)ci.size();
    std::vector<double> b(n, 0.0);
    b[br] = 1.0;
    for(int k = 1; k < n_nodes; ++k) { b[k] = 1e-4 * (1 + k % 7); }  // companion history currents

    cusolverSpHandle_t h;
    cusparseMatDescr_t descr;
    csrqrInfo_t info;
    CK(cusolverSpCreate(&h));
    CK(cusparseCreateMatDescr(&descr));
    cusparseSetMatType(descr, CUSPARSE_MATRIX_TYPE_GENERAL);
    cusparseSetMatIndexBase(descr, CUSPARSE_INDEX_BASE_ZERO);
    CK(cusolverSpCreateCsrqrInfo(&info));
    int *d_rp, *d_ci;
    CK(cudaMalloc(&d_rp, (n + 1) * sizeof(int)));
    CK(cudaMalloc(&d_ci, nnz * sizeof(int)));
    CK(cudaMemcpy(d_rp, rp.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ci, ci.data(), nnz * sizeof(int), cudaMemcpyHostToDevice));

    std::printf("{\"tool\": \"cusolverSpDcsrqrsvBatched\", \"unknowns\": %d, \"nnz\": %d, \"points\": [", n, nnz);
    bool first = true;
    for(int B = 1; B <= max_batch; B *= 4)
    {
        std::vector<double> vb((size_t)B * nnz), bb((size_t)B * n);
        for(int q = 0; q < B; ++q)
        {
            for(int e = 0; e < nnz; ++e) { vb[(size_t)q * nnz + e] = va[e] * (1.0 + 0.01 * ((q * 31 + e) % 17) * (std::fabs(va[e]) == 1.0 ? 0.0 : 1.0)); }
            for(int i = 0; i < n; ++i) { bb[(size_t)q * n + i] = b[i]; }
        }
        // the symbolic analysis is cached per sparsity pattern by the reference (prepare_csrqr_real); its buffers are sized per batch
        if(B > 1)
        {
            CK(cusolverSpDestroyCsrqrInfo(info));
            CK(cusolverSpCreateCsrqrInfo(&info));
        }
        CK(cusolverSpXcsrqrAnalysisBatched(h, n, n, nnz, descr, d_rp, d_ci, info));
        double *d_v, *d_b, *d_x;
        void* d_buf;
        size_t internal = 0, workspace = 0;
        CK(cudaMalloc(&d_v, vb.size() * sizeof(double)));
        CK(cudaMalloc(&d_b, bb.size() * sizeof(double)));
        CK(cudaMalloc(&d_x, bb.size() * sizeof(double)));
        CK(cudaMemcpy(d_v, vb.data(), vb.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_b, bb.data(), bb.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cusolverSpDcsrqrBufferInfoBatched(h, n, n, nnz, descr, d_v, d_rp, d_ci, B, info, &internal, &workspace));
        CK(cudaMalloc(&d_buf, workspace));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        int const reps = B == 1 ? 20 : 5;
        CK(cusolverSpDcsrqrsvBatched(h, n, n, nnz, descr, d_v, d_rp, d_ci, d_b, d_x, B, info, d_buf));  // warm-up
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for(int r = 0; r < reps; ++r) { CK(cusolverSpDcsrqrsvBatched(h, n, n, nnz, descr, d_v, d_rp, d_ci, d_b, d_x, B, info, d_buf)); }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // residual of system 0
        std::vector<double> x(n);
        CK(cudaMemcpy(x.data(), d_x, n * sizeof(double), cudaMemcpyDeviceToHost));
        double res = 0.0, nb = 0.0;
        for(int i = 0; i < n; ++i)
        {
            double s = -b[i];
            for(int e = rp[i]; e < rp[i + 1]; ++e) { s += vb[e] * x[ci[e]]; }
            res = std::fmax(res, std::fabs(s));
            nb = std::fmax(nb, std::fabs(b[i]));
        }
        double const per = ms / reps / B;
        std::printf("%s{\"batch\": %d, \"ms_per_call\": %.4f, \"us_per_solve\": %.3f, \"solves_per_s\": %.1f, \"workspace_mb\": %.1f, \"residual\": %.2e}", first ? "" : ", ", B,
                    ms / reps, per * 1e3, 1e3 / per, (double)workspace / 1e6, res / nb);
        first = false;
        cudaFree(d_v);
        cudaFree(d_b);
        cudaFree(d_x);
        cudaFree(d_buf);
    }
    std::printf("]}\n");
    return 0;
}
