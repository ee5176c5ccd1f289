// pe_b200_kernels.cu — the sm_100a solve kernel of the batched MNA hot path + the POD device seam.
//
// CTA = 32 lanes x G warps.  Thread (g, l) interprets warp stream g of the batch program on lane l; its data accesses
// hit lane-interleaved HBM arrays w[slot][lane] (32 consecutive doubles per warp request) and program words are
// warp-uniform (no divergence on opcodes).  The G warps of a CTA work on disjoint sub-trees of the elimination tree
// of the SAME 32 circuits and meet at CTA barriers (PE_OP_BAR) around the separator rows, which is what gives a
// 10k-instance batch enough resident warps to hide HBM latency.  One launch runs a whole analysis
// phase per lane: [prep] -> for each time step { [step]; Newton loop { eval + assemble + LU + substitution with the
// convergence test fused into the back substitution } }.  Replaces, per lane, the reference's circult::solve /
// solve_once / update_tr_step (circuit.h:363-374, 892-1527) and Eigen::SparseLU compute + solve.
//
// This TU includes no reference header (nvcc ICEs on fast_io; SURVEY.md probe table).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <utility>
#include <vector>

#include "pe_b200_interp.h"

namespace
{
    using namespace pe_interp;

    __global__ void __launch_bounds__(32 * PE_MAX_WARPS, 2) pe_b200_solve_kernel(pe_b200_run const r)
    {
        // per-lane Newton flags of the current iteration, OR-ed in by the G warps (bit 0 not converged, bit 1 singular);
        // three buffers rotate so that the reset of the next one never races with readers of the previous one
        __shared__ uint32_t s_flags[3][32];

        int const g = (int)(threadIdx.x >> 5);
        int const li = (int)(threadIdx.x & 31u);
        int64_t const lane = (int64_t)blockIdx.x * 32 + li;  // < LSu (padding lanes run on scratch rows, never store)
        bool const real_lane = lane < r.n_lanes;
        int32_t status = real_lane ? r.status[lane] : (int32_t)PE_ST_SINGULAR;
        bool const counted = real_lane && status == PE_ST_OK;

        ctx_t c;
        c.cst = r.cst;
        c.wu = r.wu + lane;
        c.wx = r.wx + (real_lane ? lane / r.ppi : 0);
        c.LSu = r.LSu;
        c.LSx = r.LSx;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol};

        if(threadIdx.x < 96) { (&s_flags[0][0])[threadIdx.x] = 0u; }
        __syncthreads();

        double t = r.t0;
        auto run_section = [&](pe_b200_section const& sec, bool live, bool check, bool& nconv, bool& fail)
        {
            if(sec.off[0] == PE_NO_SECTION) { return; }
            uint32_t const* pc = r.words + sec.off[g];
            while(run_until(pc, c, t, tol, live, check, nconv, fail) == R_BAR) { __syncthreads(); }
            __syncthreads();  // results of this section are visible to every warp of the CTA
        };

        bool ok = counted;
        uint32_t solves = 0;
        int fi = 0;
        {
            bool a = false, b = false;
            run_section(r.prep, ok, false, a, b);
        }
        for(int32_t s = 0; s < r.n_steps; ++s)
        {
            if(r.time_stepping)
            {
                // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                bool a = false, b = false;
                run_section(r.step, ok, false, a, b);
                t = t + r.dt;
            }
            bool done = !ok;
            int32_t it = 0;
            for(;;)
            {
                int const fn = fi == 2 ? 0 : fi + 1;
                if(g == 0) { s_flags[fn][li] = 0u; }
                bool nconv = false, fail = false;
                run_section(r.iter, !done, r.nonlinear != 0, nconv, fail);
                if(nconv || fail) { atomicOr(&s_flags[fi][li], (nconv ? 1u : 0u) | (fail ? 2u : 0u)); }
                __syncthreads();
                uint32_t const f = s_flags[fi][li];
                fi = fn;
                if(!done)
                {
                    ++solves;
                    if(f & 2u)
                    {
                        status = PE_ST_SINGULAR;
                        ok = false;
                        done = true;
                    }
                    else if(!r.nonlinear || !(f & 1u)) { done = true; }
                    else if(++it >= r.max_iter)
                    {
                        status = PE_ST_NO_CONVERGENCE;
                        ok = false;
                        done = true;
                    }
                }
                if(__syncthreads_and(done ? 1 : 0)) { break; }
            }
            if(r.wave != nullptr && ok && g == 0)
            {
                for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane] = ld(c, __ldg(r.probes + p)); }
            }
        }
        if(g == 0 && counted)
        {
            r.status[lane] = status;
            r.solves[lane] += solves;
        }
    }

    thread_local char g_err[256] = "";
    std::atomic<uint64_t> g_launches{0};

    // optional per-launch device timing (CUDA events on the launching stream), used by bench.py for the roofline line
    bool g_timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;

    int chk(cudaError_t e, char const* what)
    {
        if(e == cudaSuccess) { return 0; }
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
}  // namespace

extern "C"
{
    int pe_b200_dev_count(void)
    {
        int n = 0;
        if(cudaGetDeviceCount(&n) != cudaSuccess)
        {
            (void)cudaGetLastError();
            return 0;
        }
        return n;
    }

    int pe_b200_dev_set(int device) { return chk(cudaSetDevice(device), "cudaSetDevice"); }

    int pe_b200_dev_malloc(void** p, size_t bytes) { return chk(cudaMalloc(p, bytes ? bytes : 8), "cudaMalloc"); }

    int pe_b200_dev_free(void* p) { return p ? chk(cudaFree(p), "cudaFree") : 0; }

    int pe_b200_dev_memset0(void* p, size_t bytes, void* stream) { return chk(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream), "cudaMemsetAsync"); }

    int pe_b200_dev_h2d(void* dst, void const* src, size_t bytes, void* stream)
    {
        return chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync(H2D)");
    }

    int pe_b200_dev_d2h(void* dst, void const* src, size_t bytes, void* stream)
    {
        return chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(D2H)");
    }

    int pe_b200_dev_h2d_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream)
    {
        return chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpy2DAsync(H2D)");
    }

    int pe_b200_dev_d2h_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream)
    {
        return chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpy2DAsync(D2H)");
    }

    int pe_b200_dev_sync(void* stream) { return chk(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

    int pe_b200_launch(pe_b200_run const* run, void* stream)
    {
        if(run == nullptr || run->n_lanes <= 0) { return 0; }
        if(run->warps < 1 || run->warps > PE_MAX_WARPS)
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch: warps per CTA must be 1..%d", PE_MAX_WARPS);
            return 1;
        }
        int const block = 32 * run->warps;
        int const grid = (run->n_lanes + 31) / 32;
        cudaEvent_t e0{}, e1{};
        if(g_timing)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)stream);
        }
        pe_b200_solve_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*run);
        if(g_timing)
        {
            cudaEventRecord(e1, (cudaStream_t)stream);
            g_events.emplace_back(e0, e1);
        }
        g_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_solve_kernel launch");
    }

    void pe_b200_timing_enable(int on) { g_timing = on != 0; }

    // total device time (ms) of the solve-kernel launches since the last collect; waits for them to finish
    double pe_b200_timing_collect(void)
    {
        double total = 0.0;
        for(auto& [a, b]: g_events)
        {
            float ms = 0.f;
            cudaEventSynchronize(b);
            if(cudaEventElapsedTime(&ms, a, b) == cudaSuccess) { total += ms; }
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
        g_events.clear();
        return total;
    }

    char const* pe_b200_dev_last_error(void) { return g_err; }

    uint64_t pe_b200_launch_count(void) { return g_launches.load(); }
}
