// pe_b200_stream.cu — the STREAM kernel (sm_100a): one warp per lane group of 32 J lanes, one word stream per group.
//
// Compiled at run time by nvcc (pe_b200_stream_build in pe_b200_kernels.cu, driven by host/stream.cpp) together with the
// generated source of one program's iter section (PE_STREAM_SOURCE: tiles, ring stage layout, bulk-copy schedule).
// Replaces, per lane, solve_once (circuit.h:987-1527) inside the TR loop of circult::analyze (circuit.h:233-289) for
// linear circuits.  Everything outside the iter section (load / store tables, prep and step sections, status, waveform
// probes) mirrors tree_body of pe_b200_kernels.cu with one warp per group and no barrier: a lane's whole solve lives in
// one thread, so there are no cross-warp flags either.
//
// Memory path: the cold operand rows of a tile are fetched by cp.async.bulk (TMA bulk copies, one elected lane) into this
// warp's ring in shared memory, NS tiles ahead of their use, completion through one mbarrier per ring stage; results are
// stored with plain 256-byte-per-warp stores.  See csrc/pe_b200_stream.h for the ordering rules.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pe_b200_program.h"
#include "pe_b200_models.h"
#include "pe_b200_rinterp.h"
#include "pe_b200_stream.h"

#ifndef PE_STREAM_SOURCE
#error "PE_STREAM_SOURCE names the generated iter section"
#endif

namespace
{
    using namespace pe_stream;
#include PE_STREAM_SOURCE
}  // namespace

extern "C" __global__ void __launch_bounds__(256, 1) pe_b200_stream_kernel(pe_b200_rrun const r, uint32_t const ns_log)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using namespace pe_rinterp;
    constexpr int J = PE_SJ;
    constexpr uint32_t GL = 32u * J;
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t const NS = 1u << ns_log;
    uint32_t const stage_bytes = PE_STREAM_STAGE_ROWS * ROWB;
    uint32_t const ring_bytes = NS * stage_bytes;
    unsigned char* const ring = smem + (size_t)warp * (ring_bytes + 128u);  // the NS mbarriers follow the ring
    sk_ctx k;
    k.ring = reinterpret_cast<char const*>(ring) + lane * 8u;
    k.ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    k.bars = k.ring_s + ring_bytes;
    k.stage_bytes = stage_bytes;
    k.ns_mask = NS - 1u;
    k.ns_log = ns_log;
    k.seq0 = 0u;
    k.pn = 0u;
    k.fenced = -1;
    k.lane = lane;
    if(lane == 0u)
    {
        for(uint32_t s = 0; s < NS; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k.bars + 8u * s), "r"(1u) : "memory"); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    uint32_t const NG = (uint32_t)((r.n_lanes + GL - 1) / GL);
    tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol};
    for(uint32_t group = blockIdx.x + gridDim.x * warp; group < NG; group += gridDim.x * n_warps)
    {
        int64_t const glane = (int64_t)group * GL + lane;  // first of this thread's lanes; the others are + 32 j
        k.gbase = reinterpret_cast<char*>(r.wsg + (int64_t)group * r.n_slots * GL);
        k.wl = k.gbase + lane * 8u;
        auto at = [&](uint32_t slot) -> double* { return reinterpret_cast<double*>(k.wl + (size_t)slot * ROWB); };

        bool real_lane[J], counted[J], ok[J];
        int32_t status[J];
        uint32_t solves[J];
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            real_lane[j] = glane + 32 * j < r.n_lanes;
            status[j] = real_lane[j] ? r.status[glane + 32 * j] : (int32_t)PE_ST_SINGULAR;
            counted[j] = real_lane[j] && status[j] == PE_ST_OK;
            ok[j] = counted[j];
            solves[j] = 0;
        }
        // load table: persistent values -> workspace rows
        for(uint32_t e = 0; e < (uint32_t)r.n_io; ++e)
        {
            pe_b200_io const io = r.io[e];
            if(!((io.slot_kind >> 20) & PE_IO_LOAD)) { continue; }
            uint32_t const kind = (io.slot_kind >> 16) & 0xfu;
            double* const dst = at(io.slot_kind & 0xffffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                // a padding lane (beyond n_lanes) computes on the values of the last real lane: zeros would send its warp
                // through the slow paths of the FP64 division / reciprocal at every node, and the launch ends with its slowest warp
                int64_t const ln = real_lane[j] ? glane + 32 * j : (int64_t)r.n_lanes - 1;
                double v;
                if(kind == PE_IO_CONST) { v = r.cst[io.src]; }
                else
                {
                    v = kind == PE_IO_U ? r.wu[(int64_t)io.src * r.LSu + ln] : r.wx[(int64_t)io.src * r.LSx + ln / r.ppi];
                }
                dst[32 * j] = v;
            }
        }
        __syncwarp();

        rctx c;
        c.ws = reinterpret_cast<double*>(k.wl);
        c.I = (uint64_t)GL;
        c.S = 1u;
        c.C = 1u;
        c.col = 0u;
        c.stream = 0u;
        c.js = 32u;
        // an interpreted section (prep, step): the generic vector-op executor, words read straight from global memory
        auto run_section = [&](int sec, double t)
        {
            host_reader rd;
            rd.p0 = r.words + __ldg(r.sec_off + sec);      // one warp stream per section (S = 1)
            rd.q = r.words + __ldg(r.sec_off + 3 + sec);   // its side stream
            rd.C = 1u;
            rd.col = 0u;
            rd.start();
            bool nc[J], fl[J];
#pragma unroll
            for(int j = 0; j < J; ++j) { nc[j] = fl[j] = false; }
            for(;;)
            {
                int const kind = rvop<J, host_reader, false>(rd, c, t, tol, ok, false, nc, fl, true);
                if(kind == V_END || kind == V_BAD) { break; }
                if(kind == V_BAR) { rd.bar(); }
                else if(kind == V_SKIP) { rd.skip(); }
                else
                {
                    rd.close();
                }
            }
            __syncwarp();
        };

        double t = r.t0;
        if(r.has_prep) { run_section(0, t); }
        for(int32_t s = 0; s < r.n_steps; ++s)
        {
            if(r.time_stepping)
            {
                // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                if(r.has_step) { run_section(1, t); }
                t = t + r.dt;
            }
            k.enm = 0u;
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(ok[j]) { k.enm |= 1u << j; }
            }
            // one linear solve (stamp + LU + substitution): the generated tiles
            uint32_t fm = 0u;
            pe_stream_iter(k, fm);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(!ok[j]) { continue; }
                ++solves[j];
                if((fm >> j) & 1u)
                {
                    status[j] = PE_ST_SINGULAR;
                    ok[j] = false;
                }
            }
            if(r.wave != nullptr)
            {
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(!ok[j]) { continue; }
                    for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + glane + 32 * j] = at(__ldg(r.probes + p))[32 * j]; }
                }
            }
        }
        // store table: mutable values -> persistent rows
        for(uint32_t e = 0; e < (uint32_t)r.n_io; ++e)
        {
            pe_b200_io const io = r.io[e];
            if(!((io.slot_kind >> 20) & PE_IO_STORE)) { continue; }
            double const* const src = at(io.slot_kind & 0xffffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(counted[j]) { r.wu[(int64_t)io.src * r.LSu + glane + 32 * j] = src[32 * j]; }
            }
        }
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            if(counted[j])
            {
                r.status[glane + 32 * j] = status[j];
                r.solves[glane + 32 * j] += solves[j];
            }
        }
        __syncwarp();
    }
}
