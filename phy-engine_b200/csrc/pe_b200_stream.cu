// pe_b200_stream.cu — the STREAM kernel (sm_100a): one warp per lane group of PE_SGL x PE_SJ lanes, one word stream per group.
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
#include <stdio.h>

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
    constexpr uint32_t SGL = PE_SGL;      // lanes of the warp that carry a lane of the group (the others mirror them, stores off)
    constexpr uint32_t GL = SGL * J;      // lanes per group
    static_assert(SGL == 32u || J == 1, "narrow groups carry one lane per thread");
    constexpr int NB = J == 1 ? 32 : 8;  // table entries whose values are in flight together
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t const NS = 1u << ns_log;
    uint32_t const stage_bytes = PE_STREAM_STAGE_ROWS * ROWB;
    uint32_t const ring_bytes = NS * stage_bytes;
    unsigned char* const ring = smem + (size_t)warp * (ring_bytes + 128u);  // the NS mbarriers follow the ring
    sk_ctx k;
    uint32_t const gl_lane = lane & (SGL - 1u);
    bool const active = lane < SGL;
    k.ring = reinterpret_cast<char const*>(ring) + gl_lane * 8u;
    k.ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    k.bars = k.ring_s + ring_bytes;
    k.stage_bytes = stage_bytes;
    k.ns_mask = NS - 1u;
    k.ns_log = ns_log;
    k.seq0 = 0u;
    k.pn = 0u;
    k.fenced = -1;
    k.lane = lane;
    k.n_rows = (uint32_t)r.n_slots;
    k.guard = r.guard;
    k.t = r.t0;
    if(lane == 0u)
    {
        for(uint32_t s = 0; s < NS; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k.bars + 8u * s), "r"(1u) : "memory"); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    uint32_t const NG = (uint32_t)((r.n_lanes + GL - 1) / GL);
    tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};
    for(uint32_t group = blockIdx.x + gridDim.x * warp; group < NG; group += gridDim.x * n_warps)
    {
        int64_t const glane = (int64_t)group * GL + gl_lane;  // first of this thread's lanes; the others are + SGL j
        k.gbase = reinterpret_cast<char*>(r.wsg + (int64_t)group * r.n_slots * GL);
        k.wl = k.gbase + gl_lane * 8u;
        auto at = [&](uint32_t slot) -> double* { return reinterpret_cast<double*>(k.wl + (size_t)slot * ROWB); };

        bool real_lane[J], counted[J], ok[J];
        int32_t status[J];
        uint32_t solves[J];
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            real_lane[j] = active && glane + SGL * j < r.n_lanes;
            status[j] = real_lane[j] ? r.status[glane + SGL * j] : (int32_t)PE_ST_SINGULAR;
            counted[j] = real_lane[j] && status[j] == PE_ST_OK;
            ok[j] = counted[j];
            solves[j] = 0;
        }
        // load table: persistent values -> workspace rows.  Lane l fetches entry e0 + l of the table, the warp then walks the 32
        // entries in batches whose value loads are all issued before the first store (one warp per group: nothing else hides
        // the latency of a dependent table-entry -> value -> store chain).
        for(uint32_t e0 = 0; e0 < (uint32_t)r.n_io; e0 += 32u)
        {
            pe_b200_io mine{0u, 0u};
            if(e0 + lane < (uint32_t)r.n_io) { mine = r.io[e0 + lane]; }
#pragma unroll 1
            for(uint32_t u0 = 0; u0 < 32u && e0 + u0 < (uint32_t)r.n_io; u0 += NB)
            {
                double v[NB][J];
                uint32_t sk[NB];
#pragma unroll
                for(int u = 0; u < NB; ++u)
                {
                    sk[u] = __shfl_sync(0xffffffffu, mine.slot_kind, (int)(u0 + u));
                    uint32_t const src = __shfl_sync(0xffffffffu, mine.src, (int)(u0 + u));
                    uint32_t const kind = (sk[u] >> 16) & 0xfu;
                    bool const on = ((sk[u] >> 20) & PE_IO_LOAD) != 0u;
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        // a padding lane (beyond n_lanes) computes on the values of the last real lane: zeros would send its
                        // warp through the slow paths of the FP64 division / reciprocal at every node, and the launch ends
                        // with its slowest warp
                        int64_t const ln = glane + SGL * j < r.n_lanes ? glane + SGL * j : (int64_t)r.n_lanes - 1;  // idle lanes of a narrow group mirror their twin
                        v[u][j] = 0.0;
                        if(!on) { continue; }
                        if(kind == PE_IO_CONST) { v[u][j] = __ldg(r.cst + src); }
                        else
                        {
                            v[u][j] = kind == PE_IO_U ? r.wu[(int64_t)src * r.LSu + ln] : r.wx[(int64_t)src * r.LSx + ln / r.ppi];
                        }
                    }
                }
#pragma unroll
                for(int u = 0; u < NB; ++u)
                {
                    if(!((sk[u] >> 20) & PE_IO_LOAD)) { continue; }
                    double* const dst = at(sk[u] & 0xffffu);
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        if(active) { dst[SGL * j] = v[u][j]; }
                    }
                }
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
        c.js = SGL;
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
        if(r.has_prep)
        {
#ifdef PE_STREAM_PREP
            // the prep section as generated tiles (derived per-instance values, replica rows)
            k.enm = 0u;
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(ok[j]) { k.enm |= 1u << j; }
            }
            uint32_t fm0 = 0u;
            pe_stream_prep(k, fm0);
            __syncwarp();
#else
            run_section(0, t);
#endif
        }
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
            k.t = t;
#ifdef PE_STREAM_STEADY
            if(s == 0) { pe_stream_iter(k, fm); }
            else
            {
                pe_stream_iters(k, fm);  // every solve but the first of a launch: state that cannot have changed is not re-read
            }
#else
            pe_stream_iter(k, fm);
#endif
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
                    for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + glane + SGL * j] = at(__ldg(r.probes + p))[SGL * j]; }
                }
            }
        }
        // store table: mutable values -> persistent rows (batched like the load table)
        for(uint32_t e0 = 0; e0 < (uint32_t)r.n_io; e0 += 32u)
        {
            pe_b200_io mine{0u, 0u};
            if(e0 + lane < (uint32_t)r.n_io) { mine = r.io[e0 + lane]; }
#pragma unroll 1
            for(uint32_t u0 = 0; u0 < 32u && e0 + u0 < (uint32_t)r.n_io; u0 += NB)
            {
                double v[NB][J];
                uint32_t sk[NB], sr[NB];
#pragma unroll
                for(int u = 0; u < NB; ++u)
                {
                    sk[u] = __shfl_sync(0xffffffffu, mine.slot_kind, (int)(u0 + u));
                    sr[u] = __shfl_sync(0xffffffffu, mine.src, (int)(u0 + u));
                    bool const on = ((sk[u] >> 20) & PE_IO_STORE) != 0u;
                    double const* const src = at(sk[u] & 0xffffu);
#pragma unroll
                    for(int j = 0; j < J; ++j) { v[u][j] = on ? src[SGL * j] : 0.0; }
                }
#pragma unroll
                for(int u = 0; u < NB; ++u)
                {
                    if(!((sk[u] >> 20) & PE_IO_STORE)) { continue; }
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        if(counted[j]) { r.wu[(int64_t)sr[u] * r.LSu + glane + SGL * j] = v[u][j]; }
                    }
                }
            }
        }
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            if(counted[j])
            {
                r.status[glane + SGL * j] = status[j];
                r.solves[glane + SGL * j] += solves[j];
            }
        }
        __syncwarp();
    }
}
