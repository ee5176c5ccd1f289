// TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or shipped with libphyengine_b200.so.
//
// tests/emu/libpe_emu.so = the product's host side (phy-engine_b200/host/*.cpp, unchanged) linked against THIS file
// instead of the CUDA translation unit.  It implements the POD device seam of pe_b200_program.h on host memory and
// replays a launch lane by lane with the same per-thread interpreter the sm_100a kernel is built from
// (csrc/pe_b200_interp.h).  Purpose: check the symbolic phase (ordering, fill, schedules, warp streams, barriers)
// against the compiled reference on a machine without a GPU, and detect data races between the warp streams of a
// CTA (two streams touching the same slot between two barriers).  The GPU parity tests (-m gpu) never load it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>
#include <vector>

namespace emu_trace
{
    struct rec
    {
        int w_phase{-1}, w_warp{-1};
        int r_phase{-1};
        uint32_t r_mask{};
    };

    static std::unordered_map<double const*, rec> g_map;
    static int g_phase = 0, g_warp = 0;
    static uint64_t g_races = 0;
    static bool g_on = true;
    static inline bool tracked(double const*) { return g_on; }

    static inline void on_load(double const* p)
    {
        if(!tracked(p)) { return; }
        auto& r = g_map[p];
        if(r.w_phase == g_phase && r.w_warp != g_warp) { ++g_races; }
        if(r.r_phase != g_phase)
        {
            r.r_phase = g_phase;
            r.r_mask = 0;
        }
        r.r_mask |= 1u << g_warp;
    }

    static inline void on_store(double const* p)
    {
        if(!tracked(p)) { return; }
        auto& r = g_map[p];
        if(r.w_phase == g_phase && r.w_warp != g_warp) { ++g_races; }
        if(r.r_phase == g_phase && (r.r_mask & ~(1u << g_warp)) != 0u) { ++g_races; }
        r.w_phase = g_phase;
        r.w_warp = g_warp;
    }
}  // namespace emu_trace

#define PE_TRACE_LD(ptr) emu_trace::on_load(ptr)
#define PE_TRACE_ST(ptr) emu_trace::on_store(ptr)
#include "../../phy-engine_b200/csrc/pe_b200_interp.h"

namespace
{
    using namespace pe_interp;
    char g_err[256] = "";
    uint64_t g_launches = 0;
    uint64_t g_unbalanced = 0;

    // one section for one lane: the G warp streams advance phase by phase (a phase ends at PE_OP_BAR / PE_OP_END)
    void run_section(pe_b200_run const& r, pe_b200_section const& sec, ctx_t const& c, double t, tol_t const& tol, bool live, bool check, bool& nconv, bool& fail)
    {
        if(sec.off[0] == PE_NO_SECTION) { return; }
        int const G = r.warps;
        std::vector<uint32_t const*> pc(static_cast<size_t>(G));
        for(int g = 0; g < G; ++g) { pc[static_cast<size_t>(g)] = r.words + sec.off[g]; }
        for(;;)
        {
            int n_bar = 0, n_end = 0;
            for(int g = 0; g < G; ++g)
            {
                emu_trace::g_warp = g;
                bool a = false, b = false;
                int const rc = run_until(pc[static_cast<size_t>(g)], c, t, tol, live, check, a, b);
                nconv = nconv || a;
                fail = fail || b;
                if(rc == R_BAR) { ++n_bar; }
                else
                {
                    ++n_end;
                }
            }
            ++emu_trace::g_phase;
            if(n_bar != 0 && n_end != 0)
            {
                ++g_unbalanced;  // on the GPU this would be a deadlock
                return;
            }
            if(n_end != 0) { return; }
        }
    }
}  // namespace

extern "C"
{
    int pe_b200_dev_count(void) { return 1; }
    int pe_b200_dev_set(int) { return 0; }
    int pe_b200_dev_malloc(void** p, size_t bytes)
    {
        *p = calloc(bytes ? bytes : 8, 1);
        return *p ? 0 : 1;
    }
    int pe_b200_dev_free(void* p)
    {
        free(p);
        return 0;
    }
    int pe_b200_dev_memset0(void* p, size_t bytes, void*)
    {
        memset(p, 0, bytes);
        return 0;
    }
    int pe_b200_dev_h2d(void* dst, void const* src, size_t bytes, void*)
    {
        memcpy(dst, src, bytes);
        return 0;
    }
    int pe_b200_dev_d2h(void* dst, void const* src, size_t bytes, void*)
    {
        memcpy(dst, src, bytes);
        return 0;
    }
    int pe_b200_dev_sync(void*) { return 0; }
    int pe_b200_dev_h2d_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void*)
    {
        for(size_t r = 0; r < height; ++r) { memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<char const*>(src) + r * spitch, width); }
        return 0;
    }
    int pe_b200_dev_d2h_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void*)
    {
        for(size_t r = 0; r < height; ++r) { memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<char const*>(src) + r * spitch, width); }
        return 0;
    }
    char const* pe_b200_dev_last_error(void) { return g_err; }
    uint64_t pe_b200_launch_count(void) { return g_launches; }
    void pe_b200_timing_enable(int) {}
    double pe_b200_timing_collect(void) { return 0.0; }

    uint64_t pe_emu_races(void) { return emu_trace::g_races; }
    uint64_t pe_emu_unbalanced_barriers(void) { return g_unbalanced; }
    void pe_emu_trace(int on) { emu_trace::g_on = on != 0; }

    // Same control flow as pe_b200_solve_kernel (csrc/pe_b200_kernels.cu), one lane at a time.
    int pe_b200_launch(pe_b200_run const* rp, void*)
    {
        if(rp == nullptr || rp->n_lanes <= 0) { return 0; }
        pe_b200_run const& r = *rp;
        ++g_launches;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol};
        for(int64_t lane = 0; lane < r.n_lanes; ++lane)
        {
            int32_t status = r.status[lane];
            bool const counted = status == PE_ST_OK;
            ctx_t c;
            c.cst = r.cst;
            c.wu = r.wu + lane;
            c.wx = r.wx + lane / r.ppi;
            c.LSu = r.LSu;
            c.LSx = r.LSx;
            emu_trace::g_map.clear();
            emu_trace::g_phase = 0;
            double t = r.t0;
            bool ok = counted;
            uint32_t solves = 0;
            {
                bool a = false, b = false;
                run_section(r, r.prep, c, t, tol, ok, false, a, b);
            }
            for(int32_t s = 0; s < r.n_steps; ++s)
            {
                if(r.time_stepping)
                {
                    bool a = false, b = false;
                    run_section(r, r.step, c, t, tol, ok, false, a, b);
                    t = t + r.dt;
                }
                bool done = !ok;
                int32_t it = 0;
                while(!done)
                {
                    bool nconv = false, fail = false;
                    run_section(r, r.iter, c, t, tol, true, r.nonlinear != 0, nconv, fail);
                    ++solves;
                    if(fail)
                    {
                        status = PE_ST_SINGULAR;
                        ok = false;
                        done = true;
                    }
                    else if(!r.nonlinear || !nconv) { done = true; }
                    else if(++it >= r.max_iter)
                    {
                        status = PE_ST_NO_CONVERGENCE;
                        ok = false;
                        done = true;
                    }
                }
                if(r.wave != nullptr && ok)
                {
                    for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane] = ld(c, r.probes[p]); }
                }
            }
            if(counted)
            {
                r.status[lane] = status;
                r.solves[lane] += solves;
            }
        }
        return 0;
    }
}
