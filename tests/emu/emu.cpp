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
#include <dlfcn.h>
#include <cmath>
#include <algorithm>
#include <array>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace emu_trace
{
    struct rec
    {
        int w_phase{-1}, w_warp{-1};
        int r_phase{-1};
        int r_first{-1};   // first stream that read the slot in r_phase
        bool r_many{};     // ... and some other stream read it too
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
            r.r_first = g_warp;
            r.r_many = false;
        }
        else if(r.r_first != g_warp) { r.r_many = true; }
    }

    static inline void on_store(double const* p)
    {
        if(!tracked(p)) { return; }
        auto& r = g_map[p];
        if(r.w_phase == g_phase && r.w_warp != g_warp) { ++g_races; }
        if(r.r_phase == g_phase && (r.r_many || r.r_first != g_warp)) { ++g_races; }
        r.w_phase = g_phase;
        r.w_warp = g_warp;
    }
}  // namespace emu_trace

#define PE_TRACE_LD(ptr) emu_trace::on_load(ptr)
#define PE_TRACE_ST(ptr) emu_trace::on_store(ptr)
#include "../../phy-engine_b200/csrc/pe_b200_interp.h"
#include "../../phy-engine_b200/csrc/pe_b200_rinterp.h"

namespace
{
    using namespace pe_interp;
    char g_err[256] = "";
    uint64_t g_launches = 0, g_aux_launches = 0;
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
    uint64_t pe_b200_aux_launch_count(void) { return g_aux_launches; }
    void pe_b200_timing_enable(int) {}
    double pe_b200_timing_collect(void) { return 0.0; }

    size_t pe_b200_resident_smem_limit(void) { return 227 * 1024 - 1024; }

    int pe_b200_compare(double const* x, int64_t LS, int32_t n_lanes, int32_t const* ab, int32_t n_cmp, uint8_t* out, void*)
    {
        for(int32_t c = 0; c < n_cmp; ++c)
        {
            for(int64_t l = 0; l < n_lanes; ++l)
            {
                double const va = ab[2 * c] < 0 ? 0.0 : x[(int64_t)ab[2 * c] * LS + l];
                double const vb = ab[2 * c + 1] < 0 ? 0.0 : x[(int64_t)ab[2 * c + 1] * LS + l];
                out[(int64_t)c * LS + l] = va >= vb ? 1 : 0;
            }
        }
        ++g_launches;
        return 0;
    }

    int pe_b200_status_reduce(int32_t const* status, uint32_t const* solves, int64_t n_lanes, unsigned long long* out3, void*)
    {
        out3[0] = out3[1] = out3[2] = 0;
        for(int64_t i = 0; i < n_lanes; ++i)
        {
            out3[0] += status[i] != PE_ST_OK;
            out3[1] += status[i] == PE_ST_SINGULAR;
            out3[2] += solves[i];
        }
        ++g_aux_launches;
        return 0;
    }

    uint64_t pe_emu_resident_launches(void);
    uint64_t pe_emu_races(void) { return emu_trace::g_races; }
    uint64_t pe_emu_unbalanced_barriers(void) { return g_unbalanced; }
    void pe_emu_trace(int on) { emu_trace::g_on = on != 0; }

    // Same control flow as pe_b200_solve_kernel (csrc/pe_b200_kernels.cu), one lane at a time.
    int pe_b200_launch(pe_b200_run const* rp, void*)
    {
        if(rp == nullptr || rp->n_lanes <= 0) { return 0; }
        pe_b200_run const& r = *rp;
        ++g_launches;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};
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

// ---- resident programs: one CTA (S streams x I instances, workspace in "shared memory") at a time --------------------
namespace
{
    uint64_t g_resident_launches = 0;
    uint64_t g_stream_errors = 0, g_stream_launches = 0;

    template <int J>
    void run_resident(pe_b200_rrun const& r)
    {
        using namespace pe_rinterp;
        uint32_t const I = (uint32_t)r.I, IG = I / J, S = (uint32_t)r.S;
        uint32_t const T = S * IG, W = T / 32, C = 32 / IG;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};
        int64_t const n_cta = (r.n_lanes + I - 1) / I;
        bool const hbm = r.wsg != nullptr;
        uint64_t const sstride = hbm ? (uint64_t)r.LSw : (uint64_t)I;  // distance between consecutive slots
        std::vector<double> ws_local(hbm ? 0 : (size_t)r.n_slots * I);
        // column of instance j of instance-group ig inside the CTA's block of I lanes: consecutive in shared memory, 32
        // apart in the HBM form
        auto col_of = [&](uint32_t ig, int j) -> uint32_t { return hbm ? (uint32_t)j * 32u + ig : ig * (uint32_t)J + (uint32_t)j; };
        struct tstate
        {
            bool real_lane[J], counted[J], ok[J], done[J];
            int32_t status[J];
            uint32_t solves[J];
        };
        std::vector<tstate> ts(T);
        // (chunk, group) items of the tree-streaming form are replayed in the order the device hands them out
        bool const chunked = hbm && r.sched != nullptr;
        int32_t const NC = chunked ? r.n_chunks : 1;
        for(int32_t chunk = 0; chunk < NC; ++chunk)
        for(int64_t cta = 0; cta < n_cta; ++cta)
        {
            bool const first_chunk = chunk == 0, last_chunk = chunk + 1 == NC;
            int32_t const s_begin = chunked ? chunk * r.chunk_steps : 0;
            int32_t const s_end = chunked ? std::min(r.n_steps, s_begin + r.chunk_steps) : r.n_steps;
            // poison the workspace: a slot read before it is loaded or written shows up as NaN in the results
            double* const wsb = hbm ? r.wsg + cta * I : ws_local.data();
            if(!first_chunk) {}
            else if(hbm)
            {
                for(int32_t k = 0; k < r.n_slots; ++k)
                {
                    for(uint32_t i = 0; i < I; ++i) { wsb[(uint64_t)k * sstride + i] = __builtin_nan(""); }
                }
            }
            else
            {
                for(auto& v: ws_local) { v = __builtin_nan(""); }
            }
            emu_trace::g_map.clear();
            emu_trace::g_phase = 0;
            uint32_t s_flags[3][128] = {};
            for(uint32_t tid = 0; tid < T; ++tid)
            {
                uint32_t const ig = tid % IG;
                auto& st = ts[tid];
                for(int j = 0; j < J; ++j)
                {
                    int64_t const lane_j = cta * I + col_of(ig, j);
                    st.real_lane[j] = lane_j < r.n_lanes;
                    st.status[j] = st.real_lane[j] ? r.status[lane_j] : (int32_t)PE_ST_SINGULAR;
                    st.counted[j] = st.real_lane[j] && st.status[j] == PE_ST_OK;
                    st.ok[j] = st.counted[j];
                    st.solves[j] = 0;
                }
            }
            bool const trace_was = emu_trace::g_on;
            emu_trace::g_on = false;
            for(int32_t e = 0; first_chunk && e < r.n_io; ++e)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_LOAD)) { continue; }
                uint32_t const kind = (io.slot_kind >> 16) & 0xfu;
                for(uint32_t i = 0; i < I; ++i)
                {
                    int64_t const lane = cta * I + i;
                    bool const real = lane < r.n_lanes;
                    double v = 0.0;
                    if(kind == PE_IO_CONST) { v = r.cst[io.src]; }
                    else if(hbm && !real) { v = 0.0; }
                    else if(kind == PE_IO_U) { v = r.wu[(int64_t)io.src * r.LSu + lane]; }
                    else { v = r.wx[(int64_t)io.src * r.LSx + (real ? lane / r.ppi : 0)]; }
                    wsb[(uint64_t)(io.slot_kind & 0xffffu) * sstride + i] = v;
                }
            }
            emu_trace::g_on = trace_was;
            double t = chunked ? r.t_chunk[chunk] : r.t0;

            // one section: warps advance phase by phase; within a phase thread after thread (no cross-thread
            // dependency may exist inside a phase: the race detector checks exactly that)
            auto run_section = [&](int sec, bool use_done, bool check, std::vector<std::array<bool, 4>>& nconv, std::vector<std::array<bool, 4>>& fail, bool first_iter)
            {
                std::vector<host_reader> rdw(W);
                for(uint32_t w = 0; w < W; ++w)
                {
                    rdw[w].p0 = r.words + r.sec_off[sec * W + w];
                    rdw[w].q = r.words + r.sec_off[3 * W + sec * W + w];
                    rdw[w].C = C;
                    rdw[w].col = 0;
                    rdw[w].start();
                }
                for(;;)
                {
                    int n_bar = 0, n_end = 0;
                    for(uint32_t w = 0; w < W; ++w)
                    {
                        for(;;)
                        {
                            int kind = V_END;
                            host_reader after = rdw[w];
                            for(uint32_t l = 0; l < 32; ++l)
                            {
                                uint32_t const tid = w * 32 + l;
                                uint32_t const ig = tid % IG;
                                rctx c;
                                c.ws = wsb + col_of(ig, 0);
                                c.js = hbm ? 32u : 1u;
                                c.I = sstride;
                                c.S = S;
                                c.C = C;
                                c.col = l / IG;
                                c.stream = tid / IG;
                                emu_trace::g_warp = (int)(tid / IG);
                                bool en[J], nc[J], fl[J];
                                for(int j = 0; j < J; ++j)
                                {
                                    en[j] = use_done ? !ts[tid].done[j] : ts[tid].ok[j];
                                    nc[j] = fl[j] = false;
                                }
                                host_reader rd = rdw[w];
                                rd.col = c.col;
                                kind = rvop<J>(rd, c, t, tol, en, check, nc, fl, first_iter);
                                if(kind == V_OK) { rd.close(); }
                                if(kind == V_BAR) { rd.bar(); }
                                if(kind == V_SKIP) { rd.skip(); }
                                after = rd;
                                for(int j = 0; j < J; ++j)
                                {
                                    nconv[tid][j] = nconv[tid][j] || nc[j];
                                    fail[tid][j] = fail[tid][j] || fl[j];
                                }
                                if(kind == V_END || kind == V_BAR || kind == V_BAD || kind == V_SKIP) { break; }
                            }
                            if(kind == V_BAD)
                            {
                                if(getenv("PE_EMU_DEBUG")) { fprintf(stderr, "emu: bad op sec %d warp %u off %ld word %08x\n", sec, w, (long)(rdw[w].p - r.words), *rdw[w].p); }
                                for(uint32_t l = 0; l < 32; ++l)
                                {
                                    for(int j = 0; j < J; ++j) { fail[w * 32 + l][j] = true; }
                                }
                                ++n_end;
                                break;
                            }
                            if(kind == V_END)
                            {
                                ++n_end;
                                break;
                            }
                            rdw[w] = after;
                            if(kind == V_BAR)
                            {
                                ++n_bar;
                                break;
                            }
                        }
                    }
                    ++emu_trace::g_phase;
                    if(n_bar != 0 && n_end != 0)
                    {
                        ++g_unbalanced;
                        return;
                    }
                    if(n_end != 0) { return; }
                }
            };
            std::vector<std::array<bool, 4>> nconv(T), fail(T);
            auto clear_flags = [&]()
            {
                for(auto& a: nconv) { a = {false, false, false, false}; }
                for(auto& a: fail) { a = {false, false, false, false}; }
            };
            if(r.has_prep && first_chunk)
            {
                clear_flags();
                run_section(0, false, false, nconv, fail, true);
            }
            for(int32_t s = s_begin; s < s_end; ++s)
            {
                if(r.time_stepping)
                {
                    if(r.has_step)
                    {
                        clear_flags();
                        run_section(1, false, false, nconv, fail, true);
                    }
                    t = t + r.dt;
                }
                for(uint32_t tid = 0; tid < T; ++tid)
                {
                    for(int j = 0; j < J; ++j) { ts[tid].done[j] = !ts[tid].ok[j]; }
                }
                int32_t it = 0;
                int fi = 0;
                for(;;)
                {
                    for(uint32_t i = 0; i < 128; ++i) { s_flags[fi][i] = 0; }
                    clear_flags();
                    run_section(2, true, r.nonlinear != 0, nconv, fail, it == 0);
                    for(uint32_t tid = 0; tid < T; ++tid)
                    {
                        uint32_t const ig = tid % IG;
                        for(int j = 0; j < J; ++j) { s_flags[fi][col_of(ig, j)] |= (nconv[tid][j] ? 1u : 0u) | (fail[tid][j] ? 2u : 0u); }
                    }
                    ++it;
                    bool all_done = true;
                    for(uint32_t tid = 0; tid < T; ++tid)
                    {
                        uint32_t const ig = tid % IG;
                        auto& st = ts[tid];
                        for(int j = 0; j < J; ++j)
                        {
                            uint32_t const f = s_flags[fi][col_of(ig, j)];
                            if(!st.done[j])
                            {
                                ++st.solves[j];
                                if(f & 2u)
                                {
                                    st.status[j] = PE_ST_SINGULAR;
                                    st.ok[j] = false;
                                    st.done[j] = true;
                                }
                                else if(!r.nonlinear || !(f & 1u)) { st.done[j] = true; }
                                else if(it >= r.max_iter)
                                {
                                    st.status[j] = PE_ST_NO_CONVERGENCE;
                                    st.ok[j] = false;
                                    st.done[j] = true;
                                }
                            }
                            all_done = all_done && st.done[j];
                        }
                    }
                    if(all_done) { break; }
                }
                if(r.wave != nullptr)
                {
                    for(uint32_t ig = 0; ig < IG; ++ig)
                    {
                        auto const& st = ts[ig];  // stream 0
                        for(int j = 0; j < J; ++j)
                        {
                            if(!st.ok[j]) { continue; }
                            int64_t const lane = cta * I + col_of(ig, j);
                            for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane] = wsb[(uint64_t)r.probes[p] * sstride + col_of(ig, j)]; }
                        }
                    }
                }
            }
            for(int32_t e = 0; last_chunk && e < r.n_io; ++e)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_STORE)) { continue; }
                for(uint32_t ig = 0; ig < IG; ++ig)
                {
                    for(int j = 0; j < J; ++j)
                    {
                        if(ts[ig].counted[j]) { r.wu[(int64_t)io.src * r.LSu + cta * I + col_of(ig, j)] = wsb[(uint64_t)(io.slot_kind & 0xffffu) * sstride + col_of(ig, j)]; }
                    }
                }
            }
            for(uint32_t ig = 0; ig < IG; ++ig)
            {
                for(int j = 0; j < J; ++j)
                {
                    if(ts[ig].counted[j])
                    {
                        r.status[cta * I + col_of(ig, j)] = ts[ig].status[j];
                        r.solves[cta * I + col_of(ig, j)] += ts[ig].solves[j];
                    }
                }
            }
        }
    }
}  // namespace

extern "C"
{
    uint64_t pe_emu_resident_launches(void) { return g_resident_launches; }

    // the emulator replays the word interpreter only: the specialised tree kernel is GPU-only
    int pe_b200_jit_supported(void) { return 0; }
    int pe_b200_launch_jit(pe_b200_rrun const*, void const*, size_t, uint64_t, void*)
    {
        snprintf(g_err, sizeof(g_err), "pe_b200_launch_jit: not available in the emulator");
        return 1;
    }

    // ---- reduce-and-core path: the same steps as csrc/pe_b200_frontal.cu, one instance at a time on the host -----------
    char const* pe_b200_frontal_last_error(void) { return g_err; }
    int pe_b200_frontal_run(pe_b200_frontal const* fp, void*, uint64_t* n_launches)
    {
        if(fp == nullptr || fp->n_inst <= 0) { return 0; }
        pe_b200_frontal const& f = *fp;
        int64_t const B = f.B, ld = f.ld_core;
        for(int64_t inst = 0; inst < f.n_inst; ++inst)
        {
            for(int64_t k = 0; k < f.n_nodes; ++k) { f.d[k * B + inst] = f.z[k * B + inst] = 0.0; }
            for(int64_t k = 0; k < f.n_edges; ++k) { f.g[k * B + inst] = 0.0; }
            for(int64_t r = 0; r < f.n_res; ++r)
            {
                double const g = 1.0 / f.rval[r * B + inst];
                if(f.res_a[r] >= 0) { f.d[(int64_t)f.res_a[r] * B + inst] += g; }
                if(f.res_b[r] >= 0) { f.d[(int64_t)f.res_b[r] * B + inst] += g; }
                if(f.res_edge[r] >= 0) { f.g[(int64_t)f.res_edge[r] * B + inst] += g; }
            }
            for(int64_t k = 0; k < f.n_idc; ++k)
            {
                double const i = f.idc_val[k * B + inst];
                if(f.idc_p[k] >= 0) { f.z[(int64_t)f.idc_p[k] * B + inst] -= i; }
                if(f.idc_q[k] >= 0) { f.z[(int64_t)f.idc_q[k] * B + inst] += i; }
            }
            int64_t const n_ops = f.level_off_host[f.n_levels];
            for(int64_t o = 0; o < n_ops; ++o)
            {
                int32_t const* op = f.ops + o * 6;
                int32_t const k = op[0], a = op[1], b = op[2];
                double const p = f.d[(int64_t)k * B + inst];
                if(p == 0.0 || !std::isfinite(p)) { f.status[inst] = PE_ST_SINGULAR; }
                double const zk = f.z[(int64_t)k * B + inst];
                double const g1 = a >= 0 ? f.g[(int64_t)op[3] * B + inst] : 0.0, g2 = b >= 0 ? f.g[(int64_t)op[4] * B + inst] : 0.0;
                double const m1 = g1 / p, m2 = g2 / p;
                if(a >= 0)
                {
                    f.d[(int64_t)a * B + inst] -= g1 * m1;
                    f.z[(int64_t)a * B + inst] += m1 * zk;
                }
                if(b >= 0)
                {
                    f.d[(int64_t)b * B + inst] -= g2 * m2;
                    f.z[(int64_t)b * B + inst] += m2 * zk;
                }
                if(op[5] >= 0) { f.g[(int64_t)op[5] * B + inst] += g1 * m2; }
            }
            double* const M = f.M + inst * ld * ld;
            double* const c = f.c + inst * ld;
            for(int64_t k = 0; k < ld * ld; ++k) { M[k] = 0.0; }
            for(int64_t t = 0; t < ld; ++t)
            {
                c[t] = 0.0;
                if(t < f.n_core_nodes)
                {
                    M[t + t * ld] = f.d[(int64_t)f.core_unknown[t] * B + inst];
                    c[t] = f.z[(int64_t)f.core_unknown[t] * B + inst];
                }
                else if(t < f.n_core)
                {
                    int64_t const k = t - f.n_core_nodes;
                    int32_t const p = f.vdc[3 * k + 1], q = f.vdc[3 * k + 2];
                    if(p >= 0) { M[p + t * ld] = M[t + p * ld] = 1.0; }
                    if(q >= 0) { M[q + t * ld] = M[t + q * ld] = -1.0; }
                    c[t] = f.vdc_val[k * B + inst];
                }
                else
                {
                    M[t + t * ld] = 1.0;
                }
            }
            for(int64_t k = 0; k < f.n_core_edges; ++k)
            {
                int32_t const i = f.core_edges[3 * k], j = f.core_edges[3 * k + 1];
                double const g = f.g[(int64_t)f.core_edges[3 * k + 2] * B + inst];
                M[i + j * ld] = M[j + i * ld] = -g;
            }
            for(int64_t k = 0; k < ld; ++k)
            {
                double const p = M[k + k * ld];
                if(p == 0.0 || !std::isfinite(p)) { f.status[inst] = PE_ST_SINGULAR; }
                for(int64_t i = k + 1; i < ld; ++i) { M[i + k * ld] /= p; }
                for(int64_t j = k + 1; j < ld; ++j)
                {
                    double const u = M[k + j * ld];
                    if(u == 0.0) { continue; }
                    for(int64_t i = k + 1; i < ld; ++i) { M[i + j * ld] -= M[i + k * ld] * u; }
                }
            }
            for(int64_t k = 0; k < ld; ++k)
            {
                for(int64_t i = k + 1; i < ld; ++i) { c[i] -= M[i + k * ld] * c[k]; }
            }
            for(int64_t k = ld - 1; k >= 0; --k)
            {
                c[k] /= M[k + k * ld];
                for(int64_t i = 0; i < k; ++i) { c[i] -= M[i + k * ld] * c[k]; }
            }
            for(int64_t t = 0; t < f.n_core; ++t) { f.x[(int64_t)f.core_unknown[t] * f.LSx + inst] = c[t]; }
            for(int64_t o = n_ops - 1; o >= 0; --o)
            {
                int32_t const* op = f.ops + o * 6;
                int32_t const k = op[0], a = op[1], b = op[2];
                double v = f.z[(int64_t)k * B + inst];
                if(a >= 0) { v += f.g[(int64_t)op[3] * B + inst] * f.x[(int64_t)a * f.LSx + inst]; }
                if(b >= 0) { v += f.g[(int64_t)op[4] * B + inst] * f.x[(int64_t)b * f.LSx + inst]; }
                f.x[(int64_t)k * f.LSx + inst] = v / f.d[(int64_t)k * B + inst];
            }
        }
        ++g_launches;
        if(n_launches != nullptr) { *n_launches = 1; }
        if(f.phase_ms_host != nullptr) { f.phase_ms_host[0] = f.phase_ms_host[1] = f.phase_ms_host[2] = 0.0; }
        return 0;
    }

    // ---- stream kernel: the generated source is compiled for the host (stream_host.cpp) and run lane by lane ------------
    int pe_b200_stream_supported(void) { return 2; }
    uint64_t pe_emu_stream_errors(void) { return g_stream_errors; }
    uint64_t pe_emu_stream_launches(void) { return g_stream_launches; }

    int pe_b200_stream_build(char const* source_path, char const* out_path, char const* csrc_dir, int, int, char* log, size_t log_cap)
    {
        Dl_info info{};
        std::string here{"."};
        if(dladdr(reinterpret_cast<void const*>(&pe_b200_stream_supported), &info) != 0 && info.dli_fname != nullptr)
        {
            std::string p{info.dli_fname};
            auto const k = p.rfind('/');
            if(k != std::string::npos) { here = p.substr(0, k); }
        }
        std::string const logf = std::string(out_path) + ".log";
        std::string const cmd = std::string("g++ -std=c++17 -O1 -fPIC -shared -Wl,--exclude-libs,ALL -ffp-contract=off -DPE_SJ=1 '-DPE_STREAM_SOURCE=\"") + source_path + "\"' -I'" + csrc_dir + "' -o '" + out_path +
                                "' '" + here + "/stream_host.cpp' > '" + logf + "' 2>&1";
        int const rc = system(cmd.c_str());
        if(rc != 0 && log != nullptr && log_cap > 0)
        {
            log[0] = 0;
            if(FILE* f = fopen(logf.c_str(), "rb"))
            {
                size_t const n = fread(log, 1, log_cap - 1, f);
                log[n] = 0;
                fclose(f);
            }
        }
        return rc == 0 ? 0 : 1;
    }

    void pe_b200_stream_last_geometry(int* out3) { out3[0] = out3[1] = out3[2] = 0; }

    int pe_b200_launch_stream(pe_b200_rrun const* rp, void const* blob, size_t bytes, uint64_t, uint32_t n_tiles, uint32_t stage_rows, void*)
    {
        using namespace pe_rinterp;
        if(rp == nullptr || rp->n_lanes <= 0) { return 0; }
        pe_b200_rrun const& r = *rp;
        std::string const path(static_cast<char const*>(blob), bytes);
        static std::map<std::string, void*> mods;
        void* h = mods[path];
        if(h == nullptr)
        {
            h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
            if(h == nullptr)
            {
                snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: dlopen %s: %s", path.c_str(), dlerror());
                return 1;
            }
            mods[path] = h;
        }
        auto const f_tiles = reinterpret_cast<uint32_t (*)(void)>(dlsym(h, "pe_emu_stream_tiles"));
        auto const f_rows = reinterpret_cast<uint32_t (*)(void)>(dlsym(h, "pe_emu_stream_stage_rows"));
        auto const f_new = reinterpret_cast<void* (*)(uint32_t, uint32_t)>(dlsym(h, "pe_emu_stream_new"));
        auto const f_free = reinterpret_cast<void (*)(void*)>(dlsym(h, "pe_emu_stream_free"));
        auto const f_solve = reinterpret_cast<uint32_t (*)(void*, double*, uint64_t, uint32_t, uint32_t, uint64_t*, int)>(dlsym(h, "pe_emu_stream_solve"));
        auto const f_has_prep = reinterpret_cast<int (*)(void)>(dlsym(h, "pe_emu_stream_has_prep"));
        auto const f_prep = reinterpret_cast<void (*)(void*, double*, uint64_t, uint32_t, uint32_t, uint64_t*)>(dlsym(h, "pe_emu_stream_prep"));
        bool const mod_prep = f_has_prep != nullptr && f_prep != nullptr && f_has_prep() != 0;
        if(!f_tiles || !f_rows || !f_new || !f_free || !f_solve || f_tiles() != n_tiles || f_rows() != stage_rows || r.S != 1 || r.wsg == nullptr || r.nonlinear || r.cplx)
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: module / launch mismatch");
            return 1;
        }
        ++g_launches;
        ++g_stream_launches;
        bool const trace_was = emu_trace::g_on;
        emu_trace::g_on = false;
        uint32_t const GL = (uint32_t)r.I;  // 32 J lanes per group: the workspace is one block per group, ws[group][row][GL]
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};
        static int const ns_log_env = getenv("PE_EMU_STREAM_NS_LOG") ? atoi(getenv("PE_EMU_STREAM_NS_LOG")) : 2;
        for(int64_t lane = 0; lane < ((int64_t)r.n_lanes + GL - 1) / GL * GL; ++lane)
        {
            int64_t const group = lane / GL;
            uint32_t const li = (uint32_t)(lane % GL);
            double* const wl = r.wsg + group * (int64_t)r.n_slots * GL + li;
            bool const real = lane < r.n_lanes;
            int32_t status = real ? r.status[lane] : (int32_t)PE_ST_SINGULAR;
            bool const counted = real && status == PE_ST_OK;
            bool ok = counted;
            uint32_t solves = 0;
            for(int32_t k = 0; k < r.n_slots; ++k) { wl[(uint64_t)k * GL] = __builtin_nan(""); }  // poison
            for(int32_t e = 0; e < r.n_io; ++e)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_LOAD)) { continue; }
                uint32_t const kind = (io.slot_kind >> 16) & 0xfu;
                double v = 0.0;
                if(kind == PE_IO_CONST) { v = r.cst[io.src]; }
                else if(real) { v = kind == PE_IO_U ? r.wu[(int64_t)io.src * r.LSu + lane] : r.wx[(int64_t)io.src * r.LSx + lane / r.ppi]; }
                wl[(uint64_t)(io.slot_kind & 0xffffu) * GL] = v;
            }
            rctx c;
            c.ws = wl;
            c.I = GL;
            c.S = 1;
            c.C = 1;
            c.col = 0;
            c.stream = 0;
            c.js = 32;
            auto run_section = [&](int sec, double t)
            {
                host_reader rd;
                rd.p0 = r.words + r.sec_off[sec];
                rd.q = r.words + r.sec_off[3 + sec];
                rd.C = 1;
                rd.col = 0;
                rd.start();
                bool en[1] = {ok}, nc[1] = {false}, fl[1] = {false};
                for(;;)
                {
                    int const kind = rvop<1>(rd, c, t, tol, en, false, nc, fl, true);
                    if(kind == V_END || kind == V_BAD) { break; }
                    if(kind == V_BAR) { rd.bar(); }
                    else if(kind == V_SKIP) { rd.skip(); }
                    else
                    {
                        rd.close();
                    }
                }
            };
            void* const st = f_new((uint32_t)r.n_slots, (uint32_t)ns_log_env);
            if(auto const f_guard = reinterpret_cast<void (*)(void*, double)>(dlsym(h, "pe_emu_stream_set_guard")); f_guard != nullptr) { f_guard(st, r.guard); }
            auto const f_time = reinterpret_cast<void (*)(void*, double)>(dlsym(h, "pe_emu_stream_set_time"));
            double t = r.t0;
            if(r.has_prep)
            {
                if(mod_prep) { f_prep(st, wl, GL, li & 31u, ok ? 1u : 0u, &g_stream_errors); }
                else
                {
                    run_section(0, t);
                }
            }
            for(int32_t s = 0; s < r.n_steps; ++s)
            {
                if(r.time_stepping)
                {
                    if(r.has_step) { run_section(1, t); }
                    t = t + r.dt;
                }
                if(f_time != nullptr) { f_time(st, t); }
                uint32_t const fm = f_solve(st, wl, GL, li & 31u, ok ? 1u : 0u, &g_stream_errors, s == 0 ? 1 : 0);
                if(ok)
                {
                    ++solves;
                    if(fm & 1u)
                    {
                        status = PE_ST_SINGULAR;
                        ok = false;
                    }
                }
                if(r.wave != nullptr && ok)
                {
                    for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane] = wl[(uint64_t)r.probes[p] * GL]; }
                }
            }
            f_free(st);
            for(int32_t e = 0; e < r.n_io; ++e)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_STORE)) { continue; }
                if(counted) { r.wu[(int64_t)io.src * r.LSu + lane] = wl[(uint64_t)(io.slot_kind & 0xffffu) * GL]; }
            }
            if(counted)
            {
                r.status[lane] = status;
                r.solves[lane] += solves;
            }
        }
        emu_trace::g_on = trace_was;
        return 0;
    }

    int pe_b200_launch_resident(pe_b200_rrun const* rp, void*)
    {
        if(rp == nullptr || rp->n_lanes <= 0) { return 0; }
        int const I = rp->I, J = rp->J, S = rp->S;
        if(I < 1 || I > 128 || (J != 1 && J != 2 && J != 4) || I % J != 0 || S < 1 || (32 % (I / J)) != 0 || (S * (I / J)) % 32 != 0 || S * (I / J) > 1024 ||
           (rp->wsg == nullptr && I > 32) || (rp->wsg != nullptr && I != 32 * J))
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: bad geometry S=%d I=%d J=%d", S, I, J);
            return 1;
        }
        ++g_launches;
        ++g_resident_launches;
        if(J == 4) { run_resident<4>(*rp); }
        else if(J == 2) { run_resident<2>(*rp); }
        else
        {
            run_resident<1>(*rp);
        }
        return 0;
    }
}
