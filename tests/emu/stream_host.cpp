// TEST INFRASTRUCTURE ONLY — host build of the stream kernel's generated code (phy-engine_b200/host/stream.cpp).
//
// pe_b200_stream_build of the emulator (emu.cpp) compiles THIS file with g++ together with one generated source
// (-DPE_STREAM_SOURCE=...) into a shared object; the emulator then runs the iter section of every lane through it, one
// lane at a time, with the host form of csrc/pe_b200_stream.h: bulk copies are performed at issue time (the earliest
// moment the hardware could perform them) and every ordering rule of the ring is checked (a copy of a row whose last
// store is not fenced, a stage refilled before it was read, a wait on a tile nobody issued).  Never part of the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define __device__
#define __forceinline__ inline
#include "pe_b200_program.h"
#include "pe_b200_stream.h"

namespace
{
    using namespace pe_stream;
#include PE_STREAM_SOURCE

    struct lane_state
    {
        sk_ctx k{};
        std::vector<double> ring;
        std::vector<uint64_t> store_seq;
        std::vector<int64_t> stage_tile, stage_read;
    };
}  // namespace

extern "C"
{
    uint32_t pe_emu_stream_tiles(void) { return PE_STREAM_TILES; }
    uint32_t pe_emu_stream_stage_rows(void) { return PE_STREAM_STAGE_ROWS; }

    void* pe_emu_stream_new(uint32_t n_rows, uint32_t ns_log)
    {
        auto* s = new lane_state;
        uint32_t const ns = 1u << ns_log;
        s->ring.assign((size_t)ns * PE_STREAM_STAGE_ROWS, __builtin_nan(""));
        s->store_seq.assign(n_rows, 0);
        s->stage_tile.assign(ns, -1);
        s->stage_read.assign(ns, 1);
        s->k.ring = s->ring.data();
        s->k.stage_rows = PE_STREAM_STAGE_ROWS;
        s->k.ns_mask = ns - 1u;
        s->k.ns_log = ns_log;
        s->k.seq0 = 0;
        s->k.pn = 0;
        s->k.fenced = -1;
        s->k.n_rows = n_rows;
        s->k.store_seq = s->store_seq.data();
        s->k.seq_counter = 0;
        s->k.fence_seq = 0;
        s->k.stage_tile = s->stage_tile.data();
        s->k.stage_read = s->stage_read.data();
        s->k.errors = 0;
        return s;
    }

    void pe_emu_stream_free(void* p) { delete static_cast<lane_state*>(p); }
    void pe_emu_stream_set_guard(void* p, double guard) { static_cast<lane_state*>(p)->k.guard = guard; }
    void pe_emu_stream_set_time(void* p, double t) { static_cast<lane_state*>(p)->k.t = t; }

    // stores made by the interpreted sections / the load table between two solves are not tracked: every section starts
    // with a fence anyway.  Returns the pivot-failure mask; *errors accumulates the ordering violations.
    uint32_t pe_emu_stream_solve(void* p, double* wl, uint64_t GL, uint32_t lane, uint32_t enm, uint64_t* errors, int first)
    {
        auto* s = static_cast<lane_state*>(p);
        s->k.wl = wl;
        s->k.GL = GL;
        s->k.lane = lane;
        s->k.enm = enm;
        uint32_t fm = 0;
#ifdef PE_STREAM_STEADY
        if(first) { pe_stream_iter(s->k, fm); }
        else
        {
            pe_stream_iters(s->k, fm);
        }
#else
        (void)first;
        pe_stream_iter(s->k, fm);
#endif
        *errors += s->k.errors;
        s->k.errors = 0;
        return fm;
    }

    // 1 = the module holds the prep section too (pe_emu_stream_prep runs it), 0 = the caller interprets it
    int pe_emu_stream_has_prep(void)
    {
#ifdef PE_STREAM_PREP
        return 1;
#else
        return 0;
#endif
    }

    void pe_emu_stream_prep(void* p, double* wl, uint64_t GL, uint32_t lane, uint32_t enm, uint64_t* errors)
    {
#ifdef PE_STREAM_PREP
        auto* s = static_cast<lane_state*>(p);
        s->k.wl = wl;
        s->k.GL = GL;
        s->k.lane = lane;
        s->k.enm = enm;
        uint32_t fm = 0;
        pe_stream_prep(s->k, fm);
        *errors += s->k.errors;
        s->k.errors = 0;
#else
        (void)p, (void)wl, (void)GL, (void)lane, (void)enm, (void)errors;
#endif
    }
}
