// pe_b200_stream.h — run-time interface of the STREAM kernel's generated code (host/stream.cpp).
//
// The stream kernel runs one warp per lane group (32 J lanes) and one word stream per group (S = 1): the whole
// elimination / substitution sweep of a solve (circuit.h:987-1527: stamp, SparseLU factorize + solve) is one serial
// program per warp, with no barrier and no fill-in from a parallel split of the chain.  Parallelism comes from the lane
// groups only (>= 2 warps per SM), memory latency is hidden by TMA: the generated code is cut into TILES, the cold
// operand rows of a tile (parameters, device state, factors written by an earlier sweep) are fetched into a per-warp
// shared-memory ring by bulk asynchronous copies (cp.async.bulk ... mbarrier::complete_tx, one elected lane) NS tiles
// ahead of their use; values produced and consumed within a tile, or handed from one tile to the next, stay in
// registers; results needed later are stored with plain coalesced stores.
//
// This header is compiled three ways:
//   * nvcc, sm_100a (csrc/pe_b200_stream.cu): the real kernel, PE_SK_DEVICE
//   * g++ (tests/emu/stream_host.cpp, TEST INFRASTRUCTURE): the same generated source run lane by lane on host memory
//     with eager copies and hazard checks (a copy of a row whose last store has not been fenced, a ring stage
//     overwritten before it was read, a wait for a tile that was never issued all abort the run)
// Generated code only uses the j* arithmetic helpers and the SK_* macros below.
#pragma once
#include <stdint.h>

#include "pe_b200_models.h"

#ifndef PE_SJ
#define PE_SJ 1
#endif
#ifndef PE_SGL
#define PE_SGL 32  // lanes of a warp that carry a lane of the group (16 / 8: the others idle; PE_SJ must be 1 then)
#endif

namespace pe_stream
{
    struct jv
    {
        double v[PE_SJ];
    };

#if defined(__CUDA_ARCH__)
#define PE_SK_FN __device__ __forceinline__
#else
#define PE_SK_FN inline
#endif

    PE_SK_FN jv jzero()
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = 0.0; }
        return x;
    }
    // acc +/- s  (the interpreter's fma(s, +/-1.0, acc): one rounding, the same value)
    PE_SK_FN void jadd(jv& acc, jv const& s)
    {
        for(int j = 0; j < PE_SJ; ++j) { acc.v[j] = PE_ADD(acc.v[j], s.v[j]); }
    }
    PE_SK_FN void jsub(jv& acc, jv const& s)
    {
        for(int j = 0; j < PE_SJ; ++j) { acc.v[j] = PE_SUB(acc.v[j], s.v[j]); }
    }
#if defined(__CUDA_ARCH__)
#define PE_SK_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define PE_SK_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif
    // acc -/+ a b
    PE_SK_FN void jfms(jv& acc, jv const& a, jv const& b)
    {
        for(int j = 0; j < PE_SJ; ++j) { acc.v[j] = PE_SK_FMA(-a.v[j], b.v[j], acc.v[j]); }
    }
    PE_SK_FN void jfma(jv& acc, jv const& a, jv const& b)
    {
        for(int j = 0; j < PE_SJ; ++j) { acc.v[j] = PE_SK_FMA(a.v[j], b.v[j], acc.v[j]); }
    }
    PE_SK_FN void jmul(jv& acc, jv const& s)
    {
        for(int j = 0; j < PE_SJ; ++j) { acc.v[j] = PE_MUL(acc.v[j], s.v[j]); }
    }
    PE_SK_FN void jrcp(jv& acc, uint32_t& failm)
    {
        for(int j = 0; j < PE_SJ; ++j)
        {
            if(acc.v[j] == 0.0 || !isfinite(acc.v[j])) { failm |= 1u << j; }
            acc.v[j] = PE_RCP(acc.v[j]);
        }
    }
    // pivot guard (pe_b200_program.h PE_F_GUARD): an entry of L out of bounds marks the lane
    PE_SK_FN void jguard(jv const& l, double const guard, uint32_t& failm)
    {
        for(int j = 0; j < PE_SJ; ++j)
        {
            if(PE_GUARD_TRIP(fabs(l.v[j]), guard)) { failm |= 1u << j; }
        }
    }
    // CAP_STEP (capacitor.h:106-128): hist, prev_g updated in place
    PE_SK_FN void jcap(jv const& C, jv const& dt, jv const& va, jv const& vb, jv& hist, jv& prev_g)
    {
        for(int j = 0; j < PE_SJ; ++j) { pe_models::cap_step(C.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), hist.v[j], prev_g.v[j]); }
    }

    // VSIN (VAC.h:176, IAC.h:154): Vp sin(omega t + phase), the expression of pe_b200_rinterp.h
    PE_SK_FN jv jvsin(jv const& vp, jv const& om, jv const& ph, double const t)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = PE_MUL(vp.v[j], sin(PE_ADD(PE_MUL(om.v[j], t), ph.v[j]))); }
        return x;
    }
    // GEN_EVAL (generator/*.h: sawtooth / square / pulse / triangle at the section time or at t = 0), the expression of
    // pe_b200_rinterp.h
    PE_SK_FN jv jvgen(jv const& kd, jv const& ts, jv const& vh, jv const& vl, jv const& fq, jv const& du, jv const& ph, jv const& tr_, jv const& tf_, double const t)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = pe_models::gen_eval((int)kd.v[j], ts.v[j] != 0.0 ? t : 0.0, vh.v[j], vl.v[j], fq.v[j], du.v[j], ph.v[j], tr_.v[j], tf_.v[j]); }
        return x;
    }
    // IND_STEP (inductor.h:134-160): req, ueq from the previous step's voltage and branch current
    PE_SK_FN void jind(jv const& L, jv const& dt, jv const& va, jv const& vb, jv const& ib, jv& req, jv& ueq)
    {
        for(int j = 0; j < PE_SJ; ++j) { pe_models::ind_step(L.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), ib.v[j], req.v[j], ueq.v[j]); }
    }

    // CAP_STEP when prev_g is known to equal 2 C / dt already (steady variant, host/stream.cpp): the same arithmetic with the
    // freshly computed value standing in for the stored one
    PE_SK_FN void jcap_steady(jv const& C, jv const& dt, jv const& va, jv const& vb, jv& hist, jv& prev_g)
    {
        for(int j = 0; j < PE_SJ; ++j)
        {
            prev_g.v[j] = PE_DIV(PE_MUL(2.0, C.v[j]), dt.v[j]);
            pe_models::cap_step(C.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), hist.v[j], prev_g.v[j]);
        }
    }

    // ... with g_new = prev_g taken from the fetched row instead of the division (bit-identical: the row holds 2 C / dt)
    PE_SK_FN void jcap_loaded(jv const& va, jv const& vb, jv& hist, jv const& prev_g)
    {
        for(int j = 0; j < PE_SJ; ++j) { hist.v[j] = PE_SUB(PE_MUL(-PE_ADD(prev_g.v[j], prev_g.v[j]), PE_SUB(va.v[j], vb.v[j])), hist.v[j]); }
    }

    // simple value ops (pe_b200_rinterp.h: PE_OP_RECIP / MUL / SUB / COPY / MUL2DIV)
    PE_SK_FN jv jneg(jv const& a)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = -a.v[j]; }
        return x;
    }
    PE_SK_FN jv jvrecip(jv const& a)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = PE_DIV(1.0, a.v[j]); }
        return x;
    }
    PE_SK_FN jv jvmul(jv const& a, jv const& b)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = PE_MUL(a.v[j], b.v[j]); }
        return x;
    }
    PE_SK_FN jv jvsub(jv const& a, jv const& b)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = PE_SUB(a.v[j], b.v[j]); }
        return x;
    }
    PE_SK_FN jv jvcopy(jv const& a) { return a; }
    PE_SK_FN jv jvmul2div(jv const& a, jv const& b)
    {
        jv x;
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = PE_DIV(PE_MUL(2.0, a.v[j]), b.v[j]); }
        return x;
    }

    // geometry of one generated program (filled in by the generated pe_stream_geom())
    struct sk_geom
    {
        uint32_t n_tiles;     // tiles per solve
        uint32_t stage_rows;  // rows of one ring stage (max over the tiles)
    };

#if defined(__CUDA_ARCH__)
    // ---- device form ---------------------------------------------------------------------------------------------
    constexpr uint32_t ROWB = 8u * PE_SGL * PE_SJ;  // bytes of one workspace row of a lane group
    constexpr uint32_t JSTR = 8u * PE_SGL;           // distance between the J lanes of a thread

    struct sk_ctx
    {
        char* wl;           // workspace block of the group, at this thread's first lane: row r at wl + r * ROWB (+ JSTR j)
        char* gbase;        // the same block at lane 0 (source of the bulk copies)
        char const* ring;   // this warp's ring (generic address), at this thread's lane
        uint32_t ring_s;    // ... shared-window address of the ring (at lane 0)
        uint32_t bars;      // shared-window address of the NS mbarriers
        uint32_t stage_bytes;
        uint32_t ns_mask, ns_log;
        uint32_t seq0;      // running tile counter at the start of this solve (mbarrier phases run across solves)
        uint32_t pn;        // next tile to issue
        int32_t fenced;     // stores of the tiles <= fenced are ordered before later bulk copies
        uint32_t lane;
        uint32_t enm;       // store mask of this thread's J lanes
        uint32_t n_rows;    // rows of the group's workspace block (bounds checks of debug builds, -DPE_SK_DEBUG)
        double guard;       // PE_F_GUARD threshold (pe_b200_program.h)
        double t;           // time of the solve being made (sources with a time dependence: VSIN)
    };

#ifdef PE_SK_DEBUG
#define PE_SK_CHECK(cond, what, a, b)                                                                                              \
    if(!(cond)) { printf("stream kernel check failed: %s (%u, %u) block %u thread %u\n", what, (unsigned)(a), (unsigned)(b), blockIdx.x, threadIdx.x); }
#else
#define PE_SK_CHECK(cond, what, a, b)
#endif

    PE_SK_FN jv sk_ld(sk_ctx const& k, uint32_t row)
    {
        jv x;
        PE_SK_CHECK(row < k.n_rows, "sk_ld row", row, k.n_rows)
        char const* p = k.wl + (size_t)row * ROWB;
#pragma unroll
        for(int j = 0; j < PE_SJ; ++j) { x.v[j] = *reinterpret_cast<double const*>(p + JSTR * j); }
        return x;
    }
    PE_SK_FN void sk_st(sk_ctx const& k, uint32_t row, jv const& x)
    {
        PE_SK_CHECK(row < k.n_rows, "sk_st row", row, k.n_rows)
        char* p = k.wl + (size_t)row * ROWB;
#pragma unroll
        for(int j = 0; j < PE_SJ; ++j)
        {
            if((k.enm >> j) & 1u) { *reinterpret_cast<double*>(p + JSTR * j) = x.v[j]; }
        }
    }
    // shared-window address of the stage of tile t, at this thread's lane
    PE_SK_FN uint32_t sk_stage(sk_ctx const& k, uint32_t t) { return k.ring_s + (k.lane & (PE_SGL - 1u)) * 8u + ((k.seq0 + t) & k.ns_mask) * k.stage_bytes; }
    PE_SK_FN jv sk_ring(uint32_t sg, uint32_t row)
    {
        jv x;
        PE_SK_CHECK(row < PE_STREAM_STAGE_ROWS, "sk_ring row", row, PE_STREAM_STAGE_ROWS)
#pragma unroll
        for(int j = 0; j < PE_SJ; ++j) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x.v[j]) : "r"(sg + row * ROWB + JSTR * j) : "memory"); }
        return x;
    }
    PE_SK_FN void sk_wait(sk_ctx const& k, uint32_t t)
    {
        uint32_t const sq = k.seq0 + t;
        uint32_t const bar = k.bars + 8u * (sq & k.ns_mask), parity = (sq >> k.ns_log) & 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tSKW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra SKD_%=;\n\tbra SKW_%=;\n\tSKD_%=:\n\t}" ::"r"(bar),
            "r"(parity)
            : "memory");
    }
    // all ring reads of the tile are in registers: the stage may be refilled
    PE_SK_FN void sk_ring_done(sk_ctx const&, uint32_t) { __syncwarp(); }
    PE_SK_FN void sk_fence(sk_ctx& k, int32_t t_done)
    {
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        k.fenced = t_done;
    }
    PE_SK_FN void sk_tx(sk_ctx const& k, uint32_t p, uint32_t rows)
    {
        if(k.lane == 0u)
        {
            uint32_t const bar = k.bars + 8u * ((k.seq0 + p) & k.ns_mask);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows * ROWB) : "memory");
        }
    }
    PE_SK_FN void sk_end(sk_ctx const&, uint32_t) {}
    PE_SK_FN void sk_copy(sk_ctx const& k, uint32_t p, uint32_t dst_row, uint32_t src_row, uint32_t rows)
    {
        if(k.lane == 0u)
        {
            uint32_t const stg = (k.seq0 + p) & k.ns_mask;
            uint32_t const bar = k.bars + 8u * stg, dst = k.ring_s + stg * k.stage_bytes + dst_row * ROWB;
            char const* src = k.gbase + (size_t)src_row * ROWB;
            PE_SK_CHECK(src_row + rows <= k.n_rows, "sk_copy source", src_row, rows)
            PE_SK_CHECK(dst_row + rows <= PE_STREAM_STAGE_ROWS, "sk_copy stage", dst_row, rows)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(rows * ROWB), "r"(bar)
                         : "memory");
        }
    }
#else
    // ---- host (test) form: one lane at a time, eager copies, hazard checks -----------------------------------------
    struct sk_ctx
    {
        double* wl;            // workspace block of the group at this lane: row r at wl[r * GL]
        uint64_t GL;           // lanes per group (row stride in doubles)
        double* ring;          // this lane's private copy of the ring: [ns][stage_rows]
        uint32_t stage_rows;
        uint32_t ns_mask, ns_log;
        uint32_t seq0;
        uint32_t pn;
        int32_t fenced;
        uint32_t lane;
        uint32_t enm;
        double guard;
        double t;
        // checks
        uint32_t n_rows;
        uint64_t* store_seq;   // [n_rows]: sequence number of the last store to the row
        uint64_t seq_counter, fence_seq;
        int64_t* stage_tile;   // [ns]: tile whose rows the stage holds (-1 = none)
        int64_t* stage_read;   // [ns]: 1 = its rows were read (sk_ring_done)
        uint64_t errors;
    };
    inline jv sk_ld(sk_ctx const& k, uint32_t row)
    {
        jv x;
        x.v[0] = row < k.n_rows ? k.wl[(uint64_t)row * k.GL] : __builtin_nan("");
        return x;
    }
    inline void sk_st(sk_ctx& k, uint32_t row, jv const& x)
    {
        if(row >= k.n_rows)
        {
            ++k.errors;
            return;
        }
        k.store_seq[row] = ++k.seq_counter;
        if(k.enm & 1u) { k.wl[(uint64_t)row * k.GL] = x.v[0]; }
    }
    inline double const* sk_stage(sk_ctx& k, uint32_t t)
    {
        uint32_t const stg = (k.seq0 + t) & k.ns_mask;
        if(k.stage_tile[stg] != (int64_t)(k.seq0 + t)) { ++k.errors; }  // reading a stage that does not hold this tile
        return k.ring + (uint64_t)stg * k.stage_rows;
    }
    inline jv sk_ring(double const* sg, uint32_t row)
    {
        jv x;
        x.v[0] = sg[row];
        return x;
    }
    inline void sk_wait(sk_ctx& k, uint32_t t)
    {
        uint32_t const stg = (k.seq0 + t) & k.ns_mask;
        if(k.stage_tile[stg] != (int64_t)(k.seq0 + t)) { ++k.errors; }  // on the GPU: a wait that never returns
    }
    inline void sk_ring_done(sk_ctx& k, uint32_t t) { k.stage_read[(k.seq0 + t) & k.ns_mask] = 1; }
    inline void sk_fence(sk_ctx& k, int32_t t_done)
    {
        k.fence_seq = k.seq_counter;
        k.fenced = t_done;
    }
    inline void sk_tx(sk_ctx& k, uint32_t p, uint32_t rows)
    {
        uint32_t const stg = (k.seq0 + p) & k.ns_mask;
        if(k.stage_tile[stg] >= 0 && k.stage_read[stg] == 0) { ++k.errors; }  // the stage still holds unread rows
        if(rows > k.stage_rows) { ++k.errors; }
        k.stage_tile[stg] = (int64_t)(k.seq0 + p);
        k.stage_read[stg] = 0;
    }
    // end of a section: every tile was issued and consumed
    inline void sk_end(sk_ctx& k, uint32_t n_tiles)
    {
        if(k.pn != n_tiles) { ++k.errors; }
        for(uint32_t i = 0; i <= k.ns_mask; ++i)
        {
            if(k.stage_tile[i] >= 0 && k.stage_read[i] == 0) { ++k.errors; }
        }
    }
    inline void sk_copy(sk_ctx& k, uint32_t p, uint32_t dst_row, uint32_t src_row, uint32_t rows)
    {
        uint32_t const stg = (k.seq0 + p) & k.ns_mask;
        for(uint32_t r = 0; r < rows; ++r)
        {
            if(src_row + r >= k.n_rows || dst_row + r >= k.stage_rows)
            {
                ++k.errors;
                continue;
            }
            if(k.store_seq[src_row + r] > k.fence_seq) { ++k.errors; }  // the bulk copy may overtake the store
            k.ring[(uint64_t)stg * k.stage_rows + dst_row + r] = k.wl[(uint64_t)(src_row + r) * k.GL];
        }
    }
#endif
}  // namespace pe_stream
