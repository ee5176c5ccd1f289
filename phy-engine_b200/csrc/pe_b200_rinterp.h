// pe_b200_rinterp.h — per-thread executor of the vector ops of a RESIDENT batch program (pe_b200_program.h).
//
// Included by pe_b200_kernels.cu (body of the sm_100a resident kernel: thread = one word stream x J instances of the
// CTA, every operand a shared-memory slot) and by the test-only emulator tests/emu/emu.cpp, which replays the same
// vector ops thread by thread on host memory.  Replaces, per instance, solve_once (circuit.h:987-1527): stamping,
// Eigen::SparseLU factorize + solve and the Newton test of circult::solve (circuit.h:923-948).
#pragma once
#include <math.h>
#include <stdint.h>

#include "pe_b200_interp.h"

namespace pe_rinterp
{
    using pe_interp::tol_t;

    struct rctx
    {
        double* ws;    // shared-memory workspace of the CTA, already offset to this thread's first instance
        uint64_t I;    // distance (in doubles) between consecutive slots: instances per CTA (shared) or the lane stride (HBM)
        uint32_t S;    // streams per instance = distance (in slots) between re and im of a complex value
        uint32_t C;    // streams (word columns) per warp
        uint32_t col;  // this thread's column
        uint32_t stream;  // this thread's stream: operand words are relative to it (row * S + ((column - stream) mod S))
        uint32_t js{1};   // distance (in doubles) between the J instances of this thread: 1 (shared memory), 32 (HBM form)
    };

    // absolute slot of a stream-relative operand word (S is a power of two)
    PE_HD uint32_t abs_slot(rctx const& c, uint32_t w)
    {
        uint32_t const sm = c.S - 1u;
        uint32_t const r = w & 0x7fffu;
        return (r & ~sm) | ((r + c.stream) & sm);
    }

    template <int J>
    struct vd
    {
        double v[J];
    };

    template <int J>
    PE_HD double* slot_ptr(rctx const& c, uint32_t slot)
    {
        return c.ws + (uint64_t)slot * c.I;
    }

    template <int J>
    PE_HD vd<J> ldv(rctx const& c, uint32_t slot)
    {
        vd<J> r;
        double const* p = slot_ptr<J>(c, slot);
#if defined(__CUDA_ARCH__)
        if constexpr(J == 2)
        {
            if(c.js == 1u)
            {
                double2 const t = *reinterpret_cast<double2 const*>(p);
                r.v[0] = t.x;
                r.v[1] = t.y;
                return r;
            }
        }
#endif
        for(int j = 0; j < J; ++j)
        {
            PE_TRACE_LD(p + j * c.js);
            r.v[j] = p[j * c.js];
        }
        return r;
    }

    // operand word: slot | neg << 15
    template <int J>
    PE_HD vd<J> ldo(rctx const& c, uint32_t w)
    {
        vd<J> r = ldv<J>(c, abs_slot(c, w));
        if(w & PE_R_NEG)
        {
            for(int j = 0; j < J; ++j) { r.v[j] = -r.v[j]; }
        }
        return r;
    }

    template <int J>
    PE_HD void stv(rctx const& c, uint32_t slot, vd<J> const& x, bool const (&en)[J])
    {
        double* p = slot_ptr<J>(c, slot);
#if defined(__CUDA_ARCH__)
        if constexpr(J == 2)
        {
            if(c.js == 1u && en[0] && en[1])
            {
                *reinterpret_cast<double2*>(p) = make_double2(x.v[0], x.v[1]);
                return;
            }
        }
#endif
        for(int j = 0; j < J; ++j)
        {
            if(en[j])
            {
                PE_TRACE_ST(p + j * c.js);
                p[j * c.js] = x.v[j];
            }
        }
    }

    enum
    {
        V_END = 0,
        V_BAR = 1,
        V_OK = 2,
        V_BAD = 3,
        V_SKIP = 4,
    };

    // Word reader of one warp's program (host form; the sm_100a kernel has a warp-cooperative one with the same
    // interface).  A program is two word sequences: the MAIN stream holds, per vector op, [h0][mask][uniform rows...]
    // (BAR / END are single words), the SIDE stream the per-column rows (C words each) in consumption order.
    struct host_reader
    {
        uint32_t const* p0;  // start of the main stream (128-byte line aligned)
        uint32_t const* p;   // main stream, at the current op
        uint32_t const* q;   // side stream, at the first per-column row of the current op
        uint32_t C, col;
        uint32_t cur{}, m{};
        uint32_t const* qc{};
        uint32_t const* p_next{};
        uint32_t const* q_next{};

        PE_HD uint32_t head() const { return p[0]; }
        PE_HD uint32_t open(uint32_t rows)
        {
            m = p[1];
            cur = 2;
            qc = q;
            #if defined(__CUDA_ARCH__)
            uint32_t const pcnt = (uint32_t)__popc(m);
#else
            uint32_t const pcnt = (uint32_t)__builtin_popcount(m);
#endif
            p_next = p + 2 + rows - pcnt;
            q_next = q + pcnt * C;
            return m;
        }
        PE_HD uint32_t next()
        {
            uint32_t w;
            if(m & 1u)
            {
                w = qc[col];
                qc += C;
            }
            else
            {
                w = p[cur++];
            }
            m >>= 1;
            return w;
        }
        // one stream per warp (C == 1): every 32-word line of the main stream starts with two prefetch bitmaps
        PE_HD void normalize()
        {
            if(C == 1u && ((p - p0) & 31) == 0) { p += 2; }
        }
        PE_HD void start()
        {
            p = p0;
            normalize();
        }
        PE_HD void close()
        {
            p = p_next;
            q = q_next;
            normalize();
        }
        PE_HD void bar()
        {
            p += 1;
            normalize();
        }
        PE_HD void skip()
        {
            p = p0 + (((p - p0) >> 5) + 1) * 32;
            normalize();
        }
    };

    // Execute the vector op the reader stands on, for this thread.  en[j]: stores of instance j are enabled (the lane
    // is still being solved); nconv / fail accumulate the Newton test and the pivot test.  Returns V_*; after V_OK the
    // caller closes the op on the reader (rd.close()), after V_BAR it steps over the barrier word (rd.bar()).
    // On the GPU every thread of the warp consumes every row of the op before any divergence (the reader is
    // warp-cooperative); an idle column then simply does not store.
    template <int J, class R, bool FUSED = true>
    PE_HD int rvop(R& rd, rctx const& c, double t, tol_t const& tol, bool const (&en)[J], bool check, bool (&nconv)[J], bool (&fail)[J], bool first_iter = true)
    {
        uint32_t const h = rd.head();
        uint32_t const op = h & 0xffu;
        if(op == PE_OP_END) { return V_END; }
        if(op == PE_OP_BAR) { return V_BAR; }
        if(op == PE_OP_SKIP) { return V_SKIP; }
        auto fetch = [&]() -> uint32_t { return rd.next(); };
        if(op == PE_OP_DOT)
        {
            uint32_t const na = (h >> 8) & 0x1fu, nb = (h >> 18) & 0x3fu;
            rd.open(2u + na + nb);
            uint32_t const ctl = fetch();
            uint32_t const scale = fetch();
            constexpr uint32_t MA = 3, MB = 4;
            vd<J> acc;
            for(int j = 0; j < J; ++j) { acc.v[j] = 0.0; }
            if(na <= MA && nb <= MB)
            {
                // every operand word of the op is requested before the first one is used
                uint32_t sw[MA], pw[MB];
#if defined(__CUDACC__)
#pragma unroll
#endif
                for(uint32_t r = 0; r < MA; ++r)
                {
                    if(r < na) { sw[r] = fetch(); }
                }
#if defined(__CUDACC__)
#pragma unroll
#endif
                for(uint32_t r = 0; r < MB; ++r)
                {
                    if(r < nb) { pw[r] = fetch(); }
                }
                if(!(ctl & PE_R_ACTIVE)) { return V_OK; }  // every row has been consumed: idle columns may leave
#if defined(__CUDACC__)
#pragma unroll
#endif
                for(uint32_t r = 0; r < MA; ++r)
                {
                    if(r < na)
                    {
                        vd<J> const s0 = ldo<J>(c, sw[r] & 0xffffu);
                        vd<J> const s1 = ldo<J>(c, sw[r] >> 16);
                        for(int j = 0; j < J; ++j) { acc.v[j] = PE_ADD(PE_ADD(acc.v[j], s0.v[j]), s1.v[j]); }
                    }
                }
#if defined(__CUDACC__)
#pragma unroll
#endif
                for(uint32_t r = 0; r < MB; ++r)
                {
                    if(r < nb)
                    {
                        vd<J> const a = ldo<J>(c, pw[r] & 0xffffu);  // pair operands carry a sign (aliased entries)
                        vd<J> const b = ldo<J>(c, pw[r] >> 16);
                        for(int j = 0; j < J; ++j) { acc.v[j] = fma(-a.v[j], b.v[j], acc.v[j]); }
                    }
                }
            }
            else
            {
                // long op: rows are consumed one by one by the whole warp; an idle column computes on its padding
                // operands and does not store (host: it skips the loads, the race detector watches real accesses only)
                bool const idle = !(ctl & PE_R_ACTIVE);
#if !defined(__CUDA_ARCH__)
                if(idle) { return V_OK; }
#endif
                for(uint32_t r = 0; r < na; ++r)
                {
                    uint32_t const w = fetch();
                    vd<J> const s0 = ldo<J>(c, w & 0xffffu);
                    vd<J> const s1 = ldo<J>(c, w >> 16);
                    for(int j = 0; j < J; ++j) { acc.v[j] = PE_ADD(PE_ADD(acc.v[j], s0.v[j]), s1.v[j]); }
                }
                for(uint32_t r = 0; r < nb; ++r)
                {
                    uint32_t const w = fetch();
                    vd<J> const a = ldo<J>(c, w & 0xffffu);
                    vd<J> const b = ldo<J>(c, w >> 16);
                    for(int j = 0; j < J; ++j) { acc.v[j] = fma(-a.v[j], b.v[j], acc.v[j]); }
                }
                if(idle) { return V_OK; }
            }
            uint32_t const flags = ctl >> 16;
            if(flags & PE_F_SCALE)
            {
                vd<J> const sc = ldv<J>(c, abs_slot(c, scale));
                for(int j = 0; j < J; ++j) { acc.v[j] = PE_MUL(acc.v[j], sc.v[j]); }
                if(flags & PE_F_GUARD)
                {
                    // an entry of L: out of bounds when the pivot is far smaller than this entry of its column (pe_b200_program.h)
                    for(int j = 0; j < J; ++j)
                    {
                        if(PE_GUARD_TRIP(fabs(acc.v[j]), tol.guard)) { fail[j] = true; }
                    }
                }
            }
            if(flags & PE_F_RECIP)
            {
                for(int j = 0; j < J; ++j)
                {
                    if(acc.v[j] == 0.0 || !isfinite(acc.v[j])) { fail[j] = true; }
                    acc.v[j] = PE_RCP(acc.v[j]);
                }
            }
            uint32_t const dst = abs_slot(c, ctl);
            if(check && (flags & (PE_F_CHECK_V | PE_F_CHECK_I)))
            {
                // circuit.h:923-948: |new - old| > abstol + reltol * max(|new|, |old|)  => not converged
                vd<J> const xo = ldv<J>(c, dst);
                bool const br = (flags & PE_F_CHECK_I) != 0u;
                double const at = br ? tol.i_abstol : tol.v_abstol, rt = br ? tol.i_reltol : tol.v_reltol;
                for(int j = 0; j < J; ++j)
                {
                    double const tl = at + rt * fmax(fabs(acc.v[j]), fabs(xo.v[j]));
                    if(fabs(acc.v[j] - xo.v[j]) > tl) { nconv[j] = true; }
                }
            }
            stv<J>(c, dst, acc, en);
            return V_OK;
        }
        if(op == PE_OP_CDOT)
        {
            uint32_t const nre = (h >> 8) & 0x1fu, nim = (h >> 13) & 0x1fu, nb = (h >> 18) & 0x3fu;
            rd.open(2u + nre + nim + nb);
            uint32_t const ctl = fetch();
            uint32_t const scale = fetch();
            bool const idle = !(ctl & PE_R_ACTIVE);
#if !defined(__CUDA_ARCH__)
            if(idle) { return V_OK; }
#endif
            vd<J> are, aim;
            for(int j = 0; j < J; ++j) { are.v[j] = aim.v[j] = 0.0; }
            for(uint32_t r = 0; r < nre; ++r)
            {
                uint32_t const w = fetch();
                vd<J> const s0 = ldo<J>(c, w & 0xffffu);
                vd<J> const s1 = ldo<J>(c, w >> 16);
                for(int j = 0; j < J; ++j) { are.v[j] = PE_ADD(PE_ADD(are.v[j], s0.v[j]), s1.v[j]); }
            }
            for(uint32_t r = 0; r < nim; ++r)
            {
                uint32_t const w = fetch();
                vd<J> const s0 = ldo<J>(c, w & 0xffffu);
                vd<J> const s1 = ldo<J>(c, w >> 16);
                for(int j = 0; j < J; ++j) { aim.v[j] = PE_ADD(PE_ADD(aim.v[j], s0.v[j]), s1.v[j]); }
            }
            for(uint32_t r = 0; r < nb; ++r)
            {
                uint32_t const w = fetch();
                uint32_t const sa = abs_slot(c, w), sb = abs_slot(c, w >> 16);
                vd<J> const ar = ldv<J>(c, sa), ai = ldv<J>(c, sa + c.S);
                vd<J> const br = ldv<J>(c, sb), bi = ldv<J>(c, sb + c.S);
                for(int j = 0; j < J; ++j)
                {
                    are.v[j] = fma(-ar.v[j], br.v[j], are.v[j]);
                    are.v[j] = fma(ai.v[j], bi.v[j], are.v[j]);
                    aim.v[j] = fma(-ar.v[j], bi.v[j], aim.v[j]);
                    aim.v[j] = fma(-ai.v[j], br.v[j], aim.v[j]);
                }
            }
            if(idle) { return V_OK; }
            uint32_t const flags = ctl >> 16;
            if(flags & PE_F_SCALE)
            {
                uint32_t const ss = abs_slot(c, scale);
                vd<J> const sr = ldv<J>(c, ss), si = ldv<J>(c, ss + c.S);
                for(int j = 0; j < J; ++j)
                {
                    double const nr = are.v[j] * sr.v[j] - aim.v[j] * si.v[j];
                    double const ni = are.v[j] * si.v[j] + aim.v[j] * sr.v[j];
                    are.v[j] = nr;
                    aim.v[j] = ni;
                }
                if(flags & PE_F_GUARD)
                {
                    for(int j = 0; j < J; ++j)
                    {
                        if(PE_GUARD_TRIP(fabs(are.v[j]) + fabs(aim.v[j]), tol.guard)) { fail[j] = true; }
                    }
                }
            }
            if(flags & PE_F_RECIP)
            {
                for(int j = 0; j < J; ++j)
                {
                    double const mm = are.v[j] * are.v[j] + aim.v[j] * aim.v[j];
                    if(!(mm > 0.0) || !isfinite(mm)) { fail[j] = true; }
                    double const sc = 1.0 / mm;
                    are.v[j] = are.v[j] * sc;
                    aim.v[j] = -aim.v[j] * sc;
                }
            }
            uint32_t const dst = abs_slot(c, ctl);
            stv<J>(c, dst, are, en);
            stv<J>(c, dst + c.S, aim, en);
            return V_OK;
        }
        if(FUSED && op == PE_OP_CROUT2)
        {
            // fused elimination step: six DOT slots, the same arithmetic as the six DOTs one after the other; the L
            // entries are scaled by the pivot reciprocal just computed (identical to the value stored for it)
            rd.open(19u);
            uint32_t wd[19];
#if defined(__CUDACC__)
#pragma unroll
#endif
            for(uint32_t r = 0; r < 19u; ++r) { wd[r] = fetch(); }
            vd<J> piv;
            for(int j = 0; j < J; ++j) { piv.v[j] = 0.0; }
#if defined(__CUDACC__)
#pragma unroll
#endif
            for(uint32_t q = 0; q < 6u; ++q)
            {
                uint32_t const base = q == 0u ? 0u : 4u + 3u * (q - 1u);
                uint32_t const nsr = q == 0u ? 2u : 1u;
                uint32_t const ctl = wd[base];
                if(!(ctl & PE_R_ACTIVE)) { continue; }
                vd<J> acc;
                for(int j = 0; j < J; ++j) { acc.v[j] = 0.0; }
                for(uint32_t r = 0; r < nsr; ++r)
                {
                    uint32_t const w = wd[base + 1u + r];
                    vd<J> const s0 = ldo<J>(c, w & 0xffffu);
                    vd<J> const s1 = ldo<J>(c, w >> 16);
                    for(int j = 0; j < J; ++j) { acc.v[j] = PE_ADD(PE_ADD(acc.v[j], s0.v[j]), s1.v[j]); }
                }
                uint32_t const flags = ctl >> 16;
                {
                    uint32_t const w = wd[base + 1u + nsr];
                    vd<J> const a = ldo<J>(c, w & 0xffffu);
                    vd<J> const b = ldo<J>(c, w >> 16);
                    for(int j = 0; j < J; ++j) { acc.v[j] = fma(-a.v[j], b.v[j], acc.v[j]); }
                }
                if(flags & PE_F_SCALE)
                {
                    for(int j = 0; j < J; ++j) { acc.v[j] = PE_MUL(acc.v[j], piv.v[j]); }
                    if(flags & PE_F_GUARD)
                    {
                        for(int j = 0; j < J; ++j)
                        {
                            if(PE_GUARD_TRIP(fabs(acc.v[j]), tol.guard)) { fail[j] = true; }
                        }
                    }
                }
                if(flags & PE_F_RECIP)
                {
                    for(int j = 0; j < J; ++j)
                    {
                        if(acc.v[j] == 0.0 || !isfinite(acc.v[j])) { fail[j] = true; }
                        acc.v[j] = PE_RCP(acc.v[j]);
                        piv.v[j] = acc.v[j];
                    }
                }
                stv<J>(c, abs_slot(c, ctl), acc, en);
            }
            return V_OK;
        }
        // ---- value ops: a = operand rows; row 0 carries the ACTIVE bit
        uint32_t const rows = (h >> 8) & 0x1fu;
        if(op < PE_OP_RECIP || op > PE_OP_PMOS_EVAL)
        {
            // unknown opcode (also what the compiler emits for a structurally singular system)
            for(int j = 0; j < J; ++j) { fail[j] = true; }
            return V_BAD;
        }
        rd.open(rows);
        // a time-step update folded into the iter section (header bit 13) runs in the first Newton iteration only
        // (its rows are consumed all the same: a reader that feeds per-column rows from a side stream only advances that
        // stream in next())
        if((h & 0x2000u) && !first_iter)
        {
            for(uint32_t r = 0; r < rows; ++r) { (void)fetch(); }
            return V_OK;
        }
        uint32_t ow[13];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for(uint32_t r = 0; r < 13u; ++r)
        {
            ow[r] = 0u;
            if(r < rows) { ow[r] = fetch(); }
        }
        if(!(ow[0] & PE_R_VACTIVE)) { return V_OK; }
#define PE_LDR(r) ldo<J>(c, ow[r] & 0xffffu)
#define PE_STR(r, val) stv<J>(c, abs_slot(c, ow[r]), (val), en)
        switch(op)
        {
            case PE_OP_RECIP:
            {
                vd<J> a = PE_LDR(1);
                for(int j = 0; j < J; ++j) { a.v[j] = PE_DIV(1.0, a.v[j]); }
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_MUL:
            {
                vd<J> a = PE_LDR(1);
                vd<J> const b = PE_LDR(2);
                for(int j = 0; j < J; ++j) { a.v[j] = PE_MUL(a.v[j], b.v[j]); }
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_SUB:
            {
                vd<J> a = PE_LDR(1);
                vd<J> const b = PE_LDR(2);
                for(int j = 0; j < J; ++j) { a.v[j] = PE_SUB(a.v[j], b.v[j]); }
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_COPY:
            {
                vd<J> const a = PE_LDR(1);
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_VSIN:
            {
                vd<J> vp = PE_LDR(1);
                vd<J> const om = PE_LDR(2), ph = PE_LDR(3);
                for(int j = 0; j < J; ++j) { vp.v[j] = PE_MUL(vp.v[j], sin(PE_ADD(PE_MUL(om.v[j], t), ph.v[j]))); }
                PE_STR(0, vp);
                return V_OK;
            }
            case PE_OP_SINCOS:
            {
                vd<J> const vp = PE_LDR(2), ph = PE_LDR(3);
                vd<J> re, im;
                for(int j = 0; j < J; ++j)
                {
                    re.v[j] = PE_MUL(vp.v[j], cos(ph.v[j]));
                    im.v[j] = PE_MUL(vp.v[j], sin(ph.v[j]));
                }
                PE_STR(0, re);
                PE_STR(1, im);
                return V_OK;
            }
            case PE_OP_MUL2DIV:
            {
                vd<J> a = PE_LDR(1);
                vd<J> const b = PE_LDR(2);
                for(int j = 0; j < J; ++j) { a.v[j] = PE_DIV(PE_MUL(2.0, a.v[j]), b.v[j]); }
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_CAP_STEP:
            {
                vd<J> hv = PE_LDR(0), gv = PE_LDR(1);
                vd<J> const C_ = PE_LDR(2), dt = PE_LDR(3), va = PE_LDR(4), vb = PE_LDR(5);
                for(int j = 0; j < J; ++j) { pe_models::cap_step(C_.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), hv.v[j], gv.v[j]); }
                PE_STR(0, hv);
                PE_STR(1, gv);
                return V_OK;
            }
            case PE_OP_KMUT:
            {
                vd<J> const k_ = PE_LDR(1), l1 = PE_LDR(2), l2 = PE_LDR(3);
                vd<J> m_;
                for(int j = 0; j < J; ++j) { m_.v[j] = pe_models::k_mutual(k_.v[j], l1.v[j], l2.v[j]); }
                PE_STR(0, m_);
                return V_OK;
            }
            case PE_OP_KIND_STEP:
            {
                vd<J> const la = PE_LDR(3), lb = PE_LDR(4), dt = PE_LDR(5), vp = PE_LDR(6), vn = PE_LDR(7), ia = PE_LDR(8), ib = PE_LDR(9);
                vd<J> ra, rb, ue;
                for(int j = 0; j < J; ++j) { pe_models::kind_step(la.v[j], lb.v[j], dt.v[j], PE_SUB(vp.v[j], vn.v[j]), ia.v[j], ib.v[j], ra.v[j], rb.v[j], ue.v[j]); }
                PE_STR(0, ra);
                PE_STR(1, rb);
                PE_STR(2, ue);
                return V_OK;
            }
            case PE_OP_GEN_EVAL:
            {
                vd<J> const kd = PE_LDR(1), ts = PE_LDR(2), vh = PE_LDR(3), vl = PE_LDR(4), fq = PE_LDR(5), du = PE_LDR(6), ph = PE_LDR(7), tr_ = PE_LDR(8), tf_ = PE_LDR(9);
                vd<J> o;
                for(int j = 0; j < J; ++j)
                {
                    o.v[j] = pe_models::gen_eval((int)kd.v[j], ts.v[j] != 0.0 ? t : 0.0, vh.v[j], vl.v[j], fq.v[j], du.v[j], ph.v[j], tr_.v[j], tf_.v[j]);
                }
                PE_STR(0, o);
                return V_OK;
            }
            case PE_OP_RELAY_EVAL:
            {
                vd<J> en_ = PE_LDR(0), rc;
                vd<J> const vp = PE_LDR(2), vn = PE_LDR(3), von = PE_LDR(4), voff = PE_LDR(5), ro = PE_LDR(6);
                for(int j = 0; j < J; ++j) { pe_models::relay_eval(vp.v[j], vn.v[j], von.v[j], voff.v[j], ro.v[j], en_.v[j], rc.v[j]); }
                PE_STR(0, en_);
                PE_STR(1, rc);
                return V_OK;
            }
            case PE_OP_IND_STEP:
            {
                vd<J> const L = PE_LDR(2), dt = PE_LDR(3), va = PE_LDR(4), vb = PE_LDR(5), ib = PE_LDR(6);
                vd<J> req, ueq;
                for(int j = 0; j < J; ++j) { pe_models::ind_step(L.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), ib.v[j], req.v[j], ueq.v[j]); }
                PE_STR(0, req);
                PE_STR(1, ueq);
                return V_OK;
            }
            case PE_OP_PN_PREP:
            {
                vd<J> const Is = PE_LDR(5), Isr = PE_LDR(6), Ar = PE_LDR(7), N = PE_LDR(8), Tc = PE_LDR(9), Ibv = PE_LDR(10), Bv = PE_LDR(11), bs = PE_LDR(12);
                vd<J> o0, o1, o2, o3, o4;
                for(int j = 0; j < J; ++j)
                {
                    auto const d = pe_models::pn_prepare(Is.v[j], Isr.v[j], Ar.v[j], N.v[j], Tc.v[j], Ibv.v[j], Bv.v[j], bs.v[j] != 0.0);
                    o0.v[j] = d.is_eff;
                    o1.v[j] = d.isr_eff;
                    o2.v[j] = d.bv_eff;
                    o3.v[j] = d.ut;
                    o4.v[j] = d.uth;
                }
                PE_STR(0, o0);
                PE_STR(1, o1);
                PE_STR(2, o2);
                PE_STR(3, o3);
                PE_STR(4, o4);
                return V_OK;
            }
            case PE_OP_PN_EVAL:
            {
                vd<J> const udl = PE_LDR(0), va = PE_LDR(3), vb = PE_LDR(4), ise = PE_LDR(5), isr = PE_LDR(6), bve = PE_LDR(7), ut = PE_LDR(8), uth = PE_LDR(9),
                            N = PE_LDR(10), Nr = PE_LDR(11), bs = PE_LDR(12);
                vd<J> o0, o1, o2;
                for(int j = 0; j < J; ++j)
                {
                    auto const o = pe_models::pn_eval(PE_SUB(va.v[j], vb.v[j]), udl.v[j], ise.v[j], isr.v[j], bve.v[j], ut.v[j], uth.v[j], N.v[j], Nr.v[j], bs.v[j] != 0.0);
                    o0.v[j] = o.ud;
                    o1.v[j] = o.geq;
                    o2.v[j] = o.ieq;
                }
                PE_STR(0, o0);
                PE_STR(1, o1);
                PE_STR(2, o2);
                return V_OK;
            }
            case PE_OP_PN_STEP:
            {
                vd<J> hv = PE_LDR(1), gv = PE_LDR(2);
                vd<J> const va = PE_LDR(3), vb = PE_LDR(4), geq = PE_LDR(5), tt = PE_LDR(6), dt = PE_LDR(7);
                vd<J> ud;
                for(int j = 0; j < J; ++j)
                {
                    ud.v[j] = PE_SUB(va.v[j], vb.v[j]);
                    pe_models::pn_step(ud.v[j], geq.v[j], tt.v[j], dt.v[j], hv.v[j], gv.v[j]);
                }
                PE_STR(0, ud);
                PE_STR(1, hv);
                PE_STR(2, gv);
                return V_OK;
            }
            case PE_OP_PN_ACCAP:
            {
                vd<J> g = PE_LDR(1);
                vd<J> const tt = PE_LDR(2), om = PE_LDR(3);
                for(int j = 0; j < J; ++j) { g.v[j] = pe_models::pn_ac_cap(g.v[j], tt.v[j], om.v[j]); }
                PE_STR(0, g);
                return V_OK;
            }
            case PE_OP_BJT_PREP:
            {
                vd<J> a = PE_LDR(1);
                for(int j = 0; j < J; ++j) { a.v[j] = pe_models::thermal_voltage(a.v[j]); }
                PE_STR(0, a);
                return V_OK;
            }
            case PE_OP_BJT_EVAL:
            {
                vd<J> const vp = PE_LDR(4), vm = PE_LDR(5), Is = PE_LDR(6), Ar = PE_LDR(7), N = PE_LDR(8), ut = PE_LDR(9), bf = PE_LDR(10);
                vd<J> o0, o1, o2, o3;
                for(int j = 0; j < J; ++j)
                {
                    auto const o = pe_models::bjt_eval(PE_SUB(vp.v[j], vm.v[j]), Is.v[j], Ar.v[j], N.v[j], ut.v[j], bf.v[j]);
                    o0.v[j] = o.geq;
                    o1.v[j] = o.ieq_be;
                    o2.v[j] = o.gm;
                    o3.v[j] = o.ieq_c;
                }
                PE_STR(0, o0);
                PE_STR(1, o1);
                PE_STR(2, o2);
                PE_STR(3, o3);
                return V_OK;
            }
            case PE_OP_NMOS_EVAL:
            case PE_OP_PMOS_EVAL:
            {
                vd<J> const vdd = PE_LDR(3), vg = PE_LDR(4), vs = PE_LDR(5), kp = PE_LDR(6), la = PE_LDR(7), vt = PE_LDR(8);
                vd<J> o0, o1, o2;
                for(int j = 0; j < J; ++j)
                {
                    auto const o = (op == PE_OP_NMOS_EVAL) ? pe_models::nmos_eval(vdd.v[j], vg.v[j], vs.v[j], kp.v[j], la.v[j], vt.v[j])
                                                           : pe_models::pmos_eval(vdd.v[j], vg.v[j], vs.v[j], kp.v[j], la.v[j], vt.v[j]);
                    o0.v[j] = o.gm;
                    o1.v[j] = o.gds;
                    o2.v[j] = o.ieq;
                }
                PE_STR(0, o0);
                PE_STR(1, o1);
                PE_STR(2, o2);
                return V_OK;
            }
            default: break;
        }
#undef PE_LDR
#undef PE_STR
        for(int j = 0; j < J; ++j) { fail[j] = true; }
        return V_BAD;
    }
}  // namespace pe_rinterp
