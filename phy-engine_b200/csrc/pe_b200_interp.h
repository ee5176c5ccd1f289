// pe_b200_interp.h — the per-thread interpreter of one warp stream of a batch program (pe_b200_program.h).
//
// Included by pe_b200_kernels.cu, where it is the body of the sm_100a solve kernel (one thread = one lane, one warp =
// 32 consecutive lanes, program words are warp-uniform so there is no divergence on opcodes), and by the test-only
// emulator tests/emu/emu.cpp, which replays programs lane by lane on the host to check the symbolic phase without
// a GPU (never part of the shipped library).
//
// Replaces, per lane, the reference's solve_once (circuit.h:987-1527): stamping (MNA::*_ref accumulate/assign,
// mna.h:60-157), Eigen::SparseLU factorize + solve, and the Newton test of circult::solve (circuit.h:923-948).
#pragma once
#include <math.h>
#include <stdint.h>

#include "pe_b200_models.h"
#include "pe_b200_program.h"

#if defined(__CUDA_ARCH__)
#define PE_LDW(p) __ldg(p)
#else
#define PE_LDW(p) (*(p))
#endif

// test-only access tracing (the emulator's race detector); compiled out of the product
#ifndef PE_TRACE_LD
#define PE_TRACE_LD(ptr)
#endif
#ifndef PE_TRACE_ST
#define PE_TRACE_ST(ptr)
#endif

namespace pe_interp
{
    struct ctx_t
    {
        double const* cst;
        double* wu;  // + lane
        double* wx;  // + inst
        int64_t LSu, LSx;
    };

    struct tol_t
    {
        double v_abstol, v_reltol, i_abstol, i_reltol;
        double guard;  // PE_F_GUARD threshold
    };

    enum
    {
        R_END = 0,
        R_BAR = 1,
        R_BAD = 2,
    };

    PE_HD double* uaddr(ctx_t const& c, uint32_t slot) { return c.wu + (int64_t)slot * c.LSu; }

    PE_HD double* opnd_addr(ctx_t const& c, uint32_t o)
    {
        uint32_t const sp = PE_OPND_SPACE(o);
        int64_t const slot = (int64_t)PE_OPND_SLOT(o);
        if(sp == PE_SP_U) { return c.wu + slot * c.LSu; }
        if(sp == PE_SP_INSTX) { return c.wx + slot * c.LSx; }
        return const_cast<double*>(c.cst) + slot;
    }

    PE_HD double ld(ctx_t const& c, uint32_t o)
    {
        double const* p = opnd_addr(c, o);
        PE_TRACE_LD(p);
        double const v = *p;
        return (o & PE_OPND_NEG) ? -v : v;
    }

    PE_HD double ldu(ctx_t const& c, uint32_t slot)
    {
        double const* p = uaddr(c, slot);
        PE_TRACE_LD(p);
        return *p;
    }

    PE_HD void st(ctx_t const& c, uint32_t o, double v, bool live)
    {
        double* p = opnd_addr(c, o);
        if(live)
        {
            PE_TRACE_ST(p);
            *p = v;
        }
    }

    // Interpret words from pc until END or BAR.  `live` predicates every store (finished / failed / padding lanes keep
    // executing the warp-uniform stream but leave memory alone), `check` enables the Newton test, `nconv` is set when an
    // unknown moved by more than its tolerance, `fail` on a zero / non-finite pivot.
    PE_HD int run_until(uint32_t const*& pc_io, ctx_t const& c, double t, tol_t const& tol, bool live, bool check, bool& nconv, bool& fail)
    {
        uint32_t const* pc = pc_io;
        for(;;)
        {
            uint32_t const h = PE_LDW(pc);
            uint32_t const op = h & 0xffu;
            switch(op)
            {
                case PE_OP_END: pc_io = pc; return R_END;
                case PE_OP_BAR: pc_io = pc + 1; return R_BAR;
                case PE_OP_DOT:
                {
                    uint32_t const flags = (h >> 8) & 0xffu;
                    uint32_t const nsrc = (h >> 16) & 0xffu;
                    uint32_t const npair = h >> 24;
                    uint32_t const dst = PE_LDW(pc + 1);
                    uint32_t const* p = pc + 2;
                    uint32_t scale = 0;
                    if(flags & PE_F_SCALE)
                    {
                        scale = PE_LDW(p);
                        ++p;
                    }
                    double acc = 0.0;
                    for(uint32_t i = 0; i < nsrc; ++i) { acc = PE_ADD(acc, ld(c, PE_LDW(p + i))); }
                    p += nsrc;
#if defined(__CUDACC__)
#pragma unroll 2
#endif
                    for(uint32_t i = 0; i < npair; ++i)
                    {
                        double const a = ldu(c, PE_LDW(p + 2 * i));
                        double const b = ldu(c, PE_LDW(p + 2 * i + 1));
                        acc = fma(-a, b, acc);
                    }
                    p += 2 * npair;
                    if(flags & PE_F_SCALE) { acc = PE_MUL(acc, ldu(c, scale)); }
                    if((flags & PE_F_GUARD) && PE_GUARD_TRIP(fabs(acc), tol.guard)) { fail = true; }  // an entry of L out of bounds
                    if(flags & PE_F_RECIP)
                    {
                        if(acc == 0.0 || !isfinite(acc)) { fail = true; }
                        acc = PE_RCP(acc);
                    }
                    double* pd = uaddr(c, dst);
                    if(check && (flags & (PE_F_CHECK_V | PE_F_CHECK_I)))
                    {
                        // circuit.h:923-948: |new - old| > abstol + reltol * max(|new|, |old|)  => not converged
                        PE_TRACE_LD(pd);
                        double const xo = *pd;
                        bool const br = (flags & PE_F_CHECK_I) != 0u;
                        double const tl = (br ? tol.i_abstol : tol.v_abstol) + (br ? tol.i_reltol : tol.v_reltol) * fmax(fabs(acc), fabs(xo));
                        if(fabs(acc - xo) > tl) { nconv = true; }
                    }
                    if(live)
                    {
                        PE_TRACE_ST(pd);
                        *pd = acc;
                    }
                    pc = p;
                    break;
                }
                case PE_OP_CDOT:
                {
                    uint32_t const flags = (h >> 8) & 0xffu;
                    uint32_t const npair = h >> 16;
                    uint32_t const w1 = PE_LDW(pc + 1);
                    uint32_t const nre = w1 & 0xffffu, nim = w1 >> 16;
                    uint32_t const dst = PE_LDW(pc + 2);
                    uint32_t const* p = pc + 3;
                    uint32_t scale = 0;
                    if(flags & PE_F_SCALE)
                    {
                        scale = PE_LDW(p);
                        ++p;
                    }
                    double are = 0.0, aim = 0.0;
                    for(uint32_t i = 0; i < nre; ++i) { are = PE_ADD(are, ld(c, PE_LDW(p + i))); }
                    p += nre;
                    for(uint32_t i = 0; i < nim; ++i) { aim = PE_ADD(aim, ld(c, PE_LDW(p + i))); }
                    p += nim;
#if defined(__CUDACC__)
#pragma unroll 2
#endif
                    for(uint32_t i = 0; i < npair; ++i)
                    {
                        uint32_t const sa = PE_LDW(p + 2 * i), sb = PE_LDW(p + 2 * i + 1);
                        double const ar = ldu(c, sa), ai = ldu(c, sa + 1);
                        double const br = ldu(c, sb), bi = ldu(c, sb + 1);
                        are = fma(-ar, br, are);
                        are = fma(ai, bi, are);
                        aim = fma(-ar, bi, aim);
                        aim = fma(-ai, br, aim);
                    }
                    p += 2 * npair;
                    if(flags & PE_F_SCALE)
                    {
                        double const sr = ldu(c, scale), si = ldu(c, scale + 1);
                        double const nr = are * sr - aim * si;
                        double const ni = are * si + aim * sr;
                        are = nr;
                        aim = ni;
                    }
                    if((flags & PE_F_GUARD) && PE_GUARD_TRIP(fabs(are) + fabs(aim), tol.guard)) { fail = true; }
                    if(flags & PE_F_RECIP)
                    {
                        double const m = are * are + aim * aim;
                        if(!(m > 0.0) || !isfinite(m)) { fail = true; }
                        double const s = 1.0 / m;
                        are = are * s;
                        aim = -aim * s;
                    }
                    if(live)
                    {
                        double* pd = uaddr(c, dst);
                        PE_TRACE_ST(pd);
                        PE_TRACE_ST(pd + c.LSu);
                        pd[0] = are;
                        pd[c.LSu] = aim;
                    }
                    pc = p;
                    break;
                }
                case PE_OP_RECIP:
                {
                    st(c, PE_LDW(pc + 1), PE_DIV(1.0, ld(c, PE_LDW(pc + 2))), live);
                    pc += 3;
                    break;
                }
                case PE_OP_MUL:
                {
                    st(c, PE_LDW(pc + 1), PE_MUL(ld(c, PE_LDW(pc + 2)), ld(c, PE_LDW(pc + 3))), live);
                    pc += 4;
                    break;
                }
                case PE_OP_SUB:
                {
                    st(c, PE_LDW(pc + 1), PE_SUB(ld(c, PE_LDW(pc + 2)), ld(c, PE_LDW(pc + 3))), live);
                    pc += 4;
                    break;
                }
                case PE_OP_COPY:
                {
                    st(c, PE_LDW(pc + 1), ld(c, PE_LDW(pc + 2)), live);
                    pc += 3;
                    break;
                }
                case PE_OP_VSIN:
                {
                    double const vp = ld(c, PE_LDW(pc + 2));
                    double const om = ld(c, PE_LDW(pc + 3));
                    double const ph = ld(c, PE_LDW(pc + 4));
                    st(c, PE_LDW(pc + 1), PE_MUL(vp, sin(PE_ADD(PE_MUL(om, t), ph))), live);
                    pc += 5;
                    break;
                }
                case PE_OP_SINCOS:
                {
                    double const vp = ld(c, PE_LDW(pc + 3));
                    double const ph = ld(c, PE_LDW(pc + 4));
                    st(c, PE_LDW(pc + 1), PE_MUL(vp, cos(ph)), live);
                    st(c, PE_LDW(pc + 2), PE_MUL(vp, sin(ph)), live);
                    pc += 5;
                    break;
                }
                case PE_OP_MUL2DIV:
                {
                    st(c, PE_LDW(pc + 1), PE_DIV(PE_MUL(2.0, ld(c, PE_LDW(pc + 2))), ld(c, PE_LDW(pc + 3))), live);
                    pc += 4;
                    break;
                }
                case PE_OP_KMUT:
                {
                    st(c, PE_LDW(pc + 1), pe_models::k_mutual(ld(c, PE_LDW(pc + 2)), ld(c, PE_LDW(pc + 3)), ld(c, PE_LDW(pc + 4))), live);
                    pc += 5;
                    break;
                }
                case PE_OP_KIND_STEP:
                {
                    double ra, rb, ue;
                    pe_models::kind_step(ld(c, PE_LDW(pc + 4)), ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)), PE_SUB(ld(c, PE_LDW(pc + 7)), ld(c, PE_LDW(pc + 8))), ld(c, PE_LDW(pc + 9)),
                                         ld(c, PE_LDW(pc + 10)), ra, rb, ue);
                    st(c, PE_LDW(pc + 1), ra, live);
                    st(c, PE_LDW(pc + 2), rb, live);
                    st(c, PE_LDW(pc + 3), ue, live);
                    pc += 11;
                    break;
                }
                case PE_OP_GEN_EVAL:
                {
                    double const tt = ld(c, PE_LDW(pc + 3)) != 0.0 ? t : 0.0;
                    st(c, PE_LDW(pc + 1),
                       pe_models::gen_eval((int)ld(c, PE_LDW(pc + 2)), tt, ld(c, PE_LDW(pc + 4)), ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)), ld(c, PE_LDW(pc + 7)), ld(c, PE_LDW(pc + 8)),
                                           ld(c, PE_LDW(pc + 9)), ld(c, PE_LDW(pc + 10))),
                       live);
                    pc += 11;
                    break;
                }
                case PE_OP_RELAY_EVAL:
                {
                    uint32_t const oe = PE_LDW(pc + 1), orc = PE_LDW(pc + 2);
                    double en_ = ld(c, oe), rc = 0.0;
                    pe_models::relay_eval(ld(c, PE_LDW(pc + 3)), ld(c, PE_LDW(pc + 4)), ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)), ld(c, PE_LDW(pc + 7)), en_, rc);
                    st(c, oe, en_, live);
                    st(c, orc, rc, live);
                    pc += 8;
                    break;
                }
                case PE_OP_CAP_STEP:
                {
                    uint32_t const oh = PE_LDW(pc + 1), og = PE_LDW(pc + 2);
                    double const C = ld(c, PE_LDW(pc + 3));
                    double const dt = ld(c, PE_LDW(pc + 4));
                    double const v = PE_SUB(ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)));
                    double hv = ld(c, oh), gv = ld(c, og);
                    pe_models::cap_step(C, dt, v, hv, gv);
                    st(c, oh, hv, live);
                    st(c, og, gv, live);
                    pc += 7;
                    break;
                }
                case PE_OP_IND_STEP:
                {
                    double const L = ld(c, PE_LDW(pc + 3));
                    double const dt = ld(c, PE_LDW(pc + 4));
                    double const v = PE_SUB(ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)));
                    double const ib = ld(c, PE_LDW(pc + 7));
                    double req, ueq;
                    pe_models::ind_step(L, dt, v, ib, req, ueq);
                    st(c, PE_LDW(pc + 1), req, live);
                    st(c, PE_LDW(pc + 2), ueq, live);
                    pc += 8;
                    break;
                }
                case PE_OP_PN_PREP:
                {
                    auto const d = pe_models::pn_prepare(ld(c, PE_LDW(pc + 6)),
                                                         ld(c, PE_LDW(pc + 7)),
                                                         ld(c, PE_LDW(pc + 8)),
                                                         ld(c, PE_LDW(pc + 9)),
                                                         ld(c, PE_LDW(pc + 10)),
                                                         ld(c, PE_LDW(pc + 11)),
                                                         ld(c, PE_LDW(pc + 12)),
                                                         ld(c, PE_LDW(pc + 13)) != 0.0);
                    st(c, PE_LDW(pc + 1), d.is_eff, live);
                    st(c, PE_LDW(pc + 2), d.isr_eff, live);
                    st(c, PE_LDW(pc + 3), d.bv_eff, live);
                    st(c, PE_LDW(pc + 4), d.ut, live);
                    st(c, PE_LDW(pc + 5), d.uth, live);
                    pc += 14;
                    break;
                }
                case PE_OP_PN_EVAL:
                {
                    uint32_t const oud = PE_LDW(pc + 1);
                    double const v = PE_SUB(ld(c, PE_LDW(pc + 4)), ld(c, PE_LDW(pc + 5)));
                    auto const o = pe_models::pn_eval(v,
                                                      ld(c, oud),
                                                      ld(c, PE_LDW(pc + 6)),
                                                      ld(c, PE_LDW(pc + 7)),
                                                      ld(c, PE_LDW(pc + 8)),
                                                      ld(c, PE_LDW(pc + 9)),
                                                      ld(c, PE_LDW(pc + 10)),
                                                      ld(c, PE_LDW(pc + 11)),
                                                      ld(c, PE_LDW(pc + 12)),
                                                      ld(c, PE_LDW(pc + 13)) != 0.0);
                    st(c, oud, o.ud, live);
                    st(c, PE_LDW(pc + 2), o.geq, live);
                    st(c, PE_LDW(pc + 3), o.ieq, live);
                    pc += 14;
                    break;
                }
                case PE_OP_PN_STEP:
                {
                    uint32_t const oh = PE_LDW(pc + 2), og = PE_LDW(pc + 3);
                    double const v = PE_SUB(ld(c, PE_LDW(pc + 4)), ld(c, PE_LDW(pc + 5)));
                    st(c, PE_LDW(pc + 1), v, live);
                    double hv = ld(c, oh), gv = ld(c, og);
                    pe_models::pn_step(v, ld(c, PE_LDW(pc + 6)), ld(c, PE_LDW(pc + 7)), ld(c, PE_LDW(pc + 8)), hv, gv);
                    st(c, oh, hv, live);
                    st(c, og, gv, live);
                    pc += 9;
                    break;
                }
                case PE_OP_PN_ACCAP:
                {
                    st(c, PE_LDW(pc + 1), pe_models::pn_ac_cap(ld(c, PE_LDW(pc + 2)), ld(c, PE_LDW(pc + 3)), ld(c, PE_LDW(pc + 4))), live);
                    pc += 5;
                    break;
                }
                case PE_OP_BJT_PREP:
                {
                    st(c, PE_LDW(pc + 1), pe_models::thermal_voltage(ld(c, PE_LDW(pc + 2))), live);
                    pc += 3;
                    break;
                }
                case PE_OP_BJT_EVAL:
                {
                    double const v = PE_SUB(ld(c, PE_LDW(pc + 5)), ld(c, PE_LDW(pc + 6)));
                    auto const o = pe_models::bjt_eval(v,
                                                       ld(c, PE_LDW(pc + 7)),
                                                       ld(c, PE_LDW(pc + 8)),
                                                       ld(c, PE_LDW(pc + 9)),
                                                       ld(c, PE_LDW(pc + 10)),
                                                       ld(c, PE_LDW(pc + 11)));
                    st(c, PE_LDW(pc + 1), o.geq, live);
                    st(c, PE_LDW(pc + 2), o.ieq_be, live);
                    st(c, PE_LDW(pc + 3), o.gm, live);
                    st(c, PE_LDW(pc + 4), o.ieq_c, live);
                    pc += 12;
                    break;
                }
                case PE_OP_NMOS_EVAL:
                case PE_OP_PMOS_EVAL:
                {
                    double const vd = ld(c, PE_LDW(pc + 4));
                    double const vg = ld(c, PE_LDW(pc + 5));
                    double const vs = ld(c, PE_LDW(pc + 6));
                    double const kp = ld(c, PE_LDW(pc + 7));
                    double const la = ld(c, PE_LDW(pc + 8));
                    double const vt = ld(c, PE_LDW(pc + 9));
                    auto const o = (op == PE_OP_NMOS_EVAL) ? pe_models::nmos_eval(vd, vg, vs, kp, la, vt) : pe_models::pmos_eval(vd, vg, vs, kp, la, vt);
                    st(c, PE_LDW(pc + 1), o.gm, live);
                    st(c, PE_LDW(pc + 2), o.gds, live);
                    st(c, PE_LDW(pc + 3), o.ieq, live);
                    pc += 10;
                    break;
                }
                default:
                {
                    // unknown opcode (also what the compiler emits for a structurally singular system): refuse to
                    // continue silently
                    fail = true;
                    pc_io = pc;
                    return R_BAD;
                }
            }
        }
    }
}  // namespace pe_interp
