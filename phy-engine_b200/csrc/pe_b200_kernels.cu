// pe_b200_kernels.cu — sm_100a kernels for the batched MNA solve hot path + the POD device seam.
//
// One thread = one lane (independent circuit instance / frequency point).  All lanes of a warp interpret the
// same batch program word (warp-uniform fetch, no divergence on opcodes) while their data accesses hit
// lane-interleaved HBM arrays w[slot][lane] (32 consecutive doubles per warp request).  One launch runs a whole
// analysis phase per lane: [prep] -> for each time step { [step]; Newton loop { eval+assemble+LU+back-substitute
// with the convergence test fused into the back substitution } }.  Replaces, per lane, the reference's
// circult::solve / solve_once / update_tr_step (circuit.h:363-374, 892-1527) and Eigen::SparseLU compute+solve.
//
// This TU includes no reference header (nvcc ICEs on fast_io; SURVEY.md probe table).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <utility>
#include <vector>

#include "pe_b200_program.h"
#include "pe_b200_models.h"

namespace
{
    struct ctx_t
    {
        double const* cst;
        double* wi;
        double* wl;
        int64_t LSi, LSl;
        int64_t inst, lane;
    };

    __device__ __forceinline__ double* opnd_addr(ctx_t const& c, uint32_t o)
    {
        uint32_t const sp = PE_OPND_SPACE(o);
        int64_t const slot = (int64_t)PE_OPND_SLOT(o);
        if(sp == PE_SP_LANE) { return c.wl + slot * c.LSl + c.lane; }
        if(sp == PE_SP_INST) { return c.wi + slot * c.LSi + c.inst; }
        return const_cast<double*>(c.cst) + slot;
    }

    __device__ __forceinline__ int64_t opnd_stride(ctx_t const& c, uint32_t o)
    {
        uint32_t const sp = PE_OPND_SPACE(o);
        return sp == PE_SP_LANE ? c.LSl : (sp == PE_SP_INST ? c.LSi : 1);
    }

    __device__ __forceinline__ double ld(ctx_t const& c, uint32_t o)
    {
        double const v = *opnd_addr(c, o);
        return (o & PE_OPND_NEG) ? -v : v;
    }

    __device__ __forceinline__ void st(ctx_t const& c, uint32_t o, double v) { *opnd_addr(c, o) = v; }

    struct cplx
    {
        double re, im;
    };

    __device__ __forceinline__ cplx ldc(ctx_t const& c, uint32_t o)
    {
        double const* p = opnd_addr(c, o);
        cplx v;
        v.re = p[0];
        v.im = p[opnd_stride(c, o)];
        return v;
    }

    __device__ __forceinline__ void stc(ctx_t const& c, uint32_t o, cplx v)
    {
        double* p = opnd_addr(c, o);
        p[0] = v.re;
        p[opnd_stride(c, o)] = v.im;
    }

    struct tol_t
    {
        double v_abstol, v_reltol, i_abstol, i_reltol;
    };

    // Interpret one program section for this lane.  `conv` is and-ed with the per-unknown Newton test, `fail` is set
    // on a zero / non-finite pivot.
    template <bool CX>
    __device__ void run_section(uint32_t const* __restrict__ pc, ctx_t const& c, double t, tol_t const& tol, bool check, bool& conv, bool& fail)
    {
        double r_re = 0.0, r_im = 0.0;  // reciprocal of the current pivot
        for(;;)
        {
            uint32_t const h = __ldg(pc);
            uint32_t const op = h & 0xffu;
            uint32_t const n = (h >> 8) & 0x7fffffu;
            switch(op)
            {
                case PE_OP_END: return;
                case PE_OP_ASM:
                {
                    uint32_t const dst = __ldg(pc + 1);
                    double acc = 0.0;
                    for(uint32_t i = 0; i < n; ++i) { acc = PE_ADD(acc, ld(c, __ldg(pc + 2 + i))); }
                    st(c, dst, acc);
                    pc += 2 + n;
                    break;
                }
                case PE_OP_PIVOT:
                {
                    uint32_t const kk = __ldg(pc + 1);
                    if constexpr(CX)
                    {
                        cplx const p = ldc(c, kk);
                        double const m = p.re * p.re + p.im * p.im;
                        if(!(m > 0.0) || !isfinite(m)) { fail = true; }
                        double const s = 1.0 / m;
                        r_re = p.re * s;
                        r_im = -p.im * s;
                        stc(c, kk, cplx{r_re, r_im});
                    }
                    else
                    {
                        double const p = ld(c, kk);
                        if(p == 0.0 || !isfinite(p)) { fail = true; }
                        r_re = 1.0 / p;
                        st(c, kk, r_re);
                    }
                    pc += 2;
                    break;
                }
                case PE_OP_ELIM:
                {
                    uint32_t const ik = __ldg(pc + 1);
                    if constexpr(CX)
                    {
                        cplx const a = ldc(c, ik);
                        double const l_re = a.re * r_re - a.im * r_im;
                        double const l_im = a.re * r_im + a.im * r_re;
                        for(uint32_t i = 0; i < n; ++i)
                        {
                            uint32_t const ij = __ldg(pc + 2 + 2 * i);
                            uint32_t const kj = __ldg(pc + 3 + 2 * i);
                            cplx const u = ldc(c, kj);
                            cplx v = ldc(c, ij);
                            v.re -= l_re * u.re - l_im * u.im;
                            v.im -= l_re * u.im + l_im * u.re;
                            stc(c, ij, v);
                        }
                    }
                    else
                    {
                        double const l = ld(c, ik) * r_re;
                        for(uint32_t i = 0; i < n; ++i)
                        {
                            uint32_t const ij = __ldg(pc + 2 + 2 * i);
                            uint32_t const kj = __ldg(pc + 3 + 2 * i);
                            double* p = opnd_addr(c, ij);
                            *p = fma(-l, ld(c, kj), *p);
                        }
                    }
                    pc += 2 + 2 * n;
                    break;
                }
                case PE_OP_BACK:
                {
                    uint32_t const bk = __ldg(pc + 1);
                    uint32_t const kk = __ldg(pc + 2);
                    uint32_t const xk = __ldg(pc + 3);
                    if constexpr(CX)
                    {
                        cplx s = ldc(c, bk);
                        for(uint32_t i = 0; i < n; ++i)
                        {
                            cplx const u = ldc(c, __ldg(pc + 4 + 2 * i));
                            cplx const x = ldc(c, __ldg(pc + 5 + 2 * i));
                            s.re -= u.re * x.re - u.im * x.im;
                            s.im -= u.re * x.im + u.im * x.re;
                        }
                        cplx const r = ldc(c, kk);
                        stc(c, xk, cplx{s.re * r.re - s.im * r.im, s.re * r.im + s.im * r.re});
                    }
                    else
                    {
                        double s = ld(c, bk);
                        for(uint32_t i = 0; i < n; ++i) { s = fma(-ld(c, __ldg(pc + 4 + 2 * i)), ld(c, __ldg(pc + 5 + 2 * i)), s); }
                        double const xn = s * ld(c, kk);
                        double* px = opnd_addr(c, xk);
                        if(check)
                        {
                            // circuit.h:923-948: |new - old| > abstol + reltol * max(|new|, |old|)  => not converged
                            double const xo = *px;
                            bool const br = (h >> 31) != 0u;
                            double const tl = (br ? tol.i_abstol : tol.v_abstol) + (br ? tol.i_reltol : tol.v_reltol) * fmax(fabs(xn), fabs(xo));
                            if(fabs(xn - xo) > tl) { conv = false; }
                        }
                        *px = xn;
                    }
                    pc += 4 + 2 * n;
                    break;
                }
                case PE_OP_RECIP:
                {
                    st(c, __ldg(pc + 1), PE_DIV(1.0, ld(c, __ldg(pc + 2))));
                    pc += 3;
                    break;
                }
                case PE_OP_MUL:
                {
                    st(c, __ldg(pc + 1), PE_MUL(ld(c, __ldg(pc + 2)), ld(c, __ldg(pc + 3))));
                    pc += 4;
                    break;
                }
                case PE_OP_SUB:
                {
                    st(c, __ldg(pc + 1), PE_SUB(ld(c, __ldg(pc + 2)), ld(c, __ldg(pc + 3))));
                    pc += 4;
                    break;
                }
                case PE_OP_COPY:
                {
                    st(c, __ldg(pc + 1), ld(c, __ldg(pc + 2)));
                    pc += 3;
                    break;
                }
                case PE_OP_VSIN:
                {
                    double const vp = ld(c, __ldg(pc + 2));
                    double const om = ld(c, __ldg(pc + 3));
                    double const ph = ld(c, __ldg(pc + 4));
                    st(c, __ldg(pc + 1), PE_MUL(vp, sin(PE_ADD(PE_MUL(om, t), ph))));
                    pc += 5;
                    break;
                }
                case PE_OP_SINCOS:
                {
                    double const vp = ld(c, __ldg(pc + 3));
                    double const ph = ld(c, __ldg(pc + 4));
                    st(c, __ldg(pc + 1), PE_MUL(vp, cos(ph)));
                    st(c, __ldg(pc + 2), PE_MUL(vp, sin(ph)));
                    pc += 5;
                    break;
                }
                case PE_OP_MUL2DIV:
                {
                    st(c, __ldg(pc + 1), PE_DIV(PE_MUL(2.0, ld(c, __ldg(pc + 2))), ld(c, __ldg(pc + 3))));
                    pc += 4;
                    break;
                }
                case PE_OP_CAP_STEP:
                {
                    double* hist = opnd_addr(c, __ldg(pc + 1));
                    double* pg = opnd_addr(c, __ldg(pc + 2));
                    double const C = ld(c, __ldg(pc + 3));
                    double const dt = ld(c, __ldg(pc + 4));
                    double const v = PE_SUB(ld(c, __ldg(pc + 5)), ld(c, __ldg(pc + 6)));
                    double hv = *hist, gv = *pg;
                    pe_models::cap_step(C, dt, v, hv, gv);
                    *hist = hv;
                    *pg = gv;
                    pc += 7;
                    break;
                }
                case PE_OP_IND_STEP:
                {
                    double const L = ld(c, __ldg(pc + 3));
                    double const dt = ld(c, __ldg(pc + 4));
                    double const v = PE_SUB(ld(c, __ldg(pc + 5)), ld(c, __ldg(pc + 6)));
                    double const ib = ld(c, __ldg(pc + 7));
                    double req, ueq;
                    pe_models::ind_step(L, dt, v, ib, req, ueq);
                    st(c, __ldg(pc + 1), req);
                    st(c, __ldg(pc + 2), ueq);
                    pc += 8;
                    break;
                }
                case PE_OP_PN_PREP:
                {
                    auto const d = pe_models::pn_prepare(ld(c, __ldg(pc + 6)),
                                                         ld(c, __ldg(pc + 7)),
                                                         ld(c, __ldg(pc + 8)),
                                                         ld(c, __ldg(pc + 9)),
                                                         ld(c, __ldg(pc + 10)),
                                                         ld(c, __ldg(pc + 11)),
                                                         ld(c, __ldg(pc + 12)),
                                                         ld(c, __ldg(pc + 13)) != 0.0);
                    st(c, __ldg(pc + 1), d.is_eff);
                    st(c, __ldg(pc + 2), d.isr_eff);
                    st(c, __ldg(pc + 3), d.bv_eff);
                    st(c, __ldg(pc + 4), d.ut);
                    st(c, __ldg(pc + 5), d.uth);
                    pc += 14;
                    break;
                }
                case PE_OP_PN_EVAL:
                {
                    double* pud = opnd_addr(c, __ldg(pc + 1));
                    double const v = PE_SUB(ld(c, __ldg(pc + 4)), ld(c, __ldg(pc + 5)));
                    auto const o = pe_models::pn_eval(v,
                                                      *pud,
                                                      ld(c, __ldg(pc + 6)),
                                                      ld(c, __ldg(pc + 7)),
                                                      ld(c, __ldg(pc + 8)),
                                                      ld(c, __ldg(pc + 9)),
                                                      ld(c, __ldg(pc + 10)),
                                                      ld(c, __ldg(pc + 11)),
                                                      ld(c, __ldg(pc + 12)),
                                                      ld(c, __ldg(pc + 13)) != 0.0);
                    *pud = o.ud;
                    st(c, __ldg(pc + 2), o.geq);
                    st(c, __ldg(pc + 3), o.ieq);
                    pc += 14;
                    break;
                }
                case PE_OP_PN_STEP:
                {
                    double* hist = opnd_addr(c, __ldg(pc + 2));
                    double* pg = opnd_addr(c, __ldg(pc + 3));
                    double const v = PE_SUB(ld(c, __ldg(pc + 4)), ld(c, __ldg(pc + 5)));
                    st(c, __ldg(pc + 1), v);
                    double hv = *hist, gv = *pg;
                    pe_models::pn_step(v, ld(c, __ldg(pc + 6)), ld(c, __ldg(pc + 7)), ld(c, __ldg(pc + 8)), hv, gv);
                    *hist = hv;
                    *pg = gv;
                    pc += 9;
                    break;
                }
                case PE_OP_PN_ACCAP:
                {
                    st(c, __ldg(pc + 1), pe_models::pn_ac_cap(ld(c, __ldg(pc + 2)), ld(c, __ldg(pc + 3)), ld(c, __ldg(pc + 4))));
                    pc += 5;
                    break;
                }
                case PE_OP_BJT_PREP:
                {
                    st(c, __ldg(pc + 1), pe_models::thermal_voltage(ld(c, __ldg(pc + 2))));
                    pc += 3;
                    break;
                }
                case PE_OP_BJT_EVAL:
                {
                    double const v = PE_SUB(ld(c, __ldg(pc + 5)), ld(c, __ldg(pc + 6)));
                    auto const o =
                        pe_models::bjt_eval(v, ld(c, __ldg(pc + 7)), ld(c, __ldg(pc + 8)), ld(c, __ldg(pc + 9)), ld(c, __ldg(pc + 10)), ld(c, __ldg(pc + 11)));
                    st(c, __ldg(pc + 1), o.geq);
                    st(c, __ldg(pc + 2), o.ieq_be);
                    st(c, __ldg(pc + 3), o.gm);
                    st(c, __ldg(pc + 4), o.ieq_c);
                    pc += 12;
                    break;
                }
                case PE_OP_NMOS_EVAL:
                case PE_OP_PMOS_EVAL:
                {
                    double const vd = ld(c, __ldg(pc + 4));
                    double const vg = ld(c, __ldg(pc + 5));
                    double const vs = ld(c, __ldg(pc + 6));
                    double const kp = ld(c, __ldg(pc + 7));
                    double const la = ld(c, __ldg(pc + 8));
                    double const vt = ld(c, __ldg(pc + 9));
                    auto const o = (op == PE_OP_NMOS_EVAL) ? pe_models::nmos_eval(vd, vg, vs, kp, la, vt) : pe_models::pmos_eval(vd, vg, vs, kp, la, vt);
                    st(c, __ldg(pc + 1), o.gm);
                    st(c, __ldg(pc + 2), o.gds);
                    st(c, __ldg(pc + 3), o.ieq);
                    pc += 10;
                    break;
                }
                default:
                {
                    fail = true;  // unknown opcode: refuse to continue silently
                    return;
                }
            }
        }
    }

    template <bool CX>
    __global__ void __launch_bounds__(128) pe_b200_run_kernel(pe_b200_run const r)
    {
        int64_t const lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if(lane >= r.n_lanes) { return; }
        if(r.status[lane] != PE_ST_OK) { return; }

        ctx_t c;
        c.cst = r.cst;
        c.wi = r.wi;
        c.wl = r.wl;
        c.LSi = r.LSi;
        c.LSl = r.LSl;
        c.lane = lane;
        c.inst = lane / r.ppi;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol};

        bool conv = true, fail = false;
        uint32_t solves = 0;
        int32_t status = PE_ST_OK;
        double t = r.t0;

        if(r.prep != nullptr) { run_section<false>(r.prep, c, t, tol, false, conv, fail); }

        for(int32_t s = 0; s < r.n_steps && status == PE_ST_OK; ++s)
        {
            if(r.time_stepping)
            {
                // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                if(r.step != nullptr) { run_section<false>(r.step, c, t, tol, false, conv, fail); }
                t = t + r.dt;
            }
            int32_t it = 0;
            for(;;)
            {
                conv = true;
                run_section<CX>(r.iter, c, t, tol, r.nonlinear != 0, conv, fail);
                ++solves;
                if(fail)
                {
                    status = PE_ST_SINGULAR;
                    break;
                }
                if(!r.nonlinear || conv) { break; }
                if(++it >= r.max_iter)
                {
                    status = PE_ST_NO_CONVERGENCE;
                    break;
                }
            }
            if(r.wave != nullptr && status == PE_ST_OK)
            {
                for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSl + lane] = ld(c, __ldg(r.probes + p)); }
            }
        }
        r.status[lane] = status;
        r.solves[lane] += solves;
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
        // small batches: narrower blocks so that the grid still covers all 148 SMs
        int const block = (run->n_lanes <= 148 * 64) ? 32 : ((run->n_lanes <= 148 * 512) ? 64 : 128);
        int const grid = (run->n_lanes + block - 1) / block;
        cudaEvent_t e0{}, e1{};
        if(g_timing)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)stream);
        }
        if(run->cplx) { pe_b200_run_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(*run); }
        else
        {
            pe_b200_run_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(*run);
        }
        if(g_timing)
        {
            cudaEventRecord(e1, (cudaStream_t)stream);
            g_events.emplace_back(e0, e1);
        }
        g_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_run_kernel launch");
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
