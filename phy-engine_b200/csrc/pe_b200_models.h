// pe_b200_models.h — device-model arithmetic shared by the sm_100a kernels (per lane) and the host compiler
// (constant folding of broadcast parameters, nominal values for pivot selection).
//
// Every function restates the arithmetic of one reference model function, operation for operation and in the
// same association order, so that with identical inputs the stamped values are bit-identical wherever no libm
// call is involved.  Products and sums that the reference computes as separate IEEE operations are kept
// un-fused here (PE_MUL/PE_ADD/... expand to the __d*_rn intrinsics on the device, which the compiler never
// contracts into FMA; the host side is built with -ffp-contract=off).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PE_HD __host__ __device__ __forceinline__
#else
#define PE_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PE_MUL(a, b) __dmul_rn((a), (b))
#define PE_ADD(a, b) __dadd_rn((a), (b))
#define PE_SUB(a, b) __dsub_rn((a), (b))
#define PE_DIV(a, b) __ddiv_rn((a), (b))
// correctly rounded reciprocal: bit-identical to the IEEE quotient 1.0 / x, a third of its instructions
#define PE_RCP(x) __drcp_rn(x)
#else
#define PE_RCP(x) (1.0 / (x))
#define PE_MUL(a, b) ((a) * (b))
#define PE_ADD(a, b) ((a) + (b))
#define PE_SUB(a, b) ((a) - (b))
#define PE_DIV(a, b) ((a) / (b))
#endif

namespace pe_models
{
    // constants of PN_junction.h:303-306 / BJT_NPN.h:101-103
    constexpr double kKelvin = -273.15;
    constexpr double qElement = 1.6021765314e-19;
    constexpr double kBoltzmann = 1.380650524e-23;
    constexpr double sqrt2 = 1.4142135623730950488016887242096981;

    // Ut = k (T - kKelvin) / q        (PN_junction.h:308, BJT_NPN.h:104)
    PE_HD double thermal_voltage(double temp_c) { return PE_DIV(PE_MUL(kBoltzmann, PE_SUB(temp_c, kKelvin)), qElement); }

    // pn_details::limexp (PN_junction.h:10-16)
    PE_HD double limexp(double x)
    {
        if(x > 50.0) { return PE_MUL(exp(50.0), PE_ADD(1.0, PE_SUB(x, 50.0))); }
        if(x < -50.0) { return exp(-50.0); }
        return exp(x);
    }

    struct pn_derived
    {
        double is_eff, isr_eff, bv_eff, ut, uth;
    };

    // prepare_foundation_define(PN_junction) minus the Ud_last seeding (PN_junction.h:296-345)
    PE_HD pn_derived pn_prepare(double Is, double Isr, double Area, double N, double temp_c, double Ibv, double Bv, bool bv_set)
    {
        pn_derived d;
        d.is_eff = PE_MUL(Is, Area);
        d.isr_eff = PE_MUL(Isr, Area);
        d.bv_eff = Bv;
        d.ut = thermal_voltage(temp_c);
        double const nut = PE_MUL(N, d.ut);
        if(bv_set) { d.bv_eff = PE_SUB(Bv, PE_MUL(nut, log(PE_DIV(Ibv, d.is_eff)))); }
        d.uth = PE_MUL(nut, log(PE_DIV(nut, PE_MUL(sqrt2, d.is_eff))));
        return d;
    }

    // vlimit (PN_junction.h:58-109), SPICE3f5 junction limiting with breakdown mirror
    PE_HD double pn_vlimit(double Ud, double ud_last, double ute, double uth, double bv_eff, bool bv_set)
    {
        bool flag = false;
        double Ud_0, Ud_1, Ud_f;
        if(bv_set && Ud < fmin(0.0, PE_ADD(-bv_eff, PE_MUL(10.0, ute))))
        {
            Ud_0 = -PE_ADD(Ud, bv_eff);
            Ud_1 = -PE_ADD(ud_last, bv_eff);
            flag = true;
        }
        else
        {
            Ud_0 = Ud;
            Ud_1 = ud_last;
        }
        if(Ud_0 > uth && fabs(PE_SUB(Ud_0, Ud_1)) > PE_MUL(2.0, ute))
        {
            if(Ud_1 > 0)
            {
                double const arg = PE_DIV(PE_SUB(Ud_0, Ud_1), ute);
                if(arg > 0.0) { Ud_f = PE_ADD(Ud_1, PE_MUL(ute, PE_ADD(2.0, log(PE_SUB(arg, 2.0))))); }
                else
                {
                    Ud_f = PE_SUB(Ud_1, PE_MUL(ute, PE_ADD(2.0, log(PE_SUB(2.0, arg)))));
                }
            }
            else
            {
                Ud_f = PE_MUL(ute, log(PE_DIV(Ud_0, ute)));
            }
        }
        else
        {
            Ud_f = Ud_0;
            if(Ud_0 < 0.0)
            {
                double const arg = Ud_1 > 0.0 ? PE_SUB(-1.0, Ud_1) : PE_SUB(PE_MUL(2.0, Ud_1), 1.0);
                if(Ud_0 < arg) { Ud_f = arg; }
            }
        }
        if(flag) { return -PE_ADD(Ud_f, bv_eff); }
        return Ud_f;
    }

    struct pn_lin
    {
        double ud, geq, ieq;
    };

    // iterate_dc_define(PN_junction) (PN_junction.h:358-391): limited junction voltage, geq, Ieq
    PE_HD pn_lin pn_eval(double v_raw,
                         double ud_last,
                         double is_eff,
                         double isr_eff,
                         double bv_eff,
                         double ut,
                         double uth,
                         double N,
                         double Nr,
                         bool bv_set)
    {
        pn_lin o;
        double const ute = PE_MUL(N, ut);
        double const uter = PE_MUL(Nr, ut);
        double const Ud = pn_vlimit(v_raw, ud_last, ute, uth, bv_eff, bv_set);
        double Id;
        if(bv_set && Ud < -bv_eff)
        {
            double const e = limexp(PE_DIV(-PE_ADD(bv_eff, Ud), ute));
            Id = PE_MUL(-is_eff, e);
            o.geq = PE_DIV(PE_MUL(is_eff, e), ute);
        }
        else
        {
            double e = limexp(PE_DIV(Ud, ute));
            o.geq = PE_DIV(PE_MUL(is_eff, e), ute);
            Id = PE_MUL(is_eff, PE_SUB(e, 1.0));
            e = limexp(PE_DIV(Ud, uter));
            o.geq = PE_ADD(o.geq, PE_DIV(PE_MUL(isr_eff, e), uter));
            Id = PE_ADD(Id, PE_MUL(isr_eff, PE_SUB(e, 1.0)));
        }
        o.ud = Ud;
        o.ieq = PE_SUB(Id, PE_MUL(Ud, o.geq));
        return o;
    }

    // step_changed_tr_define(PN_junction) (PN_junction.h:440-476): diffusion-capacitance companion update
    PE_HD void pn_step(double v_prev, double geq, double tt, double dt, double& hist, double& prev_g)
    {
        if(!(dt > 0.0) || !(tt > 0.0) || !(geq > 0.0))
        {
            hist = 0.0;
            prev_g = 0.0;
            return;
        }
        double const cd = PE_MUL(tt, geq);
        if(!(cd > 0.0))
        {
            hist = 0.0;
            prev_g = 0.0;
            return;
        }
        double const g_new = PE_DIV(PE_MUL(2.0, cd), dt);
        hist = PE_SUB(PE_MUL(-PE_ADD(g_new, prev_g), v_prev), hist);
        prev_g = g_new;
    }

    // iterate_ac_define(PN_junction) diffusion susceptance (PN_junction.h:421-431): returns cd*omega or 0
    PE_HD double pn_ac_cap(double geq, double tt, double omega)
    {
        if(omega != 0.0 && tt > 0.0 && geq > 0.0)
        {
            double const cd = PE_MUL(tt, geq);
            if(cd > 0.0) { return PE_MUL(cd, omega); }
        }
        return 0.0;
    }

    struct bjt_lin
    {
        double geq, ieq_be, gm, ieq_c;
    };

    // iterate_dc_define(BJT_NPN) (BJT_NPN.h:116-146); PNP passes v = Ve - Vb (BJT_PNP.h:116-146).  Raw exp, no limiting.
    PE_HD bjt_lin bjt_eval(double v, double Is, double Area, double N, double ut, double BetaF)
    {
        bjt_lin o;
        double const is_eff = PE_MUL(Is, Area);
        double const ute = PE_MUL(N, ut);
        double const e = exp(PE_DIV(v, ute));
        o.geq = PE_DIV(PE_MUL(is_eff, e), ute);
        double const ibe = PE_MUL(is_eff, PE_SUB(e, 1.0));
        o.ieq_be = PE_SUB(ibe, PE_MUL(v, o.geq));
        o.gm = PE_MUL(BetaF, o.geq);
        double const ic = PE_MUL(BetaF, ibe);
        o.ieq_c = PE_SUB(ic, PE_MUL(o.gm, v));
        return o;
    }

    struct mos_lin
    {
        double gm, gds, ieq;
    };

    // iterate_dc_define(nmosfet) (nmosfet.h:84-124)
    PE_HD mos_lin nmos_eval(double Vd, double Vg, double Vs, double Kp, double lambda, double Vth)
    {
        mos_lin o;
        double const Vgs = PE_SUB(Vg, Vs);
        double const Vds = PE_SUB(Vd, Vs);
        double Id = 0.0;
        o.gm = 0.0;
        o.gds = 0.0;
        double const Vov = PE_SUB(Vgs, Vth);
        if(Vov <= 0.0) { Id = 0.0; }
        else if(Vds < Vov)
        {
            double const B = PE_SUB(PE_MUL(Vov, Vds), PE_MUL(PE_MUL(0.5, Vds), Vds));
            double const clm = PE_ADD(1.0, PE_MUL(lambda, Vds));
            Id = PE_MUL(PE_MUL(Kp, B), clm);
            o.gm = PE_MUL(PE_MUL(Kp, Vds), clm);
            o.gds = PE_MUL(Kp, PE_ADD(PE_MUL(PE_SUB(Vov, Vds), clm), PE_MUL(B, lambda)));
        }
        else
        {
            double const clm = PE_ADD(1.0, PE_MUL(lambda, Vds));
            double const hk = PE_MUL(PE_MUL(PE_MUL(0.5, Kp), Vov), Vov);
            Id = PE_MUL(hk, clm);
            o.gm = PE_MUL(PE_MUL(Kp, Vov), clm);
            o.gds = PE_MUL(hk, lambda);
        }
        o.ieq = PE_SUB(PE_SUB(Id, PE_MUL(o.gm, Vgs)), PE_MUL(o.gds, Vds));
        return o;
    }

    // iterate_dc_define(pmosfet) (pmosfet.h:84-121)
    PE_HD mos_lin pmos_eval(double Vd, double Vg, double Vs, double Kp, double lambda, double Vth)
    {
        mos_lin o;
        double const Vsg = PE_SUB(Vs, Vg);
        double const Vds = PE_SUB(Vd, Vs);
        double Id = 0.0;
        o.gm = 0.0;
        o.gds = 0.0;
        double const Vov = PE_SUB(Vsg, Vth);
        if(Vov <= 0.0) { Id = 0.0; }
        else if((-Vds) < Vov)
        {
            double const Vsd = -Vds;
            double const B = PE_SUB(PE_MUL(Vov, Vsd), PE_MUL(PE_MUL(0.5, Vsd), Vsd));
            double const clm = PE_ADD(1.0, PE_MUL(lambda, Vsd));
            double const Id_s = PE_MUL(PE_MUL(Kp, B), clm);
            Id = -Id_s;
            o.gm = PE_MUL(PE_MUL(Kp, Vsd), clm);
            double const dIddVsd = PE_MUL(Kp, PE_ADD(PE_MUL(PE_SUB(Vov, Vsd), clm), PE_MUL(B, lambda)));
            o.gds = -dIddVsd;
        }
        else
        {
            double const clm = PE_ADD(1.0, PE_MUL(lambda, -Vds));
            double const hk = PE_MUL(PE_MUL(PE_MUL(0.5, Kp), Vov), Vov);
            double const Id_s = PE_MUL(hk, clm);
            Id = -Id_s;
            o.gm = PE_MUL(PE_MUL(Kp, Vov), clm);
            o.gds = PE_MUL(hk, -lambda);
        }
        o.ieq = PE_SUB(PE_SUB(Id, PE_MUL(o.gm, Vsg)), PE_MUL(o.gds, Vds));
        return o;
    }

    // step_changed_tr_define(capacitor) (capacitor.h:106-128)
    PE_HD void cap_step(double C, double dt, double v_prev, double& hist, double& prev_g)
    {
        double const g_new = PE_DIV(PE_MUL(2.0, C), dt);
        hist = PE_SUB(PE_MUL(-PE_ADD(g_new, prev_g), v_prev), hist);
        prev_g = g_new;
    }

    // step_changed_tr_define(inductor) (inductor.h:134-160)
    PE_HD void ind_step(double L, double dt, double v_prev, double i_prev, double& req, double& ueq)
    {
        req = PE_DIV(PE_MUL(2.0, L), dt);
        ueq = PE_SUB(-v_prev, PE_MUL(req, i_prev));
    }

    // step_changed_tr_define(coupled_inductors), one winding (coupled_inductors.h:172-186)
    PE_HD void kind_step(double LA, double LB, double dt, double v_prev, double ia, double ib, double& rA, double& rB, double& ueq)
    {
        double const req_scale = PE_DIV(2.0, dt);
        rA = PE_MUL(req_scale, LA);
        rB = PE_MUL(req_scale, LB);
        ueq = PE_SUB(-v_prev, PE_ADD(PE_MUL(rA, ia), PE_MUL(rB, ib)));
    }

    // M = k sqrt(L1 L2) (coupled_inductors.h:135, :170); sqrt is correctly rounded on both sides
    PE_HD double k_mutual(double k, double L1, double L2) { return PE_MUL(k, sqrt(PE_MUL(L1, L2))); }

    // iterate_tr_define of the four generators (sawtooth.h:92-112, square.h:92-110, pulse.h:108-146, triangle.h:92-118):
    // t = fmod(tTime + phase / (2 pi) / freq, 1 / freq), then the wave shape; fmod is exact on both sides
    PE_HD double gen_eval(int kind, double t_time, double vh, double vl, double freq, double duty, double phase, double tr, double tf)
    {
        double const T = PE_DIV(1.0, freq);
        double const t0 = PE_ADD(t_time, PE_DIV(PE_DIV(phase, PE_MUL(2.0, 3.141592653589793)), freq));
        double const t = fmod(t0, T);
        if(kind == 1) { return t < PE_MUL(duty, T) ? vh : vl; }
        if(kind == 0) { return PE_ADD(vl, PE_MUL(PE_DIV(PE_SUB(vh, vl), T), t)); }
        if(kind == 3)
        {
            double const amp = PE_SUB(vh, vl);
            double const half = PE_MUL(0.5, T);
            double const s2 = PE_DIV(PE_MUL(2.0, amp), T);
            return t < half ? PE_ADD(vl, PE_MUL(s2, t)) : PE_SUB(vh, PE_MUL(s2, PE_SUB(t, half)));
        }
        double const ton = PE_MUL(duty, T);
        if(t < tr) { return PE_ADD(vl, PE_MUL(PE_DIV(PE_SUB(vh, vl), tr > 1e-30 ? tr : 1e-30), t)); }
        if(t < PE_SUB(ton, tf)) { return vh; }
        if(t < ton) { return PE_SUB(vh, PE_MUL(PE_DIV(PE_SUB(vh, vl), tf > 1e-30 ? tf : 1e-30), PE_SUB(t, PE_SUB(ton, tf)))); }
        return vl;
    }

    // iterate_dc_define(relay) (relay.h:81-94): the coil voltage of the previous solve moves the state, the contact is a
    // short (0) when engaged and mna.r_open otherwise
    PE_HD void relay_eval(double vcp, double vcn, double von, double voff, double r_open, double& engaged, double& r_contact)
    {
        double const vctrl = PE_SUB(vcp, vcn);
        if(engaged == 0.0)
        {
            if(vctrl >= von) { engaged = 1.0; }
        }
        else
        {
            if(vctrl <= voff) { engaged = 0.0; }
        }
        r_contact = engaged != 0.0 ? 0.0 : r_open;
    }
}  // namespace pe_models
