// compiler.cpp — the symbolic phase: netlist -> per-mode batch programs for the sm_100a interpreter kernels.
//
// What it restates from the reference (all host-side, integer/graph work, once per netlist):
//   * prepare() numbering                                   circuits/circuit.h:481-540
//   * the per-model stamps of iterate_{dc,ac,tr,trop}_define  model/models/**  (cited at each stamp below)
//   * accumulate-vs-assign semantics of MNA::{G,B,C,D,I,E}_ref  circuits/MNA/mna.h:60-157
//   * gmin on the node diagonal                              circuits/circuit.h:1107-1110
// What replaces Eigen::SparseLU::analyzePattern + factorize (circuit.h:1516): a Markowitz ordering with threshold
// pivoting evaluated ONCE on nominal (lane-0) values, the resulting fill pattern, and a right-looking elimination
// schedule emitted as PE_OP_PIVOT / PE_OP_ELIM / PE_OP_BACK words.  Numeric factorisation itself never runs here.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>

#include "../csrc/pe_b200_models.h"
#include "pe_host.hpp"

namespace pe_b200
{
    namespace
    {
        using cplx = std::complex<double>;

        struct val
        {
            std::uint32_t op{};  // operand word without the negate bit
            double nom{};        // lane-0 / nominal value (pivot selection only)
            bool cst{};
        };

        struct entry
        {
            std::vector<std::uint32_t> re, im;
            cplx nom{};
        };

        constexpr double k_nl_nominal = 1e-12;  // nominal conductance of a not-yet-evaluated non-linear device
        constexpr double k_pivot_tau = 1e-3;    // relative threshold (SPICE PIVREL)

        struct elem_vals
        {
            val p[k_max_attr];
            val d[6];
            val s[6];
        };

        class builder
        {
        public:
            compile_input const& in;
            compiled& out;
            netlist const& nl;
            numbering const& num;
            std::map<std::uint64_t, int> cst_index;
            std::vector<elem_vals> ev;
            int n_inst{};
            int dt_slot{-1};

            builder(compile_input const& i, compiled& o) : in{i}, out{o}, nl{*i.nl}, num{o.num} {}

            val constant(double v)
            {
                std::uint64_t bits;
                std::memcpy(&bits, &v, 8);
                auto it{cst_index.find(bits)};
                int slot;
                if(it == cst_index.end())
                {
                    slot = static_cast<int>(out.cst.size());
                    out.cst.push_back(v);
                    cst_index.emplace(bits, slot);
                }
                else
                {
                    slot = it->second;
                }
                return {PE_OPND(PE_SP_CONST, slot), v, true};
            }

            val inst_slot(double nominal) { return {PE_OPND(PE_SP_INST, n_inst++), nominal, false}; }

            static void emit(std::vector<std::uint32_t>& sec, std::uint32_t opcode, std::initializer_list<std::uint32_t> ops, std::uint32_t n = 0)
            {
                sec.push_back(opcode | (n << 8));
                for(auto o: ops) { sec.push_back(o); }
            }

            int nidx(int node) const { return node < 0 ? -1 : num.node_index[static_cast<std::size_t>(node)]; }

            val vx(int node)
            {
                int const i{nidx(node)};
                if(i < 0) { return constant(0.0); }
                return {PE_OPND(PE_SP_INST, i), 0.0, false};
            }

            static bool connected(element const& e)
            {
                for(int p{}; p < e.d->pins; ++p)
                {
                    if(e.pin_node[p] == -2) { return false; }
                }
                return true;
            }

            // ---- layout + PREP section -------------------------------------------------------------------------
            void build_layout_and_prep()
            {
                n_inst = num.unknowns();  // x occupies INST slots [0, n)
                ev.resize(nl.elems.size());
                // swept parameters
                for(std::size_t ei{}; ei < nl.elems.size(); ++ei)
                {
                    auto const& e{nl.elems[ei]};
                    for(int a{}; a < e.d->n_attr; ++a)
                    {
                        auto const key{sweep_key{static_cast<int>(ei), a}};
                        auto it{in.swept_lane0.find(key)};
                        if(it != in.swept_lane0.end())
                        {
                            ev[ei].p[a] = inst_slot(it->second);
                            out.swept_slot[key] = static_cast<int>(PE_OPND_SLOT(ev[ei].p[a].op));
                        }
                        else
                        {
                            ev[ei].p[a] = constant(e.attr[a]);
                        }
                    }
                }
                // device state (always per instance)
                for(std::size_t ei{}; ei < nl.elems.size(); ++ei)
                {
                    auto const& e{nl.elems[ei]};
                    auto& v{ev[ei]};
                    switch(e.d->code)
                    {
                        case E_CAP:
                            v.s[0] = inst_slot(0.0);
                            v.s[1] = inst_slot(in.dt > 0.0 ? 2.0 * v.p[0].nom / in.dt : 0.0);
                            break;
                        case E_IND:
                            v.s[0] = inst_slot(in.dt > 0.0 ? 2.0 * v.p[0].nom / in.dt : 0.0);
                            v.s[1] = inst_slot(0.0);
                            break;
                        case E_VAC:
                        case E_IAC: v.s[0] = inst_slot(0.0); break;
                        case E_PN:
                            v.s[0] = inst_slot(0.0);
                            v.s[1] = inst_slot(k_nl_nominal);
                            v.s[2] = inst_slot(0.0);
                            v.s[3] = inst_slot(0.0);
                            v.s[4] = inst_slot(0.0);
                            break;
                        case E_NPN:
                        case E_PNP:
                            v.s[0] = inst_slot(k_nl_nominal);
                            v.s[1] = inst_slot(0.0);
                            v.s[2] = inst_slot(k_nl_nominal * std::fabs(v.p[2].nom));
                            v.s[3] = inst_slot(0.0);
                            break;
                        case E_NMOS:
                        case E_PMOS:
                            v.s[0] = inst_slot(k_nl_nominal);
                            v.s[1] = inst_slot(k_nl_nominal);
                            v.s[2] = inst_slot(0.0);
                            break;
                        default: break;
                    }
                }
                // derived quantities: folded on the host when every input is a broadcast constant, PREP ops otherwise
                auto& prep{out.prep};
                val const temp{constant(in.env.temperature)};  // load_temperature fallback overwrites Temp (base.h:326-381)
                for(std::size_t ei{}; ei < nl.elems.size(); ++ei)
                {
                    auto const& e{nl.elems[ei]};
                    auto& v{ev[ei]};
                    if(!connected(e)) { continue; }
                    switch(e.d->code)
                    {
                        case E_RES:
                        {
                            // m_G = 1.0 / r.r (resistance.h:88)
                            if(v.p[0].cst) { v.d[0] = constant(1.0 / v.p[0].nom); }
                            else
                            {
                                v.d[0] = inst_slot(1.0 / v.p[0].nom);
                                emit(prep, PE_OP_RECIP, {v.d[0].op, v.p[0].op});
                            }
                            break;
                        }
                        case E_VAC:
                        case E_IAC:
                        {
                            // m_E = Vp (cos phase + j sin phase) (VAC.h:115-121, IAC.h:115-121)
                            if(v.p[0].cst && v.p[2].cst)
                            {
                                v.d[0] = constant(v.p[0].nom * std::cos(v.p[2].nom));
                                v.d[1] = constant(v.p[0].nom * std::sin(v.p[2].nom));
                            }
                            else
                            {
                                v.d[0] = inst_slot(v.p[0].nom * std::cos(v.p[2].nom));
                                v.d[1] = inst_slot(v.p[0].nom * std::sin(v.p[2].nom));
                                emit(prep, PE_OP_SINCOS, {v.d[0].op, v.d[1].op, v.p[0].op, v.p[2].op});
                            }
                            break;
                        }
                        case E_SWITCH:
                        {
                            // r_contact = cut_through ? 0.0 : mna.r_open (switch.h:93, circuit.h:1012)
                            double const r_open{in.env.r_open > 0.0 ? in.env.r_open : 1e12};
                            v.d[0] = constant(v.p[0].nom != 0.0 ? 0.0 : r_open);
                            break;
                        }
                        case E_PN:
                        {
                            bool const all_c{v.p[0].cst && v.p[1].cst && v.p[2].cst && v.p[5].cst && v.p[6].cst && v.p[7].cst && v.p[8].cst};
                            auto const dn{
                                pe_models::pn_prepare(v.p[0].nom, v.p[2].nom, v.p[8].nom, v.p[1].nom, temp.nom, v.p[5].nom, v.p[6].nom, v.p[7].nom != 0.0)};
                            if(all_c)
                            {
                                v.d[0] = constant(dn.is_eff);
                                v.d[1] = constant(dn.isr_eff);
                                v.d[2] = constant(dn.bv_eff);
                                v.d[3] = constant(dn.ut);
                                v.d[4] = constant(dn.uth);
                            }
                            else
                            {
                                v.d[0] = inst_slot(dn.is_eff);
                                v.d[1] = inst_slot(dn.isr_eff);
                                v.d[2] = inst_slot(dn.bv_eff);
                                v.d[3] = inst_slot(dn.ut);
                                v.d[4] = inst_slot(dn.uth);
                                emit(prep,
                                     PE_OP_PN_PREP,
                                     {v.d[0].op,
                                      v.d[1].op,
                                      v.d[2].op,
                                      v.d[3].op,
                                      v.d[4].op,
                                      v.p[0].op,
                                      v.p[2].op,
                                      v.p[8].op,
                                      v.p[1].op,
                                      temp.op,
                                      v.p[5].op,
                                      v.p[6].op,
                                      v.p[7].op});
                            }
                            // Ud_last re-seeded from the node voltages on every prepare() (PN_junction.h:351 via base.h:373)
                            emit(prep, PE_OP_SUB, {v.s[0].op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op});
                            break;
                        }
                        case E_NPN:
                        case E_PNP:
                        {
                            v.d[0] = constant(pe_models::thermal_voltage(temp.nom));  // BJT_NPN.h:104 with Temp := env temperature
                            break;
                        }
                        default: break;
                    }
                }
                prep.push_back(PE_OP_END);
                dt_slot = static_cast<int>(out.cst.size());
                out.cst.push_back(in.dt);  // dedicated (never de-duplicated) so the driver can patch it per analyze
                out.n_inst_slots = n_inst;
            }
            val dt_val() const { return {PE_OPND(PE_SP_CONST, dt_slot), in.dt, true}; }

            // ---- one program ------------------------------------------------------------------------------------
            struct pstate
            {
                prog_mode mode;
                bool cplx;
                std::map<std::pair<int, int>, entry> A;
                std::vector<entry> Z;
                std::vector<std::uint32_t> head;  // per-iteration value ops (device evaluation, sources)
                std::vector<std::uint32_t> step;
                int n_lane{};
                val omega{};
            };

            val lane_slot(pstate& ps, double nominal) { return {PE_OPND(PE_SP_LANE, ps.n_lane++), nominal, false}; }

            static std::uint32_t sg(val const& v, bool neg) { return v.op | (neg ? PE_OPND_NEG : 0u); }

            void A_add(pstate& ps, int r, int c, val const& v, bool neg, bool imag = false)
            {
                if(r < 0 || c < 0) { return; }  // ground: mna.h:62 d_temp
                auto& e{ps.A[{r, c}]};
                (imag ? e.im : e.re).push_back(sg(v, neg));
                double const x{neg ? -v.nom : v.nom};
                e.nom += imag ? cplx{0.0, x} : cplx{x, 0.0};
            }

            void A_set(pstate& ps, int r, int c, val const* re, bool re_neg, val const* im = nullptr, bool im_neg = false)
            {
                if(r < 0 || c < 0) { return; }
                auto& e{ps.A[{r, c}]};
                e.re.clear();
                e.im.clear();
                e.nom = {};
                if(re)
                {
                    e.re.push_back(sg(*re, re_neg));
                    e.nom += cplx{re_neg ? -re->nom : re->nom, 0.0};
                }
                if(im)
                {
                    e.im.push_back(sg(*im, im_neg));
                    e.nom += cplx{0.0, im_neg ? -im->nom : im->nom};
                }
            }

            void Z_add(pstate& ps, int r, val const& v, bool neg, bool imag = false)
            {
                if(r < 0) { return; }
                auto& e{ps.Z[static_cast<std::size_t>(r)]};
                (imag ? e.im : e.re).push_back(sg(v, neg));
            }

            void Z_set(pstate& ps, int r, val const* re, val const* im = nullptr)
            {
                if(r < 0) { return; }
                auto& e{ps.Z[static_cast<std::size_t>(r)]};
                e.re.clear();
                e.im.clear();
                if(re) { e.re.push_back(sg(*re, false)); }
                if(im) { e.im.push_back(sg(*im, false)); }
            }

            // two-terminal conductance pattern used by R / C / diode / gds (e.g. resistance.h:101-104)
            void G4(pstate& ps, int a, int b, val const& g, bool imag = false)
            {
                A_add(ps, a, a, g, false, imag);
                A_add(ps, a, b, g, true, imag);
                A_add(ps, b, a, g, true, imag);
                A_add(ps, b, b, g, false, imag);
            }

            // ideal-source style incidence: B(a,k)=1, B(b,k)=-1, C(k,a)=1, C(k,b)=-1 (assign; VDC.h:89-92)
            void BC(pstate& ps, int a, int b, int k)
            {
                val const one{constant(1.0)};
                A_set(ps, a, k, &one, false);
                A_set(ps, b, k, &one, true);
                A_set(ps, k, a, &one, false);
                A_set(ps, k, b, &one, true);
            }

            void stamp_element(pstate& ps, std::size_t ei)
            {
                auto const& e{nl.elems[ei]};
                auto& v{ev[ei]};
                if(!connected(e)) { return; }  // unconnected pin: the whole stamp is skipped (resistance.h:86)
                auto const mode{ps.mode};
                int const n0{nidx(e.pin_node[0])};
                int const n1{nidx(e.pin_node[1])};
                int const n2{e.d->pins > 2 ? nidx(e.pin_node[2]) : -1};
                int const n3{e.d->pins > 3 ? nidx(e.pin_node[3]) : -1};
                int const k{num.n_nodes + num.branch0[ei]};
                switch(e.d->code)
                {
                    case E_RES:
                    {
                        // resistance.h:82-110 (the ground special-cases reduce to the same four accumulations)
                        if(n0 < 0 && n1 < 0) { break; }
                        G4(ps, n0, n1, v.d[0]);
                        break;
                    }
                    case E_CAP:
                    {
                        if(mode == prog_mode::AC)
                        {
                            // z = {0, C * omega}; G += z (capacitor.h:85-102)
                            val const z{lane_slot(ps, v.p[0].nom * ps.omega.nom)};
                            emit(ps.head, PE_OP_MUL, {z.op, v.p[0].op, ps.omega.op});
                            G4(ps, n0, n1, z, true);
                        }
                        else if(mode == prog_mode::TR)
                        {
                            // capacitor.h:106-155
                            emit(ps.step, PE_OP_CAP_STEP, {v.s[0].op, v.s[1].op, v.p[0].op, dt_val().op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op});
                            G4(ps, n0, n1, v.s[1]);
                            Z_add(ps, n0, v.s[0], true);
                            Z_add(ps, n1, v.s[0], false);
                        }
                        // DC / OP / TROP: open circuit, no stamp (capacitor.h:157-161; no iterate_dc_define)
                        break;
                    }
                    case E_IND:
                    {
                        BC(ps, n0, n1, k);  // inductor.h:83-99
                        if(mode == prog_mode::AC)
                        {
                            // D(k,k) = {0, -omega * L} unless L == 0 (inductor.h:118-126); omega == 0 lanes get -0.0
                            if(!(v.p[0].cst && v.p[0].nom == 0.0))
                            {
                                val const wl{lane_slot(ps, ps.omega.nom * v.p[0].nom)};
                                emit(ps.head, PE_OP_MUL, {wl.op, ps.omega.op, v.p[0].op});
                                A_set(ps, k, k, nullptr, false, &wl, true);
                            }
                        }
                        else if(mode == prog_mode::TR)
                        {
                            // inductor.h:134-195
                            emit(ps.step,
                                 PE_OP_IND_STEP,
                                 {v.s[0].op, v.s[1].op, v.p[0].op, dt_val().op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op, PE_OPND(PE_SP_INST, k)});
                            A_set(ps, k, k, &v.s[0], true);
                            Z_set(ps, k, &v.s[1]);
                        }
                        break;
                    }
                    case E_VDC:
                    {
                        BC(ps, n0, n1, k);
                        if(mode != prog_mode::AC) { Z_set(ps, k, &v.p[0]); }  // VDC.h:93 vs :100-116
                        break;
                    }
                    case E_VAC:
                    {
                        BC(ps, n0, n1, k);
                        if(mode == prog_mode::AC) { Z_set(ps, k, &v.d[0], &v.d[1]); }  // VAC.h:143-158
                        else if(mode == prog_mode::TR || mode == prog_mode::TROP)
                        {
                            // E = Vp sin(omega t + phase) (VAC.h:162-179); TROP evaluates at t = 0 (base.h:289-292)
                            emit(ps.head, PE_OP_VSIN, {v.s[0].op, v.p[0].op, v.p[1].op, v.p[2].op});
                            Z_set(ps, k, &v.s[0]);
                        }
                        break;
                    }
                    case E_IDC:
                    {
                        if(mode != prog_mode::AC)
                        {
                            Z_add(ps, n0, v.p[0], true);  // IDC.h:84-94
                            Z_add(ps, n1, v.p[0], false);
                        }
                        break;
                    }
                    case E_IAC:
                    {
                        if(mode == prog_mode::AC)
                        {
                            Z_add(ps, n0, v.d[0], true);  // IAC.h:133-143
                            Z_add(ps, n0, v.d[1], true, true);
                            Z_add(ps, n1, v.d[0], false);
                            Z_add(ps, n1, v.d[1], false, true);
                        }
                        else if(mode == prog_mode::TR || mode == prog_mode::TROP)
                        {
                            emit(ps.head, PE_OP_VSIN, {v.s[0].op, v.p[0].op, v.p[1].op, v.p[2].op});  // IAC.h:148-160
                            Z_add(ps, n0, v.s[0], true);
                            Z_add(ps, n1, v.s[0], false);
                        }
                        break;
                    }
                    case E_VCCS:
                    {
                        // pins S,T,P,Q (VCCS.h:80-95)
                        A_add(ps, n0, n2, v.p[0], false);
                        A_add(ps, n0, n3, v.p[0], true);
                        A_add(ps, n1, n2, v.p[0], true);
                        A_add(ps, n1, n3, v.p[0], false);
                        break;
                    }
                    case E_VCVS:
                    {
                        // VCVS.h:81-102
                        BC(ps, n0, n1, k);
                        A_set(ps, k, n2, &v.p[0], true);
                        A_set(ps, k, n3, &v.p[0], false);
                        break;
                    }
                    case E_CCCS:
                    {
                        // CCCS.h:81-100
                        val const one{constant(1.0)};
                        A_set(ps, n0, k, &v.p[0], false);
                        A_set(ps, n1, k, &v.p[0], true);
                        A_set(ps, n2, k, &one, false);
                        A_set(ps, n3, k, &one, true);
                        A_set(ps, k, n2, &one, false);
                        A_set(ps, k, n3, &one, true);
                        break;
                    }
                    case E_CCVS:
                    {
                        // CCVS.h:80-106
                        int const c2{k + 1};
                        val const one{constant(1.0)};
                        A_set(ps, n0, k, &one, false);
                        A_set(ps, n1, k, &one, true);
                        A_set(ps, n2, c2, &one, false);
                        A_set(ps, n3, c2, &one, true);
                        A_set(ps, k, n0, &one, false);
                        A_set(ps, k, n1, &one, true);
                        A_set(ps, c2, n2, &one, false);
                        A_set(ps, c2, n3, &one, true);
                        A_set(ps, k, c2, &v.p[0], true);
                        break;
                    }
                    case E_SWITCH:
                    {
                        BC(ps, n0, n1, k);  // switch.h:85-104
                        A_set(ps, k, k, &v.d[0], true);
                        break;
                    }
                    case E_OPAMP:
                    {
                        // pins +,-,OUT+,OUT- (op_amp.h:64-83): B/C assign on the outputs, C accumulates -/+mu on the inputs
                        val const one{constant(1.0)};
                        A_set(ps, n2, k, &one, false);
                        A_set(ps, n3, k, &one, true);
                        A_set(ps, k, n2, &one, false);
                        A_set(ps, k, n3, &one, true);
                        A_add(ps, k, n0, v.p[0], true);
                        A_add(ps, k, n1, v.p[0], false);
                        break;
                    }
                    case E_PN:
                    {
                        bool const tt_off{v.p[9].cst && !(v.p[9].nom > 0.0)};
                        if(mode == prog_mode::AC)
                        {
                            G4(ps, n0, n1, v.s[1]);  // PN_junction.h:413-416 (geq from the bias solve)
                            if(!tt_off)
                            {
                                val const y{lane_slot(ps, 0.0)};
                                emit(ps.head, PE_OP_PN_ACCAP, {y.op, v.s[1].op, v.p[9].op, ps.omega.op});
                                G4(ps, n0, n1, y, true);
                            }
                            break;
                        }
                        // PN_junction.h:358-402
                        emit(ps.head,
                             PE_OP_PN_EVAL,
                             {v.s[0].op,
                              v.s[1].op,
                              v.s[2].op,
                              vx(e.pin_node[0]).op,
                              vx(e.pin_node[1]).op,
                              v.d[0].op,
                              v.d[1].op,
                              v.d[2].op,
                              v.d[3].op,
                              v.d[4].op,
                              v.p[1].op,
                              v.p[3].op,
                              v.p[7].op});
                        G4(ps, n0, n1, v.s[1]);
                        Z_add(ps, n0, v.s[2], true);
                        Z_add(ps, n1, v.s[2], false);
                        if(mode == prog_mode::TR)
                        {
                            // PN_junction.h:440-503; the tr_prev_g != 0 guard is value-equivalent to stamping zeros
                            emit(ps.step,
                                 PE_OP_PN_STEP,
                                 {v.s[0].op, v.s[3].op, v.s[4].op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op, v.s[1].op, v.p[9].op, dt_val().op});
                            if(!tt_off)
                            {
                                G4(ps, n0, n1, v.s[4]);
                                Z_add(ps, n0, v.s[3], true);
                                Z_add(ps, n1, v.s[3], false);
                            }
                        }
                        break;
                    }
                    case E_NPN:
                    case E_PNP:
                    {
                        // pins B,C,E.  NPN: v = Vb - Ve (BJT_NPN.h:116-159); PNP: v = Ve - Vb (BJT_PNP.h:116-159)
                        bool const npn{e.d->code == E_NPN};
                        int const nb{n0}, nc{n1}, ne{n2};
                        if(mode != prog_mode::AC)
                        {
                            int const pp{npn ? e.pin_node[0] : e.pin_node[2]};
                            int const pm{npn ? e.pin_node[2] : e.pin_node[0]};
                            emit(ps.head,
                                 PE_OP_BJT_EVAL,
                                 {v.s[0].op, v.s[1].op, v.s[2].op, v.s[3].op, vx(pp).op, vx(pm).op, v.p[0].op, v.p[4].op, v.p[1].op, v.d[0].op, v.p[2].op});
                        }
                        if(npn)
                        {
                            G4(ps, nb, ne, v.s[0]);
                            if(mode != prog_mode::AC)
                            {
                                Z_add(ps, nb, v.s[1], true);
                                Z_add(ps, ne, v.s[1], false);
                            }
                            A_add(ps, nc, nb, v.s[2], false);
                            A_add(ps, nc, ne, v.s[2], true);
                            A_add(ps, ne, nb, v.s[2], true);
                            A_add(ps, ne, ne, v.s[2], false);
                            if(mode != prog_mode::AC)
                            {
                                Z_add(ps, nc, v.s[3], true);
                                Z_add(ps, ne, v.s[3], false);
                            }
                        }
                        else
                        {
                            G4(ps, ne, nb, v.s[0]);
                            if(mode != prog_mode::AC)
                            {
                                Z_add(ps, ne, v.s[1], true);
                                Z_add(ps, nb, v.s[1], false);
                            }
                            A_add(ps, ne, ne, v.s[2], false);
                            A_add(ps, ne, nb, v.s[2], true);
                            A_add(ps, nc, ne, v.s[2], true);
                            A_add(ps, nc, nb, v.s[2], false);
                            if(mode != prog_mode::AC)
                            {
                                Z_add(ps, ne, v.s[3], true);
                                Z_add(ps, nc, v.s[3], false);
                            }
                        }
                        break;
                    }
                    case E_NMOS:
                    case E_PMOS:
                    {
                        // pins D,G,S (nmosfet.h:84-141, pmosfet.h:84-138)
                        bool const nm{e.d->code == E_NMOS};
                        int const nd{n0}, ng{n1}, ns{n2};
                        if(mode != prog_mode::AC)
                        {
                            emit(ps.head,
                                 nm ? PE_OP_NMOS_EVAL : PE_OP_PMOS_EVAL,
                                 {v.s[0].op,
                                  v.s[1].op,
                                  v.s[2].op,
                                  vx(e.pin_node[0]).op,
                                  vx(e.pin_node[1]).op,
                                  vx(e.pin_node[2]).op,
                                  v.p[0].op,
                                  v.p[1].op,
                                  v.p[2].op});
                        }
                        G4(ps, nd, ns, v.s[1]);
                        if(nm)
                        {
                            A_add(ps, nd, ng, v.s[0], false);
                            A_add(ps, nd, ns, v.s[0], true);
                            A_add(ps, ns, ng, v.s[0], true);
                            A_add(ps, ns, ns, v.s[0], false);
                        }
                        else
                        {
                            A_add(ps, nd, ns, v.s[0], false);
                            A_add(ps, nd, ng, v.s[0], true);
                            A_add(ps, ns, ns, v.s[0], true);
                            A_add(ps, ns, ng, v.s[0], false);
                        }
                        if(mode != prog_mode::AC)
                        {
                            Z_add(ps, nd, v.s[2], true);
                            Z_add(ps, ns, v.s[2], false);
                        }
                        break;
                    }
                    default: break;
                }
            }

            // ---- symbolic LU + emission ------------------------------------------------------------------------
            struct lu_step
            {
                int r, c, piv;
                std::vector<int> lrows, l_ent;  // rows i with (i,c), entry ids
                std::vector<int> ucols, u_ent;  // cols j with (r,j), entry ids
                std::vector<int> t_ent;         // lrows.size() * ucols.size() target entry ids
                std::vector<char> t_new;        // fill created at this step
            };

            void build_program(prog_mode mode)
            {
                auto& pr{out.prog[static_cast<int>(mode)]};
                pstate ps;
                ps.mode = mode;
                ps.cplx = (mode == prog_mode::AC);
                int const n{num.unknowns()};
                ps.Z.resize(static_cast<std::size_t>(n));
                if(ps.cplx)
                {
                    ps.omega = lane_slot(ps, in.omega0);  // LANE slot 0
                    pr.omega_slot = 0;
                }
                for(std::size_t ei{}; ei < nl.elems.size(); ++ei) { stamp_element(ps, ei); }
                if(in.env.g_min != 0.0)
                {
                    val const gm{constant(in.env.g_min)};  // circuit.h:1107-1110
                    for(int i{}; i < num.n_nodes; ++i) { A_add(ps, i, i, gm, false); }
                }
                pr.cplx = ps.cplx;
                pr.nnz_a = ps.A.size();

                // --- Markowitz ordering with threshold pivoting on nominal values
                std::vector<int> ent_r, ent_c;
                std::vector<cplx> nv;
                std::vector<std::map<int, int>> rows(static_cast<std::size_t>(n));  // col -> id (all entries ever)
                std::vector<std::set<int>> rowcols(static_cast<std::size_t>(n)), colrows(static_cast<std::size_t>(n));
                for(auto const& [rc, e]: ps.A)
                {
                    int const id{static_cast<int>(nv.size())};
                    ent_r.push_back(rc.first);
                    ent_c.push_back(rc.second);
                    nv.push_back(e.nom);
                    rows[static_cast<std::size_t>(rc.first)][rc.second] = id;
                    rowcols[static_cast<std::size_t>(rc.first)].insert(rc.second);
                    colrows[static_cast<std::size_t>(rc.second)].insert(rc.first);
                }
                int const n_orig{static_cast<int>(nv.size())};
                std::set<std::pair<int, int>> rq, cq;
                for(int i{}; i < n; ++i)
                {
                    rq.insert({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                    cq.insert({static_cast<int>(colrows[static_cast<std::size_t>(i)].size()), i});
                }
                std::vector<char> bnz(static_cast<std::size_t>(n), 0);
                for(int i{}; i < n; ++i) { bnz[static_cast<std::size_t>(i)] = !(ps.Z[static_cast<std::size_t>(i)].re.empty() && ps.Z[static_cast<std::size_t>(i)].im.empty()); }
                std::vector<char> step_bnz;
                std::vector<lu_step> steps;
                steps.reserve(static_cast<std::size_t>(n));
                bool singular{false};

                auto colmax = [&](int c)
                {
                    double m{};
                    for(int i: colrows[static_cast<std::size_t>(c)]) { m = std::max(m, std::abs(nv[static_cast<std::size_t>(rows[static_cast<std::size_t>(i)][c])])); }
                    return m;
                };

                for(int k{}; k < n; ++k)
                {
                    long best_cost{std::numeric_limits<long>::max()};
                    int br{-1}, bc{-1};
                    double best_ratio{};
                    auto consider = [&](int r, int c, double cm)
                    {
                        double const mag{std::abs(nv[static_cast<std::size_t>(rows[static_cast<std::size_t>(r)][c])])};
                        if(!(mag > 0.0) || mag < k_pivot_tau * cm) { return; }
                        long const cost{static_cast<long>(rowcols[static_cast<std::size_t>(r)].size() - 1) * static_cast<long>(colrows[static_cast<std::size_t>(c)].size() - 1)};
                        double const ratio{mag / cm};
                        if(cost < best_cost || (cost == best_cost && ratio > 4.0 * best_ratio))
                        {
                            best_cost = cost;
                            br = r;
                            bc = c;
                            best_ratio = ratio;
                        }
                    };
                    auto itc{cq.begin()};
                    auto itr{rq.begin()};
                    int examined{};
                    while(itc != cq.end() || itr != rq.end())
                    {
                        bool const take_col{itr == rq.end() || (itc != cq.end() && itc->first <= itr->first)};
                        int const cnt{take_col ? itc->first : itr->first};
                        if(br >= 0 && best_cost <= static_cast<long>(cnt - 1) * static_cast<long>(cnt - 1)) { break; }
                        if(take_col)
                        {
                            int const c{itc->second};
                            ++itc;
                            if(cnt == 0) { continue; }
                            double const cm{colmax(c)};
                            for(int i: colrows[static_cast<std::size_t>(c)]) { consider(i, c, cm); }
                        }
                        else
                        {
                            int const r{itr->second};
                            ++itr;
                            if(cnt == 0) { continue; }
                            for(int j: rowcols[static_cast<std::size_t>(r)]) { consider(r, j, colmax(j)); }
                        }
                        if(++examined > 64 && br >= 0) { break; }
                    }
                    if(br < 0)
                    {
                        singular = true;
                        break;
                    }
                    lu_step st;
                    st.r = br;
                    st.c = bc;
                    st.piv = rows[static_cast<std::size_t>(br)][bc];
                    for(int i: colrows[static_cast<std::size_t>(bc)])
                    {
                        if(i != br)
                        {
                            st.lrows.push_back(i);
                            st.l_ent.push_back(rows[static_cast<std::size_t>(i)][bc]);
                        }
                    }
                    for(int j: rowcols[static_cast<std::size_t>(br)])
                    {
                        if(j != bc)
                        {
                            st.ucols.push_back(j);
                            st.u_ent.push_back(rows[static_cast<std::size_t>(br)][j]);
                        }
                    }
                    // remove pivot row / col from the active structure
                    for(int j: rowcols[static_cast<std::size_t>(br)])
                    {
                        cq.erase({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                        colrows[static_cast<std::size_t>(j)].erase(br);
                        if(j != bc) { cq.insert({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j}); }
                    }
                    for(int i: st.lrows)
                    {
                        rq.erase({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                        rowcols[static_cast<std::size_t>(i)].erase(bc);
                    }
                    rq.erase({static_cast<int>(rowcols[static_cast<std::size_t>(br)].size()), br});
                    rowcols[static_cast<std::size_t>(br)].clear();
                    colrows[static_cast<std::size_t>(bc)].clear();
                    // numeric (nominal) elimination + fill
                    cplx const pv{nv[static_cast<std::size_t>(st.piv)]};
                    st.t_ent.reserve(st.lrows.size() * st.ucols.size());
                    for(std::size_t a{}; a < st.lrows.size(); ++a)
                    {
                        int const i{st.lrows[a]};
                        cplx const l{nv[static_cast<std::size_t>(st.l_ent[a])] / pv};
                        for(std::size_t b{}; b < st.ucols.size(); ++b)
                        {
                            int const j{st.ucols[b]};
                            auto& rowi{rows[static_cast<std::size_t>(i)]};
                            auto it{rowi.find(j)};
                            int id;
                            bool fresh{false};
                            if(it == rowi.end())
                            {
                                id = static_cast<int>(nv.size());
                                nv.push_back({});
                                ent_r.push_back(i);
                                ent_c.push_back(j);
                                rowi[j] = id;
                                rowcols[static_cast<std::size_t>(i)].insert(j);
                                cq.erase({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                                colrows[static_cast<std::size_t>(j)].insert(i);
                                cq.insert({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                                fresh = true;
                            }
                            else
                            {
                                id = it->second;
                            }
                            nv[static_cast<std::size_t>(id)] -= l * nv[static_cast<std::size_t>(st.u_ent[b])];
                            st.t_ent.push_back(id);
                            st.t_new.push_back(fresh ? 1 : 0);
                        }
                        rq.insert({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                        if(bnz[static_cast<std::size_t>(br)]) { bnz[static_cast<std::size_t>(i)] = 1; }
                    }
                    step_bnz.push_back(bnz[static_cast<std::size_t>(br)]);
                    steps.push_back(std::move(st));
                }
                pr.structurally_singular = singular;
                pr.nnz_lu = nv.size();
                (void)n_orig;

                // --- emission
                auto& it_sec{pr.iter};
                it_sec = std::move(ps.head);
                pr.step = std::move(ps.step);
                pr.step.push_back(PE_OP_END);
                if(singular)
                {
                    it_sec.clear();
                    it_sec.push_back(0xffu);  // unknown opcode -> every lane reports PE_ST_SINGULAR
                    it_sec.push_back(PE_OP_END);
                    pr.n_lane_slots = ps.n_lane;
                    pr.x_opnd.assign(static_cast<std::size_t>(n), PE_OPND(PE_SP_CONST, 0));
                    pr.built = true;
                    return;
                }
                int const w{ps.cplx ? 2 : 1};
                std::vector<int> eslot(nv.size(), -1), bslot(static_cast<std::size_t>(n), -1);
                auto slot_of = [&](int id)
                {
                    if(eslot[static_cast<std::size_t>(id)] < 0)
                    {
                        eslot[static_cast<std::size_t>(id)] = ps.n_lane;
                        ps.n_lane += w;
                    }
                    return PE_OPND(PE_SP_LANE, eslot[static_cast<std::size_t>(id)]);
                };
                auto emit_asm = [&](std::uint32_t dst, std::vector<std::uint32_t> const& src)
                {
                    it_sec.push_back(PE_OP_ASM | (static_cast<std::uint32_t>(src.size()) << 8));
                    it_sec.push_back(dst);
                    for(auto s: src) { it_sec.push_back(s); }
                };
                std::vector<char> assembled(static_cast<std::size_t>(n), 0);
                auto assemble_row = [&](int r)
                {
                    if(assembled[static_cast<std::size_t>(r)]) { return; }
                    assembled[static_cast<std::size_t>(r)] = 1;
                    auto lo{ps.A.lower_bound({r, std::numeric_limits<int>::min()})};
                    for(auto it{lo}; it != ps.A.end() && it->first.first == r; ++it)
                    {
                        int const id{rows[static_cast<std::size_t>(r)][it->first.second]};
                        std::uint32_t const d{slot_of(id)};
                        emit_asm(d, it->second.re);
                        if(ps.cplx) { emit_asm(d + 1, it->second.im); }
                    }
                    bslot[static_cast<std::size_t>(r)] = ps.n_lane;
                    ps.n_lane += w;
                    std::uint32_t const bd{PE_OPND(PE_SP_LANE, bslot[static_cast<std::size_t>(r)])};
                    emit_asm(bd, ps.Z[static_cast<std::size_t>(r)].re);
                    if(ps.cplx) { emit_asm(bd + 1, ps.Z[static_cast<std::size_t>(r)].im); }
                };
                static std::vector<std::uint32_t> const k_none{};
                for(std::size_t k{}; k < steps.size(); ++k)
                {
                    auto const& st{steps[k]};
                    assemble_row(st.r);
                    for(int i: st.lrows) { assemble_row(i); }
                    for(std::size_t q{}; q < st.t_ent.size(); ++q)
                    {
                        if(st.t_new[q])
                        {
                            std::uint32_t const d{slot_of(st.t_ent[q])};
                            emit_asm(d, k_none);
                            if(ps.cplx) { emit_asm(d + 1, k_none); }
                        }
                    }
                    it_sec.push_back(PE_OP_PIVOT);
                    it_sec.push_back(slot_of(st.piv));
                    bool const rb{step_bnz[k] != 0};
                    for(std::size_t a{}; a < st.lrows.size(); ++a)
                    {
                        std::uint32_t const cnt{static_cast<std::uint32_t>(st.ucols.size() + (rb ? 1 : 0))};
                        it_sec.push_back(PE_OP_ELIM | (cnt << 8));
                        it_sec.push_back(slot_of(st.l_ent[a]));
                        for(std::size_t b{}; b < st.ucols.size(); ++b)
                        {
                            it_sec.push_back(slot_of(st.t_ent[a * st.ucols.size() + b]));
                            it_sec.push_back(slot_of(st.u_ent[b]));
                        }
                        if(rb)
                        {
                            it_sec.push_back(PE_OPND(PE_SP_LANE, bslot[static_cast<std::size_t>(st.lrows[a])]));
                            it_sec.push_back(PE_OPND(PE_SP_LANE, bslot[static_cast<std::size_t>(st.r)]));
                        }
                        pr.n_fma += cnt;
                    }
                }
                // solution operands
                pr.x_opnd.resize(static_cast<std::size_t>(n));
                if(ps.cplx)
                {
                    for(int j{}; j < n; ++j)
                    {
                        pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_LANE, ps.n_lane);
                        ps.n_lane += 2;
                    }
                }
                else
                {
                    for(int j{}; j < n; ++j) { pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_INST, j); }
                }
                for(std::size_t kk{steps.size()}; kk-- > 0;)
                {
                    auto const& st{steps[kk]};
                    std::uint32_t const cnt{static_cast<std::uint32_t>(st.ucols.size())};
                    bool const is_branch{st.c >= num.n_nodes};
                    it_sec.push_back(PE_OP_BACK | (cnt << 8) | (is_branch ? 0x80000000u : 0u));
                    it_sec.push_back(PE_OPND(PE_SP_LANE, bslot[static_cast<std::size_t>(st.r)]));
                    it_sec.push_back(slot_of(st.piv));
                    it_sec.push_back(pr.x_opnd[static_cast<std::size_t>(st.c)]);
                    for(std::size_t b{}; b < st.ucols.size(); ++b)
                    {
                        it_sec.push_back(slot_of(st.u_ent[b]));
                        it_sec.push_back(pr.x_opnd[static_cast<std::size_t>(st.ucols[b])]);
                    }
                    pr.n_fma += cnt;
                }
                it_sec.push_back(PE_OP_END);
                pr.n_lane_slots = ps.n_lane;
                pr.built = true;
            }
        };
    }  // namespace

    std::unique_ptr<compiled> compile_circuit(compile_input const& in)
    {
        auto out{std::make_unique<compiled>()};
        out->num = make_numbering(*in.nl);
        out->cst.push_back(0.0);  // CONST slot 0 = 0.0 (ground voltage)
        builder b{in, *out};
        b.cst_index.emplace(0, 0);
        b.build_layout_and_prep();
        for(int m{}; m < static_cast<int>(prog_mode::COUNT); ++m) { b.build_program(static_cast<prog_mode>(m)); }
        out->dt_slot = b.dt_slot;
        return out;
    }
}  // namespace pe_b200
