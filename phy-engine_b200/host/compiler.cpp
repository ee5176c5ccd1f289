// compiler.cpp — the symbolic phase: netlist -> per-mode batch programs for the sm_100a solve kernel.
//
// What it restates from the reference (all host-side, integer/graph work, once per netlist):
//   * prepare() numbering                                   circuits/circuit.h:481-540
//   * the per-model stamps of iterate_{dc,ac,tr,trop}_define  model/models/**  (cited at each stamp below)
//   * accumulate-vs-assign semantics of MNA::{G,B,C,D,I,E}_ref  circuits/MNA/mna.h:60-157
//   * gmin on the node diagonal                              circuits/circuit.h:1107-1110
// What replaces Eigen::SparseLU::analyzePattern + factorize (circuit.h:1516):
//   1. a nested-dissection partition of the unknown graph into G-warp-sized leaves + a separator "top" (only when the
//      batch has too few lanes to fill the GPU with one warp per 32 lanes; G = 1 otherwise),
//   2. a Markowitz ordering with threshold pivoting, restricted to pivots inside one region at a time, evaluated ONCE
//      on nominal (lane-0) values, and the resulting fill pattern,
//   3. a row-wise ("dot product form") elimination schedule: every entry of L, U, the pivots, the forward-substituted
//      rhs and the solution is produced exactly once by one PE_OP_DOT, rows of different leaves on different warps,
//      separator rows after a CTA barrier.
// Numeric factorisation itself never runs here.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <queue>
#include <set>

#include "../csrc/pe_b200_models.h"
#include "pe_host.hpp"

namespace pe_b200
{
    namespace
    {
        using cplx = std::complex<double>;

        // internal operand spaces of the builder; translated to the kernel's spaces per mode at emission (xlat)
        constexpr std::uint32_t I_CONST{0}, I_INST{1}, I_LANE{2};
        constexpr std::uint32_t IOP(std::uint32_t space, int slot) { return (space << 29) | (static_cast<std::uint32_t>(slot) & 0x1fffffffu); }
        constexpr std::uint32_t IOP(std::uint32_t space, std::uint32_t slot) { return (space << 29) | (slot & 0x1fffffffu); }

        struct val
        {
            std::uint32_t op{};  // internal operand word without the negate bit
            double nom{};        // lane-0 / nominal value (pivot selection only)
            bool cst{};
        };

        struct entry
        {
            std::vector<std::uint32_t> re, im;
            cplx nom{};
        };

        struct op_t
        {
            std::vector<std::uint32_t> w;  // [opcode header][internal operand words...]
            int aff{-1};                   // unknown index the op's element hangs on (resident schedule: which stream runs it)
        };

        constexpr double k_nl_nominal = 1e-12;  // nominal conductance of a not-yet-evaluated non-linear device
        constexpr double k_pivot_tau = 1e-3;    // relative threshold (SPICE PIVREL)

        struct elem_vals
        {
            val p[k_max_attr];
            val d[6];
            val s[6];
            val bs[4][5];  // full bridge rectifier: state of its four internal junctions
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
            std::vector<op_t> prep_ops;
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
                return {IOP(I_CONST, slot), v, true};
            }

            val inst_slot(double nominal) { return {IOP(I_INST, n_inst++), nominal, false}; }

            int cur_aff{-1};
            void emit(std::vector<op_t>& sec, std::uint32_t opcode, std::initializer_list<std::uint32_t> ops)
            {
                op_t o;
                o.w.push_back(opcode);
                for(auto w: ops) { o.w.push_back(w); }
                o.aff = cur_aff;
                sec.push_back(std::move(o));
            }

            // the unknown an element's ops are scheduled next to: its first non-ground pin, else its first branch
            int affinity(std::size_t ei) const
            {
                auto const& e{nl.elems[ei]};
                for(int p{}; p < e.d->pins; ++p)
                {
                    if(e.pin_node[p] >= 0 && num.node_index[static_cast<std::size_t>(e.pin_node[p])] >= 0) { return num.node_index[static_cast<std::size_t>(e.pin_node[p])]; }
                }
                if(e.d->branches > 0) { return num.n_nodes + num.branch0[ei]; }
                return -1;
            }

            int nidx(int node) const { return node < 0 ? -1 : num.node_index[static_cast<std::size_t>(node)]; }

            val vx(int node)
            {
                int const i{nidx(node)};
                if(i < 0) { return constant(0.0); }
                return {IOP(I_INST, i), 0.0, false};
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
                double const nl_nom{in.nl_nominal > 0.0 ? in.nl_nominal : k_nl_nominal};
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
                        case E_IAC:
                        case E_GEN_SAW:
                        case E_GEN_SQUARE:
                        case E_GEN_PULSE:
                        case E_GEN_TRI: v.s[0] = inst_slot(0.0); break;
                        case E_KIND:
                            // req11, req12 (winding 1's copy), ueq1, req12 (winding 2's copy), req22, ueq2
                            for(int q{}; q < 6; ++q) { v.s[q] = inst_slot(0.0); }
                            break;
                        case E_RELAY:
                            // state: engaged (the "Engaged" attribute is its initial value), contact resistance; nominal for the
                            // pivot search = closed contact, so that the static order never pivots on D(k,k)
                            v.s[0] = inst_slot(v.p[2].nom != 0.0 ? 1.0 : 0.0);
                            v.s[1] = inst_slot(0.0);
                            break;
                        case E_PN:
                            v.s[0] = inst_slot(0.0);
                            v.s[1] = inst_slot(nl_nom);
                            v.s[2] = inst_slot(0.0);
                            v.s[3] = inst_slot(0.0);
                            v.s[4] = inst_slot(0.0);
                            break;
                        case E_BRIDGE:
                            for(auto& q: v.bs)
                            {
                                q[0] = inst_slot(0.0);
                                q[1] = inst_slot(nl_nom);
                                q[2] = inst_slot(0.0);
                                q[3] = inst_slot(0.0);
                                q[4] = inst_slot(0.0);
                            }
                            break;
                        case E_NPN:
                        case E_PNP:
                            v.s[0] = inst_slot(nl_nom);
                            v.s[1] = inst_slot(0.0);
                            v.s[2] = inst_slot(nl_nom * std::fabs(v.p[2].nom));
                            v.s[3] = inst_slot(0.0);
                            break;
                        case E_NMOS:
                        case E_PMOS:
                            v.s[0] = inst_slot(nl_nom);
                            v.s[1] = inst_slot(nl_nom);
                            v.s[2] = inst_slot(0.0);
                            break;
                        default: break;
                    }
                }
                // derived quantities: folded on the host when every input is a broadcast constant, PREP ops otherwise
                auto& prep{prep_ops};
                val const temp{constant(in.env.temperature)};  // load_temperature fallback overwrites Temp (base.h:326-381)
                for(std::size_t ei{}; ei < nl.elems.size(); ++ei)
                {
                    auto const& e{nl.elems[ei]};
                    auto& v{ev[ei]};
                    if(!connected(e)) { continue; }
                    cur_aff = affinity(ei);
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
                        case E_KIND:
                        {
                            // M = k sqrt(L1 L2) (coupled_inductors.h:135, :170)
                            if(v.p[0].cst && v.p[1].cst && v.p[2].cst) { v.d[0] = constant(pe_models::k_mutual(v.p[2].nom, v.p[0].nom, v.p[1].nom)); }
                            else
                            {
                                v.d[0] = inst_slot(pe_models::k_mutual(v.p[2].nom, v.p[0].nom, v.p[1].nom));
                                emit(prep, PE_OP_KMUT, {v.d[0].op, v.p[2].op, v.p[0].op, v.p[1].op});
                            }
                            break;
                        }
                        case E_XFMR_CT:
                        {
                            // n_half = 2 n_total (prepare_foundation, transformer_center_tap.h:72-76), invnh = 1 / n_half (:99, :109, :119)
                            if(v.p[0].cst) { v.d[0] = constant(1.0 / (2.0 * v.p[0].nom)); }
                            else
                            {
                                // one op (prep ops of an element do not depend on each other): 2 * 0.25 / n is the same real
                                // number as 1 / (2 n), both scalings by two are exact, so the quotient rounds identically
                                v.d[0] = inst_slot(1.0 / (2.0 * v.p[0].nom));
                                emit(prep, PE_OP_MUL2DIV, {v.d[0].op, constant(0.25).op, v.p[0].op});
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
                        case E_BRIDGE:
                        {
                            // the bridge exposes no "Temp" attribute, so the load_temperature fallback never touches its
                            // junctions: they keep the PN defaults, Temp = 27 C, and their Ud_last is seeded once, at the
                            // first prepare(), from the zero initial voltages (full_bridge_rectifier.h:28-52)
                            auto const* pd{find_desc(E_PN)};
                            auto const dn{pe_models::pn_prepare(pd->attr_default[0], pd->attr_default[2], pd->attr_default[8], pd->attr_default[1], pd->attr_default[4],
                                                                pd->attr_default[5], pd->attr_default[6], pd->attr_default[7] != 0.0)};
                            v.d[0] = constant(dn.is_eff);
                            v.d[1] = constant(dn.isr_eff);
                            v.d[2] = constant(dn.bv_eff);
                            v.d[3] = constant(dn.ut);
                            v.d[4] = constant(dn.uth);
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
                dt_slot = static_cast<int>(out.cst.size());
                out.cst.push_back(in.dt);  // dedicated (never de-duplicated) so the driver can patch it per analyze
                out.n_inst_slots = n_inst;
            }
            val dt_val() const { return {IOP(I_CONST, dt_slot), in.dt, true}; }

            // ---- one program ------------------------------------------------------------------------------------
            struct pstate
            {
                prog_mode mode;
                bool cplx;
                std::map<std::pair<int, int>, entry> A;
                std::vector<entry> Z;
                std::vector<op_t> head;  // per-iteration value ops (device evaluation, sources)
                std::vector<op_t> step;
                int n_lane{};
                val omega{};
            };

            val lane_slot(pstate& ps, double nominal) { return {IOP(I_LANE, ps.n_lane++), nominal, false}; }

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
                cur_aff = affinity(ei);
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
                                 {v.s[0].op, v.s[1].op, v.p[0].op, dt_val().op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op, IOP(I_INST, k)});
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
                    case E_GEN_SAW:
                    case E_GEN_SQUARE:
                    case E_GEN_PULSE:
                    case E_GEN_TRI:
                    {
                        // a voltage source whose value is a function of tr_duration (iterate_tr_define); iterate_dc evaluates it
                        // at t = 0 (DC, OP, TROP), AC drives 0 V (generator/*.h)
                        BC(ps, n0, n1, k);
                        if(mode != prog_mode::AC)
                        {
                            int const code{e.d->code};
                            double const kind{code == E_GEN_SAW ? 0.0 : (code == E_GEN_SQUARE ? 1.0 : (code == E_GEN_PULSE ? 2.0 : 3.0))};
                            bool const has_duty{code == E_GEN_SQUARE || code == E_GEN_PULSE};
                            val const zero{constant(0.0)};
                            val const& duty{has_duty ? v.p[3] : zero};
                            val const& phase{has_duty ? v.p[4] : v.p[3]};
                            val const& tr_{code == E_GEN_PULSE ? v.p[5] : zero};
                            val const& tf_{code == E_GEN_PULSE ? v.p[6] : zero};
                            emit(ps.head, PE_OP_GEN_EVAL,
                                 {v.s[0].op, constant(kind).op, constant(mode == prog_mode::TR ? 1.0 : 0.0).op, v.p[0].op, v.p[1].op, v.p[2].op, duty.op, phase.op, tr_.op, tf_.op});
                            Z_set(ps, k, &v.s[0]);
                        }
                        break;
                    }
                    case E_KIND:
                    {
                        // pins P1,P2,S1,S2, branches k1,k2 (coupled_inductors.h:92-246): both windings are B/C branches; DC / OP /
                        // TROP: shorts; TR: Thevenin companions with the 2x2 Req matrix; AC: -j omega [[L1 M],[M L2]]
                        int const k2{k + 1};
                        BC(ps, n0, n1, k);
                        BC(ps, n2, n3, k2);
                        if(mode == prog_mode::AC)
                        {
                            val const w1{lane_slot(ps, ps.omega.nom * v.p[0].nom)}, w2{lane_slot(ps, ps.omega.nom * v.p[1].nom)}, wm{lane_slot(ps, ps.omega.nom * v.d[0].nom)};
                            emit(ps.head, PE_OP_MUL, {w1.op, ps.omega.op, v.p[0].op});
                            emit(ps.head, PE_OP_MUL, {w2.op, ps.omega.op, v.p[1].op});
                            emit(ps.head, PE_OP_MUL, {wm.op, ps.omega.op, v.d[0].op});
                            A_set(ps, k, k, nullptr, false, &w1, true);
                            A_set(ps, k, k2, nullptr, false, &wm, true);
                            A_set(ps, k2, k, nullptr, false, &wm, true);
                            A_set(ps, k2, k2, nullptr, false, &w2, true);
                        }
                        else if(mode == prog_mode::TR)
                        {
                            std::uint32_t const i1{IOP(I_INST, k)}, i2{IOP(I_INST, k2)};
                            emit(ps.step, PE_OP_KIND_STEP,
                                 {v.s[0].op, v.s[1].op, v.s[2].op, v.p[0].op, v.d[0].op, dt_val().op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op, i1, i2});
                            emit(ps.step, PE_OP_KIND_STEP,
                                 {v.s[3].op, v.s[4].op, v.s[5].op, v.d[0].op, v.p[1].op, dt_val().op, vx(e.pin_node[2]).op, vx(e.pin_node[3]).op, i1, i2});
                            A_set(ps, k, k, &v.s[0], true);
                            A_set(ps, k, k2, &v.s[1], true);
                            A_set(ps, k2, k, &v.s[3], true);
                            A_set(ps, k2, k2, &v.s[4], true);
                            Z_set(ps, k, &v.s[2]);
                            Z_set(ps, k2, &v.s[5]);
                        }
                        break;
                    }
                    case E_RELAY:
                    {
                        // pins C+,C-,A,B (relay.h:74-105): the coil draws no current; the contact is a switch whose state follows
                        // the coil voltage of the previous solve (evaluated in every solve_once, AC included: iterate_ac falls
                        // back to iterate_dc)
                        double const r_open{in.env.r_open > 0.0 ? in.env.r_open : 1e12};
                        emit(ps.head, PE_OP_RELAY_EVAL, {v.s[0].op, v.s[1].op, vx(e.pin_node[0]).op, vx(e.pin_node[1]).op, v.p[0].op, v.p[1].op, constant(r_open).op});
                        BC(ps, n2, n3, k);
                        A_set(ps, k, k, &v.s[1], true);
                        break;
                    }
                    case E_XFMR:
                    {
                        // pins P,Q,S,T, branches kP,kS (transformer.h:66-99): B/C assign +-1 on both windings, the primary row
                        // accumulates -/+n on the secondary nodes (Vp - n Vs = 0), the secondary row is Is + n Ip = 0
                        int const kS{k + 1};
                        val const one{constant(1.0)};
                        A_set(ps, n0, k, &one, false);
                        A_set(ps, n1, k, &one, true);
                        A_set(ps, k, n0, &one, false);
                        A_set(ps, k, n1, &one, true);
                        A_set(ps, n2, kS, &one, false);
                        A_set(ps, n3, kS, &one, true);
                        A_add(ps, k, n2, v.p[0], true);
                        A_add(ps, k, n3, v.p[0], false);
                        A_set(ps, kS, kS, &one, false);
                        A_set(ps, kS, k, &v.p[0], false);
                        break;
                    }
                    case E_XFMR_CT:
                    {
                        // pins P,Q,S1,CT,S2, branches kP,kH1,kH2 (transformer_center_tap.h:80-125).  The coupling terms are
                        // stamped when the nominal n_half is non-zero (the reference tests n_half != 0.0 per instance; a
                        // swept ratio that reaches 0 is not supported).
                        int const n4{nidx(e.pin_node[4])};
                        int const kH1{k + 1}, kH2{k + 2};
                        val const one{constant(1.0)};
                        bool const coupled{v.p[0].nom != 0.0};
                        A_set(ps, n0, k, &one, false);
                        A_set(ps, n1, k, &one, true);
                        A_set(ps, n2, kH1, &one, false);
                        A_set(ps, n3, kH1, &one, true);
                        A_set(ps, n3, kH2, &one, false);
                        A_set(ps, n4, kH2, &one, true);
                        A_set(ps, kH1, n2, &one, false);
                        A_set(ps, kH1, n3, &one, true);
                        if(coupled)
                        {
                            A_add(ps, kH1, n0, v.d[0], true);
                            A_add(ps, kH1, n1, v.d[0], false);
                        }
                        A_set(ps, kH2, n3, &one, false);
                        A_set(ps, kH2, n4, &one, true);
                        if(coupled)
                        {
                            A_add(ps, kH2, n0, v.d[0], true);
                            A_add(ps, kH2, n1, v.d[0], false);
                            A_set(ps, k, k, &one, false);
                            A_set(ps, k, kH1, &v.d[0], false);
                            A_set(ps, k, kH2, &v.d[0], false);
                        }
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
                    case E_BRIDGE:
                    {
                        // pins A, B, +, -; D1: A->+, D2: B->+, D3: - -> A, D4: - -> B (full_bridge_rectifier.h:28-90).
                        // No iterate_tr_define: TR stamps like DC, the junction history is still stepped (:80-90).
                        auto const* pd{find_desc(E_PN)};
                        val const N{constant(pd->attr_default[1])}, Nr{constant(pd->attr_default[3])}, bvs{constant(pd->attr_default[7])}, tt{constant(pd->attr_default[9])};
                        int const pa[4]{0, 1, 3, 3}, pk[4]{2, 2, 0, 1};
                        for(int q{}; q < 4; ++q)
                        {
                            auto& bs{v.bs[q]};
                            int const na{nidx(e.pin_node[pa[q]])}, nk{nidx(e.pin_node[pk[q]])};
                            if(mode == prog_mode::AC)
                            {
                                G4(ps, na, nk, bs[1]);
                                continue;
                            }
                            emit(ps.head,
                                 PE_OP_PN_EVAL,
                                 {bs[0].op, bs[1].op, bs[2].op, vx(e.pin_node[pa[q]]).op, vx(e.pin_node[pk[q]]).op, v.d[0].op, v.d[1].op, v.d[2].op, v.d[3].op, v.d[4].op, N.op,
                                  Nr.op, bvs.op});
                            G4(ps, na, nk, bs[1]);
                            Z_add(ps, na, bs[2], true);
                            Z_add(ps, nk, bs[2], false);
                            if(mode == prog_mode::TR)
                            {
                                emit(ps.step, PE_OP_PN_STEP, {bs[0].op, bs[3].op, bs[4].op, vx(e.pin_node[pa[q]]).op, vx(e.pin_node[pk[q]]).op, bs[1].op, tt.op, dt_val().op});
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
            // ---- nested-dissection partition of the unknown graph ------------------------------------------------
            // region[i] >= 0: leaf id (rows/columns that may be pivoted inside that leaf, concurrently with the other
            // leaves); -1: the separator "top", eliminated after all leaves.  Equation i and unknown i share a region.
            static std::vector<int> partition(int n, std::vector<std::vector<int>> const& adj, int target_leaves, int min_leaf)
            {
                std::vector<int> region(static_cast<std::size_t>(n), -1);
                if(target_leaves <= 1 || n < 2 * min_leaf) { return region; }
                // hubs (supply rails, ...) go to the top straight away: they would glue every BFS level together
                std::size_t deg_sum{};
                for(auto const& a: adj) { deg_sum += a.size(); }
                std::size_t const hub_deg{std::max<std::size_t>(32, 10 * (deg_sum / static_cast<std::size_t>(n) + 1))};
                std::vector<char> in_top(static_cast<std::size_t>(n), 0);
                std::vector<int> all;
                for(int i{}; i < n; ++i)
                {
                    if(adj[static_cast<std::size_t>(i)].size() > hub_deg) { in_top[static_cast<std::size_t>(i)] = 1; }
                    else
                    {
                        all.push_back(i);
                    }
                }
                std::vector<std::vector<int>> parts;
                std::vector<char> frozen;  // parts that could not be split
                parts.push_back(std::move(all));
                frozen.push_back(0);
                std::vector<int> mark(static_cast<std::size_t>(n), -1), level(static_cast<std::size_t>(n), -1);
                int stamp{};

                auto bfs = [&](std::vector<int> const& part, int start, int my_stamp, std::vector<int>& order) -> int
                {
                    // BFS inside the vertices whose mark == my_stamp; returns the number of levels
                    order.clear();
                    for(int v: part) { level[static_cast<std::size_t>(v)] = -1; }
                    level[static_cast<std::size_t>(start)] = 0;
                    order.push_back(start);
                    int maxl{};
                    for(std::size_t h{}; h < order.size(); ++h)
                    {
                        int const v{order[h]};
                        for(int u: adj[static_cast<std::size_t>(v)])
                        {
                            if(mark[static_cast<std::size_t>(u)] != my_stamp || level[static_cast<std::size_t>(u)] >= 0) { continue; }
                            level[static_cast<std::size_t>(u)] = level[static_cast<std::size_t>(v)] + 1;
                            maxl = std::max(maxl, level[static_cast<std::size_t>(u)]);
                            order.push_back(u);
                        }
                    }
                    return maxl + 1;
                };

                while(static_cast<int>(parts.size()) < target_leaves)
                {
                    int best{-1};
                    for(std::size_t p{}; p < parts.size(); ++p)
                    {
                        if(frozen[p] || static_cast<int>(parts[p].size()) < 2 * min_leaf) { continue; }
                        if(best < 0 || parts[p].size() > parts[static_cast<std::size_t>(best)].size()) { best = static_cast<int>(p); }
                    }
                    if(best < 0) { break; }
                    auto part{std::move(parts[static_cast<std::size_t>(best)])};
                    int const st{++stamp};
                    for(int v: part) { mark[static_cast<std::size_t>(v)] = st; }
                    std::vector<int> order;
                    // connected components first
                    std::vector<std::vector<int>> comps;
                    {
                        std::vector<char> seen_local;
                        for(int v: part) { level[static_cast<std::size_t>(v)] = -1; }
                        std::vector<int> q;
                        for(int s: part)
                        {
                            if(level[static_cast<std::size_t>(s)] >= 0) { continue; }
                            q.clear();
                            q.push_back(s);
                            level[static_cast<std::size_t>(s)] = 0;
                            for(std::size_t h{}; h < q.size(); ++h)
                            {
                                for(int u: adj[static_cast<std::size_t>(q[h])])
                                {
                                    if(mark[static_cast<std::size_t>(u)] != st || level[static_cast<std::size_t>(u)] >= 0) { continue; }
                                    level[static_cast<std::size_t>(u)] = 0;
                                    q.push_back(u);
                                }
                            }
                            comps.push_back(q);
                        }
                    }
                    std::vector<int> A, B, S;
                    if(comps.size() > 1)
                    {
                        std::sort(comps.begin(), comps.end(), [](auto const& x, auto const& y) { return x.size() > y.size(); });
                        for(auto const& cmp: comps)
                        {
                            auto& dst{A.size() <= B.size() ? A : B};
                            dst.insert(dst.end(), cmp.begin(), cmp.end());
                        }
                    }
                    else
                    {
                        // pseudo-peripheral start: two BFS sweeps
                        bfs(part, part[0], st, order);
                        int const far{order.back()};
                        int const nl{bfs(part, far, st, order)};
                        if(nl >= 3)
                        {
                            std::vector<int> cnt(static_cast<std::size_t>(nl), 0);
                            for(int v: part) { ++cnt[static_cast<std::size_t>(level[static_cast<std::size_t>(v)])]; }
                            int const total{static_cast<int>(part.size())};
                            int below{cnt[0]}, best_t{-1};
                            long best_cost{std::numeric_limits<long>::max()};
                            for(int t{1}; t + 1 < nl; ++t)
                            {
                                int const above{total - below - cnt[static_cast<std::size_t>(t)]};
                                int const lo{std::min(below, above)};
                                // separator size first; strongly penalise lopsided cuts
                                long cost{static_cast<long>(cnt[static_cast<std::size_t>(t)]) * 8};
                                if(lo * 10 < total * 3) { cost += static_cast<long>(total) * 8 + static_cast<long>(std::abs(below - above)); }
                                else
                                {
                                    cost += std::abs(below - above) / 8;
                                }
                                if(lo > 0 && cost < best_cost)
                                {
                                    best_cost = cost;
                                    best_t = t;
                                }
                                below += cnt[static_cast<std::size_t>(t)];
                            }
                            if(best_t >= 0)
                            {
                                for(int v: part)
                                {
                                    int const l{level[static_cast<std::size_t>(v)]};
                                    (l < best_t ? A : (l > best_t ? B : S)).push_back(v);
                                }
                            }
                        }
                    }
                    if(A.empty() || B.empty())
                    {
                        parts[static_cast<std::size_t>(best)] = std::move(part);
                        frozen[static_cast<std::size_t>(best)] = 1;
                        continue;
                    }
                    for(int v: S) { in_top[static_cast<std::size_t>(v)] = 1; }
                    parts[static_cast<std::size_t>(best)] = std::move(A);
                    parts.push_back(std::move(B));
                    frozen.push_back(0);
                }
                if(parts.size() < 2) { return region; }
                for(std::size_t p{}; p < parts.size(); ++p)
                {
                    for(int v: parts[p]) { region[static_cast<std::size_t>(v)] = static_cast<int>(p); }
                }
                return region;
            }

            // ---- elimination tree of regions -----------------------------------------------------------------------
            // Nodes are regions of unknowns: leaves (level 0) are eliminated concurrently by different word streams, a
            // separator node after both of its children.  `stream` says which stream owns a node: leaf number for a
            // leaf, the last leaf of the left subtree for a separator (so every stream owns at most one node per level).
            struct etree
            {
                std::vector<int> region;                  // per unknown: node id
                std::vector<int> parent, level, stream;   // per node
                std::vector<int> order;                   // nodes in elimination order (level, then stream)
                int n_leaves{};
                int n_levels{1};
            };

            // v2 shape: G leaves + one top node
            static etree flat_tree(std::vector<int> const& leaf_region, int n_leaves)
            {
                etree t;
                t.region = leaf_region;
                int const root{n_leaves};
                for(auto& r: t.region)
                {
                    if(r < 0) { r = root; }
                }
                for(int l{}; l < n_leaves; ++l)
                {
                    t.parent.push_back(root);
                    t.level.push_back(0);
                    t.stream.push_back(l);
                    t.order.push_back(l);
                }
                t.parent.push_back(-1);
                t.level.push_back(n_leaves > 0 ? 1 : 0);
                t.stream.push_back(0);
                t.order.push_back(root);
                t.n_leaves = n_leaves;
                t.n_levels = n_leaves > 0 ? 2 : 1;
                return t;
            }

            // recursive nested dissection into (up to) S leaves
            static etree partition_tree(int n, std::vector<std::vector<int>> const& adj, int S, int min_leaf)
            {
                etree t;
                t.region.assign(static_cast<std::size_t>(n), -1);
                std::size_t deg_sum{};
                for(auto const& a: adj) { deg_sum += a.size(); }
                std::size_t const hub_deg{std::max<std::size_t>(32, 10 * (deg_sum / static_cast<std::size_t>(std::max(n, 1)) + 1))};
                std::vector<int> all, hubs;
                for(int i{}; i < n; ++i) { (adj[static_cast<std::size_t>(i)].size() > hub_deg ? hubs : all).push_back(i); }
                std::vector<int> mark(static_cast<std::size_t>(n), -1), lvl(static_cast<std::size_t>(n), -1);
                int stamp{};
                int leaf_counter{};

                auto bfs = [&](int start, int my_stamp, std::vector<int>& order) -> int
                {
                    for(int v: order) { lvl[static_cast<std::size_t>(v)] = -1; }
                    order.clear();
                    lvl[static_cast<std::size_t>(start)] = 0;
                    order.push_back(start);
                    int maxl{};
                    for(std::size_t h{}; h < order.size(); ++h)
                    {
                        int const v{order[h]};
                        for(int u: adj[static_cast<std::size_t>(v)])
                        {
                            if(mark[static_cast<std::size_t>(u)] != my_stamp || lvl[static_cast<std::size_t>(u)] >= 0) { continue; }
                            lvl[static_cast<std::size_t>(u)] = lvl[static_cast<std::size_t>(v)] + 1;
                            maxl = std::max(maxl, lvl[static_cast<std::size_t>(u)]);
                            order.push_back(u);
                        }
                    }
                    return maxl + 1;
                };

                // split `part` into A | Sep | B; false if it cannot be split
                auto bisect = [&](std::vector<int> const& part, std::vector<int>& A, std::vector<int>& B, std::vector<int>& Sep) -> bool
                {
                    int const st{++stamp};
                    for(int v: part)
                    {
                        mark[static_cast<std::size_t>(v)] = st;
                        lvl[static_cast<std::size_t>(v)] = -1;
                    }
                    // connected components first
                    std::vector<std::vector<int>> comps;
                    for(int s0: part)
                    {
                        if(lvl[static_cast<std::size_t>(s0)] >= 0) { continue; }
                        std::vector<int> q{s0};
                        lvl[static_cast<std::size_t>(s0)] = 0;
                        for(std::size_t h{}; h < q.size(); ++h)
                        {
                            for(int u: adj[static_cast<std::size_t>(q[h])])
                            {
                                if(mark[static_cast<std::size_t>(u)] != st || lvl[static_cast<std::size_t>(u)] >= 0) { continue; }
                                lvl[static_cast<std::size_t>(u)] = 0;
                                q.push_back(u);
                            }
                        }
                        comps.push_back(std::move(q));
                    }
                    if(comps.size() > 1)
                    {
                        std::sort(comps.begin(), comps.end(), [](auto const& x, auto const& y) { return x.size() > y.size(); });
                        if(comps[0].size() * 4 <= part.size() * 3)
                        {
                            for(auto const& cmp: comps)
                            {
                                auto& dst{A.size() <= B.size() ? A : B};
                                dst.insert(dst.end(), cmp.begin(), cmp.end());
                            }
                            return !A.empty() && !B.empty();
                        }
                        // one dominant component: cut it, the small ones ride along with the lighter side
                        std::vector<int> big{comps[0]};
                        std::vector<int> rest;
                        for(std::size_t k{1}; k < comps.size(); ++k) { rest.insert(rest.end(), comps[k].begin(), comps[k].end()); }
                        int const st2{++stamp};
                        for(int v: big) { mark[static_cast<std::size_t>(v)] = st2; }
                        std::vector<int> order{big};
                        bfs(big[0], st2, order);
                        int const far{order.back()};
                        order = big;
                        int const nl{bfs(far, st2, order)};
                        if(nl < 3) { return false; }
                        std::vector<int> cnt(static_cast<std::size_t>(nl), 0);
                        for(int v: big) { ++cnt[static_cast<std::size_t>(lvl[static_cast<std::size_t>(v)])]; }
                        int const total{static_cast<int>(big.size())};
                        int below{cnt[0]}, best_t{-1};
                        long best_cost{std::numeric_limits<long>::max()};
                        for(int tt{1}; tt + 1 < nl; ++tt)
                        {
                            int const above{total - below - cnt[static_cast<std::size_t>(tt)]};
                            long const cost{static_cast<long>(cnt[static_cast<std::size_t>(tt)]) * 8 + std::abs(below - above)};
                            if(std::min(below, above) > 0 && cost < best_cost)
                            {
                                best_cost = cost;
                                best_t = tt;
                            }
                            below += cnt[static_cast<std::size_t>(tt)];
                        }
                        if(best_t < 0) { return false; }
                        for(int v: big)
                        {
                            int const l{lvl[static_cast<std::size_t>(v)]};
                            (l < best_t ? A : (l > best_t ? B : Sep)).push_back(v);
                        }
                        auto& dst{A.size() <= B.size() ? A : B};
                        dst.insert(dst.end(), rest.begin(), rest.end());
                        return true;
                    }
                    std::vector<int> order{part};
                    bfs(part[0], st, order);
                    int const far{order.back()};
                    order = part;
                    int const nl{bfs(far, st, order)};
                    if(nl < 3) { return false; }
                    std::vector<int> cnt(static_cast<std::size_t>(nl), 0);
                    for(int v: part) { ++cnt[static_cast<std::size_t>(lvl[static_cast<std::size_t>(v)])]; }
                    int const total{static_cast<int>(part.size())};
                    int below{cnt[0]}, best_t{-1};
                    long best_cost{std::numeric_limits<long>::max()};
                    for(int tt{1}; tt + 1 < nl; ++tt)
                    {
                        int const above{total - below - cnt[static_cast<std::size_t>(tt)]};
                        int const lo{std::min(below, above)};
                        long cost{static_cast<long>(cnt[static_cast<std::size_t>(tt)]) * 8};
                        if(lo * 10 < total * 3) { cost += static_cast<long>(total) * 8 + static_cast<long>(std::abs(below - above)); }
                        else
                        {
                            cost += std::abs(below - above) / 2;
                        }
                        if(lo > 0 && cost < best_cost)
                        {
                            best_cost = cost;
                            best_t = tt;
                        }
                        below += cnt[static_cast<std::size_t>(tt)];
                    }
                    if(best_t < 0) { return false; }
                    for(int v: part)
                    {
                        int const l{lvl[static_cast<std::size_t>(v)]};
                        (l < best_t ? A : (l > best_t ? B : Sep)).push_back(v);
                    }
                    return !A.empty() && !B.empty();
                };

                auto new_node = [&](int level, int stream) -> int
                {
                    t.parent.push_back(-1);
                    t.level.push_back(level);
                    t.stream.push_back(stream);
                    return static_cast<int>(t.parent.size()) - 1;
                };

                auto build = [&](auto&& self, std::vector<int> part, int k) -> int
                {
                    std::vector<int> A, B, Sep;
                    if(k <= 1 || static_cast<int>(part.size()) < 2 * min_leaf || !bisect(part, A, B, Sep))
                    {
                        int const id{new_node(0, leaf_counter++)};
                        for(int v: part) { t.region[static_cast<std::size_t>(v)] = id; }
                        return id;
                    }
                    part.clear();
                    part.shrink_to_fit();
                    int const l{self(self, std::move(A), k / 2)};
                    int const owner{leaf_counter - 1};
                    int const r{self(self, std::move(B), k - k / 2)};
                    int const id{new_node(1 + std::max(t.level[static_cast<std::size_t>(l)], t.level[static_cast<std::size_t>(r)]), owner)};
                    t.parent[static_cast<std::size_t>(l)] = id;
                    t.parent[static_cast<std::size_t>(r)] = id;
                    for(int v: Sep) { t.region[static_cast<std::size_t>(v)] = id; }
                    return id;
                };
                int root{build(build, std::move(all), std::max(S, 1))};
                if(!hubs.empty())
                {
                    if(t.level[static_cast<std::size_t>(root)] == 0)
                    {
                        int const id{new_node(1, t.stream[static_cast<std::size_t>(root)])};
                        t.parent[static_cast<std::size_t>(root)] = id;
                        root = id;
                    }
                    for(int v: hubs) { t.region[static_cast<std::size_t>(v)] = root; }
                }
                // separators of one level are owned by consecutive streams (level 1 first), so that a level keeps whole
                // warps busy instead of one thread in every warp; every stream still owns at most one node per level
                {
                    std::vector<int> seps;
                    for(std::size_t k{}; k < t.parent.size(); ++k)
                    {
                        if(t.level[k] > 0) { seps.push_back(static_cast<int>(k)); }
                    }
                    std::stable_sort(seps.begin(),
                                     seps.end(),
                                     [&](int a, int b)
                                     {
                                         if(t.level[static_cast<std::size_t>(a)] != t.level[static_cast<std::size_t>(b)])
                                         {
                                             return t.level[static_cast<std::size_t>(a)] < t.level[static_cast<std::size_t>(b)];
                                         }
                                         return t.stream[static_cast<std::size_t>(a)] < t.stream[static_cast<std::size_t>(b)];
                                     });
                    for(std::size_t k{}; k < seps.size(); ++k) { t.stream[static_cast<std::size_t>(seps[k])] = static_cast<int>(k); }
                }
                t.n_leaves = leaf_counter;
                t.n_levels = t.level[static_cast<std::size_t>(root)] + 1;
                t.order.resize(t.parent.size());
                std::iota(t.order.begin(), t.order.end(), 0);
                std::stable_sort(t.order.begin(),
                                 t.order.end(),
                                 [&](int a, int b)
                                 {
                                     if(t.level[static_cast<std::size_t>(a)] != t.level[static_cast<std::size_t>(b)])
                                     {
                                         return t.level[static_cast<std::size_t>(a)] < t.level[static_cast<std::size_t>(b)];
                                     }
                                     return t.stream[static_cast<std::size_t>(a)] < t.stream[static_cast<std::size_t>(b)];
                                 });
                return t;
            }

            // ---- symbolic LU + emission ------------------------------------------------------------------------
            struct lu_step
            {
                int r, c, piv;
                bool guard{true};               // the pivot op carries PE_F_GUARD (false: provably safe, see "guard elision")
                int leaf;                       // v2 schedule: leaf the pivot was taken in (-1 = top)
                int node;                       // elimination-tree node the pivot was taken in
                std::vector<int> lrows, l_ent;  // rows i with (i,c), entry ids
                std::vector<int> ucols, u_ent;  // cols j with (r,j), entry ids
                std::vector<int> t_ent;         // lrows.size() * ucols.size() target entry ids
            };

            struct stream
            {
                std::vector<std::uint32_t> w;
                int tmp_slot{-1};
            };

            struct upd
            {
                int l, u;  // l: L entry id; u: U entry id (matrix entries) or pivot-row index (rhs)
                int leaf;  // v2 schedule: leaf of the elimination step that generates this update (-1 = top)
                int node;  // elimination-tree node of that step
            };

            // ================= resident (shared-memory) schedule: DESIGN.md §5 =================================
            struct resident_ctx
            {
                pstate& ps;
                program& pr;
                etree const& tree;
                int S;
                std::vector<lu_step> const& steps;
                std::vector<std::vector<upd>> const& pairs;
                std::vector<std::vector<upd>> const& ypairs;
                std::vector<char> const& ynz;
                std::vector<entry const*> const& orig;
                int n;
                std::size_t n_ent;
            };

            static int value_op_outputs(std::uint32_t opcode)
            {
                switch(opcode)
                {
                    case PE_OP_SINCOS:
                    case PE_OP_CAP_STEP:
                    case PE_OP_RELAY_EVAL:
                    case PE_OP_IND_STEP: return 2;
                    case PE_OP_PN_PREP: return 5;
                    case PE_OP_KIND_STEP:
                    case PE_OP_PN_EVAL:
                    case PE_OP_PN_STEP:
                    case PE_OP_NMOS_EVAL:
                    case PE_OP_PMOS_EVAL: return 3;
                    case PE_OP_BJT_EVAL: return 4;
                    default: return 1;
                }
            }

            void resident_singular(program& pr, pstate& ps, int S, int n)
            {
                pr.resident = true;
                pr.rS = S;
                pr.cplx = ps.cplx;
                pr.rstreams.assign(static_cast<std::size_t>(S), {});
                for(auto& rs: pr.rstreams)
                {
                    rs.sec[2].resize(1);
                    rop bad;
                    bad.opcode = 0xffu;  // unknown opcode -> every lane reports PE_ST_SINGULAR
                    rs.sec[2][0].push_back(bad);
                }
                pr.r_slots = S;
                pr.r_zero = 0;
                pr.io.clear();
                pr.io.push_back({0u | (static_cast<std::uint32_t>(PE_IO_CONST) << 16) | (PE_IO_LOAD << 20), 0u});
                pr.x_slot.assign(static_cast<std::size_t>(n), 0u);
                pr.x_opnd.assign(static_cast<std::size_t>(n), PE_OPND(PE_SP_CONST, 0));
                pr.n_lane_slots = ps.cplx ? 1 + 2 * n : 0;
                pr.omega_slot = ps.cplx ? 0 : -1;
                pr.warps = 1;
                pr.built = true;
            }

            void emit_resident(resident_ctx& rc)
            {
                auto& ps{rc.ps};
                auto& pr{rc.pr};
                auto const& tree{rc.tree};
                auto const& steps{rc.steps};
                int const S{rc.S};
                int const n{rc.n};
                std::size_t const n_ent{rc.n_ent};
                bool const cplx{ps.cplx};
                int const w{cplx ? 2 : 1};
                int const L{tree.n_levels};
                int const n_ph{2 * L};
                constexpr std::uint32_t NEG{0x80000000u};
                bool const fuse_steps{in.fuse_steps};

                pr.resident = true;
                pr.rS = S;
                pr.cplx = cplx;
                auto& RS{pr.rstreams};
                RS.assign(static_cast<std::size_t>(S), {});
                for(auto& rs: RS)
                {
                    rs.sec[0].resize(1);
                    rs.sec[1].resize(1);
                    rs.sec[2].resize(static_cast<std::size_t>(n_ph));
                }

                // lane-space values: 0 scalar, 1 re of a complex unit, 2 its im
                std::vector<std::uint8_t> lkind(static_cast<std::size_t>(ps.n_lane), 0);
                auto new_lane = [&]() -> std::uint32_t
                {
                    int const s{ps.n_lane};
                    ps.n_lane += w;
                    lkind.resize(static_cast<std::size_t>(ps.n_lane), 0);
                    if(cplx)
                    {
                        lkind[static_cast<std::size_t>(s)] = 1;
                        lkind[static_cast<std::size_t>(s) + 1] = 2;
                    }
                    return IOP(I_LANE, s);
                };
                std::vector<std::uint32_t> ekey(n_ent, 0u), ykey(static_cast<std::size_t>(n), 0u);
                std::vector<char> ehas(n_ent, 0), yhas(static_cast<std::size_t>(n), 0);
                auto key_e = [&](int id) -> std::uint32_t
                {
                    if(!ehas[static_cast<std::size_t>(id)])
                    {
                        ekey[static_cast<std::size_t>(id)] = new_lane();
                        ehas[static_cast<std::size_t>(id)] = 1;
                    }
                    return ekey[static_cast<std::size_t>(id)];
                };
                auto key_y = [&](int row) -> std::uint32_t
                {
                    if(!yhas[static_cast<std::size_t>(row)])
                    {
                        ykey[static_cast<std::size_t>(row)] = new_lane();
                        yhas[static_cast<std::size_t>(row)] = 1;
                    }
                    return ykey[static_cast<std::size_t>(row)];
                };
                std::vector<std::uint32_t> xs(static_cast<std::size_t>(n));
                for(int j{}; j < n; ++j) { xs[static_cast<std::size_t>(j)] = cplx ? new_lane() : IOP(I_INST, j); }

                auto node_stream = [&](int node) { return tree.stream[static_cast<std::size_t>(node)] % S; };
                auto fwd_phase = [&](int node) { return 1 + tree.level[static_cast<std::size_t>(node)]; };
                auto back_phase = [&](int node)
                {
                    int const l{tree.level[static_cast<std::size_t>(node)]};
                    return l == L - 1 ? L : 2 * L - 1 - l;
                };

                // ---- owners and contributions
                std::vector<int> ent_owner(n_ent, -1), row_owner(static_cast<std::size_t>(n), -1);
                for(auto const& st: steps)
                {
                    ent_owner[static_cast<std::size_t>(st.piv)] = st.node;
                    for(int e: st.u_ent) { ent_owner[static_cast<std::size_t>(e)] = st.node; }
                    for(int e: st.l_ent) { ent_owner[static_cast<std::size_t>(e)] = st.node; }
                    row_owner[static_cast<std::size_t>(st.r)] = st.node;
                }
                // contribution of node M to a target it does not own, when M generates at least two updates to it
                std::map<std::pair<std::int64_t, int>, std::uint32_t> contrib;
                std::vector<std::vector<std::pair<std::int64_t, std::uint32_t>>> node_contribs(tree.parent.size());
                auto scan_contribs = [&](std::int64_t target, std::vector<upd> const& ups, int owner)
                {
                    std::map<int, int> cnt;
                    for(auto const& u: ups)
                    {
                        if(u.node != owner) { ++cnt[u.node]; }
                    }
                    for(auto const& [m, c]: cnt)
                    {
                        if(c >= 2)
                        {
                            std::uint32_t const k{new_lane()};
                            contrib[{target, m}] = k;
                            node_contribs[static_cast<std::size_t>(m)].push_back({target, k});
                        }
                    }
                };
                for(std::size_t e{}; e < n_ent; ++e)
                {
                    if(ent_owner[e] >= 0) { scan_contribs(static_cast<std::int64_t>(e), rc.pairs[e], ent_owner[e]); }
                }
                for(int r{}; r < n; ++r) { scan_contribs(static_cast<std::int64_t>(n_ent) + r, rc.ypairs[static_cast<std::size_t>(r)], row_owner[static_cast<std::size_t>(r)]); }

                // ---- abstract DOT emission (operands are internal words; translated to slots after allocation)
                std::vector<std::uint32_t> tmp_key(static_cast<std::size_t>(S), 0u);
                std::vector<char> tmp_has(static_cast<std::size_t>(S), 0);
                std::uint32_t const dot_op{cplx ? static_cast<std::uint32_t>(PE_OP_CDOT) : static_cast<std::uint32_t>(PE_OP_DOT)};
                auto emit_dot = [&](int sj, int phase, std::uint32_t dst, std::uint32_t flags, std::uint32_t scale, std::vector<std::uint32_t> const& sre,
                                    std::vector<std::uint32_t> const& sim, std::vector<std::pair<std::uint32_t, std::uint32_t>> const& pp)
                {
                    // a vector op holds at most 32 rows (one mask bit each): ctl, scale, packed sources, pairs
                    // (two header words + ctl + scale + rows must fit the 30 program words of a line, see pack())
                    std::size_t const max_src{cplx ? 12u : 16u}, max_pair{cplx ? 12u : 16u};
                    auto& dstv{RS[static_cast<std::size_t>(sj)].sec[2][static_cast<std::size_t>(phase)]};
                    std::size_t ri{}, ii{}, pi{};
                    bool first{true};
                    for(;;)
                    {
                        std::size_t const carry{first ? 0u : 1u};
                        std::size_t const nr{std::min(sre.size() - ri, max_src - carry)};
                        std::size_t const ni{std::min(sim.size() - ii, max_src - carry)};
                        std::size_t const np{std::min(pp.size() - pi, max_pair)};
                        bool const last{ri + nr == sre.size() && ii + ni == sim.size() && pi + np == pp.size()};
                        rop o;
                        o.opcode = dot_op;
                        if(!last || !first)
                        {
                            if(!tmp_has[static_cast<std::size_t>(sj)])
                            {
                                tmp_key[static_cast<std::size_t>(sj)] = new_lane();
                                tmp_has[static_cast<std::size_t>(sj)] = 1;
                            }
                        }
                        std::uint32_t const tk{tmp_key[static_cast<std::size_t>(sj)]};
                        o.dst = last ? dst : tk;
                        o.flags = last ? flags : 0u;
                        o.scale = (o.flags & PE_F_SCALE) ? scale : 0u;
                        if(carry)
                        {
                            o.sre.push_back(tk);
                            if(cplx) { o.sim.push_back(tk + 1); }
                        }
                        o.sre.insert(o.sre.end(), sre.begin() + static_cast<std::ptrdiff_t>(ri), sre.begin() + static_cast<std::ptrdiff_t>(ri + nr));
                        o.sim.insert(o.sim.end(), sim.begin() + static_cast<std::ptrdiff_t>(ii), sim.begin() + static_cast<std::ptrdiff_t>(ii + ni));
                        o.pp.insert(o.pp.end(), pp.begin() + static_cast<std::ptrdiff_t>(pi), pp.begin() + static_cast<std::ptrdiff_t>(pi + np));
                        dstv.push_back(std::move(o));
                        ri += nr;
                        ii += ni;
                        pi += np;
                        first = false;
                        if(last) { break; }
                    }
                };

                std::vector<std::uint32_t> sre, sim;
                std::vector<std::pair<std::uint32_t, std::uint32_t>> pp;
                auto load_sources = [&](entry const* e)
                {
                    sre.clear();
                    sim.clear();
                    if(e == nullptr) { return; }
                    sre = e->re;
                    sim = e->im;
                };
                auto load_updates = [&](std::int64_t target, std::vector<upd> const& ups, int owner, bool rhs)
                {
                    pp.clear();
                    std::set<int> seen;
                    for(auto const& u: ups)
                    {
                        if(u.node != owner)
                        {
                            auto it{contrib.find({target, u.node})};
                            if(it != contrib.end())
                            {
                                if(seen.insert(u.node).second)
                                {
                                    sre.push_back(it->second);
                                    if(cplx) { sim.push_back(it->second + 1); }
                                }
                                continue;
                            }
                        }
                        pp.push_back({key_e(u.l), rhs ? key_y(u.u) : key_e(u.u)});
                    }
                };

                // ---- forward steps (Crout: pivot, row of U, column of L, rhs entry), then the node's contributions
                for(std::size_t k{}; k < steps.size(); ++k)
                {
                    auto const& st{steps[k]};
                    int const sj{node_stream(st.node)};
                    int const ph{fwd_phase(st.node)};
                    std::size_t const step_begin{RS[static_cast<std::size_t>(sj)].sec[2][static_cast<std::size_t>(ph)].size()};
                    load_sources(rc.orig[static_cast<std::size_t>(st.piv)]);
                    load_updates(st.piv, rc.pairs[static_cast<std::size_t>(st.piv)], st.node, false);
                    emit_dot(sj, ph, key_e(st.piv), PE_F_RECIP, 0, sre, sim, pp);
                    for(int e: st.u_ent)
                    {
                        load_sources(rc.orig[static_cast<std::size_t>(e)]);
                        load_updates(e, rc.pairs[static_cast<std::size_t>(e)], st.node, false);
                        // a U entry that is just (+/-) one stamped value, with no update and no contribution, is never
                        // materialised: its consumers name the value itself (pair operands carry a sign bit)
                        if(!cplx && pp.empty() && sim.empty() && sre.size() == 1 && !ehas[static_cast<std::size_t>(e)])
                        {
                            ekey[static_cast<std::size_t>(e)] = sre[0];  // internal word incl. its negate bit
                            ehas[static_cast<std::size_t>(e)] = 1;
                            ++pr.n_aliased;
                            continue;
                        }
                        emit_dot(sj, ph, key_e(e), 0, 0, sre, sim, pp);
                    }
                    for(int e: st.l_ent)
                    {
                        load_sources(rc.orig[static_cast<std::size_t>(e)]);
                        load_updates(e, rc.pairs[static_cast<std::size_t>(e)], st.node, false);
                        emit_dot(sj, ph, key_e(e), PE_F_SCALE | (st.guard ? PE_F_GUARD : 0u), key_e(st.piv), sre, sim, pp);  // an entry of L
                    }
                    if(rc.ynz[static_cast<std::size_t>(st.r)])
                    {
                        load_sources(&ps.Z[static_cast<std::size_t>(st.r)]);
                        load_updates(static_cast<std::int64_t>(n_ent) + st.r, rc.ypairs[static_cast<std::size_t>(st.r)], st.node, true);
                        emit_dot(sj, ph, key_y(st.r), 0, 0, sre, sim, pp);
                    }
                    // a step whose DOTs are all small becomes one fused op (PE_OP_CROUT2): the kernels then request every
                    // operand of the step at once and keep the pivot reciprocal in a register
                    if(fuse_steps && !cplx)
                    {
                        auto& lst{RS[static_cast<std::size_t>(sj)].sec[2][static_cast<std::size_t>(ph)]};
                        std::size_t const n_new{lst.size() - step_begin};
                        bool fits{n_new >= 2 && n_new <= 6};
                        for(std::size_t q{}; fits && q < n_new; ++q)
                        {
                            auto const& o{lst[step_begin + q]};
                            fits = o.opcode == PE_OP_DOT && o.sim.empty() && o.pp.size() <= 1 && o.sre.size() <= (q == 0 ? 4u : 2u) &&
                                   (q == 0 ? o.flags == PE_F_RECIP : (o.flags == 0u || ((o.flags & ~static_cast<std::uint32_t>(PE_F_GUARD)) == PE_F_SCALE && o.scale == lst[step_begin].dst)));
                        }
                        if(fits)
                        {
                            rop f;
                            f.opcode = PE_OP_CROUT2;
                            f.sub.assign(std::make_move_iterator(lst.begin() + static_cast<std::ptrdiff_t>(step_begin)), std::make_move_iterator(lst.end()));
                            lst.resize(step_begin);
                            lst.push_back(std::move(f));
                            ++pr.n_fused;
                        }
                    }
                }
                for(std::size_t m{}; m < node_contribs.size(); ++m)
                {
                    int const sj{node_stream(static_cast<int>(m))};
                    int const ph{fwd_phase(static_cast<int>(m))};
                    for(auto const& [target, key]: node_contribs[m])
                    {
                        bool const rhs{target >= static_cast<std::int64_t>(n_ent)};
                        auto const& ups{rhs ? rc.ypairs[static_cast<std::size_t>(target - static_cast<std::int64_t>(n_ent))] : rc.pairs[static_cast<std::size_t>(target)]};
                        sre.clear();
                        sim.clear();
                        pp.clear();
                        for(auto const& u: ups)
                        {
                            if(u.node == static_cast<int>(m)) { pp.push_back({key_e(u.l), rhs ? key_y(u.u) : key_e(u.u)}); }
                        }
                        emit_dot(sj, ph, key, 0, 0, sre, sim, pp);
                    }
                }
                // ---- back substitution, root first
                for(std::size_t k{steps.size()}; k-- > 0;)
                {
                    auto const& st{steps[k]};
                    sre.clear();
                    sim.clear();
                    pp.clear();
                    if(rc.ynz[static_cast<std::size_t>(st.r)])
                    {
                        sre.push_back(key_y(st.r));
                        if(cplx) { sim.push_back(key_y(st.r) + 1); }
                    }
                    for(std::size_t b{}; b < st.ucols.size(); ++b) { pp.push_back({key_e(st.u_ent[b]), xs[static_cast<std::size_t>(st.ucols[b])]}); }
                    std::uint32_t flags{};
                    if(!sre.empty() || !pp.empty()) { flags |= PE_F_SCALE; }
                    if(!cplx) { flags |= (st.c >= num.n_nodes) ? PE_F_CHECK_I : PE_F_CHECK_V; }
                    emit_dot(node_stream(st.node), back_phase(st.node), xs[static_cast<std::size_t>(st.c)], flags, key_e(st.piv), sre, sim, pp);
                }

                // ---- value ops: next to the stream that owns their element's unknown, same opcodes aligned
                auto deal = [&](std::vector<op_t> const& ops, int sec, int phase)
                {
                    if(ops.empty()) { return; }
                    std::vector<std::vector<op_t const*>> per(static_cast<std::size_t>(S));
                    std::size_t rr{};
                    for(auto const& o: ops)
                    {
                        int sj;
                        if(o.aff >= 0 && o.aff < n) { sj = node_stream(tree.region[static_cast<std::size_t>(o.aff)]); }
                        else
                        {
                            sj = static_cast<int>(rr++ % static_cast<std::size_t>(S));
                        }
                        per[static_cast<std::size_t>(sj)].push_back(&o);
                    }
                    // cap the imbalance: a hub node would otherwise serialise every device hanging on it
                    std::size_t const cap{2 * ((ops.size() + static_cast<std::size_t>(S) - 1) / static_cast<std::size_t>(S)) + 4};
                    std::vector<op_t const*> spill;
                    for(auto& v: per)
                    {
                        while(v.size() > cap)
                        {
                            spill.push_back(v.back());
                            v.pop_back();
                        }
                    }
                    for(auto const* o: spill)
                    {
                        auto it{std::min_element(per.begin(), per.end(), [](auto const& a, auto const& b) { return a.size() < b.size(); })};
                        it->push_back(o);
                    }
                    for(int sj{}; sj < S; ++sj)
                    {
                        auto& v{per[static_cast<std::size_t>(sj)]};
                        std::stable_sort(v.begin(), v.end(), [](op_t const* a, op_t const* b) { return a->w[0] < b->w[0]; });
                        auto& dstv{RS[static_cast<std::size_t>(sj)].sec[sec][static_cast<std::size_t>(phase)]};
                        for(auto const* o: v)
                        {
                            rop r;
                            r.opcode = o->w[0];
                            r.opnd.assign(o->w.begin() + 1, o->w.end());
                            dstv.push_back(std::move(r));
                        }
                    }
                };
                if(!cplx) { deal(prep_ops, 0, 0); }
                deal(ps.head, 2, 0);
                bool const merge_step{in.merge_step && !ps.step.empty()};
                bool step_left{false};
                if(merge_step)
                {
                    // update_tr_step (circuit.h:363-374) folded into the elimination sweep: the companion update of an
                    // element runs right before the first op of its stream that reads what it writes (first Newton iteration
                    // of the time step only), so the fresh history / conductance is consumed from L1 instead of being
                    // written in a separate section and fetched back from DRAM hundreds of ops later
                    auto reads_any = [&](rop const& o, std::vector<std::uint32_t> const& keys) -> bool
                    {
                        auto hit = [&](std::uint32_t k) { return std::find(keys.begin(), keys.end(), k & ~NEG) != keys.end(); };
                        auto one = [&](rop const& d) -> bool
                        {
                            if(d.opcode == PE_OP_DOT || d.opcode == PE_OP_CDOT)
                            {
                                for(auto k: d.sre)
                                {
                                    if(hit(k)) { return true; }
                                }
                                for(auto k: d.sim)
                                {
                                    if(hit(k)) { return true; }
                                }
                                for(auto const& [a, b]: d.pp)
                                {
                                    if(hit(a) || hit(b)) { return true; }
                                }
                                return (d.flags & PE_F_SCALE) && hit(d.scale);
                            }
                            for(auto k: d.opnd)
                            {
                                if(hit(k)) { return true; }
                            }
                            return false;
                        };
                        if(o.opcode == PE_OP_CROUT2)
                        {
                            for(auto const& sb: o.sub)
                            {
                                if(one(sb)) { return true; }
                            }
                            return false;
                        }
                        return one(o);
                    };
                    // what the head phase (device evaluation, sources: phase 0 of every stream) reads and writes: an update
                    // whose outputs it reads, or whose inputs it writes (PN_STEP <-> PN_EVAL), keeps its place in the step
                    // section, which runs before the Newton loop
                    std::vector<std::uint32_t> head_reads, head_writes;
                    for(auto const& rs: RS)
                    {
                        for(auto const& o: rs.sec[2][0])
                        {
                            int const no{value_op_outputs(o.opcode)};
                            for(std::size_t k{}; k < o.opnd.size(); ++k)
                            {
                                (static_cast<int>(k) < no ? head_writes : head_reads).push_back(o.opnd[k] & ~NEG);
                                if(static_cast<int>(k) < no) { head_reads.push_back(o.opnd[k] & ~NEG); }  // in/out operands
                            }
                        }
                    }
                    auto in_list = [](std::vector<std::uint32_t> const& v, std::uint32_t k) { return std::find(v.begin(), v.end(), k) != v.end(); };
                    std::vector<op_t> unfolded;
                    std::size_t rr{};
                    for(auto const& o: ps.step)
                    {
                        rop r;
                        r.opcode = o.w[0];
                        r.opnd.assign(o.w.begin() + 1, o.w.end());
                        r.first_only = 1;
                        std::vector<std::uint32_t> outs;
                        for(int k{}; k < value_op_outputs(r.opcode) && k < static_cast<int>(r.opnd.size()); ++k) { outs.push_back(r.opnd[static_cast<std::size_t>(k)] & ~NEG); }
                        bool clash{false};
                        for(std::size_t k{}; k < r.opnd.size(); ++k)
                        {
                            std::uint32_t const key{r.opnd[k] & ~NEG};
                            bool const is_out{static_cast<int>(k) < value_op_outputs(r.opcode)};
                            clash = clash || (is_out && in_list(head_reads, key)) || in_list(head_writes, key);
                        }
                        if(clash)
                        {
                            unfolded.push_back(o);
                            continue;
                        }
                        int sj{static_cast<int>(rr % static_cast<std::size_t>(S))}, ph{0};
                        if(o.aff >= 0 && o.aff < n)
                        {
                            int const node{tree.region[static_cast<std::size_t>(o.aff)]};
                            sj = node_stream(node);
                            ph = fwd_phase(node);
                        }
                        else
                        {
                            ++rr;
                        }
                        auto& lst{RS[static_cast<std::size_t>(sj)].sec[2][static_cast<std::size_t>(ph)]};
                        std::size_t pos{lst.size()};
                        for(std::size_t q{}; q < lst.size(); ++q)
                        {
                            if(reads_any(lst[q], outs))
                            {
                                pos = q;
                                break;
                            }
                        }
                        if(pos == lst.size() && ph != 0)
                        {
                            // nobody in that phase reads it (a consumer on another stream or level): the head phase
                            // precedes every reader
                            RS[static_cast<std::size_t>(sj)].sec[2][0].push_back(std::move(r));
                        }
                        else
                        {
                            lst.insert(lst.begin() + static_cast<std::ptrdiff_t>(pos), std::move(r));
                        }
                    }
                    deal(unfolded, 1, 0);
                    step_left = !unfolded.empty();
                }
                else
                {
                    deal(ps.step, 1, 0);
                    step_left = !ps.step.empty();
                }
                pr.has_sec[0] = !cplx && !prep_ops.empty();
                pr.has_sec[1] = step_left;
                pr.has_sec[2] = true;

                // ---- global alignment: within every phase the S op lists are merged position by position (greedy on
                // the op signature), bubbles fill the gaps.  Isomorphic sub-trees end up executing the same op at the same
                // position, which is what makes their operand rows warp-uniform after the relative slot encoding below.
                auto signature = [](rop const& o) -> std::uint64_t
                {
                    return (static_cast<std::uint64_t>(o.opcode) << 56) ^ (static_cast<std::uint64_t>(o.flags | (o.first_only ? 0x80u : 0u)) << 48) ^ (static_cast<std::uint64_t>(o.sre.size()) << 36) ^
                           (static_cast<std::uint64_t>(o.sim.size()) << 24) ^ (static_cast<std::uint64_t>(o.pp.size()) << 12) ^ static_cast<std::uint64_t>(o.opnd.size());
                };
                for(int sec{}; sec < 3; ++sec)
                {
                    std::size_t const nph{RS[0].sec[sec].size()};
                    for(std::size_t ph{}; ph < nph; ++ph)
                    {
                        std::vector<rphase> outl(static_cast<std::size_t>(S));
                        std::vector<std::size_t> idx(static_cast<std::size_t>(S), 0);
                        for(;;)
                        {
                            std::map<std::uint64_t, int> votes;
                            for(int sj{}; sj < S; ++sj)
                            {
                                auto const& l{RS[static_cast<std::size_t>(sj)].sec[sec][ph]};
                                if(idx[static_cast<std::size_t>(sj)] < l.size()) { ++votes[signature(l[idx[static_cast<std::size_t>(sj)]])]; }
                            }
                            if(votes.empty()) { break; }
                            std::uint64_t best_sig{};
                            int best{-1};
                            for(auto const& [sg, k]: votes)
                            {
                                if(k > best)
                                {
                                    best = k;
                                    best_sig = sg;
                                }
                            }
                            for(int sj{}; sj < S; ++sj)
                            {
                                auto& l{RS[static_cast<std::size_t>(sj)].sec[sec][ph]};
                                auto& o{outl[static_cast<std::size_t>(sj)]};
                                if(idx[static_cast<std::size_t>(sj)] < l.size() && signature(l[idx[static_cast<std::size_t>(sj)]]) == best_sig)
                                {
                                    o.push_back(std::move(l[idx[static_cast<std::size_t>(sj)]]));
                                    ++idx[static_cast<std::size_t>(sj)];
                                }
                                else
                                {
                                    rop b;
                                    b.bubble = 1;
                                    o.push_back(std::move(b));
                                }
                            }
                        }
                        for(int sj{}; sj < S; ++sj) { RS[static_cast<std::size_t>(sj)].sec[sec][ph] = std::move(outl[static_cast<std::size_t>(sj)]); }
                    }
                }

                // ---- shared-memory slot allocation
                std::uint32_t const zero_key{IOP(I_CONST, 0)};
                auto unit_of = [&](std::uint32_t key) -> std::pair<std::uint32_t, int>  // (first key of the unit, rows)
                {
                    std::uint32_t const k{key & ~NEG};
                    // the padding operand of complex pairs is read as (re, im): 0.0 needs a zero imaginary row as well
                    if(cplx && k == zero_key) { return {k, 2}; }
                    if((k >> 29) == I_LANE)
                    {
                        std::uint32_t const sl{k & 0x1fffffffu};
                        std::uint8_t const kd{sl < lkind.size() ? lkind[sl] : static_cast<std::uint8_t>(0)};
                        if(kd == 1) { return {k, 2}; }
                        if(kd == 2) { return {k - 1, 2}; }
                    }
                    return {k, 1};
                };
                // the scalar keys an op touches, in a canonical role order: fn(key, is_write)
                auto visit = [&](rop const& o0, auto&& fn)
                {
                  auto one = [&](rop const& o)
                  {
                    if(o.bubble) { return; }
                    if(o.opcode == PE_OP_DOT || o.opcode == PE_OP_CDOT)
                    {
                        bool const cx{o.opcode == PE_OP_CDOT};
                        fn(o.dst, true);
                        if(cx) { fn(o.dst + 1, true); }
                        if(o.flags & PE_F_SCALE)
                        {
                            fn(o.scale, false);
                            if(cx) { fn(o.scale + 1, false); }
                        }
                        for(auto k: o.sre) { fn(k & ~NEG, false); }
                        for(auto k: o.sim) { fn(k & ~NEG, false); }
                        for(auto const& [a, b]: o.pp)
                        {
                            fn(a, false);
                            fn(b, false);
                            if(cx)
                            {
                                fn(a + 1, false);
                                fn(b + 1, false);
                            }
                        }
                    }
                    else if(o.opcode != 0xffu)
                    {
                        int const no{value_op_outputs(o.opcode)};
                        for(std::size_t k{}; k < o.opnd.size(); ++k) { fn(o.opnd[k] & ~NEG, static_cast<int>(k) < no); }
                    }
                  };
                  if(o0.opcode == PE_OP_CROUT2)
                  {
                      for(auto const& sb: o0.sub) { one(sb); }
                  }
                  else
                  {
                      one(o0);
                  }
                };
                struct vinfo
                {
                    int owner{-1};
                    int rank{99};
                    bool written{};
                    bool replicated{};
                    int col{-1}, row{-1};
                    std::set<int> users;
                };
                std::map<std::uint32_t, vinfo> vals;  // key = first key of the unit
                vals[zero_key];                       // the padding operand always exists
                int const sec_order[3]{2, 1, 0};
                for(int so{}; so < 3; ++so)
                {
                    int const sec{sec_order[so]};
                    for(int sj{}; sj < S; ++sj)
                    {
                        for(auto const& ph: RS[static_cast<std::size_t>(sj)].sec[sec])
                        {
                            for(auto const& o: ph)
                            {
                                visit(o,
                                      [&](std::uint32_t key, bool wr)
                                      {
                                          auto& v{vals[unit_of(key).first]};
                                          int const rank{wr ? so : 3};
                                          if(rank < v.rank)
                                          {
                                              v.rank = rank;
                                              v.owner = sj;
                                          }
                                          v.written = v.written || wr;
                                          if((key >> 29) == I_CONST && v.users.size() < 64) { v.users.insert(sj); }
                                      });
                            }
                        }
                    }
                }
                // constants many streams read get one copy per column (relative offset 0 => warp-uniform operand words)
                int n_rep_rows{};
                {
                    std::size_t const many{std::max<std::size_t>(2, static_cast<std::size_t>(S) / 8)};
                    for(auto& [k, v]: vals)
                    {
                        if((k >> 29) != I_CONST || v.written) { continue; }
                        if(k == zero_key || (S > 1 && v.users.size() >= many))
                        {
                            v.replicated = true;
                            v.col = 0;
                            v.row = n_rep_rows;
                            n_rep_rows += unit_of(k).second;
                        }
                    }
                }
                if(vals[zero_key].owner < 0) { vals[zero_key].owner = 0; }
                std::vector<int> next_free(static_cast<std::size_t>(S), n_rep_rows);
                for(int so{}; so < 3; ++so)
                {
                    int const sec{sec_order[so]};
                    std::size_t const nph{RS[0].sec[sec].size()};
                    for(std::size_t ph{}; ph < nph; ++ph)
                    {
                        std::size_t const npos{RS[0].sec[sec][ph].size()};
                        for(std::size_t pos{}; pos < npos; ++pos)
                        {
                            // role-major: the r-th key of every active stream is allocated together, on a common row
                            std::vector<std::vector<std::uint32_t>> keys(static_cast<std::size_t>(S));
                            std::size_t max_roles{};
                            for(int sj{}; sj < S; ++sj)
                            {
                                visit(RS[static_cast<std::size_t>(sj)].sec[sec][ph][pos], [&](std::uint32_t key, bool) { keys[static_cast<std::size_t>(sj)].push_back(key); });
                                max_roles = std::max(max_roles, keys[static_cast<std::size_t>(sj)].size());
                            }
                            for(std::size_t r{}; r < max_roles; ++r)
                            {
                                int row{-1};
                                std::vector<std::pair<int, std::uint32_t>> todo;
                                for(int sj{}; sj < S; ++sj)
                                {
                                    if(r >= keys[static_cast<std::size_t>(sj)].size()) { continue; }
                                    auto const [u0, rows]{unit_of(keys[static_cast<std::size_t>(sj)][r])};
                                    (void)rows;
                                    auto& v{vals[u0]};
                                    if(v.owner != sj || v.col >= 0) { continue; }
                                    v.col = sj;  // claimed (row set below)
                                    todo.push_back({sj, u0});
                                    row = std::max(row, next_free[static_cast<std::size_t>(sj)]);
                                }
                                for(auto const& [sj, u0]: todo)
                                {
                                    vals[u0].row = row;
                                    next_free[static_cast<std::size_t>(sj)] = row + unit_of(u0).second;
                                }
                            }
                        }
                    }
                }
                for(auto& [k, v]: vals)
                {
                    if(v.col < 0)  // never touched by its owner in program order (cannot happen)
                    {
                        int const cj{static_cast<int>(std::min_element(next_free.begin(), next_free.end()) - next_free.begin())};
                        v.col = cj;
                        v.row = next_free[static_cast<std::size_t>(cj)];
                        next_free[static_cast<std::size_t>(cj)] += unit_of(k).second;
                    }
                }
                int const K{std::max(1, *std::max_element(next_free.begin(), next_free.end()))};
                pr.r_slots = K * S;
                if(pr.r_slots > PE_R_MAX_SLOTS)
                {
                    out.error = "resident program needs more than 32768 shared-memory slots";
                    pr.resident = false;  // the caller falls back to the HBM-streaming form
                    pr.built = false;
                    return;
                }
                // absolute slot of a key; operand words are stream-relative: row * S + ((col - stream) mod S), 0 for a
                // replicated constant; the interpreter adds its own stream index back (mod S)
                auto slot_of = [&](std::uint32_t key) -> std::uint32_t
                {
                    std::uint32_t const k{key & ~NEG};
                    auto const [u0, rows]{unit_of(k)};
                    (void)rows;
                    auto const& v{vals.at(u0)};
                    return static_cast<std::uint32_t>((v.row + static_cast<int>(k - u0)) * S + v.col);
                };
                auto rel_of = [&](std::uint32_t key, int sj) -> std::uint32_t
                {
                    std::uint32_t const k{key & ~NEG};
                    auto const [u0, rows]{unit_of(k)};
                    (void)rows;
                    auto const& v{vals.at(u0)};
                    int const d{v.replicated ? 0 : ((v.col - sj) % S + S) % S};
                    return static_cast<std::uint32_t>((v.row + static_cast<int>(k - u0)) * S + d) | ((key & NEG) ? PE_R_NEG : 0u);
                };
                pr.r_zero = slot_of(zero_key);  // row * S (column 0 of a replicated row): valid as a relative word for every stream
                for(int sj{}; sj < S; ++sj)
                {
                    for(auto& sec: RS[static_cast<std::size_t>(sj)].sec)
                    {
                        for(auto& ph: sec)
                        {
                            auto translate = [&](rop& o)
                            {
                                if(o.bubble) { return; }
                                if(o.opcode == PE_OP_DOT || o.opcode == PE_OP_CDOT)
                                {
                                    o.dst = rel_of(o.dst, sj);
                                    o.scale = (o.flags & PE_F_SCALE) ? rel_of(o.scale, sj) : pr.r_zero;
                                    for(auto& k: o.sre) { k = rel_of(k, sj); }
                                    for(auto& k: o.sim) { k = rel_of(k, sj); }
                                    for(auto& [a, b]: o.pp)
                                    {
                                        a = rel_of(a, sj);
                                        b = rel_of(b, sj);
                                    }
                                }
                                else if(o.opcode != 0xffu)
                                {
                                    for(auto& k: o.opnd) { k = rel_of(k, sj); }
                                }
                            };
                            for(auto& o: ph)
                            {
                                if(o.opcode == PE_OP_CROUT2)
                                {
                                    for(auto& sb: o.sub) { translate(sb); }
                                }
                                else
                                {
                                    translate(o);
                                }
                            }
                        }
                    }
                }
                // ---- load / store table
                pr.io.clear();
                auto add_io = [&](std::uint32_t slot, std::uint32_t kind, std::uint32_t fl, std::uint32_t src) { pr.io.push_back({slot | (kind << 16) | (fl << 20), src}); };
                std::vector<char> is_x(static_cast<std::size_t>(ps.n_lane) + 2, 0);
                if(cplx)
                {
                    for(int j{}; j < n; ++j) { is_x[xs[static_cast<std::size_t>(j)] & 0x1fffffffu] = 1; }
                }
                for(auto const& [k, v]: vals)
                {
                    std::uint32_t const sp{k >> 29}, sl{k & 0x1fffffffu};
                    std::uint32_t const slot{static_cast<std::uint32_t>(v.row * S + v.col)};
                    if(sp == I_CONST)
                    {
                        int const copies{v.replicated ? S : 1};
                        for(int cj{}; cj < copies; ++cj)
                        {
                            add_io(slot + static_cast<std::uint32_t>(cj), PE_IO_CONST, PE_IO_LOAD, sl);
                            if(cplx && k == zero_key) { add_io(slot + static_cast<std::uint32_t>(S + cj), PE_IO_CONST, PE_IO_LOAD, 0u); }
                        }
                    }
                    else if(sp == I_INST)
                    {
                        if(cplx) { add_io(slot, PE_IO_INSTX, PE_IO_LOAD, sl); }
                        else
                        {
                            add_io(slot, PE_IO_U, PE_IO_LOAD | (v.written ? PE_IO_STORE : 0u), sl);
                        }
                    }
                    else if(cplx && sl == static_cast<std::uint32_t>(PE_OPND_SLOT(ps.omega.op))) { add_io(slot, PE_IO_U, PE_IO_LOAD, 0u); }
                }
                pr.x_slot.resize(static_cast<std::size_t>(n));
                pr.x_opnd.resize(static_cast<std::size_t>(n));
                for(int j{}; j < n; ++j)
                {
                    // an unknown no op ever touches (cannot happen for a non-singular system) reads as the zero slot
                    std::uint32_t const key{xs[static_cast<std::size_t>(j)]};
                    bool const known{vals.find(unit_of(key).first) != vals.end()};
                    std::uint32_t const slot{known ? slot_of(key) : pr.r_zero};
                    pr.x_slot[static_cast<std::size_t>(j)] = slot;
                    if(cplx)
                    {
                        pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_U, static_cast<std::uint32_t>(1 + 2 * j));
                        if(known)
                        {
                            add_io(slot, PE_IO_U, PE_IO_STORE, static_cast<std::uint32_t>(1 + 2 * j));
                            add_io(slot + static_cast<std::uint32_t>(S), PE_IO_U, PE_IO_STORE, static_cast<std::uint32_t>(2 + 2 * j));
                        }
                    }
                    else
                    {
                        pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_U, static_cast<std::uint32_t>(j));
                    }
                }
                pr.omega_slot = cplx ? 0 : -1;
                pr.n_lane_slots = cplx ? 1 + 2 * n : 0;  // HBM lane space: omega + the solution
                pr.warps = 1;
                pr.packed_ig = -1;
                pr.built = true;
                for(auto const& st: steps)
                {
                    if(tree.level[static_cast<std::size_t>(st.node)] == 0) { ++pr.n_leaf_rows; }
                    else
                    {
                        ++pr.n_top_rows;
                    }
                }
            }

            void build_program(prog_mode mode)
            {
                auto& pr{out.prog[static_cast<int>(mode)]};
                pstate ps;
                ps.mode = mode;
                ps.cplx = (mode == prog_mode::AC);
                bool const ac{ps.cplx};
                int const n{num.unknowns()};
                int G{std::clamp(ac ? in.warps_ac : in.warps_real, 1, PE_MAX_WARPS)};
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

                // --- entries + active structure
                std::vector<int> ent_r, ent_c;
                std::vector<cplx> nv;
                std::vector<entry const*> orig;
                std::vector<std::map<int, int>> rows(static_cast<std::size_t>(n));  // col -> id (all entries ever)
                std::vector<std::set<int>> rowcols(static_cast<std::size_t>(n)), colrows(static_cast<std::size_t>(n));
                std::vector<std::vector<int>> adj(static_cast<std::size_t>(n));
                for(auto const& [rc, e]: ps.A)
                {
                    int const id{static_cast<int>(nv.size())};
                    ent_r.push_back(rc.first);
                    ent_c.push_back(rc.second);
                    nv.push_back(e.nom);
                    orig.push_back(&e);
                    rows[static_cast<std::size_t>(rc.first)][rc.second] = id;
                    rowcols[static_cast<std::size_t>(rc.first)].insert(rc.second);
                    colrows[static_cast<std::size_t>(rc.second)].insert(rc.first);
                    if(rc.first != rc.second)
                    {
                        adj[static_cast<std::size_t>(rc.first)].push_back(rc.second);
                        adj[static_cast<std::size_t>(rc.second)].push_back(rc.first);
                    }
                }
                for(auto& a: adj)
                {
                    std::sort(a.begin(), a.end());
                    a.erase(std::unique(a.begin(), a.end()), a.end());
                }

                // --- regions: elimination tree (resident: recursive dissection; v2: G leaves + one top)
                bool const resident{(ac ? in.resident_ac : in.resident_real) > 0};
                int const rS{resident ? (ac ? in.resident_ac : in.resident_real) : 0};
                etree tree;
                int n_leaves{};
                if(resident)
                {
                    tree = partition_tree(n, adj, rS, 3);
                    n_leaves = tree.n_leaves;
                    G = 1;
                }
                else
                {
                    int const min_leaf{12};
                    std::vector<int> region{partition(n, adj, G, min_leaf)};
                    for(int v: region) { n_leaves = std::max(n_leaves, v + 1); }
                    if(n_leaves < 2)
                    {
                        n_leaves = 0;
                        std::fill(region.begin(), region.end(), -1);
                        G = 1;
                    }
                    tree = flat_tree(region, n_leaves);
                }
                auto& region{tree.region};
                pr.n_leaves = static_cast<std::size_t>(n_leaves);

                // --- Markowitz ordering with threshold pivoting on nominal values, one region at a time
                int const R{static_cast<int>(tree.parent.size())};
                auto qi = [&](int idx) { return region[static_cast<std::size_t>(idx)]; };
                std::vector<std::set<std::pair<int, int>>> rq(static_cast<std::size_t>(R)), cq(static_cast<std::size_t>(R));
                for(int i{}; i < n; ++i)
                {
                    rq[static_cast<std::size_t>(qi(i))].insert({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                    cq[static_cast<std::size_t>(qi(i))].insert({static_cast<int>(colrows[static_cast<std::size_t>(i)].size()), i});
                }
                std::vector<char> row_done(static_cast<std::size_t>(n), 0), col_done(static_cast<std::size_t>(n), 0);
                std::vector<lu_step> steps;
                steps.reserve(static_cast<std::size_t>(n));
                bool singular{false};

                auto colmax = [&](int c)
                {
                    double m{};
                    for(int i: colrows[static_cast<std::size_t>(c)]) { m = std::max(m, std::abs(nv[static_cast<std::size_t>(rows[static_cast<std::size_t>(i)][c])])); }
                    return m;
                };

                for(int cur: tree.order)
                {
                    int const up{tree.parent[static_cast<std::size_t>(cur)]};
                    bool const top{up < 0};
                    auto& RQ{rq[static_cast<std::size_t>(cur)]};
                    auto& CQ{cq[static_cast<std::size_t>(cur)]};
                    for(;;)
                    {
                        long best_cost{std::numeric_limits<long>::max()};
                        int br{-1}, bc{-1};
                        double best_ratio{};
                        auto consider = [&](int r, int c, double cm)
                        {
                            if(qi(r) != cur || qi(c) != cur) { return; }
                            double const mag{std::abs(nv[static_cast<std::size_t>(rows[static_cast<std::size_t>(r)][c])])};
                            if(!(mag > 0.0) || mag < k_pivot_tau * cm) { return; }
                            long const cost{static_cast<long>(rowcols[static_cast<std::size_t>(r)].size() - 1) *
                                            static_cast<long>(colrows[static_cast<std::size_t>(c)].size() - 1)};
                            double const ratio{mag / cm};
                            if(cost < best_cost || (cost == best_cost && ratio > 4.0 * best_ratio))
                            {
                                best_cost = cost;
                                br = r;
                                bc = c;
                                best_ratio = ratio;
                            }
                        };
                        auto itc{CQ.begin()};
                        auto itr{RQ.begin()};
                        int examined{};
                        while(itc != CQ.end() || itr != RQ.end())
                        {
                            bool const take_col{itr == RQ.end() || (itc != CQ.end() && itc->first <= itr->first)};
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
                        if(br < 0) { break; }
                        lu_step st;
                        st.r = br;
                        st.c = bc;
                        st.leaf = (!resident && cur < n_leaves) ? cur : -1;
                        st.node = cur;
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
                            auto& q{cq[static_cast<std::size_t>(qi(j))]};
                            q.erase({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                            colrows[static_cast<std::size_t>(j)].erase(br);
                            if(j != bc) { q.insert({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j}); }
                        }
                        for(int i: st.lrows)
                        {
                            rq[static_cast<std::size_t>(qi(i))].erase({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                            rowcols[static_cast<std::size_t>(i)].erase(bc);
                        }
                        RQ.erase({static_cast<int>(rowcols[static_cast<std::size_t>(br)].size()), br});
                        rowcols[static_cast<std::size_t>(br)].clear();
                        colrows[static_cast<std::size_t>(bc)].clear();
                        row_done[static_cast<std::size_t>(br)] = 1;
                        col_done[static_cast<std::size_t>(bc)] = 1;
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
                                if(it == rowi.end())
                                {
                                    id = static_cast<int>(nv.size());
                                    nv.push_back({});
                                    ent_r.push_back(i);
                                    ent_c.push_back(j);
                                    orig.push_back(nullptr);
                                    rowi[j] = id;
                                    rowcols[static_cast<std::size_t>(i)].insert(j);
                                    auto& q{cq[static_cast<std::size_t>(qi(j))]};
                                    q.erase({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                                    colrows[static_cast<std::size_t>(j)].insert(i);
                                    q.insert({static_cast<int>(colrows[static_cast<std::size_t>(j)].size()), j});
                                }
                                else
                                {
                                    id = it->second;
                                }
                                nv[static_cast<std::size_t>(id)] -= l * nv[static_cast<std::size_t>(st.u_ent[b])];
                                st.t_ent.push_back(id);
                            }
                            rq[static_cast<std::size_t>(qi(i))].insert({static_cast<int>(rowcols[static_cast<std::size_t>(i)].size()), i});
                        }
                        steps.push_back(std::move(st));
                    }
                    if(!top)
                    {
                        // rows / columns of this region that found no acceptable pivot inside it are promoted to its parent
                        for(auto const& [cnt, i]: RQ) { rq[static_cast<std::size_t>(up)].insert({cnt, i}); }
                        for(auto const& [cnt, j]: CQ) { cq[static_cast<std::size_t>(up)].insert({cnt, j}); }
                        // indices whose row AND column are still active simply change region; a half-eliminated index
                        // (row done, column not, or vice versa) also moves: the restriction only looks at active ones
                        for(auto const& [cnt, i]: RQ) { region[static_cast<std::size_t>(i)] = up; }
                        for(auto const& [cnt, j]: CQ) { region[static_cast<std::size_t>(j)] = up; }
                        RQ.clear();
                        CQ.clear();
                    }
                }
                if(static_cast<int>(steps.size()) < n) { singular = true; }
                pr.structurally_singular = singular;
                pr.nnz_lu = nv.size();

                // --- guard elision: which pivots need the run-time guard on their column of L (PE_F_GUARD).  A node row that carries nothing but
                // two-terminal positive conductances (R, and C's companion in the real-valued modes) is symmetric and diagonally
                // dominant whatever the per-instance values are, eliminating such a row on its diagonal leaves the rows it updates
                // that way, and Gaussian elimination without pivoting has growth <= 2 on such a block (Wilkinson): no order can be
                // better than the static one there, the test would only cost time in the hot loop.  Every other pivot -- branch
                // rows, rows touched by any other element, rows such a pivot has updated -- keeps the guard.
                {
                    std::vector<char> clean(static_cast<std::size_t>(n), 0);
                    if(!ac && !in.guard_all)
                    {
                        for(int i{}; i < num.n_nodes; ++i) { clean[static_cast<std::size_t>(i)] = 1; }
                        for(auto const& e: nl.elems)
                        {
                            if(e.d == nullptr || e.d->code == E_RES || e.d->code == E_CAP) { continue; }
                            for(int p{}; p < e.d->pins; ++p)
                            {
                                int const u{nidx(e.pin_node[p])};
                                if(u >= 0) { clean[static_cast<std::size_t>(u)] = 0; }
                            }
                        }
                    }
                    pr.n_guarded = 0;
                    for(auto& st: steps)
                    {
                        bool const safe{st.r == st.c && clean[static_cast<std::size_t>(st.r)] != 0};
                        st.guard = !safe;
                        if(st.guard)
                        {
                            if(!st.lrows.empty()) { ++pr.n_guarded; }
                            // the rows it updates are no longer what the elements stamped (a pivot row that holds nothing but
                            // the pivot -- a grounded voltage source -- updates no matrix entry)
                            if(!st.ucols.empty())
                            {
                                for(int const i: st.lrows) { clean[static_cast<std::size_t>(i)] = 0; }
                            }
                        }
                    }
                }

                // --- operand translation to the kernel's spaces
                std::uint32_t const lane0{ac ? 0u : static_cast<std::uint32_t>(out.n_inst_slots)};
                auto xl = [&](std::uint32_t o) -> std::uint32_t
                {
                    std::uint32_t const neg{o & PE_OPND_NEG};
                    std::uint32_t const sp{(o >> 29) & 3u};
                    std::uint32_t const slot{o & 0x1fffffffu};
                    if(sp == I_CONST) { return neg | PE_OPND(PE_SP_CONST, slot); }
                    if(sp == I_INST) { return neg | (ac ? PE_OPND(PE_SP_INSTX, slot) : PE_OPND(PE_SP_U, slot)); }
                    return neg | PE_OPND(PE_SP_U, slot + lane0);
                };
                auto emit_value_op = [&](stream& S, op_t const& o)
                {
                    S.w.push_back(o.w[0]);
                    for(std::size_t k{1}; k < o.w.size(); ++k) { S.w.push_back(xl(o.w[k])); }
                };

                std::vector<stream> prep_s(static_cast<std::size_t>(G)), step_s(static_cast<std::size_t>(G)), iter_s(static_cast<std::size_t>(G));
                auto distribute = [&](std::vector<op_t> const& ops, std::vector<stream>& dst)
                {
                    for(std::size_t k{}; k < ops.size(); ++k) { emit_value_op(dst[k % static_cast<std::size_t>(G)], ops[k]); }
                };
                // PREP always addresses the instance workspace (it runs in a real-mode launch)
                if(!ac) { distribute(prep_ops, prep_s); }
                distribute(ps.step, step_s);
                distribute(ps.head, iter_s);

                auto finish = [&]()
                {
                    // word pool + section offsets
                    pr.words.clear();
                    auto place = [&](std::vector<stream>& ss, pe_b200_section& sec, bool present)
                    {
                        for(int g{}; g < PE_MAX_WARPS; ++g) { sec.off[g] = PE_NO_SECTION; }
                        if(!present) { return; }
                        for(int g{}; g < G; ++g)
                        {
                            sec.off[g] = static_cast<std::uint32_t>(pr.words.size());
                            auto& w{ss[static_cast<std::size_t>(g)].w};
                            pr.words.insert(pr.words.end(), w.begin(), w.end());
                            pr.words.push_back(PE_OP_END);
                        }
                    };
                    bool const has_prep{!ac && !prep_ops.empty()};
                    place(prep_s, pr.prep, has_prep);
                    place(step_s, pr.step, !ps.step.empty());
                    for(auto const& s: iter_s) { pr.max_warp_words = std::max(pr.max_warp_words, s.w.size()); }
                    place(iter_s, pr.iter, true);
                    pr.warps = G;
                    pr.n_lane_slots = ps.n_lane;
                    pr.built = true;
                };

                if(singular && resident)
                {
                    resident_singular(pr, ps, rS, n);
                    return;
                }
                if(singular)
                {
                    for(auto& s: iter_s)
                    {
                        s.w.clear();
                        s.w.push_back(0xffu);  // unknown opcode -> every lane reports PE_ST_SINGULAR
                    }
                    pr.x_opnd.assign(static_cast<std::size_t>(n), PE_OPND(PE_SP_CONST, 0));
                    finish();
                    return;
                }

                // --- symbolic structure in row-wise (dot product) form
                int const w{ps.cplx ? 2 : 1};
                std::size_t const n_ent{nv.size()};
                std::vector<int> pos_r(static_cast<std::size_t>(n), -1), pos_c(static_cast<std::size_t>(n), -1);
                for(std::size_t k{}; k < steps.size(); ++k)
                {
                    pos_r[static_cast<std::size_t>(steps[k].r)] = static_cast<int>(k);
                    pos_c[static_cast<std::size_t>(steps[k].c)] = static_cast<int>(k);
                }
                std::vector<std::vector<upd>> pairs(n_ent);                          // entry -> updates, in step order
                std::vector<std::vector<upd>> ypairs(static_cast<std::size_t>(n));   // row -> rhs updates (u = pivot row)
                std::vector<char> ynz(static_cast<std::size_t>(n), 0);
                for(int i{}; i < n; ++i) { ynz[static_cast<std::size_t>(i)] = !(ps.Z[static_cast<std::size_t>(i)].re.empty() && ps.Z[static_cast<std::size_t>(i)].im.empty()); }
                for(auto const& st: steps)
                {
                    std::size_t const nu{st.ucols.size()};
                    for(std::size_t a{}; a < st.lrows.size(); ++a)
                    {
                        int const i{st.lrows[a]};
                        for(std::size_t b{}; b < nu; ++b) { pairs[static_cast<std::size_t>(st.t_ent[a * nu + b])].push_back({st.l_ent[a], st.u_ent[b], st.leaf, st.node}); }
                        if(ynz[static_cast<std::size_t>(st.r)])
                        {
                            ypairs[static_cast<std::size_t>(i)].push_back({st.l_ent[a], st.r, st.leaf, st.node});
                            ynz[static_cast<std::size_t>(i)] = 1;
                        }
                    }
                }

                if(resident)
                {
                    resident_ctx rc{ps, pr, tree, rS, steps, pairs, ypairs, ynz, orig, n, n_ent};
                    emit_resident(rc);
                    for(auto const& p: pairs) { pr.n_fma += p.size(); }
                    for(auto const& p: ypairs) { pr.n_fma += p.size(); }
                    for(auto const& st: steps) { pr.n_fma += st.ucols.size(); }
                    return;
                }

                // --- slots
                std::vector<int> eslot(n_ent, -1), yslot(static_cast<std::size_t>(n), -1);
                auto new_lane = [&]()
                {
                    int const s{ps.n_lane};
                    ps.n_lane += w;
                    return s;
                };
                auto uslot_e = [&](int id) -> std::uint32_t
                {
                    if(eslot[static_cast<std::size_t>(id)] < 0) { eslot[static_cast<std::size_t>(id)] = new_lane(); }
                    return static_cast<std::uint32_t>(eslot[static_cast<std::size_t>(id)]) + lane0;
                };
                auto uslot_y = [&](int row) -> std::uint32_t
                {
                    if(yslot[static_cast<std::size_t>(row)] < 0) { yslot[static_cast<std::size_t>(row)] = new_lane(); }
                    return static_cast<std::uint32_t>(yslot[static_cast<std::size_t>(row)]) + lane0;
                };
                // solution: real -> the persistent INST slots [0, n); AC -> fresh lane slots (contiguous: the driver
                // downloads them as one block)
                std::vector<std::uint32_t> xs(static_cast<std::size_t>(n));
                pr.x_opnd.resize(static_cast<std::size_t>(n));

                auto emit_dot = [&](stream& S, std::uint32_t dst, std::uint32_t flags, std::uint32_t scale, std::vector<std::uint32_t> const& sre,
                                    std::vector<std::uint32_t> const& sim, std::vector<std::pair<std::uint32_t, std::uint32_t>> const& pp)
                {
                    std::size_t const max_src{ps.cplx ? 65535u : PE_DOT_MAX_SRC};
                    std::size_t const max_pair{ps.cplx ? 65535u : PE_DOT_MAX_PAIR};
                    std::size_t ri{}, ii{}, pi{};
                    bool first{true};
                    for(;;)
                    {
                        std::size_t const carry{first ? 0u : 1u};
                        std::size_t const nr{std::min(sre.size() - ri, max_src - carry)};
                        std::size_t const ni{std::min(sim.size() - ii, max_src - carry)};
                        std::size_t const np{std::min(pp.size() - pi, max_pair)};
                        bool const last{ri + nr == sre.size() && ii + ni == sim.size() && pi + np == pp.size()};
                        std::uint32_t d{dst};
                        if(!last || !first)
                        {
                            if(S.tmp_slot < 0) { S.tmp_slot = new_lane(); }
                        }
                        std::uint32_t const tmp{S.tmp_slot < 0 ? 0u : static_cast<std::uint32_t>(S.tmp_slot) + lane0};
                        if(!last) { d = tmp; }
                        std::uint32_t const f{last ? flags : 0u};
                        if(ps.cplx)
                        {
                            S.w.push_back(PE_OP_CDOT | (f << 8) | (static_cast<std::uint32_t>(np) << 16));
                            S.w.push_back(static_cast<std::uint32_t>(nr + carry) | (static_cast<std::uint32_t>(ni + carry) << 16));
                        }
                        else
                        {
                            S.w.push_back(PE_OP_DOT | (f << 8) | (static_cast<std::uint32_t>(nr + carry) << 16) | (static_cast<std::uint32_t>(np) << 24));
                        }
                        S.w.push_back(d);
                        if(f & PE_F_SCALE) { S.w.push_back(scale); }
                        if(carry) { S.w.push_back(PE_OPND(PE_SP_U, tmp)); }
                        for(std::size_t k{}; k < nr; ++k) { S.w.push_back(sre[ri + k]); }
                        if(ps.cplx)
                        {
                            if(carry) { S.w.push_back(PE_OPND(PE_SP_U, tmp + 1)); }
                            for(std::size_t k{}; k < ni; ++k) { S.w.push_back(sim[ii + k]); }
                        }
                        for(std::size_t k{}; k < np; ++k)
                        {
                            S.w.push_back(pp[pi + k].first);
                            S.w.push_back(pp[pi + k].second);
                        }
                        ri += nr;
                        ii += ni;
                        pi += np;
                        first = false;
                        if(last) { break; }
                    }
                };

                std::vector<std::uint32_t> sre, sim;
                std::vector<std::pair<std::uint32_t, std::uint32_t>> pp;
                auto load_sources = [&](entry const* e)
                {
                    sre.clear();
                    sim.clear();
                    if(e == nullptr) { return; }
                    for(auto o: e->re) { sre.push_back(xl(o)); }
                    for(auto o: e->im) { sim.push_back(xl(o)); }
                };
                // Updates generated by leaf steps but landing on an entry owned by a top step are summed by the leaf's own
                // warp into a contribution slot (the "update matrix" of the multifrontal method); the top step then only
                // adds the contributions.  contrib[(target, leaf)] -> slot; target = entry id, or n_ent + row for the rhs.
                std::map<std::pair<std::int64_t, int>, std::uint32_t> contrib;
                std::vector<std::vector<std::pair<std::int64_t, std::uint32_t>>> leaf_contribs(static_cast<std::size_t>(std::max(n_leaves, 1)));
                auto load_updates = [&](std::int64_t target, std::vector<upd> const& ups, bool owner_is_top, bool rhs)
                {
                    // fills pp with the updates the owner applies itself and appends contribution slots to sre / sim
                    pp.clear();
                    std::map<int, std::uint32_t> seen;
                    for(auto const& u: ups)
                    {
                        if(owner_is_top && u.leaf >= 0)
                        {
                            if(seen.find(u.leaf) == seen.end())
                            {
                                std::uint32_t const slot{static_cast<std::uint32_t>(new_lane()) + lane0};
                                seen[u.leaf] = slot;
                                contrib[{target, u.leaf}] = slot;
                                leaf_contribs[static_cast<std::size_t>(u.leaf)].push_back({target, slot});
                                sre.push_back(PE_OPND(PE_SP_U, slot));
                                if(ps.cplx) { sim.push_back(PE_OPND(PE_SP_U, slot + 1)); }
                            }
                            continue;
                        }
                        pp.push_back({uslot_e(u.l), rhs ? uslot_y(u.u) : uslot_e(u.u)});
                    }
                };

                // Crout step k: pivot, row k of U, column k of L (scaled by the fresh pivot reciprocal), rhs entry k
                auto emit_fwd_step = [&](stream& S, std::size_t k)
                {
                    auto const& st{steps[k]};
                    bool const top{st.leaf < 0};
                    load_sources(orig[static_cast<std::size_t>(st.piv)]);
                    load_updates(st.piv, pairs[static_cast<std::size_t>(st.piv)], top, false);
                    emit_dot(S, uslot_e(st.piv), PE_F_RECIP, 0, sre, sim, pp);
                    for(int e: st.u_ent)
                    {
                        load_sources(orig[static_cast<std::size_t>(e)]);
                        load_updates(e, pairs[static_cast<std::size_t>(e)], top, false);
                        emit_dot(S, uslot_e(e), 0, 0, sre, sim, pp);
                    }
                    for(int e: st.l_ent)
                    {
                        load_sources(orig[static_cast<std::size_t>(e)]);
                        load_updates(e, pairs[static_cast<std::size_t>(e)], top, false);
                        emit_dot(S, uslot_e(e), PE_F_SCALE | (st.guard ? PE_F_GUARD : 0u), uslot_e(st.piv), sre, sim, pp);  // an entry of L
                    }
                    if(ynz[static_cast<std::size_t>(st.r)])
                    {
                        load_sources(&ps.Z[static_cast<std::size_t>(st.r)]);
                        load_updates(static_cast<std::int64_t>(n_ent) + st.r, ypairs[static_cast<std::size_t>(st.r)], top, true);
                        emit_dot(S, uslot_y(st.r), 0, 0, sre, sim, pp);
                    }
                };
                // contribution of one leaf to one top-owned target: - sum of that leaf's updates
                auto emit_contrib = [&](stream& S, int leaf, std::int64_t target, std::uint32_t slot)
                {
                    bool const rhs{target >= static_cast<std::int64_t>(n_ent)};
                    auto const& ups{rhs ? ypairs[static_cast<std::size_t>(target - static_cast<std::int64_t>(n_ent))] : pairs[static_cast<std::size_t>(target)]};
                    sre.clear();
                    sim.clear();
                    pp.clear();
                    for(auto const& u: ups)
                    {
                        if(u.leaf == leaf) { pp.push_back({uslot_e(u.l), rhs ? uslot_y(u.u) : uslot_e(u.u)}); }
                    }
                    emit_dot(S, slot, 0, 0, sre, sim, pp);
                };
                auto emit_back_row = [&](stream& S, std::size_t k)
                {
                    auto const& st{steps[k]};
                    sre.clear();
                    sim.clear();
                    pp.clear();
                    if(ynz[static_cast<std::size_t>(st.r)])
                    {
                        sre.push_back(PE_OPND(PE_SP_U, uslot_y(st.r)));
                        if(ps.cplx) { sim.push_back(PE_OPND(PE_SP_U, uslot_y(st.r) + 1)); }
                    }
                    for(std::size_t b{}; b < st.ucols.size(); ++b) { pp.push_back({uslot_e(st.u_ent[b]), xs[static_cast<std::size_t>(st.ucols[b])]}); }
                    std::uint32_t flags{};
                    if(!sre.empty() || !pp.empty()) { flags |= PE_F_SCALE; }
                    if(!ps.cplx) { flags |= (st.c >= num.n_nodes) ? PE_F_CHECK_I : PE_F_CHECK_V; }
                    emit_dot(S, xs[static_cast<std::size_t>(st.c)], flags, uslot_e(st.piv), sre, sim, pp);
                };

                // --- schedule: leaves -> warps (longest processing time first), top -> warp 0
                std::vector<std::size_t> leaf_cost(static_cast<std::size_t>(std::max(n_leaves, 1)), 0);
                for(auto const& st: steps)
                {
                    if(st.leaf >= 0) { leaf_cost[static_cast<std::size_t>(st.leaf)] += 4 + st.lrows.size() * (st.ucols.size() + 1) + st.ucols.size(); }
                }
                std::vector<int> leaf_warp(static_cast<std::size_t>(std::max(n_leaves, 1)), 0);
                {
                    std::vector<int> order(static_cast<std::size_t>(n_leaves));
                    std::iota(order.begin(), order.end(), 0);
                    std::sort(order.begin(), order.end(), [&](int a, int b) { return leaf_cost[static_cast<std::size_t>(a)] > leaf_cost[static_cast<std::size_t>(b)]; });
                    std::vector<std::size_t> load(static_cast<std::size_t>(G), 0);
                    for(int lf: order)
                    {
                        int const g{static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin())};
                        leaf_warp[static_cast<std::size_t>(lf)] = g;
                        load[static_cast<std::size_t>(g)] += leaf_cost[static_cast<std::size_t>(lf)];
                    }
                }
                // x slots (AC) are allocated before any row is emitted so that they stay contiguous
                for(int j{}; j < n; ++j)
                {
                    if(ps.cplx)
                    {
                        xs[static_cast<std::size_t>(j)] = static_cast<std::uint32_t>(new_lane());
                        pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_U, xs[static_cast<std::size_t>(j)]);
                    }
                    else
                    {
                        xs[static_cast<std::size_t>(j)] = static_cast<std::uint32_t>(j);
                        pr.x_opnd[static_cast<std::size_t>(j)] = PE_OPND(PE_SP_U, static_cast<std::uint32_t>(j));
                    }
                }
                auto bar_all = [&]()
                {
                    if(G > 1)
                    {
                        for(auto& s: iter_s) { s.w.push_back(PE_OP_BAR); }
                    }
                };
                bar_all();  // device evaluation / source values are visible to every warp
                // The top steps are emitted into a side stream first: that is what discovers which contributions the
                // leaves owe; the leaf streams (which run earlier on the device) are completed afterwards.
                stream top_s;
                top_s.tmp_slot = -1;
                for(std::size_t k{}; k < steps.size(); ++k)
                {
                    if(steps[k].leaf >= 0)
                    {
                        emit_fwd_step(iter_s[static_cast<std::size_t>(leaf_warp[static_cast<std::size_t>(steps[k].leaf)])], k);
                        ++pr.n_leaf_rows;
                    }
                }
                for(std::size_t k{}; k < steps.size(); ++k)
                {
                    if(steps[k].leaf < 0)
                    {
                        emit_fwd_step(top_s, k);
                        ++pr.n_top_rows;
                    }
                }
                for(std::size_t k{steps.size()}; k-- > 0;)
                {
                    if(steps[k].leaf < 0) { emit_back_row(top_s, k); }
                }
                for(int lf{}; lf < n_leaves; ++lf)
                {
                    auto& S{iter_s[static_cast<std::size_t>(leaf_warp[static_cast<std::size_t>(lf)])]};
                    for(auto const& [target, slot]: leaf_contribs[static_cast<std::size_t>(lf)]) { emit_contrib(S, lf, target, slot); }
                }
                bar_all();
                iter_s[0].w.insert(iter_s[0].w.end(), top_s.w.begin(), top_s.w.end());
                bar_all();
                for(std::size_t k{steps.size()}; k-- > 0;)
                {
                    if(steps[k].leaf >= 0) { emit_back_row(iter_s[static_cast<std::size_t>(leaf_warp[static_cast<std::size_t>(steps[k].leaf)])], k); }
                }
                for(auto const& p: pairs) { pr.n_fma += p.size(); }
                for(auto const& p: ypairs) { pr.n_fma += p.size(); }
                for(auto const& st: steps) { pr.n_fma += st.ucols.size(); }
                finish();
            }
        };
    }  // namespace

    // Pack the (globally aligned) stream programs into warp vector ops for CTAs of rS * ig threads
    // (pe_b200_program.h): a warp carries C = 32 / ig consecutive streams.  Per position the C ops become rows of operand
    // words; a row whose words agree over every active column is stored once (warp-uniform), otherwise once per column.
    // Missing operands are padded with the zero slot (exact: x + (-0.0) = x, fma(-0, 0, x) = x).  Identical warp
    // streams (isomorphic sub-trees) share one copy of the words.
    void program::pack(int ig)
    {
        if(!resident || ig == packed_ig) { return; }
        int const C{std::max(1, 32 / std::max(ig, 1))};
        int const W{std::max(1, (rS + C - 1) / C)};
        n_warps = W;
        packed_ig = ig;
        words.clear();
        sec_off.assign(static_cast<std::size_t>(6 * W), PE_NO_SECTION);  // [main | side][section][warp]
        std::uint32_t const zero{r_zero};
        std::uint32_t const pad_src{(zero | PE_R_NEG) | ((zero | PE_R_NEG) << 16)};
        std::uint32_t const pad_pair{zero | (zero << 16)};
        max_warp_words = 0;
        std::map<std::pair<std::vector<std::uint32_t>, std::vector<std::uint32_t>>, std::pair<std::uint32_t, std::uint32_t>> seen;  // warp streams -> offsets
        std::vector<std::uint32_t> ww, sw;  // main stream (headers, masks, uniform rows) and side stream (per-column rows)
        std::vector<std::vector<std::uint32_t>> rows;  // rows[r][c]
        for(int sec{}; sec < 3; ++sec)
        {
            if(!has_sec[sec] && sec != 2) { continue; }
            for(int wv{}; wv < W; ++wv)
            {
                ww.clear();
                sw.clear();
                // C == 1 only: prefetch bitmaps of the current line, and which workspace rows are warm (touched by one
                // of the last 64 ops of this stream: they sit in L1 / L2 anyway)
                std::size_t line_bm{};
                std::size_t op_counter{};
                static std::size_t const pf_window{std::getenv("PE_B200_PF_WINDOW") ? static_cast<std::size_t>(std::atoi(std::getenv("PE_B200_PF_WINDOW"))) : 64u};
                static bool const pf_dst{std::getenv("PE_B200_PF_DST") != nullptr};
                std::map<std::uint32_t, std::size_t> last_touch;
                auto line_start = [&]()
                {
                    if(ww.size() % 32 == 0)
                    {
                        line_bm = ww.size();
                        ww.push_back(0u);
                        ww.push_back(0u);
                    }
                };
                // true when the row the operand field names is cold and not the zero pad
                auto touch = [&](std::uint32_t field, bool write) -> bool
                {
                    std::uint32_t const rel{field & 0x7fffu};
                    std::uint32_t const sm{static_cast<std::uint32_t>(rS - 1)};
                    std::uint32_t const slot{(rel & ~sm) | ((rel + static_cast<std::uint32_t>(wv)) & sm)};
                    auto it{last_touch.find(slot)};
                    bool const cold{it == last_touch.end() || op_counter - it->second > pf_window};
                    last_touch[slot] = op_counter;
                    return cold && (!write || pf_dst) && rel != zero;
                };
                std::size_t n_ph{};
                for(int c{}; c < C; ++c)
                {
                    int const sj{wv * C + c};
                    if(sj < rS) { n_ph = std::max(n_ph, rstreams[static_cast<std::size_t>(sj)].sec[sec].size()); }
                }
                for(std::size_t ph{}; ph < n_ph; ++ph)
                {
                    std::size_t n_pos{};
                    for(int c{}; c < C; ++c)
                    {
                        int const sj{wv * C + c};
                        if(sj < rS && ph < rstreams[static_cast<std::size_t>(sj)].sec[sec].size()) { n_pos = std::max(n_pos, rstreams[static_cast<std::size_t>(sj)].sec[sec][ph].size()); }
                    }
                    for(std::size_t pos{}; pos < n_pos; ++pos)
                    {
                        std::vector<rop const*> act(static_cast<std::size_t>(C), nullptr);
                        rop const* any{nullptr};
                        for(int c{}; c < C; ++c)
                        {
                            int const sj{wv * C + c};
                            if(sj >= rS || ph >= rstreams[static_cast<std::size_t>(sj)].sec[sec].size()) { continue; }
                            auto const& l{rstreams[static_cast<std::size_t>(sj)].sec[sec][ph]};
                            if(pos < l.size() && !l[pos].bubble)
                            {
                                act[static_cast<std::size_t>(c)] = &l[pos];
                                any = &l[pos];
                            }
                        }
                        if(any == nullptr) { continue; }  // no stream of this warp works at this position
                        std::uint32_t const opc{any->opcode};
                        rows.clear();
                        std::uint32_t h0{};
                        if(opc == PE_OP_DOT || opc == PE_OP_CDOT)
                        {
                            bool const cx{opc == PE_OP_CDOT};
                            std::size_t na{}, ni{}, nb{};
                            std::uint32_t ufl{};
                            for(auto const* o: act)
                            {
                                if(o == nullptr) { continue; }
                                na = std::max(na, (o->sre.size() + 1) / 2);
                                ni = std::max(ni, (o->sim.size() + 1) / 2);
                                nb = std::max(nb, o->pp.size());
                                ufl |= o->flags;
                            }
                            h0 = opc | (static_cast<std::uint32_t>(na) << 8) | (static_cast<std::uint32_t>(cx ? ni : 0) << 13) | (static_cast<std::uint32_t>(nb) << 18) | (ufl << 24);
                            auto add_row = [&](auto&& word_of)
                            {
                                rows.emplace_back(static_cast<std::size_t>(C));
                                for(int c{}; c < C; ++c) { rows.back()[static_cast<std::size_t>(c)] = word_of(act[static_cast<std::size_t>(c)]); }
                            };
                            add_row([&](rop const* o) { return o ? (o->dst | PE_R_ACTIVE | (o->flags << 16)) : 0u; });
                            add_row([&](rop const* o) { return (o && (o->flags & PE_F_SCALE)) ? o->scale : zero; });
                            auto src_rows = [&](bool im, std::size_t nr)
                            {
                                for(std::size_t r{}; r < nr; ++r)
                                {
                                    add_row(
                                        [&](rop const* o) -> std::uint32_t
                                        {
                                            if(o == nullptr) { return pad_src; }
                                            auto const& sv{im ? o->sim : o->sre};
                                            std::uint32_t const s0{2 * r < sv.size() ? sv[2 * r] : (zero | PE_R_NEG)};
                                            std::uint32_t const s1{2 * r + 1 < sv.size() ? sv[2 * r + 1] : (zero | PE_R_NEG)};
                                            return s0 | (s1 << 16);
                                        });
                                }
                            };
                            src_rows(false, na);
                            if(cx) { src_rows(true, ni); }
                            for(std::size_t r{}; r < nb; ++r)
                            {
                                add_row([&](rop const* o) { return (o && r < o->pp.size()) ? (o->pp[r].first | (o->pp[r].second << 16)) : pad_pair; });
                            }
                        }
                        else if(opc == PE_OP_CROUT2)
                        {
                            h0 = opc | (19u << 8);
                            for(std::size_t q{}; q < 6; ++q)
                            {
                                auto sub_of = [&](rop const* o) -> rop const* { return (o && q < o->sub.size()) ? &o->sub[q] : nullptr; };
                                rows.emplace_back(static_cast<std::size_t>(C));
                                for(int c{}; c < C; ++c)
                                {
                                    auto const* sb{sub_of(act[static_cast<std::size_t>(c)])};
                                    rows.back()[static_cast<std::size_t>(c)] = sb ? (sb->dst | PE_R_ACTIVE | (sb->flags << 16)) : 0u;
                                }
                                for(std::size_t r{}; r < (q == 0 ? 2u : 1u); ++r)
                                {
                                    rows.emplace_back(static_cast<std::size_t>(C));
                                    for(int c{}; c < C; ++c)
                                    {
                                        auto const* sb{sub_of(act[static_cast<std::size_t>(c)])};
                                        std::uint32_t s0{zero | PE_R_NEG}, s1{zero | PE_R_NEG};
                                        if(sb)
                                        {
                                            if(2 * r < sb->sre.size()) { s0 = sb->sre[2 * r]; }
                                            if(2 * r + 1 < sb->sre.size()) { s1 = sb->sre[2 * r + 1]; }
                                        }
                                        rows.back()[static_cast<std::size_t>(c)] = s0 | (s1 << 16);
                                    }
                                }
                                rows.emplace_back(static_cast<std::size_t>(C));
                                for(int c{}; c < C; ++c)
                                {
                                    auto const* sb{sub_of(act[static_cast<std::size_t>(c)])};
                                    rows.back()[static_cast<std::size_t>(c)] = (sb && !sb->pp.empty()) ? (sb->pp[0].first | (sb->pp[0].second << 16)) : pad_pair;
                                }
                            }
                        }
                        else
                        {
                            std::size_t nr{};
                            for(auto const* o: act)
                            {
                                if(o) { nr = std::max(nr, o->opnd.size()); }
                            }
                            h0 = opc | (static_cast<std::uint32_t>(nr) << 8) | (any->first_only ? 0x2000u : 0u);  // bit 13: first Newton iteration only
                            for(std::size_t r{}; r < nr; ++r)
                            {
                                rows.emplace_back(static_cast<std::size_t>(C));
                                for(int c{}; c < C; ++c)
                                {
                                    auto const* o{act[static_cast<std::size_t>(c)]};
                                    std::uint32_t wd{(o && r < o->opnd.size()) ? o->opnd[r] : zero};
                                    if(o && r == 0) { wd |= PE_R_VACTIVE; }
                                    rows.back()[static_cast<std::size_t>(c)] = wd;
                                }
                            }
                        }
                        // uniform rows: row 0 carries the activity bit, so it is uniform only when every column works; the
                        // other rows are never looked at by an idle column, which simply copies an active one
                        std::uint32_t mask{};
                        for(std::size_t r{}; r < rows.size() && r < 32; ++r)
                        {
                            auto& rw{rows[r]};
                            bool uni{true};
                            std::uint32_t first{};
                            bool have{false};
                            for(int c{}; c < C; ++c)
                            {
                                bool const on{act[static_cast<std::size_t>(c)] != nullptr};
                                // rows that carry an activity bit (row 0; every ctl row of a fused step) must reach an
                                // idle column as its own (inactive) word
                                bool const ctl_row{r == 0 || (opc == PE_OP_CROUT2 && r >= 4 && (r - 4) % 3 == 0)};
                                if(ctl_row && !on) { uni = false; }
                                if(!on) { continue; }
                                if(!have)
                                {
                                    first = rw[static_cast<std::size_t>(c)];
                                    have = true;
                                }
                                else if(rw[static_cast<std::size_t>(c)] != first) { uni = false; }
                            }
                            if(C == 1) { uni = true; }
                            if(!uni) { mask |= 1u << r; }
                            else
                            {
                                rw.assign(1, first);
                            }
                        }
                        if(C == 1)
                        {
                            // one stream per warp: an op never straddles two 32-word lines of the main stream, and every
                            // line starts with two bitmaps: which of its words name (in their low / high half) a
                            // workspace row worth prefetching (pe_b200_program.h)
                            line_start();
                            std::size_t const len{2 + rows.size()};
                            if(ww.size() % 32 + len > 32)
                            {
                                ww.push_back(PE_OP_SKIP);
                                while(ww.size() % 32 != 0) { ww.push_back(PE_OP_SKIP); }
                                line_start();
                            }
                            ++op_counter;
                            bool const dotlike{opc == PE_OP_DOT || opc == PE_OP_CDOT};
                            rop const* const o{any};
                            for(std::size_t r{}; r < rows.size(); ++r)
                            {
                                std::size_t const wi{(ww.size() + 2 + r) % 32};
                                std::uint32_t const word{rows[r][0]};
                                bool lo{}, hi{};
                                if(dotlike)
                                {
                                    if(r == 0)
                                    {
                                        // the destination: written here, warm from now on
                                        if(touch(word, true)) { ww[line_bm] |= 1u << wi; }
                                    }
                                    else if(r == 1) { lo = (o->flags & PE_F_SCALE) != 0u; }
                                    else
                                    {
                                        lo = hi = true;
                                    }
                                }
                                else
                                {
                                    lo = true;
                                }
                                if(lo && touch(word, false)) { ww[line_bm] |= 1u << wi; }
                                if(hi && touch(word >> 16, false)) { ww[line_bm + 1] |= 1u << wi; }
                            }
                        }
                        ww.push_back(h0);
                        ww.push_back(mask);
                        for(std::size_t r{}; r < rows.size(); ++r)
                        {
                            auto& dst{(mask >> r) & 1u ? sw : ww};
                            dst.insert(dst.end(), rows[r].begin(), rows[r].end());
                        }
                    }
                    if(C == 1) { line_start(); }
                    if(ph + 1 < n_ph) { ww.push_back(PE_OP_BAR); }
                }
                if(C == 1) { line_start(); }
                ww.push_back(PE_OP_END);
                max_warp_words = std::max(max_warp_words, ww.size() + sw.size());
                auto key{std::make_pair(ww, sw)};
                auto it{seen.find(key)};
                if(it == seen.end())
                {
                    // main streams start on a 128-byte line; both streams are read a few lines / rows ahead
                    while(words.size() % 32 != 0) { words.push_back(PE_OP_END); }
                    std::uint32_t const o_main{static_cast<std::uint32_t>(words.size())};
                    words.insert(words.end(), ww.begin(), ww.end());
                    for(int k{}; k < 128; ++k) { words.push_back(PE_OP_END); }
                    while(words.size() % 32 != 0) { words.push_back(PE_OP_END); }
                    std::uint32_t const o_side{static_cast<std::uint32_t>(words.size())};
                    words.insert(words.end(), sw.begin(), sw.end());
                    for(int k{}; k < 8 * C; ++k) { words.push_back(0u); }
                    it = seen.emplace(std::move(key), std::make_pair(o_main, o_side)).first;
                }
                sec_off[static_cast<std::size_t>(sec * W + wv)] = it->second.first;
                sec_off[static_cast<std::size_t>(3 * W + sec * W + wv)] = it->second.second;
            }
        }
        for(int k{}; k < 64; ++k) { words.push_back(PE_OP_END); }
    }

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
        for(int m{}; m < static_cast<int>(prog_mode::COUNT); ++m)
        {
            if(static_cast<prog_mode>(m) != prog_mode::AC) { out->n_real_lane_slots = std::max(out->n_real_lane_slots, out->prog[static_cast<std::size_t>(m)].n_lane_slots); }
        }
        return out;
    }
}  // namespace pe_b200
