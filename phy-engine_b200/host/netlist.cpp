// netlist.cpp — element catalogue, netlist construction and the prepare() numbering.
// Follows: dll_api.h:51-135 (codes), src/dll_main.cpp:1707-1956 (positional properties), the per-model
// set_attribute_define / get_attribute_name_define tables, circuits/circuit.h:481-540 (numbering).
#include <cctype>
#include <cstring>
#include <numbers>

#include "pe_host.hpp"

namespace pe_b200
{
    namespace
    {
        // attribute order = the idx of set_attribute(idx, ...) in the reference model headers
        constexpr elem_desc k_descs[] = {
            {E_RES, "Resistance", 2, 0, 1, {"R"}, {10.0}, 1, false},
            {E_CAP, "Capacitor", 2, 0, 1, {"C"}, {1e-5}, 1, false},
            {E_IND, "Inductor", 2, 1, 1, {"L"}, {1e-5}, 1, false},
            {E_VDC, "VDC", 2, 1, 1, {"V"}, {5.0}, 1, false},
            {E_VAC, "VAC", 2, 1, 3, {"Vp", "freq", "phase"}, {5.0, 50.0, 0.0}, 3, false},
            {E_IDC, "IDC", 2, 0, 1, {"I"}, {0.2}, 1, false},
            {E_IAC, "IAC", 2, 0, 3, {"Ip", "freq", "phase"}, {0.2, 50.0, 0.0}, 3, false},
            {E_VCCS, "VCCS", 4, 0, 1, {"G"}, {1.0}, 1, false},
            {E_VCVS, "VCVS", 4, 1, 1, {"Mu"}, {1.0}, 1, false},
            {E_CCCS, "CCCS", 4, 1, 1, {"alpha"}, {10.0}, 1, false},
            {E_CCVS, "CCVS", 4, 2, 1, {"r"}, {10.0}, 1, false},
            {E_SWITCH, "Single-Pole Switch", 2, 1, 1, {"Cut Through"}, {0.0}, 1, false},
            {E_PN,
             "PN Junction",
             2,
             0,
             10,
             {"Is", "N", "Isr", "Nr", "Temp", "Ibv", "Bv", "Bv_set", "Area", "tt"},
             {1e-14, 1.0, 0.0, 2.0, 27.0, 1e-3, 40.0, 1.0, 1.0, 0.0},
             9,
             true},
            // pins P, Q (primary), S, T (secondary); branches kP, kS; n = Vp / Vs (transformer.h:8-19)
            {E_XFMR, "Transformer", 4, 2, 1, {"n"}, {1.0}, 1, false},
            // pins P1, P2 (winding 1), S1, S2 (winding 2); branches k1, k2 (coupled_inductors.h:8-30)
            {E_KIND, "Coupled Inductors", 4, 2, 3, {"L1", "L2", "k"}, {1e-3, 1e-3, 0.99}, 3, false},
            // pins P, Q, S1, CT, S2; branches kP, kH1, kH2; n_total = Vp / V(S1 - S2) (transformer_center_tap.h:8-22)
            {E_XFMR_CT, "Transformer Center Tap", 5, 3, 1, {"n_total"}, {1.0}, 1, false},
            {E_OPAMP, "OpAmp", 4, 1, 1, {"mu"}, {1.0e5}, 1, false},
            // pins C+, C- (coil), A, B (contact); one branch; Von, Voff, Engaged (relay.h:8-22: non-linear, so solve() iterates)
            {E_RELAY, "Relay", 4, 1, 3, {"Von", "Voff", "Engaged"}, {5.0, 3.0, 0.0}, 2, true},
            // pins A, B (analog inputs), o (digital output): no MNA stamp; its state is vA >= vB at digital_clk (comparator.h:73-108)
            {E_CMP, "Comparator", 3, 0, 2, {"Ll", "Hl"}, {0.0, 5.0}, 2, false, true},
            // time-domain generators: pins +, -, one branch (generator/*.h); phase in radians
            {E_GEN_SAW, "Sawtooth Wave Generator", 2, 1, 4, {"Vh", "Vl", "freq", "phase"}, {5.0, 0.0, 1e3, 0.0}, 4, false},
            {E_GEN_SQUARE, "Square Wave Generator", 2, 1, 5, {"Vh", "Vl", "freq", "duty", "phase"}, {5.0, 0.0, 1e3, 0.5, 0.0}, 5, false},
            {E_GEN_PULSE, "Pulse Wave Generator", 2, 1, 7, {"Vh", "Vl", "freq", "duty", "phase", "tr", "tf"}, {5.0, 0.0, 1e3, 0.5, 0.0, 0.0, 0.0}, 7, false},
            {E_GEN_TRI, "Triangle Wave Generator", 2, 1, 4, {"Vh", "Vl", "freq", "phase"}, {5.0, 0.0, 1e3, 0.0}, 4, false},
            {E_NPN, "NPN BJT", 3, 0, 5, {"Is", "N", "BetaF", "Temp", "Area"}, {1e-16, 1.0, 100.0, 27.0, 1.0}, 5, true},
            {E_PNP, "PNP BJT", 3, 0, 5, {"Is", "N", "BetaF", "Temp", "Area"}, {1e-16, 1.0, 100.0, 27.0, 1.0}, 5, true},
            {E_NMOS, "NMOSFET", 3, 0, 3, {"Kp", "lambda", "Vth"}, {1e-3, 0.0, 1.0}, 3, true},
            {E_PMOS, "PMOSFET", 3, 0, 3, {"Kp", "lambda", "Vth"}, {1e-3, 0.0, 1.0}, 3, true},
            // four default PN junctions D1: A->+, D2: B->+, D3: - -> A, D4: - -> B; no attributes (full_bridge_rectifier.h:10-24)
            {E_BRIDGE, "Full Bridge Rectifier", 4, 0, 0, {}, {}, 0, true},
        };

        bool ieq(char a, char b) noexcept { return std::tolower(static_cast<unsigned char>(a)) == std::tolower(static_cast<unsigned char>(b)); }
    }  // namespace

    elem_desc const* find_desc(int code) noexcept
    {
        for(auto const& d: k_descs)
        {
            if(d.code == code) { return &d; }
        }
        return nullptr;
    }

    bool netlist::set_attribute(int ei, int idx, double v)
    {
        if(ei < 0 || ei >= static_cast<int>(elems.size())) { return false; }
        auto& e{elems[static_cast<std::size_t>(ei)]};
        if(idx < 0 || idx >= e.d->n_attr) { return false; }
        double stored{v};
        if(e.d->code == E_VAC || e.d->code == E_IAC)
        {
            // VAC.h:39-50: freq[Hz] -> omega, phase[deg] -> rad
            if(idx == 1) { stored = v * (2.0 * std::numbers::pi); }
            else if(idx == 2) { stored = v * (std::numbers::pi / 180.0); }
        }
        else if((e.d->code == E_SWITCH && idx == 0) || (e.d->code == E_PN && idx == 7))
        {
            stored = (v != 0.0) ? 1.0 : 0.0;  // boolean attributes (dll_main.cpp:2890)
        }
        e.attr[idx] = stored;
        return true;
    }

    int netlist::add_model(int code, double const* props, std::size_t* consumed)
    {
        auto const* d{find_desc(code)};
        if(consumed) { *consumed = 0; }
        if(d == nullptr) { return -1; }
        element e{};
        e.d = d;
        for(int i{}; i < d->n_attr; ++i) { e.attr[i] = d->attr_default[i]; }
        elems.push_back(e);
        int const ei{static_cast<int>(elems.size()) - 1};
        if(props != nullptr)
        {
            // the C ABI streams properties in attribute order (dll_main.cpp:1707-1956); VAC/IAC/switch/PN conversions are
            // the same as set_attribute's
            for(int i{}; i < d->n_cabi_props; ++i) { (void)set_attribute(ei, i, props[i]); }
            if(consumed) { *consumed = static_cast<std::size_t>(d->n_cabi_props); }
        }
        else if(d->code == E_VAC || d->code == E_IAC)
        {
            // defaults are stored raw in the reference struct (m_omega{50.0}, m_phase{0.0}: VAC.h:17-19)
            elems.back().attr[1] = 50.0;
            elems.back().attr[2] = 0.0;
        }
        return ei;
    }

    bool netlist::add_to_node(int ei, int pin, int node)
    {
        if(ei < 0 || ei >= static_cast<int>(elems.size())) { return false; }
        auto& e{elems[static_cast<std::size_t>(ei)]};
        if(pin < 0 || pin >= e.d->pins) { return false; }
        if(node < -1 || node >= n_created_nodes) { return false; }
        e.pin_node[pin] = node;
        return true;
    }

    int netlist::find_attribute(int ei, char const* name, std::size_t name_size) const
    {
        if(ei < 0 || ei >= static_cast<int>(elems.size()) || name == nullptr) { return -1; }
        auto const& e{elems[static_cast<std::size_t>(ei)]};
        for(int i{}; i < e.d->n_attr; ++i)
        {
            char const* an{e.d->attr_name[i]};
            if(std::strlen(an) != name_size) { continue; }
            bool ok{true};
            for(std::size_t k{}; k < name_size; ++k)
            {
                if(!ieq(an[k], name[k]))
                {
                    ok = false;
                    break;
                }
            }
            if(ok) { return i; }
        }
        return -1;
    }

    bool netlist::has_nonlinear() const
    {
        for(auto const& e: elems)
        {
            if(e.d->nonlinear) { return true; }
        }
        return false;
    }

    numbering make_numbering(netlist const& nl)
    {
        numbering nb;
        std::vector<int> analog_pins(static_cast<std::size_t>(nl.n_created_nodes), 0);
        for(auto const& e: nl.elems)
        {
            if(e.d->digital) { continue; }  // pins of digital models do not make a node analog (circuit.h:481-540)
            for(int p{}; p < e.d->pins; ++p)
            {
                if(e.pin_node[p] >= 0) { ++analog_pins[static_cast<std::size_t>(e.pin_node[p])]; }
            }
        }
        nb.node_index.assign(static_cast<std::size_t>(nl.n_created_nodes), -2);
        for(int n{}; n < nl.n_created_nodes; ++n)
        {
            if(analog_pins[static_cast<std::size_t>(n)] != 0) { nb.node_index[static_cast<std::size_t>(n)] = nb.n_nodes++; }
        }
        nb.branch0.resize(nl.elems.size());
        for(std::size_t i{}; i < nl.elems.size(); ++i)
        {
            nb.branch0[i] = nb.n_branches;
            nb.n_branches += nl.elems[i].d->branches;
        }
        return nb;
    }
}  // namespace pe_b200
