// c_api.cpp — the extern "C" surface declared in include/phy_engine_b200.h.
// Part 1 mirrors src/dll_main.cpp of the reference entry point by entry point (citations in the header); the
// handle is a pe_b200::circuit instead of a phy_engine::circult, everything a caller can observe through the
// C ABI (wire format, component order, prefix-sum layouts, return codes) is kept.
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>

#include "../../include/phy_engine_b200.h"
#include "pe_host.hpp"

namespace
{
    using namespace pe_b200;

    // netlist::model_pos of the reference: models live in 4 KiB chunks of sizeof(model_base) = 80 bytes
    // (netlist/netlist.h:20-21) -> 51 models per chunk; component i sits at (i % 51, i / 51).
    constexpr std::size_t k_chunk{4096 / 80};

    int elem_of(circuit const& c, std::size_t vec_pos, std::size_t chunk_pos)
    {
        if(vec_pos >= k_chunk) { return -1; }
        std::size_t const e{chunk_pos * k_chunk + vec_pos};
        return e < c.nl.elems.size() ? static_cast<int>(e) : -1;
    }

    int uf_find(int x, std::vector<int>& parent, std::vector<char>& visited)
    {
        if(parent[static_cast<std::size_t>(x)] != x) { parent[static_cast<std::size_t>(x)] = uf_find(parent[static_cast<std::size_t>(x)], parent, visited); }
        visited[static_cast<std::size_t>(x)] = 1;
        return parent[static_cast<std::size_t>(x)];
    }

    // build_netlist_from_wires (dll_main.cpp:1529-1705): nets = union-find over pin slots, ground forced root,
    // nodes created in ascending root-slot order.
    void wire_up(netlist& nl, int const* elements, int ele_size, int const* wires, int wire_count, std::vector<int> const& elem_of_ele)
    {
        std::vector<std::size_t> base(static_cast<std::size_t>(ele_size) + 1, 0), pins(static_cast<std::size_t>(ele_size), 0);
        for(int i{}; i < ele_size; ++i)
        {
            if(elements[i] && elem_of_ele[static_cast<std::size_t>(i)] >= 0)
            {
                pins[static_cast<std::size_t>(i)] = static_cast<std::size_t>(nl.elems[static_cast<std::size_t>(elem_of_ele[static_cast<std::size_t>(i)])].d->pins);
            }
            base[static_cast<std::size_t>(i) + 1] = base[static_cast<std::size_t>(i)] + pins[static_cast<std::size_t>(i)];
        }
        int const total{static_cast<int>(base[static_cast<std::size_t>(ele_size)])};
        int const GROUND{total};
        std::vector<int> parent(static_cast<std::size_t>(total) + 1);
        std::vector<char> visited(static_cast<std::size_t>(total) + 1, 0);
        for(int i{}; i <= total; ++i) { parent[static_cast<std::size_t>(i)] = i; }
        for(int w{}; w < wire_count; ++w)
        {
            int const e1{wires[w * 4]}, p1{wires[w * 4 + 1]}, e2{wires[w * 4 + 2]}, p2{wires[w * 4 + 3]};
            if(e1 < 0 || e2 < 0 || e1 >= ele_size || e2 >= ele_size) { continue; }
            int n1, n2;
            if(!elements[e1]) { n1 = GROUND; }
            else
            {
                if(p1 < 0 || static_cast<std::size_t>(p1) >= pins[static_cast<std::size_t>(e1)]) { continue; }
                n1 = static_cast<int>(base[static_cast<std::size_t>(e1)]) + p1;
            }
            if(!elements[e2]) { n2 = GROUND; }
            else
            {
                if(p2 < 0 || static_cast<std::size_t>(p2) >= pins[static_cast<std::size_t>(e2)]) { continue; }
                n2 = static_cast<int>(base[static_cast<std::size_t>(e2)]) + p2;
            }
            int const r1{uf_find(n1, parent, visited)};
            int const r2{uf_find(n2, parent, visited)};
            if(r1 != r2)
            {
                if(r1 == GROUND) { parent[static_cast<std::size_t>(r2)] = r1; }
                else if(r2 == GROUND) { parent[static_cast<std::size_t>(r1)] = r2; }
                else
                {
                    parent[static_cast<std::size_t>(r2)] = r1;
                }
            }
        }
        std::unordered_map<int, int> node_map;
        for(int i{}; i < total; ++i)
        {
            if(parent[static_cast<std::size_t>(i)] == i && visited[static_cast<std::size_t>(i)]) { node_map[i] = nl.create_node(); }
        }
        node_map[GROUND] = -1;
        for(int i{}; i < ele_size; ++i)
        {
            if(!elements[i] || elem_of_ele[static_cast<std::size_t>(i)] < 0) { continue; }
            for(std::size_t p{}; p < pins[static_cast<std::size_t>(i)]; ++p)
            {
                int const id{static_cast<int>(base[static_cast<std::size_t>(i)] + p)};
                if(!visited[static_cast<std::size_t>(id)]) { continue; }
                int const root{uf_find(id, parent, visited)};
                (void)nl.add_to_node(elem_of_ele[static_cast<std::size_t>(i)], static_cast<int>(p), node_map[root]);
            }
        }
    }

    double to_internal(int code, int idx, double v)
    {
        if(code == E_VAC || code == E_IAC)
        {
            if(idx == 1) { return v * (2.0 * 3.14159265358979323846264338327950288); }
            if(idx == 2) { return v * (3.14159265358979323846264338327950288 / 180.0); }
        }
        if((code == E_SWITCH && idx == 0) || (code == E_PN && idx == 7)) { return v != 0.0 ? 1.0 : 0.0; }
        return v;
    }

    int sample_impl(circuit& c, std::size_t* vec_pos, std::size_t* chunk_pos, std::size_t comp_size, double* voltage, std::size_t* vo, double* current, std::size_t* co)
    {
        bool const have{!c.x_host.empty()};
        for(std::size_t i{}; i < comp_size; ++i)
        {
            int const ei{elem_of(c, vec_pos[i], chunk_pos[i])};
            if(ei < 0) { continue; }
            auto const& e{c.nl.elems[static_cast<std::size_t>(ei)]};
            for(int j{}; j < e.d->pins; ++j)
            {
                double v{};
                int const node{e.pin_node[j]};
                if(have && node >= 0)
                {
                    int const u{c.num_host.node_index[static_cast<std::size_t>(node)]};
                    if(u >= 0) { v = c.x_host[static_cast<std::size_t>(u)]; }
                }
                voltage[static_cast<std::size_t>(j) + vo[i]] = v;
            }
            for(int j{}; j < e.d->branches; ++j)
            {
                double v{};
                if(have) { v = c.x_host[static_cast<std::size_t>(c.num_host.n_nodes + c.num_host.branch0[static_cast<std::size_t>(ei)] + j)]; }
                current[static_cast<std::size_t>(j) + co[i]] = v;
            }
        }
        return 0;
    }
}  // namespace

extern "C"
{
    char const* phy_engine_last_error(void) { return pe_b200::last_error(); }

    void phy_engine_clear_error(void) { pe_b200::set_last_error({}); }

    void phy_engine_string_free(char* s) { delete[] s; }

    void* create_circuit(int* elements, size_t ele_size, int* wires, size_t wires_size, double* properties, size_t** vec_pos, size_t** chunk_pos, size_t* comp_size)
    {
        set_last_error({});
        if(vec_pos == nullptr || chunk_pos == nullptr || comp_size == nullptr)
        {
            set_last_error("create_circuit: output pointers are null");
            return nullptr;
        }
        *vec_pos = nullptr;
        *chunk_pos = nullptr;
        *comp_size = 0;
        if(elements == nullptr || properties == nullptr)
        {
            set_last_error("create_circuit: elements/properties are null");
            return nullptr;
        }
        auto* c{new(std::nothrow) circuit{}};
        if(c == nullptr) { return nullptr; }
        // defaults of dll_main.cpp:2530-2535
        c->at = analyze_type::TR;
        c->tr.t_step = 1e-6;
        c->tr.t_stop = 1e-6;
        *vec_pos = static_cast<size_t*>(std::malloc((ele_size ? ele_size : 1) * sizeof(size_t)));
        *chunk_pos = static_cast<size_t*>(std::malloc((ele_size ? ele_size : 1) * sizeof(size_t)));
        std::vector<int> elem_of_ele(ele_size, -1);
        double const* prop{properties};
        std::size_t k{};
        for(std::size_t i{}; i < ele_size; ++i)
        {
            if(!elements[i]) { continue; }
            std::size_t used{};
            int const ei{c->nl.add_model(elements[i], prop, &used)};
            if(ei < 0)
            {
                set_last_error("create_circuit: element code " + std::to_string(elements[i]) + " at index " + std::to_string(i) +
                               " is outside the B200 hot path (SURVEY.md §8)");
                std::free(*vec_pos);
                std::free(*chunk_pos);
                *vec_pos = nullptr;
                *chunk_pos = nullptr;
                delete c;
                return nullptr;
            }
            prop += used;
            elem_of_ele[i] = ei;
            (*vec_pos)[k] = static_cast<std::size_t>(ei) % k_chunk;
            (*chunk_pos)[k] = static_cast<std::size_t>(ei) / k_chunk;
            ++k;
        }
        *comp_size = k;
        int const wire_count{static_cast<int>(wires_size / 4)};
        if(wires != nullptr && wire_count > 0) { wire_up(c->nl, elements, static_cast<int>(ele_size), wires, wire_count, elem_of_ele); }
        return c;
    }

    void* create_circuit_ex(int* elements,
                            size_t ele_size,
                            int* wires,
                            size_t wires_size,
                            double* properties,
                            char const* const*,
                            size_t const*,
                            size_t,
                            size_t const*,
                            size_t const*,
                            size_t** vec_pos,
                            size_t** chunk_pos,
                            size_t* comp_size)
    {
        return create_circuit(elements, ele_size, wires, wires_size, properties, vec_pos, chunk_pos, comp_size);
    }

    void destroy_circuit(void* circuit_ptr, size_t* vec_pos, size_t* chunk_pos)
    {
        delete static_cast<circuit*>(circuit_ptr);
        std::free(vec_pos);
        std::free(chunk_pos);
    }

    int circuit_set_analyze_type(void* p, uint32_t v)
    {
        if(p == nullptr) { return 1; }
        static_cast<circuit*>(p)->at = static_cast<analyze_type>(v);
        return 0;
    }

    int circuit_set_tr(void* p, double t_step, double t_stop)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->tr.t_step = t_step;
        c->tr.t_stop = t_stop;
        return 0;
    }

    int circuit_set_ac_omega(void* p, double omega)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->ac.sweep = sweep_type::single;  // dll_main.cpp:2162
        c->ac.omega = omega;
        return 0;
    }

    int circuit_set_temperature(void* p, double t)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->env.temperature = t;
        ++c->param_rev;
        return 0;
    }

    int circuit_set_tnom(void* p, double t)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->env.norm_temperature = t;
        ++c->param_rev;
        return 0;
    }

    int circuit_set_model_double_by_name(void* p, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double value)
    {
        if(p == nullptr || name == nullptr || name_size == 0) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        int const ei{elem_of(*c, vec_pos, chunk_pos)};
        if(ei < 0) { return 2; }
        int const idx{c->nl.find_attribute(ei, name, name_size)};
        if(idx < 0) { return 3; }
        (void)c->nl.set_attribute(ei, idx, value);
        ++c->param_rev;
        return 0;
    }

    int circuit_analyze(void* p)
    {
        if(p == nullptr) { return 1; }
        return static_cast<circuit*>(p)->analyze() ? 0 : 1;
    }

    // circult::digital_clk() (circuit.h:298-354) for the one digital model in scope, the comparator (comparator.h:73-108): the
    // state of its output node becomes vA >= vB of the last analog solution.  One instance: evaluated on the host from the
    // solution analyze() downloaded (the batch path does the same on the device: circuit_batch_digital_clk).  A comparator
    // whose output sits on an ANALOG node would drive it through an ideal source (digital -> analog, circuit.h:1015-1022):
    // that driver is not part of this path and the call says so instead of ticking silently.
    int circuit_digital_clk(void* p)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->digital_state.resize(static_cast<std::size_t>(c->nl.n_created_nodes), 2);  // 2 = indeterminate until first driven
        auto volt = [&](int node) -> double
        {
            if(node < 0 || c->x_host.empty()) { return 0.0; }
            int const u{c->num_host.node_index[static_cast<std::size_t>(node)]};
            return u >= 0 ? c->x_host[static_cast<std::size_t>(u)] : 0.0;
        };
        for(auto const& e: c->nl.elems)
        {
            if(e.d->code != E_CMP) { continue; }
            int const a{e.pin_node[0]}, b{e.pin_node[1]}, o{e.pin_node[2]};
            if(a == -2 || b == -2 || o == -2) { continue; }  // comparator.h:85: all three pins must be connected
            if(o >= 0 && !c->num_host.node_index.empty() && c->num_host.node_index[static_cast<std::size_t>(o)] >= 0)
            {
                set_last_error("circuit_digital_clk: a comparator output on an analog node (digital -> analog driver) is outside the B200 hot path");
                return 1;
            }
            if(o >= 0) { c->digital_state[static_cast<std::size_t>(o)] = volt(a) >= volt(b) ? 1 : 0; }
        }
        return 0;
    }

    namespace
    {
        // digital read-out of circuit_sample*: a pin on a pure digital node reports the node's state (dll_main.cpp:2346-2355)
        void sample_digital(circuit& c, std::size_t* vec_pos, std::size_t* chunk_pos, std::size_t comp_size, std::size_t const* dg, int mode, void* out)
        {
            for(std::size_t i{}; i < comp_size; ++i)
            {
                int const ei{elem_of(c, vec_pos[i], chunk_pos[i])};
                if(ei < 0) { continue; }
                auto const& e{c.nl.elems[static_cast<std::size_t>(ei)]};
                for(int j{}; j < e.d->pins; ++j)
                {
                    int const node{e.pin_node[j]};
                    bool const analog{node == -1 || (node >= 0 && !c.num_host.node_index.empty() && c.num_host.node_index[static_cast<std::size_t>(node)] >= 0)};
                    int st{2};
                    if(node >= 0 && !analog && static_cast<std::size_t>(node) < c.digital_state.size()) { st = c.digital_state[static_cast<std::size_t>(node)]; }
                    std::size_t const k{static_cast<std::size_t>(j) + dg[i]};
                    if(mode == 0) { static_cast<bool*>(out)[k] = !analog && st == 1; }
                    else if(mode == 1) { static_cast<std::uint8_t*>(out)[k] = (!analog && st == 1) ? 1 : 0; }
                    else
                    {
                        static_cast<std::uint8_t*>(out)[k] = static_cast<std::uint8_t>(analog ? 2 : st);  // analog pins report X (dll_api.h:224-226)
                    }
                }
            }
        }
    }  // namespace

    int circuit_sample_layout(void* p, size_t* vec_pos, size_t* chunk_pos, size_t comp_size, size_t* vo, size_t* co, size_t* dg)
    {
        if(p == nullptr || vec_pos == nullptr || chunk_pos == nullptr || vo == nullptr || co == nullptr || dg == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        vo[0] = co[0] = dg[0] = 0;
        for(std::size_t i{}; i < comp_size; ++i)
        {
            int const ei{elem_of(*c, vec_pos[i], chunk_pos[i])};
            std::size_t const np{ei < 0 ? 0u : static_cast<std::size_t>(c->nl.elems[static_cast<std::size_t>(ei)].d->pins)};
            std::size_t const nb{ei < 0 ? 0u : static_cast<std::size_t>(c->nl.elems[static_cast<std::size_t>(ei)].d->branches)};
            vo[i + 1] = vo[i] + np;
            co[i + 1] = co[i] + nb;
            dg[i + 1] = dg[i] + np;
        }
        return 0;
    }

    int circuit_sample(void* p, size_t* vec_pos, size_t* chunk_pos, size_t comp_size, double* voltage, size_t* vo, double* current, size_t* co, bool* digital, size_t* dg)
    {
        if(p == nullptr || vec_pos == nullptr || chunk_pos == nullptr || voltage == nullptr || vo == nullptr || current == nullptr || co == nullptr ||
           digital == nullptr || dg == nullptr)
        {
            return 1;
        }
        if(circuit_sample_layout(p, vec_pos, chunk_pos, comp_size, vo, co, dg) != 0) { return 1; }
        for(std::size_t i{}; i < dg[comp_size]; ++i) { digital[i] = false; }
        sample_digital(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, dg, 0, digital);
        return sample_impl(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, voltage, vo, current, co);
    }

    int circuit_sample_u8(void* p, size_t* vec_pos, size_t* chunk_pos, size_t comp_size, double* voltage, size_t* vo, double* current, size_t* co, uint8_t* digital, size_t* dg)
    {
        if(p == nullptr || vec_pos == nullptr || chunk_pos == nullptr || voltage == nullptr || vo == nullptr || current == nullptr || co == nullptr ||
           digital == nullptr || dg == nullptr)
        {
            return 1;
        }
        if(circuit_sample_layout(p, vec_pos, chunk_pos, comp_size, vo, co, dg) != 0) { return 1; }
        for(std::size_t i{}; i < dg[comp_size]; ++i) { digital[i] = 0; }
        sample_digital(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, dg, 1, digital);
        return sample_impl(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, voltage, vo, current, co);
    }

    int circuit_sample_digital_state_u8(void* p,
                                        size_t* vec_pos,
                                        size_t* chunk_pos,
                                        size_t comp_size,
                                        double* voltage,
                                        size_t* vo,
                                        double* current,
                                        size_t* co,
                                        uint8_t* digital,
                                        size_t* dg)
    {
        if(p == nullptr || vec_pos == nullptr || chunk_pos == nullptr || voltage == nullptr || vo == nullptr || current == nullptr || co == nullptr ||
           digital == nullptr || dg == nullptr)
        {
            return 1;
        }
        if(circuit_sample_layout(p, vec_pos, chunk_pos, comp_size, vo, co, dg) != 0) { return 1; }
        for(std::size_t i{}; i < dg[comp_size]; ++i) { digital[i] = 2; }  // analog pins report X (dll_api.h:224-226)
        sample_digital(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, dg, 2, digital);
        return sample_impl(*static_cast<circuit*>(p), vec_pos, chunk_pos, comp_size, voltage, vo, current, co);
    }

    int circuit_set_model_digital(void* p, size_t vec_pos, size_t chunk_pos, size_t, uint8_t)
    {
        if(p == nullptr) { return 1; }
        return elem_of(*static_cast<circuit*>(p), vec_pos, chunk_pos) < 0 ? 2 : 3;
    }

    int analyze_circuit(void* p,
                        size_t* vec_pos,
                        size_t* chunk_pos,
                        size_t comp_size,
                        int* changed_ele,
                        size_t* changed_ind,
                        double* changed_prop,
                        size_t prop_size,
                        double* voltage,
                        size_t* vo,
                        double* current,
                        size_t* co,
                        bool* digital,
                        size_t* dg)
    {
        if(p && vec_pos && chunk_pos && voltage && vo && current && co && digital && dg)
        {
            auto* c{static_cast<circuit*>(p)};
            for(std::size_t i{}; i < prop_size; ++i)
            {
                auto const ce{static_cast<std::size_t>(changed_ele[i])};
                int const ei{elem_of(*c, vec_pos[ce], chunk_pos[ce])};
                if(ei >= 0 && c->nl.set_attribute(ei, static_cast<int>(changed_ind[i]), changed_prop[i])) { ++c->param_rev; }
            }
            if(!c->analyze()) { return 1; }
            return circuit_sample(p, vec_pos, chunk_pos, comp_size, voltage, vo, current, co, digital, dg);
        }
        return 0;  // dll_main.cpp:2933
    }

    // ---- additive surface ----------------------------------------------------------------------------------------
    int circuit_set_env(void* p, double const* e)
    {
        if(p == nullptr || e == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->env.V_eps_max = e[0];
        c->env.V_epsr_max = e[1];
        c->env.I_eps_max = e[2];
        c->env.I_epsr_max = e[3];
        c->env.g_min = e[4];
        c->env.r_open = e[5];
        c->env.temperature = e[6];
        c->env.norm_temperature = e[7];
        ++c->param_rev;
        return 0;
    }

    int circuit_set_ac_sweep(void* p, int sweep, double w0, double w1, size_t points)
    {
        if(p == nullptr || sweep < 0 || sweep > 2) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        c->ac.sweep = static_cast<sweep_type>(sweep);
        c->ac.omega_start = w0;
        c->ac.omega_stop = w1;
        c->ac.points = points;
        return 0;
    }

    int circuit_unknown_count(void* p, size_t* n_nodes, size_t* n_branches)
    {
        if(p == nullptr) { return 1; }
        auto const nb{make_numbering(static_cast<circuit*>(p)->nl)};
        if(n_nodes) { *n_nodes = static_cast<size_t>(nb.n_nodes); }
        if(n_branches) { *n_branches = static_cast<size_t>(nb.n_branches); }
        return 0;
    }

    long long circuit_pin_unknown(void* p, size_t vec_pos, size_t chunk_pos, size_t pin)
    {
        if(p == nullptr) { return -2; }
        auto* c{static_cast<circuit*>(p)};
        int const ei{elem_of(*c, vec_pos, chunk_pos)};
        if(ei < 0) { return -2; }
        auto const& e{c->nl.elems[static_cast<std::size_t>(ei)]};
        if(pin >= static_cast<size_t>(e.d->pins)) { return -2; }
        int const node{e.pin_node[pin]};
        if(node < 0) { return node; }
        return make_numbering(c->nl).node_index[static_cast<std::size_t>(node)];
    }

    long long circuit_branch_unknown(void* p, size_t vec_pos, size_t chunk_pos, size_t br)
    {
        if(p == nullptr) { return -2; }
        auto* c{static_cast<circuit*>(p)};
        int const ei{elem_of(*c, vec_pos, chunk_pos)};
        if(ei < 0 || br >= static_cast<size_t>(c->nl.elems[static_cast<std::size_t>(ei)].d->branches)) { return -2; }
        auto const nb{make_numbering(c->nl)};
        return nb.n_nodes + nb.branch0[static_cast<std::size_t>(ei)] + static_cast<long long>(br);
    }

    int circuit_get_solution(void* p, double* re, double* im)
    {
        if(p == nullptr || re == nullptr) { return 1; }
        auto* c{static_cast<circuit*>(p)};
        for(std::size_t i{}; i < c->x_host.size(); ++i)
        {
            re[i] = c->x_host[i];
            if(im) { im[i] = c->xi_host[i]; }
        }
        return 0;
    }

    void* circuit_batch_create(void* p, size_t n)
    {
        if(p == nullptr || n == 0)
        {
            set_last_error("circuit_batch_create: null circuit or zero instances");
            return nullptr;
        }
        auto* b{new(std::nothrow) batch{}};
        if(b == nullptr) { return nullptr; }
        b->parent = static_cast<circuit*>(p);
        b->n_inst = n;
        auto const& d{default_path()};
        b->res_S = d.res_S;
        b->res_I = d.res_I;
        b->res_J = d.res_J;
        b->subtree_warps = d.subtree_warps;
        b->res_ws = d.res_ws;
        b->res_chunks = d.res_chunks;
        circuit_batch_set_tuning(b, d.tuning);
        return b;
    }

    void circuit_batch_destroy(void* b) { delete static_cast<batch*>(b); }

    int circuit_batch_set_device(void* b, int device)
    {
        if(b == nullptr) { return 1; }
        static_cast<batch*>(b)->device = device;
        return 0;
    }

    int circuit_batch_set_subtree_warps(void* b, int warps)
    {
        if(b == nullptr || warps < 0 || warps > PE_MAX_WARPS || (warps & (warps - 1)) != 0) { return 1; }
        static_cast<batch*>(b)->subtree_warps = warps;
        return 0;
    }

    int circuit_batch_set_resident(void* b, int streams, int instances_per_cta, int instances_per_thread)
    {
        auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
        if(b == nullptr || streams < -1 || streams > 1024 || (streams > 0 && !pow2(streams))) { return 1; }
        if(instances_per_cta < 0 || instances_per_cta > 32 || (instances_per_cta > 0 && !pow2(instances_per_cta))) { return 1; }
        if(instances_per_thread < 0 || (instances_per_thread > 2 && instances_per_thread != 4)) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        bp->res_S = streams;
        bp->res_I = instances_per_cta;
        bp->res_J = instances_per_thread;
        return 0;
    }

    int circuit_batch_set_workspace(void* b, int where)
    {
        if(b == nullptr || where < 0 || where > 2) { return 1; }
        static_cast<batch*>(b)->res_ws = where;
        return 0;
    }

    int circuit_batch_set_chunks(void* b, int chunks)
    {
        if(b == nullptr || chunks < 0 || chunks > 32) { return 1; }
        static_cast<batch*>(b)->res_chunks = chunks;
        return 0;
    }

    int circuit_batch_set_tuning(void* b, unsigned flags)
    {
        if(b == nullptr || flags > 255u || (flags & 48u) == 48u || (flags & 192u) == 192u) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        bp->res_prefetch = ((flags & 1u) ? 1 : 0) | ((flags & 2u) ? 2 : 0) | ((flags & 4u) ? 0 : 4);
        bp->res_fuse = (flags & 8u) ? 1 : 0;
        bp->res_jit = (flags & 16u) ? 1 : ((flags & 32u) ? -1 : 0);
        bp->res_stream = (flags & 64u) ? 1 : ((flags & 128u) ? -1 : 0);  // stream kernel: required / forbidden
        return 0;
    }

    int circuit_batch_last_kernel(void* b) { return b == nullptr ? -1 : static_cast<batch*>(b)->last_jit; }

    // shard of an AC sweep: solve only the points [first, first + count) of the sweep set by circuit_batch_set_ac_sweep
    // (count = 0: all of them).  The omega table is still built in full, sequentially, so every rank sees the reference's values.
    int circuit_batch_set_ac_slice(void* b, size_t first, size_t count)
    {
        if(b == nullptr) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        bp->ac_slice_first = first;
        bp->ac_slice_count = count;
        return 0;
    }

    // circuits of at least this many unknowns (default 20 000) that consist of resistors and DC sources take the
    // reduce-and-core path for DC / OP; process-wide
    int phy_engine_b200_set_frontal_min(size_t n_unknowns)
    {
        frontal_min_unknowns() = n_unknowns;
        return 0;
    }

    // reduce-and-core path (config A): info[11] = unknowns, eliminated nodes, levels, core rows, edges (fill included), kernel
    // launches of the last solve, leading dimension of the dense core, edges inside the core, device microseconds of the
    // reduction / the core LU / the substitutions of the last solve; 1 = the batch did not take it
    int circuit_batch_frontal_info(void* b, int64_t* info)
    {
        if(b == nullptr || info == nullptr) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        if(!bp->frontal) { return 1; }
        frontal_stats(bp->frontal.get(), info);
        return 0;
    }

    // stream kernel of the last launch: [0] last kernel (2 = stream), [1] warps per CTA, [2] ring stages, [3] shared memory
    // per CTA, [4] tiles per solve, [5] rows per ring stage
    int circuit_batch_stream_info(void* b, int mode, int64_t* info)
    {
        if(b == nullptr || info == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        int g[3]{};
        pe_b200_stream_last_geometry(g);
        info[0] = bp->last_jit;
        info[1] = g[0];
        info[2] = g[1];
        info[3] = g[2];
        info[4] = info[5] = 0;
        if(bp->cc)
        {
            auto const& pr{bp->cc->prog[static_cast<std::size_t>(mode)]};
            info[4] = pr.stream_tiles;
            info[5] = pr.stream_stage_rows;
        }
        return 0;
    }

    // batch-state checkpoint (pe_host.hpp): call with buffer == NULL to learn the size; 0 = ok, 1 = error, 2 = buffer too small
    int circuit_batch_save_state(void* b, void* buffer, size_t capacity, size_t* size)
    {
        if(b == nullptr || size == nullptr) { return 1; }
        std::vector<unsigned char> blob;
        if(!static_cast<batch*>(b)->save_state(blob)) { return 1; }
        *size = blob.size();
        if(buffer == nullptr) { return 0; }
        if(capacity < blob.size()) { return 2; }
        std::memcpy(buffer, blob.data(), blob.size());
        return 0;
    }

    int circuit_batch_load_state(void* b, void const* buffer, size_t size)
    {
        if(b == nullptr || buffer == nullptr) { return 1; }
        return static_cast<batch*>(b)->load_state(static_cast<unsigned char const*>(buffer), size) ? 0 : 1;
    }

    // pivot safety net (pe_host.hpp): guard = 1 / the largest multiplier |l| a guarded column may hold (0 = off, < 0 = default
    // 2^-20), rounds = re-orderings tried before the last, unguarded one (< 0 = default 3)
    int circuit_batch_set_pivot_guard(void* b, double guard, int rounds)
    {
        if(b == nullptr || guard >= 1.0 || rounds > 16) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        bp->pivot_guard = guard < 0.0 ? PE_GUARD_DEFAULT : guard;
        bp->rescue_rounds = rounds < 0 ? 3 : std::max(rounds, 1);
        bp->rescues.clear();  // sub-batches built under the old setting no longer own anything
        return 0;
    }

    // info[6] = guarded pivots of the program of `mode`, instances the guard flagged so far, instances a re-ordered sub-batch
    // solved, instances that went through the last (unguarded) round, sub-batch analyze() calls, live sub-batches
    int circuit_batch_rescue_info(void* b, int mode, int64_t* info)
    {
        if(b == nullptr || info == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        info[0] = bp->cc ? static_cast<std::int64_t>(bp->cc->prog[static_cast<std::size_t>(mode)].n_guarded) : -1;
        info[1] = static_cast<std::int64_t>(bp->stat_guard_trips);
        info[2] = static_cast<std::int64_t>(bp->stat_rescued);
        info[3] = static_cast<std::int64_t>(bp->stat_unguarded);
        info[4] = static_cast<std::int64_t>(bp->stat_rescue_launches);
        info[5] = static_cast<std::int64_t>(bp->rescues.size());
        return 0;
    }

    int circuit_batch_resident_info(void* b, int mode, int64_t* info)
    {
        if(b == nullptr || info == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* bp{static_cast<batch*>(b)};
        if(!bp->cc) { return 1; }
        auto const& pr{bp->cc->prog[static_cast<std::size_t>(mode)]};
        int I{}, J{};
        bool const fits{pr.resident && bp->pick_geometry(pr, I, J)};
        info[0] = pr.resident ? 1 : 0;
        info[1] = pr.rS;
        info[2] = pr.r_slots;
        info[3] = fits ? I : 0;
        info[4] = fits ? J : 0;
        info[5] = static_cast<int64_t>(pr.io.size());
        info[6] = bp->last_S;
        info[7] = bp->last_I;
        info[8] = bp->last_J;
        std::size_t nph{};
        if(pr.resident && !pr.rstreams.empty()) { nph = pr.rstreams[0].sec[2].size(); }
        info[9] = static_cast<int64_t>(nph);
        info[10] = static_cast<int64_t>(pr.words.size());
        info[11] = static_cast<int64_t>(pr.max_warp_words);
        info[12] = (pr.resident && bp->use_hbm(pr)) ? 1 : 0;
        return 0;
    }

    size_t circuit_batch_resident_secoff(void* b, int mode, uint32_t* out)
    {
        if(b == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 0; }
        auto* bp{static_cast<batch*>(b)};
        if(!bp->cc) { return 0; }
        auto const& so{bp->cc->prog[static_cast<std::size_t>(mode)].sec_off};
        if(out != nullptr) { std::memcpy(out, so.data(), so.size() * sizeof(std::uint32_t)); }
        return so.size();
    }

    int circuit_batch_set_stream(void* b, void* s)
    {
        if(b == nullptr) { return 1; }
        static_cast<batch*>(b)->stream = s;
        return 0;
    }

    int circuit_batch_set_param(void* bp, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double const* values)
    {
        if(bp == nullptr || name == nullptr || values == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        int const ei{elem_of(*b->parent, vec_pos, chunk_pos)};
        if(ei < 0) { return 2; }
        int const idx{b->parent->nl.find_attribute(ei, name, name_size)};
        if(idx < 0) { return 3; }
        int const code{b->parent->nl.elems[static_cast<std::size_t>(ei)].d->code};
        if(code == E_SWITCH)
        {
            set_last_error("circuit_batch_set_param: switch state cannot be swept per instance");
            return 1;
        }
        auto& v{b->sweeps[{ei, idx}]};
        v.resize(b->n_inst);
        for(std::size_t i{}; i < b->n_inst; ++i) { v[i] = to_internal(code, idx, values[i]); }
        if((code == E_RES || code == E_CAP) && idx == 0)
        {
            // the compiler leaves the pivot guard off rows that only carry positive conductances (compiler.cpp "guard elision")
            for(double const x: v)
            {
                if(!(x > 0.0)) { b->guard_all = true; }
            }
        }
        b->sweeps_dirty = true;
        ++b->sweeps_rev;
        return 0;
    }

    int circuit_batch_set_params(void* bp, size_t n_params, size_t const* vec_pos, size_t const* chunk_pos, char const* const* names, double const* values)
    {
        if(bp == nullptr || (n_params != 0 && (vec_pos == nullptr || chunk_pos == nullptr || names == nullptr || values == nullptr))) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        bool direct{b->cc && !b->device_stale && !b->layout_pending && b->d_wi.p != nullptr};
        for(size_t k{}; k < n_params && direct; ++k)
        {
            // fast path only when every parameter already has a device row and needs no unit conversion
            int const ei{elem_of(*b->parent, vec_pos[k], chunk_pos[k])};
            if(ei < 0 || names[k] == nullptr) { return 2; }
            int const idx{b->parent->nl.find_attribute(ei, names[k], std::strlen(names[k]))};
            if(idx < 0) { return 3; }
            int const code{b->parent->nl.elems[static_cast<std::size_t>(ei)].d->code};
            if(code == E_VAC || code == E_IAC || code == E_SWITCH || (code == E_PN && idx == 7)) { direct = false; }
            else if(b->cc->swept_slot.find({ei, idx}) == b->cc->swept_slot.end()) { direct = false; }
        }
        if(!direct)
        {
            for(size_t k{}; k < n_params; ++k)
            {
                int const rc{circuit_batch_set_param(bp, vec_pos[k], chunk_pos[k], names[k], names[k] ? std::strlen(names[k]) : 0, values + k * b->n_inst)};
                if(rc != 0) { return rc; }
            }
            return 0;
        }
        if(pe_b200_dev_set(b->device) != 0) { return 1; }
        ++b->sweeps_rev;
        // parameters whose device rows are equally spaced travel as one strided 2-D copy (one DMA descriptor instead of one
        // per 80 KB row: a table of 2000 parameters is a handful of copies)
        std::vector<std::int64_t> slot(n_params);
        for(size_t k{}; k < n_params; ++k)
        {
            int const ei{elem_of(*b->parent, vec_pos[k], chunk_pos[k])};
            int const idx{b->parent->nl.find_attribute(ei, names[k], std::strlen(names[k]))};
            auto& v{b->sweeps[{ei, idx}]};
            v.assign(1, values[k * b->n_inst]);  // nominal (lane-0) value for the symbolic phase; the full row lives on the device only
            slot[k] = static_cast<std::int64_t>(b->cc->swept_slot.at({ei, idx}));
        }
        for(size_t k{}; k < n_params;)
        {
            // a run of parameters whose device rows are equally spaced (every R of a ladder: rows 0, 2, 4, ...) = one copy
            // whose destination pitch is that spacing
            std::int64_t const d{k + 1 < n_params ? slot[k + 1] - slot[k] : 1};
            size_t e{k + 1};
            if(d > 0)
            {
                while(e < n_params && slot[e] - slot[e - 1] == d) { ++e; }
            }
            auto* dst{static_cast<double*>(b->d_wi.p) + slot[k] * b->LSi};
            std::size_t const dpitch{static_cast<std::size_t>(d > 0 ? d : 1) * static_cast<std::size_t>(b->LSi) * sizeof(double)};
            if(pe_b200_dev_h2d_2d(dst, dpitch, values + k * b->n_inst, b->n_inst * sizeof(double), b->n_inst * sizeof(double), e - k, b->stream) != 0)
            {
                set_last_error(std::string{"circuit_batch_set_params: "} + pe_b200_dev_last_error());
                return 1;
            }
            k = e;
        }
        // the caller's buffer may be reused as soon as we return
        return pe_b200_dev_sync(b->stream) == 0 ? 0 : 1;
    }

    int circuit_batch_set_ac_sweep(void* bp, int sweep, double w0, double w1, size_t points)
    {
        if(bp == nullptr || sweep < 0 || sweep > 2) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        b->ac.sweep = static_cast<sweep_type>(sweep);
        b->ac.omega_start = w0;
        b->ac.omega_stop = w1;
        b->ac.omega = w0;
        b->ac.points = points;
        return 0;
    }

    int circuit_batch_set_probes(void* bp, size_t const* u, size_t n)
    {
        if(bp == nullptr || (n != 0 && u == nullptr)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        auto const nb{make_numbering(b->parent->nl)};
        b->probes.clear();
        for(size_t i{}; i < n; ++i)
        {
            if(u[i] >= static_cast<size_t>(nb.unknowns())) { return 1; }
            b->probes.push_back(static_cast<int>(u[i]));
        }
        return 0;
    }

    int circuit_batch_prepare(void* bp)
    {
        if(bp == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->ensure_compiled() || !b->upload_sweeps()) { return 1; }
        return pe_b200_dev_sync(b->stream) == 0 ? 0 : 1;
    }

    int circuit_batch_reset_state(void* bp)
    {
        if(bp == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->ensure_compiled()) { return 1; }
        b->rescues.clear();  // the sub-batches of the pivot safety net hold state of the transient that is being reset
        // INST layout: x [0,n) | swept parameters [n, n+ns) | device state + derived values
        std::size_t const n{static_cast<std::size_t>(b->cc->num.unknowns())};
        std::size_t const ns{b->cc->swept_slot.size()};
        std::size_t const row{static_cast<std::size_t>(b->LSi) * sizeof(double)};
        auto* base{static_cast<char*>(b->d_wi.p)};
        if(pe_b200_dev_memset0(base, n * row, b->stream) != 0) { return 1; }
        std::size_t const tail{static_cast<std::size_t>(b->cc->n_inst_slots) - n - ns};
        if(tail != 0 && pe_b200_dev_memset0(base + (n + ns) * row, tail * row, b->stream) != 0) { return 1; }
        b->tr_duration = 0.0;
        b->last_step = 0.0;
        return 0;
    }

    int circuit_batch_analyze(void* bp)
    {
        if(bp == nullptr) { return 1; }
        return static_cast<batch*>(bp)->analyze() ? 0 : 1;
    }

    size_t circuit_batch_lanes(void* bp) { return bp ? static_cast<batch*>(bp)->last_lanes : 0; }

    size_t circuit_batch_points(void* bp) { return bp ? static_cast<batch*>(bp)->last_points : 0; }

    uint64_t circuit_batch_total_solves(void* bp) { return bp ? static_cast<batch*>(bp)->total_solves : 0; }

    double circuit_batch_tr_duration(void* bp) { return bp ? static_cast<batch*>(bp)->tr_duration : 0.0; }

    int circuit_batch_solution(void* bp, double* x) { return (bp && x && static_cast<batch*>(bp)->get_solution(x)) ? 0 : 1; }

    int circuit_batch_solution_soa(void* bp, double* x) { return (bp && x && static_cast<batch*>(bp)->get_solution_soa(x)) ? 0 : 1; }

    int circuit_batch_ac_solution(void* bp, double* x) { return (bp && x && static_cast<batch*>(bp)->get_ac_solution(x)) ? 0 : 1; }

    int circuit_batch_ac_solution_lanes(void* bp, size_t const* lanes, size_t n, double* x)
    {
        return (bp && (n == 0 || (lanes && x)) && static_cast<batch*>(bp)->get_ac_solution_lanes(lanes, n, x)) ? 0 : 1;
    }

    int circuit_batch_ac_omegas(void* bp, double* om)
    {
        if(bp == nullptr || om == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        for(std::size_t i{}; i < b->ac_omegas.size(); ++i) { om[i] = b->ac_omegas[i]; }
        return 0;
    }

    int circuit_batch_status(void* bp, int32_t* st) { return (bp && st && static_cast<batch*>(bp)->get_status(st)) ? 0 : 1; }

    int circuit_batch_newton_iters(void* bp, uint32_t* n) { return (bp && n && static_cast<batch*>(bp)->get_solves(n)) ? 0 : 1; }

    int circuit_batch_digital_clk(void* bp) { return (bp && static_cast<batch*>(bp)->digital_clk()) ? 0 : 1; }

    size_t circuit_batch_comparator_count(void* bp) { return bp ? static_cast<batch*>(bp)->n_cmp : 0; }

    int circuit_batch_comparator_states(void* bp, uint8_t* out) { return (bp && out && static_cast<batch*>(bp)->get_comparator_states(out)) ? 0 : 1; }

    int circuit_batch_waveform(void* bp, double* w) { return (bp && w && static_cast<batch*>(bp)->get_wave(w)) ? 0 : 1; }

    int circuit_batch_stats(void* bp, int mode, size_t* n_unknowns, size_t* nnz_a, size_t* nnz_lu, size_t* n_fma, size_t* n_lane_slots, size_t* n_inst_slots)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        auto const& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        if(n_unknowns) { *n_unknowns = static_cast<size_t>(b->cc->num.unknowns()); }
        if(nnz_a) { *nnz_a = pr.nnz_a; }
        if(nnz_lu) { *nnz_lu = pr.nnz_lu; }
        if(n_fma) { *n_fma = pr.n_fma; }
        if(n_lane_slots) { *n_lane_slots = static_cast<size_t>(pr.n_lane_slots); }
        if(n_inst_slots) { *n_inst_slots = static_cast<size_t>(b->cc->n_inst_slots); }
        return 0;
    }

    int circuit_batch_param_device_ptr(void* bp, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double** dptr)
    {
        if(bp == nullptr || name == nullptr || dptr == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        int const ei{elem_of(*b->parent, vec_pos, chunk_pos)};
        if(ei < 0) { return 2; }
        int const idx{b->parent->nl.find_attribute(ei, name, name_size)};
        if(idx < 0) { return 3; }
        auto it{b->cc->swept_slot.find({ei, idx})};
        if(it == b->cc->swept_slot.end()) { return 3; }
        *dptr = static_cast<double*>(b->d_wi.p) + static_cast<std::int64_t>(it->second) * b->LSi;
        return 0;
    }

    int circuit_batch_solution_device_ptr(void* bp, double** x0, size_t* lane_stride)
    {
        if(bp == nullptr || x0 == nullptr) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        *x0 = static_cast<double*>(b->d_wi.p);
        if(lane_stride) { *lane_stride = static_cast<size_t>(b->LSi); }
        return 0;
    }

    // ---- introspection (symbolic phase only; no device needed) ----------------------------------------------------
    int circuit_batch_compile_host(void* bp)
    {
        if(bp == nullptr) { return 1; }
        bool lc{};
        auto* b{static_cast<batch*>(bp)};
        if(!b->compile_host(lc)) { return 1; }
        if(lc) { b->layout_pending = true; }
        return 0;
    }

    // tooling for the stream kernel of a compiled batch (circuit_batch_compile_host first; no device needed): the generated
    // source (returns its length; copies at most cap bytes; 0 = this mode's program does not take the stream kernel) and
    // build-or-fetch of its module, so that __graft_entry__.build() ships the cubin of the bench configuration to the GPU box.
    // stats[8] (optional) = tiles per solve, rows per ring stage, loops, ops in loops, ops, rows fetched per solve, bulk copies
    // per solve, rows stored per solve.
    size_t circuit_batch_stream_source(void* bp, int mode, char* out, size_t cap, uint64_t* stats)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 0; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc || !b->stream_mode) { return 0; }
        auto const& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        if(!pr.stream_laid_out) { return 0; }
        stream_geom g{};
        std::string const src{stream_generate(pr, g)};
        if(out != nullptr && cap > 0) { std::memcpy(out, src.data(), std::min(cap, src.size())); }
        if(stats != nullptr)
        {
            std::uint64_t const v[8]{g.n_tiles, g.stage_rows, g.n_loops, g.loop_ops, g.n_ops, g.rows_fetched, g.n_copies, g.rows_stored};
            std::memcpy(stats, v, sizeof(v));
        }
        return src.size();
    }

    int circuit_batch_stream_build(void* bp, int mode)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc || !b->stream_mode)
        {
            set_last_error("stream: this batch does not take the stream kernel (small batch, small or nonlinear circuit, or ops the generator does not cover)");
            return 1;
        }
        auto& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        int I{}, J{};
        if(!pr.stream_laid_out || !b->pick_geometry(pr, I, J))
        {
            set_last_error("stream: the program of this mode holds ops the stream kernel does not cover");
            return 1;
        }
        stream_geom g{};
        std::string const src{stream_generate(pr, g)};
        if(src.empty() || !stream_compile(src, J, I / J, pr.stream_blob, pr.stream_key, pr.stream_error))
        {
            set_last_error(src.empty() ? std::string{"stream: generator failed"} : pr.stream_error);
            return 1;
        }
        pr.stream_state = 1;
        pr.stream_j = I;
        pr.stream_tiles = g.n_tiles;
        pr.stream_stage_rows = g.stage_rows;
        return 0;
    }

    // tooling: the generated source of the specialised kernel's iter section (returns its length; copies at most cap bytes)
    size_t circuit_batch_jit_source(void* bp, int mode, char* out, size_t cap)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 0; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 0; }
        auto const& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        if(!jit_supported(pr)) { return 0; }
        std::string const src{jit_generate(pr, jit_load_distance())};
        if(out != nullptr && cap > 0) { std::memcpy(out, src.data(), std::min(cap, src.size())); }
        return src.size();
    }

    // build (or find in the cache) the specialised kernel of a compiled program without a device: nvcc cross-compiles, so
    // __graft_entry__.build() can ship the cubin of the bench configuration to the GPU box.  cluster = CTAs per lane group.
    int circuit_batch_jit_build(void* bp, int mode, int cluster)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT) || (cluster != 1 && cluster != 2)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        auto& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        if(!jit_supported(pr))
        {
            set_last_error("jit: the iter section holds ops the specialised kernel does not cover");
            return 1;
        }
        if(!jit_compile(jit_generate(pr, jit_load_distance()), cluster, pr.jit_cubin, pr.jit_key, pr.jit_error))
        {
            set_last_error(pr.jit_error);
            return 1;
        }
        pr.jit_state = 1;
        pr.jit_cl = cluster;
        return 0;
    }

    size_t circuit_batch_program_words(void* bp, int mode)
    {
        if(bp == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 0; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 0; }
        return b->cc->prog[static_cast<std::size_t>(mode)].words.size();
    }

    int circuit_batch_program_copy(void* bp, int mode, uint32_t* out)
    {
        if(bp == nullptr || out == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        auto const& v{b->cc->prog[static_cast<std::size_t>(mode)].words};
        std::memcpy(out, v.data(), v.size() * 4);
        return 0;
    }

    size_t circuit_batch_const_count(void* bp)
    {
        auto* b{static_cast<batch*>(bp)};
        return (b && b->cc) ? b->cc->cst.size() : 0;
    }

    int circuit_batch_const_copy(void* bp, double* out)
    {
        auto* b{static_cast<batch*>(bp)};
        if(b == nullptr || !b->cc || out == nullptr) { return 1; }
        std::memcpy(out, b->cc->cst.data(), b->cc->cst.size() * sizeof(double));
        return 0;
    }

    // info[0..15] = cplx, structurally_singular, n_lane_slots, omega_slot, n_inst_slots, dt_slot, x_slot0, n_unknowns, warps,
    // n_real_lane_slots, n_leaves, n_leaf_rows, n_top_rows, max_warp_words, nnz_a, nnz_lu;
    // info[16 + 16 * s + g] = word offset of warp g's stream of section s (0 prep, 1 step, 2 iter), -1 if absent
    int circuit_batch_program_info(void* bp, int mode, int64_t* info)
    {
        if(bp == nullptr || info == nullptr || mode < 0 || mode >= static_cast<int>(prog_mode::COUNT)) { return 1; }
        auto* b{static_cast<batch*>(bp)};
        if(!b->cc) { return 1; }
        auto const& pr{b->cc->prog[static_cast<std::size_t>(mode)]};
        info[0] = pr.cplx;
        info[1] = pr.structurally_singular;
        info[2] = pr.n_lane_slots;
        info[3] = pr.omega_slot;
        info[4] = b->cc->n_inst_slots;
        info[5] = b->cc->dt_slot;
        info[6] = pr.x_opnd.empty() ? 0 : static_cast<int64_t>(PE_OPND_SLOT(pr.x_opnd[0]));
        info[7] = b->cc->num.unknowns();
        info[8] = pr.warps;
        info[9] = b->cc->n_real_lane_slots;
        info[10] = static_cast<int64_t>(pr.n_leaves);
        info[11] = static_cast<int64_t>(pr.n_leaf_rows);
        info[12] = static_cast<int64_t>(pr.n_top_rows);
        info[13] = static_cast<int64_t>(pr.max_warp_words);
        info[14] = static_cast<int64_t>(pr.nnz_a);
        info[15] = static_cast<int64_t>(pr.nnz_lu);
        pe_b200_section const* secs[3]{&pr.prep, &pr.step, &pr.iter};
        for(int s{}; s < 3; ++s)
        {
            for(int g{}; g < PE_MAX_WARPS; ++g)
            {
                auto const o{secs[s]->off[g]};
                info[16 + 16 * s + g] = o == PE_NO_SECTION ? -1 : static_cast<int64_t>(o);
            }
        }
        return 0;
    }

    long long circuit_batch_swept_slot(void* bp, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size)
    {
        auto* b{static_cast<batch*>(bp)};
        if(b == nullptr || !b->cc || name == nullptr) { return -1; }
        int const ei{elem_of(*b->parent, vec_pos, chunk_pos)};
        if(ei < 0) { return -1; }
        int const idx{b->parent->nl.find_attribute(ei, name, name_size)};
        auto it{b->cc->swept_slot.find({ei, idx})};
        return it == b->cc->swept_slot.end() ? -1 : it->second;
    }

    // host copy of one swept parameter row in internal units (what upload_sweeps() sends to the device)
    int circuit_batch_swept_values(void* bp, long long slot, double* out)
    {
        auto* b{static_cast<batch*>(bp)};
        if(b == nullptr || !b->cc || out == nullptr) { return 1; }
        for(auto const& [k, v]: b->sweeps)
        {
            auto it{b->cc->swept_slot.find(k)};
            if(it != b->cc->swept_slot.end() && it->second == slot)
            {
                std::memcpy(out, v.data(), v.size() * sizeof(double));
                return 0;
            }
        }
        return 1;
    }

    int phy_engine_b200_set_default_path(int streams, int instances_per_cta, int instances_per_thread, int subtree_warps, int workspace, unsigned tuning)
    {
        if(workspace < 0 || workspace > 2 || tuning > 255u || (tuning & 48u) == 48u || (tuning & 192u) == 192u) { return 1; }
        auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
        if(streams < -1 || streams > 1024 || (streams > 0 && !pow2(streams))) { return 1; }
        if(instances_per_cta < 0 || instances_per_cta > 32 || (instances_per_cta > 0 && !pow2(instances_per_cta))) { return 1; }
        if(instances_per_thread < 0 || (instances_per_thread > 2 && instances_per_thread != 4)) { return 1; }
        if(subtree_warps < 0 || subtree_warps > PE_MAX_WARPS || (subtree_warps > 0 && !pow2(subtree_warps))) { return 1; }
        auto& d{default_path()};
        d.res_S = streams;
        d.res_I = instances_per_cta;
        d.res_J = instances_per_thread;
        d.subtree_warps = subtree_warps;
        d.res_ws = workspace;
        d.tuning = tuning;
        return 0;
    }

    int phy_engine_b200_device_count(void) { return pe_b200_dev_count(); }

    uint64_t phy_engine_b200_launch_count(void) { return pe_b200_launch_count(); }
    uint64_t phy_engine_b200_aux_launch_count(void) { return pe_b200_aux_launch_count(); }

    void phy_engine_b200_timing(int on) { pe_b200_timing_enable(on); }

    double phy_engine_b200_kernel_ms(void) { return pe_b200_timing_collect(); }
}
