// frontal.cpp — symbolic phase and driver of the REDUCE-AND-CORE path (kernels: csrc/pe_b200_frontal.cu).
//
// Config A of BASELINE.json (benchmark/series_parallel.cpp): ONE linear DC circuit of ~1e5 resistors whose random node
// merges give the LU real fill — the static dot-product programs of the batch kernels would need ~1e10 operand words.
// Here the parallelism comes from the elimination DAG instead (SURVEY.md §7 hard part 8):
//   * every node of degree <= 2 is eliminated symbolically, level by level (a level = pairwise non-adjacent nodes, chosen
//     greedily; a chain between two junctions halves per level); every fill edge gets its number now
//   * the nodes that remain (junctions of degree >= 3, terminals of voltage sources) and the source branches form the dense
//     core, nodes first so that no pivoting is needed (Schur complement of an M-matrix, then the quasi-definite border)
// Scope: DC / OP of netlists made of resistors, DC voltage sources and DC current sources (resistance.h:82-110,
// VDC.h:82-96, IDC.h:84-94); per-instance sweeps of their values are supported (Monte-Carlo batches).  Anything else keeps
// the compiled-program kernels.
#include "pe_host.hpp"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <set>

namespace pe_b200
{
    namespace
    {
        constexpr int k_nb{64};  // block size of the dense LU (csrc/pe_b200_frontal.cu)

        std::int64_t round_up(std::int64_t v, std::int64_t m) { return (v + m - 1) / m * m; }
    }  // namespace

    // circuits below this many unknowns keep the compiled-program kernels (process-wide; phy_engine_b200_set_frontal_min)
    std::size_t& frontal_min_unknowns()
    {
        static std::size_t v{std::getenv("PE_B200_FRONTAL_MIN") != nullptr ? static_cast<std::size_t>(std::atoll(std::getenv("PE_B200_FRONTAL_MIN"))) : 20000u};
        return v;
    }

    bool frontal_applicable(circuit const& c, std::size_t n_unknowns)
    {
        if(c.at != analyze_type::DC && c.at != analyze_type::OP) { return false; }
        if(n_unknowns < frontal_min_unknowns()) { return false; }
        if(c.env.g_min != 0.0) { return false; }
        for(auto const& e: c.nl.elems)
        {
            if(e.d->code != E_RES && e.d->code != E_VDC && e.d->code != E_IDC) { return false; }
            // an element stamps only when all its pins are connected (resistance.h:86); half-connected ones would leave a node
            // without any stamp: not this path's business
            int connected{};
            for(int p{}; p < e.d->pins; ++p) { connected += e.pin_node[p] != -2 ? 1 : 0; }
            if(connected != 0 && connected != e.d->pins) { return false; }
        }
        return true;
    }

    struct frontal_state
    {
        // symbolic
        std::vector<int> res_elem, idc_elem, vdc_elem;
        std::vector<std::int32_t> res_a, res_b, res_edge, idc_p, idc_q, ops, level_off, core_unknown, core_edges, vdc;
        int n_nodes{}, n_unknowns{}, n_edges{}, n_core{}, n_core_nodes{};
        std::int64_t ld_core{};
        std::uint64_t structure_rev{};
        // device
        device_buf d_rval, d_idc, d_vdcv, d_ia, d_ib, d_ie, d_ip, d_iq, d_ops, d_cu, d_ce, d_vdc, d_d, d_z, d_g, d_M, d_c;
        std::size_t n_inst_built{};
        numbering num{};
        std::uint64_t num_rev{};
        std::uint64_t val_sweeps_rev{}, val_param_rev{};
        std::size_t val_n_inst{};
        void const* val_batch{};
        std::uint64_t launches{};
        double phase_ms[3]{};  // device time of the last solve: reduce / core LU / substitutions
    };

    frontal_state* frontal_new() { return new frontal_state; }
    void frontal_delete(frontal_state* s) { delete s; }
    void frontal_stats(frontal_state const* s, std::int64_t* out)
    {
        out[0] = s->n_unknowns;
        out[1] = static_cast<std::int64_t>(s->ops.size() / 6);
        out[2] = static_cast<std::int64_t>(s->level_off.size()) - 1;
        out[3] = s->n_core;
        out[4] = s->n_edges;
        out[5] = static_cast<std::int64_t>(s->launches);
        out[6] = s->ld_core;
        out[7] = static_cast<std::int64_t>(s->core_edges.size() / 3);
        for(int i{}; i < 3; ++i) { out[8 + i] = static_cast<std::int64_t>(s->phase_ms[i] * 1000.0); }  // microseconds
    }

    // symbolic phase: pure integer / graph work on the host
    static void frontal_build(frontal_state& s, netlist const& nl, numbering const& num)
    {
        // symbolic tables start empty (the device buffers are kept and re-used)
        for(auto* v: {&s.res_a, &s.res_b, &s.res_edge, &s.idc_p, &s.idc_q, &s.ops, &s.level_off, &s.core_unknown, &s.core_edges, &s.vdc}) { v->clear(); }
        s.res_elem.clear();
        s.idc_elem.clear();
        s.vdc_elem.clear();
        s.n_edges = 0;
        s.n_nodes = num.n_nodes;
        s.n_unknowns = num.unknowns();
        auto unk = [&](int node) { return node >= 0 ? num.node_index[static_cast<std::size_t>(node)] : -1; };
        std::map<std::pair<int, int>, int> edge_id;
        auto edge_of = [&](int a, int b) -> int
        {
            if(a > b) { std::swap(a, b); }
            auto const it{edge_id.find({a, b})};
            if(it != edge_id.end()) { return it->second; }
            int const id{s.n_edges++};
            edge_id.emplace(std::make_pair(a, b), id);
            return id;
        };
        std::vector<std::set<int>> adj(static_cast<std::size_t>(s.n_nodes));
        std::vector<char> core(static_cast<std::size_t>(s.n_nodes), 0);
        for(std::size_t ei{}; ei < nl.elems.size(); ++ei)
        {
            auto const& e{nl.elems[ei]};
            if(e.pin_node[0] == -2) { continue; }  // unconnected element: no stamp
            int const a{unk(e.pin_node[0])}, b{unk(e.pin_node[1])};
            if(e.d->code == E_RES)
            {
                if(a < 0 && b < 0) { continue; }
                s.res_elem.push_back(static_cast<int>(ei));
                if(a == b)
                {
                    // both pins on one node: the four stamps cancel (up to rounding); nothing to solve for
                    s.res_a.push_back(-1);
                    s.res_b.push_back(-1);
                    s.res_edge.push_back(-1);
                    continue;
                }
                s.res_a.push_back(a);
                s.res_b.push_back(b);
                if(a >= 0 && b >= 0)
                {
                    s.res_edge.push_back(edge_of(a, b));
                    adj[static_cast<std::size_t>(a)].insert(b);
                    adj[static_cast<std::size_t>(b)].insert(a);
                }
                else
                {
                    s.res_edge.push_back(-1);
                }
            }
            else if(e.d->code == E_IDC)
            {
                s.idc_elem.push_back(static_cast<int>(ei));
                s.idc_p.push_back(a);
                s.idc_q.push_back(b);
            }
            else
            {
                s.vdc_elem.push_back(static_cast<int>(ei));
                if(a >= 0) { core[static_cast<std::size_t>(a)] = 1; }
                if(b >= 0) { core[static_cast<std::size_t>(b)] = 1; }
            }
        }
        // level-scheduled elimination of the nodes of degree <= 2
        std::vector<char> gone(static_cast<std::size_t>(s.n_nodes), 0);
        std::vector<int> cand;
        for(int k{}; k < s.n_nodes; ++k)
        {
            if(!core[static_cast<std::size_t>(k)] && adj[static_cast<std::size_t>(k)].size() <= 2) { cand.push_back(k); }
        }
        s.level_off.push_back(0);
        std::vector<char> blocked(static_cast<std::size_t>(s.n_nodes), 0);
        while(!cand.empty())
        {
            // a maximal set of pairwise non-adjacent candidates
            std::vector<int> picked, touched;
            for(int const k: cand)
            {
                if(gone[static_cast<std::size_t>(k)] || blocked[static_cast<std::size_t>(k)] || adj[static_cast<std::size_t>(k)].size() > 2) { continue; }
                picked.push_back(k);
                for(int const a: adj[static_cast<std::size_t>(k)])
                {
                    blocked[static_cast<std::size_t>(a)] = 1;
                    touched.push_back(a);
                }
            }
            if(picked.empty()) { break; }
            std::vector<int> next;
            for(int const k: picked)
            {
                auto& nk{adj[static_cast<std::size_t>(k)]};
                int a{-1}, b{-1};
                if(!nk.empty()) { a = *nk.begin(); }
                if(nk.size() == 2) { b = *std::next(nk.begin()); }
                int const eka{a >= 0 ? edge_of(k, a) : -1}, ekb{b >= 0 ? edge_of(k, b) : -1};
                int eab{-1};
                if(a >= 0 && b >= 0)
                {
                    eab = edge_of(a, b);
                    adj[static_cast<std::size_t>(a)].insert(b);
                    adj[static_cast<std::size_t>(b)].insert(a);
                }
                for(int const v: {a, b})
                {
                    if(v >= 0) { adj[static_cast<std::size_t>(v)].erase(k); }
                }
                gone[static_cast<std::size_t>(k)] = 1;
                for(std::int32_t const v: {k, a, b, eka, ekb, eab}) { s.ops.push_back(v); }
                for(int const v: {a, b})
                {
                    if(v >= 0) { next.push_back(v); }
                }
            }
            for(int const v: touched) { blocked[static_cast<std::size_t>(v)] = 0; }
            s.level_off.push_back(static_cast<std::int32_t>(s.ops.size() / 6));
            // candidates of the next round: what was skipped this round, and the neighbours whose degree may have dropped
            for(int const k: cand)
            {
                if(!gone[static_cast<std::size_t>(k)]) { next.push_back(k); }
            }
            std::sort(next.begin(), next.end());
            next.erase(std::unique(next.begin(), next.end()), next.end());
            cand.clear();
            for(int const k: next)
            {
                if(!gone[static_cast<std::size_t>(k)] && !core[static_cast<std::size_t>(k)] && adj[static_cast<std::size_t>(k)].size() <= 2) { cand.push_back(k); }
            }
        }
        // the dense core: remaining nodes first, then the source branches
        std::vector<int> core_row(static_cast<std::size_t>(s.n_nodes), -1);
        for(int k{}; k < s.n_nodes; ++k)
        {
            if(!gone[static_cast<std::size_t>(k)])
            {
                core_row[static_cast<std::size_t>(k)] = static_cast<int>(s.core_unknown.size());
                s.core_unknown.push_back(k);
            }
        }
        s.n_core_nodes = static_cast<int>(s.core_unknown.size());
        for(int const ei: s.vdc_elem)
        {
            auto const& e{nl.elems[static_cast<std::size_t>(ei)]};
            int const a{unk(e.pin_node[0])}, b{unk(e.pin_node[1])};
            s.vdc.push_back(0);
            s.vdc.push_back(a >= 0 ? core_row[static_cast<std::size_t>(a)] : -1);
            s.vdc.push_back(b >= 0 ? core_row[static_cast<std::size_t>(b)] : -1);
            s.core_unknown.push_back(num.n_nodes + num.branch0[static_cast<std::size_t>(ei)]);
        }
        s.n_core = static_cast<int>(s.core_unknown.size());
        s.ld_core = round_up(std::max(s.n_core, 1), k_nb);
        for(int k{}; k < s.n_nodes; ++k)
        {
            if(gone[static_cast<std::size_t>(k)]) { continue; }
            for(int const v: adj[static_cast<std::size_t>(k)])
            {
                if(v > k && !gone[static_cast<std::size_t>(v)])
                {
                    s.core_edges.push_back(core_row[static_cast<std::size_t>(k)]);
                    s.core_edges.push_back(core_row[static_cast<std::size_t>(v)]);
                    s.core_edges.push_back(edge_of(k, v));
                }
            }
        }
    }

    // one DC solve of every instance of the batch; the solution lands in the rows x[unknown][instance] of b.d_wi
    bool frontal_run(batch& b, frontal_state& s)
    {
        circuit& c{*b.parent};
        auto fail = [&](std::string m)
        {
            b.error = std::move(m);
            set_last_error(b.error);
            return false;
        };
        if(pe_b200_dev_count() <= 0) { return fail("no CUDA device visible: the B200 MNA path has no CPU fallback"); }
        if(pe_b200_dev_set(b.device) != 0) { return fail(std::string{"set device: "} + pe_b200_dev_last_error()); }
        if(s.num_rev != c.structure_rev)
        {
            s.num = make_numbering(c.nl);  // O(elements): once per structure, not once per solve
            s.num_rev = c.structure_rev;
        }
        numbering const& num{s.num};
        bool const rebuild{s.structure_rev != c.structure_rev || s.n_unknowns != num.unknowns()};
        if(rebuild)
        {
            frontal_build(s, c.nl, num);
            s.structure_rev = c.structure_rev;
            s.n_inst_built = 0;
        }
        std::size_t const n_inst{b.n_inst};
        std::int64_t const B{static_cast<std::int64_t>((n_inst + 31) / 32 * 32)};
        b.LSi = B;
        if(!b.cc || b.cc_structure_rev != c.structure_rev)
        {
            // a program-less compiled record: numbering only (the accessors of the batch read x from the first rows of d_wi)
            b.cc = std::make_unique<compiled>();
            b.cc->num = num;
            b.cc->n_inst_slots = num.unknowns();
            b.cc_structure_rev = c.structure_rev;
        }
        std::size_t const n{static_cast<std::size_t>(num.unknowns())};
        auto up = [&](device_buf& d, void const* p, std::size_t bytes) -> bool
        { return d.ensure(std::max<std::size_t>(bytes, 8)) && (bytes == 0 || pe_b200_dev_h2d(d.p, p, bytes, b.stream) == 0); };
        std::size_t const n_res{s.res_elem.size()}, n_idc{s.idc_elem.size()}, n_vdc{s.vdc_elem.size()};
        if(rebuild || s.n_inst_built != n_inst)
        {
            bool ok{up(s.d_ia, s.res_a.data(), n_res * 4) && up(s.d_ib, s.res_b.data(), n_res * 4) && up(s.d_ie, s.res_edge.data(), n_res * 4) &&
                    up(s.d_ip, s.idc_p.data(), n_idc * 4) && up(s.d_iq, s.idc_q.data(), n_idc * 4) && up(s.d_ops, s.ops.data(), s.ops.size() * 4) &&
                    up(s.d_cu, s.core_unknown.data(), s.core_unknown.size() * 4) && up(s.d_ce, s.core_edges.data(), s.core_edges.size() * 4) &&
                    up(s.d_vdc, s.vdc.data(), s.vdc.size() * 4)};
            std::size_t const BB{static_cast<std::size_t>(B)};
            ok = ok && s.d_d.ensure(std::max<std::size_t>(static_cast<std::size_t>(s.n_nodes), 1) * BB * 8) && s.d_z.ensure(std::max<std::size_t>(static_cast<std::size_t>(s.n_nodes), 1) * BB * 8) &&
                 s.d_g.ensure(std::max<std::size_t>(static_cast<std::size_t>(s.n_edges), 1) * BB * 8) &&
                 s.d_M.ensure(static_cast<std::size_t>(s.ld_core) * static_cast<std::size_t>(s.ld_core) * n_inst * 8) && s.d_c.ensure(static_cast<std::size_t>(s.ld_core) * n_inst * 8) &&
                 b.d_wi.ensure(std::max<std::size_t>(n, 1) * BB * 8) && b.d_status.ensure(BB * 4) && b.d_solves.ensure(BB * 4);
            if(!ok) { return fail(std::string{"frontal: device allocation / upload: "} + pe_b200_dev_last_error()); }
            s.n_inst_built = n_inst;
        }
        // element values: the netlist's, overridden per instance by the sweeps
        auto values = [&](std::vector<int> const& elems, device_buf& d) -> bool
        {
            std::vector<double> v(std::max<std::size_t>(elems.size(), 1) * static_cast<std::size_t>(B), 1.0);
            for(std::size_t k{}; k < elems.size(); ++k)
            {
                auto const sw{b.sweeps.find({elems[k], 0})};
                for(std::size_t i{}; i < n_inst; ++i)
                {
                    v[k * static_cast<std::size_t>(B) + i] = (sw != b.sweeps.end() && sw->second.size() >= n_inst) ? sw->second[i] : c.nl.elems[static_cast<std::size_t>(elems[k])].attr[0];
                }
            }
            return up(d, v.data(), v.size() * 8) && pe_b200_dev_sync(b.stream) == 0;  // `v` is pageable host memory
        };
        // the value tables (lane-interleaved, 26 MB for config A) travel again only when a parameter changed
        bool const values_current{!rebuild && s.val_sweeps_rev == b.sweeps_rev && s.val_param_rev == c.param_rev && s.val_n_inst == n_inst && s.val_batch == &b};
        if(!values_current)
        {
            if(!values(s.res_elem, s.d_rval) || !values(s.idc_elem, s.d_idc) || !values(s.vdc_elem, s.d_vdcv)) { return fail(std::string{"frontal: upload values: "} + pe_b200_dev_last_error()); }
            s.val_sweeps_rev = b.sweeps_rev;
            s.val_param_rev = c.param_rev;
            s.val_n_inst = n_inst;
            s.val_batch = &b;
        }
        if(pe_b200_dev_memset0(b.d_status.p, b.d_status.bytes, b.stream) != 0 || pe_b200_dev_memset0(b.d_solves.p, b.d_solves.bytes, b.stream) != 0 ||
           pe_b200_dev_memset0(b.d_wi.p, std::max<std::size_t>(n, 1) * static_cast<std::size_t>(B) * 8, b.stream) != 0)
        {
            return fail(std::string{"frontal: zero: "} + pe_b200_dev_last_error());
        }
        pe_b200_frontal f{};
        f.n_inst = static_cast<std::int32_t>(n_inst);
        f.B = B;
        f.n_nodes = s.n_nodes;
        f.n_res = static_cast<std::int32_t>(n_res);
        f.n_idc = static_cast<std::int32_t>(n_idc);
        f.n_vdc = static_cast<std::int32_t>(n_vdc);
        f.n_edges = s.n_edges;
        f.rval = static_cast<double const*>(s.d_rval.p);
        f.res_a = static_cast<std::int32_t const*>(s.d_ia.p);
        f.res_b = static_cast<std::int32_t const*>(s.d_ib.p);
        f.res_edge = static_cast<std::int32_t const*>(s.d_ie.p);
        f.idc_p = static_cast<std::int32_t const*>(s.d_ip.p);
        f.idc_q = static_cast<std::int32_t const*>(s.d_iq.p);
        f.idc_val = static_cast<double const*>(s.d_idc.p);
        f.ops = static_cast<std::int32_t const*>(s.d_ops.p);
        f.level_off_host = s.level_off.data();
        f.n_levels = static_cast<std::int32_t>(s.level_off.size()) - 1;
        f.n_core = s.n_core;
        f.n_core_nodes = s.n_core_nodes;
        f.n_core_edges = static_cast<std::int32_t>(s.core_edges.size() / 3);
        f.ld_core = s.ld_core;
        f.core_unknown = static_cast<std::int32_t const*>(s.d_cu.p);
        f.core_edges = static_cast<std::int32_t const*>(s.d_ce.p);
        f.vdc = static_cast<std::int32_t const*>(s.d_vdc.p);
        f.vdc_val = static_cast<double const*>(s.d_vdcv.p);
        f.d = static_cast<double*>(s.d_d.p);
        f.z = static_cast<double*>(s.d_z.p);
        f.g = static_cast<double*>(s.d_g.p);
        f.M = static_cast<double*>(s.d_M.p);
        f.c = static_cast<double*>(s.d_c.p);
        f.x = static_cast<double*>(b.d_wi.p);
        f.LSx = B;
        f.status = static_cast<std::int32_t*>(b.d_status.p);
        f.phase_ms_host = s.phase_ms;
        std::uint64_t nl{};
        if(pe_b200_frontal_run(&f, b.stream, &nl) != 0) { return fail(std::string{"frontal: "} + pe_b200_frontal_last_error()); }
        s.launches = nl;
        b.last_lanes = n_inst;
        b.last_points = 1;
        b.last_cplx = false;
        b.last_LSl = B;
        b.last_jit = 3;
        std::vector<std::int32_t> st(n_inst);
        if(pe_b200_dev_d2h(st.data(), b.d_status.p, n_inst * 4, b.stream) != 0 || pe_b200_dev_sync(b.stream) != 0) { return fail(std::string{"frontal: status: "} + pe_b200_dev_last_error()); }
        b.total_solves = n_inst;
        std::vector<std::uint32_t> ones(n_inst, 1u);
        (void)pe_b200_dev_h2d(b.d_solves.p, ones.data(), n_inst * 4, b.stream);
        (void)pe_b200_dev_sync(b.stream);
        bool all_ok{true};
        for(auto const v: st) { all_ok = all_ok && v == PE_ST_OK; }
        if(!all_ok) { return fail("analyze: at least one lane failed (singular matrix)"); }
        return true;
    }
}  // namespace pe_b200
