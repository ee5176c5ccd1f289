// TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or shipped with the product library.
//
// oracle/_ref/libpe_ref.so = the UNMODIFIED Phy-Engine reference (its own C ABI, src/dll_main.cpp, compiled from
// the sources where they lie under /root/reference) plus the small `ref_*` probe functions below, which only
// *read* reference state or *drive* the reference's own public members (`circult::prepare / solve_once /
// update_tr_step`, all public: include/phy_engine/circuits/circuit.h:60-179,363,468,892,987).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Build: see oracle/Makefile (g++ -std=c++23 -I/root/reference/include). No reference source is copied here:
// the single #include below pulls the reference TU in at compile time, exactly like the reference's own
// test/0008.dll/dll_main_smoke.cpp:33 does.

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "src/dll_main.cpp"  // resolved through -I/root/reference

namespace
{
    using circ_t = ::phy_engine::circult;

    // Newton driver with a solve counter. Mirrors circult::solve() (circuit.h:892-985) statement by statement, but
    // around the reference's *own* solve_once(); used only so that tests can compare iteration counts.
    // tests/test_oracle_ref.py checks that it returns bit-identical state to circult::analyze().
    bool counted_solve(circ_t& c, ::std::uint64_t& count) noexcept
    {
        if(c.at == ::phy_engine::analyze_type::AC)
        {
            ++count;
            return c.solve_once();
        }
        if(!c.has_nonlinear_device())
        {
            ++count;
            return c.solve_once();
        }

        constexpr ::std::size_t max_iter{64};
        auto const& env{c.env};
        double const v_abstol{env.V_eps_max > 0.0 ? env.V_eps_max : 1e-6};
        double const v_reltol{env.V_epsr_max > 0.0 ? env.V_epsr_max : 1e-3};
        double const i_abstol{env.I_eps_max > 0.0 ? env.I_eps_max : 1e-12};
        double const i_reltol{env.I_epsr_max > 0.0 ? env.I_epsr_max : v_reltol};

        ::std::vector<::std::complex<double>> pn(c.node_counter), pb(c.size_t_to_branch_p.size());
        for(::std::size_t iter{}; iter < max_iter; ++iter)
        {
            for(::std::size_t i{}; i < c.node_counter; ++i) { pn[i] = c.size_t_to_node_p.index_unchecked(i)->node_information.an.voltage; }
            for(::std::size_t i{}; i < pb.size(); ++i) { pb[i] = c.size_t_to_branch_p.index_unchecked(i)->current; }
            ++count;
            if(!c.solve_once()) { return false; }
            bool converged{true};
            for(::std::size_t i{}; i < c.node_counter; ++i)
            {
                auto const v_new{c.size_t_to_node_p.index_unchecked(i)->node_information.an.voltage};
                double const tol{v_abstol + v_reltol * ::std::max(::std::abs(v_new), ::std::abs(pn[i]))};
                if(::std::abs(v_new - pn[i]) > tol)
                {
                    converged = false;
                    break;
                }
            }
            if(converged)
            {
                for(::std::size_t i{}; i < pb.size(); ++i)
                {
                    auto const i_new{c.size_t_to_branch_p.index_unchecked(i)->current};
                    double const tol{i_abstol + i_reltol * ::std::max(::std::abs(i_new), ::std::abs(pb[i]))};
                    if(::std::abs(i_new - pb[i]) > tol)
                    {
                        converged = false;
                        break;
                    }
                }
            }
            // check_convergence(): no in-scope model defines it; save_op(): only bsim3v32 (out of scope) defines it.
            if(converged)
            {
                if(c.at == ::phy_engine::analyze_type::OP || c.at == ::phy_engine::analyze_type::DC || c.at == ::phy_engine::analyze_type::TROP)
                {
                    for(auto& i: c.nl.models)
                    {
                        for(auto m{i.begin}; m != i.curr; ++m)
                        {
                            if(m->type != ::phy_engine::model::model_type::normal || m->ptr == nullptr) { continue; }
                            (void)m->ptr->save_op();
                        }
                    }
                }
                return true;
            }
        }
        return false;
    }

    bool counted_ac(circ_t& c, ::std::uint64_t& count) noexcept
    {
        auto& ac{c.analyzer_setting.ac};
        using sweep_t = ::phy_engine::analyzer::AC::sweep_type;
        if(ac.sweep == sweep_t::single || ac.points <= 1) { return counted_solve(c, count); }
        c.ac_sweep_results.clear();
        if(ac.sweep == sweep_t::linear)
        {
            double const step{(ac.omega_stop - ac.omega_start) / static_cast<double>(ac.points - 1)};
            for(::std::size_t idx{}; idx < ac.points; ++idx)
            {
                ac.omega = ac.omega_start + step * static_cast<double>(idx);
                if(!counted_solve(c, count)) { return false; }
                c.ac_sweep_results.push_back({ac.omega, c.capture_solution_vector()});
            }
            return true;
        }
        if(ac.sweep == sweep_t::log)
        {
            if(ac.omega_start <= 0.0 || ac.omega_stop <= 0.0) { return false; }
            double const ratio{::std::pow(ac.omega_stop / ac.omega_start, 1.0 / static_cast<double>(ac.points - 1))};
            double omega{ac.omega_start};
            for(::std::size_t idx{}; idx < ac.points; ++idx)
            {
                ac.omega = omega;
                if(!counted_solve(c, count)) { return false; }
                c.ac_sweep_results.push_back({ac.omega, c.capture_solution_vector()});
                omega *= ratio;
            }
            return true;
        }
        return counted_solve(c, count);
    }

    // Mirrors circult::analyze() (circuit.h:179-296) with counted_solve in place of solve().
    bool counted_analyze(circ_t& c, ::std::uint64_t& count) noexcept
    {
        using at_t = ::phy_engine::analyze_type;
        switch(c.at)
        {
            case at_t::OP: [[fallthrough]];
            case at_t::DC:
            {
                c.prepare();
                return counted_solve(c, count);
            }
            case at_t::AC:
            {
                c.prepare();
                c.clear_ac_sweep_results();
                if(c.has_nonlinear_device())
                {
                    auto const saved{c.at};
                    c.at = at_t::OP;
                    bool const ok{counted_solve(c, count)};
                    c.at = saved;
                    if(!ok) { return false; }
                }
                return counted_ac(c, count);
            }
            case at_t::ACOP:
            {
                c.prepare();
                c.clear_ac_sweep_results();
                auto const saved{c.at};
                c.at = at_t::OP;
                if(!counted_solve(c, count))
                {
                    c.at = saved;
                    return false;
                }
                c.at = at_t::AC;
                bool const ok{counted_ac(c, count)};
                c.at = saved;
                return ok;
            }
            case at_t::TR: [[fallthrough]];
            case at_t::TROP:
            {
                auto const dt{c.analyzer_setting.tr.t_step};
                if(dt <= 0.0) { return false; }
                auto const t_stop{c.analyzer_setting.tr.t_stop};
                c.prepare();
                auto const saved{c.at};
                if(c.at == at_t::TROP)
                {
                    if(!counted_solve(c, count)) { return false; }
                    c.at = at_t::TR;
                }
                auto const end_time{c.tr_duration + t_stop};
                for(; c.tr_duration < end_time;)
                {
                    c.update_tr_step(dt);
                    auto const prev{c.tr_duration};
                    c.tr_duration = prev + dt;
                    if(!counted_solve(c, count))
                    {
                        c.tr_duration = prev;
                        c.at = saved;
                        return false;
                    }
                }
                c.at = saved;
                return true;
            }
            default: return false;
        }
    }

    void apply_env(circ_t& c, double const* env10) noexcept
    {
        if(env10 == nullptr) { return; }
        c.env.V_eps_max = env10[0];
        c.env.V_epsr_max = env10[1];
        c.env.I_eps_max = env10[2];
        c.env.I_epsr_max = env10[3];
        c.env.g_min = env10[4];
        c.env.r_open = env10[5];
        c.env.temperature = env10[6];
        c.env.norm_temperature = env10[7];
    }
}  // namespace

extern "C"
{
    // --- probes on a reference circuit handle (void* = phy_engine::circult*, as returned by create_circuit) ---

    int ref_set_env(void* p, double const* env8)
    {
        if(p == nullptr || env8 == nullptr) { return 1; }
        apply_env(*static_cast<circ_t*>(p), env8);
        return 0;
    }

    int ref_set_ac_sweep(void* p, int sweep, double w_start, double w_stop, ::std::size_t points)
    {
        if(p == nullptr) { return 1; }
        auto& ac{static_cast<circ_t*>(p)->analyzer_setting.ac};
        ac.sweep = static_cast<::phy_engine::analyzer::AC::sweep_type>(sweep);
        ac.omega_start = w_start;
        ac.omega_stop = w_stop;
        ac.points = points;
        return 0;
    }

    int ref_counts(void* p, ::std::size_t* n_nodes, ::std::size_t* n_branches)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circ_t*>(p)};
        if(n_nodes) { *n_nodes = c->node_counter; }
        if(n_branches) { *n_branches = c->branch_counter; }
        return 0;
    }

    // x as interleaved (re,im) pairs, length 2*(nodes+branches), unknown order = node_index then node_counter+branch.index
    int ref_get_solution(void* p, double* x)
    {
        if(p == nullptr || x == nullptr) { return 1; }
        auto* c{static_cast<circ_t*>(p)};
        auto const v{c->capture_solution_vector()};
        for(::std::size_t i{}; i < v.size(); ++i)
        {
            x[2 * i] = v.index_unchecked(i).real();
            x[2 * i + 1] = v.index_unchecked(i).imag();
        }
        return 0;
    }

    ::std::size_t ref_ac_result_count(void* p) { return p ? static_cast<circ_t*>(p)->ac_sweep_results.size() : 0; }

    // omegas[points], x[points][2*n]
    int ref_ac_results(void* p, double* omegas, double* x)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circ_t*>(p)};
        ::std::size_t const n{c->node_counter + c->branch_counter};
        for(::std::size_t k{}; k < c->ac_sweep_results.size(); ++k)
        {
            auto const& r{c->ac_sweep_results.index_unchecked(k)};
            if(omegas) { omegas[k] = r.omega; }
            if(x)
            {
                for(::std::size_t i{}; i < r.x.size() && i < n; ++i)
                {
                    x[(k * n + i) * 2] = r.x.index_unchecked(i).real();
                    x[(k * n + i) * 2 + 1] = r.x.index_unchecked(i).imag();
                }
            }
        }
        return 0;
    }

    double ref_tr_duration(void* p) { return p ? static_cast<circ_t*>(p)->tr_duration : 0.0; }

    int ref_reset(void* p)
    {
        if(p == nullptr) { return 1; }
        static_cast<circ_t*>(p)->reset();
        return 0;
    }

    int ref_prepare(void* p)
    {
        if(p == nullptr) { return 1; }
        static_cast<circ_t*>(p)->prepare();
        return 0;
    }

    int ref_solve_once(void* p)
    {
        if(p == nullptr) { return 1; }
        return static_cast<circ_t*>(p)->solve_once() ? 0 : 1;
    }

    int ref_analyze_counted(void* p, ::std::uint64_t* solves)
    {
        if(p == nullptr) { return 1; }
        ::std::uint64_t cnt{};
        bool const ok{counted_analyze(*static_cast<circ_t*>(p), cnt)};
        if(solves) { *solves = cnt; }
        return ok ? 0 : 1;
    }

    // unknown index of a pin's node: >=0 node_index; -1 ground; -2 unconnected / not analog
    long long ref_pin_node_index(void* p, ::std::size_t vec_pos, ::std::size_t chunk_pos, ::std::size_t pin)
    {
        if(p == nullptr) { return -2; }
        auto* c{static_cast<circ_t*>(p)};
        auto* m{get_model(c->nl, ::phy_engine::netlist::model_pos{vec_pos, chunk_pos})};
        if(m == nullptr || m->ptr == nullptr) { return -2; }
        auto const pv{m->ptr->generate_pin_view()};
        if(pin >= pv.size || pv.pins[pin].nodes == nullptr) { return -2; }
        auto const* n{pv.pins[pin].nodes};
        if(n == &c->nl.ground_node) { return -1; }
        if(n->num_of_analog_node == 0) { return -2; }
        return static_cast<long long>(n->node_index);
    }

    long long ref_branch_index(void* p, ::std::size_t vec_pos, ::std::size_t chunk_pos, ::std::size_t br)
    {
        if(p == nullptr) { return -2; }
        auto* c{static_cast<circ_t*>(p)};
        auto* m{get_model(c->nl, ::phy_engine::netlist::model_pos{vec_pos, chunk_pos})};
        if(m == nullptr || m->ptr == nullptr) { return -2; }
        auto const bv{m->ptr->generate_branch_view()};
        if(br >= bv.size) { return -2; }
        return static_cast<long long>(bv.branches[br].index);
    }

    // The stamped system left in c.mna after the last solve_once() (circuit.h:999-1003 keeps it until the next call).
    ::std::size_t ref_mna_nnz(void* p)
    {
        if(p == nullptr) { return 0; }
        auto* c{static_cast<circ_t*>(p)};
        ::std::size_t nnz{};
        for(auto const& row: c->mna.A) { nnz += row.size(); }
        return nnz;
    }

    // rows/cols/vals(re,im) in row-major order; z[2*n] dense rhs
    int ref_mna_dump(void* p, int* rows, int* cols, double* vals, double* z)
    {
        if(p == nullptr) { return 1; }
        auto* c{static_cast<circ_t*>(p)};
        ::std::size_t k{};
        ::std::size_t const n{c->mna.A.size()};
        for(::std::size_t r{}; r < n; ++r)
        {
            for(auto const& [col, v]: c->mna.A[r])
            {
                rows[k] = static_cast<int>(r);
                cols[k] = static_cast<int>(col);
                vals[2 * k] = v.real();
                vals[2 * k + 1] = v.imag();
                ++k;
            }
        }
        if(z)
        {
            for(::std::size_t i{}; i < 2 * n; ++i) { z[i] = 0.0; }
            for(auto const& [r, v]: c->mna.Z)
            {
                if(r < n)
                {
                    z[2 * r] = v.real();
                    z[2 * r + 1] = v.imag();
                }
            }
        }
        return 0;
    }

    ::std::size_t ref_sizeof_model_base(void) { return sizeof(::phy_engine::model::model_base); }
    ::std::size_t ref_sizeof_node(void) { return sizeof(::phy_engine::model::node_t); }

    // --- batch CPU baseline: the reference's own analyze() over many independent instances, T worker threads ---
    //
    // Each worker owns its circuits (circult is not thread-safe; distinct circuits on distinct threads are fine).
    // Instance i = base netlist with overrides (comp over_comp[k], attribute over_name[k]) := over_vals[k*n_inst + i].
    // A fresh circuit is created per instance (state-free start, like a new Monte-Carlo sample); only the time spent
    // inside analyze() is accumulated, so netlist construction does not count against the reference.
    // Returns 0; seconds_out = max over workers of their summed analyze() time.
    // x_out: n_inst * 2 * n doubles (final state) or, for an AC sweep with points>1, n_inst * points * 2 * n.
    int ref_run_batch(int* elements,
                      ::std::size_t ele_size,
                      int* wires,
                      ::std::size_t wires_size,
                      double* properties,
                      ::std::uint32_t analyze_type,
                      double t_step,
                      double t_stop,
                      int ac_sweep,
                      double w_start,
                      double w_stop,
                      ::std::size_t points,
                      double const* env8,
                      ::std::size_t n_inst,
                      ::std::size_t n_over,
                      ::std::size_t const* over_comp,
                      char const* const* over_name,
                      double const* over_vals,
                      int threads,
                      double* x_out,
                      ::std::uint64_t* solves_out,
                      int* ok_out,
                      double* seconds_out)
    {
        if(threads < 1) { threads = 1; }
        ::std::vector<double> tsec(static_cast<::std::size_t>(threads), 0.0);
        ::std::atomic<::std::size_t> next{0};
        auto worker = [&](int tid)
        {
            using clock = ::std::chrono::steady_clock;
            for(;;)
            {
                ::std::size_t const i{next.fetch_add(1)};
                if(i >= n_inst) { break; }
                ::std::size_t* vp{};
                ::std::size_t* cp{};
                ::std::size_t cs{};
                void* h{create_circuit(elements, ele_size, wires, wires_size, properties, &vp, &cp, &cs)};
                if(h == nullptr)
                {
                    if(ok_out) { ok_out[i] = -1; }
                    continue;
                }
                auto* c{static_cast<circ_t*>(h)};
                apply_env(*c, env8);
                c->at = static_cast<::phy_engine::analyze_type>(analyze_type);
                c->analyzer_setting.tr.t_step = t_step;
                c->analyzer_setting.tr.t_stop = t_stop;
                c->analyzer_setting.ac.sweep = static_cast<::phy_engine::analyzer::AC::sweep_type>(ac_sweep);
                c->analyzer_setting.ac.omega_start = w_start;
                c->analyzer_setting.ac.omega_stop = w_stop;
                c->analyzer_setting.ac.omega = w_start;
                c->analyzer_setting.ac.points = points;
                for(::std::size_t k{}; k < n_over; ++k)
                {
                    auto const ci{over_comp[k]};
                    if(ci >= cs) { continue; }
                    (void)circuit_set_model_double_by_name(h, vp[ci], cp[ci], over_name[k], ::std::strlen(over_name[k]), over_vals[k * n_inst + i]);
                }
                ::std::uint64_t cnt{};
                auto const t0{clock::now()};
                bool const ok{counted_analyze(*c, cnt)};
                auto const t1{clock::now()};
                tsec[static_cast<::std::size_t>(tid)] += ::std::chrono::duration<double>(t1 - t0).count();
                if(solves_out) { solves_out[i] = cnt; }
                if(ok_out) { ok_out[i] = ok ? 1 : 0; }
                if(x_out)
                {
                    ::std::size_t const n{c->node_counter + c->branch_counter};
                    bool const sweep{(c->at == ::phy_engine::analyze_type::AC || c->at == ::phy_engine::analyze_type::ACOP) && ac_sweep != 0 && points > 1};
                    if(sweep)
                    {
                        ref_ac_results(h, nullptr, x_out + i * points * 2 * n);
                    }
                    else
                    {
                        ref_get_solution(h, x_out + i * 2 * n);
                    }
                }
                destroy_circuit(h, vp, cp);
            }
        };
        ::std::vector<::std::thread> th;
        for(int t{1}; t < threads; ++t) { th.emplace_back(worker, t); }
        worker(0);
        for(auto& t: th) { t.join(); }
        double mx{};
        for(double s: tsec) { mx = ::std::max(mx, s); }
        if(seconds_out) { *seconds_out = mx; }
        return 0;
    }
}
