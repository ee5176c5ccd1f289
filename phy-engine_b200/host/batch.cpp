// batch.cpp — the analysis driver: circult::analyze() (circuits/circuit.h:179-296) re-stated as a launch plan
// over the sm_100a kernels, for B independent instances at once.  The host only sequences phases and moves
// parameters/results; every solve runs on the device.  No CUDA device => analyze() fails with an error.
#include <algorithm>
#include <functional>
#include <cmath>
#include <cstring>

#include "pe_host.hpp"

#include <cstdio>
#include <cstdlib>

namespace pe_b200
{
    namespace
    {
        thread_local std::string g_last_error;

        std::int64_t round_up32(std::size_t n) { return static_cast<std::int64_t>((n + 31) / 32 * 32); }

        bool dev_fail(std::string& err, char const* what)
        {
            err = std::string{what} + ": " + pe_b200_dev_last_error();
            set_last_error(err);
            return false;
        }
    }  // namespace

    path_defaults& default_path()
    {
        static path_defaults d;
        return d;
    }

    void set_last_error(std::string s) { g_last_error = std::move(s); }

    char const* last_error() { return g_last_error.c_str(); }

    device_buf::~device_buf() { release(); }

    void device_buf::release()
    {
        if(p) { (void)pe_b200_dev_free(p); }
        p = nullptr;
        bytes = 0;
    }

    bool device_buf::ensure(std::size_t n)
    {
        if(n <= bytes && p != nullptr) { return true; }
        release();
        if(pe_b200_dev_malloc(&p, n) != 0)
        {
            p = nullptr;
            return false;
        }
        bytes = n;
        return true;
    }

    bool batch::compile_host(bool& layout_change)
    {
        std::vector<sweep_key> keys;
        for(auto const& [k, v]: sweeps) { keys.push_back(k); }
        layout_change = !cc || cc_structure_rev != parent->structure_rev || keys != layout_keys;
        auto const& acs{ac.points > 0 || ac.omega != 0.0 || ac.omega_start != 0.0 ? ac : parent->ac};
        std::size_t const ac_points{(acs.sweep == sweep_type::single || acs.points <= 1) ? 1u : acs.points};
        int const n_unk{make_numbering(parent->nl).unknowns()};
        int const w_real{pick_warps(n_inst, n_unk)};
        int const w_ac{pick_warps(n_inst * ac_points, n_unk)};
        int r_real{pick_streams(n_unk)};
        int r_ac{pick_streams(n_unk)};
        // Stream kernel (host/stream.cpp): large batches of large linear circuits are compiled with ONE stream per lane group
        // (no parallel split of the elimination, no fill-in from it) and run one warp per group with TMA-fed tiles.  The
        // program of the analysis to be run must consist of ops the generator covers, else the default geometry is used.
        auto const at_now{parent->at};
        bool const real_analysis{at_now == analyze_type::TR || at_now == analyze_type::TROP || at_now == analyze_type::OP || at_now == analyze_type::DC};
        bool want_stream{res_stream >= 0 && res_S == 0 && res_ws != 1 && real_analysis && n_unk > 64 && !parent->nl.has_nonlinear() && pe_b200_stream_supported() != 0 &&
                         stream_rejected_rev != parent->structure_rev &&
                         (res_stream == 1 || (n_inst >= 4096 && std::getenv("PE_B200_NO_STREAM") == nullptr))};
        if(want_stream) { r_real = 1; }
        stream_mode = want_stream;  // decides where a one-stream program keeps its workspace (use_hbm)
        bool const need{layout_change || cc_param_rev != parent->param_rev || cc_dt != parent->tr.t_step || w_real != cc_warps_real || w_ac != cc_warps_ac ||
                        r_real != cc_res_real || r_ac != cc_res_ac || res_fuse != cc_fuse || guard_all != cc_guard_all};
        if(!need) { return true; }
        layout_keys = keys;

        compile_input in;
        in.nl = &parent->nl;
        in.env = parent->env;
        in.dt = parent->tr.t_step;
        auto const& a{ac.points > 0 || ac.omega != 0.0 || ac.omega_start != 0.0 ? ac : parent->ac};
        in.omega0 = (a.sweep == sweep_type::single || a.points <= 1) ? a.omega : std::sqrt(std::fabs(a.omega_start * a.omega_stop));
        if(omega0_override > 0.0) { in.omega0 = omega0_override; }
        if(!(in.omega0 > 0.0)) { in.omega0 = 1.0; }
        in.guard_all = guard_all;
        in.nl_nominal = nl_nominal;
        for(auto const& [k, v]: sweeps) { in.swept_lane0[k] = v.empty() ? 0.0 : v[0]; }
        in.warps_real = w_real;
        in.warps_ac = w_ac;
        cc_warps_real = w_real;
        cc_warps_ac = w_ac;
        cc_res_real = r_real;
        cc_res_ac = r_ac;
        cc_fuse = res_fuse;
        cc_guard_all = guard_all;
        for(auto const& e: parent->nl.elems)
        {
            // guard elision (compiler.cpp) rests on positive R / C values
            if(e.d != nullptr && (e.d->code == E_RES || e.d->code == E_CAP) && !(e.attr[0] > 0.0)) { in.guard_all = true; }
        }
        for(int attempt{}; attempt < 3; ++attempt)
        {
            in.resident_real = r_real;
            in.resident_ac = r_ac;
            in.fuse_steps = res_fuse != 0;
            cc = compile_circuit(in);
            if(!cc)
            {
                error = "compile failed";
                set_last_error(error);
                return false;
            }
            // a resident program whose workspace does not fit the shared memory of one CTA falls back to the
            // HBM-streaming form (all real-valued modes together: they share the instance workspace layout)
            bool redo{false};
            bool real_fits{true};
            for(int m{}; m < static_cast<int>(prog_mode::COUNT); ++m)
            {
                auto const& pr{cc->prog[static_cast<std::size_t>(m)]};
                bool const is_ac{static_cast<prog_mode>(m) == prog_mode::AC};
                if((is_ac ? r_ac : r_real) <= 0) { continue; }
                int I{}, J{};
                bool const fits{pr.built && pr.resident && pick_geometry(pr, I, J)};
                if(!fits)
                {
                    if(is_ac) { r_ac = 0; }
                    else
                    {
                        real_fits = false;
                    }
                    redo = true;
                }
            }
            if(!real_fits) { r_real = 0; }
            if(want_stream && !redo)
            {
                auto const primary{(at_now == analyze_type::TR || at_now == analyze_type::TROP) ? prog_mode::TR : prog_mode::DC};
                auto& pp{cc->prog[static_cast<std::size_t>(primary)]};
                if(pp.built && pp.resident && pp.rS == 1 && stream_profitable(pp))
                {
                    // every real-valued program the generator covers is re-laid out for it (the others run the interpreter on
                    // the same workspace geometry)
                    for(auto const m: {prog_mode::DC, prog_mode::TR, prog_mode::TROP})
                    {
                        auto& q{cc->prog[static_cast<std::size_t>(m)]};
                        if(q.built && q.resident && q.rS == 1 && stream_profitable(q)) { (void)stream_prepare(q); }
                    }
                }
                else
                {
                    if(std::getenv("PE_B200_STREAM_DEBUG") != nullptr)
                    {
                        std::fprintf(stderr, "stream: rejected (built %d resident %d rS %d supported %d)\n", (int)pp.built, (int)pp.resident, pp.rS, (int)stream_supported(pp));
                    }
                    want_stream = false;
                    stream_mode = false;
                    stream_rejected_rev = parent->structure_rev;
                    r_real = pick_streams(n_unk);
                    cc_res_real = r_real;
                    redo = true;
                }
            }
            if(!redo) { break; }
        }
        stream_mode = want_stream;
        cc_structure_rev = parent->structure_rev;
        cc_param_rev = parent->param_rev;
        cc_dt = parent->tr.t_step;
        uploaded.fill(false);
        uploaded_ig.fill(0);
        device_stale = true;
        return true;
    }

    int batch::pick_warps(std::size_t lanes, int n_unknowns) const
    {
        // G warps cooperate on each block of 32 lanes.  With plenty of lanes one warp per block already fills the GPU;
        // a small batch of large circuits needs G > 1 to have enough resident warps to hide HBM latency.
        int g{subtree_warps};
        if(g <= 0)
        {
            std::size_t const blocks{(lanes + 31) / 32};
            std::size_t const want{148u * 40u};  // resident warps we would like to have
            g = 1;
            while(g < PE_MAX_WARPS && blocks * static_cast<std::size_t>(g) * 2 <= want) { g *= 2; }
        }
        g = std::clamp(g, 1, PE_MAX_WARPS);
        while(g > 1 && n_unknowns < 48 * g) { g /= 2; }  // leaves need enough rows to amortise the barriers
        return g;
    }

    int batch::pick_streams(int n_unknowns) const
    {
        if(res_S < 0) { return 0; }
        if(res_S > 0) { return res_S; }
        // small circuits: one thread per instance, workspace in shared memory (no barriers, no divergence, no HBM
        // traffic inside the Newton / time loops)
        if(n_unknowns <= 64 && res_ws != 2) { return 1; }
        if(res_ws == 1)
        {
            // shared-memory workspace, thread per stream: leaves of ~4 rows
            int s{1};
            while(s < 256 && s * 2 * 4 <= n_unknowns) { s *= 2; }
            return s;
        }
        // larger circuits stream their workspace through HBM, warp per sub-tree: enough sub-trees to keep ~32 warps per
        // 32 lanes in flight (latency hiding), leaves of at least ~16 rows
        int s{1};
        while(s < 32 && s * 2 * 16 <= n_unknowns) { s *= 2; }
        return s;
    }

    bool batch::use_hbm(program const& pr) const
    {
        if(res_ws == 2) { return true; }
        if(res_ws == 1) { return false; }
        return pr.rS > 1 || stream_mode;
    }

    bool batch::pick_geometry(program const& pr, int& I, int& J) const
    {
        if(use_hbm(pr))
        {
            // tree-streaming form: 32 J lanes per CTA, J of them per thread (0 = choose: two when every sub-tree warp of a
            // full-width CTA still leaves the GPU enough independent groups)
            // lanes per thread: the more lanes share one decode and the longer a workspace row (128 lanes = 1 KB), the
            // better, as long as there are enough groups for all SMs: 128-lane groups run on a cluster of two CTAs
            std::size_t const lanes_total{n_inst * std::max<std::size_t>(last_points_hint, 1)};
            // narrower programs (fewer sub-tree warps per group): two lanes per thread (512-byte rows) as soon as the batch still
            // fills every SM several times over -- config D, 1e6 frequency points on 8 streams: 40.9 -> 57.1 M points/s
            if(stream_mode && pr.rS == 1)
            {
                // stream kernel: one warp per lane group, one lane per thread (two lanes per thread halve the warps: 26.8 vs
                // 49.2 M solves/s on config B)
                J = res_J > 0 ? res_J : 1;
                std::size_t const lanes_total{n_inst * std::max<std::size_t>(last_points_hint, 1)};
                // lanes per group: full warps.  Narrow groups (16 / 8 lanes with the other lanes idle: twice / four times the
                // warps) were measured on config B, 10 000 lanes: 49.2 (32) / 45.9 (16) / 23.9 (8) M solves/s -- the memory system
                // wants 256-byte rows more than the schedulers want warps; PE_B200_STREAM_GL keeps the switch for experiments
                int gl{32};
                (void)lanes_total;
                if(char const* e{std::getenv("PE_B200_STREAM_GL")}; e != nullptr && J == 1)
                {
                    int const v{std::atoi(e)};
                    if(v == 8 || v == 16 || v == 32) { gl = v; }
                }
                I = gl * J;
                return true;
            }
            J = res_J > 0 ? res_J
                          : ((pr.rS >= 32 && lanes_total >= 74u * 128u) ? 4
                                                                        : (((pr.rS >= 32 && lanes_total >= 148u * 64u) || lanes_total >= 148u * 64u * 8u) ? 2 : 1));
            I = 32 * J;
            return pr.rS <= 32;
        }
        std::size_t const limit{pe_b200_resident_smem_limit()};
        std::size_t const per_inst{static_cast<std::size_t>(std::max(pr.r_slots, 1)) * sizeof(double)};
        int const S{pr.rS};
        int const i_fit{static_cast<int>(std::min<std::size_t>(limit / per_inst, 32))};
        int const i_min{std::max(1, 32 / S)};
        if(i_fit < i_min) { return false; }
        auto pow2_floor = [](int v)
        {
            int p{1};
            while(p * 2 <= v) { p *= 2; }
            return p;
        };
        if(res_I > 0)
        {
            I = res_I;
            J = res_J > 0 ? res_J : 1;
            if(I > i_fit || I % J != 0 || I / J < i_min || (I & (I - 1)) != 0 || S * (I / J) > 1024) { return false; }
            return true;
        }
        J = res_J > 0 ? res_J : ((S >= 32 && i_fit >= 2) ? 2 : 1);
        // prefer two CTAs per SM (one hides the barrier stalls of the other) when the workspace allows it
        int i{pow2_floor(i_fit)};
        if(i / 2 >= i_min * J && i / 2 >= J && S * (i / 2 / J) >= 128) { i /= 2; }
        while(i > i_min * J && S * (i / J) > 512) { i /= 2; }
        if(i < J) { J = 1; }
        if(i / J < i_min) { J = 1; }
        if(i / J < i_min) { return false; }
        I = i;
        return true;
    }

    bool batch::ensure_compiled()
    {
        if(pe_b200_dev_count() <= 0)
        {
            error = "no CUDA device visible: the B200 MNA path has no CPU fallback";
            set_last_error(error);
            return false;
        }
        if(pe_b200_dev_set(device) != 0) { return dev_fail(error, "set device"); }
        bool layout_change{};
        if(!compile_host(layout_change)) { return false; }
        if(layout_change) { layout_pending = true; }
        if(!device_stale) { return true; }
        device_stale = false;
        bool const lc{layout_pending};
        layout_pending = false;

        LSi = round_up32(n_inst);
        // instance workspace = persistent INST slots followed by the scratch of the real-valued programs (one U space)
        std::size_t const wi_bytes{static_cast<std::size_t>(cc->n_inst_slots + cc->n_real_lane_slots) * static_cast<std::size_t>(LSi) * sizeof(double)};
        if(lc)
        {
            if(!d_wi.ensure(wi_bytes)) { return dev_fail(error, "alloc instance workspace"); }
            if(pe_b200_dev_memset0(d_wi.p, wi_bytes, stream) != 0) { return dev_fail(error, "zero instance workspace"); }
            sweeps_dirty = true;
            tr_duration = 0.0;
            last_step = 0.0;
        }
        else if(d_wi.bytes < wi_bytes)
        {
            // derived slots were added behind the persistent ones: grow, keeping state is not possible -> reset
            if(!d_wi.ensure(wi_bytes)) { return dev_fail(error, "alloc instance workspace"); }
            if(pe_b200_dev_memset0(d_wi.p, wi_bytes, stream) != 0) { return dev_fail(error, "zero instance workspace"); }
            sweeps_dirty = true;
        }
        if(!d_status.ensure(static_cast<std::size_t>(LSi) * 4) || !d_solves.ensure(static_cast<std::size_t>(LSi) * 4)) { return dev_fail(error, "alloc status"); }
        return true;
    }

    bool batch::upload_sweeps()
    {
        if(!sweeps_dirty) { return true; }
        for(auto const& [k, v]: sweeps)
        {
            auto it{cc->swept_slot.find(k)};
            if(it == cc->swept_slot.end()) { continue; }
            auto* dst{static_cast<double*>(d_wi.p) + static_cast<std::int64_t>(it->second) * LSi};
            if(v.size() < n_inst)
            {
                error = "a swept parameter was set through the device-direct path and the workspace was re-laid out: set it again";
                set_last_error(error);
                return false;
            }
            if(pe_b200_dev_h2d(dst, v.data(), n_inst * sizeof(double), stream) != 0) { return dev_fail(error, "upload sweep"); }
        }
        sweeps_dirty = false;
        return true;
    }

    bool batch::run_phase(prog_mode m, bool with_prep, bool nonlinear, int n_steps, bool time_stepping, double t0, double dt, std::size_t lanes, int ppi)
    {
        int const mi{static_cast<int>(m)};
        auto& pr{cc->prog[static_cast<std::size_t>(mi)]};
        if(pr.resident) { return run_phase_resident(m, with_prep, nonlinear, n_steps, time_stepping, t0, dt, lanes, ppi); }
        last_I = last_J = last_S = 0;
        auto& d_w{d_words[static_cast<std::size_t>(mi)]};
        if(!uploaded[static_cast<std::size_t>(mi)])
        {
            if(!d_w.ensure(pr.words.size() * 4)) { return dev_fail(error, "alloc program"); }
            if(pe_b200_dev_h2d(d_w.p, pr.words.data(), pr.words.size() * 4, stream) != 0) { return dev_fail(error, "upload program"); }
            uploaded[static_cast<std::size_t>(mi)] = true;
        }
        // constants (dt patched in place)
        if(cc->dt_slot >= 0) { cc->cst[static_cast<std::size_t>(cc->dt_slot)] = dt; }
        if(!d_cst.ensure(cc->cst.size() * sizeof(double))) { return dev_fail(error, "alloc const table"); }
        if(pe_b200_dev_h2d(d_cst.p, cc->cst.data(), cc->cst.size() * sizeof(double), stream) != 0) { return dev_fail(error, "upload const table"); }

        std::int64_t const LSl{pr.cplx ? round_up32(lanes) : LSi};
        if(pr.cplx)
        {
            std::size_t const wl_bytes{static_cast<std::size_t>(std::max(pr.n_lane_slots, 1)) * static_cast<std::size_t>(LSl) * sizeof(double)};
            if(!d_wl.ensure(wl_bytes)) { return dev_fail(error, "alloc lane workspace"); }
        }
        if(!d_status.ensure(static_cast<std::size_t>(LSl) * 4) || !d_solves.ensure(static_cast<std::size_t>(LSl) * 4)) { return dev_fail(error, "alloc status"); }

        pe_b200_run r{};
        r.words = static_cast<std::uint32_t const*>(d_w.p);
        r.prep = pr.prep;
        if(!with_prep)
        {
            for(auto& o: r.prep.off) { o = PE_NO_SECTION; }
        }
        r.step = pr.step;
        r.iter = pr.iter;
        r.cst = static_cast<double const*>(d_cst.p);
        r.wu = static_cast<double*>(pr.cplx ? d_wl.p : d_wi.p);
        r.wx = static_cast<double*>(d_wi.p);
        r.warps = pr.warps;
        r.status = static_cast<std::int32_t*>(d_status.p);
        r.solves = static_cast<std::uint32_t*>(d_solves.p);
        r.wave = nullptr;
        r.probes = nullptr;
        r.n_probe = 0;
        if(time_stepping && !probes.empty())
        {
            std::vector<std::uint32_t> po;
            for(int u: probes) { po.push_back(pr.x_opnd[static_cast<std::size_t>(u)]); }
            if(!d_probes.ensure(po.size() * 4) || !d_wave.ensure(static_cast<std::size_t>(n_steps) * po.size() * static_cast<std::size_t>(LSl) * sizeof(double)))
            {
                return dev_fail(error, "alloc waveform store");
            }
            if(pe_b200_dev_h2d(d_probes.p, po.data(), po.size() * 4, stream) != 0) { return dev_fail(error, "upload probes"); }
            r.wave = static_cast<double*>(d_wave.p);
            r.probes = static_cast<std::uint32_t const*>(d_probes.p);
            r.n_probe = static_cast<std::int32_t>(po.size());
            wave_steps = static_cast<std::size_t>(n_steps);
            wave_pitch = LSl;
            // rows of lanes that fail and of the steps after a failure are never written: they read as zero, not as whatever
            // the allocation held
            if(pe_b200_dev_memset0(d_wave.p, static_cast<std::size_t>(n_steps) * po.size() * static_cast<std::size_t>(LSl) * sizeof(double), stream) != 0)
            {
                return dev_fail(error, "zero waveform store");
            }
        }
        r.LSu = LSl;
        r.LSx = LSi;
        r.n_lanes = static_cast<std::int32_t>(lanes);
        r.ppi = ppi;
        r.cplx = pr.cplx ? 1 : 0;
        r.nonlinear = nonlinear ? 1 : 0;
        r.max_iter = 64;  // circuit.h:898
        r.n_steps = n_steps;
        r.time_stepping = time_stepping ? 1 : 0;
        r.t0 = t0;
        r.dt = dt;
        auto const& env{parent->env};
        // default tolerances of circult::solve() (circuit.h:900-903)
        r.guard = pivot_guard;
        r.v_abstol = env.V_eps_max > 0.0 ? env.V_eps_max : 1e-6;
        r.v_reltol = env.V_epsr_max > 0.0 ? env.V_epsr_max : 1e-3;
        r.i_abstol = env.I_eps_max > 0.0 ? env.I_eps_max : 1e-12;
        r.i_reltol = env.I_epsr_max > 0.0 ? env.I_epsr_max : r.v_reltol;
        last_LSl = LSl;
        last_lanes = lanes;
        last_cplx = pr.cplx;
        if(pe_b200_launch(&r, stream) != 0) { return dev_fail(error, "launch"); }
        return true;
    }

    bool batch::run_phase_resident(prog_mode m, bool with_prep, bool nonlinear, int n_steps, bool time_stepping, double t0, double dt, std::size_t lanes, int ppi)
    {
        int const mi{static_cast<int>(m)};
        auto& pr{cc->prog[static_cast<std::size_t>(mi)]};
        int I{}, J{};
        if(!pick_geometry(pr, I, J))
        {
            error = "resident program does not fit the shared memory of one CTA";
            set_last_error(error);
            return false;
        }
        int const ig{(stream_mode && pr.rS == 1) ? 32 : I / J};  // the stream kernel reads its interpreted sections one stream per warp
        auto& d_w{d_words[static_cast<std::size_t>(mi)]};
        auto& d_so{d_secoff[static_cast<std::size_t>(mi)]};
        auto& d_i{d_io[static_cast<std::size_t>(mi)]};
        if(!uploaded[static_cast<std::size_t>(mi)] || uploaded_ig[static_cast<std::size_t>(mi)] != ig)
        {
            pr.pack(ig);
            if(!d_w.ensure(pr.words.size() * 4) || !d_so.ensure(pr.sec_off.size() * 4) || !d_i.ensure(std::max<std::size_t>(pr.io.size(), 1) * sizeof(pe_b200_io)))
            {
                return dev_fail(error, "alloc program");
            }
            if(pe_b200_dev_h2d(d_w.p, pr.words.data(), pr.words.size() * 4, stream) != 0 ||
               pe_b200_dev_h2d(d_so.p, pr.sec_off.data(), pr.sec_off.size() * 4, stream) != 0 ||
               (!pr.io.empty() && pe_b200_dev_h2d(d_i.p, pr.io.data(), pr.io.size() * sizeof(pe_b200_io), stream) != 0))
            {
                return dev_fail(error, "upload program");
            }
            uploaded[static_cast<std::size_t>(mi)] = true;
            uploaded_ig[static_cast<std::size_t>(mi)] = ig;
        }
        if(cc->dt_slot >= 0) { cc->cst[static_cast<std::size_t>(cc->dt_slot)] = dt; }
        if(!d_cst.ensure(cc->cst.size() * sizeof(double))) { return dev_fail(error, "alloc const table"); }
        if(pe_b200_dev_h2d(d_cst.p, cc->cst.data(), cc->cst.size() * sizeof(double), stream) != 0) { return dev_fail(error, "upload const table"); }

        std::int64_t const LSl{pr.cplx ? round_up32(lanes) : LSi};
        if(pr.cplx)
        {
            std::size_t const wl_bytes{static_cast<std::size_t>(std::max(pr.n_lane_slots, 1)) * static_cast<std::size_t>(LSl) * sizeof(double)};
            if(!d_wl.ensure(wl_bytes)) { return dev_fail(error, "alloc lane workspace"); }
        }
        if(!d_status.ensure(static_cast<std::size_t>(LSl) * 4) || !d_solves.ensure(static_cast<std::size_t>(LSl) * 4)) { return dev_fail(error, "alloc status"); }

        pe_b200_rrun r{};
        r.words = static_cast<std::uint32_t const*>(d_w.p);
        r.sec_off = static_cast<std::uint32_t const*>(d_so.p);
        r.io = static_cast<pe_b200_io const*>(d_i.p);
        r.n_io = static_cast<std::int32_t>(pr.io.size());
        r.has_prep = (with_prep && pr.has_sec[0]) ? 1 : 0;
        r.has_step = pr.has_sec[1] ? 1 : 0;
        r.cst = static_cast<double const*>(d_cst.p);
        r.wu = static_cast<double*>(pr.cplx ? d_wl.p : d_wi.p);
        r.wx = static_cast<double*>(d_wi.p);
        r.status = static_cast<std::int32_t*>(d_status.p);
        r.solves = static_cast<std::uint32_t*>(d_solves.p);
        r.wave = nullptr;
        r.probes = nullptr;
        r.n_probe = 0;
        if(time_stepping && !probes.empty())
        {
            std::vector<std::uint32_t> po;
            for(int u: probes) { po.push_back(pr.x_slot[static_cast<std::size_t>(u)]); }
            if(!d_probes.ensure(po.size() * 4) || !d_wave.ensure(static_cast<std::size_t>(n_steps) * po.size() * static_cast<std::size_t>(LSl) * sizeof(double)))
            {
                return dev_fail(error, "alloc waveform store");
            }
            if(pe_b200_dev_h2d(d_probes.p, po.data(), po.size() * 4, stream) != 0) { return dev_fail(error, "upload probes"); }
            if(pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "sync"); }  // `po` is pageable host memory
            r.wave = static_cast<double*>(d_wave.p);
            r.probes = static_cast<std::uint32_t const*>(d_probes.p);
            r.n_probe = static_cast<std::int32_t>(po.size());
            wave_steps = static_cast<std::size_t>(n_steps);
            wave_pitch = LSl;
            // rows of lanes that fail and of the steps after a failure are never written: they read as zero, not as whatever
            // the allocation held
            if(pe_b200_dev_memset0(d_wave.p, static_cast<std::size_t>(n_steps) * po.size() * static_cast<std::size_t>(LSl) * sizeof(double), stream) != 0)
            {
                return dev_fail(error, "zero waveform store");
            }
        }
        r.LSu = LSl;
        r.LSx = LSi;
        r.n_lanes = static_cast<std::int32_t>(lanes);
        r.ppi = ppi;
        r.S = pr.rS;
        r.I = I;
        r.J = J;
        r.n_slots = pr.r_slots;
        r.wsg = nullptr;
        r.LSw = 0;
        if(use_hbm(pr))
        {
            std::int64_t const LSw{static_cast<std::int64_t>((lanes + static_cast<std::size_t>(I) - 1) / static_cast<std::size_t>(I) * static_cast<std::size_t>(I))};
            if(!d_ws.ensure(static_cast<std::size_t>(pr.r_slots) * static_cast<std::size_t>(LSw) * sizeof(double))) { return dev_fail(error, "alloc HBM workspace"); }
            r.wsg = static_cast<double*>(d_ws.p);
            r.LSw = LSw;
            r.prefetch = res_prefetch;
            r.regs128 = pr.n_fused > 0 ? 1 : 0;
            // 128-lane groups whose sub-tree warps do not fit one CTA of 512 threads are run by a cluster of two CTAs
            r.cluster = (J == 4 && pr.rS * 32 > 512) ? 2 : 1;
            // dynamic (group, chunk) scheduling of long time loops: up to 32 chunks of at least 4 steps
            int nc{res_chunks > 0 ? res_chunks : std::min(32, n_steps / 4)};
            nc = std::clamp(nc, 1, 32);
            if(time_stepping && n_steps > 1 && nc > 1)
            {
                int const cs{(n_steps + nc - 1) / nc};
                nc = (n_steps + cs - 1) / cs;
                std::size_t const groups{(lanes + static_cast<std::size_t>(I) - 1) / static_cast<std::size_t>(I)};
                if(!d_sched.ensure((1 + groups) * sizeof(std::uint32_t))) { return dev_fail(error, "alloc scheduler"); }
                if(pe_b200_dev_memset0(d_sched.p, (1 + groups) * sizeof(std::uint32_t), stream) != 0) { return dev_fail(error, "zero scheduler"); }
                r.sched = static_cast<std::uint32_t*>(d_sched.p);
                r.chunk_steps = cs;
                r.n_chunks = nc;
                double tc{t0};
                for(int k{}; k < nc; ++k)
                {
                    r.t_chunk[k] = tc;
                    for(int q{}; q < cs; ++q) { tc = tc + dt; }  // the reference's accumulation (circuit.h:243-248)
                }
            }
        }
        r.cplx = pr.cplx ? 1 : 0;
        r.nonlinear = nonlinear ? 1 : 0;
        r.max_iter = 64;  // circuit.h:898
        r.n_steps = n_steps;
        r.time_stepping = time_stepping ? 1 : 0;
        r.t0 = t0;
        r.dt = dt;
        auto const& env{parent->env};
        r.guard = pivot_guard;
        r.v_abstol = env.V_eps_max > 0.0 ? env.V_eps_max : 1e-6;
        r.v_reltol = env.V_epsr_max > 0.0 ? env.V_epsr_max : 1e-3;
        r.i_abstol = env.I_eps_max > 0.0 ? env.I_eps_max : 1e-12;
        r.i_reltol = env.I_epsr_max > 0.0 ? env.I_epsr_max : r.v_reltol;
        last_LSl = LSl;
        last_lanes = lanes;
        last_cplx = pr.cplx;
        last_I = I;
        last_J = J;
        last_S = pr.rS;
        // Stream kernel: one-stream linear real programs (stream_mode).  The module is generated and compiled on first use
        // (seconds: the sweeps are rolled loops) and cached; when it cannot be had the interpreter runs the same program.
        if(stream_mode && r.wsg != nullptr && pr.rS == 1 && !nonlinear && !pr.cplx && pe_b200_stream_supported() != 0)
        {
            if(pr.stream_state == 0 || (pr.stream_state == 1 && pr.stream_j != I))
            {
                pr.stream_state = -1;
                stream_geom g{};
                std::string const src{pr.stream_laid_out ? stream_generate(pr, g) : std::string{}};  // laid out = accepted by stream_profitable()
                if(src.empty()) { pr.stream_error = "stream: the iter section holds ops the stream kernel does not cover"; }
                else if(stream_compile(src, J, I / J, pr.stream_blob, pr.stream_key, pr.stream_error))
                {
                    pr.stream_state = 1;
                    pr.stream_j = I;
                    pr.stream_tiles = g.n_tiles;
                    pr.stream_stage_rows = g.stage_rows;
                }
            }
            if(pr.stream_state == 1)
            {
                r.sched = nullptr;
                r.n_chunks = 1;
                if(pe_b200_launch_stream(&r, pr.stream_blob.data(), pr.stream_blob.size(), pr.stream_key, pr.stream_tiles, pr.stream_stage_rows, stream) == 0)
                {
                    last_jit = 2;
                    return true;
                }
                pr.stream_state = -1;
                pr.stream_error = std::string{"stream: "} + pe_b200_dev_last_error();
            }
            if(res_stream == 1)
            {
                error = pr.stream_error;
                set_last_error(error);
                return false;
            }
        }
        // Specialised kernel: linear real programs on 128-lane groups whose iter section is DOT / CAP_STEP only.  Required
        // (res_jit = 1): built on the spot if it is not in the cache (minutes for a 1000-node circuit), a failure is an
        // error; automatic: large batches take it when the cache holds it, the word interpreter runs otherwise.
        bool const jit_auto{res_jit == 0 && lanes >= 4096 && std::getenv("PE_B200_NO_JIT") == nullptr};
        if(r.wsg != nullptr && J == 4 && !nonlinear && !pr.cplx && (res_jit == 1 || jit_auto) && pe_b200_jit_supported() != 0)
        {
            if(pr.jit_state == 0 || (pr.jit_state == 1 && pr.jit_cl != r.cluster))
            {
                pr.jit_state = -1;
                if(jit_supported(pr))
                {
                    pr.jit_cl = r.cluster;
                    if(jit_compile(jit_generate(pr, jit_load_distance()), r.cluster, pr.jit_cubin, pr.jit_key, pr.jit_error, res_jit == 1)) { pr.jit_state = 1; }
                }
                else
                {
                    pr.jit_error = "jit: the iter section holds ops the specialised kernel does not cover";
                }
            }
            if(pr.jit_state == 1)
            {
                if(pe_b200_launch_jit(&r, pr.jit_cubin.data(), pr.jit_cubin.size(), pr.jit_key, stream) == 0)
                {
                    last_jit = 1;
                    return true;
                }
                // a cached module that does not load or launch (stale, wrong architecture, corrupt): required -> error,
                // automatic -> the interpreting kernel runs the program
                pr.jit_state = -1;
                pr.jit_error = std::string{"jit: "} + pe_b200_dev_last_error();
                if(res_jit == 1) { return dev_fail(error, "launch (specialised kernel)"); }
            }
            if(res_jit == 1)
            {
                error = pr.jit_error;
                set_last_error(error);
                return false;
            }
        }
        last_jit = 0;
        if(pe_b200_launch_resident(&r, stream) != 0) { return dev_fail(error, "launch"); }
        return true;
    }

    // one analyze() of this batch's own program; st / sv = per-lane status and solve counters (empty when nothing was launched)
    bool batch::analyze_main(std::vector<std::int32_t>& st, std::vector<std::uint32_t>& sv)
    {
        st.clear();
        sv.clear();
        last_points_hint = 1;
        // one huge linear DC circuit (config A): the parallelism is in the elimination DAG, not in static per-lane programs
        if(frontal_applicable(*parent, static_cast<std::size_t>(make_numbering(parent->nl).unknowns())))
        {
            if(!frontal) { frontal = std::shared_ptr<frontal_state>(frontal_new(), frontal_delete); }
            return frontal_run(*this, *frontal);
        }
        if(frontal)
        {
            frontal.reset();
            cc.reset();  // the program-less record of the other path
        }
        if(!ensure_compiled() || !upload_sweeps()) { return false; }
        auto const at{parent->at};
        int const n{cc->num.unknowns()};
        if(n == 0) { return true; }  // circuit.h:1116-1120: nothing to solve
        bool const nonlin{parent->nl.has_nonlinear()};
        // status / solve counters start clean for every analyze()
        if(pe_b200_dev_memset0(d_status.p, d_status.bytes, stream) != 0 || pe_b200_dev_memset0(d_solves.p, d_solves.bytes, stream) != 0)
        {
            return dev_fail(error, "zero status");
        }
        last_points = 1;

        auto run_tr = [&](bool with_prep) -> bool
        {
            double const dt{parent->tr.t_step};
            // step count with the reference's floating-point time accumulation (circuit.h:242-248)
            double const end_time{tr_duration + parent->tr.t_stop};
            double t{tr_duration};
            int steps{};
            while(t < end_time)
            {
                t = t + dt;
                ++steps;
            }
            if(steps == 0)
            {
                if(with_prep) { return run_phase(prog_mode::TR, true, nonlin, 0, false, tr_duration, dt, n_inst, 1); }
                return true;
            }
            if(!run_phase(prog_mode::TR, with_prep, nonlin, steps, true, tr_duration, dt, n_inst, 1)) { return false; }
            tr_duration = t;
            last_step = dt;
            return true;
        };

        auto run_ac = [&]() -> bool
        {
            auto const& a{(ac.points > 0 || ac.omega != 0.0) ? ac : parent->ac};
            auto same_sweep = [](ac_setting const& x, ac_setting const& y)
            { return x.sweep == y.sweep && x.omega == y.omega && x.omega_start == y.omega_start && x.omega_stop == y.omega_stop && x.points == y.points; };
            // an unchanged sweep keeps its table (host and device): building 1e6 points sequentially, replicating and uploading
            // them from pageable memory costs as much as a third of the solve itself
            bool const table_cached{lane_omegas.empty() && !ac_omegas.empty() && om_ptr != nullptr && same_sweep(om_key, a) && om_first == ac_slice_first && om_count == ac_slice_count && om_n_inst == n_inst};
            if(table_cached)
            {
                std::size_t const P{ac_omegas.size()};
                std::size_t const lanes{n_inst * P};
                std::int64_t const LSl{round_up32(lanes)};
                auto const& pr{cc->prog[static_cast<int>(prog_mode::AC)]};
                auto* const dst{static_cast<double*>(d_wl.p) + static_cast<std::int64_t>(pr.omega_slot) * LSl};
                if(d_wl.p != nullptr && static_cast<void*>(dst) == om_ptr && d_wl.bytes >= static_cast<std::size_t>(std::max(pr.n_lane_slots, 1)) * static_cast<std::size_t>(LSl) * sizeof(double) &&
                   d_status.bytes >= static_cast<std::size_t>(LSl) * 4 && d_solves.bytes >= static_cast<std::size_t>(LSl) * 4)
                {
                    if(pe_b200_dev_memset0(d_status.p, d_status.bytes, stream) != 0 || pe_b200_dev_memset0(d_solves.p, d_solves.bytes, stream) != 0)
                    {
                        return dev_fail(error, "zero status");
                    }
                    last_points = P;
                    last_points_hint = P;
                    return run_phase(prog_mode::AC, false, false, 1, false, 0.0, parent->tr.t_step, lanes, static_cast<int>(P));
                }
            }
            om_ptr = nullptr;
            ac_omegas.clear();
            if(!lane_omegas.empty())
            {
                // sub-batch of the pivot safety net: every "instance" is one flagged frequency point with its own omega
                if(lane_omegas.size() != n_inst)
                {
                    error = "per-lane omegas: size mismatch";
                    set_last_error(error);
                    return false;
                }
                std::size_t const lanes{n_inst};
                std::int64_t const LSl{round_up32(lanes)};
                auto const& pr{cc->prog[static_cast<int>(prog_mode::AC)]};
                if(!d_wl.ensure(static_cast<std::size_t>(std::max(pr.n_lane_slots, 1)) * static_cast<std::size_t>(LSl) * sizeof(double))) { return dev_fail(error, "alloc lane workspace"); }
                auto* dst{static_cast<double*>(d_wl.p) + static_cast<std::int64_t>(pr.omega_slot) * LSl};
                if(pe_b200_dev_h2d(dst, lane_omegas.data(), lanes * sizeof(double), stream) != 0 || pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "upload omega"); }
                if(!d_status.ensure(static_cast<std::size_t>(LSl) * 4) || !d_solves.ensure(static_cast<std::size_t>(LSl) * 4)) { return dev_fail(error, "alloc status"); }
                if(pe_b200_dev_memset0(d_status.p, d_status.bytes, stream) != 0 || pe_b200_dev_memset0(d_solves.p, d_solves.bytes, stream) != 0) { return dev_fail(error, "zero status"); }
                ac_omegas.push_back(lane_omegas[0]);
                last_points = 1;
                last_points_hint = 1;
                return run_phase(prog_mode::AC, false, false, 1, false, 0.0, parent->tr.t_step, lanes, 1);
            }
            if(a.sweep == sweep_type::single || a.points <= 1) { ac_omegas.push_back(a.omega); }
            else if(a.sweep == sweep_type::linear)
            {
                // circuit.h:399-410
                double const step{(a.omega_stop - a.omega_start) / static_cast<double>(a.points - 1)};
                for(std::size_t i{}; i < a.points; ++i) { ac_omegas.push_back(a.omega_start + step * static_cast<double>(i)); }
            }
            else
            {
                // circuit.h:412-428: cumulative product (the drift is part of the contract)
                if(a.omega_start <= 0.0 || a.omega_stop <= 0.0)
                {
                    error = "log sweep needs positive omega_start/omega_stop";
                    set_last_error(error);
                    return false;
                }
                double const ratio{std::pow(a.omega_stop / a.omega_start, 1.0 / static_cast<double>(a.points - 1))};
                double om{a.omega_start};
                for(std::size_t i{}; i < a.points; ++i)
                {
                    ac_omegas.push_back(om);
                    om *= ratio;
                }
            }
            if(ac_slice_count > 0)
            {
                // multi-GPU sharding of a sweep: the table is built sequentially in full, exactly like the reference's
                // cumulative product, and this rank keeps its contiguous block of points
                std::size_t const lo{std::min(ac_slice_first, ac_omegas.size())}, hi{std::min(ac_omegas.size(), lo + ac_slice_count)};
                ac_omegas = std::vector<double>(ac_omegas.begin() + static_cast<std::ptrdiff_t>(lo), ac_omegas.begin() + static_cast<std::ptrdiff_t>(hi));
                if(ac_omegas.empty()) { return true; }
            }
            std::size_t const P{ac_omegas.size()};
            std::size_t const lanes{n_inst * P};
            std::int64_t const LSl{round_up32(lanes)};
            auto const& pr{cc->prog[static_cast<int>(prog_mode::AC)]};
            if(!d_wl.ensure(static_cast<std::size_t>(std::max(pr.n_lane_slots, 1)) * static_cast<std::size_t>(LSl) * sizeof(double)))
            {
                return dev_fail(error, "alloc lane workspace");
            }
            std::vector<double> om(lanes);
            for(std::size_t i{}; i < n_inst; ++i) { std::copy(ac_omegas.begin(), ac_omegas.end(), om.begin() + static_cast<std::ptrdiff_t>(i * P)); }
            auto* dst{static_cast<double*>(d_wl.p) + static_cast<std::int64_t>(pr.omega_slot) * LSl};
            if(pe_b200_dev_h2d(dst, om.data(), lanes * sizeof(double), stream) != 0) { return dev_fail(error, "upload omega"); }
            if(pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "sync"); }  // `om` is pageable host memory
            om_key = a;
            om_first = ac_slice_first;
            om_count = ac_slice_count;
            om_n_inst = n_inst;
            om_ptr = dst;
            // the AC phase re-uses status/solves with `lanes` entries
            if(!d_status.ensure(static_cast<std::size_t>(LSl) * 4) || !d_solves.ensure(static_cast<std::size_t>(LSl) * 4)) { return dev_fail(error, "alloc status"); }
            if(pe_b200_dev_memset0(d_status.p, d_status.bytes, stream) != 0 || pe_b200_dev_memset0(d_solves.p, d_solves.bytes, stream) != 0)
            {
                return dev_fail(error, "zero status");
            }
            last_points = P;
            last_points_hint = P;
            return run_phase(prog_mode::AC, false, false, 1, false, 0.0, parent->tr.t_step, lanes, static_cast<int>(P));
        };

        bool ok{true};
        std::uint64_t op_solves{};
        bool op_failed{false};
        switch(at)
        {
            case analyze_type::OP:
            case analyze_type::DC: ok = run_phase(prog_mode::DC, true, nonlin, 1, false, tr_duration, parent->tr.t_step, n_inst, 1); break;
            case analyze_type::TR:
            {
                if(parent->tr.t_step <= 0.0)
                {
                    error = "TR needs t_step > 0";
                    set_last_error(error);
                    return false;
                }
                ok = run_tr(true);
                break;
            }
            case analyze_type::TROP:
            {
                if(parent->tr.t_step <= 0.0)
                {
                    error = "TROP needs t_step > 0";
                    set_last_error(error);
                    return false;
                }
                // 1) operating point with C open / L short, sources at t = 0 (circuit.h:264-266; base.h:289-292)
                ok = run_phase(prog_mode::TROP, true, nonlin, 1, false, 0.0, parent->tr.t_step, n_inst, 1);
                // 2) transient from there (lanes that failed stay flagged and are skipped)
                if(ok) { ok = run_tr(false); }
                break;
            }
            case analyze_type::AC:
            case analyze_type::ACOP:
            {
                // bias point first when needed (circuit.h:199-209, 213-224), then the sweep
                bool const need_op{at == analyze_type::ACOP || nonlin};
                if(need_op) { ok = run_phase(prog_mode::DC, true, nonlin, 1, false, tr_duration, parent->tr.t_step, n_inst, 1); }
                else
                {
                    ok = run_phase(prog_mode::DC, true, false, 0, false, tr_duration, parent->tr.t_step, n_inst, 1);  // prep only
                }
                if(ok && need_op)
                {
                    std::vector<std::int32_t> st(n_inst);
                    std::vector<std::uint32_t> sv(n_inst);
                    if(pe_b200_dev_d2h(st.data(), d_status.p, n_inst * 4, stream) != 0 || pe_b200_dev_d2h(sv.data(), d_solves.p, n_inst * 4, stream) != 0 ||
                       pe_b200_dev_sync(stream) != 0)
                    {
                        return dev_fail(error, "download bias status");
                    }
                    for(std::size_t i{}; i < n_inst; ++i)
                    {
                        op_solves += sv[i];
                        if(st[i] != PE_ST_OK) { op_failed = true; }
                    }
                    if(op_failed)
                    {
                        error = "AC: bias-point solve failed on at least one instance";
                        set_last_error(error);
                        total_solves = op_solves;
                        return false;
                    }
                }
                if(ok) { ok = run_ac(); }
                break;
            }
            default: ok = false; break;
        }
        if(!ok) { return false; }

        // one host sync per analyze(): three counters reduced on the device; the per-lane arrays travel only when a lane failed
        // (a sweep of a million points would read back 8 MB to learn that nothing happened)
        unsigned long long red[3]{};
        if(!d_red.ensure(sizeof(red)) ||
           pe_b200_status_reduce(static_cast<std::int32_t const*>(d_status.p), static_cast<std::uint32_t const*>(d_solves.p), static_cast<std::int64_t>(last_lanes),
                                 static_cast<unsigned long long*>(d_red.p), stream) != 0 ||
           pe_b200_dev_d2h(red, d_red.p, sizeof(red), stream) != 0 || pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "reduce status");
        }
        total_solves = op_solves;
        st.assign(last_lanes, 0);
        last_fast_rescue = false;
        if(red[0] == 0 && rescues.empty())
        {
            sv.clear();  // fetched on demand (get_solves)
            total_solves += red[2];
            return true;
        }
        // a single-shot analysis repeated on unchanged inputs flags the same lanes again: when the device counts say so (nothing
        // but guard trips, as many as the sub-batches own) the 8 bytes per lane stay on the device here too
        if(!rescues.empty() && rescue_fast_valid && red[0] == red[1] && red[1] == rescue_main_singular && at != analyze_type::TR && at != analyze_type::TROP)
        {
            sv.clear();
            total_solves += red[2] - rescue_main_owned_solves;
            last_fast_rescue = true;
            return true;
        }
        sv.assign(last_lanes, 0);
        if(pe_b200_dev_d2h(st.data(), d_status.p, last_lanes * 4, stream) != 0 || pe_b200_dev_d2h(sv.data(), d_solves.p, last_lanes * 4, stream) != 0 ||
           pe_b200_dev_sync(stream) != 0)
        {
            st.clear();
            sv.clear();
            return dev_fail(error, "download status");
        }
        return true;
    }

    // The pivot safety net (pe_host.hpp): lanes the guard flagged are solved again in sub-batches ordered on their own values.
    // The unit is the LANE: an instance of a real-valued analysis, one (instance, frequency point) of an AC sweep -- a sweep of a
    // million points with a few bad ones re-runs those few, each as a one-point "instance" of the sub-batch with its own omega.
    bool batch::run_rescues(std::vector<std::int32_t>& st, std::vector<std::uint32_t>& sv, bool fresh_state)
    {
        std::size_t const P{st.size() / n_inst};  // lanes per instance (points of an AC sweep, else 1)
        if(P == 0 || P * n_inst != st.size()) { return true; }
        std::vector<char> taken(st.size(), 0);
        std::uint64_t main_singular{}, main_singular_solves{};
        for(std::size_t l{}; l < st.size(); ++l)
        {
            if(st[l] == PE_ST_SINGULAR)
            {
                ++main_singular;
                main_singular_solves += sv[l];
            }
        }
        rescue_fast_valid = false;
        auto record = [&]()
        {
            // the fast path of analyze_main() applies when every lane the main batch flags is owned by a sub-batch
            std::uint64_t owned{};
            for(auto const& rs: rescues)
            {
                for(char const o: rs.owned) { owned += o ? 1u : 0u; }
            }
            rescue_main_singular = main_singular;
            rescue_main_owned_solves = main_singular_solves;
            rescue_fast_valid = owned == main_singular;
        };
        auto merge = [&](rescue_set& rs)
        {
            auto& sb{*rs.b};
            if(sb.last_status.size() != rs.inst.size()) { return; }
            if(sb.last_solves.size() != sb.last_status.size())
            {
                // every lane of the sub-batch was fine: its counters were reduced on the device, the per-lane ones are fetched here
                std::vector<std::uint32_t> fetched(sb.last_status.size(), 0);
                if(!sb.get_solves(fetched.data())) { return; }
                sb.last_solves = std::move(fetched);
            }
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                if(!rs.owned[k] || rs.inst[k] >= st.size()) { continue; }
                taken[rs.inst[k]] = 1;
                st[rs.inst[k]] = sb.last_status[k];
                sv[rs.inst[k]] = sb.last_solves[k];
            }
        };
        // lanes that moved to a sub-batch in an earlier call continue there (a transient keeps its state in the sub-batch)
        for(auto& rs: rescues)
        {
            rs.b->stream = stream;
            (void)rs.b->analyze();
            ++stat_rescue_launches;
            if(rs.b->last_status.empty())
            {
                error = "rescue batch: " + rs.b->error;
                set_last_error(error);
                return false;
            }
            merge(rs);
        }
        std::vector<std::size_t> F;
        for(std::size_t l{}; l < st.size(); ++l)
        {
            if(!taken[l] && st[l] == PE_ST_SINGULAR) { F.push_back(l); }
        }
        if(F.empty())
        {
            record();
            return true;
        }
        stat_guard_trips += F.size();
        auto const at{parent->at};
        if((at == analyze_type::TR || at == analyze_type::TROP) && !fresh_state) { return true; }  // the state the transient continues from is gone
        bool const nonlin{parent->nl.has_nonlinear()};
        static double const nl_table[3]{1.0, 1e-3, 1e-6};
        // rows that live on the device only (circuit_batch_set_params fast path) are fetched once
        std::map<sweep_key, std::vector<double>> fetched_rows;
        for(auto const& [key, v]: sweeps)
        {
            if(v.size() >= n_inst) { continue; }
            auto const it{cc->swept_slot.find(key)};
            if(it == cc->swept_slot.end()) { continue; }
            auto& row{fetched_rows[key]};
            row.resize(n_inst);
            if(pe_b200_dev_d2h(row.data(), static_cast<double const*>(d_wi.p) + static_cast<std::int64_t>(it->second) * LSi, n_inst * sizeof(double), stream) != 0 ||
               pe_b200_dev_sync(stream) != 0)
            {
                return dev_fail(error, "download parameter row");
            }
        }
        for(int round{1}; round <= rescue_rounds && !F.empty(); ++round)
        {
            bool const last{round == rescue_rounds};
            rescue_set rs;
            rs.inst = F;
            rs.b = std::make_unique<batch>();
            auto& sb{*rs.b};
            sb.parent = parent;
            sb.n_inst = F.size();
            sb.device = device;
            sb.stream = stream;
            sb.is_rescue = true;
            sb.pivot_guard = last ? 0.0 : pivot_guard;
            sb.guard_all = guard_all;
            sb.nl_nominal = (nonlin && !last) ? nl_table[(round - 1) % 3] : nl_nominal;
            sb.probes = probes;
            if(last_cplx && !ac_omegas.empty())
            {
                // one frequency point per unit; the order is chosen at the first flagged point's omega
                sb.lane_omegas.resize(F.size());
                for(std::size_t k{}; k < F.size(); ++k) { sb.lane_omegas[k] = ac_omegas[F[k] % P]; }
                sb.omega0_override = std::fabs(sb.lane_omegas[0]);
                sb.ac.sweep = sweep_type::single;
                sb.ac.omega = sb.lane_omegas[0];
                sb.ac.points = 1;
            }
            for(auto const& [key, v]: sweeps)
            {
                auto& dst{sb.sweeps[key]};
                dst.resize(F.size());
                auto const fr{fetched_rows.find(key)};
                auto const& src{fr != fetched_rows.end() ? fr->second : v};
                if(src.size() >= n_inst)
                {
                    for(std::size_t k{}; k < F.size(); ++k) { dst[k] = src[F[k] / P]; }
                }
                else
                {
                    std::fill(dst.begin(), dst.end(), src.empty() ? 0.0 : src[0]);
                }
            }
            (void)sb.analyze();
            ++stat_rescue_launches;
            if(sb.last_status.size() != F.size())
            {
                error = "rescue batch: " + sb.error;
                set_last_error(error);
                return false;
            }
            rs.owned.assign(F.size(), 0);
            std::vector<std::size_t> next;
            for(std::size_t k{}; k < F.size(); ++k)
            {
                if(last || sb.last_status[k] != PE_ST_SINGULAR)
                {
                    rs.owned[k] = 1;
                    ++(last ? stat_unguarded : stat_rescued);
                }
                else
                {
                    next.push_back(F[k]);
                }
            }
            merge(rs);
            rescues.push_back(std::move(rs));
            F = std::move(next);
        }
        rescues_sweeps_rev = sweeps_rev;
        rescues_param_rev = parent->param_rev;
        rescues_structure_rev = parent->structure_rev;
        rescues_at = parent->at;
        rescues_ac = (ac.points > 0 || ac.omega != 0.0) ? ac : parent->ac;
        rescues_slice_first = ac_slice_first;
        rescues_slice_count = ac_slice_count;
        record();
        return true;
    }

    bool batch::analyze()
    {
        error.clear();
        if(parent == nullptr || n_inst == 0)
        {
            error = "empty batch";
            set_last_error(error);
            return false;
        }
        // sub-batches of an earlier call own their lanes as long as nothing they were built from has changed
        auto const& ac_now{(ac.points > 0 || ac.omega != 0.0) ? ac : parent->ac};
        auto same_ac = [](ac_setting const& x, ac_setting const& y)
        { return x.sweep == y.sweep && x.omega == y.omega && x.omega_start == y.omega_start && x.omega_stop == y.omega_stop && x.points == y.points; };
        if(!rescues.empty() && (rescues_sweeps_rev != sweeps_rev || rescues_param_rev != parent->param_rev || rescues_structure_rev != parent->structure_rev || rescues_at != parent->at ||
                                !same_ac(rescues_ac, ac_now) || rescues_slice_first != ac_slice_first || rescues_slice_count != ac_slice_count))
        {
            rescues.clear();  // their lanes are numbered by the analysis (points of THIS sweep) they were built in
        }
        bool const fresh_state{tr_duration == 0.0};
        std::vector<std::int32_t> st;
        std::vector<std::uint32_t> sv;
        bool const launched{analyze_main(st, sv)};
        if(!launched && st.empty())
        {
            last_status.clear();
            return false;
        }
        if(frontal || st.empty())
        {
            last_status = st;
            last_solves = sv;
            return launched;
        }
        if(last_fast_rescue)
        {
            // the sub-batches run again, their lanes' status and counters replace the main batch's; every other lane is fine
            bool sub_ok{true};
            for(auto& rs: rescues)
            {
                rs.b->stream = stream;
                (void)rs.b->analyze();
                ++stat_rescue_launches;
                if(rs.b->last_status.size() != rs.inst.size())
                {
                    error = "rescue batch: " + rs.b->error;
                    set_last_error(error);
                    return false;
                }
                bool all_owned{true};
                for(char const o: rs.owned) { all_owned = all_owned && o != 0; }
                std::vector<std::uint32_t> ssv;
                if(!all_owned)
                {
                    ssv.assign(rs.inst.size(), 0);
                    if(!rs.b->get_solves(ssv.data())) { return false; }
                }
                else
                {
                    total_solves += rs.b->total_solves;
                }
                for(std::size_t k{}; k < rs.inst.size(); ++k)
                {
                    if(!rs.owned[k] || rs.inst[k] >= st.size()) { continue; }
                    st[rs.inst[k]] = rs.b->last_status[k];
                    sub_ok = sub_ok && rs.b->last_status[k] == PE_ST_OK;
                    if(!all_owned) { total_solves += ssv[k]; }
                }
            }
            last_status = std::move(st);
            last_solves.clear();  // per-lane counters on demand (get_solves merges the sub-batches' lanes)
            if(!sub_ok)
            {
                error = "analyze: at least one lane failed (no convergence or singular matrix)";
                set_last_error(error);
            }
            return sub_ok;
        }
        if(!sv.empty() && !is_rescue && pivot_guard > 0.0 && !run_rescues(st, sv, fresh_state)) { return false; }
        bool all_ok{true};
        for(std::size_t i{}; i < sv.size(); ++i)  // empty: every lane is fine and total_solves already holds the device's sum
        {
            total_solves += sv[i];
            if(st[i] != PE_ST_OK) { all_ok = false; }
        }
        last_status = std::move(st);
        last_solves = std::move(sv);
        if(!all_ok)
        {
            error = "analyze: at least one lane failed (no convergence or singular matrix)";
            set_last_error(error);
        }
        return all_ok;
    }

    bool batch::get_status(std::int32_t* st)
    {
        if(last_lanes == 0) { return false; }
        if(last_status.size() == last_lanes)
        {
            std::copy(last_status.begin(), last_status.end(), st);  // merged with the rescue sub-batches by analyze()
            return true;
        }
        if(pe_b200_dev_d2h(st, d_status.p, last_lanes * 4, stream) != 0 || pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "download status"); }
        return true;
    }

    bool batch::get_solves(std::uint32_t* sv)
    {
        if(last_lanes == 0) { return false; }
        if(last_solves.size() == last_lanes)
        {
            std::copy(last_solves.begin(), last_solves.end(), sv);
            return true;
        }
        if(pe_b200_dev_d2h(sv, d_solves.p, last_lanes * 4, stream) != 0 || pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "download solves"); }
        for(auto& rs: rescues)
        {
            std::vector<std::uint32_t> ssv(rs.inst.size(), 0);
            if(rs.b->last_lanes != rs.inst.size() || !rs.b->get_solves(ssv.data())) { continue; }
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                if(rs.owned[k] && rs.inst[k] < last_lanes) { sv[rs.inst[k]] = ssv[k]; }
            }
        }
        return true;
    }

    bool batch::get_solution(double* x)
    {
        if(!cc) { return false; }
        std::size_t const n{static_cast<std::size_t>(cc->num.unknowns())};
        if(n == 0) { return true; }
        // x lives in INST slots [0, n): rows of LSi doubles -> packed [n][n_inst], then transposed on the host
        std::vector<double> tmp(n * n_inst);
        if(pe_b200_dev_d2h_2d(tmp.data(), n_inst * sizeof(double), d_wi.p, static_cast<std::size_t>(LSi) * sizeof(double), n_inst * sizeof(double), n, stream) != 0 ||
           pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "download solution");
        }
        for(std::size_t j{}; j < n; ++j)
        {
            for(std::size_t i{}; i < n_inst; ++i) { x[i * n + j] = tmp[j * n_inst + i]; }
        }
        // rows of the lanes a rescue sub-batch owns (pivot safety net)
        for(auto const& rs: rescues)
        {
            if(!rs.b->lane_omegas.empty()) { continue; }  // units of an AC sub-batch are frequency points: see get_ac_solution
            std::vector<double> sx(rs.inst.size() * n);
            if(!rs.b->get_solution(sx.data())) { return false; }
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                if(rs.owned[k] && rs.inst[k] < n_inst) { std::copy(sx.begin() + static_cast<std::ptrdiff_t>(k * n), sx.begin() + static_cast<std::ptrdiff_t>((k + 1) * n), x + rs.inst[k] * n); }
            }
        }
        return true;
    }

    bool batch::get_solution_soa(double* x)
    {
        if(!cc) { return false; }
        std::size_t const n{static_cast<std::size_t>(cc->num.unknowns())};
        if(n == 0) { return true; }
        // device-native layout: unknown-major rows of n_inst doubles, no host transpose
        if(pe_b200_dev_d2h_2d(x, n_inst * sizeof(double), d_wi.p, static_cast<std::size_t>(LSi) * sizeof(double), n_inst * sizeof(double), n, stream) != 0 ||
           pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "download solution");
        }
        for(auto const& rs: rescues)
        {
            std::size_t const m{rs.inst.size()};
            if(!rs.b->lane_omegas.empty()) { continue; }
            std::vector<double> sx(m * n);
            if(!rs.b->get_solution_soa(sx.data())) { return false; }
            for(std::size_t k{}; k < m; ++k)
            {
                if(!rs.owned[k] || rs.inst[k] >= n_inst) { continue; }
                for(std::size_t j{}; j < n; ++j) { x[j * n_inst + rs.inst[k]] = sx[j * m + k]; }
            }
        }
        return true;
    }

    bool batch::get_ac_solution(double* x)
    {
        if(!cc || !last_cplx) { return false; }
        auto const& pr{cc->prog[static_cast<int>(prog_mode::AC)]};
        std::size_t const n{static_cast<std::size_t>(cc->num.unknowns())};
        std::size_t const lanes{last_lanes};
        if(n == 0) { return true; }
        // x_opnd slots are contiguous: [x0.re, x0.im, x1.re, ...] rows of LSl doubles
        std::size_t const slot0{PE_OPND_SLOT(pr.x_opnd[0])};
        std::vector<double> tmp(2 * n * lanes);
        auto const* src{static_cast<double const*>(d_wl.p) + static_cast<std::int64_t>(slot0) * last_LSl};
        if(pe_b200_dev_d2h_2d(tmp.data(), lanes * sizeof(double), src, static_cast<std::size_t>(last_LSl) * sizeof(double), lanes * sizeof(double), 2 * n, stream) != 0 ||
           pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "download AC solution");
        }
        for(std::size_t j{}; j < 2 * n; ++j)
        {
            for(std::size_t l{}; l < lanes; ++l) { x[l * 2 * n + j] = tmp[j * lanes + l]; }
        }
        for(auto const& rs: rescues)
        {
            if(!rs.b->last_cplx || rs.b->last_lanes != rs.inst.size()) { continue; }
            std::vector<double> sx(rs.b->last_lanes * 2 * n);
            if(!rs.b->get_ac_solution(sx.data())) { return false; }
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                if(rs.owned[k] && rs.inst[k] < lanes) { std::copy(sx.begin() + static_cast<std::ptrdiff_t>(k * 2 * n), sx.begin() + static_cast<std::ptrdiff_t>((k + 1) * 2 * n), x + rs.inst[k] * 2 * n); }
            }
        }
        return true;
    }

    // complex solution of selected lanes of the last AC analyze(): x[k][unknown][re, im] for lanes[k] (a sweep of a million points is
    // 3 GB per instance; a caller that wants a sample of it does not have to download all of it)
    bool batch::get_ac_solution_lanes(std::size_t const* lanes, std::size_t n_sel, double* x)
    {
        if(!cc || !last_cplx) { return false; }
        auto const& pr{cc->prog[static_cast<int>(prog_mode::AC)]};
        std::size_t const n{static_cast<std::size_t>(cc->num.unknowns())};
        if(n == 0 || n_sel == 0) { return true; }
        std::size_t const slot0{PE_OPND_SLOT(pr.x_opnd[0])};
        auto const* base{static_cast<double const*>(d_wl.p) + static_cast<std::int64_t>(slot0) * last_LSl};
        for(std::size_t k{}; k < n_sel; ++k)
        {
            if(lanes[k] >= last_lanes)
            {
                error = "ac_solution_lanes: lane out of range";
                set_last_error(error);
                return false;
            }
            // one column of the [2 n rows][lanes] block: 2 n elements, one per row
            if(pe_b200_dev_d2h_2d(x + k * 2 * n, sizeof(double), base + lanes[k], static_cast<std::size_t>(last_LSl) * sizeof(double), sizeof(double), 2 * n, stream) != 0)
            {
                return dev_fail(error, "download AC lanes");
            }
        }
        if(pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "download AC lanes"); }
        // lanes a sub-batch of the pivot safety net owns
        for(auto const& rs: rescues)
        {
            if(!rs.b->last_cplx || rs.b->last_lanes != rs.inst.size()) { continue; }
            std::vector<std::size_t> sub_lane, dst;
            for(std::size_t k{}; k < n_sel; ++k)
            {
                auto const it{std::lower_bound(rs.inst.begin(), rs.inst.end(), lanes[k])};
                if(it != rs.inst.end() && *it == lanes[k] && rs.owned[static_cast<std::size_t>(it - rs.inst.begin())])
                {
                    sub_lane.push_back(static_cast<std::size_t>(it - rs.inst.begin()));
                    dst.push_back(k);
                }
            }
            if(sub_lane.empty()) { continue; }
            std::vector<double> sx(sub_lane.size() * 2 * n);
            if(!rs.b->get_ac_solution_lanes(sub_lane.data(), sub_lane.size(), sx.data())) { return false; }
            for(std::size_t q{}; q < dst.size(); ++q) { std::copy(sx.begin() + static_cast<std::ptrdiff_t>(q * 2 * n), sx.begin() + static_cast<std::ptrdiff_t>((q + 1) * 2 * n), x + dst[q] * 2 * n); }
        }
        return true;
    }

    bool batch::digital_clk()
    {
        if(!cc || d_wi.p == nullptr)
        {
            error = "digital_clk: analyze first";
            set_last_error(error);
            return false;
        }
        // comparators in element order; pin unknowns: -1 ground / no analog node (reads 0 V like an unconnected node)
        std::vector<std::int32_t> ab;
        auto const& nl{parent->nl};
        for(auto const& e: nl.elems)
        {
            if(e.d->code != E_CMP) { continue; }
            for(int p{}; p < 2; ++p)
            {
                int const node{e.pin_node[p]};
                int const u{node < 0 ? -1 : cc->num.node_index[static_cast<std::size_t>(node)]};
                ab.push_back(u < 0 ? -1 : u);
            }
        }
        n_cmp = ab.size() / 2;
        if(n_cmp == 0) { return true; }
        if(!d_cmp_idx.ensure(ab.size() * 4) || !d_cmp_out.ensure(n_cmp * static_cast<std::size_t>(LSi))) { return dev_fail(error, "alloc comparator buffers"); }
        if(pe_b200_dev_h2d(d_cmp_idx.p, ab.data(), ab.size() * 4, stream) != 0 || pe_b200_dev_sync(stream) != 0) { return dev_fail(error, "upload comparator table"); }
        if(pe_b200_compare(static_cast<double const*>(d_wi.p), LSi, static_cast<std::int32_t>(n_inst), static_cast<std::int32_t const*>(d_cmp_idx.p),
                           static_cast<std::int32_t>(n_cmp), static_cast<std::uint8_t*>(d_cmp_out.p), stream) != 0)
        {
            return dev_fail(error, "comparator kernel");
        }
        return true;
    }

    bool batch::get_comparator_states(std::uint8_t* out)
    {
        if(n_cmp == 0) { return true; }
        if(d_cmp_out.p == nullptr) { return false; }
        std::vector<std::uint8_t> tmp(n_cmp * n_inst);
        if(pe_b200_dev_d2h_2d(tmp.data(), n_inst, d_cmp_out.p, static_cast<std::size_t>(LSi), n_inst, n_cmp, stream) != 0 || pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "download comparator states");
        }
        for(std::size_t c{}; c < n_cmp; ++c)
        {
            for(std::size_t i{}; i < n_inst; ++i) { out[i * n_cmp + c] = tmp[c * n_inst + i]; }
        }
        // instances a sub-batch of the pivot safety net owns: their comparators read THAT solution
        for(auto& rs: rescues)
        {
            if(!rs.b->lane_omegas.empty()) { continue; }
            std::vector<std::uint8_t> sub(rs.inst.size() * n_cmp);
            if(!rs.b->digital_clk() || rs.b->n_cmp != n_cmp || !rs.b->get_comparator_states(sub.data())) { return false; }
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                if(rs.owned[k] && rs.inst[k] < n_inst) { std::copy(sub.begin() + static_cast<std::ptrdiff_t>(k * n_cmp), sub.begin() + static_cast<std::ptrdiff_t>((k + 1) * n_cmp), out + rs.inst[k] * n_cmp); }
            }
        }
        return true;
    }

    bool batch::get_wave(double* w)
    {
        if(probes.empty() || wave_steps == 0 || d_wave.p == nullptr) { return false; }
        std::size_t const rows{wave_steps * probes.size()};
        if(pe_b200_dev_d2h_2d(w, n_inst * sizeof(double), d_wave.p, static_cast<std::size_t>(wave_pitch) * sizeof(double), n_inst * sizeof(double), rows, stream) != 0 ||
           pe_b200_dev_sync(stream) != 0)
        {
            return dev_fail(error, "download waveform");
        }
        for(auto const& rs: rescues)
        {
            std::size_t const m{rs.inst.size()};
            if(rs.b->wave_steps != wave_steps || rs.b->probes.size() != probes.size() || !rs.b->lane_omegas.empty()) { continue; }
            std::vector<double> sw(rows * m);
            if(!rs.b->get_wave(sw.data())) { continue; }
            for(std::size_t k{}; k < m; ++k)
            {
                if(!rs.owned[k]) { continue; }
                for(std::size_t q{}; q < rows; ++q) { w[q * n_inst + rs.inst[k]] = sw[q * m + k]; }
            }
        }
        return true;
    }

    // ---- checkpoint --------------------------------------------------------------------------------------------------------
    namespace
    {
        constexpr std::uint64_t k_ckpt_magic{0x3130544b43324250ull};  // "PB2CKT01"
        template <typename T>
        void put(std::vector<unsigned char>& o, T const& v)
        {
            auto const* p{reinterpret_cast<unsigned char const*>(&v)};
            o.insert(o.end(), p, p + sizeof(T));
        }
        template <typename T>
        bool get(unsigned char const*& p, unsigned char const* e, T& v)
        {
            if(static_cast<std::size_t>(e - p) < sizeof(T)) { return false; }
            std::memcpy(&v, p, sizeof(T));
            p += sizeof(T);
            return true;
        }
    }  // namespace

    // what a blob must agree on with the batch that loads it: the netlist (element codes, wiring), the per-instance parameter
    // keys (they decide the row layout) and the analysis step
    std::uint64_t batch::fingerprint() const
    {
        std::uint64_t h{1469598103934665603ull};
        auto mix = [&](std::uint64_t v)
        {
            for(int k{}; k < 8; ++k)
            {
                h ^= (v >> (8 * k)) & 0xffu;
                h *= 1099511628211ull;
            }
        };
        for(auto const& e: parent->nl.elems)
        {
            mix(static_cast<std::uint64_t>(e.d != nullptr ? e.d->code : -1));
            for(int const pn: e.pin_node) { mix(static_cast<std::uint64_t>(static_cast<std::int64_t>(pn))); }
        }
        for(auto const& [k, v]: sweeps)
        {
            mix(static_cast<std::uint64_t>(k.first));
            mix(static_cast<std::uint64_t>(k.second));
        }
        mix(static_cast<std::uint64_t>(n_inst));
        return h;
    }

    bool batch::save_state(std::vector<unsigned char>& out)
    {
        if(frontal || !cc || d_wi.p == nullptr)
        {
            error = "save_state: nothing to save (analyze first; the reduce-and-core path keeps no state)";
            set_last_error(error);
            return false;
        }
        std::size_t const rows{static_cast<std::size_t>(cc->n_inst_slots)};
        put(out, k_ckpt_magic);
        put(out, fingerprint());
        put(out, static_cast<std::uint64_t>(n_inst));
        put(out, static_cast<std::uint64_t>(rows));
        put(out, tr_duration);
        put(out, last_step);
        std::uint64_t n_sets{};
        for(auto const& rs: rescues) { n_sets += rs.b->lane_omegas.empty() ? 1u : 0u; }  // AC sub-batches (frequency points) keep no state
        put(out, n_sets);
        std::size_t const at{out.size()};
        out.resize(at + rows * n_inst * sizeof(double));
        if(rows != 0 && (pe_b200_dev_d2h_2d(out.data() + at, n_inst * sizeof(double), d_wi.p, static_cast<std::size_t>(LSi) * sizeof(double), n_inst * sizeof(double), rows, stream) != 0 ||
                         pe_b200_dev_sync(stream) != 0))
        {
            return dev_fail(error, "download state");
        }
        for(auto& rs: rescues)
        {
            if(!rs.b->lane_omegas.empty()) { continue; }
            put(out, static_cast<std::uint64_t>(rs.inst.size()));
            for(std::size_t k{}; k < rs.inst.size(); ++k)
            {
                put(out, static_cast<std::uint64_t>(rs.inst[k]));
                put(out, static_cast<std::uint8_t>(rs.owned[k]));
            }
            put(out, rs.b->pivot_guard);
            put(out, rs.b->nl_nominal);
            put(out, rs.b->omega0_override);
            if(!rs.b->save_state(out))
            {
                error = rs.b->error;
                return false;
            }
        }
        return true;
    }

    bool batch::load_state(unsigned char const* p, std::size_t n)
    {
        unsigned char const* const e{p + n};
        std::function<bool(batch&, unsigned char const*&)> load = [&](batch& b, unsigned char const*& q) -> bool
        {
            std::uint64_t magic{}, fp{}, ni{}, rows{}, nres{};
            double t{}, ls{};
            if(!get(q, e, magic) || !get(q, e, fp) || !get(q, e, ni) || !get(q, e, rows) || !get(q, e, t) || !get(q, e, ls) || !get(q, e, nres) || magic != k_ckpt_magic)
            {
                error = "load_state: not a batch checkpoint (or truncated)";
                return false;
            }
            std::size_t const bytes{static_cast<std::size_t>(rows) * static_cast<std::size_t>(ni) * sizeof(double)};
            if(static_cast<std::size_t>(e - q) < bytes)
            {
                error = "load_state: truncated";
                return false;
            }
            unsigned char const* const data{q};
            q += bytes;
            // the parameter rows of the blob are the authority: host copies first (they decide the elimination order), then the
            // program, then every persistent row as it was
            if(ni != b.n_inst)
            {
                error = "load_state: instance count differs";
                return false;
            }
            if(fp != b.fingerprint())
            {
                error = "load_state: the checkpoint belongs to another netlist / parameter set";
                return false;
            }
            if(!b.ensure_compiled())
            {
                error = b.error;
                return false;
            }
            if(rows != static_cast<std::uint64_t>(b.cc->n_inst_slots))
            {
                error = "load_state: row layout differs (another library version?)";
                return false;
            }
            auto const* rowsd{reinterpret_cast<double const*>(data)};
            for(auto const& [key, slot]: b.cc->swept_slot)
            {
                auto& v{b.sweeps[key]};
                v.assign(rowsd + static_cast<std::size_t>(slot) * ni, rowsd + (static_cast<std::size_t>(slot) + 1) * ni);
            }
            ++b.sweeps_rev;
            b.sweeps_dirty = false;
            if(rows != 0 && (pe_b200_dev_h2d_2d(b.d_wi.p, static_cast<std::size_t>(b.LSi) * sizeof(double), data, ni * sizeof(double), ni * sizeof(double), rows, b.stream) != 0 ||
                             pe_b200_dev_sync(b.stream) != 0))
            {
                error = std::string{"load_state: "} + pe_b200_dev_last_error();
                return false;
            }
            b.tr_duration = t;
            b.last_step = ls;
            b.rescues.clear();
            for(std::uint64_t r{}; r < nres; ++r)
            {
                std::uint64_t m{};
                if(!get(q, e, m) || m > ni)
                {
                    error = "load_state: truncated";
                    return false;
                }
                rescue_set rs;
                rs.inst.resize(m);
                rs.owned.resize(m);
                for(std::uint64_t k{}; k < m; ++k)
                {
                    std::uint64_t i{};
                    std::uint8_t o{};
                    if(!get(q, e, i) || !get(q, e, o) || i >= ni)
                    {
                        error = "load_state: truncated";
                        return false;
                    }
                    rs.inst[k] = static_cast<std::size_t>(i);
                    rs.owned[k] = static_cast<char>(o);
                }
                rs.b = std::make_unique<batch>();
                auto& sb{*rs.b};
                sb.parent = b.parent;
                sb.n_inst = static_cast<std::size_t>(m);
                sb.device = b.device;
                sb.stream = b.stream;
                sb.is_rescue = true;
                sb.guard_all = b.guard_all;
                sb.ac = b.ac;
                sb.ac_slice_first = b.ac_slice_first;
                sb.ac_slice_count = b.ac_slice_count;
                sb.probes = b.probes;
                if(!get(q, e, sb.pivot_guard) || !get(q, e, sb.nl_nominal) || !get(q, e, sb.omega0_override))
                {
                    error = "load_state: truncated";
                    return false;
                }
                // its elimination order was chosen on ITS first instance: the sweeps must be there before it compiles
                for(auto const& [key, v]: b.sweeps)
                {
                    auto& d{sb.sweeps[key]};
                    d.resize(m);
                    for(std::uint64_t k{}; k < m; ++k) { d[k] = v[rs.inst[k]]; }
                }
                if(!load(sb, q)) { return false; }
                b.rescues.push_back(std::move(rs));
            }
            b.rescues_sweeps_rev = b.sweeps_rev;
            b.rescues_param_rev = b.parent->param_rev;
            b.rescues_structure_rev = b.parent->structure_rev;
            b.rescues_at = b.parent->at;
            b.rescues_ac = (b.ac.points > 0 || b.ac.omega != 0.0) ? b.ac : b.parent->ac;
            b.rescues_slice_first = b.ac_slice_first;
            b.rescues_slice_count = b.ac_slice_count;
            return true;
        };
        unsigned char const* q{p};
        if(frontal || parent == nullptr)
        {
            error = "load_state: not available on this batch";
            set_last_error(error);
            return false;
        }
        if(!load(*this, q))
        {
            set_last_error(error);
            return false;
        }
        return true;
    }

    bool circuit::analyze()
    {
        if(!solo)
        {
            solo = std::make_unique<batch>();
            solo->parent = this;
            solo->n_inst = 1;
            auto const& d{default_path()};
            solo->res_S = d.res_S;
            solo->res_I = d.res_I;
            solo->res_J = d.res_J;
            solo->subtree_warps = d.subtree_warps;
            solo->res_ws = d.res_ws;
            solo->res_chunks = d.res_chunks;
            solo->res_prefetch = ((d.tuning & 1u) ? 1 : 0) | ((d.tuning & 2u) ? 2 : 0) | ((d.tuning & 4u) ? 0 : 4);
            solo->res_fuse = (d.tuning & 8u) ? 1 : 0;
            solo->res_jit = (d.tuning & 16u) ? 1 : ((d.tuning & 32u) ? -1 : 0);
            solo->res_stream = (d.tuning & 64u) ? 1 : ((d.tuning & 128u) ? -1 : 0);
        }
        solo->ac = {};
        bool const ok{solo->analyze()};
        if(!solo->cc) { return false; }
        num_host = solo->cc->num;
        std::size_t const n{static_cast<std::size_t>(num_host.unknowns())};
        x_host.assign(n, 0.0);
        xi_host.assign(n, 0.0);
        if(n == 0) { return ok; }
        if(solo->last_cplx && solo->last_lanes >= 1)
        {
            // node_t::an.voltage keeps the complex AC solution of the last point (circuit.h:1521)
            std::vector<double> all(solo->last_lanes * 2 * n);
            if(!solo->get_ac_solution(all.data())) { return false; }
            std::size_t const l{solo->last_lanes - 1};
            for(std::size_t j{}; j < n; ++j)
            {
                x_host[j] = all[l * 2 * n + 2 * j];
                xi_host[j] = all[l * 2 * n + 2 * j + 1];
            }
        }
        else
        {
            if(!solo->get_solution(x_host.data())) { return false; }
        }
        return ok;
    }
}  // namespace pe_b200
