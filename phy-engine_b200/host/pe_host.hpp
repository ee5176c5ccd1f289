// pe_host.hpp — C++23 host side of the B200 MNA hot path: netlist model, symbolic analysis ("compiler") and the
// batch driver.  Mirrors the reference's phy_engine::circult / netlist / analyze surface for the analog path
// (include/phy_engine/circuits/circuit.h, netlist/operation.h) but owns no numeric solve: every linear solve,
// Newton iteration, time step and frequency point runs in the sm_100a kernels behind pe_b200_program.h.
// There is no CPU solve path; without a CUDA device analyze() fails loudly.
#pragma once
#include <array>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../csrc/pe_b200_program.h"

namespace pe_b200
{
    // element codes = phy_engine_element_code (dll_api.h:51-135)
    enum : int
    {
        E_RES = 1,
        E_CAP = 2,
        E_IND = 3,
        E_VDC = 4,
        E_VAC = 5,
        E_IDC = 6,
        E_IAC = 7,
        E_VCCS = 8,
        E_VCVS = 9,
        E_CCCS = 10,
        E_CCVS = 11,
        E_SWITCH = 12,
        E_PN = 13,
        E_XFMR = 14,  // ideal transformer (transformer.h)
        E_KIND = 15,  // coupled inductors (coupled_inductors.h)
        E_XFMR_CT = 16,  // centre-tapped transformer (transformer_center_tap.h)
        E_OPAMP = 17,
        E_RELAY = 18,  // controller/relay.h
        E_CMP = 19,  // comparator: the analog -> digital boundary (controller/comparator.h)
        E_GEN_SAW = 20,  // generator/sawtooth.h
        E_GEN_SQUARE = 21,
        E_GEN_PULSE = 22,
        E_GEN_TRI = 23,
        E_NPN = 50,
        E_PNP = 51,
        E_NMOS = 52,
        E_PMOS = 53,
        E_BRIDGE = 54,
    };

    // analyze_type (circuits/analyze.h)
    enum class analyze_type : std::uint32_t
    {
        OP = 0,
        DC,
        AC,
        ACOP,
        TR,
        TROP
    };

    enum class sweep_type : int
    {
        single = 0,
        linear = 1,
        log = 2
    };

    constexpr int k_max_attr = 10;
    constexpr int k_max_pins = 5;

    struct elem_desc
    {
        int code;
        char const* name;
        int pins;
        int branches;
        int n_attr;
        char const* attr_name[k_max_attr];
        double attr_default[k_max_attr];
        int n_cabi_props;  // how many doubles create_circuit() consumes (dll_main.cpp:1707-1956)
        bool nonlinear;    // model_device_type::non_linear
        bool digital{false};  // model_device_type::digital: its pins are not analog pins (no node numbering, no stamps)
    };

    elem_desc const* find_desc(int code) noexcept;

    struct element
    {
        elem_desc const* d{};
        double attr[k_max_attr]{};  // internal units (VAC/IAC: omega rad/s, phase rad — VAC.h:29-50)
        int pin_node[k_max_pins]{-2, -2, -2, -2, -2};  // -2 unconnected, -1 ground, >= 0 node id (creation order)
    };

    // environment (circuits/environment/environment.h:7-22)
    struct environment
    {
        double V_eps_max{};
        double V_epsr_max{};
        double I_eps_max{};
        double I_epsr_max{};
        double charge_eps_max{};
        double g_min{};
        double r_open{};
        double t_TOEF{};
        double temperature{27.0};
        double norm_temperature{27.0};
    };

    struct ac_setting
    {
        sweep_type sweep{sweep_type::single};
        double omega{};
        double omega_start{};
        double omega_stop{};
        std::size_t points{};
    };

    struct tr_setting
    {
        double t_stop{};
        double t_step{};
    };

    struct netlist
    {
        std::vector<element> elems;
        int n_created_nodes{};

        // returns element index or -1 (unsupported code)
        int add_model(int code, double const* cabi_props, std::size_t* consumed);
        int create_node() { return n_created_nodes++; }
        bool add_to_node(int elem, int pin, int node);  // node -1 = ground
        // set_attribute semantics of the reference models (unit conversions included); false if idx invalid
        bool set_attribute(int elem, int idx, double v);
        int find_attribute(int elem, char const* name, std::size_t name_size) const;  // case-insensitive; -1 if absent
        bool has_nonlinear() const;
    };

    // node / branch numbering of circult::prepare() (circuit.h:481-540)
    struct numbering
    {
        std::vector<int> node_index;  // per created node: unknown index or -2 (no analog pin)
        std::vector<int> branch0;     // per element: first branch index (branch_counter order)
        int n_nodes{};
        int n_branches{};
        int unknowns() const { return n_nodes + n_branches; }
    };

    numbering make_numbering(netlist const& nl);

    enum class prog_mode : int
    {
        DC = 0,  // OP and DC share stamps for every in-scope model (base.h:266-281 fallback)
        TR = 1,
        TROP = 2,
        AC = 3,
        COUNT = 4
    };

    // ---- resident (shared-memory) form of a program: abstract per-stream op lists + slot allocation, packed into
    // warp vector-op words once the launch geometry (instance groups per CTA) is known.  See pe_b200_program.h.
    struct rop
    {
        std::uint32_t opcode{};
        std::uint32_t flags{};
        std::uint8_t bubble{};  // no op for this stream at this (globally aligned) position
        std::uint8_t first_only{};  // a time-step update folded into the iter section: runs in the first Newton iteration only
        std::uint32_t dst{}, scale{};                               // shared-memory slots
        // operand words are stream-relative: row * S + ((column - stream) mod S), | neg << 15 for sources
        std::vector<std::uint32_t> sre, sim;
        std::vector<std::pair<std::uint32_t, std::uint32_t>> pp;
        std::vector<std::uint32_t> opnd;  // value ops, encoding order
        std::vector<rop> sub;             // PE_OP_CROUT2: the DOTs of the fused elimination step (pivot first)
    };
    using rphase = std::vector<rop>;
    struct rstream
    {
        std::vector<rphase> sec[3];  // prep, step, iter; every stream holds the same number of phases per section
    };

    struct program
    {
        // resident form (resident == true: `words` hold vector ops, see pack())
        bool resident{};
        int rS{1};                          // streams per instance
        int r_slots{};                      // shared-memory slots per instance (a multiple of rS)
        std::uint32_t r_zero{};             // slot holding 0.0 (padding operand)
        std::vector<rstream> rstreams;      // [rS]
        std::vector<pe_b200_io> io;
        std::vector<std::uint32_t> x_slot;  // per unknown: shared-memory slot of the solution (cplx: re, im at + rS)
        int packed_ig{-1};                  // instance groups per CTA the words were packed for
        int n_warps{};
        std::vector<std::uint32_t> sec_off;  // [3][n_warps]
        bool has_sec[3]{};
        void pack(int ig);                   // (re)build words / sec_off for CTAs of rS * ig threads
        // stream kernel (host/stream.cpp): one-stream programs whose iter section is generated as tiled, TMA-fed code
        int stream_state{0};                 // 0 = not tried, 1 = module ready, -1 = not available
        bool stream_laid_out{};              // stream_prepare() has re-laid the workspace rows out (planes + replicas)
        int stream_j{0};                     // lanes per group (GL x J) the module was built for
        std::uint64_t stream_key{};
        std::vector<char> stream_blob;       // cubin (device) / path of the host module (emulator)
        std::string stream_error;
        std::uint32_t stream_tiles{}, stream_stage_rows{};
        // specialised kernel of the iter section (host/jit.cpp): 0 = not tried, 1 = cubin ready, -1 = not available
        int jit_state{0};
        int jit_cl{0};
        std::uint64_t jit_key{};
        std::vector<char> jit_cubin;
        std::string jit_error;

        bool built{};
        bool cplx{};
        bool structurally_singular{};
        int warps{1};                     // G: word streams per section = warps per CTA
        std::vector<std::uint32_t> words;  // word pool of this mode (prep + step + iter streams)
        pe_b200_section prep{}, step{}, iter{};
        int n_lane_slots{};
        std::vector<std::uint32_t> x_opnd;  // per unknown: where the solution lands (cplx: slot of re, im = +1)
        int omega_slot{-1};                 // U slot holding omega (AC)
        // statistics for the roofline and the schedule
        std::size_t nnz_a{}, nnz_lu{}, n_fma{};
        std::size_t n_guarded{};  // pivots that carry the run-time guard (PE_F_GUARD)
        std::size_t n_leaf_rows{}, n_top_rows{}, n_leaves{};
        std::size_t n_fused{};    // elimination steps emitted as one fused PE_OP_CROUT2
        std::size_t n_aliased{};  // U entries that are a signed copy of one stamped value and were never materialised
        std::size_t max_warp_words{};  // longest iter stream (critical path in words)
    };

    using sweep_key = std::pair<int, int>;  // (element, attribute)

    struct compiled
    {
        numbering num;
        int n_inst_slots{};
        int n_real_lane_slots{};  // max over the real-valued modes (their scratch shares the instance workspace)
        std::vector<double> cst;
        int dt_slot{-1};  // CONST slot holding the transient step (patched by the driver before every run)
        std::map<sweep_key, int> swept_slot;  // INST slot of each per-instance parameter
        std::array<program, static_cast<int>(prog_mode::COUNT)> prog;
        std::string error;
    };

    struct compile_input
    {
        netlist const* nl{};
        environment env{};
        double dt{};      // TR step (nominal values of companions)
        double omega0{};  // representative omega for AC pivot selection
        // per-instance parameters: lane-0 value (nominal) for each swept (element, attribute)
        std::map<sweep_key, double> swept_lane0;
        int warps_real{1};  // requested warps per CTA for the real-valued programs (DC/TR/TROP)
        int warps_ac{1};    // ... for the AC program
        int resident_real{0};  // > 0: build the real-valued programs in resident form with this many streams
        int resident_ac{0};    // > 0: ... the AC program
        bool fuse_steps{false};  // resident programs: emit small elimination steps as one fused op
        bool merge_step{true};   // resident programs: fold the per-time-step companion updates into the iter section
        bool guard_all{false};   // keep the pivot guard on every pivot (a non-positive R / C value was seen: no row is provably safe)
        double nl_nominal{0.0};  // > 0: conductance a not-yet-evaluated non-linear device enters the ordering with (default 1e-12 S)
    };

    // Symbolic phase: numbering, stamp maps, Markowitz/threshold pivot order on nominal values, fill pattern, slot
    // assignment and the batch program for every mode.  Pure host integer/graph work.
    std::unique_ptr<compiled> compile_circuit(compile_input const& in);

    // ---- batch driver -------------------------------------------------------------------------------------------
    struct device_buf
    {
        void* p{};
        std::size_t bytes{};
        device_buf() = default;
        device_buf(device_buf const&) = delete;
        device_buf& operator= (device_buf const&) = delete;
        ~device_buf();
        bool ensure(std::size_t n);  // (re)allocate if too small; contents undefined after growth
        void release();
    };

    struct circuit;

    // host/jit.cpp: run-time specialisation of the tree-streaming kernel
    bool jit_supported(program const& pr);
    std::string jit_generate(program const& pr, int load_distance);
    bool jit_compile(std::string const& gen, int CL, std::vector<char>& cubin, std::uint64_t& key, std::string& err, bool allow_compile = true);
    int jit_load_distance();

    // host/stream.cpp: the stream kernel (one warp per lane group, one stream per group, TMA-fed tiles)
    struct stream_geom
    {
        std::uint32_t n_tiles{}, stage_rows{};
        std::size_t n_loops{}, loop_ops{}, n_ops{};
        std::size_t rows_fetched{}, n_copies{}, rows_stored{};  // per solve
    };
    bool stream_supported(program const& pr);
    bool stream_profitable(program const& pr);  // ... and periodic enough for the generated code to be a few rolled loops
    // re-lays the workspace rows of a one-stream program out for the stream kernel (contiguous rows per tile, replica rows
    // of shared read-only operands); every executor of the program sees the same, value-equivalent program afterwards
    bool stream_prepare(program& pr);
    std::string stream_generate(program const& pr, stream_geom& g);
    bool stream_compile(std::string const& gen, int J, int GL, std::vector<char>& blob, std::uint64_t& key, std::string& err, bool allow_compile = true);

    // host/frontal.cpp: reduce-and-core path for one huge linear DC circuit per instance (config A)
    struct frontal_state;
    struct batch;
    bool frontal_applicable(circuit const& c, std::size_t n_unknowns);
    std::size_t& frontal_min_unknowns();
    frontal_state* frontal_new();
    void frontal_delete(frontal_state* s);
    bool frontal_run(batch& b, frontal_state& s);
    void frontal_stats(frontal_state const* s, std::int64_t* out11);  // unknowns, eliminated nodes, levels, core rows, edges, launches, ld, core edges, device us of reduce / LU / substitutions

    struct batch
    {
        circuit* parent{};
        std::shared_ptr<frontal_state> frontal;  // set once analyze() took the reduce-and-core path
        std::size_t n_inst{};
        std::int64_t LSi{};  // padded instance count
        int device{0};
        void* stream{};  // cudaStream_t chosen by the caller (0 = default stream)

        std::map<sweep_key, std::vector<double>> sweeps;  // host copies of per-instance parameters
        bool sweeps_dirty{true};
        bool layout_valid{};

        // batch AC sweep (lanes = n_inst * points)
        ac_setting ac{};
        std::size_t ac_slice_first{}, ac_slice_count{};  // count > 0: only this block of the sweep's points is solved (one rank's shard)

        std::unique_ptr<compiled> cc;
        std::vector<sweep_key> layout_keys;
        std::uint64_t cc_structure_rev{}, cc_param_rev{};
        double cc_dt{-1.0};
        bool device_stale{true};    // host program newer than what the device holds
        bool layout_pending{true};  // INST layout changed since the device workspace was sized

        device_buf d_wi, d_wl, d_cst, d_status, d_solves, d_wave, d_probes, d_red;
        std::array<device_buf, static_cast<int>(prog_mode::COUNT)> d_words;
        std::array<bool, static_cast<int>(prog_mode::COUNT)> uploaded{};
        int subtree_warps{0};  // 0 = choose from the lane count; else the requested G (power of two, <= PE_MAX_WARPS)
        int cc_warps_real{-1}, cc_warps_ac{-1};
        int pick_warps(std::size_t lanes, int n_unknowns) const;
        // resident (shared-memory) path: streams per instance (-1 = never, 0 = choose, else forced), instances per CTA
        // and per thread (0 = choose)
        int res_S{0}, res_I{0}, res_J{0};
        int res_ws{0};  // where the tree-scheduled program keeps its workspace: 0 = choose, 1 = shared memory, 2 = HBM
        device_buf d_ws;  // HBM workspace of the tree-streaming form [slots][lanes]
        device_buf d_sched;  // work-item counter + per-group chunk counters of the tree-streaming form
        // tree-streaming form, bit set: 1 = line-ahead L2 operand prefetch (measured neutral at J = 2: off), 2 = ... two lines
        // ahead, 4 = L1 re-fetch of every DOT result right after its store (+3 %: on)
        int res_prefetch{4};
        std::size_t last_points_hint{1};  // frequency points per instance of the AC sweep being launched (lane count = n_inst * points)
        int res_jit{0};      // specialised (run-time compiled) tree-streaming kernel: 0 = automatic, 1 = required, -1 = off
        int res_fuse{0};     // emit small elimination steps as one fused op (PE_OP_CROUT2); measured slower on config B (register file bound): off
        int cc_fuse{-1};
        int res_chunks{0};   // chunks the time loop is cut into for dynamic scheduling: 0 = choose, 1 = static (one CTA per group)
        bool use_hbm(program const& pr) const;
        int cc_res_real{-1}, cc_res_ac{-1};
        int pick_streams(int n_unknowns) const;
        bool pick_geometry(program const& pr, int& I, int& J) const;
        std::array<device_buf, static_cast<int>(prog_mode::COUNT)> d_secoff, d_io;
        std::array<int, static_cast<int>(prog_mode::COUNT)> uploaded_ig{};
        int res_stream{0};   // stream kernel: 0 = automatic, 1 = required, -1 = off
        bool stream_mode{};  // the real-valued programs were compiled with one stream per lane group for the stream kernel
        std::uint64_t stream_rejected_rev{};  // structure revision whose program the stream generator does not cover
        int last_jit{};  // 1 = the last tree-streaming launch ran the specialised (run-time compiled) kernel, 2 = the stream kernel, 3 = the reduce-and-core path
        int last_I{}, last_J{}, last_S{};  // geometry of the last resident launch (0 = the HBM-streaming kernel ran)

        // results of the last analyze()
        std::size_t last_lanes{};
        std::size_t last_points{1};
        bool last_cplx{};
        std::int64_t last_LSl{};
        std::vector<double> ac_omegas;  // per point
        // the omega table on the device is kept while the sweep, the slice, the instance count and the buffer stay the same
        ac_setting om_key{};
        std::size_t om_first{}, om_count{}, om_n_inst{};
        void* om_ptr{};
        std::uint64_t total_solves{};
        double tr_duration{};
        double last_step{};

        // waveform probes (unknown indices), recorded per time step when non-empty
        std::vector<int> probes;
        std::size_t wave_steps{};
        std::int64_t wave_pitch{};  // lane pitch of d_wave when it was recorded (a later AC analyze() changes last_LSl)

        std::string error;

        // ---- pivot safety net.  The elimination order is static (chosen on lane-0 values); pivots that are not provably safe
        // carry a run-time guard (PE_F_GUARD, pe_b200_program.h), a lane whose pivot loses more than -log2(pivot_guard) bits to
        // cancellation is flagged like a singular one, and analyze() re-runs the flagged lanes in a sub-batch whose order is
        // chosen on THEIR values (lane 0 of the sub-batch = the first flagged lane); what still fails goes to the next round,
        // the last round runs with the guard off (what the reference does: Eigen re-pivots per solve and has no such test,
        // SparseLU_pivotL.h:76-107).  The sub-batches stay: they own their lanes until a parameter changes (a transient continues
        // in them), and the download functions merge their rows in.
        double pivot_guard{PE_GUARD_DEFAULT};  // 0 = off (no guard, no rescue)
        int rescue_rounds{3};
        bool is_rescue{};
        bool guard_all{}, cc_guard_all{};  // compile with the guard on every pivot (set when a non-positive R / C value is seen)
        double nl_nominal{};     // > 0: ordering conductance of not-yet-evaluated non-linear devices (rescue rounds vary it)
        double omega0_override{};  // > 0: omega the AC order is chosen on
        std::vector<double> lane_omegas;  // sub-batch of flagged frequency points: one omega per "instance" (one point each)
        std::uint64_t sweeps_rev{1};  // bumped by every per-instance parameter write
        struct rescue_set
        {
            std::unique_ptr<batch> b;
            std::vector<std::size_t> inst;   // sub-batch instance k = LANE inst[k] of this batch (an instance; one point of an AC sweep)
            std::vector<char> owned;         // [k]: its results replace this batch's
        };
        std::vector<rescue_set> rescues;
        // what the main batch looked like when the sub-batches were built: lanes it flagged, the solve counters of those lanes
        std::uint64_t rescue_main_singular{}, rescue_main_owned_solves{};
        bool rescue_fast_valid{}, last_fast_rescue{};
        std::uint64_t rescues_sweeps_rev{}, rescues_param_rev{}, rescues_structure_rev{};
        analyze_type rescues_at{};
        ac_setting rescues_ac{};
        std::size_t rescues_slice_first{}, rescues_slice_count{};
        std::uint64_t stat_guard_trips{}, stat_rescued{}, stat_rescue_launches{}, stat_unguarded{};
        std::vector<std::int32_t> last_status;  // merged status / solve counters of the last analyze()
        std::vector<std::uint32_t> last_solves;
        bool analyze_main(std::vector<std::int32_t>& st, std::vector<std::uint32_t>& sv);
        bool run_rescues(std::vector<std::int32_t>& st, std::vector<std::uint32_t>& sv, bool fresh_state);

        // ---- checkpoint (the reference persists a single circuit in pe_nl_fileformat.h:584-657, 805-1046): the
        // persistent per-instance rows (solution, swept parameters, companion / device state), the clock of the transient and
        // the sub-batches of the pivot safety net, as one self-describing blob a batch of the same netlist can resume from
        bool save_state(std::vector<unsigned char>& out);
        bool load_state(unsigned char const* p, std::size_t n);
        std::uint64_t fingerprint() const;

        bool analyze();
        bool compile_host(bool& layout_change);  // symbolic phase only, no device needed
        bool ensure_compiled();
        bool upload_sweeps();
        bool run_phase(prog_mode m, bool with_prep, bool nonlinear, int n_steps, bool time_stepping, double t0, double dt, std::size_t lanes, int ppi);
        bool run_phase_resident(prog_mode m, bool with_prep, bool nonlinear, int n_steps, bool time_stepping, double t0, double dt, std::size_t lanes, int ppi);
        bool run_prep_only();
        // downloads
        bool get_solution(double* x /* [n_inst][n] */);
        bool get_solution_soa(double* x /* [n][n_inst] */);
        bool get_ac_solution(double* x /* [lanes][n][2] */);
        bool get_ac_solution_lanes(std::size_t const* lanes, std::size_t n_sel, double* x /* [n_sel][n][2] */);
        bool get_status(std::int32_t* st /* [lanes] */);
        bool get_solves(std::uint32_t* sv /* [lanes] */);
        bool get_wave(double* w /* [steps][probes][n_inst] */);
        // mixed-signal boundary, analog -> digital: update_digital_clk of every comparator of the netlist (vA >= vB,
        // controller/comparator.h:73-108) for every instance, on the device
        device_buf d_cmp_idx, d_cmp_out;
        std::size_t n_cmp{};
        bool digital_clk();
        bool get_comparator_states(std::uint8_t* out /* [n_inst][n_cmp] */);
    };

    struct circuit
    {
        environment env{};
        netlist nl{};
        analyze_type at{analyze_type::TR};
        ac_setting ac{};
        tr_setting tr{};
        std::uint64_t structure_rev{1};  // bumped when elements / connections change (device state is reset)
        std::uint64_t param_rev{1};      // bumped when a broadcast parameter / environment value changes

        std::unique_ptr<batch> solo;   // the n = 1 batch behind circuit_analyze()
        std::vector<double> x_host;    // last solution (re) of the solo batch, unknown order
        std::vector<double> xi_host;   // imaginary parts (AC)
        numbering num_host;
        std::vector<std::int8_t> digital_state;  // per created node: 0 / 1 after circuit_digital_clk drove it, 2 = indeterminate

        bool analyze();
    };

    // process-wide defaults a new batch starts from (tests and benches run every path through the same API calls)
    struct path_defaults
    {
        int res_S{0}, res_I{0}, res_J{0};
        int subtree_warps{0};
        int res_ws{0};
        int res_chunks{0};
        unsigned tuning{0};  // circuit_batch_set_tuning flags
    };
    path_defaults& default_path();

    void set_last_error(std::string s);
    char const* last_error();
}  // namespace pe_b200
