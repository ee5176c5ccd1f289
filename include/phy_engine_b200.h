/* phy_engine_b200.h — C ABI of libphyengine_b200.so, the B200-native drop-in for Phy-Engine's analog MNA solve path.
 *
 * Part 1 re-declares, with identical names / signatures / return conventions, the analog-path entry points of the
 * reference's C ABI (reference: include/phy_engine/dll_api.h, implemented in src/dll_main.cpp).  A caller bound to
 * the reference library (python/phy_engine/_ffi.py:41-140, phy_lab_wrapper/pe_sim.h:415-428) can load this library
 * instead and every analyze() runs on the GPU.  Part 2 is additive: batch (Monte-Carlo / sweep) handles that the
 * reference does not have (SURVEY.md §8b last row).  Plain pointers and sizes only; no C++/torch types.
 *
 * Error convention (dll_main.cpp:2141-2259, 2467-2490): int 0 = ok, 1 = null/invalid argument or analyze failed,
 * 2 = model not found, 3 = attribute not found; creators return NULL and set the thread-local error string.
 * There is no CPU fallback: without a CUDA device every analyze entry point returns 1 with an explanatory error.
 */
#ifndef PHY_ENGINE_B200_H
#define PHY_ENGINE_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    /* ===== Part 1: reference-compatible surface ================================================================== */

    /* dll_api.h:43-47 */
    char const* phy_engine_last_error(void);
    void phy_engine_clear_error(void);
    void phy_engine_string_free(char* s);

    /* dll_api.h:143-150 / dll_main.cpp:2492-2640.  Element codes: dll_api.h:51-135.  In scope: 1-13 (R, C, L, VDC, VAC,
     * IDC, IAC, VCCS, VCVS, CCCS, CCVS, SPST switch, PN junction), 14-16 (ideal transformer, coupled inductors,
     * centre-tapped transformer), 17 (op-amp), 18 (relay), 19 (comparator: the analog -> digital boundary), 20-23
     * (sawtooth / square / pulse / triangle generators), 50-53 (BJT NPN / PNP, level-1 NMOS / PMOS), 54 (full bridge
     * rectifier).  Any other code makes the creator fail with "element code N is outside the B200 hot path".
     * vec_pos / chunk_pos are malloc()ed and released by destroy_circuit(), like the reference. */
    void* create_circuit(int* elements,
                         size_t ele_size,
                         int* wires,
                         size_t wires_size,
                         double* properties,
                         size_t** vec_pos,
                         size_t** chunk_pos,
                         size_t* comp_size);

    /* dll_api.h:156-168 (Verilog text tables are accepted and ignored; Verilog element codes are out of scope) */
    void* create_circuit_ex(int* elements,
                            size_t ele_size,
                            int* wires,
                            size_t wires_size,
                            double* properties,
                            char const* const* texts,
                            size_t const* text_sizes,
                            size_t text_count,
                            size_t const* element_src_index,
                            size_t const* element_top_index,
                            size_t** vec_pos,
                            size_t** chunk_pos,
                            size_t* comp_size);

    /* dll_api.h:170 */
    void destroy_circuit(void* circuit_ptr, size_t* vec_pos, size_t* chunk_pos);

    /* dll_api.h:173-180 / dll_main.cpp:2141-2267 */
    int circuit_set_analyze_type(void* circuit_ptr, uint32_t analyze_type_value);
    int circuit_set_tr(void* circuit_ptr, double t_step, double t_stop);
    int circuit_set_ac_omega(void* circuit_ptr, double omega);
    int circuit_set_temperature(void* circuit_ptr, double temp_c);
    int circuit_set_tnom(void* circuit_ptr, double tnom_c);
    int circuit_set_model_double_by_name(void* circuit_ptr, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double value);
    int circuit_analyze(void* circuit_ptr);
    int circuit_digital_clk(void* circuit_ptr);

    /* dll_api.h:186-236 / dll_main.cpp:2269-2465 */
    int circuit_sample_layout(void* circuit_ptr,
                              size_t* vec_pos,
                              size_t* chunk_pos,
                              size_t comp_size,
                              size_t* voltage_ord,
                              size_t* current_ord,
                              size_t* digital_ord);
    int circuit_sample(void* circuit_ptr,
                       size_t* vec_pos,
                       size_t* chunk_pos,
                       size_t comp_size,
                       double* voltage,
                       size_t* voltage_ord,
                       double* current,
                       size_t* current_ord,
                       bool* digital,
                       size_t* digital_ord);
    int circuit_sample_u8(void* circuit_ptr,
                          size_t* vec_pos,
                          size_t* chunk_pos,
                          size_t comp_size,
                          double* voltage,
                          size_t* voltage_ord,
                          double* current,
                          size_t* current_ord,
                          uint8_t* digital,
                          size_t* digital_ord);
    int circuit_sample_digital_state_u8(void* circuit_ptr,
                                        size_t* vec_pos,
                                        size_t* chunk_pos,
                                        size_t comp_size,
                                        double* voltage,
                                        size_t* voltage_ord,
                                        double* current,
                                        size_t* current_ord,
                                        uint8_t* digital,
                                        size_t* digital_ord);

    /* dll_api.h:239 (no digital model is in scope: always returns 3) */
    int circuit_set_model_digital(void* circuit_ptr, size_t vec_pos, size_t chunk_pos, size_t attribute_index, uint8_t state);

    /* dll_api.h:242-255 / dll_main.cpp:2899-2934 */
    int analyze_circuit(void* circuit_ptr,
                        size_t* vec_pos,
                        size_t* chunk_pos,
                        size_t comp_size,
                        int* changed_ele,
                        size_t* changed_ind,
                        double* changed_prop,
                        size_t prop_size,
                        double* voltage,
                        size_t* voltage_ord,
                        double* current,
                        size_t* current_ord,
                        bool* digital,
                        size_t* digital_ord);

    /* ===== Part 2: additive surface ============================================================================== */

    /* C++-only knobs of the reference made reachable from C: phy_engine::environment (environment.h:7-22) and the AC
     * sweep settings (analyzer/AC.h).  env8 = V_eps_max, V_epsr_max, I_eps_max, I_epsr_max, g_min, r_open,
     * temperature, norm_temperature. */
    int circuit_set_env(void* circuit_ptr, double const* env8);
    int circuit_set_ac_sweep(void* circuit_ptr, int sweep_type, double omega_start, double omega_stop, size_t points);

    /* unknown vector layout of circult::prepare() (circuit.h:481-540): node_index, then node_counter + branch.index */
    int circuit_unknown_count(void* circuit_ptr, size_t* n_nodes, size_t* n_branches);
    long long circuit_pin_unknown(void* circuit_ptr, size_t vec_pos, size_t chunk_pos, size_t pin); /* -1 ground, -2 n/c */
    long long circuit_branch_unknown(void* circuit_ptr, size_t vec_pos, size_t chunk_pos, size_t branch);
    /* last solution of circuit_analyze(): re[n], im[n] (im may be NULL; zero outside AC) */
    int circuit_get_solution(void* circuit_ptr, double* re, double* im);

    /* Batch of independent instances of one netlist (Monte-Carlo / parameter sweep / frequency sweep). */
    void* circuit_batch_create(void* circuit_ptr, size_t n_instances);
    void circuit_batch_destroy(void* batch);
    int circuit_batch_set_device(void* batch, int device);
    int circuit_batch_set_stream(void* batch, void* cuda_stream);
    /* warps per CTA cooperating on each block of 32 instances (sub-tree parallel LU): 0 = automatic from the lane count,
     * else a power of two <= 16 */
    int circuit_batch_set_subtree_warps(void* batch, int warps);
    /* resident (shared-memory) solve path: `streams` word streams per instance (-1 = never use it, 0 = automatic, else a
     * power of two), `instances_per_cta` (0 = automatic, else a power of two <= 32), `instances_per_thread` (0 = automatic,
     * 1 or 2).  A circuit whose workspace does not fit one CTA's shared memory runs on the HBM-streaming kernel. */
    int circuit_batch_set_resident(void* batch, int streams, int instances_per_cta, int instances_per_thread);
    /* where a tree-scheduled program keeps its workspace: 0 = automatic (shared memory for small circuits, HBM with one
     * warp per sub-tree otherwise), 1 = shared memory, 2 = HBM */
    int circuit_batch_set_workspace(void* batch, int where);
    /* tree-streaming kernel: number of chunks the time loop is cut into for dynamic (group, chunk) scheduling over
     * persistent CTAs: 0 = automatic, 1 = static (one CTA per 32-lane group runs the whole launch), up to 32 */
    int circuit_batch_set_chunks(void* batch, int chunks);
    /* tuning flags of the tree-scheduled kernels (defaults = 0): bit 0 line-ahead L2 prefetch of operand rows, bit 1 ... two
     * lines ahead, bit 2 no L1 re-fetch of a DOT result after its store, bit 3 fuse small elimination steps into one op
     * (PE_OP_CROUT2; takes effect at the next compile), bit 4 require / bit 5 forbid the specialised kernel, bit 6 require /
     * bit 7 forbid the stream kernel */
    int circuit_batch_set_tuning(void* batch, unsigned flags);
    /* which kernel the last tree-streaming launch of the batch ran: 0 = the word interpreter, 1 = the specialised kernel
     * (iter section compiled to straight-line sm_100a code at run time, cached under jit_cache/ next to the library;
     * tuning bit 4 requires it, bit 5 forbids it, default: large linear batches), -1 = bad handle */
    int circuit_batch_last_kernel(void* batch);
    /* stream kernel (large batches of large linear circuits: one warp per lane group, one word stream per group, the iter
     * section generated as tiled code whose cold operand rows arrive by TMA bulk copies; tuning bit 6 requires it, bit 7
     * forbids it; circuit_batch_last_kernel reports 2).  info[6] = last kernel, warps per CTA, ring stages, shared memory
     * per CTA (bytes), tiles per solve, rows per ring stage */
    int circuit_batch_stream_info(void* batch, int mode, int64_t* info);
    /* pivot safety net.  The elimination order is static (chosen on the first instance's values; the reference re-pivots in
     * every solve, Eigen SparseLU_pivotL.h:76-107).  Pivots that are not provably safe are tested on the device: a lane with an
     * entry of L above 1 / `guard` (the pivot is that much smaller than an entry of its column) is flagged, and analyze() solves the flagged
     * instances again in a sub-batch whose order is chosen on THEIR values; what still fails goes to the next of `rounds`
     * rounds, the last of which runs unguarded.  guard: 0 = off, < 0 = default (2^-20); rounds < 0 = default (3).
     * info[6] = guarded pivots of the program, instances flagged, instances solved by a re-ordered sub-batch, instances that
     * took the unguarded round, sub-batch runs, live sub-batches */
    int circuit_batch_set_pivot_guard(void* batch, double guard, int rounds);
    /* batch-state checkpoint (the reference persists ONE circuit with its netlist, pe_nl_fileformat.h:584-657, 805-1046; a
     * batch persists what its instances carry between analyze() calls): solution, per-instance parameters, companion / device
     * state of every instance, the clock of the transient and the sub-batches of the pivot safety net, in one blob.  A batch of
     * the same netlist, parameter keys and instance count resumes from it bit-identically (load checks a fingerprint).
     * save: buffer == NULL returns the size in *size; 0 = ok, 1 = error, 2 = buffer too small */
    int circuit_batch_save_state(void* batch, void* buffer, size_t capacity, size_t* size);
    int circuit_batch_load_state(void* batch, void const* buffer, size_t size);
    int circuit_batch_rescue_info(void* batch, int mode, int64_t* info);
    /* reduce-and-core path: DC / OP of one huge linear circuit (resistors, DC sources; >= 20 000 unknowns, config A of
     * BASELINE.json) per instance: level-scheduled elimination of the degree <= 2 nodes, dense LU of the rest on the FP64
     * tensor cores (circuit_batch_last_kernel reports 3).  info[11] = unknowns, eliminated nodes, levels, core rows, edges
     * (fill included), kernel launches of the last solve, leading dimension of the dense core, edges inside the core, device
     * microseconds of the reduction / the core LU / the substitutions of the last solve */
    int circuit_batch_frontal_info(void* batch, int64_t* info);
    int phy_engine_b200_set_frontal_min(size_t n_unknowns); /* size threshold of that path (default 20 000 unknowns) */
    /* tooling (after circuit_batch_compile_host; no device needed): generated source of the stream kernel's sections and
     * build-or-fetch of its module.  stats[8] = tiles per solve, rows per ring stage, loops, ops in loops, ops, rows
     * fetched / bulk copies / rows stored per solve */
    size_t circuit_batch_stream_source(void* batch, int mode, char* out, size_t cap, uint64_t* stats);
    int circuit_batch_stream_build(void* batch, int mode);
    /* tooling for the specialised kernel of a compiled batch (circuit_batch_compile_host first; no device needed): the
     * generated source of the iter section (returns its length, copies at most cap bytes), and build-or-fetch of its cubin
     * (cluster = CTAs per lane group, 1 or 2) */
    size_t circuit_batch_jit_source(void* batch, int mode, char* out, size_t cap);
    int circuit_batch_jit_build(void* batch, int mode, int cluster);
    /* info[13] = resident, streams, workspace slots per instance, I, J (0 = does not fit), io entries,
     * last launch S / I / J (0 = the flat HBM-streaming kernel ran), phases of the iter section, words, longest warp
     * stream, workspace in HBM (1) or shared memory (0) */
    int circuit_batch_resident_info(void* batch, int mode, int64_t* info);
    /* word offsets [3][n_warps] of the prep / step / iter stream of each warp of the packed resident program (0xffffffff =
     * absent); returns the number of entries (out may be NULL) */
    size_t circuit_batch_resident_secoff(void* batch, int mode, uint32_t* out);
    /* per-instance values of one model attribute; values[n_instances] in the attribute's public unit */
    int circuit_batch_set_param(void* batch, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double const* values);
    /* many parameters in one call: values[n_params][n_instances].  Once the batch is prepared and every parameter
     * already has a device row, rows are copied host->device straight from `values` (pinned memory recommended). */
    int circuit_batch_set_params(void* batch, size_t n_params, size_t const* vec_pos, size_t const* chunk_pos, char const* const* names, double const* values);
    int circuit_batch_set_ac_sweep(void* batch, int sweep_type, double omega_start, double omega_stop, size_t points);
    /* record these unknowns after every transient step (waveform store [steps][n_probes][n_instances]) */
    int circuit_batch_set_probes(void* batch, size_t const* unknowns, size_t n_probes);
    /* compile + allocate + upload parameters without running (so device pointers below are valid) */
    int circuit_batch_prepare(void* batch);
    /* zero x and all device state, keep the per-instance parameters (a fresh Monte-Carlo start) */
    int circuit_batch_reset_state(void* batch);
    /* analysis type / settings are those of the parent circuit.  0 = every lane ok; 1 = error or some lane failed. */
    int circuit_batch_analyze(void* batch);

    size_t circuit_batch_lanes(void* batch);  /* n_instances * frequency points of the last analyze */
    size_t circuit_batch_points(void* batch);
    uint64_t circuit_batch_total_solves(void* batch); /* sum of solve_once-equivalents in the last analyze */
    double circuit_batch_tr_duration(void* batch);
    int circuit_batch_solution(void* batch, double* x);        /* [n_instances][n] real state */
    int circuit_batch_solution_soa(void* batch, double* x);    /* [n][n_instances]: the device-native layout, no transpose */
    int circuit_batch_ac_solution(void* batch, double* x);     /* [lanes][n][2] */
    /* ... of selected lanes only (lane = instance * points + point): x[n_lanes][n][2]; a sample of a sweep too large to download */
    int circuit_batch_ac_solution_lanes(void* batch, size_t const* lanes, size_t n_lanes, double* x);
    int circuit_batch_ac_omegas(void* batch, double* omegas);  /* [points] */
    /* one rank's shard of a sweep: only the points [first, first + count) are solved (count = 0: all); the omega table is
     * built in full with the reference's cumulative product (circuit.h:412-428) and sliced */
    int circuit_batch_set_ac_slice(void* batch, size_t first, size_t count);
    int circuit_batch_status(void* batch, int32_t* status);    /* [lanes] 0 ok, 1 no convergence, 2 singular */
    int circuit_batch_newton_iters(void* batch, uint32_t* n);  /* [lanes] solves per lane */
    int circuit_batch_waveform(void* batch, double* w);        /* [steps][n_probes][n_instances] */
    /* mixed-signal boundary, analog -> digital: the update_digital_clk of every comparator (element 19) of the netlist,
     * vA >= vB (controller/comparator.h:73-108), for every instance, on the device; states [n_instances][n_comparators]
     * in element order.  Digital logic behind the comparators stays with the caller. */
    int circuit_batch_digital_clk(void* batch);
    size_t circuit_batch_comparator_count(void* batch);
    int circuit_batch_comparator_states(void* batch, uint8_t* states);
    /* program statistics for roofline accounting: mode 0 DC/OP, 1 TR, 2 TROP, 3 AC */
    int circuit_batch_stats(void* batch, int mode, size_t* n_unknowns, size_t* nnz_a, size_t* nnz_lu, size_t* n_fma, size_t* n_lane_slots, size_t* n_inst_slots);
    /* HBM-resident access: device row of a swept parameter (n_instances doubles) and of the real solution
     * (unknown j at x0 + j * lane_stride) */
    int circuit_batch_param_device_ptr(void* batch, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size, double** dptr);
    int circuit_batch_solution_device_ptr(void* batch, double** x0, size_t* lane_stride);

    /* ===== Part 3: introspection of the symbolic phase (no device needed; used by tests and DESIGN.md tooling) ===== */
    int circuit_batch_compile_host(void* batch);
    size_t circuit_batch_program_words(void* batch, int mode);
    int circuit_batch_program_copy(void* batch, int mode, uint32_t* out);
    size_t circuit_batch_const_count(void* batch);
    int circuit_batch_const_copy(void* batch, double* out);
    /* info[64]: [0..15] = cplx, structurally_singular, n_lane_slots, omega_slot, n_inst_slots, dt_slot, x_slot0, n_unknowns,
     * warps, n_real_lane_slots, n_leaves, n_leaf_rows, n_top_rows, max_warp_words, nnz_a, nnz_lu;
     * [16 + 16 * s + g] = word offset of warp g's stream of section s (0 prep, 1 step, 2 iter), -1 if absent */
    int circuit_batch_program_info(void* batch, int mode, int64_t* info);
    long long circuit_batch_swept_slot(void* batch, size_t vec_pos, size_t chunk_pos, char const* name, size_t name_size);
    int circuit_batch_swept_values(void* batch, long long slot, double* out);

    /* process-wide defaults every batch created afterwards (including the one behind circuit_analyze) starts from:
     * the arguments of circuit_batch_set_resident, circuit_batch_set_subtree_warps, circuit_batch_set_workspace and
     * circuit_batch_set_tuning */
    int phy_engine_b200_set_default_path(int streams, int instances_per_cta, int instances_per_thread, int subtree_warps, int workspace, unsigned tuning);
    int phy_engine_b200_device_count(void);
    uint64_t phy_engine_b200_launch_count(void);     /* solve-kernel launches of this process */
    uint64_t phy_engine_b200_aux_launch_count(void); /* launches of the small helper kernels (status reduction) */
    /* device-side timing of the solve kernels (CUDA events on the launching stream): enable, then read-and-reset the
     * accumulated milliseconds (waits for the launches to finish) */
    void phy_engine_b200_timing(int on);
    double phy_engine_b200_kernel_ms(void);

    /* ===== Part 3: the rest of dll_api.h (dll_api.h:252-440) ====================================================== */
    /* Verilog synthesis options / runtime, PhysicsLab experiment and adapter handles, PE -> PhysicsLab conversion,
     * auto-layout: outside the analog solve path.  Exported with the reference's signatures so that callers that bind the
     * whole ABI (python/phy_engine/_ffi.py:150-420) load this library; host/c_api_forward.cpp forwards each call to a build
     * of the reference's own library when PHY_ENGINE_REF_LIB names one and fails with the reference's error convention
     * (NULL / 1 / 0 + phy_engine_last_error()) otherwise.  The verilog_synth_* options work natively. */
    void verilog_synth_set_opt_level(uint8_t level);
    uint8_t verilog_synth_get_opt_level(void);
    void verilog_synth_set_assume_binary_inputs(bool value);
    bool verilog_synth_get_assume_binary_inputs(void);
    void verilog_synth_set_allow_inout(bool value);
    bool verilog_synth_get_allow_inout(void);
    void verilog_synth_set_allow_multi_driver(bool value);
    bool verilog_synth_get_allow_multi_driver(void);
    void verilog_synth_set_optimize_wires(bool value);
    bool verilog_synth_get_optimize_wires(void);
    void verilog_synth_set_optimize_mul2(bool value);
    bool verilog_synth_get_optimize_mul2(void);
    void verilog_synth_set_optimize_adders(bool value);
    bool verilog_synth_get_optimize_adders(void);
    void verilog_synth_set_loop_unroll_limit(size_t n);
    size_t verilog_synth_get_loop_unroll_limit(void);
    void* verilog_runtime_create(char const* src, size_t src_size, char const* top, size_t top_size, char const* const* include_dirs, size_t const* include_dir_sizes, size_t include_dir_count);
    void verilog_runtime_destroy(void* runtime_ptr);
    uint64_t verilog_runtime_get_tick(void* runtime_ptr);
    int verilog_runtime_reset(void* runtime_ptr);
    int verilog_runtime_step(void* runtime_ptr, uint64_t tick, uint8_t process_sequential);
    int verilog_runtime_tick(void* runtime_ptr);
    size_t verilog_runtime_module_count(void* runtime_ptr);
    size_t verilog_runtime_port_count(void* runtime_ptr);
    size_t verilog_runtime_signal_count(void* runtime_ptr);
    size_t verilog_runtime_preprocessed_size(void* runtime_ptr);
    int verilog_runtime_copy_preprocessed(void* runtime_ptr, char* out, size_t out_size);
    size_t verilog_runtime_top_module_name_size(void* runtime_ptr);
    int verilog_runtime_copy_top_module_name(void* runtime_ptr, char* out, size_t out_size);
    size_t verilog_runtime_module_name_size(void* runtime_ptr, size_t module_index);
    int verilog_runtime_copy_module_name(void* runtime_ptr, size_t module_index, char* out, size_t out_size);
    size_t verilog_runtime_port_name_size(void* runtime_ptr, size_t port_index);
    int verilog_runtime_copy_port_name(void* runtime_ptr, size_t port_index, char* out, size_t out_size);
    uint8_t verilog_runtime_port_dir(void* runtime_ptr, size_t port_index);
    uint8_t verilog_runtime_get_port_value(void* runtime_ptr, size_t port_index);
    int verilog_runtime_set_port_value(void* runtime_ptr, size_t port_index, uint8_t state);
    size_t verilog_runtime_signal_name_size(void* runtime_ptr, size_t signal_index);
    int verilog_runtime_copy_signal_name(void* runtime_ptr, size_t signal_index, char* out, size_t out_size);
    uint8_t verilog_runtime_get_signal_value(void* runtime_ptr, size_t signal_index);
    int verilog_runtime_set_signal_value(void* runtime_ptr, size_t signal_index, uint8_t state);
    void* pl_experiment_create(int type_value);
    void* pl_experiment_load_from_string(char const* sav_json, size_t sav_json_size);
    void* pl_experiment_load_from_file(char const* path, size_t path_size);
    void pl_experiment_destroy(void* experiment_ptr);
    char* pl_experiment_dump(void* experiment_ptr, int indent);
    int pl_experiment_save(void* experiment_ptr, char const* path, size_t path_size, int indent);
    char* pl_experiment_add_circuit_element(void* experiment_ptr, char const* model_id, size_t model_id_size, double x, double y, double z, uint8_t element_xyz_coords, uint8_t is_big_element, uint8_t participate_in_layout);
    int pl_experiment_connect(void* experiment_ptr, char const* src_id, size_t src_id_size, int src_pin, char const* dst_id, size_t dst_id_size, int dst_pin, int color_value);
    int pl_experiment_clear_wires(void* experiment_ptr);
    int pl_experiment_set_xyz_precision(void* experiment_ptr, int decimals);
    int pl_experiment_set_element_xyz(void* experiment_ptr, uint8_t enabled, double origin_x, double origin_y, double origin_z);
    int pl_experiment_set_camera(void* experiment_ptr, double vision_center_x, double vision_center_y, double vision_center_z, double target_rotation_x, double target_rotation_y, double target_rotation_z);
    int pl_experiment_set_element_property_number(void* experiment_ptr, char const* element_id, size_t element_id_size, char const* key, size_t key_size, double value);
    int pl_experiment_set_element_label(void* experiment_ptr, char const* element_id, size_t element_id_size, char const* label, size_t label_size);
    int pl_experiment_set_element_position(void* experiment_ptr, char const* element_id, size_t element_id_size, double x, double y, double z, uint8_t element_xyz_coords);
    int pl_experiment_merge(void* dst_experiment_ptr, void* src_experiment_ptr, double offset_x, double offset_y, double offset_z);
    void* pl_pe_circuit_build(void* experiment_ptr);
    void pl_pe_circuit_destroy(void* pe_circuit_ptr);
    size_t pl_pe_circuit_comp_size(void* pe_circuit_ptr);
    int pl_pe_circuit_set_analyze_type(void* pe_circuit_ptr, uint32_t analyze_type_value);
    int pl_pe_circuit_set_tr(void* pe_circuit_ptr, double t_step, double t_stop);
    int pl_pe_circuit_set_ac_omega(void* pe_circuit_ptr, double omega);
    int pl_pe_circuit_analyze(void* pe_circuit_ptr);
    int pl_pe_circuit_digital_clk(void* pe_circuit_ptr);
    int pl_pe_circuit_sync_inputs_from_pl(void* pe_circuit_ptr, void* experiment_ptr);
    int pl_pe_circuit_write_back_to_pl(void* pe_circuit_ptr, void* experiment_ptr);
    int pl_pe_circuit_write_back_to_pl_ex(void* pe_circuit_ptr, void* experiment_ptr, double logic_output_low, double logic_output_high, double logic_output_x, double logic_output_z);
    int pl_pe_circuit_sample_layout(void* pe_circuit_ptr, size_t* voltage_ord, size_t* current_ord, size_t* digital_ord);
    int pl_pe_circuit_sample_u8(void* pe_circuit_ptr, double* voltage, size_t* voltage_ord, double* current, size_t* current_ord, uint8_t* digital, size_t* digital_ord);
    int pl_pe_circuit_sample_digital_state_u8(void* pe_circuit_ptr, double* voltage, size_t* voltage_ord, double* current, size_t* current_ord, uint8_t* digital, size_t* digital_ord);
    void* pe_to_pl_convert(void* circuit_ptr, double fixed_x, double fixed_y, double fixed_z, uint8_t element_xyz_coords, uint8_t keep_pl_macros, uint8_t include_linear, uint8_t include_ground, uint8_t generate_wires, uint8_t keep_unknown_as_placeholders, uint8_t drop_dangling_logic_inputs);
    int pl_experiment_auto_layout(void* experiment_ptr, double corner0_x, double corner0_y, double corner0_z, double corner1_x, double corner1_y, double corner1_z, double z_fixed, int backend_value, int mode_value, double step_x, double step_y, double margin_x, double margin_y, size_t* out_grid_w, size_t* out_grid_h, size_t* out_fixed_obstacles, size_t* out_placed, size_t* out_skipped);

#ifdef __cplusplus
}
#endif
#endif
