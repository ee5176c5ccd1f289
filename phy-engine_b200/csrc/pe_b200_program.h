// pe_b200_program.h — the POD seam between the C++23 host side (netlist -> symbolic analysis -> batch program)
// and the sm_100a kernels.  Plain C types only: nvcc cannot include any phy_engine / fast_io header
// (SURVEY.md probe table), so nothing here refers to them.
//
// Execution model: one *lane* = one independent circuit solve stream (a Monte-Carlo instance, or one
// (instance, frequency point) pair in an AC sweep).  All per-lane doubles live in *lane-interleaved* HBM arrays
// w[slot][lane], so that the 32 threads of a warp (= 32 consecutive lanes) touching slot s read 32 consecutive
// doubles (one fully coalesced 256-byte request).  A CTA is 32 lanes x G warps: the G warps cooperate on the SAME
// 32 lanes, each interpreting its own word stream (its sub-trees of the elimination tree) with CTA barriers where
// the streams meet.
//
// Operand spaces:
//   CONST  cst[slot]                         values shared by every lane (broadcast parameters, folded constants)
//   U      wu[slot * LSu + lane]             the lane-strided workspace of this launch
//   INSTX  wx[slot * LSx + lane / ppi]       per-instance values seen from a frequency-point lane (AC sweeps only)
// In the real-valued modes (DC/OP/TR/TROP) lanes == instances and everything (x, parameters, device state, matrix
// and rhs scratch) is one U workspace; in AC the U workspace holds the per-point complex scratch and INSTX the
// per-instance values.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C"
{
#endif

    // ---- operand encoding (uint32) ---------------------------------------------------------------------------------
    //   bit 31      negate
    //   bits 30:29  space
    //   bits 28:0   slot
    enum
    {
        PE_SP_CONST = 0,
        PE_SP_U = 1,
        PE_SP_INSTX = 2,
    };
#define PE_OPND(space, slot) ((((uint32_t)(space)) << 29) | ((uint32_t)(slot) & 0x1fffffffu))
#define PE_OPND_NEG 0x80000000u
#define PE_OPND_SPACE(o) (((o) >> 29) & 3u)
#define PE_OPND_SLOT(o) ((o) & 0x1fffffffu)

    // ---- opcodes; header word: bits 7:0 opcode, the rest op-specific --------------------------------------------
    enum pe_b200_opcode
    {
        PE_OP_END = 0,
        PE_OP_BAR = 1,  // CTA barrier: every warp stream of a section holds the same number of them

        // The one numeric-factorisation primitive.  Every matrix entry of L, U, the pivots, the forward-substituted
        // rhs and the solution is produced exactly once by
        //     v = sum_i (+/-)src_i  -  sum_p a_p * b_p ;   [v *= scale] ;  [v = 1 / v] ;  store
        // (row-wise "dot product" form of LU on the fixed pattern: stamp assembly, elimination, forward and back
        // substitution are all this op).  header = op | flags << 8 | nsrc << 16 | npair << 24
        //   words: [dst (U slot)] [scale (U slot), if F_SCALE] [src operand x nsrc] [(a, b) U slots x npair]
        PE_OP_DOT = 2,
        // complex variant (AC): slots hold re at s, im at s + 1.
        //   header = op | flags << 8 | npair << 16 ; second word = nsrc_re | nsrc_im << 16
        //   words: [dst] [scale, if F_SCALE] [src_re x nsrc_re] [src_im x nsrc_im] [(a, b) x npair]
        PE_OP_CDOT = 3,
        // resident programs packed for one stream per warp: the rest of this 32-word line of the main stream is padding
        // (a vector op never straddles two lines, so that the warp reader hands out its words with one shuffle each)
        PE_OP_SKIP = 4,
        // resident programs: one whole Crout elimination step whose pivot has at most two U and two L entries, fused:
        // six DOT slots [pivot | up to five of: U entries, L entries, the rhs entry], semantically the six DOTs in a row,
        // with the L entries scaled by the fresh pivot reciprocal from a register.  19 rows (header field a = 19):
        //   slot 0: [ctl][src][src][pair]      slots 1..5: [ctl][src][pair]      (ctl = dst | ACTIVE << 15 | flags << 16)
        // The kernels request all 26 operands of the step before they use the first one.
        PE_OP_CROUT2 = 5,

        // scalar value ops (real), generic operands
        PE_OP_RECIP = 10,    // [dst][a]          dst = 1.0 / a
        PE_OP_MUL = 11,      // [dst][a][b]       dst = a * b
        PE_OP_SUB = 12,      // [dst][a][b]       dst = a - b
        PE_OP_COPY = 13,     // [dst][a]
        PE_OP_VSIN = 14,     // [dst][Vp][omega][phase]   dst = Vp * sin(omega * t + phase)      (VAC.h:176, IAC.h:154)
        PE_OP_SINCOS = 15,   // [dre][dim][Vp][phase]     dre = Vp cos(phase), dim = Vp sin(phase) (VAC.h:115-121)
        PE_OP_MUL2DIV = 16,  // [dst][a][b]       dst = 2.0 * a / b                              (2C/dt, 2L/dt)

        PE_OP_KMUT = 17,     // [dst][k][L1][L2]  dst = k * sqrt(L1 * L2)    (mutual inductance, coupled_inductors.h:135, :170)

        // trapezoidal companions (capacitor.h:106-128, inductor.h:134-160)
        PE_OP_CAP_STEP = 20,  // [hist][prev_g][C][dt][va][vb]
        PE_OP_IND_STEP = 21,  // [req][ueq][L][dt][va][vb][ib]

        // one winding of a pair of coupled inductors (coupled_inductors.h:160-200): rA = (2 / dt) LA, rB = (2 / dt) LB,
        // ueq = -(vp - vn) - (rA ia + rB ib); winding 1: (LA, LB) = (L1, M), winding 2: (M, L2), (ia, ib) = (i1, i2) for both
        PE_OP_KIND_STEP = 23,  // [rA][rB][ueq] <- [LA][LB][dt][vp][vn][ia][ib]

        // time-domain generators (generator/{sawtooth,square,pulse,triangle}.h: iterate_tr_define; DC = the value at t = 0)
        PE_OP_GEN_EVAL = 24,  // [dst] <- [kind 0 saw / 1 square / 2 pulse / 3 triangle][tsel 1 = section time, 0 = t = 0][Vh][Vl][freq][duty][phase][tr][tf]

        // relay (controller/relay.h:74-105): hysteresis on the coil voltage, contact resistance 0 / r_open
        PE_OP_RELAY_EVAL = 22,  // [engaged][r_contact] <- [vcp][vcn][Von][Voff][r_open]   (engaged: 0.0 / 1.0, updated in place)

        // PN junction (PN_junction.h)
        PE_OP_PN_PREP = 30,   // [is_eff][isr_eff][bv_eff][ut][uth] <- [Is][Isr][Area][N][Temp][Ibv][Bv][bv_set]
        PE_OP_PN_EVAL = 31,   // [ud_last][geq][ieq] <- [va][vb][is_eff][isr_eff][bv_eff][ut][uth][N][Nr][bv_set]
        PE_OP_PN_STEP = 32,   // [ud_last][hist][prev_g] <- [va][vb][geq][tt][dt]
        PE_OP_PN_ACCAP = 33,  // [dst] <- [geq][tt][omega]     dst = (omega!=0 && tt>0 && geq>0 && tt*geq>0) ? tt*geq*omega : 0

        // BJT (BJT_NPN.h:116-159; PNP = same with the controlling voltage operands swapped)
        PE_OP_BJT_PREP = 40,  // [ut] <- [Temp]
        PE_OP_BJT_EVAL = 41,  // [geq][ieq_be][gm][ieq_c] <- [vp][vm][Is][Area][N][ut][BetaF]

        // level-1 MOSFET (nmosfet.h:84-141, pmosfet.h:84-138)
        PE_OP_NMOS_EVAL = 50,  // [gm][gds][ieq] <- [vd][vg][vs][Kp][lambda][Vth]
        PE_OP_PMOS_EVAL = 51,  // [gm][gds][ieq] <- [vd][vg][vs][Kp][lambda][Vth]
    };

    // DOT / CDOT flags
    enum
    {
        PE_F_SCALE = 1,    // multiply by the value at the scale slot (a stored pivot reciprocal)
        PE_F_RECIP = 2,    // store the reciprocal; a zero / non-finite value marks the lane singular
        PE_F_CHECK_V = 4,  // Newton test against the old value at dst with the voltage tolerances (circuit.h:923-933)
        PE_F_CHECK_I = 8,  // ... with the branch-current tolerances (circuit.h:937-947)
        // pivot guard, on the ops that produce an entry of L (PE_F_SCALE by a pivot that is not provably safe, compiler.cpp "guard
        // elision"): the lane is marked singular when |l| * guard > 1, i.e. when the pivot is more than 1 / guard times smaller than
        // an entry of its column.  The order was chosen with threshold pivoting (|l| <= 1e3 on the nominal values), and a partial-
        // pivoting LU would have taken another row here: the static elimination order does not suit this lane's values.
        // The host re-runs such lanes with an order chosen on THEIR values (batch.cpp "rescue"); Eigen's SparseLU re-pivots per
        // solve instead (SparseLU_pivotL.h:76-107).  guard == 0 switches the test off.
        PE_F_GUARD = 16,
    };
#define PE_GUARD_DEFAULT 9.5367431640625e-07 /* 2^-20: multipliers above ~1e6 */
#define PE_GUARD_TRIP(l_abs, guard) ((l_abs) * (guard) > 1.0)
#define PE_DOT_MAX_SRC 255u
#define PE_DOT_MAX_PAIR 255u
#define PE_MAX_WARPS 16

    // ---- lane status ------------------------------------------------------------------------------------------------
    enum
    {
        PE_ST_OK = 0,
        PE_ST_NO_CONVERGENCE = 1,  // 64 Newton iterations without convergence (circuit.h:898,984)
        PE_ST_SINGULAR = 2,        // zero / non-finite pivot (circuit.h:1517 factorizationIsOk() == false)
    };

    // A section = one word stream per warp of the CTA; off[g] is the start of warp g's stream inside `words`
    // (PE_NO_SECTION: the section is absent).
#define PE_NO_SECTION 0xffffffffu
    typedef struct pe_b200_section
    {
        uint32_t off[PE_MAX_WARPS];
    } pe_b200_section;

    // ---- one kernel launch = one analysis phase over all lanes -----------------------------------------------------
    typedef struct pe_b200_run
    {
        uint32_t const* words;  // device: program word pool
        pe_b200_section prep;   // once per launch
        pe_b200_section step;   // once per time step, before the Newton loop (TR only)
        pe_b200_section iter;   // one linearised MNA solve: eval + assemble + factor + substitute (+ convergence test)
        double const* cst;
        double* wu;
        double* wx;
        int32_t* status;   // [n_lanes]  (in/out: lanes with status != 0 on entry keep their status and are not counted)
        uint32_t* solves;  // [n_lanes]  += number of solve_once-equivalents executed
        double* wave;      // optional waveform store [n_steps][n_probe][LSu] (NULL = off)
        uint32_t const* probes;  // [n_probe] operand words

        int64_t LSu;  // lane stride (elements) of wu; a multiple of 32, >= n_lanes
        int64_t LSx;  // stride of wx
        int32_t n_lanes;
        int32_t ppi;        // lanes per instance (frequency points per instance in an AC sweep, else 1)
        int32_t warps;      // G: warps per CTA = word streams per section (1..PE_MAX_WARPS)
        int32_t cplx;       // informational: 1 when the iter section is a complex (AC) program
        int32_t nonlinear;  // 0: single solve per step; 1: Newton loop with convergence test
        int32_t max_iter;   // 64
        int32_t n_steps;    // >= 0
        int32_t n_probe;
        int32_t time_stepping;  // 1: run `step` section and advance t by dt before each step's solve
        double t0;              // tr_duration at entry
        double dt;
        double v_abstol, v_reltol, i_abstol, i_reltol;
        double guard;  // PE_F_GUARD threshold (0 = off)
    } pe_b200_run;

    // ================================================================================================================
    // RESIDENT programs: the whole per-instance workspace of a CTA lives in shared memory for the whole
    // launch (all time steps / Newton iterations / the frequency point), HBM is touched only by the load of the
    // persistent values at kernel entry and the store of the mutable ones at exit.
    //
    // CTA = S streams x I instances.  Thread (s, ig) runs word stream s for the J = I / IG instances
    // [ig * J, ig * J + J) of the CTA (tid = s * IG + ig).  A warp therefore carries C = 32 / IG consecutive streams;
    // its program is a sequence of VECTOR OPS: one warp-uniform header followed by rows of C words, column c belonging
    // to stream (warp * C + c).  All C streams execute the same opcode with the same (padded) trip counts, so there
    // is no divergence; every operand is a 15-bit shared-memory slot (+ negate bit), the value of slot q for instance
    // i of the CTA lives at ws[q * I + i].  Slots are allocated "interleaved": the k-th value owned by stream s is
    // slot k * S + s, so the threads of a warp touching "their" k-th value hit consecutive addresses (no bank
    // conflicts), and a complex value (re at q) has its imaginary part at q + S.
    //
    //   per warp and section, two word sequences: the MAIN stream holds, per vector op, [h0][mask][uniform rows ...]
    //   (BAR / END / SKIP are single words); the SIDE stream holds the per-column rows (C words each) in consumption order
    //   h0                op | a << 8 | a2 << 13 | b << 18 | flags << 24
    //                     DOT / CDOT: a = rows of packed (re-)sources, a2 = rows of packed im-sources (CDOT), b = rows of
    //                     pairs, flags = union of the streams' flags; value ops: a = number of operand rows
    //   mask              bit r set = row r is per-column (side stream), clear = warp-uniform (one word, main stream)
    //   DOT / CDOT rows   [ctl] [scale] [src_re x a] [src_im x a2] [pair x b]
    //   value-op rows     one operand per row, the same order as the v2 encoding; row 0 carries VACTIVE
    //   ctl word          dst | ACTIVE << 15 | flags << 16
    //   src word          s0 | neg0 << 15 | s1 << 16 | neg1 << 31      (padding: the -0.0 operand)
    //   pair word         a | b << 16                                   (padding: (zero, zero))
    //   every operand is STREAM-RELATIVE: row * S + ((column - stream) mod S); the executing stream adds its own index
    //   back (mod S), which is what makes the rows of isomorphic sub-trees identical (warp-uniform)
    //   one stream per warp (C = 1): every 32-word line of the main stream starts with two prefetch bitmaps (bit l of
    //   word 0 / word 1: the low / high half of word l names a workspace row that is cold when used) and an op never
    //   straddles two lines (PE_OP_SKIP pads)
#define PE_R_SLOT(w) ((w) & 0x7fffu)
#define PE_R_NEG 0x8000u
#define PE_R_ACTIVE 0x8000u       /* in a DOT ctl word */
#define PE_R_VACTIVE 0x10000u     /* in row 0 of a value op */
#define PE_R_MAX_SLOTS 32768
#define PE_R_MAX_WARPS 32

    // load / store table of a resident program: which shared-memory slots are filled from (and written back to) HBM
    enum
    {
        PE_IO_CONST = 0,  // cst[src]
        PE_IO_U = 1,      // wu[src * LSu + lane]
        PE_IO_INSTX = 2,  // wx[src * LSx + lane / ppi]
    };
#define PE_IO_LOAD 1u
#define PE_IO_STORE 2u
    typedef struct pe_b200_io
    {
        uint32_t slot_kind;  // slot | kind << 16 | (LOAD | STORE) << 20
        uint32_t src;
    } pe_b200_io;

    typedef struct pe_b200_rrun
    {
        uint32_t const* words;
        uint32_t const* sec_off;  // device: [3][n_warps] offsets of the prep / step / iter stream of each warp (PE_NO_SECTION)
        pe_b200_io const* io;
        int32_t n_io;
        int32_t has_prep, has_step;
        double const* cst;
        double* wu;
        double* wx;
        int32_t* status;
        uint32_t* solves;
        double* wave;
        uint32_t const* probes;  // [n_probe] shared-memory slots
        int64_t LSu, LSx;
        int32_t n_lanes;
        int32_t ppi;
        int32_t S;        // streams per instance
        int32_t I;        // instances per CTA
        int32_t J;        // instances per thread (1 or 2; I % J == 0)
        int32_t n_slots;  // shared-memory slots per instance (a multiple of S)
        // Tree-streaming form of the same program: the workspace ws[slot][lane] lives in HBM
        // (wsg != NULL, lane stride LSw), a CTA is S warps x 32 lanes (I = 32, J = 1), warp s runs stream s.
        double* wsg;
        int64_t LSw;
        // Dynamic (group, chunk) scheduling of the tree-streaming form: the time loop of every 32-lane group is cut into
        // n_chunks chunks of chunk_steps steps; persistent CTAs take (chunk, group) items from an atomic counter in
        // chunk-major order, so a group migrates between SMs and the last wave is one chunk long instead of one run.
        // sched = NULL: one CTA per group runs the whole launch.
        int32_t cluster;      // tree-streaming form: CTAs per lane group (1, or 2 = thread-block cluster; needs J = 4); the S
                              // sub-tree warps are split evenly over them
        int32_t regs128;      // tree-streaming form: 1 = the program has fused elimination steps, give CTAs of <= 512 threads 128 registers
        int32_t prefetch;     // tree-streaming form: 1 = L2 prefetch of the operands of the next line of program words
        int32_t chunk_steps;
        int32_t n_chunks;
        uint32_t* sched;      // device: [0] next item, [1 + g] chunks of group g completed (zeroed before the launch)
        double t_chunk[32];   // tr_duration at the start of chunk c (the host accumulates t = t + dt like circuit.h:243-248)
        int32_t cplx;
        int32_t nonlinear;
        int32_t max_iter;
        int32_t n_steps;
        int32_t n_probe;
        int32_t time_stepping;
        double t0, dt;
        double v_abstol, v_reltol, i_abstol, i_reltol;
        double guard;  // PE_F_GUARD threshold (0 = off)
    } pe_b200_rrun;

    // ---- device seam (implemented in pe_b200_kernels.cu; everything CUDA stays behind these) ---------------------
    int pe_b200_dev_count(void);
    int pe_b200_dev_set(int device);
    int pe_b200_dev_malloc(void** p, size_t bytes);
    int pe_b200_dev_free(void* p);
    int pe_b200_dev_memset0(void* p, size_t bytes, void* stream);
    int pe_b200_dev_h2d(void* dst, void const* src, size_t bytes, void* stream);
    int pe_b200_dev_d2h(void* dst, void const* src, size_t bytes, void* stream);
    int pe_b200_dev_sync(void* stream);
    // strided 2-D copies for [slot][lane] <-> packed host arrays
    int pe_b200_dev_h2d_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream);
    int pe_b200_dev_d2h_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream);
    int pe_b200_launch(pe_b200_run const* run, void* stream);
    // analog -> digital boundary: out[c * LS + lane] = (x[a_c][lane] >= x[b_c][lane]) for the comparators c (ab = pairs of
    // unknown indices, -1 = ground = 0 V), x = lane-interleaved solution rows x[unknown * LS + lane]
    int pe_b200_compare(double const* x, int64_t LS, int32_t n_lanes, int32_t const* ab, int32_t n_cmp, uint8_t* out, void* stream);
    // out[0] = lanes whose status is not PE_ST_OK, out[1] = lanes marked PE_ST_SINGULAR, out[2] = sum of the solve counters (device
    // memory, 3 x uint64, zeroed by the call): a batch of a million lanes reads back 24 bytes instead of 8 MB when nothing failed
    int pe_b200_status_reduce(int32_t const* status, uint32_t const* solves, int64_t n_lanes, unsigned long long* out3, void* stream);
    int pe_b200_launch_resident(pe_b200_rrun const* run, void* stream);
    // specialised tree-streaming kernel (host/jit.cpp): the same launch with the kernel taken from a cubin (loaded once
    // per process, device and key); geometry J = 4 only
    int pe_b200_jit_supported(void);
    int pe_b200_launch_jit(pe_b200_rrun const* run, void const* cubin, size_t bytes, uint64_t key, void* stream);
    // stream kernel (host/stream.cpp, csrc/pe_b200_stream.cu): one warp per lane group of GL J lanes (GL = 32, or 16 / 8 with
    // the other lanes of the warp idle: run->I = GL J), S = 1, the iter
    // section taken from a run-time compiled module (device: cubin bytes; emulator: path of a host shared object).
    // pe_b200_stream_build turns generated source into that module with the toolchain of the seam's implementation
    // (device: nvcc -cubin for sm_100a; emulator: g++ -shared); returns 0 and writes `out_path`, or != 0 with the
    // compiler's output in log[0 .. log_cap).
    int pe_b200_stream_supported(void);
    int pe_b200_stream_build(char const* source_path, char const* out_path, char const* csrc_dir, int J, int GL, char* log, size_t log_cap);
    int pe_b200_launch_stream(pe_b200_rrun const* run, void const* blob, size_t bytes, uint64_t key, uint32_t n_tiles, uint32_t stage_rows, void* stream);
    void pe_b200_stream_last_geometry(int* out3);  // warps per CTA, ring stages, shared memory per CTA of the last stream launch
    // ---- reduce-and-core path (csrc/pe_b200_frontal.cu, host/frontal.cpp): one huge linear DC circuit per instance ------
    // Values are lane-interleaved over the instances: v[id * B + instance].  The symbolic phase numbers the resistor graph's
    // edges (fill edges included), schedules the elimination of every node of degree <= 2 in levels of pairwise
    // non-adjacent nodes, and maps what is left (the core) to a dense matrix, nodes first, source branches last.
    typedef struct pe_b200_frontal
    {
        int32_t n_inst;
        int64_t B;  // lane stride of the value arrays (>= n_inst)
        int32_t n_nodes, n_res, n_idc, n_vdc, n_edges;
        double const* rval;       // [n_res][B] resistances
        int32_t const* res_a;     // [n_res] node unknown of pin A (-1 = ground)
        int32_t const* res_b;
        int32_t const* res_edge;  // [n_res] edge the resistor contributes to (-1: to ground / self loop)
        int32_t const* idc_p;     // [n_idc] node unknowns (+, -)
        int32_t const* idc_q;
        double const* idc_val;    // [n_idc][B]
        int32_t const* ops;       // [n_ops][6]: k, a, b, edge(k,a), edge(k,b), edge(a,b)   (-1 = absent)
        int32_t const* level_off_host;  // HOST pointer: [n_levels + 1] first op of every level
        int32_t n_levels;
        int32_t n_core, n_core_nodes, n_core_edges;
        int64_t ld_core;              // n_core rounded up to the LU block size (64)
        int32_t const* core_unknown;  // [n_core] unknown index of every core row (nodes, then branches)
        int32_t const* core_edges;    // [n_core_edges][3]: core row i, core row j, edge id
        int32_t const* vdc;           // [n_vdc][3]: unused, core row of node +, core row of node - (-1 = ground)
        double const* vdc_val;        // [n_vdc][B]
        double* d;   // [n_nodes][B] diagonals (work)
        double* z;   // [n_nodes][B] right-hand side of the node rows (work)
        double* g;   // [n_edges][B] edge conductances (work)
        double* M;   // [n_inst][ld_core * ld_core] dense core, column-major
        double* c;   // [n_inst][ld_core]
        double* x;   // solution rows x[unknown * LSx + instance]
        int64_t LSx;
        int32_t* status;  // [n_inst]
        double* phase_ms_host;  // HOST pointer, optional: [3] device time of reduce / core LU / substitutions (the call then waits for the stream)
    } pe_b200_frontal;
    int pe_b200_frontal_run(pe_b200_frontal const* f, void* stream, uint64_t* n_launches);
    char const* pe_b200_frontal_last_error(void);

    // largest dynamic shared memory (bytes) one CTA of the resident kernel may use on the current device
    size_t pe_b200_resident_smem_limit(void);
    char const* pe_b200_dev_last_error(void);
    // number of kernels this library has launched so far in this process (bench.py's gpu_launches evidence)
    uint64_t pe_b200_launch_count(void);      // solve-kernel launches
    uint64_t pe_b200_aux_launch_count(void);  // launches of the small helper kernels (status reduction)
    void pe_b200_timing_enable(int on);
    double pe_b200_timing_collect(void);

#ifdef __cplusplus
}
#endif
