// pe_b200_program.h — the POD seam between the C++23 host side (netlist -> symbolic analysis -> batch program)
// and the sm_100a kernels.  Plain C types only: nvcc cannot include any phy_engine / fast_io header
// (SURVEY.md probe table), so nothing here refers to them.
//
// Execution model: one *lane* = one independent circuit solve stream (a Monte-Carlo instance, or one
// (instance, frequency point) pair in an AC sweep).  Every lane interprets the same batch program; all per-lane
// doubles live in *lane-interleaved* HBM arrays  w[slot][lane]  so that a warp touching slot s reads 32 consecutive
// doubles (one fully coalesced 256-byte request).
//
// Three operand spaces:
//   CONST  cst[slot]                         values shared by every lane (broadcast parameters, folded constants)
//   INST   wi[slot * LSi + inst]             per-instance state: swept parameters, device state, real solution x
//   LANE   wl[slot * LSl + lane]             per-lane scratch: matrix/rhs values (LU in place), AC solution, omega
// with inst = lane / ppi (ppi = frequency points per instance; 1 outside AC sweeps).
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
        PE_SP_INST = 1,
        PE_SP_LANE = 2,
    };
#define PE_OPND(space, slot) ((((uint32_t)(space)) << 29) | ((uint32_t)(slot) & 0x1fffffffu))
#define PE_OPND_NEG 0x80000000u
#define PE_OPND_SPACE(o) (((o) >> 29) & 3u)
#define PE_OPND_SLOT(o) ((o) & 0x1fffffffu)

    // ---- opcodes; header word = opcode | (n << 8) -----------------------------------------------------------------
    // Operand lists follow the header as uint32 operand words.  "cx" = in a complex program the slot and slot+1 hold
    // (re, im).
    enum pe_b200_opcode
    {
        PE_OP_END = 0,

        // assembly (stamp gather): dst = sum_i (+/-) src_i, accumulated in list order = the reference's model order
        // (mna.h:60-157 accumulate semantics; "assign" stamps restart the list on the host side).  n = #src.
        PE_OP_ASM = 1,  // [dst][src * n]

        // LU on the fixed pattern, right-looking, augmented with the rhs column
        PE_OP_PIVOT = 2,  // [kk]                r = 1 / w[kk]; w[kk] = r                       (cx)
        PE_OP_ELIM = 3,   // [ik][(ij,kj) * n]   l = w[ik] * r; w[ij] -= l * w[kj]              (cx)
        // back substitution fused with the Newton convergence test (circuit.h:923-948).  flag bit 0 of n's top bit:
        // header = opcode | (n << 8) | (is_branch << 31)
        PE_OP_BACK = 4,  // [bk][kk][xk][(kj,xj) * n]  x = (w[bk] - sum w[kj]*x[xj]) * w[kk]      (cx)

        // scalar value ops (real)
        PE_OP_RECIP = 10,   // [dst][a]          dst = 1.0 / a
        PE_OP_MUL = 11,     // [dst][a][b]       dst = a * b
        PE_OP_SUB = 12,     // [dst][a][b]       dst = a - b
        PE_OP_COPY = 13,    // [dst][a]
        PE_OP_VSIN = 14,    // [dst][Vp][omega][phase]   dst = Vp * sin(omega * t + phase)      (VAC.h:176, IAC.h:154)
        PE_OP_SINCOS = 15,  // [dre][dim][Vp][phase]     dre = Vp cos(phase), dim = Vp sin(phase) (VAC.h:115-121)
        PE_OP_MUL2DIV = 16, // [dst][a][b]       dst = 2.0 * a / b                              (2C/dt, 2L/dt)

        // trapezoidal companions (capacitor.h:106-128, inductor.h:134-160)
        PE_OP_CAP_STEP = 20,  // [hist][prev_g][C][dt][va][vb]
        PE_OP_IND_STEP = 21,  // [req][ueq][L][dt][va][vb][ib]

        // PN junction (PN_junction.h)
        PE_OP_PN_PREP = 30,  // [is_eff][isr_eff][bv_eff][ut][uth] <- [Is][Isr][Area][N][Temp][Ibv][Bv][bv_set]
        PE_OP_PN_EVAL = 31,  // [ud_last][geq][ieq] <- [va][vb][is_eff][isr_eff][bv_eff][ut][uth][N][Nr][bv_set]
        PE_OP_PN_STEP = 32,  // [ud_last][hist][prev_g] <- [va][vb][geq][tt][dt]
        PE_OP_PN_ACCAP = 33, // [dst] <- [geq][tt][omega]     dst = (omega!=0 && tt>0 && geq>0 && tt*geq>0) ? tt*geq*omega : 0

        // BJT (BJT_NPN.h:116-159; PNP = same with the controlling voltage operands swapped)
        PE_OP_BJT_PREP = 40,  // [ut] <- [Temp]
        PE_OP_BJT_EVAL = 41,  // [geq][ieq_be][gm][ieq_c] <- [vp][vm][Is][Area][N][ut][BetaF]

        // level-1 MOSFET (nmosfet.h:84-141, pmosfet.h:84-138)
        PE_OP_NMOS_EVAL = 50,  // [gm][gds][ieq] <- [vd][vg][vs][Kp][lambda][Vth]
        PE_OP_PMOS_EVAL = 51,  // [gm][gds][ieq] <- [vd][vg][vs][Kp][lambda][Vth]
    };

    // ---- lane status ------------------------------------------------------------------------------------------------
    enum
    {
        PE_ST_OK = 0,
        PE_ST_NO_CONVERGENCE = 1,  // 64 Newton iterations without convergence (circuit.h:898,984)
        PE_ST_SINGULAR = 2,        // zero / non-finite pivot (circuit.h:1517 factorizationIsOk() == false)
    };

    // ---- one kernel launch = one analysis phase over all lanes -----------------------------------------------------
    typedef struct pe_b200_run
    {
        // device pointers
        uint32_t const* prep;  // once per launch (may be NULL)
        uint32_t const* step;  // once per time step, before the Newton loop (TR only; may be NULL)
        uint32_t const* iter;  // one linearised MNA solve: eval + assemble + factor + back-substitute(+check)
        double const* cst;
        double* wi;
        double* wl;
        int32_t* status;   // [n_lanes]  (in/out: lanes with status != 0 on entry are skipped)
        uint32_t* solves;  // [n_lanes]  += number of solve_once-equivalents executed
        double* wave;      // optional waveform store [n_steps][n_probe][LSl] (NULL = off)
        uint32_t const* probes;  // [n_probe] operand words

        int64_t LSi;  // lane stride (elements) of wi
        int64_t LSl;  // lane stride of wl
        int32_t n_lanes;
        int32_t ppi;  // lanes per instance
        int32_t cplx;       // 0 real program, 1 complex program (AC)
        int32_t nonlinear;  // 0: single solve per step; 1: Newton loop with convergence test
        int32_t max_iter;   // 64
        int32_t n_steps;    // >= 1
        int32_t n_probe;
        int32_t time_stepping;  // 1: run `step` section and advance t by dt before each step's solve
        double t0;              // tr_duration at entry
        double dt;
        double v_abstol, v_reltol, i_abstol, i_reltol;
    } pe_b200_run;

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
    char const* pe_b200_dev_last_error(void);
    // number of kernels this library has launched so far in this process (bench.py's gpu_launches evidence)
    uint64_t pe_b200_launch_count(void);
    void pe_b200_timing_enable(int on);
    double pe_b200_timing_collect(void);

#ifdef __cplusplus
}
#endif
