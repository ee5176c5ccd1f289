// pe_b200_kernels.cu — the sm_100a kernels of the batched MNA hot path + the POD device seam.
//
//   pe_b200_tree_kernel      tree-streaming kernel, the hot path of large circuits (configs B, D): the tree-scheduled
//                            program (pe_b200_program.h, "RESIDENT programs") with its workspace ws[row][lane] in HBM,
//                            CTA = S sub-tree warps x 32 J lanes, warp-cooperative word decode, persistent CTAs with
//                            dynamic (group, chunk) scheduling of the time loop.
//   pe_b200_resident_kernel  the same programs with the workspace in shared memory for the whole launch (small
//                            circuits, configs C, E: no HBM traffic inside the Newton / time loops).
//   pe_b200_solve_kernel     first-generation flat streaming interpreter (fallback for programs of > 32 768 rows).
//
// One launch runs a whole analysis phase per lane: [prep] -> for each time step { [step]; Newton loop { device
// evaluation + assembly + LU + substitution with the convergence test fused into the back substitution } }.  Replaces,
// per lane, the reference's circult::solve / solve_once / update_tr_step (circuit.h:363-374, 892-1527) and
// Eigen::SparseLU compute + solve.
//
// This TU includes no reference header (nvcc ICEs on fast_io; SURVEY.md probe table).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <unistd.h>

#include "pe_b200_interp.h"
#include "pe_b200_rinterp.h"

namespace
{
    using namespace pe_interp;

#ifndef PE_JIT
    __global__ void __launch_bounds__(32 * PE_MAX_WARPS, 2) pe_b200_solve_kernel(pe_b200_run const r)
    {
        // per-lane Newton flags of the current iteration, OR-ed in by the G warps (bit 0 not converged, bit 1 singular);
        // three buffers rotate so that the reset of the next one never races with readers of the previous one
        __shared__ uint32_t s_flags[3][32];

        int const g = (int)(threadIdx.x >> 5);
        int const li = (int)(threadIdx.x & 31u);
        int64_t const lane = (int64_t)blockIdx.x * 32 + li;  // < LSu (padding lanes run on scratch rows, never store)
        bool const real_lane = lane < r.n_lanes;
        int32_t status = real_lane ? r.status[lane] : (int32_t)PE_ST_SINGULAR;
        bool const counted = real_lane && status == PE_ST_OK;

        ctx_t c;
        c.cst = r.cst;
        c.wu = r.wu + lane;
        c.wx = r.wx + (real_lane ? lane / r.ppi : 0);
        c.LSu = r.LSu;
        c.LSx = r.LSx;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};

        if(threadIdx.x < 96) { (&s_flags[0][0])[threadIdx.x] = 0u; }
        __syncthreads();

        double t = r.t0;
        auto run_section = [&](pe_b200_section const& sec, bool live, bool check, bool& nconv, bool& fail)
        {
            if(sec.off[0] == PE_NO_SECTION) { return; }
            uint32_t const* pc = r.words + sec.off[g];
            while(run_until(pc, c, t, tol, live, check, nconv, fail) == R_BAR) { __syncthreads(); }
            __syncthreads();  // results of this section are visible to every warp of the CTA
        };

        bool ok = counted;
        uint32_t solves = 0;
        int fi = 0;
        {
            bool a = false, b = false;
            run_section(r.prep, ok, false, a, b);
        }
        for(int32_t s = 0; s < r.n_steps; ++s)
        {
            if(r.time_stepping)
            {
                // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                bool a = false, b = false;
                run_section(r.step, ok, false, a, b);
                t = t + r.dt;
            }
            bool done = !ok;
            int32_t it = 0;
            for(;;)
            {
                int const fn = fi == 2 ? 0 : fi + 1;
                if(g == 0) { s_flags[fn][li] = 0u; }
                bool nconv = false, fail = false;
                run_section(r.iter, !done, r.nonlinear != 0, nconv, fail);
                if(nconv || fail) { atomicOr(&s_flags[fi][li], (nconv ? 1u : 0u) | (fail ? 2u : 0u)); }
                __syncthreads();
                uint32_t const f = s_flags[fi][li];
                fi = fn;
                if(!done)
                {
                    ++solves;
                    if(f & 2u)
                    {
                        status = PE_ST_SINGULAR;
                        ok = false;
                        done = true;
                    }
                    else if(!r.nonlinear || !(f & 1u)) { done = true; }
                    else if(++it >= r.max_iter)
                    {
                        status = PE_ST_NO_CONVERGENCE;
                        ok = false;
                        done = true;
                    }
                }
                if(__syncthreads_and(done ? 1 : 0)) { break; }
            }
            if(r.wave != nullptr && ok && g == 0)
            {
                for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane] = ld(c, __ldg(r.probes + p)); }
            }
        }
        if(g == 0 && counted)
        {
            r.status[lane] = status;
            r.solves[lane] += solves;
        }
    }

    // ---- resident kernel ---------------------------------------------------------------------------
    // CTA = S streams x I instances, workspace ws[slot][I] in dynamic shared memory for the whole launch.  Thread
    // (stream, ig) executes the vector ops of its warp for the J instances [ig * J, ig * J + J) of the CTA.
    // Warp-cooperative reader of a warp's program (same interface as pe_rinterp::host_reader).
    // MAIN stream (headers, masks, warp-uniform rows): three consecutive 128-byte lines are held in registers, lane l
    // keeping word l of each; a word is handed to the whole warp with one shuffle, and the line after next is already in
    // flight while the current one is consumed, so L2 latency never sits on the critical path of an op.
    // SIDE stream (per-column rows, C words each): every lane reads the word of its own column (one coalesced line per
    // row); the next four rows are kept in flight in a register queue.
#endif  // PE_JIT

    struct warp_reader
    {
        uint32_t const* main;
        uint32_t const* side;
        uint32_t w0, w1, w2, line, pos;
        uint32_t q0, q1, q2, q3, srow;
        uint32_t C, col, lane;
        uint32_t cur, m, pos_next;

        __device__ __forceinline__ void init(uint32_t const* main_, uint32_t const* side_, uint32_t C_, uint32_t col_, uint32_t lane_)
        {
            main = main_;
            side = side_;
            C = C_;
            col = col_;
            lane = lane_;
            line = 0;
            pos = C_ == 1u ? 2u : 0u;  // one stream per warp: every line starts with two prefetch bitmaps
            w0 = __ldg(main + lane);
            w1 = __ldg(main + 32 + lane);
            w2 = __ldg(main + 64 + lane);
            srow = 0;
            q0 = __ldg(side + col);
            q1 = __ldg(side + C + col);
            q2 = __ldg(side + 2 * C + col);
            q3 = __ldg(side + 3 * C + col);
        }
        __device__ __forceinline__ uint32_t peek(uint32_t a) const
        {
            uint32_t const l = (a >> 5) - line;
            uint32_t const v = l == 0u ? w0 : (l == 1u ? w1 : w2);
            return __shfl_sync(0xffffffffu, v, (int)(a & 31u));
        }
        __device__ __forceinline__ uint32_t head() const { return peek(pos); }
        __device__ __forceinline__ uint32_t open(uint32_t rows)
        {
            m = peek(pos + 1u);
            cur = pos + 2u;
            pos_next = pos + 2u + rows - (uint32_t)__popc(m);
            return m;
        }
        __device__ __forceinline__ uint32_t next()
        {
            uint32_t w;
            if(m & 1u)
            {
                w = q0;
                q0 = q1;
                q1 = q2;
                q2 = q3;
                q3 = __ldg(side + (srow + 4u) * C + col);
                ++srow;
            }
            else
            {
                w = peek(cur);
                ++cur;
            }
            m >>= 1;
            return w;
        }
        __device__ __forceinline__ void advance(uint32_t np)
        {
            if(C == 1u && (np & 31u) == 0u) { np += 2u; }
            pos = np;
            while((pos >> 5) > line)
            {
                w0 = w1;
                w1 = w2;
                ++line;
                w2 = __ldg(main + (line + 2u) * 32u + lane);
            }
        }
        __device__ __forceinline__ void close() { advance(pos_next); }
        __device__ __forceinline__ void bar() { advance(pos + 1u); }
        __device__ __forceinline__ void skip() { advance(((pos >> 5) + 1u) << 5); }
    };

#ifndef PE_JIT
    template <int J, int MAXT>
    __global__ void __launch_bounds__(MAXT, 1) pe_b200_resident_kernel(pe_b200_rrun const r)
    {
        extern __shared__ __align__(16) double ws[];
        __shared__ uint32_t s_flags[3][32];

        using namespace pe_rinterp;
        uint32_t const I = (uint32_t)r.I, IG = I / J, S = (uint32_t)r.S;
        uint32_t const tid = threadIdx.x;
        uint32_t const ig = tid % IG, stream = tid / IG;
        uint32_t const warp = tid >> 5, n_warps = blockDim.x >> 5;
        int64_t const lane0 = (int64_t)blockIdx.x * I + ig * J;

        rctx c;
        c.ws = ws + ig * J;
        c.I = (uint64_t)I;
        c.S = S;
        c.C = 32u / IG;
        c.col = (tid & 31u) / IG;
        c.stream = stream;
        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};

        bool real_lane[J], counted[J], ok[J];
        int32_t status[J];
        uint32_t solves[J];
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            real_lane[j] = lane0 + j < r.n_lanes;
            status[j] = real_lane[j] ? r.status[lane0 + j] : (int32_t)PE_ST_SINGULAR;
            counted[j] = real_lane[j] && status[j] == PE_ST_OK;
            ok[j] = counted[j];
            solves[j] = 0;
        }

        // ---- load: persistent values, broadcast constants, per-instance values seen from a frequency-point lane
        for(uint32_t e = stream; e < (uint32_t)r.n_io; e += S)
        {
            pe_b200_io const io = r.io[e];
            uint32_t const fl = io.slot_kind >> 20;
            if(!(fl & PE_IO_LOAD)) { continue; }
            uint32_t const kind = (io.slot_kind >> 16) & 0xfu;
            double* dst = c.ws + (uint64_t)(io.slot_kind & 0xffffu) * c.I;
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                int64_t const lane = lane0 + j;
                double v;
                if(kind == PE_IO_CONST) { v = r.cst[io.src]; }
                else if(kind == PE_IO_U) { v = r.wu[(int64_t)io.src * r.LSu + lane]; }
                else { v = r.wx[(int64_t)io.src * r.LSx + (real_lane[j] ? lane / r.ppi : 0)]; }
                dst[j] = v;
            }
        }
        if(tid < 96) { (&s_flags[0][0])[tid] = 0u; }
        __syncthreads();

        double t = r.t0;
        auto run_section = [&](int sec, bool const (&en)[J], bool check, bool (&nconv)[J], bool (&fail)[J], bool first_iter)
        {
            warp_reader rd;
            rd.init(r.words + __ldg(r.sec_off + sec * n_warps + warp), r.words + __ldg(r.sec_off + (3 + sec) * n_warps + warp), c.C, c.col, tid & 31u);
            for(;;)
            {
                int const k = rvop<J>(rd, c, t, tol, en, check, nconv, fail, first_iter);
                if(k == V_END || k == V_BAD) { break; }
                if(k == V_BAR)
                {
                    __syncthreads();
                    rd.bar();
                }
                else if(k == V_SKIP) { rd.skip(); }
                else
                {
                    rd.close();
                }
            }
            __syncthreads();  // results of this section are visible to every stream of the CTA
        };

        int fi = 0;
        if(r.has_prep)
        {
            bool a[J], b[J];
#pragma unroll
            for(int j = 0; j < J; ++j) { a[j] = b[j] = false; }
            run_section(0, ok, false, a, b, true);
        }
        for(int32_t s = 0; s < r.n_steps; ++s)
        {
            if(r.time_stepping)
            {
                // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                if(r.has_step)
                {
                    bool a[J], b[J];
#pragma unroll
                    for(int j = 0; j < J; ++j) { a[j] = b[j] = false; }
                    run_section(1, ok, false, a, b, true);
                }
                t = t + r.dt;
            }
            bool done[J];
#pragma unroll
            for(int j = 0; j < J; ++j) { done[j] = !ok[j]; }
            int32_t it = 0;
            for(;;)
            {
                int const fn = fi == 2 ? 0 : fi + 1;
                if(stream == 0)
                {
#pragma unroll
                    for(int j = 0; j < J; ++j) { s_flags[fn][ig * J + j] = 0u; }
                }
                bool nconv[J], fail[J], en[J];
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    nconv[j] = fail[j] = false;
                    en[j] = !done[j];
                }
                run_section(2, en, r.nonlinear != 0, nconv, fail, it == 0);
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(nconv[j] || fail[j]) { atomicOr(&s_flags[fi][ig * J + j], (nconv[j] ? 1u : 0u) | (fail[j] ? 2u : 0u)); }
                }
                __syncthreads();
                bool all_done = true;
                ++it;
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    uint32_t const f = s_flags[fi][ig * J + j];
                    if(!done[j])
                    {
                        ++solves[j];
                        if(f & 2u)
                        {
                            status[j] = PE_ST_SINGULAR;
                            ok[j] = false;
                            done[j] = true;
                        }
                        else if(!r.nonlinear || !(f & 1u)) { done[j] = true; }
                        else if(it >= r.max_iter)
                        {
                            status[j] = PE_ST_NO_CONVERGENCE;
                            ok[j] = false;
                            done[j] = true;
                        }
                    }
                    all_done = all_done && done[j];
                }
                fi = fn;
                if(__syncthreads_and(all_done ? 1 : 0)) { break; }
            }
            if(r.wave != nullptr && stream == 0)
            {
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(!ok[j]) { continue; }
                    for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + lane0 + j] = c.ws[(uint64_t)__ldg(r.probes + p) * c.I + j]; }
                }
            }
        }
        // ---- store the mutable persistent values back
        for(uint32_t e = stream; e < (uint32_t)r.n_io; e += S)
        {
            pe_b200_io const io = r.io[e];
            if(!((io.slot_kind >> 20) & PE_IO_STORE)) { continue; }
            double const* src = c.ws + (uint64_t)(io.slot_kind & 0xffffu) * c.I;
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(counted[j]) { r.wu[(int64_t)io.src * r.LSu + lane0 + j] = src[j]; }
            }
        }
        if(stream == 0)
        {
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(counted[j])
                {
                    r.status[lane0 + j] = status[j];
                    r.solves[lane0 + j] += solves[j];
                }
            }
        }
    }

#endif  // PE_JIT

    // ---- tree-streaming kernel: the tree-scheduled program with its workspace in HBM ---------------
    // CTA = S warps x 32 lanes: warp s runs word stream s (a sub-tree of the elimination tree) for the CTA's 32
    // instances, thread = one instance.  ws[slot][lane] is lane-interleaved in HBM (one coalesced 256-byte request per
    // warp access).  The program words are warp-uniform, so the warp decodes them cooperatively: lane l holds word l of
    // the current 32-word line (and of the next one, already in flight), pre-decodes both 16-bit operand fields of it
    // once, and an operand costs the whole warp one shuffle + one address multiply-add + the load.
    struct line_reader
    {
        uint32_t const* base;
        uint32_t lane, stream, sm;
        uint32_t w, nw, n2w;  // this lane's raw word of the current line and of the next two (in flight)
        uint32_t dlo, dhi;    // the two operand fields of w decoded: absolute slot | neg << 31
        uint32_t line, off;   // current line, offset of the current op inside it
        // generic-reader state (ops the fast path does not take)
        uint32_t cur, m;
        bool fresh;  // a line was taken since the caller last looked (it then prefetches the operands of the next one)

        __device__ __forceinline__ uint32_t dec(uint32_t f) const
        {
            uint32_t const r = f & 0x7fffu;
            return ((r & ~sm) | ((r + stream) & sm)) | ((f & 0x8000u) << 16);
        }
        __device__ __forceinline__ void decode()
        {
            dlo = dec(w & 0xffffu);
            dhi = dec(w >> 16);
        }
        // Every 32-word line starts with two bitmaps: bit l of word 0 (word 1) says that the low (high) half of word l
        // names a workspace row that is cold by the time it is used.  Lane l asks L2 for the rows of its word of the
        // NEXT line (a row = this group's 32 J lanes = 2 J 128-byte lines; complex values: the imaginary row too), a
        // line's worth of ops before they are used.
        template <int J>
        __device__ __forceinline__ void prefetch_next(char const* group_base, uint32_t LS8, uint32_t S, bool cplx, bool far) const
        {
            uint32_t const pw = far ? n2w : nw;  // one or two lines ahead
            uint32_t const bl = __shfl_sync(0xffffffffu, pw, 0), bh = __shfl_sync(0xffffffffu, pw, 1);
            uint32_t const f[2] = {pw & 0xffffu, pw >> 16};
            uint32_t const on[2] = {(bl >> lane) & 1u, (bh >> lane) & 1u};
#pragma unroll
            for(int h = 0; h < 2; ++h)
            {
                if(!on[h]) { continue; }
                char const* p = group_base + (uint64_t)(dec(f[h]) & 0x7fffu) * LS8;
#pragma unroll
                for(int k = 0; k < 2 * J; ++k) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128 * k)); }
                if(cplx)
                {
                    p += (uint64_t)S * LS8;
#pragma unroll
                    for(int k = 0; k < 2 * J; ++k) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128 * k)); }
                }
            }
        }
        __device__ __forceinline__ void init(uint32_t const* b, uint32_t lane_, uint32_t stream_, uint32_t S)
        {
            base = b;
            lane = lane_;
            stream = stream_;
            sm = S - 1u;
            line = 0;
            off = 2;
            w = __ldg(base + lane);
            nw = __ldg(base + 32 + lane);
            n2w = __ldg(base + 64 + lane);
            decode();
            fresh = true;
        }
        __device__ __forceinline__ uint32_t raw(uint32_t k) const { return __shfl_sync(0xffffffffu, w, (int)(off + k)); }
        __device__ __forceinline__ uint32_t lo(uint32_t k) const { return __shfl_sync(0xffffffffu, dlo, (int)(off + k)); }
        __device__ __forceinline__ uint32_t hi(uint32_t k) const { return __shfl_sync(0xffffffffu, dhi, (int)(off + k)); }
        __device__ __forceinline__ void next_line()
        {
            w = nw;
            nw = n2w;
            ++line;
            off = 2;
            n2w = __ldg(base + (line + 2u) * 32u + lane);
            decode();
            fresh = true;
        }
        __device__ __forceinline__ void adv(uint32_t n)
        {
            off += n;
            if(off >= 32u) { next_line(); }
        }
        // pe_rinterp reader interface (mask is always 0: one stream per warp has no per-column rows)
        __device__ __forceinline__ uint32_t head() const { return raw(0); }
        __device__ __forceinline__ uint32_t open(uint32_t rows)
        {
            cur = 2u;
            m = rows;
            return 0u;
        }
        __device__ __forceinline__ uint32_t next() { return raw(cur++); }
        __device__ __forceinline__ void close() { adv(2u + m); }
        __device__ __forceinline__ void bar() { adv(1u); }
        __device__ __forceinline__ void skip() { next_line(); }
    };

    // address of slot `slot` of this thread's lane in the HBM workspace (by-value functor: lives in registers)
    struct lane_ws
    {
        char* wl;
        uint32_t LS8;
        __device__ __forceinline__ double* operator() (uint32_t slot) const { return reinterpret_cast<double*>(wl + (uint64_t)slot * LS8); }
    };

    // one DOT with NA rows of packed sources and NB pairs: every operand load is issued before the first value is used
    // one DOT with NA rows of packed sources and NB pairs for the J instances of this thread (lanes 32 apart: 256 bytes
    // apart in a workspace row): every operand load is issued before the first value is used
    template <int J, int NA, int NB>
    __device__ __forceinline__ void tree_dot(line_reader const& rd, lane_ws const& at, pe_b200_rrun const& r, bool const (&en)[J], bool check, bool (&nconv)[J], bool (&fail)[J])
    {
        // words of the op: [h][mask][ctl][scale][src x NA][pair x NB]
        uint32_t const flags = rd.raw(2) >> 16;
        uint32_t const dst = rd.lo(2) & 0x7fffu;
        uint32_t g[2 * NA + 1];
        double sv[2 * NA + 1][J], av[NB + 1][J], bv[NB + 1][J];
#pragma unroll
        for(int i = 0; i < NA; ++i)
        {
            g[2 * i] = rd.lo(4 + i);
            g[2 * i + 1] = rd.hi(4 + i);
            double const* p0 = at(g[2 * i] & 0x7fffu);
            double const* p1 = at(g[2 * i + 1] & 0x7fffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                sv[2 * i][j] = p0[32 * j];
                sv[2 * i + 1][j] = p1[32 * j];
            }
        }
#pragma unroll
        uint32_t gp[NB + 1];  // sign of the product of pair i (aliased entries carry a negate bit)
#pragma unroll
        for(int i = 0; i < NB; ++i)
        {
            uint32_t const ga = rd.lo(4 + NA + i), gb = rd.hi(4 + NA + i);
            gp[i] = (ga ^ gb) & 0x80000000u;
            double const* pa = at(ga & 0x7fffu);
            double const* pb = at(gb & 0x7fffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                av[i][j] = pa[32 * j];
                bv[i][j] = pb[32 * j];
            }
        }
        double sc[J], xo[J];
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            sc[j] = 1.0;
            xo[j] = 0.0;
        }
        if(flags & PE_F_SCALE)
        {
            double const* ps = at(rd.lo(3) & 0x7fffu);
#pragma unroll
            for(int j = 0; j < J; ++j) { sc[j] = ps[32 * j]; }
        }
        bool const chk = check && (flags & (PE_F_CHECK_V | PE_F_CHECK_I));
        double* const pd = at(dst);
        if(chk)
        {
#pragma unroll
            for(int j = 0; j < J; ++j) { xo[j] = pd[32 * j]; }
        }
        double acc[J];
#pragma unroll
        for(int j = 0; j < J; ++j) { acc[j] = 0.0; }
        // acc +/- s as fma(s, +/-1.0, acc): one rounding, the same value as the add, no select on the operand
#pragma unroll
        for(int i = 0; i < 2 * NA; ++i)
        {
            double const sg = __hiloint2double((int)(0x3ff00000u | (g[i] & 0x80000000u)), 0);
#pragma unroll
            for(int j = 0; j < J; ++j) { acc[j] = fma(sv[i][j], sg, acc[j]); }
        }
#pragma unroll
        for(int i = 0; i < NB; ++i)
        {
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                // -(+/-a)(+/-b): flipping the sign bit of a is exact
                double const a = __hiloint2double(__double2hiint(av[i][j]) ^ (int)(0x80000000u ^ gp[i]), __double2loint(av[i][j]));
                acc[j] = fma(a, bv[i][j], acc[j]);
            }
        }
        if(flags & PE_F_SCALE)
        {
#pragma unroll
            for(int j = 0; j < J; ++j) { acc[j] = PE_MUL(acc[j], sc[j]); }
        }
        if(flags & PE_F_GUARD)
        {
            // an entry of L: out of bounds when the pivot is far smaller than this entry of its column (pe_b200_program.h)
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(PE_GUARD_TRIP(fabs(acc[j]), r.guard)) { fail[j] = true; }
            }
        }
        if(flags & PE_F_RECIP)
        {
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(acc[j] == 0.0 || !isfinite(acc[j])) { fail[j] = true; }
                acc[j] = PE_RCP(acc[j]);
            }
        }
        if(chk)
        {
            // circuit.h:923-948: |new - old| > abstol + reltol * max(|new|, |old|)  => not converged
            bool const br = (flags & PE_F_CHECK_I) != 0u;
            double const at_ = br ? r.i_abstol : r.v_abstol, rt_ = br ? r.i_reltol : r.v_reltol;
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                double const tl = at_ + rt_ * fmax(fabs(acc[j]), fabs(xo[j]));
                if(fabs(acc[j] - xo[j]) > tl) { nconv[j] = true; }
            }
        }
#pragma unroll
        for(int j = 0; j < J; ++j)
        {
            if(en[j]) { pd[32 * j] = acc[j]; }
        }
        // Stores write through to L2 and do not allocate in L1, while the value just produced is what the next few ops
        // read: ask for the line back right away, so that they find it in L1 (~40 clk) instead of L2 (~300 clk).
        if(r.prefetch & 4)
        {
#pragma unroll
            for(int j = 0; j < J; ++j) { asm volatile("prefetch.global.L1 [%0];" ::"l"(pd + 32 * j)); }
        }
    }

    // fused elimination step (PE_OP_CROUT2): six DOT slots [pivot | U entries, L entries, rhs entry]; all 26 operand
    // rows of the step are requested before the first is used, the pivot reciprocal stays in a register
    template <int J>
    __device__ __forceinline__ void tree_crout2(line_reader const& rd, lane_ws const& at, double guard, bool const (&en)[J], bool (&fail)[J])
    {
        // words: [h][mask] then slot 0: [ctl][src][src][pair], slots 1..5: [ctl][src][pair]
        uint32_t gs[14], gp[6];
        double sv[14][J], pa[6][J], pb[6][J];
#pragma unroll
        for(int q = 0; q < 6; ++q)
        {
            int const base = q == 0 ? 0 : 4 + 3 * (q - 1);
            int const nsr = q == 0 ? 2 : 1;
            int const s0 = q == 0 ? 0 : 4 + 2 * (q - 1);
#pragma unroll
            for(int r = 0; r < nsr; ++r)
            {
                gs[s0 + 2 * r] = rd.lo(2 + base + 1 + r);
                gs[s0 + 2 * r + 1] = rd.hi(2 + base + 1 + r);
                double const* p0 = at(gs[s0 + 2 * r] & 0x7fffu);
                double const* p1 = at(gs[s0 + 2 * r + 1] & 0x7fffu);
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    sv[s0 + 2 * r][j] = p0[32 * j];
                    sv[s0 + 2 * r + 1][j] = p1[32 * j];
                }
            }
            uint32_t const ga = rd.lo(2 + base + 1 + nsr), gb = rd.hi(2 + base + 1 + nsr);
            gp[q] = (ga ^ gb) & 0x80000000u;
            double const* a = at(ga & 0x7fffu);
            double const* b = at(gb & 0x7fffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                pa[q][j] = a[32 * j];
                pb[q][j] = b[32 * j];
            }
        }
        double piv[J];
#pragma unroll
        for(int j = 0; j < J; ++j) { piv[j] = 0.0; }
#pragma unroll
        for(int q = 0; q < 6; ++q)
        {
            int const base = q == 0 ? 0 : 4 + 3 * (q - 1);
            int const nsr = q == 0 ? 2 : 1;
            int const s0 = q == 0 ? 0 : 4 + 2 * (q - 1);
            uint32_t const ctl = rd.raw(2 + base);
            if(!(ctl & PE_R_ACTIVE)) { continue; }
            uint32_t const flags = ctl >> 16;
            double acc[J];
#pragma unroll
            for(int j = 0; j < J; ++j) { acc[j] = 0.0; }
#pragma unroll
            for(int i = 0; i < 2 * nsr; ++i)
            {
                double const sg = __hiloint2double((int)(0x3ff00000u | (gs[s0 + i] & 0x80000000u)), 0);
#pragma unroll
                for(int j = 0; j < J; ++j) { acc[j] = fma(sv[s0 + i][j], sg, acc[j]); }
            }
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                double const a = __hiloint2double(__double2hiint(pa[q][j]) ^ (int)(0x80000000u ^ gp[q]), __double2loint(pa[q][j]));
                acc[j] = fma(a, pb[q][j], acc[j]);
            }
            if(flags & PE_F_SCALE)
            {
#pragma unroll
                for(int j = 0; j < J; ++j) { acc[j] = PE_MUL(acc[j], piv[j]); }
            }
            if(flags & PE_F_GUARD)
            {
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(PE_GUARD_TRIP(fabs(acc[j]), guard)) { fail[j] = true; }
                }
            }
            if(flags & PE_F_RECIP)
            {
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(acc[j] == 0.0 || !isfinite(acc[j])) { fail[j] = true; }
                    acc[j] = PE_RCP(acc[j]);
                    piv[j] = acc[j];
                }
            }
            double* const pd = at(rd.lo(2 + base) & 0x7fffu);
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                if(en[j]) { pd[32 * j] = acc[j]; }
            }
        }
    }

    // CTA = S warps x (32 x J) lanes: thread (warp s, lane l) runs stream s for lanes l, l + 32, ... of the group
    // FUSED: the program may hold fused elimination steps (PE_OP_CROUT2); the code for them is kept out of the other
    // instances, whose register budget it would blow
    // CL: CTAs per lane group.  CL = 2: a thread-block cluster of two CTAs (two SMs) shares a group, each CTA running half
    // of its S sub-tree warps; the barriers of the program are cluster barriers (release / acquire at cluster scope), the
    // per-lane Newton flags of the two CTAs are combined through distributed shared memory.  That is what lets 128-lane
    // groups (1 KB workspace rows, twice the DRAM efficiency of 512 B rows) still cover all 148 SMs with 10 000 lanes.
    template <int CL>
    __device__ __forceinline__ void group_sync()
    {
        if constexpr(CL == 1) { __syncthreads(); }
        else
        {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }

#ifdef PE_JIT
    // ---- run-time specialisation (host/jit.cpp): this file is compiled again by nvcc at run time with -DPE_JIT; the
    // generated source (PE_JIT_SOURCE) holds the iter section of one tree-scheduled program as straight-line code --
    // one function per (sub-tree stream, phase), operand rows addressed by constants, results of the last few ops
    // forwarded in registers, the loads of an op issued PE_JIT_D ops ahead -- and replaces the word interpreter for that
    // section.  Everything around it (scheduling, time / Newton loops, flags, I/O) is the code of the tree kernel below.
    struct jv
    {
        double v[4];
    };
    __device__ __forceinline__ char* jopq(char* p) { return p; }
    __device__ __forceinline__ jv jld(char const* p)
    {
        jv x;
#pragma unroll
        for(int j = 0; j < 4; ++j) { x.v[j] = *reinterpret_cast<double const*>(p + 256 * j); }
        return x;
    }
    __device__ __forceinline__ void jst(char* p, jv const& x, uint32_t enm)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j)
        {
            if((enm >> j) & 1u) { *reinterpret_cast<double*>(p + 256 * j) = x.v[j]; }
        }
    }
    __device__ __forceinline__ jv jzero()
    {
        jv x;
#pragma unroll
        for(int j = 0; j < 4; ++j) { x.v[j] = 0.0; }
        return x;
    }
    // acc +/- s  (the interpreter's fma(s, +/-1.0, acc): the same single rounding)
    __device__ __forceinline__ void jadd(jv& acc, jv const& s)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { acc.v[j] = __dadd_rn(acc.v[j], s.v[j]); }
    }
    __device__ __forceinline__ void jsub(jv& acc, jv const& s)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { acc.v[j] = __dsub_rn(acc.v[j], s.v[j]); }
    }
    // acc -/+ a b
    __device__ __forceinline__ void jfms(jv& acc, jv const& a, jv const& b)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { acc.v[j] = __fma_rn(-a.v[j], b.v[j], acc.v[j]); }
    }
    __device__ __forceinline__ void jfma(jv& acc, jv const& a, jv const& b)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { acc.v[j] = __fma_rn(a.v[j], b.v[j], acc.v[j]); }
    }
    __device__ __forceinline__ void jmul(jv& acc, jv const& s)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { acc.v[j] = PE_MUL(acc.v[j], s.v[j]); }
    }
    __device__ __forceinline__ void jrcp(jv& acc, uint32_t& failm)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j)
        {
            if(acc.v[j] == 0.0 || !isfinite(acc.v[j])) { failm |= 1u << j; }
            acc.v[j] = PE_RCP(acc.v[j]);
        }
    }
    // pivot guard (pe_b200_program.h PE_F_GUARD): an entry of L out of bounds marks the lane
    __device__ __forceinline__ void jguard(jv const& l, double const guard, uint32_t& failm)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j)
        {
            if(PE_GUARD_TRIP(fabs(l.v[j]), guard)) { failm |= 1u << j; }
        }
    }
    // CAP_STEP (capacitor.h:106-128): hist, prev_g updated in place
    __device__ __forceinline__ void jcap(jv const& C, jv const& dt, jv const& va, jv const& vb, jv& hist, jv& prev_g)
    {
#pragma unroll
        for(int j = 0; j < 4; ++j) { pe_models::cap_step(C.v[j], dt.v[j], PE_SUB(va.v[j], vb.v[j]), hist.v[j], prev_g.v[j]); }
    }
#include PE_JIT_SOURCE
#endif  // PE_JIT

    template <int J, bool FUSED, int CL>
    __device__ __forceinline__ void tree_body(pe_b200_rrun const& r)
    {
        __shared__ uint32_t s_flags[3][32 * J];
        __shared__ uint32_t s_item;
        using namespace pe_rinterp;
        uint32_t const tid = threadIdx.x, lane = tid & 31u, n_warps_cta = blockDim.x >> 5;
        uint32_t const rank = CL == 1 ? 0u : (uint32_t)cooperative_groups::this_cluster().block_rank();
        // warp = index of this warp among the S sub-tree warps of the group (over the CTAs of the cluster)
        uint32_t const warp = rank * n_warps_cta + (tid >> 5), n_warps = n_warps_cta * CL;
        uint32_t const first_block = blockIdx.x / CL;  // static assignment: one group per cluster
        // the other CTA's shared memory (CL = 2)
        uint32_t const* peer_flags = &s_flags[0][0];
        uint32_t const* lead_item = &s_item;
        if constexpr(CL == 2)
        {
            peer_flags = cooperative_groups::this_cluster().map_shared_rank(&s_flags[0][0], rank ^ 1u);
            lead_item = cooperative_groups::this_cluster().map_shared_rank(&s_item, 0u);
        }
        uint32_t const S = (uint32_t)r.S;
        uint32_t const GL = 32u * J;  // lanes per group
        uint32_t const NG = (uint32_t)((r.n_lanes + GL - 1) / GL), NC = r.sched != nullptr ? (uint32_t)r.n_chunks : 1u;
        for(;;)
        {
            // ---- next work item: (chunk c of the time loop, lane group g)
            uint32_t item = first_block;
            if(r.sched != nullptr)
            {
                if(tid == 0 && rank == 0u) { s_item = atomicAdd(r.sched, 1u); }
                group_sync<CL>();
                item = *lead_item;
                if constexpr(CL == 2) { group_sync<CL>(); }  // the leader may overwrite s_item only after both CTAs have read it
            }
            if(item >= NC * NG) { break; }
            uint32_t const chunk = item / NG, group = item - chunk * NG;
            if(r.sched != nullptr && chunk > 0u)
            {
                // chunk c of a group starts when its chunk c - 1 (run by some other CTA) has published its results; the
                // acquire load also drops whatever this SM's L1 still holds of the group's workspace
                if(tid == 0)  // thread 0 of every CTA of the cluster: each SM drops its own L1
                {
                    uint32_t v;
                    for(;;)
                    {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(r.sched + 1 + group) : "memory");
                        if(v >= chunk) { break; }
                        __nanosleep(256);
                    }
                }
                __syncthreads();
            }
            int64_t const glane = (int64_t)group * GL + lane;  // first of this thread's lanes; the others are + 32 j
#ifdef PE_JIT
            // specialised kernel: the workspace of a lane group is one block, ws[group][slot][32 J lanes] (1 KB rows whose
            // addresses are a column base plus a load / store immediate, host/jit.cpp)
            lane_ws const at{reinterpret_cast<char*>(r.wsg + (int64_t)group * r.n_slots * GL + lane), GL * 8u};
#else
            lane_ws const at{reinterpret_cast<char*>(r.wsg + glane), (uint32_t)(r.LSw * 8)};
#endif
            bool const first_chunk = chunk == 0u, last_chunk = chunk + 1u == NC;
            int32_t const s_begin = r.sched != nullptr ? (int32_t)chunk * r.chunk_steps : 0;
            int32_t const s_end = r.sched != nullptr ? min(r.n_steps, s_begin + r.chunk_steps) : r.n_steps;

            bool real_lane[J], counted[J], ok[J], done[J];
            int32_t status[J];
            uint32_t solves[J];
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                real_lane[j] = glane + 32 * j < r.n_lanes;
                status[j] = real_lane[j] ? r.status[glane + 32 * j] : (int32_t)PE_ST_SINGULAR;
                counted[j] = real_lane[j] && status[j] == PE_ST_OK;
                ok[j] = counted[j];
                done[j] = false;
                solves[j] = 0;
            }

            for(uint32_t e = warp; first_chunk && e < (uint32_t)r.n_io; e += n_warps)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_LOAD)) { continue; }
                uint32_t const kind = (io.slot_kind >> 16) & 0xfu;
                double* const dst = at(io.slot_kind & 0xffffu);
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    double v = 0.0;
                    if(kind == PE_IO_CONST) { v = r.cst[io.src]; }
                    else if(real_lane[j])
                    {
                        v = kind == PE_IO_U ? r.wu[(int64_t)io.src * r.LSu + glane + 32 * j] : r.wx[(int64_t)io.src * r.LSx + (glane + 32 * j) / r.ppi];
                    }
                    dst[32 * j] = v;
                }
            }
            if(tid < 96 * J) { (&s_flags[0][0])[tid] = 0u; }
            group_sync<CL>();

            // One copy of the interpreter serves the three sections: stage 0 = prep, 1 = the step section of time step
            // s, 2 = one Newton iteration (the iter section).  The sequencing is uniform over the CTA.
            double t = r.sched != nullptr ? r.t_chunk[chunk] : r.t0;
            int32_t s = s_begin, it = 0;
            int stage = (r.has_prep && first_chunk) ? 0 : 1;
            int fi = 0;
            for(;;)
            {
                int sec;
                bool en[J], check;
                if(stage == 0)
                {
                    sec = 0;
                    check = false;
#pragma unroll
                    for(int j = 0; j < J; ++j) { en[j] = ok[j]; }
                }
                else if(stage == 1)
                {
                    if(s >= s_end) { break; }
                    if(!(r.time_stepping && r.has_step))
                    {
                        if(r.time_stepping) { t = t + r.dt; }
#pragma unroll
                        for(int j = 0; j < J; ++j) { done[j] = !ok[j]; }
                        it = 0;
                        stage = 2;
                        continue;
                    }
                    // update_tr_step(dt) then tr_duration = prev + dt  (circuit.h:243-248)
                    sec = 1;
                    check = false;
#pragma unroll
                    for(int j = 0; j < J; ++j) { en[j] = ok[j]; }
                }
                else
                {
                    sec = 2;
                    check = r.nonlinear != 0;
#pragma unroll
                    for(int j = 0; j < J; ++j) { en[j] = !done[j]; }
                    if((tid >> 5) == 0)  // first warp of every CTA: its own flag buffer
                    {
#pragma unroll
                        for(int j = 0; j < J; ++j) { s_flags[fi == 2 ? 0 : fi + 1][lane + 32 * j] = 0u; }
                    }
                }

                // ---- run the section
                bool nconv[J], fail[J];
#pragma unroll
                for(int j = 0; j < J; ++j) { nconv[j] = fail[j] = false; }
#ifdef PE_JIT
                if(sec == 2)
                {
                    pe_jit_iter<CL>(warp, at.wl, r.guard, en, fail);
                    group_sync<CL>();
                }
                else
#endif
                {
                    line_reader rd;
                    rd.init(r.words + __ldg(r.sec_off + sec * n_warps + warp), lane, warp, S);
                    for(;;)
                    {
                        if(rd.fresh)
                        {
                            rd.fresh = false;
                            if(r.prefetch & 1) { rd.template prefetch_next<J>(at.wl - lane * 8u, at.LS8, S, r.cplx != 0, (r.prefetch & 2) != 0); }
                        }
                        uint32_t const h = rd.raw(0);
                        uint32_t const op = h & 0xffu;
                        uint32_t const na = (h >> 8) & 0x1fu, nb = (h >> 18) & 0x3fu;
                        if(op == PE_OP_DOT && na <= 2u && nb <= 3u)
                        {
                            switch(na * 4u + nb)
                            {
                                case 0: tree_dot<J, 0, 0>(rd, at, r, en, check, nconv, fail); break;
                                case 1: tree_dot<J, 0, 1>(rd, at, r, en, check, nconv, fail); break;
                                case 2: tree_dot<J, 0, 2>(rd, at, r, en, check, nconv, fail); break;
                                case 3: tree_dot<J, 0, 3>(rd, at, r, en, check, nconv, fail); break;
                                case 4: tree_dot<J, 1, 0>(rd, at, r, en, check, nconv, fail); break;
                                case 5: tree_dot<J, 1, 1>(rd, at, r, en, check, nconv, fail); break;
                                case 6: tree_dot<J, 1, 2>(rd, at, r, en, check, nconv, fail); break;
                                case 7: tree_dot<J, 1, 3>(rd, at, r, en, check, nconv, fail); break;
                                case 8: tree_dot<J, 2, 0>(rd, at, r, en, check, nconv, fail); break;
                                case 9: tree_dot<J, 2, 1>(rd, at, r, en, check, nconv, fail); break;
                                case 10: tree_dot<J, 2, 2>(rd, at, r, en, check, nconv, fail); break;
                                default: tree_dot<J, 2, 3>(rd, at, r, en, check, nconv, fail); break;
                            }
                            rd.adv(4u + na + nb);
                            continue;
                        }
                        if constexpr(FUSED)
                        {
                            if(op == PE_OP_CROUT2)
                            {
                                tree_crout2<J>(rd, at, r.guard, en, fail);
                                rd.adv(21u);
                                continue;
                            }
                        }
                        if(op == PE_OP_CAP_STEP)
                        {
                            if((h & 0x2000u) && it > 0)  // folded time-step update: first Newton iteration only
                            {
                                rd.adv(8u);
                                continue;
                            }
                            // [h][mask][hist][prev_g][C][dt][va][vb]  (capacitor.h:106-128)
                            double* const ph = at(rd.lo(2) & 0x7fffu);
                            double* const pg = at(rd.lo(3) & 0x7fffu);
                            double const* const pc = at(rd.lo(4) & 0x7fffu);
                            double const* const pt = at(rd.lo(5) & 0x7fffu);
                            double const* const pa = at(rd.lo(6) & 0x7fffu);
                            double const* const pb = at(rd.lo(7) & 0x7fffu);
                            double hv[J], gv[J], cv[J], tv[J], va[J], vb[J];
#pragma unroll
                            for(int j = 0; j < J; ++j)
                            {
                                hv[j] = ph[32 * j];
                                gv[j] = pg[32 * j];
                                cv[j] = pc[32 * j];
                                tv[j] = pt[32 * j];
                                va[j] = pa[32 * j];
                                vb[j] = pb[32 * j];
                            }
#pragma unroll
                            for(int j = 0; j < J; ++j)
                            {
                                pe_models::cap_step(cv[j], tv[j], PE_SUB(va[j], vb[j]), hv[j], gv[j]);
                                if(en[j])
                                {
                                    ph[32 * j] = hv[j];
                                    pg[32 * j] = gv[j];
                                }
                            }
                            rd.adv(8u);
                            continue;
                        }
                        // everything else goes through the generic vector-op executor (pe_b200_rinterp.h)
                        rctx c;
#ifdef PE_JIT
                        c.ws = reinterpret_cast<double*>(at.wl);
                        c.I = (uint64_t)GL;
#else
                        c.ws = r.wsg + glane;
                        c.I = (uint64_t)r.LSw;
#endif
                        c.S = S;
                        c.C = 1;
                        c.col = 0;
                        c.stream = warp;
                        c.js = 32;
                        tol_t const tol{r.v_abstol, r.v_reltol, r.i_abstol, r.i_reltol, r.guard};
                        int const k = rvop<J, line_reader, false>(rd, c, t, tol, en, check, nconv, fail, it == 0);
                        if(k == V_END || k == V_BAD) { break; }
                        if(k == V_BAR)
                        {
                            group_sync<CL>();
                            rd.bar();
                        }
                        else if(k == V_SKIP) { rd.skip(); }
                        else
                        {
                            rd.close();
                        }
                    }
                    group_sync<CL>();  // results of this section are visible to every warp of the group
                }

                // ---- what comes next
                if(stage == 0) { stage = 1; }
                else if(stage == 1)
                {
                    t = t + r.dt;
#pragma unroll
                    for(int j = 0; j < J; ++j) { done[j] = !ok[j]; }
                    it = 0;
                    stage = 2;
                }
                else
                {
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        if(nconv[j] || fail[j]) { atomicOr(&s_flags[fi][lane + 32 * j], (nconv[j] ? 1u : 0u) | (fail[j] ? 2u : 0u)); }
                    }
                    group_sync<CL>();
                    ++it;
                    bool all_done = true;
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        uint32_t f = s_flags[fi][lane + 32 * j];
                        if constexpr(CL == 2) { f |= peer_flags[fi * 32 * J + lane + 32 * j]; }
                        if(!done[j])
                        {
                            ++solves[j];
                            if(f & 2u)
                            {
                                status[j] = PE_ST_SINGULAR;
                                ok[j] = false;
                                done[j] = true;
                            }
                            else if(!r.nonlinear || !(f & 1u)) { done[j] = true; }
                            else if(it >= r.max_iter)
                            {
                                status[j] = PE_ST_NO_CONVERGENCE;
                                ok[j] = false;
                                done[j] = true;
                            }
                        }
                        all_done = all_done && done[j];
                    }
                    fi = fi == 2 ? 0 : fi + 1;
                    if(__syncthreads_and(all_done ? 1 : 0))
                    {
                        if(r.wave != nullptr && warp == 0)
                        {
#pragma unroll
                            for(int j = 0; j < J; ++j)
                            {
                                if(!ok[j]) { continue; }
                                for(int32_t p = 0; p < r.n_probe; ++p) { r.wave[((int64_t)s * r.n_probe + p) * r.LSu + glane + 32 * j] = at(__ldg(r.probes + p))[32 * j]; }
                            }
                        }
                        ++s;
                        stage = 1;
                    }
                }
            }
            for(uint32_t e = warp; last_chunk && e < (uint32_t)r.n_io; e += n_warps)
            {
                pe_b200_io const io = r.io[e];
                if(!((io.slot_kind >> 20) & PE_IO_STORE)) { continue; }
                double const* const src = at(io.slot_kind & 0xffffu);
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(counted[j]) { r.wu[(int64_t)io.src * r.LSu + glane + 32 * j] = src[32 * j]; }
                }
            }
            if(warp == 0)
            {
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    if(counted[j])
                    {
                        r.status[glane + 32 * j] = status[j];
                        r.solves[glane + 32 * j] += solves[j];
                    }
                }
            }
            if(r.sched == nullptr) { break; }
            // publish the chunk: every thread's stores are visible device-wide before the group's counter moves
            __threadfence();
            group_sync<CL>();
            if(tid == 0 && rank == 0u) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(r.sched + 1 + group), "r"(chunk + 1u) : "memory"); }
        }
        if constexpr(CL == 2) { group_sync<CL>(); }  // neither CTA leaves while the other may still read its shared memory
    }

#ifndef PE_JIT
    template <int J, int MAXT, int MINB, bool FUSED, int CL>
    __global__ void __launch_bounds__(MAXT, MINB) pe_b200_tree_kernel(pe_b200_rrun const r)
    {
        tree_body<J, FUSED, CL>(r);
    }

    // mixed-signal boundary, analog -> digital (controller/comparator.h:88-101): one thread per (comparator, lane)
    __global__ void pe_b200_compare_kernel(double const* __restrict__ x, int64_t LS, int32_t n_lanes, int32_t const* __restrict__ ab, int32_t n_cmp,
                                           uint8_t* __restrict__ out)
    {
        int64_t const lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int32_t const cmp = (int32_t)blockIdx.y;
        if(lane >= n_lanes || cmp >= n_cmp) { return; }
        int32_t const a = __ldg(ab + 2 * cmp), b = __ldg(ab + 2 * cmp + 1);
        double const va = a < 0 ? 0.0 : x[(int64_t)a * LS + lane];
        double const vb = b < 0 ? 0.0 : x[(int64_t)b * LS + lane];
        out[(int64_t)cmp * LS + lane] = va >= vb ? 1 : 0;
    }

    __global__ void pe_b200_status_reduce_kernel(int32_t const* __restrict__ status, uint32_t const* __restrict__ solves, int64_t n, unsigned long long* __restrict__ out)
    {
        unsigned long long bad = 0, sing = 0, sum = 0;
        for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        {
            int32_t const s = status[i];
            bad += s != PE_ST_OK;
            sing += s == PE_ST_SINGULAR;
            sum += solves[i];
        }
#pragma unroll
        for(int d = 16; d > 0; d >>= 1)
        {
            bad += __shfl_down_sync(0xffffffffu, bad, d);
            sing += __shfl_down_sync(0xffffffffu, sing, d);
            sum += __shfl_down_sync(0xffffffffu, sum, d);
        }
        if((threadIdx.x & 31u) == 0u)
        {
            if(bad) { atomicAdd(out, bad); }
            if(sing) { atomicAdd(out + 1, sing); }
            if(sum) { atomicAdd(out + 2, sum); }
        }
    }

    thread_local char g_err[256] = "";
    std::atomic<uint64_t> g_launches{0}, g_aux_launches{0};

    // optional per-launch device timing (CUDA events on the launching stream), used by bench.py for the roofline line
    std::atomic<bool> g_timing{false};
    // launches come from several host threads (one per batch handle / stream / device): the event list is shared
    std::mutex g_events_mu;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
    void record_events(cudaEvent_t e0, cudaEvent_t e1)
    {
        std::lock_guard<std::mutex> lk(g_events_mu);
        g_events.emplace_back(e0, e1);
    }

    int chk(cudaError_t e, char const* what)
    {
        // debugging aid (PE_B200_SYNC_ALL): wait for the device after every seam call and name the first one that fails
        static bool const sync_all = std::getenv("PE_B200_SYNC_ALL") != nullptr;
        if(sync_all && e == cudaSuccess)
        {
            e = cudaDeviceSynchronize();
            fprintf(stderr, "seam: %s -> %s\n", what, cudaGetErrorString(e));
        }
        if(e == cudaSuccess) { return 0; }
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
#endif  // PE_JIT
}  // namespace

#ifdef PE_JIT
extern "C" __global__ void __launch_bounds__(512, 1) pe_b200_jit_kernel(pe_b200_rrun const r) { tree_body<4, false, PE_JIT_CL>(r); }
#else

extern "C"
{
    int pe_b200_dev_count(void)
    {
        int n = 0;
        if(cudaGetDeviceCount(&n) != cudaSuccess)
        {
            (void)cudaGetLastError();
            return 0;
        }
        return n;
    }

    int pe_b200_dev_set(int device) { return chk(cudaSetDevice(device), "cudaSetDevice"); }

    int pe_b200_dev_malloc(void** p, size_t bytes) { return chk(cudaMalloc(p, bytes ? bytes : 8), "cudaMalloc"); }

    int pe_b200_dev_free(void* p) { return p ? chk(cudaFree(p), "cudaFree") : 0; }

    int pe_b200_dev_memset0(void* p, size_t bytes, void* stream) { return chk(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream), "cudaMemsetAsync"); }

    int pe_b200_dev_h2d(void* dst, void const* src, size_t bytes, void* stream)
    {
        return chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync(H2D)");
    }

    int pe_b200_dev_d2h(void* dst, void const* src, size_t bytes, void* stream)
    {
        return chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(D2H)");
    }

    int pe_b200_dev_h2d_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream)
    {
        return chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpy2DAsync(H2D)");
    }

    int pe_b200_dev_d2h_2d(void* dst, size_t dpitch, void const* src, size_t spitch, size_t width, size_t height, void* stream)
    {
        return chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpy2DAsync(D2H)");
    }

    int pe_b200_dev_sync(void* stream) { return chk(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

    int pe_b200_launch(pe_b200_run const* run, void* stream)
    {
        if(run == nullptr || run->n_lanes <= 0) { return 0; }
        if(run->warps < 1 || run->warps > PE_MAX_WARPS)
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch: warps per CTA must be 1..%d", PE_MAX_WARPS);
            return 1;
        }
        int const block = 32 * run->warps;
        int const grid = (run->n_lanes + 31) / 32;
        cudaEvent_t e0{}, e1{};
        if(g_timing)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)stream);
        }
        pe_b200_solve_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*run);
        if(g_timing)
        {
            cudaEventRecord(e1, (cudaStream_t)stream);
            record_events(e0, e1);
        }
        g_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_solve_kernel launch");
    }

    size_t pe_b200_resident_smem_limit(void)
    {
        int dev = 0, v = 0;
        if(cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        {
            (void)cudaGetLastError();
            return 0;
        }
        return v > 1024 ? (size_t)v - 1024 : 0;  // static shared memory of the kernel (flags) comes out of the same budget
    }

    // specialised kernels (host/jit.cpp): cubin -> kernel handle, loaded once per process and key
    static thread_local void const* g_jit_kernel = nullptr;
    static thread_local int g_stream_geom[3] = {0, 0, 0};

    int pe_b200_jit_supported(void) { return 1; }

    int pe_b200_launch_jit(pe_b200_rrun const* run, void const* cubin, size_t bytes, uint64_t key, void* stream)
    {
        static std::mutex mu;
        static std::map<std::pair<int, uint64_t>, cudaKernel_t> loaded;
        int dev = 0;
        if(chk(cudaGetDevice(&dev), "cudaGetDevice") != 0) { return 1; }
        cudaKernel_t k{};
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = loaded.find({dev, key});
            if(it == loaded.end())
            {
                if(cubin == nullptr || bytes == 0)
                {
                    snprintf(g_err, sizeof(g_err), "pe_b200_launch_jit: no cubin");
                    return 1;
                }
                cudaLibrary_t lib{};
                if(chk(cudaLibraryLoadData(&lib, cubin, nullptr, nullptr, 0, nullptr, nullptr, 0), "cudaLibraryLoadData") != 0) { return 1; }
                if(chk(cudaLibraryGetKernel(&k, lib, "pe_b200_jit_kernel"), "cudaLibraryGetKernel(pe_b200_jit_kernel)") != 0) { return 1; }
                it = loaded.emplace(std::make_pair(dev, key), k).first;
            }
            k = it->second;
        }
        g_jit_kernel = (void const*)k;
        int const rc = pe_b200_launch_resident(run, stream);
        g_jit_kernel = nullptr;
        return rc;
    }

    // ---- stream kernel (csrc/pe_b200_stream.cu + generated source, host/stream.cpp) ------------------------------------
    int pe_b200_stream_supported(void) { return 1; }

    // nvcc -cubin for sm_100a; the module holds pe_b200_stream_kernel specialised for one program and J lanes per thread
    int pe_b200_stream_build(char const* source_path, char const* out_path, char const* csrc_dir, int J, int GL, char* log, size_t log_cap)
    {
        char const* nv = std::getenv("PE_B200_NVCC");
        std::string nvcc = nv != nullptr ? nv : "/usr/local/cuda/bin/nvcc";
        if(::access(nvcc.c_str(), X_OK) != 0) { nvcc = "nvcc"; }
        std::string const logf = std::string(out_path) + ".log";
        // debugging aid: bounds checks with device printf (PE_B200_STREAM_CHECKS)
        std::string dbg_flags;
        if(std::getenv("PE_B200_STREAM_CHECKS") != nullptr) { dbg_flags += "-DPE_SK_DEBUG "; }
        // a module is a few rolled loops (seconds to compile); the time limit only guards against a pathological program
        std::string const cmd = "timeout 600 " + nvcc + " -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin " + dbg_flags + "-DPE_SJ=" + std::to_string(J) + " -DPE_SGL=" + std::to_string(GL) +
                                " '-DPE_STREAM_SOURCE=\"" + source_path +
                                "\"' -I'" + csrc_dir + "' -o '" + out_path + "' '" + csrc_dir + "/pe_b200_stream.cu' > '" + logf + "' 2>&1";
        int const rc = std::system(cmd.c_str());
        if(rc != 0 && log != nullptr && log_cap > 0)
        {
            log[0] = 0;
            if(FILE* f = std::fopen(logf.c_str(), "rb"))
            {
                size_t const n = std::fread(log, 1, log_cap - 1, f);
                log[n] = 0;
                std::fclose(f);
            }
        }
        return rc == 0 ? 0 : 1;
    }

    int pe_b200_launch_stream(pe_b200_rrun const* run, void const* blob, size_t bytes, uint64_t key, uint32_t n_tiles, uint32_t stage_rows, void* stream)
    {
        (void)n_tiles;
        if(run == nullptr || run->n_lanes <= 0) { return 0; }
        static std::mutex mu;
        static std::map<std::pair<int, uint64_t>, cudaKernel_t> loaded;
        int dev = 0;
        if(chk(cudaGetDevice(&dev), "cudaGetDevice") != 0) { return 1; }
        int const J = run->J;
        int const GLJ = run->I;  // lanes per group
        if(run->S != 1 || run->wsg == nullptr || (J != 1 && J != 2 && J != 4) || (GLJ != 32 * J && !(J == 1 && (GLJ == 16 || GLJ == 8))) || run->nonlinear != 0 || run->cplx != 0)
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: needs a real linear one-stream program in HBM form (S=%d I=%d J=%d)", run->S, run->I, J);
            return 1;
        }
        cudaKernel_t k{};
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = loaded.find({dev, key});
            if(it == loaded.end())
            {
                if(blob == nullptr || bytes == 0)
                {
                    snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: no module");
                    return 1;
                }
                cudaLibrary_t lib{};
                if(chk(cudaLibraryLoadData(&lib, blob, nullptr, nullptr, 0, nullptr, nullptr, 0), "cudaLibraryLoadData") != 0) { return 1; }
                if(chk(cudaLibraryGetKernel(&k, lib, "pe_b200_stream_kernel"), "cudaLibraryGetKernel(pe_b200_stream_kernel)") != 0) { return 1; }
                it = loaded.emplace(std::make_pair(dev, key), k).first;
            }
            k = it->second;
        }
        int sms = 0, smem_max = 0;
        if(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
           sms < 1)
        {
            (void)cudaGetLastError();
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: device query failed");
            return 1;
        }
        // geometry: one CTA per SM, W warps each = lane groups resident per SM; every warp owns a ring of NS stages
        int const GL = GLJ;
        long long const NG = ((long long)run->n_lanes + GL - 1) / GL;
        size_t const stage_bytes = (size_t)stage_rows * 8u * (size_t)GL;
        static int const w_max = std::getenv("PE_B200_STREAM_WARPS") ? std::atoi(std::getenv("PE_B200_STREAM_WARPS")) : 8;
        static int const ns_cap = std::getenv("PE_B200_STREAM_NS") ? std::atoi(std::getenv("PE_B200_STREAM_NS")) : 8;
        int W = (int)std::min<long long>(std::max(1, std::min(w_max, 8)), (NG + sms - 1) / sms);
        uint32_t ns_log = 0;
        for(;;)
        {
            size_t const budget = (size_t)smem_max / (size_t)W;
            ns_log = 0;
            while(ns_log < 4 && (2u << ns_log) <= (uint32_t)ns_cap && (size_t)(2u << ns_log) * stage_bytes + 128u <= budget) { ++ns_log; }
            if(((size_t)(1u << ns_log) * stage_bytes + 128u <= budget && ns_log >= 1) || W == 1) { break; }
            --W;  // fewer resident groups per SM, deeper rings; the warps then walk over several groups
        }
        size_t const smem = (size_t)W * ((size_t)(1u << ns_log) * stage_bytes + 128u);
        if(ns_log < 1 || smem > (size_t)smem_max)
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_stream: a ring stage of %zu bytes does not fit shared memory", stage_bytes);
            return 1;
        }
        static bool const dbg_sync = std::getenv("PE_B200_STREAM_SYNC") != nullptr;
        if(dbg_sync)
        {
            fprintf(stderr, "stream launch: lanes %d GL %d groups %lld W %d NS %u stage %zu B smem %zu B; before: %s\n", run->n_lanes, GL, NG, W, 1u << ns_log, stage_bytes, smem,
                    cudaGetErrorString(cudaStreamSynchronize((cudaStream_t)stream)));
        }
        if(chk(cudaFuncSetAttribute((void const*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(stream smem)") != 0) { return 1; }
        cudaEvent_t e0{}, e1{};
        if(g_timing)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)stream);
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)std::min<long long>(sms, NG));
        cfg.blockDim = dim3((unsigned)(32 * W));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cfg.numAttrs = 0;
        pe_b200_rrun arg = *run;
        uint32_t nsl = ns_log;
        void* kargs[2] = {&arg, &nsl};
        cudaError_t const le = cudaLaunchKernelExC(&cfg, (void const*)k, kargs);
        if(g_timing)
        {
            cudaEventRecord(e1, (cudaStream_t)stream);
            record_events(e0, e1);
        }
        if(le != cudaSuccess) { return chk(le, "pe_b200_stream_kernel launch"); }
        if(dbg_sync) { fprintf(stderr, "stream launch: after: %s\n", cudaGetErrorString(cudaStreamSynchronize((cudaStream_t)stream))); }
        g_launches.fetch_add(1);
        g_stream_geom[0] = W;
        g_stream_geom[1] = 1 << ns_log;
        g_stream_geom[2] = (int)smem;
        return chk(cudaGetLastError(), "pe_b200_stream_kernel launch");
    }

    // geometry of the last stream launch of this thread: warps per CTA, ring stages, dynamic shared memory per CTA
    void pe_b200_stream_last_geometry(int* out3)
    {
        for(int i = 0; i < 3; ++i) { out3[i] = g_stream_geom[i]; }
    }

    int pe_b200_launch_resident(pe_b200_rrun const* run, void* stream)
    {
        if(run == nullptr || run->n_lanes <= 0) { return 0; }
        int const I = run->I, J = run->J, S = run->S;
        if(I < 1 || I > 128 || (J != 1 && J != 2 && J != 4) || I % J != 0 || S < 1 || (32 % (I / J)) != 0 || (S * (I / J)) % 32 != 0 || S * (I / J) > 1024 ||
           (run->wsg == nullptr && (I > 32 || J > 2)))
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: bad geometry S=%d I=%d J=%d", S, I, J);
            return 1;
        }
        bool const hbm{run->wsg != nullptr};
        if(hbm && (I != 32 * J || S > 32))
        {
            snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: the HBM form needs I = 32 J and S <= 32 (S=%d I=%d J=%d)", S, I, J);
            return 1;
        }
        int const block = S * (I / J);
        int const grid = (run->n_lanes + I - 1) / I;
        size_t const smem = hbm ? 0 : (size_t)run->n_slots * (size_t)I * sizeof(double);
        // two register budgets: CTAs of up to 512 threads get 128 registers per thread, larger ones 64
        if(hbm)
        {
            cudaEvent_t e0{}, e1{};
            if(g_timing)
            {
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaEventRecord(e0, (cudaStream_t)stream);
            }
            void const* tk = nullptr;
            // register budgets: 64 per thread by default; programs with fused elimination steps (26 J operand rows in
            // flight per thread) and 128-lane groups get 128 in CTAs of at most 512 threads
            bool const fused{run->regs128 != 0};
            int const CL{run->cluster == 2 ? 2 : 1};
            int const cta_threads{block / CL};
            if(J == 4)
            {
                if(cta_threads > 512)
                {
                    snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: J = 4 needs at most 16 sub-tree warps per CTA");
                    return 1;
                }
                tk = CL == 2 ? (fused ? (void const*)pe_b200_tree_kernel<4, 512, 1, true, 2> : (void const*)pe_b200_tree_kernel<4, 512, 1, false, 2>)
                             : (fused ? (void const*)pe_b200_tree_kernel<4, 512, 1, true, 1> : (void const*)pe_b200_tree_kernel<4, 512, 1, false, 1>);
                if(g_jit_kernel != nullptr) { tk = g_jit_kernel; }
            }
            else if(CL == 2)
            {
                snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: clusters are used with J = 4 only");
                return 1;
            }
            else if(J == 2)
            {
                tk = fused ? (block <= 512 ? (void const*)pe_b200_tree_kernel<2, 512, 1, true, 1> : (void const*)pe_b200_tree_kernel<2, 1024, 1, true, 1>)
                           : (block <= 256 ? (void const*)pe_b200_tree_kernel<2, 256, 4, false, 1>
                                           : (block <= 512 ? (void const*)pe_b200_tree_kernel<2, 512, 2, false, 1> : (void const*)pe_b200_tree_kernel<2, 1024, 1, false, 1>));
            }
            else
            {
                tk = fused ? (block <= 512 ? (void const*)pe_b200_tree_kernel<1, 512, 1, true, 1> : (void const*)pe_b200_tree_kernel<1, 1024, 1, true, 1>)
                           : (block <= 256 ? (void const*)pe_b200_tree_kernel<1, 256, 4, false, 1>
                                           : (block <= 512 ? (void const*)pe_b200_tree_kernel<1, 512, 2, false, 1> : (void const*)pe_b200_tree_kernel<1, 1024, 1, false, 1>));
            }
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute attr[1];
            cfg.blockDim = dim3((unsigned)cta_threads);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = (cudaStream_t)stream;
            cfg.numAttrs = 0;
            if(CL == 2)
            {
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
            }
            int tgroups = grid;  // CTAs (CL = 1) or clusters (CL = 2) to launch
            if(run->sched != nullptr)
            {
                // persistent CTAs / clusters, all co-resident (an item may wait for the previous chunk of its group, which is
                // always already running)
                int dev = 0, sms = 0, occ = 0, max_clusters = 0;
                bool ok = cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
                if(ok && CL == 1) { ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tk, cta_threads, 0) == cudaSuccess && occ >= 1; }
                if(ok && CL == 2)
                {
                    cfg.gridDim = dim3(2u * (unsigned)sms);
                    ok = cudaOccupancyMaxActiveClusters(&max_clusters, tk, &cfg) == cudaSuccess && max_clusters >= 1;
                }
                if(!ok)
                {
                    (void)cudaGetLastError();
                    snprintf(g_err, sizeof(g_err), "pe_b200_launch_resident: occupancy query failed");
                    return 1;
                }
                long long const items = (long long)run->n_chunks * grid;
                tgroups = (int)std::min<long long>(items, CL == 2 ? (long long)max_clusters : (long long)occ * sms);
            }
            cfg.gridDim = dim3((unsigned)(tgroups * CL));
            pe_b200_rrun arg = *run;
            void* kargs[1] = {&arg};
            cudaError_t const le = cudaLaunchKernelExC(&cfg, tk, kargs);
            if(g_timing)
            {
                cudaEventRecord(e1, (cudaStream_t)stream);
                record_events(e0, e1);
            }
            if(le != cudaSuccess) { return chk(le, "pe_b200_tree_kernel launch"); }
            g_launches.fetch_add(1);
            return chk(cudaGetLastError(), "pe_b200_tree_kernel launch");
        }
        auto kern = block <= 512 ? (J == 2 ? pe_b200_resident_kernel<2, 512> : pe_b200_resident_kernel<1, 512>)
                                 : (J == 2 ? pe_b200_resident_kernel<2, 1024> : pe_b200_resident_kernel<1, 1024>);
        if(chk(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(smem)") != 0) { return 1; }
        cudaEvent_t e0{}, e1{};
        if(g_timing)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)stream);
        }
        kern<<<grid, block, smem, (cudaStream_t)stream>>>(*run);
        if(g_timing)
        {
            cudaEventRecord(e1, (cudaStream_t)stream);
            record_events(e0, e1);
        }
        g_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_resident_kernel launch");
    }

    int pe_b200_compare(double const* x, int64_t LS, int32_t n_lanes, int32_t const* ab, int32_t n_cmp, uint8_t* out, void* stream)
    {
        if(n_lanes <= 0 || n_cmp <= 0) { return 0; }
        dim3 const grid((unsigned)((n_lanes + 255) / 256), (unsigned)n_cmp);
        pe_b200_compare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, LS, n_lanes, ab, n_cmp, out);
        g_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_compare_kernel launch");
    }

    int pe_b200_status_reduce(int32_t const* status, uint32_t const* solves, int64_t n_lanes, unsigned long long* out3, void* stream)
    {
        if(chk(cudaMemsetAsync(out3, 0, 3 * sizeof(unsigned long long), (cudaStream_t)stream), "status reduce: zero") != 0) { return 1; }
        if(n_lanes <= 0) { return 0; }
        unsigned const grid = (unsigned)std::min<int64_t>((n_lanes + 255) / 256, 148 * 8);
        pe_b200_status_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(status, solves, n_lanes, out3);
        g_aux_launches.fetch_add(1);
        return chk(cudaGetLastError(), "pe_b200_status_reduce_kernel launch");
    }

    void pe_b200_timing_enable(int on) { g_timing = on != 0; }

    // total device time (ms) of the solve-kernel launches since the last collect; waits for them to finish
    double pe_b200_timing_collect(void)
    {
        double total = 0.0;
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
        {
            std::lock_guard<std::mutex> lk(g_events_mu);
            ev.swap(g_events);
        }
        for(auto& [a, b]: ev)
        {
            float ms = 0.f;
            cudaEventSynchronize(b);
            if(cudaEventElapsedTime(&ms, a, b) == cudaSuccess) { total += ms; }
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
        return total;
    }

    char const* pe_b200_dev_last_error(void) { return g_err; }

    uint64_t pe_b200_launch_count(void) { return g_launches.load(); }
    uint64_t pe_b200_aux_launch_count(void) { return g_aux_launches.load(); }
}
#endif  // PE_JIT
