// pe_b200_frontal.cu — the sm_100a kernels of the REDUCE-AND-CORE path: one huge linear DC circuit per
// instance (config A of BASELINE.json, benchmark/series_parallel.cpp: a ring of 100 003 random resistors with 9 000 random
// node merges, n ~ 9e4 unknowns), where the parallelism is in the elimination DAG, not in the batch.
//
// Replaces, per instance, solve_once (circuit.h:987-1527: stamping of resistance.h:82-110 / VDC.h:82-96 / IDC.h:84-94 and
// Eigen::SparseLU compute + solve) by
//   1. fr_assemble       conductances g = 1 / r into node diagonals and edge values (parallel resistors share an edge)
//   2. fr_eliminate      LEVEL-SCHEDULED elimination of every node of degree <= 2 (series / parallel reduction): the nodes of
//                        a level are pairwise non-adjacent, thread = (node, instance); a chain of length L between two
//                        junctions vanishes in ~log2 L levels (cyclic reduction), the symbolic phase (host/frontal.cpp) knows
//                        every fill edge in advance
//   3. fr_core_assemble  what is left (junction nodes, source nodes, source branches) as a dense matrix, nodes first
//   4. fr_lu_*           blocked right-looking dense LU without pivoting (the node block is the Schur complement of an
//                        M-matrix: symmetric positive definite; the branch rows come last: quasi-definite), trailing update
//                        = FP64 tensor-core contraction (DMMA mma.sync m8n8k4; tcgen05 has no FP64 kind)
//   5. fr_solve_core     forward / back substitution of the core
//   6. fr_back           level-scheduled back substitution of the eliminated nodes, results into the solution rows
// Instances (Monte-Carlo resistor draws) are independent: values are lane-interleaved v[id][instance].
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>

#include "pe_b200_program.h"

namespace
{
    constexpr int NB = 64;  // block size of the dense LU

    __global__ void fr_assemble(pe_b200_frontal const f)
    {
        int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t const B = f.B;
        int64_t const r = t / B, inst = t - r * B;
        if(inst >= f.n_inst) { return; }
        if(r < f.n_res)
        {
            double const g = 1.0 / f.rval[r * B + inst];
            int32_t const a = f.res_a[r], b = f.res_b[r], e = f.res_edge[r];
            if(a >= 0) { atomicAdd(f.d + (int64_t)a * B + inst, g); }
            if(b >= 0) { atomicAdd(f.d + (int64_t)b * B + inst, g); }
            if(e >= 0) { atomicAdd(f.g + (int64_t)e * B + inst, g); }
        }
        else if(r < f.n_res + f.n_idc)
        {
            // IDC.h:84-94: I(+) -= I, I(-) += I
            int64_t const k = r - f.n_res;
            double const i = f.idc_val[k * B + inst];
            int32_t const p = f.idc_p[k], q = f.idc_q[k];
            if(p >= 0) { atomicAdd(f.z + (int64_t)p * B + inst, -i); }
            if(q >= 0) { atomicAdd(f.z + (int64_t)q * B + inst, i); }
        }
    }

    // one level of the elimination DAG: op = (k, a, b, e_ka, e_kb, e_ab); row k: d_k x_k - g_ka x_a - g_kb x_b = z_k
    __global__ void fr_eliminate(pe_b200_frontal const f, int32_t op0, int32_t n_ops)
    {
        int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t const B = f.B;
        int64_t const o = t / B, inst = t - o * B;
        if(o >= n_ops || inst >= f.n_inst) { return; }
        int32_t const* op = f.ops + (int64_t)(op0 + o) * 6;
        int32_t const k = op[0], a = op[1], b = op[2], eka = op[3], ekb = op[4], eab = op[5];
        double const p = f.d[(int64_t)k * B + inst];
        if(p == 0.0 || !isfinite(p)) { f.status[inst] = PE_ST_SINGULAR; }
        double const zk = f.z[(int64_t)k * B + inst];
        double const g1 = a >= 0 ? f.g[(int64_t)eka * B + inst] : 0.0;
        double const g2 = b >= 0 ? f.g[(int64_t)ekb * B + inst] : 0.0;
        double const m1 = g1 / p, m2 = g2 / p;
        if(a >= 0)
        {
            atomicAdd(f.d + (int64_t)a * B + inst, -g1 * m1);
            atomicAdd(f.z + (int64_t)a * B + inst, m1 * zk);
        }
        if(b >= 0)
        {
            atomicAdd(f.d + (int64_t)b * B + inst, -g2 * m2);
            atomicAdd(f.z + (int64_t)b * B + inst, m2 * zk);
        }
        if(eab >= 0) { atomicAdd(f.g + (int64_t)eab * B + inst, g1 * m2); }
    }

    // the reverse: x_k = (z_k + g_ka x_a + g_kb x_b) / d_k
    __global__ void fr_back(pe_b200_frontal const f, int32_t op0, int32_t n_ops)
    {
        int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t const B = f.B;
        int64_t const o = t / B, inst = t - o * B;
        if(o >= n_ops || inst >= f.n_inst) { return; }
        int32_t const* op = f.ops + (int64_t)(op0 + o) * 6;
        int32_t const k = op[0], a = op[1], b = op[2], eka = op[3], ekb = op[4];
        double v = f.z[(int64_t)k * B + inst];
        if(a >= 0) { v = fma(f.g[(int64_t)eka * B + inst], f.x[(int64_t)a * f.LSx + inst], v); }
        if(b >= 0) { v = fma(f.g[(int64_t)ekb * B + inst], f.x[(int64_t)b * f.LSx + inst], v); }
        f.x[(int64_t)k * f.LSx + inst] = v / f.d[(int64_t)k * B + inst];
    }

    // dense core of one instance (blockIdx.y), column-major M[i + j * ld], right-hand side c
    __global__ void fr_core_assemble(pe_b200_frontal const f)
    {
        int64_t const inst = blockIdx.y;
        int64_t const B = f.B, nc = f.n_core, ld = f.ld_core;
        double* const M = f.M + inst * ld * ld;
        double* const c = f.c + inst * ld;
        int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if(t < f.n_core_nodes)
        {
            int32_t const u = f.core_unknown[t];
            M[t + t * ld] = f.d[(int64_t)u * B + inst];
            c[t] = f.z[(int64_t)u * B + inst];
        }
        else if(t < nc)
        {
            // a source branch: VDC.h:82-96 (B, C assigned +-1, E = V)
            int64_t const k = t - f.n_core_nodes;
            int32_t const p = f.vdc[3 * k + 1], q = f.vdc[3 * k + 2];
            if(p >= 0)
            {
                M[p + t * ld] = 1.0;
                M[t + p * ld] = 1.0;
            }
            if(q >= 0)
            {
                M[q + t * ld] = -1.0;
                M[t + q * ld] = -1.0;
            }
            c[t] = f.vdc_val[k * B + inst];
        }
        else if(t < nc + f.n_core_edges)
        {
            int64_t const k = t - nc;
            int32_t const i = f.core_edges[3 * k], j = f.core_edges[3 * k + 1], e = f.core_edges[3 * k + 2];
            double const g = f.g[(int64_t)e * B + inst];
            M[i + j * ld] = -g;
            M[j + i * ld] = -g;
        }
        else if(t < nc + f.n_core_edges + (ld - nc))
        {
            // padding rows / columns up to a multiple of the block size: identity
            int64_t const k = nc + (t - nc - f.n_core_edges);
            M[k + k * ld] = 1.0;
            c[k] = 0.0;
        }
    }

    // ---- blocked right-looking LU without pivoting (ld is a multiple of NB) --------------------------------------------
    // diagonal block k: in-place LU in shared memory
    __global__ void __launch_bounds__(NB) fr_lu_diag(double* __restrict__ Mall, int64_t ld, int32_t k0, int32_t* status)
    {
        // thread i keeps ROW i of the block in registers (its loads and stores are coalesced across the threads).  Step k: the
        // owner of row k publishes it, every thread below divides its own entry of column k by the pivot and applies the row:
        // ONE barrier per step (the published row is double-buffered), one division per thread, 64 - k FMAs
        __shared__ double urow[2][NB];
        double* const M = Mall + (int64_t)blockIdx.z * ld * ld;
        int const i = threadIdx.x;
        double row[NB];
#pragma unroll
        for(int j = 0; j < NB; ++j) { row[j] = M[(k0 + i) + (int64_t)(k0 + j) * ld]; }
#pragma unroll
        for(int k = 0; k < NB; ++k)
        {
            if(i == k)
            {
#pragma unroll
                for(int j = 0; j < NB; ++j)
                {
                    if(j >= k) { urow[k & 1][j] = row[j]; }
                }
            }
            __syncthreads();
            double const p = urow[k & 1][k];
            if(i == k && (p == 0.0 || !isfinite(p))) { status[blockIdx.z] = PE_ST_SINGULAR; }
            if(i > k)
            {
                double const l = row[k] / p;
                row[k] = l;
#pragma unroll
                for(int j = 0; j < NB; ++j)
                {
                    if(j > k) { row[j] = fma(-l, urow[k & 1][j], row[j]); }
                }
            }
        }
#pragma unroll
        for(int j = 0; j < NB; ++j) { M[(k0 + i) + (int64_t)(k0 + j) * ld] = row[j]; }
    }

    // panels of block step k: blockIdx.x < nbl: L21 block row = A21 U11^-1, else U12 block column = L11^-1 A12
    __global__ void __launch_bounds__(NB) fr_lu_panel(double* __restrict__ Mall, int64_t ld, int32_t k0, int32_t nbl)
    {
        __shared__ double s[NB][NB + 1];
        double* const M = Mall + (int64_t)blockIdx.z * ld * ld;
        int const tid = threadIdx.x;
        for(int e = tid; e < NB * NB; e += NB) { s[e % NB][e / NB] = M[(k0 + e % NB) + (int64_t)(k0 + e / NB) * ld]; }
        __syncthreads();
        if((int)blockIdx.x < nbl)
        {
            // row r of the block row below the diagonal: l_rj = (a_rj - sum_{t<j} l_rt u_tj) / u_jj
            int64_t const r = k0 + NB + (int64_t)blockIdx.x * NB + tid;
            // all 64 loads first, all 64 stores last: interleaved, every load would wait for the store before it (the compiler
            // cannot tell they do not alias) -- 64 serial round trips to memory per thread
            double l[NB];
#pragma unroll
            for(int j = 0; j < NB; ++j) { l[j] = M[r + (int64_t)(k0 + j) * ld]; }
#pragma unroll
            for(int j = 0; j < NB; ++j)
            {
                double v = l[j];
#pragma unroll
                for(int t = 0; t < NB; ++t)
                {
                    if(t < j) { v = fma(-l[t], s[t][j], v); }
                }
                l[j] = v / s[j][j];
            }
#pragma unroll
            for(int j = 0; j < NB; ++j) { M[r + (int64_t)(k0 + j) * ld] = l[j]; }
        }
        else
        {
            // column c of the block column right of the diagonal: u_ic = a_ic - sum_{t<i} l_it u_tc
            int64_t const c = k0 + NB + (int64_t)(blockIdx.x - nbl) * NB + tid;
            double u[NB];
#pragma unroll
            for(int i = 0; i < NB; ++i) { u[i] = M[(k0 + i) + c * ld]; }
#pragma unroll
            for(int i = 0; i < NB; ++i)
            {
                double v = u[i];
#pragma unroll
                for(int t = 0; t < NB; ++t)
                {
                    if(t < i) { v = fma(-s[i][t], u[t], v); }
                }
                u[i] = v;
            }
#pragma unroll
            for(int i = 0; i < NB; ++i) { M[(k0 + i) + c * ld] = u[i]; }
        }
    }

    // trailing update A22 -= L21 U12 with FP64 tensor cores: CTA = 128 x 128 tile of A22, 8 warps as 4 (rows) x 2 (columns), a
    // warp owns 32 x 64 = 4 x 8 fragments of m8n8 accumulators; the K = 64 panel is contracted in 16 mma.sync.m8n8k4 steps.
    // A (L21, column-major) and B (U12, column-major) are staged in shared memory one k-slab of 16 at a time.
    __global__ void __launch_bounds__(256) fr_lu_update(double* __restrict__ Mall, int64_t ld, int32_t k0)
    {
        __shared__ double sa[16][128 + 4];  // sa[k][row]
        __shared__ double sb[16][128 + 4];  // sb[k][col]
        double* const M = Mall + (int64_t)blockIdx.z * ld * ld;
        int64_t const r0 = k0 + NB + (int64_t)blockIdx.x * 128, c0 = k0 + NB + (int64_t)blockIdx.y * 128;
        int64_t const n_rem = ld - (k0 + NB);
        int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        int const wr = (warp & 3) * 32, wc = (warp >> 2) * 64;
        int64_t const rr_left = k0 + NB + n_rem - r0, cc_left = k0 + NB + n_rem - c0;
        int const rows_here = rr_left < 128 ? (int)rr_left : 128, cols_here = cc_left < 128 ? (int)cc_left : 128;
        double acc[4][8][2];
#pragma unroll
        for(int i = 0; i < 4; ++i)
#pragma unroll
            for(int j = 0; j < 8; ++j) { acc[i][j][0] = acc[i][j][1] = 0.0; }
        for(int ks = 0; ks < NB; ks += 16)
        {
            __syncthreads();
            for(int e = tid; e < 16 * 128; e += 256)
            {
                int const kk = e / 128, rr = e % 128;
                sa[kk][rr] = rr < rows_here ? M[(r0 + rr) + (int64_t)(k0 + ks + kk) * ld] : 0.0;
            }
            for(int e = tid; e < 16 * 128; e += 256)
            {
                int const cc = e / 16, kk = e % 16;
                sb[kk][cc] = cc < cols_here ? M[(k0 + ks + kk) + (c0 + cc) * ld] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for(int k4 = 0; k4 < 16; k4 += 4)
            {
                double a[4], b[8];
#pragma unroll
                for(int i = 0; i < 4; ++i) { a[i] = sa[k4 + (lane & 3)][wr + 8 * i + (lane >> 2)]; }
#pragma unroll
                for(int j = 0; j < 8; ++j) { b[j] = sb[k4 + (lane & 3)][wc + 8 * j + (lane >> 2)]; }
#pragma unroll
                for(int i = 0; i < 4; ++i)
#pragma unroll
                    for(int j = 0; j < 8; ++j)
                    {
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                     : "d"(a[i]), "d"(b[j]));
                    }
            }
        }
        // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1}
#pragma unroll
        for(int i = 0; i < 4; ++i)
#pragma unroll
            for(int j = 0; j < 8; ++j)
#pragma unroll
                for(int h = 0; h < 2; ++h)
                {
                    int const rr = wr + 8 * i + (lane >> 2), cc = wc + 8 * j + 2 * (lane & 3) + h;
                    if(rr < rows_here && cc < cols_here)
                    {
                        double* const p = M + (r0 + rr) + (c0 + cc) * ld;
                        *p = *p - acc[i][j][h];
                    }
                }
    }

    // Second version of the trailing update (the first one ran at 5 of the 37 TFLOP/s the DMMA pipe sustains -- tools/micro/
    // fp64_rate.cu -- because its epilogue was a chain of 64 dependent read-modify-write round trips per thread, C(i,j) -= acc,
    // that the compiler may not reorder, and because the tensor pipe idled during every slab load):
    //   * the C tile is loaded INTO the accumulators at the start (64 independent loads in flight) and the L21 fragments are
    //     negated, so that the DMMAs compute (-L21) U12 + C and the epilogue is 64 plain stores;
    //   * the k-slabs are double-buffered: slab s + 1 arrives by cp.async while slab s is contracted.
    // WN = 8-column fragments per warp: CTA tile = 128 rows x TC = 16 WN columns (8 warps as 4 x 2).  WN = 8: 128 x 128, 128
    // accumulator registers, one CTA per SM; WN = 4: 128 x 64, 64 accumulator registers, two CTAs per SM (the prologue and the
    // epilogue of one tile overlap the DMMAs of the other).
    template <int WN>
    constexpr int upd_smem() { return 2 * 16 * (132 + 16 * WN + 4) * (int)sizeof(double); }
    __device__ __forceinline__ void cp_async8(void* smem_dst, void const* src, bool valid)
    {
        uint32_t const d = (uint32_t)__cvta_generic_to_shared(smem_dst);
        int const n = valid ? 8 : 0;  // 0 source bytes: the destination is filled with zeros
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
    }
    template <int WN>
    __global__ void __launch_bounds__(256, WN == 8 ? 1 : 2) fr_lu_update2(double* __restrict__ Mall, int64_t ld, int32_t k0, int32_t part, int32_t ty)
    {
        constexpr int TC = 16 * WN, SBW = TC + 4;
        // part 0: every tile (grid tx x ty).  Look-ahead split: part 1 = the tiles of the first tile row and the first tile column
        // (1-D grid of ty + tx - 1), which hold the next step's panel; part 2 = all the others (grid tx - 1 x ty - 1)
        int bx = (int)blockIdx.x, by = (int)blockIdx.y;
        if(part == 1)
        {
            int const t = (int)blockIdx.x;
            bx = t < ty ? 0 : t - ty + 1;
            by = t < ty ? t : 0;
        }
        else if(part == 2)
        {
            ++bx;
            ++by;
        }
        extern __shared__ double upd_smem[];
        double(*const sa)[16][132] = reinterpret_cast<double(*)[16][132]>(upd_smem);                  // sa[buf][k][row]
        double(*const sb)[16][SBW] = reinterpret_cast<double(*)[16][SBW]>(upd_smem + 2 * 16 * 132);  // sb[buf][k][col]
        double* const M = Mall + (int64_t)blockIdx.z * ld * ld;
        int64_t const r0 = k0 + NB + (int64_t)bx * 128, c0 = k0 + NB + (int64_t)by * TC;
        int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        int const wr = (warp & 3) * 32, wc = (warp >> 2) * (8 * WN);
        int64_t const rr_left = ld - r0, cc_left = ld - c0;
        int const rows_here = rr_left < 128 ? (int)rr_left : 128, cols_here = cc_left < TC ? (int)cc_left : TC;
        double acc[4][WN][2];
        // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1}
#pragma unroll
        for(int i = 0; i < 4; ++i)
#pragma unroll
            for(int j = 0; j < WN; ++j)
#pragma unroll
                for(int h = 0; h < 2; ++h)
                {
                    int const rr = wr + 8 * i + (lane >> 2), cc = wc + 8 * j + 2 * (lane & 3) + h;
                    acc[i][j][h] = (rr < rows_here && cc < cols_here) ? M[(r0 + rr) + (c0 + cc) * ld] : 0.0;
                }
        // one k-slab (16 columns of L21, 16 rows of U12) straight into shared memory: 8-byte cp.async, zero fill past the edge
        auto fetch = [&](int ks, int buf)
        {
#pragma unroll
            for(int q = 0; q < 8; ++q)
            {
                int const e = tid + 256 * q;
                int const kk = e >> 7, rr = e & 127;
                bool const in = rr < rows_here;
                cp_async8(&sa[buf][kk][rr], in ? M + (r0 + rr) + (int64_t)(k0 + ks + kk) * ld : M, in);
            }
#pragma unroll
            for(int q = 0; q < TC / 16; ++q)
            {
                int const e = tid + 256 * q;
                int const cc = e >> 4, kk = e & 15;
                bool const in = cc < cols_here;
                cp_async8(&sb[buf][kk][cc], in ? M + (k0 + ks + kk) + (c0 + cc) * ld : M, in);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fetch(0, 0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 1
        for(int sl = 0; sl < NB / 16; ++sl)
        {
            int const buf = sl & 1;
            if(sl + 1 < NB / 16) { fetch((sl + 1) * 16, buf ^ 1); }  // in flight while this slab is contracted; the other buffer was last read before the barrier that ended the previous round
#pragma unroll
            for(int k4 = 0; k4 < 16; k4 += 4)
            {
                double a[4], b[WN];
#pragma unroll
                for(int i = 0; i < 4; ++i) { a[i] = -sa[buf][k4 + (lane & 3)][wr + 8 * i + (lane >> 2)]; }  // (-L21) U12 + C
#pragma unroll
                for(int j = 0; j < WN; ++j) { b[j] = sb[buf][k4 + (lane & 3)][wc + 8 * j + (lane >> 2)]; }
#pragma unroll
                for(int i = 0; i < 4; ++i)
#pragma unroll
                    for(int j = 0; j < WN; ++j)
                    {
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                     : "d"(a[i]), "d"(b[j]));
                    }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
#pragma unroll
        for(int i = 0; i < 4; ++i)
#pragma unroll
            for(int j = 0; j < WN; ++j)
#pragma unroll
                for(int h = 0; h < 2; ++h)
                {
                    int const rr = wr + 8 * i + (lane >> 2), cc = wc + 8 * j + 2 * (lane & 3) + h;
                    if(rr < rows_here && cc < cols_here) { M[(r0 + rr) + (c0 + cc) * ld] = acc[i][j][h]; }
                }
    }

    // forward (unit lower) and back (upper) substitution of one instance's core, one CTA per instance
    __global__ void __launch_bounds__(1024) fr_solve_core(double const* __restrict__ Mall, double* __restrict__ call, int64_t ld)
    {
        __shared__ double y[NB];
        double const* const M = Mall + (int64_t)blockIdx.x * ld * ld;
        double* const c = call + (int64_t)blockIdx.x * ld;
        int const tid = threadIdx.x;
        for(int64_t k0 = 0; k0 < ld; k0 += NB)
        {
            if(tid < 32)
            {
                // y = L11^-1 c_k: lane i owns rows i and i + 32
                double v0 = c[k0 + tid], v1 = c[k0 + tid + 32];
                for(int j = 0; j < NB; ++j)
                {
                    double const yj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31);
                    if(tid > j) { v0 = fma(-M[(k0 + tid) + (k0 + j) * ld], yj, v0); }
                    if(tid + 32 > j) { v1 = fma(-M[(k0 + tid + 32) + (k0 + j) * ld], yj, v1); }
                }
                y[tid] = v0;
                y[tid + 32] = v1;
                c[k0 + tid] = v0;
                c[k0 + tid + 32] = v1;
            }
            __syncthreads();
            for(int64_t r = k0 + NB + tid; r < ld; r += 1024)
            {
                double v = c[r];
#pragma unroll 8
                for(int j = 0; j < NB; ++j) { v = fma(-M[r + (k0 + j) * ld], y[j], v); }
                c[r] = v;
            }
            __syncthreads();
        }
        for(int64_t k0 = ld - NB; k0 >= 0; k0 -= NB)
        {
            if(tid < 32)
            {
                double v0 = c[k0 + tid], v1 = c[k0 + tid + 32];
                for(int j = NB - 1; j >= 0; --j)
                {
                    double const piv = M[(k0 + j) + (k0 + j) * ld];
                    double xj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31) / piv;
                    if(tid == (j & 31))
                    {
                        if(j < 32) { v0 = xj; }
                        else
                        {
                            v1 = xj;
                        }
                    }
                    if(tid < j) { v0 = fma(-M[(k0 + tid) + (k0 + j) * ld], xj, v0); }
                    if(tid + 32 < j) { v1 = fma(-M[(k0 + tid + 32) + (k0 + j) * ld], xj, v1); }
                }
                y[tid] = v0;
                y[tid + 32] = v1;
                c[k0 + tid] = v0;
                c[k0 + tid + 32] = v1;
            }
            __syncthreads();
            for(int64_t r = tid; r < k0; r += 1024)
            {
                double v = c[r];
#pragma unroll 8
                for(int j = 0; j < NB; ++j) { v = fma(-M[r + (k0 + j) * ld], y[j], v); }
                c[r] = v;
            }
            __syncthreads();
        }
    }

    // ---- the same substitutions, one block step per launch pair, so that the rows outside the diagonal block are spread over the
    // whole GPU (fr_solve_core walks the 456 MB of a 7 552-row factor with ONE CTA: 16.5 ms; this way it is a stream of small
    // launches: the factor is read once at HBM speed).  Arithmetic and its order per row are those of fr_solve_core.
    // y_k = L11^-1 c_k (unit lower): the 64 x 64 block is staged in shared memory by the whole CTA (all its loads in flight at
    // once; read from global inside the recurrence, every step waited for memory), then one warp runs the recurrence, lane i
    // owning rows i and i + 32
    __global__ void __launch_bounds__(256) fr_fwd_diag(double const* __restrict__ Mall, double* __restrict__ call, int64_t ld, int64_t k0)
    {
        __shared__ double s[NB][NB + 1];
        double const* const M = Mall + (int64_t)blockIdx.x * ld * ld;
        double* const c = call + (int64_t)blockIdx.x * ld;
        int const tid = threadIdx.x;
        {
            double t[NB * NB / 256];  // every load of the block in flight before the first store
#pragma unroll
            for(int q = 0; q < NB * NB / 256; ++q)
            {
                int const e = tid + 256 * q;
                t[q] = M[(k0 + e % NB) + (k0 + e / NB) * ld];
            }
#pragma unroll
            for(int q = 0; q < NB * NB / 256; ++q)
            {
                int const e = tid + 256 * q;
                s[e % NB][e / NB] = t[q];
            }
        }
        __syncthreads();
        if(tid >= 32) { return; }
        double v0 = c[k0 + tid], v1 = c[k0 + tid + 32];
        for(int j = 0; j < NB; ++j)
        {
            double const yj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31);
            if(tid > j) { v0 = fma(-s[tid][j], yj, v0); }
            if(tid + 32 > j) { v1 = fma(-s[tid + 32][j], yj, v1); }
        }
        c[k0 + tid] = v0;
        c[k0 + tid + 32] = v1;
    }
    // x_k = U11^-1 y_k
    __global__ void __launch_bounds__(256) fr_bwd_diag(double const* __restrict__ Mall, double* __restrict__ call, int64_t ld, int64_t k0)
    {
        __shared__ double s[NB][NB + 1];
        double const* const M = Mall + (int64_t)blockIdx.x * ld * ld;
        double* const c = call + (int64_t)blockIdx.x * ld;
        int const tid = threadIdx.x;
        {
            double t[NB * NB / 256];  // every load of the block in flight before the first store
#pragma unroll
            for(int q = 0; q < NB * NB / 256; ++q)
            {
                int const e = tid + 256 * q;
                t[q] = M[(k0 + e % NB) + (k0 + e / NB) * ld];
            }
#pragma unroll
            for(int q = 0; q < NB * NB / 256; ++q)
            {
                int const e = tid + 256 * q;
                s[e % NB][e / NB] = t[q];
            }
        }
        __syncthreads();
        if(tid >= 32) { return; }
        double v0 = c[k0 + tid], v1 = c[k0 + tid + 32];
        for(int j = NB - 1; j >= 0; --j)
        {
            double const piv = s[j][j];
            double xj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31) / piv;
            if(tid == (j & 31))
            {
                if(j < 32) { v0 = xj; }
                else
                {
                    v1 = xj;
                }
            }
            if(tid < j) { v0 = fma(-s[tid][j], xj, v0); }
            if(tid + 32 < j) { v1 = fma(-s[tid + 32][j], xj, v1); }
        }
        c[k0 + tid] = v0;
        c[k0 + tid + 32] = v1;
    }
    // c[r] -= sum_j M[r, k0 + j] c[k0 + j] for the rows r0 <= r < r1 (below the block: forward; above it: backward)
    __global__ void __launch_bounds__(256) fr_subst_update(double const* __restrict__ Mall, double* __restrict__ call, int64_t ld, int64_t k0, int64_t r0, int64_t r1)
    {
        __shared__ double y[NB];
        double const* const M = Mall + (int64_t)blockIdx.y * ld * ld;
        double* const c = call + (int64_t)blockIdx.y * ld;
        if(threadIdx.x < NB) { y[threadIdx.x] = c[k0 + threadIdx.x]; }
        __syncthreads();
        int64_t const r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if(r >= r1) { return; }
        double m[NB];  // the row's 64 entries in flight together
#pragma unroll
        for(int j = 0; j < NB; ++j) { m[j] = M[r + (k0 + j) * ld]; }
        double v = c[r];
#pragma unroll
        for(int j = 0; j < NB; ++j) { v = fma(-m[j], y[j], v); }
        c[r] = v;
    }

    __global__ void fr_core_scatter(pe_b200_frontal const f)
    {
        int64_t const inst = blockIdx.y;
        int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if(t >= f.n_core) { return; }
        f.x[(int64_t)f.core_unknown[t] * f.LSx + inst] = f.c[inst * f.ld_core + t];
    }

    thread_local char g_ferr[256] = "";
    int fchk(cudaError_t e, char const* what)
    {
        if(e == cudaSuccess) { return 0; }
        snprintf(g_ferr, sizeof(g_ferr), "%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
}  // namespace

extern "C"
{
    char const* pe_b200_frontal_last_error(void) { return g_ferr; }

    // launches of one frontal solve; *n_launches (optional) receives their number
    int pe_b200_frontal_run(pe_b200_frontal const* fp, void* stream, uint64_t* n_launches)
    {
        if(fp == nullptr || fp->n_inst <= 0) { return 0; }
        pe_b200_frontal const& f = *fp;
        cudaStream_t const st = (cudaStream_t)stream;
        uint64_t nl = 0;
        int64_t const B = f.B;
        auto blocks = [](int64_t n) { return (unsigned)((n + 255) / 256); };
        cudaEvent_t ev[4]{};
        bool const timing = f.phase_ms_host != nullptr;
        if(timing)
        {
            for(auto& e: ev) { cudaEventCreate(&e); }
            cudaEventRecord(ev[0], st);
        }
        if(fchk(cudaMemsetAsync(f.d, 0, sizeof(double) * (size_t)f.n_nodes * (size_t)B, st), "memset d") != 0) { return 1; }
        if(fchk(cudaMemsetAsync(f.z, 0, sizeof(double) * (size_t)f.n_nodes * (size_t)B, st), "memset z") != 0) { return 1; }
        if(fchk(cudaMemsetAsync(f.g, 0, sizeof(double) * (size_t)std::max(f.n_edges, 1) * (size_t)B, st), "memset g") != 0) { return 1; }
        fr_assemble<<<blocks((int64_t)(f.n_res + f.n_idc) * B), 256, 0, st>>>(f);
        ++nl;
        for(int32_t l = 0; l < f.n_levels; ++l)
        {
            int32_t const a = f.level_off_host[l], n = f.level_off_host[l + 1] - a;
            if(n <= 0) { continue; }
            fr_eliminate<<<blocks((int64_t)n * B), 256, 0, st>>>(f, a, n);
            ++nl;
        }
        int64_t const ld = f.ld_core;
        if(timing) { cudaEventRecord(ev[1], st); }
        if(ld > 0)
        {
            if(fchk(cudaMemsetAsync(f.M, 0, sizeof(double) * (size_t)ld * (size_t)ld * (size_t)f.n_inst, st), "memset core") != 0) { return 1; }
            fr_core_assemble<<<dim3(blocks(f.n_core + f.n_core_edges + (ld - f.n_core)), (unsigned)f.n_inst), 256, 0, st>>>(f);
            ++nl;
            static bool const update_v1 = std::getenv("PE_B200_FRONTAL_UPDATE_V1") != nullptr;  // the first trailing update, kept for A/B runs
            static bool const update_wide = std::getenv("PE_B200_FRONTAL_UPDATE_WIDE") != nullptr;  // 128 x 64 tiles, two CTAs per SM (default: 36.5 ms per solve of config A) / 128 x 128 tiles, one CTA per SM (40.0 ms)
            if(!update_v1 && (fchk(cudaFuncSetAttribute(fr_lu_update2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem<8>()), "fr_lu_update2 shared memory") != 0 ||
                              fchk(cudaFuncSetAttribute(fr_lu_update2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem<4>()), "fr_lu_update2 shared memory") != 0))
            {
                return 1;
            }
            // second stream + two events for the look-ahead of the default update (PE_B200_FRONTAL_NO_LOOKAHEAD: one stream)
            static bool const lookahead = std::getenv("PE_B200_FRONTAL_NO_LOOKAHEAD") == nullptr;
            cudaStream_t st2 = nullptr;
            cudaEvent_t e_panel{}, e_rest{};
            bool rest_pending = false;
            if(lookahead && !update_v1 && !update_wide)
            {
                if(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking) != cudaSuccess) { st2 = nullptr; }
                else if(cudaEventCreateWithFlags(&e_panel, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e_rest, cudaEventDisableTiming) != cudaSuccess)
                {
                    cudaStreamDestroy(st2);
                    st2 = nullptr;
                }
            }
            for(int64_t k0 = 0; k0 < ld; k0 += NB)
            {
                fr_lu_diag<<<dim3(1, 1, (unsigned)f.n_inst), NB, 0, st>>>(f.M, ld, (int32_t)k0, f.status);
                ++nl;
                int64_t const rem = ld - k0 - NB;
                if(rem <= 0) { break; }
                int32_t const nbl = (int32_t)(rem / NB);
                fr_lu_panel<<<dim3((unsigned)(2 * nbl), 1, (unsigned)f.n_inst), NB, 0, st>>>(f.M, ld, (int32_t)k0, nbl);
                unsigned const tb = (unsigned)((rem + 127) / 128);
                if(update_v1) { fr_lu_update<<<dim3(tb, tb, (unsigned)f.n_inst), 256, 0, st>>>(f.M, ld, (int32_t)k0); }
                else
                {
                    if(update_wide) { fr_lu_update2<8><<<dim3(tb, tb, (unsigned)f.n_inst), 256, upd_smem<8>(), st>>>(f.M, ld, (int32_t)k0, 0, (int32_t)tb); }
                    else
                    {
                        unsigned const ty = (unsigned)((rem + 63) / 64);
                        if(st2 != nullptr && tb > 1 && ty > 1)
                        {
                            // look-ahead: the first tile row / column on this stream (the next diagonal block and panel follow them),
                            // everything else on the second stream, beside that diagonal block and panel
                            cudaEventRecord(e_panel, st);
                            if(rest_pending) { cudaStreamWaitEvent(st, e_rest, 0); }  // the rest of the step before: its tiles are updated again now
                            fr_lu_update2<4><<<dim3(ty + tb - 1, 1, (unsigned)f.n_inst), 256, upd_smem<4>(), st>>>(f.M, ld, (int32_t)k0, 1, (int32_t)ty);
                            cudaStreamWaitEvent(st2, e_panel, 0);
                            fr_lu_update2<4><<<dim3(tb - 1, ty - 1, (unsigned)f.n_inst), 256, upd_smem<4>(), st2>>>(f.M, ld, (int32_t)k0, 2, (int32_t)ty);
                            cudaEventRecord(e_rest, st2);
                            rest_pending = true;
                            ++nl;
                        }
                        else
                        {
                            if(rest_pending)
                            {
                                cudaStreamWaitEvent(st, e_rest, 0);
                                rest_pending = false;
                            }
                            fr_lu_update2<4><<<dim3(tb, ty, (unsigned)f.n_inst), 256, upd_smem<4>(), st>>>(f.M, ld, (int32_t)k0, 0, (int32_t)ty);
                        }
                    }
                }
                nl += 2;
            }
            if(rest_pending) { cudaStreamWaitEvent(st, e_rest, 0); }
            if(st2 != nullptr)
            {
                cudaEventDestroy(e_panel);
                cudaEventDestroy(e_rest);
                cudaStreamDestroy(st2);  // returns at once; the stream goes when its work is done
            }
            if(timing) { cudaEventRecord(ev[2], st); }
            static bool const one_cta_subst = std::getenv("PE_B200_FRONTAL_ONE_CTA_SUBST") != nullptr;  // the first version, kept for A/B runs
            if(one_cta_subst)
            {
                fr_solve_core<<<(unsigned)f.n_inst, 1024, 0, st>>>(f.M, f.c, ld);
                ++nl;
            }
            else
            {
                for(int64_t k0 = 0; k0 < ld; k0 += NB)
                {
                    fr_fwd_diag<<<(unsigned)f.n_inst, 256, 0, st>>>(f.M, f.c, ld, k0);
                    ++nl;
                    if(k0 + NB < ld)
                    {
                        fr_subst_update<<<dim3(blocks(ld - k0 - NB), (unsigned)f.n_inst), 256, 0, st>>>(f.M, f.c, ld, k0, k0 + NB, ld);
                        ++nl;
                    }
                }
                for(int64_t k0 = ld - NB; k0 >= 0; k0 -= NB)
                {
                    fr_bwd_diag<<<(unsigned)f.n_inst, 256, 0, st>>>(f.M, f.c, ld, k0);
                    ++nl;
                    if(k0 > 0)
                    {
                        fr_subst_update<<<dim3(blocks(k0), (unsigned)f.n_inst), 256, 0, st>>>(f.M, f.c, ld, k0, 0, k0);
                        ++nl;
                    }
                }
            }
            fr_core_scatter<<<dim3(blocks(f.n_core), (unsigned)f.n_inst), 256, 0, st>>>(f);
            ++nl;
        }
        for(int32_t l = f.n_levels - 1; l >= 0; --l)
        {
            int32_t const a = f.level_off_host[l], n = f.level_off_host[l + 1] - a;
            if(n <= 0) { continue; }
            fr_back<<<blocks((int64_t)n * B), 256, 0, st>>>(f, a, n);
            ++nl;
        }
        if(n_launches != nullptr) { *n_launches = nl; }
        if(timing)
        {
            cudaEventRecord(ev[3], st);
            cudaEventSynchronize(ev[3]);
            for(int i = 0; i < 3; ++i)
            {
                float ms = 0.f;
                f.phase_ms_host[i] = cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess ? (double)ms : 0.0;
            }
            for(auto& e: ev) { cudaEventDestroy(e); }
        }
        return fchk(cudaGetLastError(), "frontal launch");
    }
}
