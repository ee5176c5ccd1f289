// stream.cpp — code generator of the STREAM kernel (run-time interface: csrc/pe_b200_stream.h).
//
// Input: the resident form of a real, linear program built with ONE stream per lane group (S = 1): its iter section
// (assembly + LU + substitution of one solve_once, circuit.h:987-1527) is a serial list of DOT / CAP_STEP ops over
// workspace rows.  Output: CUDA source in which that list is cut into TILES; a tile's cold operand rows arrive in a
// shared-memory ring by bulk asynchronous copies issued NS tiles ahead, everything produced inside a tile or handed to the
// next tile travels in registers, periodic runs of tiles (the elimination of a chain of like nodes) are rolled into
// loops.  stream_prepare() first re-lays the workspace rows out so that the rows one tile fetches are contiguous (one
// bulk copy per tile instead of one per row: 64 issue cycles each on the consuming warp), adding read-only replica rows
// where two sweeps need the same parameter.  The arithmetic of every op is the interpreter's, operation for operation.
#include "pe_host.hpp"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

namespace pe_b200
{
    namespace
    {
        constexpr std::uint32_t k_slot_mask{0x7fffu};
        constexpr int k_max_tile_rows{64};   // ring rows one tile may fetch (config B: 32 / 44 / 56 / 64 / 80 rows -> 43.0 / 46.2 / 47.5 / 49.1 / 46.4 M solves/s)
        constexpr int k_max_tile_ops{64};
        constexpr int k_max_single_ops{12};  // ops per tile outside the periodic runs
        constexpr int k_max_inv{8};

        int env_int(char const* name, int dflt)
        {
            char const* v{std::getenv(name)};
            return v != nullptr ? std::atoi(v) : dflt;
        }

        struct sop
        {
            rop const* o{};
            std::vector<std::uint32_t> rd, wr;  // slots
            bool steady_cap{};  // CAP_STEP in the steady variant: prev_g equals 2 C / dt already and is not stored again
            bool steady_load{};  // ... and is fetched instead of being recomputed
        };

        // rows written by the step section (a replica of such a row would go stale between two solves)
        std::vector<std::uint8_t> step_written(program const& pr)
        {
            std::vector<std::uint8_t> w(PE_R_MAX_SLOTS + 1, 0);
            for(auto const& ph: pr.rstreams[0].sec[1])
            {
                for(auto const& o: ph)
                {
                    if(o.bubble) { continue; }
                    if(o.opcode == PE_OP_DOT || o.opcode == PE_OP_CDOT) { w[o.dst & k_slot_mask] = 1; }
                    else
                    {
                        for(auto const f: o.opnd) { w[f & k_slot_mask] = 1; }  // conservative: outputs come first, inputs are marked too
                    }
                }
            }
            return w;
        }

        bool simple_value_op(std::uint32_t op) { return op == PE_OP_RECIP || op == PE_OP_MUL || op == PE_OP_SUB || op == PE_OP_COPY || op == PE_OP_MUL2DIV; }

        // ops of one section in program order (one stream: the phases simply follow each other)
        bool flatten(program const& pr, int sec, std::vector<sop>& ops, bool steady = false)
        {
            ops.clear();
            if(pr.rstreams.size() != 1) { return false; }
            for(auto const& ph: pr.rstreams[0].sec[sec])
            {
                for(auto const& o: ph)
                {
                    if(o.bubble) { continue; }
                    sop j;
                    j.o = &o;
                    if(o.opcode == PE_OP_DOT)
                    {
                        // the Newton-test flags of the solution ops are inert here: the stream kernel runs linear programs only
                        for(auto const s: o.sre) { j.rd.push_back(s & k_slot_mask); }
                        for(auto const& pp: o.pp)
                        {
                            j.rd.push_back(pp.first & k_slot_mask);
                            j.rd.push_back(pp.second & k_slot_mask);
                        }
                        if(o.flags & PE_F_SCALE) { j.rd.push_back(o.scale & k_slot_mask); }
                        j.wr.push_back(o.dst & k_slot_mask);
                    }
                    else if(o.opcode == PE_OP_CAP_STEP && o.opnd.size() == 6)
                    {
                        for(auto const w: o.opnd) { j.rd.push_back(w & k_slot_mask); }
                        j.wr.push_back(o.opnd[0] & k_slot_mask);
                        j.wr.push_back(o.opnd[1] & k_slot_mask);
                    }
                    else if(o.opcode == PE_OP_VSIN && o.opnd.size() == 4)
                    {
                        // [dst] <- [Vp][omega][phase] at the time of the solve (VAC.h:176, IAC.h:154)
                        for(std::size_t i{1}; i < 4; ++i) { j.rd.push_back(o.opnd[i] & k_slot_mask); }
                        j.wr.push_back(o.opnd[0] & k_slot_mask);
                    }
                    else if(o.opcode == PE_OP_GEN_EVAL && o.opnd.size() == 10)
                    {
                        // [dst] <- [kind][tsel][Vh][Vl][freq][duty][phase][tr][tf] at the time of the solve (generator/*.h)
                        for(std::size_t i{1}; i < 10; ++i) { j.rd.push_back(o.opnd[i] & k_slot_mask); }
                        j.wr.push_back(o.opnd[0] & k_slot_mask);
                    }
                    else if(o.opcode == PE_OP_IND_STEP && o.opnd.size() == 7)
                    {
                        // [req][ueq] <- [L][dt][va][vb][ib] (inductor.h:134-160)
                        for(std::size_t i{2}; i < 7; ++i) { j.rd.push_back(o.opnd[i] & k_slot_mask); }
                        j.wr.push_back(o.opnd[0] & k_slot_mask);
                        j.wr.push_back(o.opnd[1] & k_slot_mask);
                    }
                    else if(simple_value_op(o.opcode) && o.opnd.size() == (o.opcode == PE_OP_RECIP || o.opcode == PE_OP_COPY ? 2u : 3u))
                    {
                        for(std::size_t i{1}; i < o.opnd.size(); ++i) { j.rd.push_back(o.opnd[i] & k_slot_mask); }
                        j.wr.push_back(o.opnd[0] & k_slot_mask);
                    }
                    else
                    {
                        if(std::getenv("PE_B200_STREAM_DEBUG") != nullptr) { std::fprintf(stderr, "stream: section %d op %u (%zu operands) is not covered\n", sec, o.opcode, o.opnd.size()); }
                        return false;
                    }
                    ops.push_back(std::move(j));
                }
            }
            if(steady)
            {
                // Steady variant of the iter section (every solve of a launch but the first): the trapezoidal companion of a
                // capacitor whose C and dt no op of the step / iter sections writes finds prev_g == 2 C / dt (the first solve
                // of the launch stored exactly that, capacitor.h:106-128), so the row is neither fetched nor stored again.
                std::vector<std::uint8_t> w(PE_R_MAX_SLOTS + 1, 0);
                for(auto const& o: ops)
                {
                    for(auto const sl: o.wr) { w[sl] = 1; }
                }
                auto const sw{step_written(pr)};
                for(auto& o: ops)
                {
                    if(o.o->opcode != PE_OP_CAP_STEP) { continue; }
                    auto const c{o.rd[2]}, dt{o.rd[3]};
                    if(w[c] || w[dt] || sw[c] || sw[dt] || o.rd[1] == o.rd[0]) { continue; }
                    o.steady_cap = true;
                    // mode 0: prev_g (= 2 C / dt) is recomputed, its row is not fetched; mode 1: the row is fetched and stands in
                    // for the division (8 bytes more per capacitor and solve, ~40 instructions less on the warp's serial path)
                    static int const mode{env_int("PE_B200_STREAM_STEADY_MODE", 0)};  // measured on config B: 49.1 M solves/s (mode 0) vs 46.5 (mode 1)
                    o.steady_load = mode == 1;
                    if(!o.steady_load) { o.rd.erase(o.rd.begin() + 1); }
                }
            }
            return !ops.empty();
        }

        bool same_shape(sop const& a, sop const& b)
        {
            rop const &x{*a.o}, &y{*b.o};
            if(a.steady_cap != b.steady_cap || a.steady_load != b.steady_load || x.opcode != y.opcode || x.flags != y.flags || a.rd.size() != b.rd.size() || a.wr.size() != b.wr.size() || x.sre.size() != y.sre.size() || x.pp.size() != y.pp.size())
            {
                return false;
            }
            for(std::size_t i{}; i < x.sre.size(); ++i)
            {
                if((x.sre[i] ^ y.sre[i]) & PE_R_NEG) { return false; }
            }
            for(std::size_t i{}; i < x.pp.size(); ++i)
            {
                if(((x.pp[i].first ^ x.pp[i].second) ^ (y.pp[i].first ^ y.pp[i].second)) & PE_R_NEG) { return false; }
            }
            if(x.opnd.size() != y.opnd.size()) { return false; }
            for(std::size_t i{}; i < x.opnd.size(); ++i)
            {
                if((x.opnd[i] ^ y.opnd[i]) & PE_R_NEG) { return false; }
            }
            return true;
        }

        // op a + m * per repeats op a with every row shifted by m * (the shift between a and a + per)
        bool shifted(std::vector<sop> const& ops, int a, int per, int m)
        {
            sop const &x{ops[static_cast<std::size_t>(a)]}, &y{ops[static_cast<std::size_t>(a + per)]}, &z{ops[static_cast<std::size_t>(a + m * per)]};
            if(!same_shape(x, z)) { return false; }
            auto ok = [&](std::uint32_t s0, std::uint32_t s1, std::uint32_t sm_) -> bool
            {
                std::int64_t const d{static_cast<std::int64_t>(s1) - static_cast<std::int64_t>(s0)};
                return static_cast<std::int64_t>(sm_) == static_cast<std::int64_t>(s0) + m * d;
            };
            for(std::size_t i{}; i < x.rd.size(); ++i)
            {
                if(!ok(x.rd[i], y.rd[i], z.rd[i])) { return false; }
            }
            for(std::size_t i{}; i < x.wr.size(); ++i)
            {
                if(!ok(x.wr[i], y.wr[i], z.wr[i])) { return false; }
            }
            return true;
        }

        // longest periodic run starting at k0: period per (ops), n_it repetitions; returns the covered ops (0 = none)
        int find_run(std::vector<sop> const& ops, int k0, int& per, int& n_it)
        {
            int const n{static_cast<int>(ops.size())};
            int best{};
            for(int p{1}; p <= 16 && k0 + 2 * p <= n; ++p)
            {
                if(!same_shape(ops[static_cast<std::size_t>(k0)], ops[static_cast<std::size_t>(k0 + p)])) { continue; }
                int m{1};
                for(;; ++m)
                {
                    if(k0 + (m + 1) * p > n) { break; }
                    bool all{true};
                    for(int q{}; q < p && all; ++q) { all = shifted(ops, k0 + q, p, m); }
                    if(!all) { break; }
                }
                if(m >= 8 && m * p > best)
                {
                    best = m * p;
                    per = p;
                    n_it = m;
                }
            }
            return best;
        }

        enum cls_t : std::uint8_t
        {
            C_INV = 0,   // launch-invariant row held in a register for the whole solve
            C_SAME = 1,  // produced earlier in the same tile
            C_PREV = 2,  // produced by the previous tile
            C_MEM = 3,   // fetched into the ring by a bulk copy
        };
        struct rsrc
        {
            cls_t cls{C_MEM};
            int kw{-1}, ow{};  // SAME / PREV: writer op and its output index
            int wt{-1};        // MEM: tile of the last writer inside this section (-1: an earlier solve)
        };

        struct region
        {
            int op0{};    // first op of the first tile
            int per{};    // ops per tile
            int n{1};     // tiles (1 = a single tile, emitted straight-line)
            int tile0{};  // global index of the first tile
            bool loop() const { return n > 1; }
        };

        struct plan
        {
            std::vector<sop> ops;
            std::vector<region> regs;
            std::vector<int> tile_of;               // per op
            std::vector<int> op_run;                // per op: periodic run it belongs to (-1 = none)
            std::vector<std::array<int, 3>> runs;   // (first op, period in ops, periods)
            std::vector<int> reg_of_tile;           // per tile
            std::vector<std::vector<rsrc>> src;     // per op, per read
            std::vector<std::uint32_t> inv;         // invariant rows, in register order
            std::map<std::uint32_t, int> inv_idx;
            std::vector<std::vector<std::uint32_t>> mem;  // per tile: sorted distinct MEM rows
            std::vector<int> wmax;                  // per tile: latest writer tile of its MEM rows (-1 = none)
            std::vector<std::uint8_t> written;      // per slot: written by the iter section
            int n_tiles{};
            int stage_rows{};
        };

        // cut the op list into regions / tiles
        void build_regions(plan& p, std::set<std::uint32_t> const& inv)
        {
            auto const& ops{p.ops};
            int const n{static_cast<int>(ops.size())};
            p.regs.clear();
            p.op_run.assign(ops.size(), -1);
            p.runs.clear();
            int const max_rows{std::clamp(env_int("PE_B200_STREAM_ROWS", k_max_tile_rows), 4, 96)};
            auto singles = [&](int a, int b)
            {
                while(a < b)
                {
                    // greedy: up to k_max_single_ops ops and max_rows distinct cold rows
                    std::set<std::uint32_t> rows, wr;
                    int e{a};
                    while(e < b && e - a < k_max_single_ops)
                    {
                        std::set<std::uint32_t> add;
                        for(auto const s: ops[static_cast<std::size_t>(e)].rd)
                        {
                            if(inv.count(s) == 0 && wr.count(s) == 0 && rows.count(s) == 0) { add.insert(s); }
                        }
                        if(e > a && static_cast<int>(rows.size() + add.size()) > max_rows) { break; }
                        rows.insert(add.begin(), add.end());
                        for(auto const s: ops[static_cast<std::size_t>(e)].wr) { wr.insert(s); }
                        ++e;
                    }
                    region r;
                    r.op0 = a;
                    r.per = e - a;
                    r.n = 1;
                    p.regs.push_back(r);
                    a = e;
                }
            };
            int k{}, pend{};
            while(k < n)
            {
                int per{}, n_it{};
                if(find_run(ops, k, per, n_it) > 0)
                {
                    // periods per tile: as many as keep the cold rows of a tile (rows read before this run's code wrote them) within
                    // the ring stage
                    auto cold_rows = [&](int U_) -> int
                    {
                        std::set<std::uint32_t> wr, cold;
                        for(int q{}; q < per; ++q)
                        {
                            for(auto const s: ops[static_cast<std::size_t>(k + q)].wr) { wr.insert(s); }
                        }
                        for(int q{per}; q < (1 + U_) * per && k + q < n; ++q)
                        {
                            for(auto const s: ops[static_cast<std::size_t>(k + q)].rd)
                            {
                                if(inv.count(s) == 0 && wr.count(s) == 0) { cold.insert(s); }
                            }
                            for(auto const s: ops[static_cast<std::size_t>(k + q)].wr) { wr.insert(s); }
                        }
                        return static_cast<int>(cold.size());
                    };
                    int U{std::clamp(std::min(16, k_max_tile_ops / per), 1, std::max(1, n_it - 1))};
                    while(U > 1 && cold_rows(U) > max_rows) { --U; }
                    int const nt{n_it / U};
                    if(nt >= 4)
                    {
                        for(int q{}; q < n_it * per; ++q) { p.op_run[static_cast<std::size_t>(k + q)] = static_cast<int>(p.runs.size()); }
                        p.runs.push_back({k, per, n_it});
                        singles(k - pend, k);
                        pend = 0;
                        // the first tile of the run is peeled (its operands come from the code before the run)
                        region r0;
                        r0.op0 = k;
                        r0.per = per * U;
                        r0.n = 1;
                        p.regs.push_back(r0);
                        region r;
                        r.op0 = k + per * U;
                        r.per = per * U;
                        r.n = nt - 1;
                        p.regs.push_back(r);
                        k += nt * per * U;
                        continue;
                    }
                }
                ++pend;
                ++k;
            }
            singles(n - pend, n);
        }

        void number_tiles(plan& p)
        {
            p.tile_of.assign(p.ops.size(), -1);
            p.reg_of_tile.clear();
            int t{};
            for(std::size_t r{}; r < p.regs.size(); ++r)
            {
                auto& rg{p.regs[r]};
                rg.tile0 = t;
                for(int it{}; it < rg.n; ++it, ++t)
                {
                    for(int q{}; q < rg.per; ++q) { p.tile_of[static_cast<std::size_t>(rg.op0 + it * rg.per + q)] = t; }
                    p.reg_of_tile.push_back(static_cast<int>(r));
                }
            }
            p.n_tiles = t;
        }

        // where every read takes its value from
        void classify(plan& p)
        {
            auto const& ops{p.ops};
            p.src.assign(ops.size(), {});
            p.mem.assign(static_cast<std::size_t>(p.n_tiles), {});
            p.wmax.assign(static_cast<std::size_t>(p.n_tiles), -1);
            std::map<std::uint32_t, std::pair<int, int>> last_w;
            std::vector<std::set<std::uint32_t>> mem_set(static_cast<std::size_t>(p.n_tiles));
            for(std::size_t k{}; k < ops.size(); ++k)
            {
                int const T{p.tile_of[k]};
                for(auto const s: ops[k].rd)
                {
                    rsrc r;
                    auto const w{last_w.find(s)};
                    if(p.inv_idx.count(s) != 0) { r.cls = C_INV; }
                    else if(w != last_w.end() && p.tile_of[static_cast<std::size_t>(w->second.first)] == T)
                    {
                        r.cls = C_SAME;
                        r.kw = w->second.first;
                        r.ow = w->second.second;
                    }
                    else if(w != last_w.end() && p.tile_of[static_cast<std::size_t>(w->second.first)] == T - 1)
                    {
                        r.cls = C_PREV;
                        r.kw = w->second.first;
                        r.ow = w->second.second;
                    }
                    else
                    {
                        r.cls = C_MEM;
                        r.wt = w != last_w.end() ? p.tile_of[static_cast<std::size_t>(w->second.first)] : -1;
                        mem_set[static_cast<std::size_t>(T)].insert(s);
                        p.wmax[static_cast<std::size_t>(T)] = std::max(p.wmax[static_cast<std::size_t>(T)], r.wt);
                    }
                    p.src[k].push_back(r);
                }
                for(std::size_t i{}; i < ops[k].wr.size(); ++i) { last_w[ops[k].wr[i]] = {static_cast<int>(k), static_cast<int>(i)}; }
            }
            p.stage_rows = 1;
            for(int t{}; t < p.n_tiles; ++t)
            {
                p.mem[static_cast<std::size_t>(t)].assign(mem_set[static_cast<std::size_t>(t)].begin(), mem_set[static_cast<std::size_t>(t)].end());
                p.stage_rows = std::max(p.stage_rows, static_cast<int>(p.mem[static_cast<std::size_t>(t)].size()));
            }
        }

        // a loop region is uniform when every iteration takes every operand from the same place as the first one
        bool uniform(plan const& p, region const& rg)
        {
            for(int it{1}; it < rg.n; ++it)
            {
                for(int q{}; q < rg.per; ++q)
                {
                    auto const& a{p.src[static_cast<std::size_t>(rg.op0 + q)]};
                    auto const& b{p.src[static_cast<std::size_t>(rg.op0 + it * rg.per + q)]};
                    for(std::size_t i{}; i < a.size(); ++i)
                    {
                        if(a[i].cls != b[i].cls) { return false; }
                        if(a[i].cls == C_SAME || a[i].cls == C_PREV)
                        {
                            if(a[i].ow != b[i].ow || a[i].kw + it * rg.per != b[i].kw) { return false; }
                        }
                    }
                }
                // the rows fetched move rigidly: same count, row j of iteration `it` = row j of iteration 0 + it * d_j
                auto const& m0{p.mem[static_cast<std::size_t>(rg.tile0)]};
                auto const& m1{p.mem[static_cast<std::size_t>(rg.tile0 + 1)]};
                auto const& mi{p.mem[static_cast<std::size_t>(rg.tile0 + it)]};
                if(mi.size() != m0.size()) { return false; }
                for(std::size_t j{}; j < m0.size(); ++j)
                {
                    std::int64_t const d{static_cast<std::int64_t>(m1[j]) - static_cast<std::int64_t>(m0[j])};
                    if(static_cast<std::int64_t>(mi[j]) != static_cast<std::int64_t>(m0[j]) + it * d) { return false; }
                }
            }
            // no tile of the loop may fetch a row a tile of the loop has written (the copies run ahead of the stores)
            for(int it{}; it < rg.n; ++it)
            {
                if(p.wmax[static_cast<std::size_t>(rg.tile0 + it)] >= rg.tile0) { return false; }
            }
            // a carried value must come from the same op position in the peeled tile before the loop
            for(int q{}; q < rg.per; ++q)
            {
                for(auto const& r: p.src[static_cast<std::size_t>(rg.op0 + q)])
                {
                    if(r.cls == C_PREV && (r.kw < rg.op0 - rg.per || r.kw >= rg.op0)) { return false; }
                }
            }
            return true;
        }

        // invariant rows: never written by the iter section, read by many ops
        std::vector<std::uint32_t> pick_invariants(std::vector<sop> const& ops)
        {
            std::map<std::uint32_t, int> cnt;
            std::set<std::uint32_t> wr;
            for(auto const& o: ops)
            {
                for(auto const s: o.wr) { wr.insert(s); }
                for(auto const s: o.rd) { ++cnt[s]; }
            }
            std::vector<std::pair<int, std::uint32_t>> cand;
            for(auto const& [s, c]: cnt)
            {
                if(c >= 8 && wr.count(s) == 0) { cand.push_back({c, s}); }
            }
            std::sort(cand.begin(), cand.end(), [](auto const& a, auto const& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; });
            std::vector<std::uint32_t> out;
            for(std::size_t i{}; i < cand.size() && i < static_cast<std::size_t>(k_max_inv); ++i) { out.push_back(cand[i].second); }
            return out;
        }

        bool make_plan(program const& pr, plan& p, int sec = 2, bool steady = false)
        {
            if(!flatten(pr, sec, p.ops, steady)) { return false; }
            p.inv = pick_invariants(p.ops);
            p.inv_idx.clear();
            for(std::size_t i{}; i < p.inv.size(); ++i) { p.inv_idx[p.inv[i]] = static_cast<int>(i); }
            std::set<std::uint32_t> const invs(p.inv.begin(), p.inv.end());
            build_regions(p, invs);
            bool const dbg{std::getenv("PE_B200_STREAM_DEBUG") != nullptr};
            if(dbg)
            {
                std::fprintf(stderr, "stream: %zu ops, %zu invariants, regions:", p.ops.size(), p.inv.size());
                for(auto const& rg: p.regs) { std::fprintf(stderr, " [%d+%dx%d]", rg.op0, rg.n, rg.per); }
                std::fprintf(stderr, "\n");
            }
            for(int round{}; round < 8; ++round)
            {
                number_tiles(p);
                classify(p);
                bool changed{};
                std::vector<region> next;
                for(auto const& rg: p.regs)
                {
                    if(rg.loop() && !uniform(p, rg))
                    {
                        if(dbg) { std::fprintf(stderr, "stream: loop at op %d (%d x %d) is not uniform\n", rg.op0, rg.n, rg.per); }
                        for(int it{}; it < rg.n; ++it)
                        {
                            region s;
                            s.op0 = rg.op0 + it * rg.per;
                            s.per = rg.per;
                            s.n = 1;
                            next.push_back(s);
                        }
                        changed = true;
                    }
                    else
                    {
                        next.push_back(rg);
                    }
                }
                if(!changed) { break; }
                p.regs = std::move(next);  // the tiles stay the same: only the form of the code changes
            }
            number_tiles(p);
            classify(p);
            p.written.assign(PE_R_MAX_SLOTS + 1, 0);
            for(auto const& o: p.ops)
            {
                for(auto const s: o.wr) { p.written[s] = 1; }
            }
            return p.stage_rows <= 96;
        }

        // slots read outside the generated code: by the interpreted sections, the load / store tables, the solution read-out
        std::vector<std::uint8_t> kept_slots(program const& pr)
        {
            std::vector<std::uint8_t> keep(PE_R_MAX_SLOTS + 1, 0);
            auto k = [&](std::uint32_t f) { keep[f & k_slot_mask] = 1; };
            for(int sec{}; sec < 3; ++sec)
            {
                for(auto const& ph: pr.rstreams[0].sec[sec])
                {
                    for(auto const& o: ph)
                    {
                        if(o.bubble || sec == 2) { continue; }
                        k(o.dst);
                        k(o.scale);
                        for(auto const w: o.sre) { k(w); }
                        for(auto const w: o.sim) { k(w); }
                        for(auto const& pp: o.pp)
                        {
                            k(pp.first);
                            k(pp.second);
                        }
                        for(auto const w: o.opnd) { k(w); }
                        for(auto const& sb: o.sub) { k(sb.dst); }
                    }
                }
            }
            for(auto const& io: pr.io) { k(io.slot_kind & 0xffffu); }
            for(auto const x: pr.x_slot) { k(x); }
            return keep;
        }

        // ---- code generation -------------------------------------------------------------------------------------
        struct emitter
        {
            plan const& p;
            std::string sfx;  // "_iter" / "_prep"
            std::vector<std::uint8_t> needed;  // per slot: some load takes it from memory, or it is read outside this code
            std::ostringstream out;

            std::string valname(int k, int o) const
            {
                int const T{p.tile_of[static_cast<std::size_t>(k)]};
                auto const& rg{p.regs[static_cast<std::size_t>(p.reg_of_tile[static_cast<std::size_t>(T)])]};
                int const q{(k - rg.op0) % rg.per};
                std::ostringstream s;
                if(rg.loop()) { s << "r" << p.reg_of_tile[static_cast<std::size_t>(T)] << "_w" << q << "_" << o; }
                else
                {
                    s << "t" << T << "_w" << q << "_" << o;
                }
                return s.str();
            }

            static std::string op_text(rop const& o, std::vector<std::string> const& x, std::vector<std::string> const& w, std::vector<std::string> const& st, bool declare, int steady_cap = 0)
            {
                std::ostringstream t;
                char const* const dv{declare ? "jv " : ""};
                if(o.opcode == PE_OP_DOT)
                {
                    std::size_t q{};
                    t << dv << w[0] << " = jzero();";
                    for(auto const s: o.sre) { t << ((s & PE_R_NEG) ? " jsub(" : " jadd(") << w[0] << ", " << x[q++] << ");"; }
                    for(auto const& pp: o.pp)
                    {
                        bool const pos{((pp.first ^ pp.second) & PE_R_NEG) != 0u};  // -(+-a)(+-b)
                        t << (pos ? " jfma(" : " jfms(") << w[0] << ", " << x[q] << ", " << x[q + 1] << ");";
                        q += 2;
                    }
                    if(o.flags & PE_F_SCALE) { t << " jmul(" << w[0] << ", " << x[q++] << ");"; }
                    if((o.flags & PE_F_SCALE) && (o.flags & PE_F_GUARD)) { t << " jguard(" << w[0] << ", k.guard, fm);"; }  // pe_b200_program.h: an entry of L out of bounds
                    if(o.flags & PE_F_RECIP) { t << " jrcp(" << w[0] << ", fm);"; }
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                else if(o.opcode == PE_OP_CAP_STEP && steady_cap == 2)  // operands [hist][prev_g][C][dt][va][vb]; prev_g == 2 C / dt is taken as loaded
                {
                    t << dv << w[0] << " = " << x[0] << "; " << dv << w[1] << " = " << x[1] << "; jcap_loaded(" << x[4] << ", " << x[5] << ", " << w[0] << ", " << w[1] << ");";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                else if(o.opcode == PE_OP_CAP_STEP && steady_cap == 1)  // operands [hist][C][dt][va][vb]; prev_g := 2 C / dt, kept in its register only
                {
                    t << dv << w[0] << " = " << x[0] << "; " << dv << w[1] << " = jzero(); jcap_steady(" << x[1] << ", " << x[2] << ", " << x[3] << ", " << x[4] << ", " << w[0] << ", " << w[1] << ");";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                else if(o.opcode == PE_OP_CAP_STEP)  // [hist][prev_g][C][dt][va][vb]
                {
                    t << dv << w[0] << " = " << x[0] << "; " << dv << w[1] << " = " << x[1] << "; jcap(" << x[2] << ", " << x[3] << ", " << x[4] << ", " << x[5] << ", " << w[0] << ", " << w[1]
                      << ");";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                    if(!st[1].empty()) { t << " sk_st(k, " << st[1] << ", " << w[1] << ");"; }
                }
                else if(o.opcode == PE_OP_VSIN)  // [dst] <- [Vp][omega][phase]
                {
                    t << dv << w[0] << " = jvsin(" << x[0] << ", " << x[1] << ", " << x[2] << ", k.t);";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                else if(o.opcode == PE_OP_GEN_EVAL)  // [dst] <- [kind][tsel][Vh][Vl][freq][duty][phase][tr][tf]
                {
                    t << dv << w[0] << " = jvgen(";
                    for(std::size_t i{}; i < 9; ++i) { t << x[i] << ", "; }
                    t << "k.t);";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                else if(o.opcode == PE_OP_IND_STEP)  // [req][ueq] <- [L][dt][va][vb][ib]
                {
                    t << dv << w[0] << " = jzero(); " << dv << w[1] << " = jzero(); jind(" << x[0] << ", " << x[1] << ", " << x[2] << ", " << x[3] << ", " << x[4] << ", " << w[0] << ", " << w[1] << ");";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                    if(!st[1].empty()) { t << " sk_st(k, " << st[1] << ", " << w[1] << ");"; }
                }
                else  // simple value ops: [dst][a]([b]), operands carry a negate bit
                {
                    auto arg = [&](std::size_t i) { return (o.opnd[i + 1] & PE_R_NEG) ? "jneg(" + x[i] + ")" : x[i]; };
                    char const* fn{o.opcode == PE_OP_RECIP ? "jvrecip" : o.opcode == PE_OP_MUL ? "jvmul" : o.opcode == PE_OP_SUB ? "jvsub" : o.opcode == PE_OP_COPY ? "jvcopy" : "jvmul2div"};
                    t << dv << w[0] << " = " << fn << "(" << arg(0);
                    if(x.size() > 1) { t << ", " << arg(1); }
                    t << ");";
                    if(!st[0].empty()) { t << " sk_st(k, " << st[0] << ", " << w[0] << ");"; }
                }
                return t.str();
            }

            // first tile p > t whose cold rows were written by tile >= a (it may only be issued once those stores are done)
            int limit_after(int a, int t) const
            {
                for(int q{t + 1}; q < p.n_tiles; ++q)
                {
                    if(p.wmax[static_cast<std::size_t>(q)] >= a) { return q; }
                }
                return p.n_tiles;
            }

            // copies of tile t: (ring row, first workspace row, rows)
            std::vector<std::array<std::uint32_t, 3>> copies(int t) const
            {
                std::vector<std::array<std::uint32_t, 3>> c;
                auto const& m{p.mem[static_cast<std::size_t>(t)]};
                for(std::size_t j{}; j < m.size(); ++j)
                {
                    if(!c.empty() && c.back()[1] + c.back()[2] == m[j]) { ++c.back()[2]; }
                    else
                    {
                        c.push_back({static_cast<std::uint32_t>(j), m[j], 1u});
                    }
                }
                return c;
            }

            void emit_tile_head(std::string const& tvar, std::string const& lim, std::string const& ind)
            {
                out << ind << "if(k.pn <= " << tvar << ") { pe_stream_produce_upto" << sfx << "(k, " << tvar << " + 1u, (int32_t)" << tvar << " - 1); }\n";
                out << ind << "sk_wait(k, " << tvar << ");\n";
                out << ind << "auto const sg = sk_stage(k, " << tvar << ");\n";
                (void)lim;
            }

            void emit_single(int ri)
            {
                auto const& rg{p.regs[static_cast<std::size_t>(ri)]};
                int const T{rg.tile0};
                auto const& m{p.mem[static_cast<std::size_t>(T)]};
                out << "    // tile " << T << ": ops " << rg.op0 << " .. " << (rg.op0 + rg.per - 1) << "\n";
                out << "    {\n";
                emit_tile_head(std::to_string(T) + "u", "", "        ");
                for(std::size_t j{}; j < m.size(); ++j) { out << "        jv const m" << j << " = sk_ring(sg, " << j << "u);\n"; }
                out << "        sk_ring_done(k, " << T << "u);\n";
                out << "        { uint32_t tg = " << T << "u + k.ns_mask + 2u; if(tg > " << limit_after(T, T) << "u) { tg = " << limit_after(T, T) << "u; } pe_stream_produce_upto" << sfx << "(k, tg, " << (T - 1)
                    << "); }\n";
                std::map<std::uint32_t, std::size_t> row_of;
                for(std::size_t j{}; j < m.size(); ++j) { row_of[m[j]] = j; }
                std::ostringstream body;
                for(int q{}; q < rg.per; ++q)
                {
                    int const kk{rg.op0 + q};
                    auto const& j{p.ops[static_cast<std::size_t>(kk)]};
                    std::vector<std::string> x, w, st;
                    for(std::size_t i{}; i < j.rd.size(); ++i)
                    {
                        auto const& r{p.src[static_cast<std::size_t>(kk)][i]};
                        if(r.cls == C_INV) { x.push_back("inv" + std::to_string(p.inv_idx.at(j.rd[i]))); }
                        else if(r.cls == C_MEM) { x.push_back("m" + std::to_string(row_of.at(j.rd[i]))); }
                        else
                        {
                            x.push_back(valname(r.kw, r.ow));
                        }
                    }
                    for(std::size_t i{}; i < j.wr.size(); ++i)
                    {
                        w.push_back(valname(kk, static_cast<int>(i)));
                        st.push_back(needed[j.wr[i]] ? std::to_string(j.wr[i]) + "u" : std::string{});
                    }
                    // results are declared at function scope (the next tile may take them from their registers)
                    for(auto const& v: w) { decl << "    jv " << v << ";\n"; }
                    body << "        " << op_text(*j.o, x, w, st, false, j.steady_cap ? (j.steady_load ? 2 : 1) : 0) << "\n";
                }
                out << body.str() << "    }\n";
            }

            void emit_loop(int ri)
            {
                auto const& rg{p.regs[static_cast<std::size_t>(ri)]};
                int const T0{rg.tile0};
                auto const& m0{p.mem[static_cast<std::size_t>(T0)]};
                auto const& m1{p.mem[static_cast<std::size_t>(T0 + 1)]};
                std::string const pre{"r" + std::to_string(ri) + "_"};
                out << "    // tiles " << T0 << " .. " << (T0 + rg.n - 1) << ": " << rg.n << " x ops " << rg.op0 << " .. " << (rg.op0 + rg.per - 1) << "\n";
                // carried values: seeded from the peeled tile before the loop
                std::map<std::string, std::string> carried;  // loop variable -> copy made at the top of the body
                for(int q{}; q < rg.per; ++q)
                {
                    auto const& sr{p.src[static_cast<std::size_t>(rg.op0 + q)]};
                    for(auto const& r: sr)
                    {
                        if(r.cls != C_PREV) { continue; }
                        int const qw{r.kw - (rg.op0 - rg.per)};
                        std::string const lv{pre + "w" + std::to_string(qw) + "_" + std::to_string(r.ow)};
                        if(carried.count(lv) == 0)
                        {
                            carried[lv] = pre + "c" + std::to_string(carried.size());
                            out << "    " << lv << " = " << valname(r.kw, r.ow) << ";\n";
                        }
                    }
                }
                int const lim{limit_after(T0, T0 + rg.n - 1)};
                out << "#pragma unroll 1\n    for(uint32_t it = 0u; it < " << rg.n << "u; ++it)\n    {\n";
                out << "        uint32_t const t = " << T0 << "u + it;\n";
                emit_tile_head("t", "", "        ");
                for(std::size_t j{}; j < m0.size(); ++j) { out << "        jv const m" << j << " = sk_ring(sg, " << j << "u);\n"; }
                out << "        sk_ring_done(k, t);\n";
                out << "        { uint32_t tg = t + k.ns_mask + 2u; if(tg > " << lim << "u) { tg = " << lim << "u; } pe_stream_produce_upto" << sfx << "(k, tg, (int32_t)t - 1); }\n";
                for(auto const& [lv, cv]: carried) { out << "        jv const " << cv << " = " << lv << ";\n"; }
                std::map<std::uint32_t, std::size_t> row_of;
                for(std::size_t j{}; j < m0.size(); ++j) { row_of[m0[j]] = j; }
                (void)m1;
                for(int q{}; q < rg.per; ++q)
                {
                    int const kk{rg.op0 + q};
                    auto const& j{p.ops[static_cast<std::size_t>(kk)]};
                    auto const& j1{p.ops[static_cast<std::size_t>(kk + rg.per)]};
                    std::vector<std::string> x, w, st;
                    for(std::size_t i{}; i < j.rd.size(); ++i)
                    {
                        auto const& r{p.src[static_cast<std::size_t>(kk)][i]};
                        if(r.cls == C_INV) { x.push_back("inv" + std::to_string(p.inv_idx.at(j.rd[i]))); }
                        else if(r.cls == C_MEM) { x.push_back("m" + std::to_string(row_of.at(j.rd[i]))); }
                        else if(r.cls == C_SAME) { x.push_back(pre + "w" + std::to_string(r.kw - rg.op0) + "_" + std::to_string(r.ow)); }
                        else
                        {
                            x.push_back(carried.at(pre + "w" + std::to_string(r.kw - (rg.op0 - rg.per)) + "_" + std::to_string(r.ow)));
                        }
                    }
                    for(std::size_t i{}; i < j.wr.size(); ++i)
                    {
                        std::string const v{pre + "w" + std::to_string(q) + "_" + std::to_string(i)};
                        w.push_back(v);
                        decl << "    jv " << v << ";\n";
                        std::int64_t const d{static_cast<std::int64_t>(j1.wr[i]) - static_cast<std::int64_t>(j.wr[i])};
                        bool any{};
                        for(int it{}; it < rg.n && !any; ++it) { any = needed[static_cast<std::size_t>(static_cast<std::int64_t>(j.wr[i]) + it * d)] != 0; }
                        if(!any) { st.push_back({}); }
                        else
                        {
                            std::ostringstream a;
                            a << "(uint32_t)(" << j.wr[i] << " + (int32_t)it * " << d << ")";
                            st.push_back(a.str());
                        }
                    }
                    out << "        " << op_text(*j.o, x, w, st, false, j.steady_cap ? (j.steady_load ? 2 : 1) : 0) << "\n";
                }
                out << "    }\n";
            }

            std::ostringstream decl;

            void emit_produce()
            {
                out << "__device__ __forceinline__ void pe_stream_produce" << sfx << "(sk_ctx& k, uint32_t const p, int32_t const t_done)\n{\n";
                bool first{true};
                std::size_t ri{};
                while(ri < p.regs.size())
                {
                    auto const& rg{p.regs[ri]};
                    if(rg.loop())
                    {
                        int const T0{rg.tile0};
                        auto const& m0{p.mem[static_cast<std::size_t>(T0)]};
                        auto const& m1{p.mem[static_cast<std::size_t>(T0 + 1)]};
                        int wm{-1};
                        for(int it{}; it < rg.n; ++it) { wm = std::max(wm, p.wmax[static_cast<std::size_t>(T0 + it)]); }
                        out << "    " << (first ? "" : "else ") << "if(p < " << (T0 + rg.n) << "u)\n    {\n";
                        out << "        int32_t const it = (int32_t)(p - " << T0 << "u);\n";
                        if(wm >= 0) { out << "        if(" << wm << " > k.fenced) { sk_fence(k, t_done); }\n"; }
                        out << "        sk_tx(k, p, " << m0.size() << "u);\n";
                        // rows that stay adjacent in every iteration travel as one copy
                        std::size_t j{};
                        while(j < m0.size())
                        {
                            std::int64_t const d{static_cast<std::int64_t>(m1[j]) - static_cast<std::int64_t>(m0[j])};
                            std::size_t e{j + 1};
                            while(e < m0.size() && m0[e] == m0[e - 1] + 1u && static_cast<std::int64_t>(m1[e]) - static_cast<std::int64_t>(m0[e]) == d) { ++e; }
                            out << "        sk_copy(k, p, " << j << "u, (uint32_t)(" << m0[j] << " + it * " << d << "), " << (e - j) << "u);\n";
                            j = e;
                        }
                        out << "    }\n";
                        first = false;
                        ++ri;
                        continue;
                    }
                    // a run of single tiles: one switch
                    std::size_t re{ri};
                    while(re < p.regs.size() && !p.regs[re].loop()) { ++re; }
                    int const Ta{p.regs[ri].tile0}, Tb{p.regs[re - 1].tile0 + 1};
                    out << "    " << (first ? "" : "else ") << "if(p < " << Tb << "u)\n    {\n        switch(p)\n        {\n";
                    for(int T{Ta}; T < Tb; ++T)
                    {
                        out << "            case " << T << "u:";
                        if(p.wmax[static_cast<std::size_t>(T)] >= 0) { out << " if(" << p.wmax[static_cast<std::size_t>(T)] << " > k.fenced) { sk_fence(k, t_done); }"; }
                        out << " sk_tx(k, p, " << p.mem[static_cast<std::size_t>(T)].size() << "u);";
                        for(auto const& c: copies(T)) { out << " sk_copy(k, p, " << c[0] << "u, " << c[1] << "u, " << c[2] << "u);"; }
                        out << " break;\n";
                    }
                    out << "            default: break;\n        }\n    }\n";
                    first = false;
                    ri = re;
                }
                out << "}\n";
                out << "__device__ __forceinline__ void pe_stream_produce_upto" << sfx << "(sk_ctx& k, uint32_t target, int32_t const t_done)\n{\n"
                    << "    if(target > " << p.n_tiles << "u) { target = " << p.n_tiles << "u; }\n"
                    << "    while(k.pn < target)\n    {\n        pe_stream_produce" << sfx << "(k, k.pn, t_done);\n        ++k.pn;\n    }\n}\n";
            }

            std::string run()
            {
                std::ostringstream src;
                src << "// section" << sfx << ": " << p.n_tiles << " tiles\n";
                emit_produce();
                src << out.str();
                out.str(std::string{});
                for(std::size_t ri{}; ri < p.regs.size(); ++ri)
                {
                    if(p.regs[ri].loop()) { emit_loop(static_cast<int>(ri)); }
                    else
                    {
                        emit_single(static_cast<int>(ri));
                    }
                }
                src << "__device__ __forceinline__ void pe_stream" << sfx << "(sk_ctx& k, uint32_t& fm)\n{\n";
                // the section starts with the ring empty: every store of the earlier sections / solves is ordered before the
                // bulk copies that follow
                src << "    sk_fence(k, -1);\n    k.pn = 0u;\n";
                for(std::size_t i{}; i < p.inv.size(); ++i) { src << "    jv const inv" << i << " = sk_ld(k, " << p.inv[i] << "u);\n"; }
                // prologue: the tiles whose rows no tile of this section writes, up to the ring depth
                src << "    { uint32_t tg = k.ns_mask + 1u; if(tg > " << limit_after(0, -1) << "u) { tg = " << limit_after(0, -1) << "u; } pe_stream_produce_upto" << sfx << "(k, tg, -1); }\n";
                src << decl.str() << out.str();
                src << "    sk_end(k, " << p.n_tiles << "u);\n    k.seq0 += " << p.n_tiles << "u;\n}\n";
                return src.str();
            }
        };
    }  // namespace

    namespace
    {
        // every slot field of an op through f (sources / pairs / value operands keep their flag bits)
        template <class F>
        void remap_rop(rop& o, F&& f)
        {
            if(o.bubble) { return; }
            auto field = [&](std::uint32_t& w) { w = (w & ~k_slot_mask) | f(w & k_slot_mask); };
            if(o.opcode == PE_OP_DOT || o.opcode == PE_OP_CDOT)
            {
                field(o.dst);
                field(o.scale);
                for(auto& w: o.sre) { field(w); }
                for(auto& w: o.sim) { field(w); }
                for(auto& pp: o.pp)
                {
                    field(pp.first);
                    field(pp.second);
                }
            }
            else
            {
                for(auto& w: o.opnd) { field(w); }
            }
            for(auto& sb: o.sub) { remap_rop(sb, f); }
        }

    }  // namespace

    bool stream_prepare(program& pr)
    {
        if(pr.stream_laid_out) { return true; }
        if(!stream_supported(pr)) { return false; }
        static bool const no_layout{std::getenv("PE_B200_STREAM_NO_LAYOUT") != nullptr};
        static bool const no_replica{std::getenv("PE_B200_STREAM_NO_REPLICA") != nullptr};
        pr.stream_laid_out = true;
        if(no_layout) { return true; }
        // ---- 1. replica rows: a read-only family of rows that two periodic sweeps fetch (the conductances of a chain:
        // the pivot of the forward sweep and the back substitution both read them) gets a copy per extra sweep, so that
        // each sweep finds its rows next to each other.  The copies are made by COPY ops appended to the prep section
        // and by the load table, i.e. by every executor of the program.
        {
            plan p;
            if(!make_plan(pr, p)) { return false; }
            auto const stepw{step_written(pr)};
            std::vector<std::pair<std::uint32_t, std::uint32_t>> replicas;  // (source row, new row)
            // owner of a fetched row = the periodic run (or lone op) that reads it first
            std::map<std::uint32_t, int> owner;
            for(std::size_t k{}; k < p.ops.size(); ++k)
            {
                for(std::size_t i{}; i < p.ops[k].rd.size(); ++i)
                {
                    if(p.src[k][i].cls != C_MEM) { continue; }
                    owner.emplace(p.ops[k].rd[i], p.op_run[k] >= 0 ? p.op_run[k] : -2 - static_cast<int>(k));
                }
            }
            for(std::size_t rho{}; rho < p.runs.size() && !no_replica; ++rho)
            {
                auto const [op0, per, n_it]{p.runs[rho]};
                // operand positions of the period whose rows mostly belong to another run and never change
                std::set<std::pair<int, std::size_t>> repl;
                for(int q{}; q < per; ++q)
                {
                    for(std::size_t i{}; i < p.ops[static_cast<std::size_t>(op0 + q)].rd.size(); ++i)
                    {
                        int total{}, foreign{};
                        bool constant{true};
                        for(int m{}; m < n_it; ++m)
                        {
                            auto const kk{static_cast<std::size_t>(op0 + m * per + q)};
                            if(p.src[kk][i].cls != C_MEM) { continue; }
                            auto const sl{p.ops[kk].rd[i]};
                            ++total;
                            if(owner.at(sl) != static_cast<int>(rho)) { ++foreign; }
                            if(p.written[sl] || stepw[sl]) { constant = false; }
                        }
                        if(total > 0 && foreign * 2 > total && constant && pr.r_slots + total <= PE_R_MAX_SLOTS - 1) { repl.insert({q, i}); }
                    }
                }
                if(repl.empty()) { continue; }
                // new rows in the order the sweep reads them: the replica family is as regular as the original
                std::map<std::uint32_t, std::uint32_t> fresh;
                for(int m{}; m < n_it; ++m)
                {
                    for(int q{}; q < per; ++q)
                    {
                        auto const kk{static_cast<std::size_t>(op0 + m * per + q)};
                        bool hit{};
                        for(std::size_t i{}; i < p.ops[kk].rd.size(); ++i)
                        {
                            if(repl.count({q, i}) == 0 || p.src[kk][i].cls != C_MEM) { continue; }
                            auto const sl{p.ops[kk].rd[i]};
                            hit = true;
                            if(fresh.count(sl) == 0)
                            {
                                fresh[sl] = static_cast<std::uint32_t>(pr.r_slots++);
                                replicas.push_back({sl, fresh[sl]});
                            }
                        }
                        if(hit)
                        {
                            remap_rop(*const_cast<rop*>(p.ops[kk].o), [&](std::uint32_t f) { auto const z{fresh.find(f)}; return z != fresh.end() ? z->second : f; });
                        }
                    }
                }
            }
            if(!replicas.empty())
            {
                auto& prep{pr.rstreams[0].sec[0]};
                if(prep.empty()) { prep.emplace_back(); }
                for(auto const& [s, ns]: replicas)
                {
                    rop c;
                    c.opcode = PE_OP_COPY;
                    c.opnd = {ns, s};
                    prep.back().push_back(std::move(c));
                }
                pr.has_sec[0] = true;
                std::map<std::uint32_t, std::uint32_t> rep;
                for(auto const& [s, ns]: replicas) { rep[s] = ns; }
                std::size_t const n_io{pr.io.size()};
                for(std::size_t e{}; e < n_io; ++e)
                {
                    auto const io{pr.io[e]};
                    auto const it{rep.find(io.slot_kind & 0xffffu)};
                    if(it == rep.end() || !((io.slot_kind >> 20) & PE_IO_LOAD)) { continue; }
                    pr.io.push_back({it->second | (io.slot_kind & 0xf0000u) | (PE_IO_LOAD << 20), io.src});
                }
            }
        }
        // ---- 2. planes: the rows a tile fetches become consecutive rows, tile after tile in program order
        {
            plan p;
            if(!make_plan(pr, p)) { return false; }
            constexpr std::uint32_t none{0xffffffffu};
            std::vector<std::uint32_t> perm(static_cast<std::size_t>(pr.r_slots), none);
            std::uint32_t next{};
            auto take = [&](std::size_t s)
            {
                if(s < perm.size() && perm[s] == none) { perm[s] = next++; }
            };
            // rows fetched by bulk copies come first, in the order the program first reads them: the rows of a tile end up
            // next to each other (one copy per tile) and a regular family of rows stays regular.  The steady variant of the iter
            // section (all solves of a launch but the first) is laid out first.
            {
                plan pst;
                if(make_plan(pr, pst, 2, true))
                {
                    for(std::size_t k{}; k < pst.ops.size(); ++k)
                    {
                        for(std::size_t i{}; i < pst.ops[k].rd.size(); ++i)
                        {
                            if(pst.src[k][i].cls == C_MEM) { take(pst.ops[k].rd[i]); }
                        }
                    }
                }
            }
            for(std::size_t k{}; k < p.ops.size(); ++k)
            {
                for(std::size_t i{}; i < p.ops[k].rd.size(); ++i)
                {
                    if(p.src[k][i].cls == C_MEM) { take(p.ops[k].rd[i]); }
                }
            }
            // then the rows the prep section fetches (its tiles are generated the same way)
            {
                plan pq;
                if(pr.has_sec[0] && make_plan(pr, pq, 0))
                {
                    for(std::size_t k{}; k < pq.ops.size(); ++k)
                    {
                        for(std::size_t i{}; i < pq.ops[k].rd.size(); ++i)
                        {
                            if(pq.src[k][i].cls == C_MEM) { take(pq.ops[k].rd[i]); }
                        }
                    }
                }
            }
            for(std::size_t s{}; s < perm.size(); ++s) { take(s); }
            auto f = [&](std::uint32_t slot) -> std::uint32_t { return slot < perm.size() ? perm[slot] : slot; };
            for(auto& sec: pr.rstreams[0].sec)
            {
                for(auto& ph: sec)
                {
                    for(auto& o: ph) { remap_rop(o, f); }
                }
            }
            for(auto& io: pr.io) { io.slot_kind = (io.slot_kind & ~0xffffu) | f(io.slot_kind & 0xffffu); }
            for(auto& x: pr.x_slot) { x = f(x); }
            pr.r_zero = f(pr.r_zero);
        }
        pr.packed_ig = -1;  // the words are re-packed from the re-laid ops
        return true;
    }

    // opcode histogram of a section (tooling / debug)
    std::map<std::uint32_t, std::size_t> stream_section_ops(program const& pr, int sec)
    {
        std::map<std::uint32_t, std::size_t> h;
        if(pr.rstreams.empty()) { return h; }
        for(auto const& ph: pr.rstreams[0].sec[sec])
        {
            for(auto const& o: ph)
            {
                if(!o.bubble) { ++h[o.opcode]; }
            }
        }
        return h;
    }

    bool stream_supported(program const& pr)
    {
        if(!pr.resident || pr.cplx || pr.rS != 1 || pr.rstreams.size() != 1) { return false; }
        std::vector<sop> ops;
        return flatten(pr, 2, ops);
    }

    // The stream kernel pays off where the iter section is periodic: a long chain of like elimination steps becomes a few
    // rolled loops.  Irregular programs would come out as thousands of straight-line tiles (minutes of nvcc, no locality for
    // the bulk copies): they keep the tree-scheduled kernels.
    bool stream_profitable(program const& pr)
    {
        plan p;
        if(!stream_supported(pr) || !make_plan(pr, p)) { return false; }
        std::size_t in_loops{}, straight{};
        for(auto const& rg: p.regs)
        {
            if(rg.loop()) { in_loops += static_cast<std::size_t>(rg.n) * static_cast<std::size_t>(rg.per); }
            else
            {
                straight += static_cast<std::size_t>(rg.per);
            }
        }
        return in_loops * 10 >= p.ops.size() * 7 && straight <= 1200;
    }

    std::string stream_generate(program const& pr, stream_geom& g)
    {
        plan p;
        if(!stream_supported(pr) || !make_plan(pr, p)) { return {}; }
        bool const dbg{std::getenv("PE_B200_STREAM_DEBUG") != nullptr};
        if(dbg)
        {
            for(int sec{}; sec < 2; ++sec)
            {
                std::fprintf(stderr, "stream: section %d ops:", sec);
                for(auto const& [op, n]: stream_section_ops(pr, sec)) { std::fprintf(stderr, " %u x %zu", op, n); }
                std::fprintf(stderr, "\n");
            }
        }
        emitter e{p, "_iter", kept_slots(pr), {}, {}};
        // a store is needed when some load takes the row from memory (now or in the next solve)
        for(std::size_t k{}; k < p.ops.size(); ++k)
        {
            for(std::size_t i{}; i < p.ops[k].rd.size(); ++i)
            {
                if(p.src[k][i].cls == C_MEM || p.src[k][i].cls == C_INV) { e.needed[p.ops[k].rd[i]] = 1; }
            }
        }
        static bool const keep_all{std::getenv("PE_B200_STREAM_KEEP_STORES") != nullptr};
        if(keep_all) { std::fill(e.needed.begin(), e.needed.end(), 1); }
        std::string const iter_src{e.run()};
        // steady variant (see flatten): generated when it differs from the full one
        std::string steady_src;
        plan ps;
        static bool const no_steady{std::getenv("PE_B200_STREAM_NO_STEADY") != nullptr};
        bool gen_steady{!no_steady && make_plan(pr, ps, 2, true)};
        if(gen_steady)
        {
            gen_steady = false;
            for(auto const& o: ps.ops) { gen_steady = gen_steady || o.steady_cap; }
        }
        if(gen_steady)
        {
            emitter es{ps, "_iters", kept_slots(pr), {}, {}};
            for(std::size_t k{}; k < ps.ops.size(); ++k)
            {
                for(std::size_t i{}; i < ps.ops[k].rd.size(); ++i)
                {
                    if(ps.src[k][i].cls == C_MEM || ps.src[k][i].cls == C_INV) { es.needed[ps.ops[k].rd[i]] = 1; }
                }
            }
            // rows the FULL variant fetches stay stored too (the next launch starts with it)
            for(std::size_t sl{}; sl < es.needed.size(); ++sl) { es.needed[sl] = es.needed[sl] || e.needed[sl]; }
            if(keep_all) { std::fill(es.needed.begin(), es.needed.end(), 1); }
            steady_src = es.run();
        }
        // the prep section (derived per-instance values: conductances, replica rows) as tiles too when the generator covers
        // its ops; the kernel interprets it otherwise
        std::string prep_src;
        plan pp;
        static bool const no_prep{std::getenv("PE_B200_STREAM_NO_PREP") != nullptr};
        bool const gen_prep{!no_prep && pr.has_sec[0] && make_plan(pr, pp, 0)};
        if(gen_prep)
        {
            emitter ep{pp, "_prep", std::vector<std::uint8_t>(PE_R_MAX_SLOTS + 1, 1), {}, {}};  // every result of prep is kept
            prep_src = ep.run();
        }
        std::ostringstream src;
        g.n_tiles = static_cast<std::uint32_t>(p.n_tiles);
        g.stage_rows = static_cast<std::uint32_t>(std::max({p.stage_rows, gen_prep ? pp.stage_rows : 1, gen_steady ? ps.stage_rows : 1}));
        src << "// generated by host/stream.cpp: one-stream program as tiles, ring stage = " << g.stage_rows << " rows\n";
        src << "#define PE_STREAM_TILES " << p.n_tiles << "u\n#define PE_STREAM_STAGE_ROWS " << g.stage_rows << "u\n";
        if(gen_prep) { src << "#define PE_STREAM_PREP 1\n"; }
        if(gen_steady) { src << "#define PE_STREAM_STEADY 1\n"; }
        src << iter_src << steady_src << prep_src;
        g.n_loops = 0;
        g.loop_ops = 0;
        for(auto const& rg: p.regs)
        {
            if(rg.loop())
            {
                ++g.n_loops;
                g.loop_ops += static_cast<std::size_t>(rg.n) * static_cast<std::size_t>(rg.per);
            }
        }
        g.n_ops = p.ops.size();
        std::size_t rows{}, cps{}, st{};
        for(int t{}; t < p.n_tiles; ++t)
        {
            rows += p.mem[static_cast<std::size_t>(t)].size();
            cps += e.copies(t).size();
        }
        for(auto const& o: p.ops)
        {
            for(auto const s_: o.wr) { st += e.needed[s_] ? 1u : 0u; }
        }
        g.rows_fetched = rows;
        g.n_copies = cps;
        g.rows_stored = st;
        return src.str();
    }
}  // namespace pe_b200
