// jit.cpp — run-time specialisation of the tree-streaming kernel (DESIGN.md §5 "specialised kernel").
//
// The iter section of a tree-scheduled program (assembly + LU + substitution of one solve_once, circuit.h:987-1527) is
// turned into straight-line CUDA source: one function per (sub-tree stream, phase), deduplicated over isomorphic
// sub-trees (operand rows are named relative to the stream, exactly like the packed words), operand rows addressed by
// constants, results of the last few ops forwarded in registers, the loads of an op issued D ops before it runs.  The
// source is compiled for sm_100a together with csrc/pe_b200_kernels.cu (-DPE_JIT: scheduling, time loop, flags and I/O
// are the tree kernel's own code) and cached as a cubin next to the library.  The arithmetic of every op is the
// interpreter's, operation for operation, so results are bit-identical to the interpreted kernels.
#include "pe_host.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

namespace pe_b200
{
    namespace
    {
        std::uint64_t fnv(std::uint64_t h, void const* p, std::size_t n)
        {
            auto const* b{static_cast<unsigned char const*>(p)};
            for(std::size_t i{}; i < n; ++i)
            {
                h ^= b[i];
                h *= 1099511628211ull;
            }
            return h;
        }

        // cache files are only ever read when they are regular files owned by this user (no symlinks: O_NOFOLLOW)
        bool read_file(std::string const& path, std::string& out, bool owned = false)
        {
            int const fd{::open(path.c_str(), O_RDONLY | (owned ? O_NOFOLLOW : 0) | O_CLOEXEC)};
            if(fd < 0) { return false; }
            struct stat st{};
            if(::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (owned && st.st_uid != ::geteuid()))
            {
                ::close(fd);
                return false;
            }
            out.resize(static_cast<std::size_t>(st.st_size));
            std::size_t got{};
            while(got < out.size())
            {
                ssize_t const k{::read(fd, out.data() + got, out.size() - got)};
                if(k <= 0) { break; }
                got += static_cast<std::size_t>(k);
            }
            ::close(fd);
            return got == out.size();
        }

        // write under a per-process temporary name, then rename into place (concurrent ranks compile the same key)
        bool write_file_atomic(std::string const& path, std::string const& text)
        {
            std::string const tmp{path + ".tmp" + std::to_string(static_cast<long>(::getpid()))};
            int const fd{::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_NOFOLLOW | O_CLOEXEC, 0600)};
            if(fd < 0) { return false; }
            std::size_t put{};
            while(put < text.size())
            {
                ssize_t const k{::write(fd, text.data() + put, text.size() - put)};
                if(k <= 0) { break; }
                put += static_cast<std::size_t>(k);
            }
            ::close(fd);
            if(put != text.size() || ::rename(tmp.c_str(), path.c_str()) != 0)
            {
                ::unlink(tmp.c_str());
                return false;
            }
            return true;
        }

        // a directory this user owns and nobody else can write to
        bool private_dir(std::string const& d)
        {
            struct stat st{};
            return ::lstat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode) && st.st_uid == ::geteuid() && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0;
        }

        std::string lib_dir()
        {
            Dl_info info{};
            if(dladdr(reinterpret_cast<void const*>(&jit_supported), &info) != 0 && info.dli_fname != nullptr)
            {
                std::string p{info.dli_fname};
                auto const k{p.rfind('/')};
                return k == std::string::npos ? std::string{"."} : p.substr(0, k);
            }
            return ".";
        }

        int env_int(char const* name, int dflt)
        {
            char const* v{std::getenv(name)};
            return v != nullptr ? std::atoi(v) : dflt;
        }

        // Where compiled kernels are cached: <library directory>/jit_cache when this user owns it and nobody else can write
        // to it (the in-tree build), else a per-user directory ($XDG_CACHE_HOME/pe_b200 or ~/.cache/pe_b200, mode 0700).
        // A directory that is a symlink, belongs to somebody else or is group / world writable is refused: the files in
        // it are loaded as GPU code.  Empty string = no usable cache.
        std::string cache_dir(std::string& err)
        {
            std::string const a{lib_dir() + "/jit_cache"};
            (void)::mkdir(a.c_str(), 0700);
            if(private_dir(a) && ::access(a.c_str(), W_OK) == 0) { return a; }
            std::string base;
            if(char const* x{std::getenv("XDG_CACHE_HOME")}; x != nullptr && x[0] == '/') { base = x; }
            else if(char const* h{std::getenv("HOME")}; h != nullptr && h[0] == '/')
            {
                base = std::string{h} + "/.cache";
                (void)::mkdir(base.c_str(), 0700);
            }
            if(!base.empty())
            {
                std::string const b{base + "/pe_b200"};
                (void)::mkdir(b.c_str(), 0700);
                if(private_dir(b) && ::access(b.c_str(), W_OK) == 0) { return b; }
            }
            err = "jit: no private cache directory (tried " + a + " and $XDG_CACHE_HOME / ~/.cache): refusing a shared one";
            return {};
        }

        struct jop
        {
            rop const* o{};
            std::vector<std::uint32_t> reads;   // rel slots, in the op's operand order (sign bits stripped)
            std::vector<std::uint32_t> writes;  // rel slots
        };

        // ---- code generation of one (stream, phase) function ------------------------------------------------------
        // Relative columns are the function's parameters k0, k1, ... (numbered in order of first use): sub-trees that
        // differ only in where their separators live share one function.  The workspace of a 128-lane group is one
        // block, ws[group][slot][128 lanes]: a row is 1 KB and an operand address is a column base (one pointer per column
        // and 4 MB anchor) plus a constant that fits the immediate field of the load / store.
        //
        // Periodic runs of ops (the elimination of a chain of like nodes repeats one pattern with all rows shifted by a
        // constant) are rolled into a loop: the code of one period with moving row pointers, so that the instruction
        // footprint of a sub-tree is a few KB instead of a few hundred KB (the straight-line form stalls on instruction
        // fetch).  Forwarding and the load-ahead pipeline run across the back edge in named registers.
        struct gen
        {
            std::vector<jop> const& ops;
            int S, D;
            int K;  // forwarding window (>= D): a result is taken from its register by the ops up to K positions later
            std::uint32_t sm;
            std::vector<std::uint32_t>& cols;
            std::uint32_t stream{};                        // absolute slot = (rel & ~sm) | ((rel + stream) & sm)
            std::vector<std::uint8_t>* mem_reads{};        // pass 1: absolute slots some load takes from memory
            std::vector<std::uint8_t> const* needed{};     // pass 2: a store to a slot nobody loads is left out
            std::set<std::pair<std::size_t, std::uint32_t>> anchors;
            std::ostringstream body;
            int n_regions{};

            gen(std::vector<jop> const& o, int S_, int D_, int K_, std::vector<std::uint32_t>& c) : ops{o}, S{S_}, D{D_}, K{std::max(K_, D_)}, sm{static_cast<std::uint32_t>(S_ - 1)}, cols{c}
            {
                cols.clear();
            }
            std::uint32_t abs_slot(std::int64_t rel) const { return (static_cast<std::uint32_t>(rel) & ~sm) | ((static_cast<std::uint32_t>(rel) + stream) & sm); }
            void note_read(std::int64_t rel)
            {
                if(mem_reads != nullptr && rel >= 0 && abs_slot(rel) < mem_reads->size()) { (*mem_reads)[abs_slot(rel)] = 1; }
            }
            bool store_needed(std::int64_t rel) const { return needed == nullptr || abs_slot(rel) >= needed->size() || (*needed)[abs_slot(rel)] != 0; }

            std::size_t col_of(std::uint32_t rel)
            {
                std::size_t ci{};
                while(ci < cols.size() && cols[ci] != (rel & sm)) { ++ci; }
                if(ci == cols.size()) { cols.push_back(rel & sm); }
                return ci;
            }
            std::string addr(std::uint32_t rel)
            {
                std::size_t const ci{col_of(rel)};
                std::uint64_t const off{static_cast<std::uint64_t>(rel & ~sm) * 1024u};
                std::uint32_t const an{static_cast<std::uint32_t>(off >> 22)};
                anchors.insert({ci, an});
                std::ostringstream a;
                a << "c" << ci << "_" << an << " + " << (off & 0x3fffffu);
                return a.str();
            }

            // the statement(s) of one op: x = operand expressions, w = names of its result variables, st = store addresses
            static std::string op_text(rop const& o, std::vector<std::string> const& x, std::vector<std::string> const& w, std::vector<std::string> const& st, bool declare)
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
                    if((o.flags & PE_F_SCALE) && (o.flags & PE_F_GUARD)) { t << " jguard(" << w[0] << ", guard, fm);"; }  // pe_b200_program.h: an entry of L out of bounds
                    if(o.flags & PE_F_RECIP) { t << " jrcp(" << w[0] << ", fm);"; }
                    if(!st[0].empty()) { t << " jst(" << st[0] << ", " << w[0] << ", enm);"; }
                }
                else  // PE_OP_CAP_STEP: [hist][prev_g][C][dt][va][vb]
                {
                    t << dv << w[0] << " = " << x[0] << "; " << dv << w[1] << " = " << x[1] << "; jcap(" << x[2] << ", " << x[3] << ", " << x[4] << ", " << x[5] << ", " << w[0] << ", " << w[1]
                      << ");";
                    if(!st[0].empty()) { t << " jst(" << st[0] << ", " << w[0] << ", enm);"; }
                    if(!st[1].empty()) { t << " jst(" << st[1] << ", " << w[1] << ", enm);"; }
                }
                return t.str();
            }

            // a DOT with many terms (the update a sub-tree owes its separators sums one product per eliminated node): its
            // own loads are made a few terms at a time (PE_B200_JIT_BLOCK, default 8; 3 measured the same) right where they are consumed, in the op's own order of accumulation,
            // instead of all ahead of the op (hundreds of registers).  ld[i] = address to load operand i from ("" = x[i] is a
            // register that already holds it)
            static constexpr std::size_t long_reads{12};
            static std::string long_dot_text(rop const& o, std::vector<std::string> const& x, std::vector<std::string> const& ld, std::string const& w, std::string const& st)
            {
                std::ostringstream t;
                t << "jv " << w << " = jzero();\n";
                struct term
                {
                    int kind;  // 0 add, 1 sub, 2 fms, 3 fma, 4 mul
                    std::size_t a, b;
                };
                std::vector<term> terms;
                std::size_t q{};
                for(auto const s: o.sre) { terms.push_back({(s & PE_R_NEG) ? 1 : 0, q, q}), ++q; }
                for(auto const& pp: o.pp)
                {
                    terms.push_back({((pp.first ^ pp.second) & PE_R_NEG) ? 3 : 2, q, q + 1});
                    q += 2;
                }
                if(o.flags & PE_F_SCALE) { terms.push_back({4, q, q}), ++q; }
                static std::size_t const blk{static_cast<std::size_t>(std::clamp(env_int("PE_B200_JIT_BLOCK", 8), 1, 16))};
                for(std::size_t b{}; b < terms.size(); b += blk)
                {
                    std::size_t const e{std::min(terms.size(), b + blk)};
                    t << "    {\n";
                    std::set<std::size_t> loaded;
                    for(std::size_t k{b}; k < e; ++k)
                    {
                        for(std::size_t const i: {terms[k].a, terms[k].b})
                        {
                            if(!ld[i].empty() && loaded.insert(i).second) { t << "        jv const " << x[i] << " = jld(" << ld[i] << ");\n"; }
                        }
                    }
                    t << "       ";
                    for(std::size_t k{b}; k < e; ++k)
                    {
                        static char const* const fn[5]{"jadd", "jsub", "jfms", "jfma", "jmul"};
                        t << " " << fn[terms[k].kind] << "(" << w << ", " << x[terms[k].a];
                        if(terms[k].kind == 2 || terms[k].kind == 3) { t << ", " << x[terms[k].b]; }
                        t << ");";
                    }
                    t << "\n        asm volatile(\"\" ::: \"memory\");  // keeps the loads of the next block behind this one\n    }\n";
                }
                t << "   ";
                if((o.flags & PE_F_SCALE) && (o.flags & PE_F_GUARD)) { t << " jguard(" << w << ", guard, fm);"; }
                if(o.flags & PE_F_RECIP) { t << " jrcp(" << w << ", fm);"; }
                if(!st.empty()) { t << " jst(" << st << ", " << w << ", enm);"; }
                return t.str();
            }

            // where the value of read i of op k comes from, for a self-contained run of ops starting at `first`
            struct src_t
            {
                int kind{};  // 0 = its own load, 1 = result of op k - d (output o), 2 = the load made for read o of op k - d
                int d{}, o{};
                bool operator== (src_t const& b) const { return kind == b.kind && d == b.d && o == b.o; }
            };
            static bool is_long(jop const& j) { return j.o->opcode == PE_OP_DOT && j.reads.size() > long_reads; }
            std::vector<std::vector<src_t>> analyse(int first, int count) const
            {
                std::map<std::uint32_t, std::pair<int, int>> last_write, last_load;  // rel -> (op, output / read index)
                std::vector<std::vector<src_t>> res(static_cast<std::size_t>(count));
                for(int k{first}; k < first + count; ++k)
                {
                    auto const& j{ops[static_cast<std::size_t>(k)]};
                    for(std::size_t i{}; i < j.reads.size(); ++i)
                    {
                        std::uint32_t const rel{j.reads[i]};
                        auto const w{last_write.find(rel)};
                        auto const l{last_load.find(rel)};
                        src_t sr{};
                        if(w != last_write.end() && k - w->second.first <= K) { sr = {1, k - w->second.first, w->second.second}; }
                        else if(l != last_load.end() && k - l->second.first <= D && (w == last_write.end() || w->second.first < l->second.first - D))
                        {
                            sr = {2, k - l->second.first, l->second.second};
                        }
                        else if(!is_long(j))  // the loads of a long op live in its own blocks only
                        {
                            last_load[rel] = {k, static_cast<int>(i)};
                        }
                        res[static_cast<std::size_t>(k - first)].push_back(sr);
                    }
                    for(std::size_t i{}; i < j.writes.size(); ++i) { last_write[j.writes[i]] = {k, static_cast<int>(i)}; }
                }
                return res;
            }

            static bool same_shape(jop const& a, jop const& b)
            {
                rop const &x{*a.o}, &y{*b.o};
                if(x.opcode != y.opcode || x.flags != y.flags || a.reads.size() != b.reads.size() || a.writes.size() != b.writes.size() || x.sre.size() != y.sre.size() ||
                   x.pp.size() != y.pp.size())
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
                return true;
            }
            // op b repeats op a with every row shifted by m * (the shift between a and a + period)
            bool shifted(int a, int per, int m) const
            {
                jop const &x{ops[static_cast<std::size_t>(a)]}, &y{ops[static_cast<std::size_t>(a + per)]}, &z{ops[static_cast<std::size_t>(a + m * per)]};
                if(!same_shape(x, z)) { return false; }
                auto ok = [&](std::uint32_t s0, std::uint32_t s1, std::uint32_t sm_) -> bool
                {
                    std::int64_t const d{static_cast<std::int64_t>(s1) - static_cast<std::int64_t>(s0)};
                    return (s0 & sm) == (s1 & sm) && static_cast<std::int64_t>(sm_) == static_cast<std::int64_t>(s0) + m * d;
                };
                for(std::size_t i{}; i < x.reads.size(); ++i)
                {
                    if(!ok(x.reads[i], y.reads[i], z.reads[i])) { return false; }
                }
                for(std::size_t i{}; i < x.writes.size(); ++i)
                {
                    if(!ok(x.writes[i], y.writes[i], z.writes[i])) { return false; }
                }
                return true;
            }

            // longest periodic run starting at k0: period per (in ops), iterations n_it; returns covered ops (0 = none)
            int find_run(int k0, int& per, int& n_it) const
            {
                int const n{static_cast<int>(ops.size())};
                int best{};
                for(int p{1}; p <= 24 && k0 + 2 * p <= n; ++p)
                {
                    int m{1};
                    for(;; ++m)
                    {
                        if(k0 + (m + 1) * p > n) { break; }
                        bool all{true};
                        for(int q{}; q < p && all; ++q) { all = shifted(k0 + q, p, m); }
                        if(!all) { break; }
                    }
                    if(m >= 4 && m * p > best)
                    {
                        best = m * p;
                        per = p;
                        n_it = m;
                    }
                }
                return best;
            }

            void straight(int first, int count)
            {
                if(count <= 0) { return; }
                auto const srcs{analyse(first, count)};
                std::vector<std::vector<std::string>> loads_at(static_cast<std::size_t>(count));
                std::vector<std::vector<std::string>> opnd(static_cast<std::size_t>(count));
                for(int r{}; r < count; ++r)
                {
                    int const k{first + r};
                    auto const& j{ops[static_cast<std::size_t>(k)]};
                    for(std::size_t i{}; i < j.reads.size(); ++i)
                    {
                        auto const& sr{srcs[static_cast<std::size_t>(r)][i]};
                        std::ostringstream v;
                        if(sr.kind == 1) { v << "w" << (k - sr.d) << "_" << sr.o; }
                        else if(sr.kind == 2) { v << "l" << (k - sr.d) << "_" << sr.o; }
                        else
                        {
                            v << "l" << k << "_" << i;
                            note_read(j.reads[i]);
                            if(!is_long(j)) { loads_at[static_cast<std::size_t>(std::max(0, r - D))].push_back("jv const " + v.str() + " = jld(" + addr(j.reads[i]) + ");"); }
                        }
                        opnd[static_cast<std::size_t>(r)].push_back(v.str());
                    }
                }
                for(int r{}; r < count; ++r)
                {
                    int const k{first + r};
                    for(auto const& st: loads_at[static_cast<std::size_t>(r)]) { body << "    " << st << "\n"; }
                    auto const& j{ops[static_cast<std::size_t>(k)]};
                    std::vector<std::string> w, st;
                    for(std::size_t i{}; i < j.writes.size(); ++i)
                    {
                        w.push_back("w" + std::to_string(k) + "_" + std::to_string(i));
                        st.push_back(store_needed(j.writes[i]) ? addr(j.writes[i]) : std::string{});
                    }
                    if(is_long(j))
                    {
                        std::vector<std::string> ld;
                        for(std::size_t i{}; i < j.reads.size(); ++i) { ld.push_back(srcs[static_cast<std::size_t>(r)][i].kind == 0 ? addr(j.reads[i]) : std::string{}); }
                        body << "    " << long_dot_text(*j.o, opnd[static_cast<std::size_t>(r)], ld, w[0], st[0]) << "\n";
                        continue;
                    }
                    body << "    " << op_text(*j.o, opnd[static_cast<std::size_t>(r)], w, st, true) << "\n";
                }
            }

            // ops [k0, k0 + n_it * per) as a loop of n_it iterations; false = the run does not qualify (nothing emitted)
            bool rolled(int k0, int per, int n_it)
            {
                if(per < std::max(2 * D + 1, K + D + 1) || n_it < 3) { return false; }
                for(int q{}; q < per; ++q)
                {
                    if(is_long(ops[static_cast<std::size_t>(k0 + q)])) { return false; }
                }
                auto const srcs{analyse(k0, per * n_it)};
                // steady state = iteration 1; every later iteration must take its operands from the same places
                for(int m{2}; m < n_it; ++m)
                {
                    for(int q{}; q < per; ++q)
                    {
                        if(!(srcs[static_cast<std::size_t>(m * per + q)] == srcs[static_cast<std::size_t>(per + q)])) { return false; }
                    }
                }
                int const rid{n_regions};
                std::string const pre{"r" + std::to_string(rid) + "_"};
                // moving pointers: one per (column, row shift per iteration); the others are fixed addresses
                struct mov_t
                {
                    std::size_t ci;
                    std::int64_t delta;      // slots per iteration
                    std::uint32_t base;      // lowest slot (iteration 0) of the operands using it
                };
                std::vector<mov_t> movs;
                auto delta_of = [&](int q, bool wr, std::size_t i) -> std::int64_t
                {
                    auto const &a{ops[static_cast<std::size_t>(k0 + q)]}, &b{ops[static_cast<std::size_t>(k0 + per + q)]};
                    return wr ? static_cast<std::int64_t>(b.writes[i]) - static_cast<std::int64_t>(a.writes[i]) : static_cast<std::int64_t>(b.reads[i]) - static_cast<std::int64_t>(a.reads[i]);
                };
                auto mov_of = [&](std::uint32_t slot0, std::int64_t delta) -> std::size_t
                {
                    std::size_t const ci{col_of(slot0)};
                    for(std::size_t m{}; m < movs.size(); ++m)
                    {
                        if(movs[m].ci == ci && movs[m].delta == delta)
                        {
                            movs[m].base = std::min(movs[m].base, slot0 & ~sm);
                            return m;
                        }
                    }
                    movs.push_back({ci, delta, slot0 & ~sm});
                    return movs.size() - 1;
                };
                for(int q{}; q < per; ++q)
                {
                    auto const& a{ops[static_cast<std::size_t>(k0 + q)]};
                    for(std::size_t i{}; i < a.reads.size(); ++i)
                    {
                        if(delta_of(q, false, i) != 0) { mov_of(a.reads[i], delta_of(q, false, i)); }
                    }
                    for(std::size_t i{}; i < a.writes.size(); ++i)
                    {
                        if(delta_of(q, true, i) != 0) { mov_of(a.writes[i], delta_of(q, true, i)); }
                    }
                }
                bool fits{true};
                // address of an operand in iteration it + ahead (ahead = 0 / 1), as an expression valid inside the loop
                auto loop_addr = [&](std::uint32_t slot0, std::int64_t delta, int ahead) -> std::string
                {
                    if(delta == 0) { return addr(slot0); }
                    std::size_t const m{mov_of(slot0, delta)};
                    std::int64_t const imm{(static_cast<std::int64_t>(slot0 & ~sm) - static_cast<std::int64_t>(movs[m].base) + ahead * delta) * 1024};
                    if(imm >= (1 << 22) || imm <= -(1 << 22)) { fits = false; }
                    return pre + "m" + std::to_string(m) + " + " + std::to_string(imm);
                };
                auto wname = [&](int q, int o) { return pre + "w" + std::to_string(q) + "_" + std::to_string(o); };
                auto lname = [&](int q, int i) { return pre + "l" + std::to_string(q) + "_" + std::to_string(i); };
                std::ostringstream decl, init, prologue, loop;
                std::set<std::string> declared;
                auto declare = [&](std::string const& v)
                {
                    if(declared.insert(v).second) { decl << "    jv " << v << ";\n"; }
                };
                // operand names of the steady state; values carried over the back edge are seeded from memory before the loop
                std::vector<std::vector<std::string>> opnd(static_cast<std::size_t>(per));
                std::vector<std::vector<std::pair<std::string, std::string>>> loads_of(static_cast<std::size_t>(per));  // per op: (variable, read index) of its own loads
                for(int q{}; q < per; ++q)
                {
                    auto const& a{ops[static_cast<std::size_t>(k0 + q)]};
                    for(std::size_t i{}; i < a.reads.size(); ++i)
                    {
                        auto const& sr{srcs[static_cast<std::size_t>(per + q)][i]};
                        std::string v;
                        if(sr.kind == 0)
                        {
                            v = lname(q, static_cast<int>(i));
                            loads_of[static_cast<std::size_t>(q)].push_back({v, std::to_string(i)});
                            for(int m{}; m < n_it; ++m) { note_read(static_cast<std::int64_t>(a.reads[i]) + m * delta_of(q, false, i)); }
                        }
                        else
                        {
                            int const qp{((q - sr.d) % per + per) % per};
                            v = sr.kind == 1 ? wname(qp, sr.o) : lname(qp, sr.o);
                            if(q - sr.d < 0)
                            {
                                init << "    " << v << " = jld(" << loop_addr(a.reads[i], delta_of(q, false, i), 0) << ");\n";
                                note_read(a.reads[i]);
                            }
                        }
                        declare(v);
                        opnd[static_cast<std::size_t>(q)].push_back(v);
                    }
                    for(std::size_t i{}; i < a.writes.size(); ++i) { declare(wname(q, static_cast<int>(i))); }
                }
                auto emit_loads = [&](std::ostringstream& o, int q, int ahead, char const* indent)
                {
                    auto const& a{ops[static_cast<std::size_t>(k0 + q)]};
                    for(auto const& [v, is]: loads_of[static_cast<std::size_t>(q)])
                    {
                        std::size_t const i{static_cast<std::size_t>(std::stoi(is))};
                        o << indent << v << " = jld(" << loop_addr(a.reads[i], delta_of(q, false, i), ahead) << ");\n";
                    }
                };
                for(int q{}; q < std::min(D, per); ++q) { emit_loads(prologue, q, 0, "    "); }
                for(int q{}; q < per; ++q)
                {
                    int const q2{q + D};
                    if(q2 < per) { emit_loads(loop, q2, 0, "        "); }
                    else if(!loads_of[static_cast<std::size_t>(q2 - per)].empty())
                    {
                        loop << "        if(it + 1u < " << n_it << "u)\n        {\n";
                        emit_loads(loop, q2 - per, 1, "            ");
                        loop << "        }\n";
                    }
                    auto const& a{ops[static_cast<std::size_t>(k0 + q)]};
                    std::vector<std::string> w, st;
                    for(std::size_t i{}; i < a.writes.size(); ++i)
                    {
                        w.push_back(wname(q, static_cast<int>(i)));
                        bool any{};
                        for(int m{}; m < n_it && !any; ++m) { any = store_needed(static_cast<std::int64_t>(a.writes[i]) + m * delta_of(q, true, i)); }
                        st.push_back(any ? loop_addr(a.writes[i], delta_of(q, true, i), 0) : std::string{});
                    }
                    loop << "        " << op_text(*a.o, opnd[static_cast<std::size_t>(q)], w, st, false) << "\n";
                }
                if(!fits) { return false; }
                ++n_regions;
                body << "    // ops " << k0 << " .. " << (k0 + per * n_it - 1) << ": " << n_it << " x " << per << "\n" << decl.str();
                for(std::size_t m{}; m < movs.size(); ++m)
                {
                    body << "    char* " << pre << "m" << m << " = jopq(wl + (((k" << movs[m].ci << " + stream) & " << sm << "u) * 1024u + " << static_cast<std::uint64_t>(movs[m].base) * 1024u << "ull));\n";
                }
                body << init.str() << prologue.str() << "#pragma unroll 1\n    for(uint32_t it = 0u; it < " << n_it << "u; ++it)\n    {\n" << loop.str();
                for(std::size_t m{}; m < movs.size(); ++m) { body << "        " << pre << "m" << m << " += " << movs[m].delta * 1024 << "ll;\n"; }
                body << "    }\n";
                return true;
            }

            std::string run(bool roll)
            {
                int const n{static_cast<int>(ops.size())};
                int k{}, pending{};  // `pending` ops before k wait to be emitted as straight-line code
                while(k < n)
                {
                    int per{}, n_it{};
                    if(roll && find_run(k, per, n_it) > 0)
                    {
                        // a period long enough for the pipeline: unroll short periods; the operand sources (which loads are
                        // shared between neighbouring ops) may repeat with a multiple of the period only
                        int u{1};
                        while(per * u < std::max(2 * D + 1, K + D + 1)) { ++u; }
                        bool done{};
                        for(int tries{}; tries < 4 && !done; ++tries, ++u)
                        {
                            int const P{per * u}, N{n_it / u};
                            if(N < 3 || P > 32) { break; }
                            std::string const saved{body.str()};
                            auto const saved_regions{n_regions};
                            straight(k - pending, pending);
                            if(rolled(k, P, N))
                            {
                                pending = 0;
                                k += P * N;
                                done = true;
                                break;
                            }
                            body.str(saved);
                            body.seekp(0, std::ios::end);
                            n_regions = saved_regions;
                        }
                        if(done) { continue; }
                    }
                    ++pending;
                    ++k;
                }
                straight(n - pending, pending);
                std::ostringstream out;
                for(auto const& [ci, an]: anchors)
                {
                    out << "    char* const c" << ci << "_" << an << " = jopq(wl + (((k" << ci << " + stream) & " << sm << "u) * 1024u + " << (static_cast<std::uint64_t>(an) << 22) << "ull));\n";
                }
                out << "    uint32_t fm = 0u;\n" << body.str() << "    return fm;\n";
                return out.str();
            }
        };

        std::string gen_body(std::vector<jop> const& ops, int S, int D, std::vector<std::uint32_t>& cols, std::uint32_t stream, std::vector<std::uint8_t>* mem_reads,
                             std::vector<std::uint8_t> const* needed)
        {
            static bool const roll{std::getenv("PE_B200_JIT_NOROLL") == nullptr};
            static int const K{env_int("PE_B200_JIT_K", 1)};
            gen g{ops, S, D, K, cols};
            g.stream = stream;
            g.mem_reads = mem_reads;
            g.needed = needed;
            return g.run(roll);
        }
    }  // namespace

    bool jit_supported(program const& pr)
    {
        if(!pr.resident || pr.cplx || pr.rS < 1 || (pr.rS & (pr.rS - 1)) != 0 || pr.rstreams.size() != static_cast<std::size_t>(pr.rS)) { return false; }
        std::size_t n_ops{};
        for(auto const& st: pr.rstreams)
        {
            for(auto const& ph: st.sec[2])
            {
                for(auto const& o: ph)
                {
                    if(o.bubble) { continue; }
                    if(o.opcode != PE_OP_DOT && o.opcode != PE_OP_CAP_STEP) { return false; }
                    if(o.opcode == PE_OP_CAP_STEP && o.opnd.size() != 6) { return false; }
                    ++n_ops;
                }
            }
        }
        return n_ops > 0;
    }

    std::string jit_generate(program const& pr, int D)
    {
        int const S{pr.rS};
        std::size_t n_ph{};
        for(auto const& st: pr.rstreams) { n_ph = std::max(n_ph, st.sec[2].size()); }
        std::map<std::string, int> fid;                       // body -> function id
        std::vector<std::string> bodies;
        std::vector<std::size_t> n_cols;                      // per function: number of column parameters
        std::vector<std::vector<int>> call(n_ph, std::vector<int>(static_cast<std::size_t>(S), -1));  // [phase][stream] -> function
        std::vector<std::vector<std::vector<std::uint32_t>>> call_cols(n_ph, std::vector<std::vector<std::uint32_t>>(static_cast<std::size_t>(S)));
        // Pass 0 finds the slots some load takes from memory; a store to any other slot is dead (its value only ever
        // travels in registers, e.g. the L entries of a refactorisation) unless the slot is read outside the generated
        // code: by the interpreted sections, the load / store tables, the solution read-out or the waveform probes.
        std::uint32_t const smk{static_cast<std::uint32_t>(S - 1)};
        std::vector<std::uint8_t> needed(static_cast<std::size_t>(pr.r_slots) + 2u * static_cast<std::size_t>(S), 0);
        auto keep = [&](std::uint32_t abs)
        {
            if(abs < needed.size()) { needed[abs] = 1; }
        };
        for(int sj{}; sj < S; ++sj)
        {
            auto abs_of = [&](std::uint32_t field) { return ((field & 0x7fffu) & ~smk) | (((field & 0x7fffu) + static_cast<std::uint32_t>(sj)) & smk); };
            for(int sec{}; sec < 3; ++sec)
            {
                for(auto const& ph: pr.rstreams[static_cast<std::size_t>(sj)].sec[sec])
                {
                    for(auto const& o: ph)
                    {
                        if(o.bubble) { continue; }
                        if(sec == 2)
                        {
                            if(o.flags & (PE_F_CHECK_V | PE_F_CHECK_I)) { keep(abs_of(o.dst)); }
                            continue;
                        }
                        keep(abs_of(o.dst));
                        keep(abs_of(o.scale));
                        for(auto const w: o.sre) { keep(abs_of(w)); }
                        for(auto const w: o.sim) { keep(abs_of(w)); }
                        for(auto const& pp: o.pp)
                        {
                            keep(abs_of(pp.first));
                            keep(abs_of(pp.second));
                        }
                        for(auto const w: o.opnd) { keep(abs_of(w)); }
                        for(auto const& sb: o.sub) { keep(abs_of(sb.dst)); }
                    }
                }
            }
        }
        for(auto const& io: pr.io) { keep(io.slot_kind & 0xffffu); }
        for(auto const x: pr.x_slot) { keep(x); }
        static bool const keep_stores{std::getenv("PE_B200_JIT_KEEP_STORES") != nullptr};
        for(int pass{}; pass < 2; ++pass)
        for(std::size_t ph{}; ph < n_ph; ++ph)
        {
            for(int sj{}; sj < S; ++sj)
            {
                auto const& secs{pr.rstreams[static_cast<std::size_t>(sj)].sec[2]};
                if(ph >= secs.size()) { continue; }
                std::vector<jop> ops;
                for(auto const& o: secs[ph])
                {
                    if(o.bubble) { continue; }
                    jop j;
                    j.o = &o;
                    if(o.opcode == PE_OP_DOT)
                    {
                        for(auto const s: o.sre) { j.reads.push_back(s & 0x7fffu); }
                        for(auto const& pp: o.pp)
                        {
                            j.reads.push_back(pp.first & 0x7fffu);
                            j.reads.push_back(pp.second & 0x7fffu);
                        }
                        if(o.flags & PE_F_SCALE) { j.reads.push_back(o.scale & 0x7fffu); }
                        j.writes.push_back(o.dst & 0x7fffu);
                    }
                    else
                    {
                        for(auto const w: o.opnd) { j.reads.push_back(w & 0x7fffu); }
                        j.writes.push_back(o.opnd[0] & 0x7fffu);
                        j.writes.push_back(o.opnd[1] & 0x7fffu);
                    }
                    ops.push_back(std::move(j));
                }
                if(ops.empty()) { continue; }
                std::vector<std::uint32_t> cols;
                if(pass == 0)
                {
                    (void)gen_body(ops, S, D, cols, static_cast<std::uint32_t>(sj), &needed, nullptr);
                    continue;
                }
                std::string body{gen_body(ops, S, D, cols, static_cast<std::uint32_t>(sj), nullptr, keep_stores ? nullptr : &needed)};
                auto it{fid.find(body)};
                if(it == fid.end())
                {
                    it = fid.emplace(body, static_cast<int>(bodies.size())).first;
                    bodies.push_back(std::move(body));
                    n_cols.push_back(cols.size());
                }
                call[ph][static_cast<std::size_t>(sj)] = it->second;
                call_cols[ph][static_cast<std::size_t>(sj)] = std::move(cols);
            }
        }
        std::ostringstream src;
        src << "// generated by host/jit.cpp: iter section of a tree-scheduled program, S = " << S << ", load distance " << D << "\n";
        for(std::size_t f{}; f < bodies.size(); ++f)
        {
            src << "__device__ __forceinline__ uint32_t jf" << f << "(char* const wl, uint32_t const stream, uint32_t const enm, double const guard";
            for(std::size_t ci{}; ci < n_cols[f]; ++ci) { src << ", uint32_t const k" << ci; }
            src << ")\n{\n" << bodies[f] << "}\n";
        }
        src << "template <int CL>\n__device__ __forceinline__ void pe_jit_iter(uint32_t const warp, char* const wl, double const guard, bool const (&en)[4], bool (&fail)[4])\n{\n"
            << "    uint32_t const enm = (en[0] ? 1u : 0u) | (en[1] ? 2u : 0u) | (en[2] ? 4u : 0u) | (en[3] ? 8u : 0u);\n    uint32_t fm = 0u;\n";
        for(std::size_t ph{}; ph < n_ph; ++ph)
        {
            // every function is inlined exactly once per phase (a callee would have to move the memory descriptor into a
            // uniform register before every access): the first switch picks the function and its column arguments for this
            // warp's stream, the second one runs it
            std::size_t max_cols{};
            std::set<int> used;
            for(int sj{}; sj < S; ++sj)
            {
                if(call[ph][static_cast<std::size_t>(sj)] >= 0)
                {
                    used.insert(call[ph][static_cast<std::size_t>(sj)]);
                    max_cols = std::max(max_cols, call_cols[ph][static_cast<std::size_t>(sj)].size());
                }
            }
            src << "    {\n        uint32_t f = 0xffffffffu";
            for(std::size_t ci{}; ci < max_cols; ++ci) { src << ", a" << ci << " = 0u"; }
            src << ";\n        switch(warp)\n        {\n";
            for(int sj{}; sj < S; ++sj)
            {
                if(call[ph][static_cast<std::size_t>(sj)] < 0) { continue; }
                src << "            case " << sj << ": f = " << call[ph][static_cast<std::size_t>(sj)] << "u;";
                auto const& cc{call_cols[ph][static_cast<std::size_t>(sj)]};
                for(std::size_t ci{}; ci < cc.size(); ++ci) { src << " a" << ci << " = " << cc[ci] << "u;"; }
                src << " break;\n";
            }
            src << "            default: break;\n        }\n        switch(f)\n        {\n";
            for(int const f: used)
            {
                src << "            case " << f << ": fm |= jf" << f << "(wl, warp, enm, guard";
                for(std::size_t ci{}; ci < n_cols[static_cast<std::size_t>(f)]; ++ci) { src << ", a" << ci; }
                src << "); break;\n";
            }
            src << "            default: break;\n        }\n    }\n";
            if(ph + 1 < n_ph) { src << "    group_sync<CL>();\n"; }
        }
        src << "#pragma unroll\n    for(int j = 0; j < 4; ++j)\n    {\n        if((fm >> j) & 1u) { fail[j] = true; }\n    }\n}\n";
        return src.str();
    }

    // compile (or fetch from the cache next to the library) the specialised kernel; cubin = the file's bytes
    bool jit_compile(std::string const& gen, int CL, std::vector<char>& cubin, std::uint64_t& key, std::string& err, bool allow_compile)
    {
        std::string const dir{lib_dir()};
        std::string const csrc{dir + "/csrc"};
        std::uint64_t h{1469598103934665603ull};
        h = fnv(h, gen.data(), gen.size());
        h = fnv(h, &CL, sizeof(CL));
        for(char const* f: {"pe_b200_kernels.cu", "pe_b200_program.h", "pe_b200_models.h", "pe_b200_interp.h", "pe_b200_rinterp.h"})
        {
            std::string text;
            if(!read_file(csrc + "/" + f, text))
            {
                err = "jit: kernel source " + csrc + "/" + f + " not found";
                return false;
            }
            h = fnv(h, text.data(), text.size());
        }
        key = h;
        char name[64];
        std::snprintf(name, sizeof(name), "pe_jit_%016llx", static_cast<unsigned long long>(h));
        std::string const cache{cache_dir(err)};
        if(cache.empty()) { return false; }
        std::string const cub{cache + "/" + name + ".cubin"};
        std::string bytes;
        if(read_file(cub, bytes, true) && !bytes.empty())
        {
            cubin.assign(bytes.begin(), bytes.end());
            return true;
        }
        if(!allow_compile)
        {
            err = "jit: " + cub + " is not in the cache (tools/jit_prebuild.py or tuning bit 4 builds it)";
            return false;
        }
        std::string const inc{cache + "/" + name + ".inc"};
        if(!write_file_atomic(inc, gen))
        {
            err = "jit: cannot write " + inc;
            return false;
        }
        char const* nv{std::getenv("PE_B200_NVCC")};
        std::string nvcc{nv != nullptr ? nv : "/usr/local/cuda/bin/nvcc"};
        if(::access(nvcc.c_str(), X_OK) != 0) { nvcc = "nvcc"; }
        std::string const tmp{cub + ".tmp" + std::to_string(static_cast<long>(::getpid()))};
        std::string const log{tmp + ".log"};
        std::ostringstream cmd;
        cmd << nvcc << " -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -DPE_JIT -DPE_JIT_CL=" << CL << " '-DPE_JIT_SOURCE=\"" << inc << "\"' -I'" << csrc
            << "' -o '" << tmp << "' '" << csrc << "/pe_b200_kernels.cu' > '" << log << "' 2>&1";
        int const rc{std::system(cmd.str().c_str())};
        if(rc != 0 || !read_file(tmp, bytes, true) || bytes.empty())
        {
            std::string l;
            read_file(log, l, true);
            err = "jit: nvcc failed (" + std::to_string(WIFEXITED(rc) ? WEXITSTATUS(rc) : -1) + "): " + l.substr(0, 600);
            ::unlink(tmp.c_str());
            return false;
        }
        ::rename(tmp.c_str(), cub.c_str());
        ::rename(log.c_str(), (cache + "/" + name + ".log").c_str());
        cubin.assign(bytes.begin(), bytes.end());
        return true;
    }

    // the stream kernel's module (host/stream.cpp): generated source + csrc/pe_b200_stream.cu -> cubin (device seam) or a
    // host shared object (the emulator's seam), cached under a key of everything that goes into it
    bool stream_compile(std::string const& gen, int J, int GL, std::vector<char>& blob, std::uint64_t& key, std::string& err, bool allow_compile)
    {
        int const kind{pe_b200_stream_supported()};  // 1 = device (cubin bytes travel), 2 = emulator (the path travels)
        if(kind == 0)
        {
            err = "stream: not available in this build";
            return false;
        }
        std::string const dir{lib_dir()};
        std::string csrc{dir + "/csrc"};
        if(char const* e{std::getenv("PE_B200_CSRC")}; e != nullptr) { csrc = e; }
        else if(::access((csrc + "/pe_b200_stream.cu").c_str(), R_OK) != 0 && ::access((dir + "/../../phy-engine_b200/csrc/pe_b200_stream.cu").c_str(), R_OK) == 0)
        {
            csrc = dir + "/../../phy-engine_b200/csrc";  // the test emulator lives in tests/emu
        }
        std::uint64_t h{1469598103934665603ull};
        h = fnv(h, gen.data(), gen.size());
        h = fnv(h, &J, sizeof(J));
        h = fnv(h, &GL, sizeof(GL));
        int const checks{std::getenv("PE_B200_STREAM_CHECKS") != nullptr ? 1 : 0};  // debug build with bounds checks
        h = fnv(h, &checks, sizeof(checks));
        h = fnv(h, &kind, sizeof(kind));
        for(char const* f: {"pe_b200_stream.cu", "pe_b200_stream.h", "pe_b200_program.h", "pe_b200_models.h", "pe_b200_interp.h", "pe_b200_rinterp.h"})
        {
            std::string text;
            if(!read_file(csrc + "/" + f, text))
            {
                err = "stream: kernel source " + csrc + "/" + f + " not found";
                return false;
            }
            h = fnv(h, text.data(), text.size());
        }
        key = h;
        char name[64];
        std::snprintf(name, sizeof(name), "pe_stream_%016llx", static_cast<unsigned long long>(h));
        std::string const cache{cache_dir(err)};
        if(cache.empty()) { return false; }
        std::string const mod{cache + "/" + name + (kind == 2 ? ".so" : ".cubin")};
        std::string bytes;
        auto deliver = [&]() -> bool
        {
            if(!read_file(mod, bytes, true) || bytes.empty()) { return false; }
            if(kind == 2) { blob.assign(mod.begin(), mod.end()); }
            else
            {
                blob.assign(bytes.begin(), bytes.end());
            }
            return true;
        };
        if(deliver()) { return true; }
        if(!allow_compile)
        {
            err = "stream: " + mod + " is not in the cache";
            return false;
        }
        std::string const inc{cache + "/" + name + ".inc"};
        if(!write_file_atomic(inc, gen))
        {
            err = "stream: cannot write " + inc;
            return false;
        }
        std::string const tmp{mod + ".tmp" + std::to_string(static_cast<long>(::getpid()))};
        std::vector<char> log(2048, 0);
        if(pe_b200_stream_build(inc.c_str(), tmp.c_str(), csrc.c_str(), J, GL, log.data(), log.size()) != 0)
        {
            err = std::string{"stream: compiler failed: "} + log.data();
            ::unlink(tmp.c_str());
            return false;
        }
        ::unlink((tmp + ".log").c_str());
        if(::rename(tmp.c_str(), mod.c_str()) != 0 || !deliver())
        {
            err = "stream: cannot install " + mod;
            return false;
        }
        return true;
    }

    int jit_load_distance() { return std::clamp(env_int("PE_B200_JIT_D", 0), 0, 8); }
}  // namespace pe_b200
