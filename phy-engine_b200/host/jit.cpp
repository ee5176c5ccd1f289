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

        bool read_file(std::string const& path, std::string& out)
        {
            std::ifstream f(path, std::ios::binary);
            if(!f) { return false; }
            std::ostringstream ss;
            ss << f.rdbuf();
            out = ss.str();
            return true;
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

        struct jop
        {
            rop const* o{};
            std::vector<std::uint32_t> reads;   // rel slots, in the op's operand order (sign bits stripped)
            std::vector<std::uint32_t> writes;  // rel slots
        };

        // source of one (stream, phase) function body; `S` = streams per instance, D = load distance = forwarding window
        // The relative columns a function touches are its parameters k0, k1, ... (numbered in order of first use), so that
        // sub-trees which differ only in where their separators live share one function.
        std::string gen_body(std::vector<jop> const& ops, int S, int D, std::vector<std::uint32_t>& cols)
        {
            std::uint32_t const sm{static_cast<std::uint32_t>(S - 1)};
            std::ostringstream out;
            cols.clear();
            // The specialised kernel keeps the workspace of a 128-lane group in one block, ws[group][slot][128 lanes]: a
            // row is 1 KB, the address of an operand is a column base (one register pair per column and 4 MB anchor) plus
            // a constant that fits the immediate field of the load / store.
            std::set<std::pair<std::size_t, std::uint32_t>> anchors;
            auto addr = [&](std::uint32_t rel) -> std::string
            {
                std::size_t ci{};
                while(ci < cols.size() && cols[ci] != (rel & sm)) { ++ci; }
                if(ci == cols.size()) { cols.push_back(rel & sm); }
                std::uint64_t const off{static_cast<std::uint64_t>(rel & ~sm) * 1024u};
                std::uint32_t const an{static_cast<std::uint32_t>(off >> 22)};
                anchors.insert({ci, an});
                std::ostringstream a;
                a << "c" << ci << "_" << an << " + " << (off & 0x3fffffu);
                return a.str();
            };
            int const n{static_cast<int>(ops.size())};
            // value of rel slot as seen by op k: a forwarded register (written / loaded within the last D ops) or a load
            // issued at op max(0, k - D)
            std::map<std::uint32_t, std::pair<int, std::string>> last_write;  // rel -> (op, variable)
            std::map<std::uint32_t, std::pair<int, std::string>> last_load;   // rel -> (op the load was made for, variable)
            std::vector<std::vector<std::string>> opnd(static_cast<std::size_t>(n));
            std::vector<std::vector<std::string>> loads_at(static_cast<std::size_t>(n));  // load statements emitted before op p runs
            for(int k{}; k < n; ++k)
            {
                auto const& j{ops[static_cast<std::size_t>(k)]};
                for(std::size_t i{}; i < j.reads.size(); ++i)
                {
                    std::uint32_t const rel{j.reads[i]};
                    auto const w{last_write.find(rel)};
                    if(w != last_write.end() && k - w->second.first <= D)
                    {
                        opnd[static_cast<std::size_t>(k)].push_back(w->second.second);
                        continue;
                    }
                    auto const l{last_load.find(rel)};
                    if(l != last_load.end() && k - l->second.first <= D && (w == last_write.end() || w->second.first < l->second.first - D))
                    {
                        opnd[static_cast<std::size_t>(k)].push_back(l->second.second);
                        continue;
                    }
                    std::ostringstream v;
                    v << "l" << k << "_" << i;
                    std::ostringstream st;
                    st << "jv const " << v.str() << " = jld(" << addr(rel) << ");";
                    loads_at[static_cast<std::size_t>(std::max(0, k - D))].push_back(st.str());
                    last_load[rel] = {k, v.str()};
                    opnd[static_cast<std::size_t>(k)].push_back(v.str());
                }
                for(std::size_t i{}; i < j.writes.size(); ++i)
                {
                    std::ostringstream v;
                    v << "w" << k << "_" << i;
                    last_write[j.writes[i]] = {k, v.str()};
                }
            }
            std::ostringstream body;
            for(int k{}; k < n; ++k)
            {
                for(auto const& s: loads_at[static_cast<std::size_t>(k)]) { body << "    " << s << "\n"; }
                auto const& j{ops[static_cast<std::size_t>(k)]};
                auto const& x{opnd[static_cast<std::size_t>(k)]};
                rop const& o{*j.o};
                if(o.opcode == PE_OP_DOT)
                {
                    std::size_t q{};
                    body << "    jv w" << k << "_0 = jzero();";
                    for(auto const s: o.sre) { body << ((s & PE_R_NEG) ? " jsub(" : " jadd(") << "w" << k << "_0, " << x[q++] << ");"; }
                    for(auto const& pp: o.pp)
                    {
                        bool const pos{((pp.first ^ pp.second) & PE_R_NEG) != 0u};  // -(+-a)(+-b)
                        body << (pos ? " jfma(" : " jfms(") << "w" << k << "_0, " << x[q] << ", " << x[q + 1] << ");";
                        q += 2;
                    }
                    if(o.flags & PE_F_SCALE) { body << " jmul(w" << k << "_0, " << x[q++] << ");"; }
                    if(o.flags & PE_F_RECIP) { body << " jrcp(w" << k << "_0, fm);"; }
                    body << " jst(" << addr(j.writes[0]) << ", w" << k << "_0, enm);\n";
                }
                else  // PE_OP_CAP_STEP: [hist][prev_g][C][dt][va][vb]
                {
                    body << "    jv w" << k << "_0 = " << x[0] << ", w" << k << "_1 = " << x[1] << "; jcap(" << x[2] << ", " << x[3] << ", " << x[4] << ", " << x[5] << ", w" << k
                         << "_0, w" << k << "_1); jst(" << addr(j.writes[0]) << ", w" << k << "_0, enm); jst(" << addr(j.writes[1]) << ", w" << k << "_1, enm);\n";
                }
            }
            for(auto const& [ci, an]: anchors)
            {
                out << "    char* const c" << ci << "_" << an << " = wl + (((k" << ci << " + stream) & " << sm << "u) * 1024u + " << (static_cast<std::uint64_t>(an) << 22) << "ull);\n";
            }
            out << "    uint32_t fm = 0u;\n" << body.str() << "    return fm;\n";
            return out.str();
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
                std::string body{gen_body(ops, S, D, cols)};
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
            src << "__device__ __noinline__ uint32_t jf" << f << "(char* const wl, uint32_t const stream, uint32_t const enm";
            for(std::size_t ci{}; ci < n_cols[f]; ++ci) { src << ", uint32_t const k" << ci; }
            src << ")\n{\n" << bodies[f] << "}\n";
        }
        src << "template <int CL>\n__device__ __forceinline__ void pe_jit_iter(uint32_t const warp, char* const wl, bool const (&en)[4], bool (&fail)[4])\n{\n"
            << "    uint32_t const enm = (en[0] ? 1u : 0u) | (en[1] ? 2u : 0u) | (en[2] ? 4u : 0u) | (en[3] ? 8u : 0u);\n    uint32_t fm = 0u;\n";
        for(std::size_t ph{}; ph < n_ph; ++ph)
        {
            // streams that call the same function with the same columns share a case
            std::map<std::pair<int, std::vector<std::uint32_t>>, std::vector<int>> by_f;
            for(int sj{}; sj < S; ++sj)
            {
                if(call[ph][static_cast<std::size_t>(sj)] >= 0) { by_f[{call[ph][static_cast<std::size_t>(sj)], call_cols[ph][static_cast<std::size_t>(sj)]}].push_back(sj); }
            }
            src << "    switch(warp)\n    {\n";
            for(auto const& [fc, ss]: by_f)
            {
                src << "       ";
                for(int const sj: ss) { src << " case " << sj << ":"; }
                src << " fm |= jf" << fc.first << "(wl, warp, enm";
                for(auto const c: fc.second) { src << ", " << c << "u"; }
                src << "); break;\n";
            }
            src << "        default: break;\n    }\n";
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
        std::string cache{dir + "/jit_cache"};
        if(::mkdir(cache.c_str(), 0755) != 0 && ::access(cache.c_str(), W_OK) != 0)
        {
            cache = "/tmp/pe_b200_jit_cache";
            ::mkdir(cache.c_str(), 0755);
        }
        std::string const cub{cache + "/" + name + ".cubin"};
        std::string bytes;
        if(read_file(cub, bytes) && !bytes.empty())
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
        {
            std::ofstream f(inc, std::ios::binary);
            f << gen;
            if(!f)
            {
                err = "jit: cannot write " + inc;
                return false;
            }
        }
        char const* nv{std::getenv("PE_B200_NVCC")};
        std::string nvcc{nv != nullptr ? nv : "/usr/local/cuda/bin/nvcc"};
        if(::access(nvcc.c_str(), X_OK) != 0) { nvcc = "nvcc"; }
        std::string const tmp{cub + ".tmp" + std::to_string(static_cast<long>(::getpid()))};
        std::string const log{cache + "/" + name + ".log"};
        std::ostringstream cmd;
        cmd << nvcc << " -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -DPE_JIT -DPE_JIT_CL=" << CL << " '-DPE_JIT_SOURCE=\"" << inc << "\"' -I'" << csrc
            << "' -o '" << tmp << "' '" << csrc << "/pe_b200_kernels.cu' > '" << log << "' 2>&1";
        int const rc{std::system(cmd.str().c_str())};
        if(rc != 0 || !read_file(tmp, bytes) || bytes.empty())
        {
            std::string l;
            read_file(log, l);
            err = "jit: nvcc failed (" + std::to_string(WIFEXITED(rc) ? WEXITSTATUS(rc) : -1) + "): " + l.substr(0, 600);
            ::unlink(tmp.c_str());
            return false;
        }
        ::rename(tmp.c_str(), cub.c_str());
        cubin.assign(bytes.begin(), bytes.end());
        return true;
    }

    int jit_load_distance() { return std::clamp(env_int("PE_B200_JIT_D", 2), 0, 8); }
}  // namespace pe_b200
