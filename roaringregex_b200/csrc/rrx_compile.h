// Host-side regex compiler for the rrx scanner: ERE subset -> reference-equivalent automaton ->
// flattened position automaton -> device program blob.
//
// What it replaces in the reference (all host code there too):
//   Parser.cpp:16-170   RRegex::RRegex / build_NFA / bracket_expression  (the language, quirks included)
//   NFA.cc:42-71,108-189  the construction algebra (skip, *=, |=, *, +)  (restated on tracked atoms)
//   regex.h:16-75       MemoryPool (the table) -> here: the flattened device program
// See SURVEY.md Appendix A for the semantics and A.4 for why the flattening is exact.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "rrx_prog.h"

namespace rrx {

// Growable bitset over 32-bit words (bit i of the set = bit (i&31) of word i>>5).
struct Bits {
    std::vector<uint32_t> w;
    Bits() {}
    explicit Bits(uint32_t nbits) : w((nbits + 31) / 32, 0u) {}
    void ensure(uint32_t bit) { if ((bit >> 5) >= w.size()) w.resize((bit >> 5) + 1, 0u); }
    void add(uint32_t bit) { ensure(bit); w[bit >> 5] |= 1u << (bit & 31); }
    void del(uint32_t bit) { if ((bit >> 5) < w.size()) w[bit >> 5] &= ~(1u << (bit & 31)); }
    bool has(uint32_t bit) const { return (bit >> 5) < w.size() && ((w[bit >> 5] >> (bit & 31)) & 1u); }
    bool any() const { for (uint32_t x : w) if (x) return true; return false; }
    uint32_t count() const { uint32_t c = 0; for (uint32_t x : w) c += (uint32_t)__builtin_popcount(x); return c; }
    bool intersects(const Bits& o) const {
        size_t n = w.size() < o.w.size() ? w.size() : o.w.size();
        for (size_t i = 0; i < n; i++) if (w[i] & o.w[i]) return true;
        return false;
    }
    void unite(const Bits& o) {
        if (o.w.size() > w.size()) w.resize(o.w.size(), 0u);
        for (size_t i = 0; i < o.w.size(); i++) w[i] |= o.w[i];
    }
    bool equals(const Bits& o) const {
        size_t n = w.size() > o.w.size() ? w.size() : o.w.size();
        for (size_t i = 0; i < n; i++) {
            uint32_t a = i < w.size() ? w[i] : 0u, b = i < o.w.size() ? o.w[i] : 0u;
            if (a != b) return false;
        }
        return true;
    }
    std::vector<uint32_t> list() const {
        std::vector<uint32_t> out;
        for (size_t i = 0; i < w.size(); i++) {
            uint32_t x = w[i];
            while (x) { out.push_back((uint32_t)(i * 32 + __builtin_ctz(x))); x &= x - 1; }
        }
        return out;
    }
    Bits shifted_up(uint32_t r) const { Bits o; for (uint32_t b : list()) o.add(b + r); return o; }
    uint32_t word(uint32_t i) const { return i < w.size() ? w[i] : 0u; }
};

// One atom of the pattern (a literal, '.', or a bracket class; or a {m,n}/+ copy of one).  In the
// reference every transition ever created belongs to exactly one atom and the edge set is always
//   U_k pre(k) x class(k) x post(k)        
// over RAW reference states, so tracking (pre, class, post) per atom reproduces the raw tables.
struct Atom {
    uint64_t cls[2];   // 128-bit byte class
    uint32_t raw_init; // raw state the atom was created at (its dead-end final is raw_init+1)
    Bits pre, post;    // raw-state sets
};

// The automaton after parsing, in both views.
struct Automaton {
    uint32_t n_raw = 0;         // reference raw state count (Parser.cpp:163 states_n)
    Bits raw_finals;            // reference final_states
    std::vector<Atom> atoms;    // positions 0..P-1, in pattern order
    // flattened (position) view, START excluded; N_flat = P + 1
    uint32_t P = 0;
    bool nullable = false;      // raw initial state is final (empty string in the language)
    Bits first, accept;         // P bits
    std::vector<Bits> follow;   // follow[k], P bits; all edges, including those into `first`
};

// Parse `pattern` exactly as the reference would; throws std::runtime_error with the reference's
// messages ("invalid expression!" Parser.cpp:36, "invalid expression" Parser.cpp:155), and for the
// inputs on which the reference has undefined behaviour (operator without operand, non-ASCII).
Automaton compile_ere(const char* pattern);

// Reference set type for this raw state count (Parser.cpp:165-168): 0 BitSet<1>, 1 BitSet<2>,
// 2 BitSet<4>, 3 Roaring.
int reference_width_class(uint32_t n_raw);

// ---- device program ------------------------------------------------------------------------------
// Layout: rrx_prog.h (ProgHeader + tables).
struct Program {
    std::vector<uint32_t> blob;   // ProgHeader followed by the tables
    const ProgHeader& hdr() const { return *reinterpret_cast<const ProgHeader*>(blob.data()); }
    std::string pattern;
};

// `reduce`: merge positions that are indistinguishable (same predecessors+class-union rule, or
// same class/successors/accept) before laying out the device program.  Language-preserving.
// `allow_donly`: let the compiler choose the thread-owned candidate scan (rrx_prog.h `donly`) where it applies.
Program build_program(const Automaton& a, const std::string& pattern, bool reduce, bool allow_donly = true);

}  // namespace rrx
