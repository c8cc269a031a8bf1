// Host regex compiler: see rrx_compile.h.  From-scratch code; reference file:line citations say
// which behaviour each piece reproduces.
#include "rrx_compile.h"

#include <algorithm>
#include <cstring>
#include <map>

namespace rrx {
namespace {

enum Op { OP_CONCAT, OP_BRACKET, OP_OR };   // regex.h:15

struct Sub {                  // PseudoNFA{initial_state,size} + final_states (regex.h:77-96,177-178)
    uint32_t init = 0, size = 0;
    Bits finals;
};

[[noreturn]] void fail(const char* msg) { throw std::runtime_error(msg); }

// The construction algebra of NFA.cc, restated on atoms (pre, class, post) instead of on the
// dense forward/backward tables.  Invariant: raw edges = U_k pre(k) x cls(k) x post(k).
struct Builder {
    std::vector<Atom> atoms;

    // skip<true>(n,m), NFA.cc:108-121: n inherits every outgoing edge of m.
    void inherit_out(uint32_t n, uint32_t m) {
        for (Atom& a : atoms) if (a.pre.has(m)) a.pre.add(n);
    }
    // skip<false>(n,m), NFA.cc:108-121: m inherits every incoming edge of n.
    void inherit_in(uint32_t n, uint32_t m) {
        for (Atom& a : atoms) if (a.post.has(n)) a.post.add(m);
    }
    Sub empty(uint32_t cur) {                                   // NFA.cc:42-49
        Sub s; s.init = cur; s.size = 1; s.finals.add(cur); return s;
    }
    Sub atom(uint32_t cur, uint64_t lo, uint64_t hi) {          // NFA.cc:50-71
        Atom a; a.cls[0] = lo; a.cls[1] = hi; a.raw_init = cur;
        a.pre.add(cur); a.post.add(cur + 1);
        atoms.push_back(a);
        Sub s; s.init = cur; s.size = 2; s.finals.add(cur + 1); return s;
    }
    void concat(Sub& a, const Sub& b) {                         // NFA.cc:122-137
        a.size += b.size;
        for (uint32_t f : a.finals.list()) inherit_in(f, b.init);
        bool a_null = a.finals.has(a.init);
        if (a_null) inherit_out(a.init, b.init);
        bool both = a_null && b.finals.has(b.init);
        a.finals = b.finals;
        if (both) a.finals.add(a.init);
    }
    void unite(Sub& a, const Sub& b) {                          // NFA.cc:138-149
        a.size += b.size;
        a.finals.unite(b.finals);
        inherit_out(a.init, b.init);
        if (b.finals.has(b.init)) a.finals.add(a.init);
    }
    void star(Sub& a) {                                         // NFA.cc:150-157
        for (uint32_t f : a.finals.list()) inherit_in(f, a.init);
        a.finals.add(a.init);
    }
    Sub clone(const Sub& a, uint32_t r) {                       // NFA.cc:177-185 (+BitSet.cc:168-180)
        size_t n = atoms.size();
        for (size_t i = 0; i < n; i++) {
            if (atoms[i].raw_init < a.init || atoms[i].raw_init >= a.init + a.size) continue;
            Atom c;
            c.cls[0] = atoms[i].cls[0]; c.cls[1] = atoms[i].cls[1];
            c.raw_init = atoms[i].raw_init + r;
            c.pre = atoms[i].pre.shifted_up(r);
            c.post = atoms[i].post.shifted_up(r);
            atoms.push_back(c);
        }
        Sub s; s.init = a.init + r; s.size = a.size; s.finals = a.finals.shifted_up(r);
        return s;
    }
};

struct Parser {
    const unsigned char* p;
    int ps;
    Builder b;
    std::vector<Sub> nfas;
    std::vector<Op> ops;

    unsigned char at(int i) const { return i >= 0 && i < ps ? p[i] : 0; }
    uint32_t next_initial() const { return nfas.empty() ? 0u : nfas.back().init + nfas.back().size; }  // Parser.cpp:84-86
    void need_top() const { if (nfas.empty()) fail("invalid expression: operator without operand"); }

    // C strtol(…,10) on the NUL-terminated pattern: value and end index.
    long str_to_long(int pos, int* end) const {
        int i = pos;
        while (at(i) == ' ' || (at(i) >= '\t' && at(i) <= '\r')) i++;
        bool neg = false;
        if (at(i) == '+' || at(i) == '-') { neg = at(i) == '-'; i++; }
        int start = i;
        long v = 0;
        while (at(i) >= '0' && at(i) <= '9') { v = v * 10 + (at(i) - '0'); if (v > 100000) fail("invalid expression: repeat count too large"); i++; }
        if (i == start) { *end = pos; return 0; }
        *end = i;
        return neg ? -v : v;
    }

    // Bracket expression, Parser.cpp:16-39.  `pos` at '['; returns index of the closing byte.
    int bracket(int pos, uint64_t cls[2]) const {
        const int start = pos, rem = ps - start;
        cls[0] = cls[1] = 0;
        auto add = [&](unsigned c) { if (c < 128) cls[c >> 6] |= 1ull << (c & 63); };
        pos++;
        bool complement = at(pos) == '^';     // not skipped: the caret joins the set (A.3 quirk 3)
        bool escaped = false;
        while (pos + 1 < ps && !(at(pos) == ']' && !escaped)) {
            if (!escaped) {
                unsigned char nxt = at(pos + 1);
                if (nxt != ']' && pos + 2 < ps) {
                    unsigned char nn = at(pos + 2);
                    if (nxt == '-' && nn != ']') {
                        for (unsigned c = at(pos); c <= nn; c++) add(c);
                        pos += 3;
                        continue;
                    }
                }
            }
            escaped = !escaped && at(pos) == '\\';
            add(at(pos));                     // a backslash is itself a member (A.3 quirk 4)
            pos++;
        }
        if (pos - start == rem) fail("invalid expression!");                 // Parser.cpp:36
        if (complement) { cls[0] = ~cls[0]; cls[1] = ~cls[1]; }
        return pos;
    }

    void push_atom(uint64_t lo, uint64_t hi) {
        nfas.push_back(b.atom(next_initial(), lo, hi));
        ops.push_back(OP_CONCAT);
    }
    void repeat() {                                                          // Parser.cpp:80-83
        need_top();
        Sub c = b.clone(nfas.back(), nfas.back().size);
        nfas.push_back(c);
        ops.push_back(OP_CONCAT);
    }
    void optional_top() {                                                    // Parser.cpp:120-122
        need_top();
        Sub e = b.empty(next_initial());
        b.unite(nfas.back(), e);
    }
    // Fold the stack right-to-left down to the innermost open bracket, Parser.cpp:49-79.
    void clear_stack() {
        need_top();
        Sub cur = nfas.back();
        if (!ops.empty()) {
            ops.pop_back();
            while (ops.size() > 1 && ops.back() != OP_BRACKET) {
                if (ops.back() == OP_CONCAT) {
                    nfas.pop_back(); ops.pop_back();
                    need_top();
                    b.concat(nfas.back(), cur);
                    cur = nfas.back();
                } else {   // OP_OR
                    Sub right = cur;
                    ops.pop_back();
                    if (ops.empty()) fail("invalid expression");
                    ops.pop_back();
                    nfas.pop_back();
                    need_top();
                    cur = nfas.back();
                    while (ops.size() > 1 && ops.back() == OP_CONCAT) {
                        nfas.pop_back(); ops.pop_back();
                        need_top();
                        b.concat(nfas.back(), cur);
                        cur = nfas.back();
                    }
                    b.unite(cur, right);
                }
            }
        }
        if (ops.empty()) fail("invalid expression");
        ops.pop_back();
        ops.push_back(OP_CONCAT);
        nfas.pop_back();
        nfas.push_back(cur);
    }

    Sub run() {                                                              // Parser.cpp:41-160
        ops.push_back(OP_BRACKET);
        bool escaped = false;
        int cp = 0;
        for (;;) {                                   // do { … } while (++cp - p < ps)
            unsigned char ch = at(cp);
            if (!escaped && ch == '\\') {
                escaped = true;
                if (++cp < ps) continue;
                break;
            }
            unsigned char sel = escaped ? 0 : ch;
            switch (sel) {
                case '[': {
                    uint64_t cls[2];
                    cp = bracket(cp, cls);
                    push_atom(cls[0], cls[1]);
                    break;
                }
                case '(': ops.push_back(OP_BRACKET); break;
                case '|': ops.push_back(OP_OR); break;
                case ')': clear_stack(); break;
                case '.': push_atom(~0ull, ~0ull); break;       // all 128 bytes incl. 0x00 and '\n'
                case '*': need_top(); b.star(nfas.back()); break;
                case '+': repeat(); b.star(nfas.back()); break;  // A+ = A A'*
                case '?': optional_top(); break;
                case '{': {                                     // Parser.cpp:123-141
                    int e1 = 0, e2 = 0;
                    long m = str_to_long(++cp, &e1);
                    for (long i = 0; i < m - 1; i++) repeat();
                    if (at(e1) != '}') {
                        long n = str_to_long(e1 + 1, &e2);
                        if (!n) { repeat(); b.star(nfas.back()); }
                        else if (n > m) {
                            repeat();
                            optional_top();
                            for (long k = m + 1; k < n; k++) repeat();
                        }
                        cp = e2;
                    } else cp = e1;
                    break;
                }
                case '^':
                case '$': push_atom(1ull, 0ull); break;          // both anchors: literal byte 0x00
                default: push_atom(ch < 64 ? 1ull << ch : 0ull, ch >= 64 ? 1ull << (ch - 64) : 0ull); break;
            }
            escaped = false;
            if (!(++cp < ps)) break;
        }
        clear_stack();
        if (nfas.size() != 1) fail("invalid expression");                   // Parser.cpp:155
        return nfas.back();
    }
};

}  // namespace

int reference_width_class(uint32_t n_raw) {
    return n_raw > 256 ? 3 : n_raw > 128 ? 2 : n_raw > 64 ? 1 : 0;          // Parser.cpp:165-168
}

Automaton compile_ere(const char* pattern) {
    if (!pattern) fail("invalid expression");
    Parser ps;
    ps.p = reinterpret_cast<const unsigned char*>(pattern);
    ps.ps = (int)std::strlen(pattern);
    for (int i = 0; i < ps.ps; i++)
        if (ps.p[i] >= 0x80) fail("invalid expression: pattern must be 7-bit ASCII");
    Sub top = ps.run();

    Automaton a;
    a.n_raw = top.size;
    a.raw_finals = top.finals;
    a.atoms = std::move(ps.b.atoms);
    a.P = (uint32_t)a.atoms.size();
    a.nullable = top.finals.has(top.init);
    a.first = Bits(a.P);
    a.accept = Bits(a.P);
    a.follow.assign(a.P, Bits(a.P));
    // Flattening, SURVEY.md A.4: positions = atoms (+ a virtual START that is never a target).
    for (uint32_t k = 0; k < a.P; k++) {
        const Atom& ak = a.atoms[k];
        if (ak.pre.has(top.init)) a.first.add(k);
        if (ak.post.intersects(top.finals)) a.accept.add(k);
        for (uint32_t j = 0; j < a.P; j++)
            if (a.atoms[j].pre.intersects(ak.post)) a.follow[k].add(j);
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------
// Device program construction
// ---------------------------------------------------------------------------------------------------
namespace {

struct Pos {                 // working copy of the position automaton (possibly reduced)
    uint64_t cls[2];
    Bits follow;             // over current position ids
    bool first, accept;
};

// Language-preserving reduction of the position automaton (LAZY-v1 depends on the language only).
// Two kinds of steps, repeated until nothing changes:
//  merge (R): positions with the same class, accept flag and successor set are folded into one
//             (forward-bisimulation quotient; e.g. the two copies the reference makes for "x+").
//  merge (L): positions with the same predecessor set (incl. START), successor set and accept flag
//             are always entered by the same moves and behave identically afterwards, so one
//             position with the union of their classes replaces them (e.g. a_i / b_i of every
//             level of "(a|b){1,300}c").
//  prune:     an edge p->q is dropped when p also has an edge to some q' != q that can do
//             everything q can: class(q) <= class(q'), accept(q) => accept(q'),
//             follow(q) <= follow(q').  Each removal is checked on the current automaton; any
//             accepting run through p->q can be rerouted through q'.  This turns the dense
//             "skip any number of optional copies" graph the reference builds for A{m,n}
//             (A^m (A?)^(n-m), Parser.cpp:123-141) into a chain.  START's edges (`first`) are
//             pruned by the same rule.
bool merge_round(std::vector<Pos>& v) {
    const size_t n = v.size();
    std::vector<Bits> pred(n, Bits((uint32_t)n));
    for (size_t k = 0; k < n; k++) for (uint32_t t : v[k].follow.list()) pred[t].add((uint32_t)k);
    std::vector<int> rep(n, -1);
    bool any = false;
    for (int pass = 0; pass < 2 && !any; pass++) {
        std::map<std::vector<uint32_t>, int> seen;
        for (size_t k = 0; k < n; k++) {
            std::vector<uint32_t> key;
            key.push_back(v[k].accept ? 1u : 0u);
            if (pass == 0) {
                key.push_back((uint32_t)v[k].cls[0]); key.push_back((uint32_t)(v[k].cls[0] >> 32));
                key.push_back((uint32_t)v[k].cls[1]); key.push_back((uint32_t)(v[k].cls[1] >> 32));
            } else {
                key.push_back(v[k].first ? 1u : 0u);
                key.push_back((v[k].cls[0] == 1ull && v[k].cls[1] == 0ull) ? 1u : 0u);     // an anchor (class {0x00}) only merges with anchors
                for (uint32_t w = 0; w < (n + 31) / 32; w++) key.push_back(pred[k].word(w));
            }
            for (uint32_t w = 0; w < (n + 31) / 32; w++) key.push_back(v[k].follow.word(w));
            auto it = seen.find(key);
            if (it == seen.end()) seen.emplace(std::move(key), (int)k);
            else { rep[k] = it->second; any = true; }
        }
        if (any && pass == 1)
            for (size_t k = 0; k < n; k++) if (rep[k] >= 0) { v[rep[k]].cls[0] |= v[k].cls[0]; v[rep[k]].cls[1] |= v[k].cls[1]; }
    }
    if (!any) return false;
    for (size_t k = 0; k < n; k++) if (rep[k] >= 0 && v[k].first) v[rep[k]].first = true;
    std::vector<uint32_t> newid(n, 0);
    uint32_t m = 0;
    for (size_t k = 0; k < n; k++) if (rep[k] < 0) newid[k] = m++;
    for (size_t k = 0; k < n; k++) if (rep[k] >= 0) newid[k] = newid[rep[k]];
    std::vector<Pos> nv;
    nv.reserve(m);
    for (size_t k = 0; k < n; k++) {
        if (rep[k] >= 0) continue;
        Pos q = v[k];
        Bits nf(m);
        for (uint32_t t : v[k].follow.list()) nf.add(newid[t]);
        q.follow = nf;
        nv.push_back(q);
    }
    v.swap(nv);
    return true;
}

bool subset_of(const Bits& a, const Bits& b) {
    for (size_t i = 0; i < a.w.size(); i++) if (a.w[i] & ~b.word((uint32_t)i)) return false;
    return true;
}
bool dominated(const Pos& q, const Pos& by) {
    if ((q.cls[0] & ~by.cls[0]) || (q.cls[1] & ~by.cls[1])) return false;
    if (q.accept && !by.accept) return false;
    return subset_of(q.follow, by.follow);
}
bool prune_round(std::vector<Pos>& v) {
    const size_t n = v.size();
    bool any = false;
    // START's edges
    {
        std::vector<uint32_t> fs;
        for (size_t k = 0; k < n; k++) if (v[k].first) fs.push_back((uint32_t)k);
        for (uint32_t q : fs) {
            for (uint32_t q2 : fs) {
                if (q2 == q || !v[q2].first || !v[q].first) continue;
                if (dominated(v[q], v[q2])) { v[q].first = false; any = true; break; }
            }
        }
    }
    for (size_t p = 0; p < n; p++) {
        std::vector<uint32_t> succ = v[p].follow.list();
        for (uint32_t q : succ) {
            for (uint32_t q2 : succ) {
                if (q2 == q || !v[p].follow.has(q2)) continue;
                if (dominated(v[q], v[q2])) { v[p].follow.del(q); any = true; break; }
            }
        }
    }
    return any;
}
// Drop positions that cannot be reached from START or cannot reach an accepting position.
bool trim_round(std::vector<Pos>& v) {
    const size_t n = v.size();
    std::vector<char> fwd(n, 0), bwd(n, 0);
    std::vector<uint32_t> st;
    for (size_t k = 0; k < n; k++) if (v[k].first) { fwd[k] = 1; st.push_back((uint32_t)k); }
    while (!st.empty()) { uint32_t k = st.back(); st.pop_back(); for (uint32_t t : v[k].follow.list()) if (!fwd[t]) { fwd[t] = 1; st.push_back(t); } }
    std::vector<std::vector<uint32_t>> pred(n);
    for (size_t k = 0; k < n; k++) for (uint32_t t : v[k].follow.list()) pred[t].push_back((uint32_t)k);
    for (size_t k = 0; k < n; k++) if (v[k].accept) { bwd[k] = 1; st.push_back((uint32_t)k); }
    while (!st.empty()) { uint32_t k = st.back(); st.pop_back(); for (uint32_t t : pred[k]) if (!bwd[t]) { bwd[t] = 1; st.push_back(t); } }
    std::vector<uint32_t> newid(n, 0);
    uint32_t m = 0;
    for (size_t k = 0; k < n; k++) if (fwd[k] && bwd[k]) newid[k] = m++;
    if (m == n || m == 0) return false;     // keep at least the shape for an empty language
    std::vector<Pos> nv;
    for (size_t k = 0; k < n; k++) {
        if (!(fwd[k] && bwd[k])) continue;
        Pos q = v[k];
        Bits nf(m);
        for (uint32_t t : v[k].follow.list()) if (fwd[t] && bwd[t]) nf.add(newid[t]);
        q.follow = nf;
        nv.push_back(q);
    }
    v.swap(nv);
    return true;
}
void reduce_positions(std::vector<Pos>& v) {
    for (int guard = 0; guard < 4096; guard++) {
        bool ch = false;
        while (merge_round(v)) ch = true;
        if (prune_round(v)) ch = true;
        if (trim_round(v)) ch = true;
        if (!ch) break;
    }
}

struct Decomp {
    struct Term { uint32_t kind, r; Bits mask; };
    struct Bicl { Bits src, dst; };
    std::vector<Term> terms;
    std::vector<Bicl> bicl;
    bool use_rows = false;
    Bits rowsrc;
    std::vector<Bits> rows;      // uncovered edges per source
};

// Cover the edge relation `edges[k]` (targets of k) by shift terms, bicliques and (if those run
// out) per-source rows.  Every structure only ever contains true edges, so any cover is exact.
Decomp decompose(const std::vector<Bits>& edges, uint32_t P, uint32_t NW) {
    Decomp d;
    std::vector<Bits> rem = edges;
    // distance histogram
    std::map<int, uint32_t> hist;
    auto dist_key = [&](uint32_t k, uint32_t t) -> int {
        int dd = (int)t - (int)k;
        if (NW == 1) return ((dd % 32) + 32) % 32;           // rotation amount
        return dd;
    };
    for (uint32_t k = 0; k < P; k++) for (uint32_t t : edges[k].list()) hist[dist_key(k, t)]++;
    std::vector<std::pair<uint32_t, int>> order;
    for (auto& kv : hist) {
        if (NW > 1 && (kv.first > 31 || kv.first < -31)) continue;   // multiword shifts stay within a neighbour word
        order.push_back({kv.second, kv.first});
    }
    std::sort(order.begin(), order.end(), [](auto& a, auto& b) { return a.first != b.first ? a.first > b.first : std::abs(a.second) < std::abs(b.second); });
    for (auto& oc : order) {
        if (d.terms.size() >= RRX_MAX_TERMS) break;
        if (oc.first < 2 && !d.terms.empty()) break;         // single edges are left to the bicliques
        Decomp::Term t;
        int key = oc.second;
        t.mask = Bits(P);
        for (uint32_t k = 0; k < P; k++)
            for (uint32_t tg : edges[k].list())
                if (dist_key(k, tg) == key) { t.mask.add(tg); rem[k].del(tg); }
        if (NW == 1) { t.kind = TERM_ROT; t.r = (uint32_t)key; }
        else if (key == 0) { t.kind = TERM_SELF; t.r = 0; }
        else if (key > 0) { t.kind = TERM_SHL; t.r = (uint32_t)key; }
        else { t.kind = TERM_SHR; t.r = (uint32_t)(-key); }
        d.terms.push_back(t);
    }
    // bicliques: sources with identical remaining target sets
    std::vector<bool> done(P, false);
    for (uint32_t k = 0; k < P; k++) {
        if (done[k] || !rem[k].any()) continue;
        Decomp::Bicl bc;
        bc.src = Bits(P); bc.dst = rem[k];
        for (uint32_t j = k; j < P; j++)
            if (!done[j] && rem[j].equals(rem[k])) { bc.src.add(j); done[j] = true; }
        // a lone edge is cheaper as a shift term (shift+and) than as a biclique (and+test+select)
        if (bc.src.count() == 1 && bc.dst.count() == 1 && d.terms.size() < 4) {
            const uint32_t sk = bc.src.list()[0], tk = bc.dst.list()[0];
            const int key = dist_key(sk, tk);
            if (NW == 1 || (key <= 31 && key >= -31)) {
                Decomp::Term t;
                t.mask = Bits(P);
                t.mask.add(tk);
                if (NW == 1) { t.kind = TERM_ROT; t.r = (uint32_t)key; }
                else if (key == 0) { t.kind = TERM_SELF; t.r = 0; }
                else if (key > 0) { t.kind = TERM_SHL; t.r = (uint32_t)key; }
                else { t.kind = TERM_SHR; t.r = (uint32_t)(-key); }
                d.terms.push_back(t);
                continue;
            }
        }
        d.bicl.push_back(bc);
    }
    if (d.bicl.size() > RRX_MAX_BICL) {
        // too irregular: keep the cheapest bicliques (most edges) and push the rest to rows
        std::sort(d.bicl.begin(), d.bicl.end(), [](const Decomp::Bicl& a, const Decomp::Bicl& b) {
            return (uint64_t)a.src.count() * a.dst.count() > (uint64_t)b.src.count() * b.dst.count(); });
        d.use_rows = true;
        d.rowsrc = Bits(P);
        d.rows.assign(P, Bits(P));
        for (size_t i = RRX_MAX_BICL; i < d.bicl.size(); i++)
            for (uint32_t s : d.bicl[i].src.list()) { d.rowsrc.add(s); d.rows[s].unite(d.bicl[i].dst); }
        d.bicl.resize(RRX_MAX_BICL);
    }
    if (d.rows.empty()) d.rows.assign(P, Bits(P));
    if (d.rowsrc.w.empty()) d.rowsrc = Bits(P);
    return d;
}

void put_bits(std::vector<uint32_t>& blob, uint32_t off, const Bits& b, uint32_t NW) {
    for (uint32_t w = 0; w < NW; w++) blob[off + w] = b.word(w);
}

}  // namespace

namespace {
// ---- cut analysis for the candidate-driven row scan (rrx_tile.cuh phase_a_sparse) -----------------
// For a byte value v let K_v = positions whose class is exactly {v}.  K_v is a CUT when no activation
// chain (start at a first-position, follow pruned edges) reaches an accept position without
// visiting K_v.  Then D = positions every chain to which passes through K_v ("dominated"; accept is
// inside D), and the scan only has to track D-states, which are empty except shortly after an
// occurrence of v.  Among the cuts the cheapest by a static estimate (typical frequency of v x
// expected length of the dominated run + of the backward verification) is chosen.
struct Cut {
    bool ok = false;
    uint32_t byte = 0;
    bool run_broad = false;      // a run of dominated states can loop on a broad byte class: it need never die
    bool ver_broad = false;      // the backward verification walk can loop on a broad byte class without finding a start
    Bits D;
    struct Group { Bits k, pred; uint32_t flags; };
    std::vector<Group> groups;
};

uint32_t byte_freq_score(uint32_t v) {
    if (v == ' ') return 10;
    if (v >= 'a' && v <= 'z') return 8;
    if (v >= '0' && v <= '9') return 4;
    if (v >= 'A' && v <= 'Z') return 3;
    if (v == '.' || v == ',' || v == '-' || v == '_' || v == '/' || v == ':' || v == '=' || v == '"' || v == '\'') return 3;
    if (v == '\n' || v == '\t') return 2;
    return 1;
}

// longest path (in nodes) inside `nodes` starting anywhere in `from`, following `edges`; capped at
// `cap`, which is also returned when a cycle is reachable
uint32_t longest_run(const std::vector<Bits>& edges, const Bits& nodes, const Bits& from, uint32_t cap) {
    const uint32_t P = (uint32_t)edges.size();
    std::vector<int> colour(P, 0);
    std::vector<uint32_t> depth(P, 0);
    uint32_t best = 0;
    bool cyclic = false;
    for (uint32_t root : from.list()) {
        if (!nodes.has(root) || colour[root]) continue;
        std::vector<std::pair<uint32_t, size_t>> st;
        st.push_back({root, 0});
        colour[root] = 1;
        while (!st.empty() && !cyclic) {
            const uint32_t k = st.back().first;
            std::vector<uint32_t> succ;
            for (uint32_t t : edges[k].list()) if (nodes.has(t)) succ.push_back(t);
            if (st.back().second < succ.size()) {
                const uint32_t t = succ[st.back().second++];
                if (colour[t] == 1) cyclic = true;
                else if (colour[t] == 0) { colour[t] = 1; st.push_back({t, 0}); }
            } else {
                uint32_t d = 1;
                for (uint32_t t : succ) d = std::max(d, depth[t] + 1);
                depth[k] = d;
                colour[k] = 2;
                st.pop_back();
            }
        }
        if (cyclic) return cap;
        best = std::max(best, depth[root]);
    }
    return std::min(best, cap);
}

// Does the sub-graph of `nodes` reachable from `from` contain a cycle through a position whose byte class
// is broad (>= 96 of the 128 byte values, e.g. '.')?  Such a loop keeps a walk alive on almost any text.
bool has_broad_cycle(const std::vector<Pos>& pos, const std::vector<Bits>& edges, const Bits& nodes, const Bits& from) {
    const uint32_t P = (uint32_t)edges.size();
    Bits reach(P);
    std::vector<uint32_t> work;
    for (uint32_t k : from.list()) if (nodes.has(k) && !reach.has(k)) { reach.add(k); work.push_back(k); }
    while (!work.empty()) {
        const uint32_t k = work.back();
        work.pop_back();
        for (uint32_t t : edges[k].list()) if (nodes.has(t) && !reach.has(t)) { reach.add(t); work.push_back(t); }
    }
    for (uint32_t b : reach.list()) {
        if (__builtin_popcountll(pos[b].cls[0]) + __builtin_popcountll(pos[b].cls[1]) < 96) continue;
        Bits seen(P);
        std::vector<uint32_t> st;
        for (uint32_t t : edges[b].list()) if (nodes.has(t) && !seen.has(t)) { seen.add(t); st.push_back(t); }
        while (!st.empty()) {
            const uint32_t k = st.back();
            st.pop_back();
            if (k == b) return true;
            for (uint32_t t : edges[k].list()) if (nodes.has(t) && !seen.has(t)) { seen.add(t); st.push_back(t); }
        }
        if (seen.has(b)) return true;
    }
    return false;
}

Cut find_cut(const std::vector<Pos>& pos, const std::vector<Bits>& fol, const std::vector<Bits>& rfol,
             const Bits& first, const Bits& accept) {
    const uint32_t P = (uint32_t)pos.size();
    Cut best;
    uint64_t best_cost = ~0ull;
    Bits all(P);
    for (uint32_t k = 0; k < P; k++) all.add(k);
    for (uint32_t v = 1; v < 128; v++) {
        Bits K(P);
        for (uint32_t k = 0; k < P; k++) {
            const uint64_t lo = v < 64 ? 1ull << v : 0ull, hi = v < 64 ? 0ull : 1ull << (v - 64);
            if (pos[k].cls[0] == lo && pos[k].cls[1] == hi) K.add(k);
        }
        if (!K.any()) continue;
        // positions reachable from first while avoiding K
        Bits nd(P);
        std::vector<uint32_t> work;
        for (uint32_t k : first.list()) if (!K.has(k)) { nd.add(k); work.push_back(k); }
        while (!work.empty()) {
            const uint32_t k = work.back();
            work.pop_back();
            for (uint32_t t : fol[k].list()) if (!K.has(t) && !nd.has(t)) { nd.add(t); work.push_back(t); }
        }
        if (nd.intersects(accept)) continue;                 // some match avoids v: not a cut
        Cut c;
        c.ok = true;
        c.byte = v;
        c.D = Bits(P);
        for (uint32_t k = 0; k < P; k++) if (!nd.has(k)) c.D.add(k);
        // the K positions that can be entered at all: first-positions, or with a predecessor
        bool too_many = false;
        for (uint32_t k : K.list()) {
            const bool is_first = first.has(k);
            if (!is_first && !rfol[k].any()) continue;       // unreachable
            Bits pred = is_first ? Bits(P) : rfol[k];
            bool placed = false;
            for (auto& g : c.groups)
                if (g.flags == (is_first ? 1u : 0u) && g.pred.equals(pred)) { g.k.add(k); placed = true; break; }
            if (!placed) {
                if (c.groups.size() == RRX_MAX_GROUPS) { too_many = true; break; }
                Cut::Group g;
                g.k = Bits(P); g.k.add(k); g.pred = pred; g.flags = is_first ? 1u : 0u;
                c.groups.push_back(g);
            }
        }
        if (too_many || c.groups.empty()) continue;
        // static cost estimate.  A run of dominated states that can loop on a broad class ("z.*q" cut at z)
        // leaves every row open — far worse than a verification walk that loops on one ("z.*q" cut at q: the
        // walk back to the last z has a skip state and costs about what any cyclic walk is charged, 16).
        const uint32_t run = longest_run(fol, c.D, K, 16);
        uint32_t ver = 0;
        Bits nonfirst(P);
        for (uint32_t k = 0; k < P; k++) if (!first.has(k)) nonfirst.add(k);
        for (auto& g : c.groups) if (!(g.flags & 1u)) {
            ver = std::max(ver, longest_run(rfol, all, g.pred, 16));
            if (has_broad_cycle(pos, rfol, nonfirst, g.pred)) c.ver_broad = true;
        }
        c.run_broad = has_broad_cycle(pos, fol, c.D, K);
        // ties: the smaller dominated set (more of the automaton is skipped between candidates)
        const uint64_t cost = (uint64_t)byte_freq_score(v) * ((c.run_broad ? 1024u : run) + ver) * 1024 + c.D.count();
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    // not worth it when candidates are expected every few bytes and each starts a long run
    if (best.ok && best_cost >= 8ull * 24ull * 1024ull) best.ok = false;
    return best;
}

// ---- forward skip state of the sparse row scan ------------------------------------------------------
// Runs of dominated states start at an injected cut group and evolve by  X' = follow'(X) & D & B[c]
// (no injection inside a run unless another candidate byte comes up).  Explore the reachable X and
// look for one that a whole range of bytes maps to itself, match-free.
struct FSkip {
    uint32_t n = 0, lo = 0, hi = 0;
    Bits G;
};

FSkip find_forward_skip(const std::vector<Pos>& pos, const std::vector<Bits>& fol, const Cut& cut, const Bits& accept) {
    FSkip best;
    if (!cut.ok) return best;
    const uint32_t P = (uint32_t)pos.size();
    std::vector<Bits> Bc(128, Bits(P));
    for (uint32_t c = 0; c < 128; c++)
        for (uint32_t k = 0; k < P; k++)
            if (((pos[k].cls[c >> 6] >> (c & 63)) & 1ull) && cut.D.has(k)) Bc[c].add(k);
    auto step = [&](const Bits& X, uint32_t c) {
        Bits nx(P), out(P);
        for (uint32_t k : X.list()) nx.unite(fol[k]);
        for (uint32_t k : nx.list()) if (Bc[c].has(k)) out.add(k);
        return out;
    };
    std::vector<Bits> seen, queue;
    for (auto& g : cut.groups) {
        bool known = false;
        for (auto& s : seen) if (s.equals(g.k)) known = true;
        if (!known) { seen.push_back(g.k); queue.push_back(g.k); }
    }
    for (size_t qi = 0; qi < queue.size() && qi < 48; qi++) {
        const Bits X = queue[qi];
        if (!X.any()) continue;
        if (!X.intersects(accept)) {
            uint32_t run_lo = 0, run_len = 0;
            for (uint32_t c = 1; c <= 128; c++) {
                const bool stay = c < 128 && c != cut.byte && step(X, c).equals(X);
                if (stay) { if (!run_len) run_lo = c; run_len++; }
                else {
                    if (run_len >= 4 && run_len > best.n) { best.n = run_len; best.lo = run_lo; best.hi = run_lo + run_len - 1; best.G = X; }
                    run_len = 0;
                }
            }
        }
        if (seen.size() < 96)
            for (uint32_t c = 1; c < 128; c++) {
                if (c == cut.byte) continue;
                Bits nx = step(X, c);
                if (!nx.any()) continue;
                bool known = false;
                for (auto& s : seen) if (s.equals(nx)) { known = true; break; }
                if (!known) { seen.push_back(nx); queue.push_back(nx); }
            }
    }
    return best;
}

// ---- reverse-walk skip state -----------------------------------------------------------------------
// The lazy start scan walks the reversed automaton from the accepting positions:  x = R & B[c];
// found if x meets first; R = rfollow'(x).  Explore the reverse state sets reachable from `accept`
// (breadth first, bounded) and look for one that almost every byte leaves unchanged: in that state
// the walk can run over the text looking only for the few bytes that matter ("z.*q": after the q the
// state is {z, .} until a z turns up).
struct RSkip {
    uint32_t n = 0;
    uint32_t bytes[2] = {0, 0};
    Bits F;
};

RSkip find_reverse_skip(const std::vector<Pos>& pos, const std::vector<Bits>& rfol, const Bits& first, const Bits& accept,
                        const std::vector<Bits>& more_seeds) {
    const uint32_t P = (uint32_t)pos.size();
    std::vector<Bits> Bc(128, Bits(P));
    for (uint32_t c = 0; c < 128; c++)
        for (uint32_t k = 0; k < P; k++)
            if ((pos[k].cls[c >> 6] >> (c & 63)) & 1ull) Bc[c].add(k);
    auto step = [&](const Bits& R, uint32_t c, bool* hit) {
        Bits x(P), nx(P);
        for (uint32_t k : R.list()) if (Bc[c].has(k)) x.add(k);
        *hit = x.intersects(first);
        for (uint32_t k : x.list()) nx.unite(rfol[k]);
        return nx;
    };
    std::vector<Bits> seen;
    std::vector<Bits> queue;
    queue.push_back(accept);
    seen.push_back(accept);
    for (auto& s : more_seeds) {          // the verification walks of the candidate scan start at the cut groups' predecessors
        bool known = false;
        for (auto& k : seen) if (k.equals(s)) known = true;
        if (!known && s.any()) { queue.push_back(s); seen.push_back(s); }
    }
    RSkip best;
    for (size_t qi = 0; qi < queue.size() && qi < 48; qi++) {
        const Bits R = queue[qi];
        if (!R.any()) continue;
        std::vector<uint32_t> odd;           // bytes that change R or find a start
        for (uint32_t c = 0; c < 128 && odd.size() <= 2; c++) {
            bool hit = false;
            Bits nx = step(R, c, &hit);
            if (hit || !nx.equals(R)) odd.push_back(c);
        }
        if (odd.size() >= 1 && odd.size() <= 2 && (best.n == 0 || odd.size() < best.n)) {
            best.n = (uint32_t)odd.size();
            best.bytes[0] = odd[0];
            best.bytes[1] = odd.size() > 1 ? odd[1] : odd[0];
            best.F = R;
        }
        if (seen.size() < 96)
            for (uint32_t c = 0; c < 128; c++) {
                bool hit = false;
                Bits nx = step(R, c, &hit);
                if (hit || !nx.any()) continue;                 // the walk ends there
                bool known = false;
                for (auto& s : seen) if (s.equals(nx)) { known = true; break; }
                if (!known) { seen.push_back(nx); queue.push_back(nx); }
            }
    }
    return best;
}

}  // namespace

Program build_program(const Automaton& a, const std::string& pattern, bool reduce, bool allow_donly) {
    std::vector<Pos> pos(a.P);
    for (uint32_t k = 0; k < a.P; k++) {
        pos[k].cls[0] = a.atoms[k].cls[0]; pos[k].cls[1] = a.atoms[k].cls[1];
        pos[k].follow = a.follow[k];
        pos[k].first = a.first.has(k);
        pos[k].accept = a.accept.has(k);
    }
    if (reduce) reduce_positions(pos);
    const uint32_t P = (uint32_t)pos.size();
    if (P == 0) fail("invalid expression");
    if (P > 1024) fail("invalid expression: more than 1024 positions");
    // supported state-set widths: 1,2,4,8,16,32 words of 32 bits
    uint32_t NW = 1;
    while (NW * 32 < P) NW *= 2;
    if (NW > 1) {
        // Multi-word automaton: number the positions a cut dominates FIRST (stable otherwise), so that everything
        // the thread-owned scan carries fits the first CW words (rrx_prog.h).  Renumbering changes no language.
        // (Done whether or not the thread-owned scan is allowed: a scanner's fallback program must number the positions
        // like its primary one — entry sets and composites pass between the two.)
        Bits first0(P), accept0(P);
        std::vector<Bits> fol0(P, Bits(P)), rfol0(P, Bits(P));
        for (uint32_t k = 0; k < P; k++) { if (pos[k].first) first0.add(k); if (pos[k].accept) accept0.add(k); }
        for (uint32_t k = 0; k < P; k++)
            for (uint32_t t : pos[k].follow.list())
                if (!first0.has(t)) { fol0[k].add(t); rfol0[t].add(k); }
        const Cut c0 = find_cut(pos, fol0, rfol0, first0, accept0);
        if (c0.ok && !c0.run_broad && c0.D.count() < P) {
            std::vector<uint32_t> order, newid(P);
            for (uint32_t k = 0; k < P; k++) if (c0.D.has(k)) order.push_back(k);
            for (uint32_t k = 0; k < P; k++) if (!c0.D.has(k)) order.push_back(k);
            for (uint32_t i = 0; i < P; i++) newid[order[i]] = i;
            std::vector<Pos> np(P);
            for (uint32_t i = 0; i < P; i++) {
                np[i] = pos[order[i]];
                Bits nf(P);
                for (uint32_t t : pos[order[i]].follow.list()) nf.add(newid[t]);
                np[i].follow = nf;
            }
            pos.swap(np);
        }
    }

    Bits first(P), accept(P), all(P);
    for (uint32_t k = 0; k < P; k++) {
        if (pos[k].first) first.add(k);
        if (pos[k].accept) accept.add(k);
        all.add(k);
    }
    // pruned edges E' and their transpose
    std::vector<Bits> fol(P, Bits(P)), rfol(P, Bits(P));
    for (uint32_t k = 0; k < P; k++)
        for (uint32_t t : pos[k].follow.list())
            if (!first.has(t)) { fol[k].add(t); rfol[t].add(k); }
    Decomp fd = decompose(fol, P, NW), rd = decompose(rfol, P, NW);

    ProgHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = RRX_PROG_MAGIC; h.P = P; h.NW = NW;
    h.n_terms = (uint32_t)fd.terms.size(); h.n_bicl = (uint32_t)fd.bicl.size(); h.use_rows = fd.use_rows;
    h.n_rterms = (uint32_t)rd.terms.size(); h.n_rbicl = (uint32_t)rd.bicl.size(); h.use_rrows = rd.use_rows;
    h.n_raw = a.n_raw; h.ref_width_class = (uint32_t)reference_width_class(a.n_raw); h.nullable = a.nullable;
    h.n_flat = a.P + 1;
    {   // longest path of the pruned follow graph (iterative DFS with colours); cyclic -> 0xffffffff
        std::vector<int> colour(P, 0), depth(P, 0);
        bool cyclic = false;
        uint32_t best = 0;
        for (uint32_t root = 0; root < P && !cyclic; root++) {
            if (colour[root]) continue;
            std::vector<std::pair<uint32_t, size_t>> st;
            std::vector<std::vector<uint32_t>> succ_cache;
            st.push_back({root, 0});
            colour[root] = 1;
            while (!st.empty() && !cyclic) {
                uint32_t k = st.back().first;
                std::vector<uint32_t> succ = fol[k].list();
                if (st.back().second < succ.size()) {
                    uint32_t t = succ[st.back().second++];
                    if (colour[t] == 1) cyclic = true;
                    else if (colour[t] == 0) { colour[t] = 1; st.push_back({t, 0}); }
                } else {
                    int d = 0;
                    for (uint32_t t : succ) d = std::max(d, depth[t] + 1);
                    depth[k] = d;
                    best = std::max(best, (uint32_t)d);
                    colour[k] = 2;
                    st.pop_back();
                }
            }
        }
        h.static_depth = cyclic ? 0xffffffffu : best;
    }
    uint32_t off = (uint32_t)(sizeof(ProgHeader) / 4);
    auto take = [&](uint32_t words) { uint32_t o = off; off += words; return o; };
    h.off_first = take(NW); h.off_accept = take(NW); h.off_all = take(NW);
    h.off_terms = take(RRX_MAX_TERMS * (1 + NW)); h.off_bicl = take(RRX_MAX_BICL * 2 * NW); h.off_rowsrc = take(NW);
    h.off_rterms = take(RRX_MAX_TERMS * (1 + NW)); h.off_rbicl = take(RRX_MAX_BICL * 2 * NW); h.off_rrowsrc = take(NW);
    const Cut cut = find_cut(pos, fol, rfol, first, accept);
    h.off_sparse = take(NW + RRX_MAX_GROUPS * (2 * NW + 1));
    h.sparse_ok = cut.ok ? 1u : 0u;
    h.cand_byte = cut.byte;
    h.n_groups = cut.ok ? (uint32_t)cut.groups.size() : 0u;
    std::vector<Bits> vseeds;
    if (cut.ok) for (auto& g : cut.groups) if (!(g.flags & 1u)) vseeds.push_back(g.pred);
    const RSkip rskip = find_reverse_skip(pos, rfol, first, accept, vseeds);
    // thread-owned candidate scan (rrx_prog.h `donly`): runs of dominated states must die by themselves, and the
    // verification walk must be short or have a skip state to run on
    h.donly = (allow_donly && cut.ok && !cut.run_broad && (!cut.ver_broad || rskip.n >= 1)) ? 1u : 0u;
    h.CW = NW;
    if (h.donly) {
        uint32_t top = 0;
        for (uint32_t k : cut.D.list()) top = k;
        uint32_t cw = 1;
        while (cw * 32 <= top) cw *= 2;
        h.CW = cw < NW ? cw : NW;
    }
    const FSkip fskip = find_forward_skip(pos, fol, cut, accept);
    h.off_fskip = take(NW);
    h.fskip_n = fskip.n; h.fskip_lo = fskip.lo; h.fskip_hi = fskip.hi;
    h.off_rskip = take(NW);
    h.rskip_n = rskip.n;
    h.rskip_bytes = rskip.bytes[0] | (rskip.bytes[1] << 8);
    h.off_B = take(256 * NW);
    h.off_follow = take(P * NW); h.off_rfollow = take(P * NW);
    h.off_rows = take(P * NW); h.off_rrows = take(P * NW);
    h.off_rfull = take(P * NW);
    h.off_anch = take(NW);
    h.total_words = off;

    Program pr;
    pr.pattern = pattern;
    pr.blob.assign(off, 0u);
    std::memcpy(pr.blob.data(), &h, sizeof(h));
    put_bits(pr.blob, h.off_first, first, NW);
    put_bits(pr.blob, h.off_accept, accept, NW);
    put_bits(pr.blob, h.off_all, all, NW);
    if (cut.ok) {
        put_bits(pr.blob, h.off_sparse, cut.D, NW);
        for (size_t g = 0; g < cut.groups.size(); g++) {
            const uint32_t o = h.off_sparse + NW + (uint32_t)g * (2 * NW + 1);
            put_bits(pr.blob, o, cut.groups[g].k, NW);
            put_bits(pr.blob, o + NW, cut.groups[g].pred, NW);
            pr.blob[o + 2 * NW] = cut.groups[g].flags;
        }
    }
    {
        Bits anch(P);
        for (uint32_t k = 0; k < P; k++) if (pos[k].cls[0] == 1ull && pos[k].cls[1] == 0ull) anch.add(k);
        put_bits(pr.blob, h.off_anch, anch, NW);
        reinterpret_cast<ProgHeader*>(pr.blob.data())->n_anchor = anch.count();
    }
    if (rskip.n) put_bits(pr.blob, h.off_rskip, rskip.F, NW);
    if (fskip.n) put_bits(pr.blob, h.off_fskip, fskip.G, NW);
    auto put_decomp = [&](const Decomp& d, uint32_t o_terms, uint32_t o_bicl, uint32_t o_rowsrc, uint32_t o_rows) {
        for (size_t i = 0; i < d.terms.size(); i++) {
            uint32_t o = o_terms + (uint32_t)i * (1 + NW);
            pr.blob[o] = (d.terms[i].kind << 8) | d.terms[i].r;
            put_bits(pr.blob, o + 1, d.terms[i].mask, NW);
        }
        for (size_t i = 0; i < d.bicl.size(); i++) {
            uint32_t o = o_bicl + (uint32_t)i * 2 * NW;
            put_bits(pr.blob, o, d.bicl[i].src, NW);
            put_bits(pr.blob, o + NW, d.bicl[i].dst, NW);
        }
        put_bits(pr.blob, o_rowsrc, d.rowsrc, NW);
        for (uint32_t k = 0; k < P; k++) put_bits(pr.blob, o_rows + k * NW, d.rows[k], NW);
    };
    put_decomp(fd, h.off_terms, h.off_bicl, h.off_rowsrc, h.off_rows);
    put_decomp(rd, h.off_rterms, h.off_rbicl, h.off_rrowsrc, h.off_rrows);
    for (uint32_t c = 0; c < 128; c++)
        for (uint32_t k = 0; k < P; k++)
            if ((pos[k].cls[c >> 6] >> (c & 63)) & 1ull) pr.blob[h.off_B + c * NW + (k >> 5)] |= 1u << (k & 31);
    std::vector<Bits> rfull(P, Bits(P));
    for (uint32_t k = 0; k < P; k++) for (uint32_t t : pos[k].follow.list()) rfull[t].add(k);
    for (uint32_t k = 0; k < P; k++) {
        put_bits(pr.blob, h.off_follow + k * NW, fol[k], NW);
        put_bits(pr.blob, h.off_rfollow + k * NW, rfol[k], NW);
        put_bits(pr.blob, h.off_rfull + k * NW, rfull[k], NW);
    }
    return pr;
}

}  // namespace rrx
