// Drop-in C++ front end: the reference's public interface (RoaringRegex src/inc/regex.h:100-126,
// 212-228) re-expressed over the C ABI of librrx.so, so that code written against the reference —
// e.g. its test program src/test/main.cpp — compiles unchanged against this header.
//
//   Regex::RRegex(const char* pattern)            regex.h:224 / Parser.cpp:161; throws std::runtime_error
//                                                 with the reference's messages (Parser.cpp:36,155)
//   IteratorWrapper get_acceptance_iter(char*)    regex.h:225: `++` consumes the NUL-terminated text,
//                                                 `*` = Match{text, text+len} iff the whole string matches
//   IteratorWrapper get_lazy_iter(char*)          NEW, same pattern: the lazy-match iterator the
//                                                 reference's README promises and src/Iterator.cpp leaves
//                                                 empty; `++` advances to the next LAZY-v1 match,
//                                                 `*` returns it, or nullopt when exhausted
//   IteratorWrapper get_nonoverlap_iter(char*)    NEW: the finditer-style view of the same list (SURVEY.md Appendix D.2:
//                                                 p = 0; for (s,e) in LAZY-v1: if s >= p { emit; p = e }), filtered on the device
//   Match{char* start; char* end; str()}          regex.h:100-105, half-open, pointers into the caller's text
//   RegexIterator / IteratorWrapper / IterFactoryBase   regex.h:106-126 (virtual destructors added: C4 of
//                                                 SURVEY.md Appendix C is undefined behaviour in the reference)
//
// Lifetime rules are the reference's: iterators must not outlive the RRegex, the text is borrowed.
// All matching runs on the GPU (librrx has no CPU path); failures surface as std::runtime_error.
#pragma once
#include <cstring>
#include <iostream>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>

#include "rrx.h"

namespace Regex {

class Match {
public:
    char* start;
    char* end;
    std::string str() { return std::string(start, (size_t)(end - start)); }
};

class RegexIterator {
public:
    virtual RegexIterator& operator++(int) = 0;          // process forward
    virtual RegexIterator* create_copy() = 0;
    virtual std::optional<Match> operator*() = 0;
    virtual void print() = 0;
    virtual ~RegexIterator() {}
};

class IteratorWrapper : public RegexIterator {
    std::unique_ptr<RegexIterator> concrete;
public:
    IteratorWrapper(RegexIterator* ptr) : concrete(std::unique_ptr<RegexIterator>(ptr)) {}
    IteratorWrapper(IteratorWrapper& to_copy) : concrete(std::unique_ptr<RegexIterator>(to_copy.create_copy())) {}
    IteratorWrapper(IteratorWrapper&& o) = default;
    IteratorWrapper& operator++(int) { (*concrete)++; return *this; }
    std::optional<Match> operator*() { return **concrete; }
    void print() { concrete->print(); }
    RegexIterator* create_copy() { return concrete->create_copy(); }
};

class IterFactoryBase {
public:
    virtual IteratorWrapper get_acceptance_iter(char* c) = 0;
    virtual IteratorWrapper get_lazy_iter(char* c) = 0;
    virtual IteratorWrapper get_nonoverlap_iter(char* c) = 0;
    virtual ~IterFactoryBase() {}
};

namespace detail {
inline void check(int rc) { if (rc != RRX_OK) throw std::runtime_error(rrx_last_error()); }

inline void print_program(const rrx_program* p) {
    rrx_info i;
    rrx_program_info(p, &i);
    std::cout << "raw states: " << i.n_raw << "  flattened: " << i.n_flat << "  device positions: " << i.n_device
              << " (" << i.words << " x 32-bit words)  follow terms/bicliques: " << i.n_terms << "/" << i.n_bicliques << std::endl;
}

class AcceptanceIterator : public RegexIterator {
    const rrx_program* prog;
    char* text;
    char* current;
    int accepted;       // -1: nothing consumed yet
public:
    AcceptanceIterator(char* t, const rrx_program* p) : prog(p), text(t), current(t), accepted(-1) {}
    RegexIterator& operator++(int) {
        const size_t n = std::strlen(current);
        current += n;
        check(rrx_accepts_host(prog, reinterpret_cast<const uint8_t*>(text), (uint64_t)(current - text), &accepted));
        return *this;
    }
    std::optional<Match> operator*() {
        if (accepted < 0) {
            rrx_info i;
            rrx_program_info(prog, &i);
            return i.nullable ? std::optional<Match>(Match{text, text}) : std::nullopt;
        }
        return accepted ? std::optional<Match>(Match{text, current}) : std::nullopt;
    }
    RegexIterator* create_copy() { return new AcceptanceIterator(*this); }
    void print() { print_program(prog); }
};

class LazyIterator : public RegexIterator {
    const rrx_program* prog;
    char* text;
    int mode;                              // RRX_MODE_LAZY_V1 or RRX_MODE_NO_OVERLAP
    std::shared_ptr<rrx_matches> list;     // produced by ONE device scan on first use, shared by copies
    const int64_t* pairs = nullptr;
    uint64_t count = 0;
    int64_t pos = -1;
    void ensure() {
        if (list) return;
        rrx_matches* m = nullptr;
        check(rrx_scan_host(prog, reinterpret_cast<const uint8_t*>(text), std::strlen(text), 0, mode, &m));
        list = std::shared_ptr<rrx_matches>(m, rrx_matches_free);
        count = rrx_matches_count(m);
        pairs = rrx_matches_data(m);
        if (count && !pairs) throw std::runtime_error(rrx_last_error());
    }
public:
    LazyIterator(char* t, const rrx_program* p, int mode_ = RRX_MODE_LAZY_V1) : prog(p), text(t), mode(mode_) {}
    RegexIterator& operator++(int) { ensure(); if (pos < (int64_t)count) pos++; return *this; }
    std::optional<Match> operator*() {
        ensure();
        if (pos < 0 || pos >= (int64_t)count) return std::nullopt;
        return Match{text + pairs[2 * pos], text + pairs[2 * pos + 1]};
    }
    RegexIterator* create_copy() { return new LazyIterator(*this); }
    void print() { print_program(prog); }
};

class Factory : public IterFactoryBase {
    const rrx_program* prog;
public:
    explicit Factory(const rrx_program* p) : prog(p) {}
    IteratorWrapper get_acceptance_iter(char* c) { return IteratorWrapper(new AcceptanceIterator(c, prog)); }
    IteratorWrapper get_lazy_iter(char* c) { return IteratorWrapper(new LazyIterator(c, prog)); }
    IteratorWrapper get_nonoverlap_iter(char* c) { return IteratorWrapper(new LazyIterator(c, prog, RRX_MODE_NO_OVERLAP)); }
};
}  // namespace detail

class RRegex {
    std::unique_ptr<rrx_program, void (*)(rrx_program*)> prog;
public:
    std::unique_ptr<IterFactoryBase> iter_factory;
    RRegex(const char* p) : prog(nullptr, rrx_program_free) {
        rrx_program* h = nullptr;
        if (rrx_compile(p, &h) != RRX_OK) throw std::runtime_error(rrx_last_error());
        prog.reset(h);
        iter_factory = std::make_unique<detail::Factory>(h);
    }
    IteratorWrapper get_acceptance_iter(char* c) { return iter_factory->get_acceptance_iter(c); }
    IteratorWrapper get_lazy_iter(char* c) { return iter_factory->get_lazy_iter(c); }
    IteratorWrapper get_nonoverlap_iter(char* c) { return iter_factory->get_nonoverlap_iter(c); }
};

}  // namespace Regex
