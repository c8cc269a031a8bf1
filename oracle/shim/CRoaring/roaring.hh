// TEST INFRASTRUCTURE ONLY — stand-in for the un-vendored CRoaring dependency.
//
// The reference clones CRoaring `master` (unpinned) at build time and amalgamates it
// (/root/reference/Makefile:21-23); the result is included from src/inc/regex.h:8 and
// src/Parser.cpp:8.  There is no network here, so this header supplies a small set type with
// exactly the member functions the reference calls (call sites listed in SURVEY.md §8c).  It
// exists so that the reference's own sources compile UNMODIFIED into oracle/_ref/; it is only
// *executed* for regexes with more than 256 raw states, a path that is defective in the
// reference anyway and that no test relies on.
//
// Requirements honoured: all-zero bytes are a valid empty set (the reference memsets its table
// and never constructs the objects, regex.h:29); trivially destructible storage.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace roaring {

class Roaring {
public:
    static constexpr uint32_t kBits = 4096;
    static constexpr uint32_t kWords = kBits / 64;
    uint64_t w[kWords];

    Roaring() { std::memset(w, 0, sizeof(w)); }
    Roaring(const Roaring&) = default;
    Roaring(Roaring&&) = default;
    Roaring& operator=(const Roaring&) = default;
    Roaring& operator=(Roaring&&) = default;
    ~Roaring() {}

    static Roaring bitmapOf(size_t n, ...) {
        Roaring r;
        va_list ap;
        va_start(ap, n);
        for (size_t i = 0; i < n; i++) r.add(va_arg(ap, uint32_t));
        va_end(ap);
        return r;
    }
    void add(uint32_t x) { if (x < kBits) w[x >> 6] |= 1ull << (x & 63); }
    bool contains(uint32_t x) const { return x < kBits && ((w[x >> 6] >> (x & 63)) & 1ull); }
    uint64_t cardinality() const {
        uint64_t c = 0;
        for (uint32_t i = 0; i < kWords; i++) c += (uint64_t)__builtin_popcountll(w[i]);
        return c;
    }
    uint64_t and_cardinality(const Roaring& o) const {
        uint64_t c = 0;
        for (uint32_t i = 0; i < kWords; i++) c += (uint64_t)__builtin_popcountll(w[i] & o.w[i]);
        return c;
    }
    void toUint32Array(uint32_t* out) const {
        for (uint32_t i = 0; i < kWords; i++) {
            uint64_t v = w[i];
            while (v) { *out++ = i * 64 + (uint32_t)__builtin_ctzll(v); v &= v - 1; }
        }
    }
    static Roaring fastunion(size_t n, const Roaring** in) {
        Roaring r;
        for (size_t k = 0; k < n; k++)
            for (uint32_t i = 0; i < kWords; i++) r.w[i] |= in[k]->w[i];
        return r;
    }
    Roaring& operator|=(const Roaring& o) {
        for (uint32_t i = 0; i < kWords; i++) w[i] |= o.w[i];
        return *this;
    }
    void printf() const {
        std::printf("{");
        bool first = true;
        for (uint32_t x = 0; x < kBits; x++)
            if (contains(x)) { std::printf(first ? "%u" : ",%u", x); first = false; }
        std::printf("}");
    }

    class const_iterator {
        const Roaring* s;
        uint32_t pos;
        void seek() { while (pos < kBits && !s->contains(pos)) pos++; }
    public:
        const_iterator(const Roaring* set, uint32_t p) : s(set), pos(p) { seek(); }
        uint32_t operator*() const { return pos; }
        const_iterator& operator++() { pos++; seek(); return *this; }
        const_iterator operator++(int) { const_iterator t = *this; ++*this; return t; }
        bool operator!=(const const_iterator& o) const { return pos != o.pos; }
        bool operator==(const const_iterator& o) const { return pos == o.pos; }
    };
    const_iterator begin() const { return const_iterator(this, 0); }
    const_iterator end() const { return const_iterator(this, kBits); }
};

}  // namespace roaring
