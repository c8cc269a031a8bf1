// TEST INFRASTRUCTURE ONLY — thin extern "C" driver around the UNMODIFIED reference library.
//
// Compiled by oracle/Makefile together with the reference's own sources from where they lie
// (/root/reference/src/{BitSet.cc,NFA.cc,Parser.cpp,Iterator.cpp}) into oracle/_ref/libref.so.
// It calls only the reference's public API, exactly as its test program does
// (/root/reference/src/test/main.cpp:22-31):  RRegex r(pattern); r.get_acceptance_iter(text)++;
// (*iter).has_value().  Nothing under roaringregex_b200/ links or loads this file.
#ifdef REF_UNITY
// Unity build: the reference's four sources, included from where they lie (-I /root/reference/src).
// The per-byte shift() lives in NFA.cc and the iterator loop in regex.h; without cross-file
// inlining (the reference asks for -flto, which this image's gcc cannot run) the scan is ~10x slower.
#include "BitSet.cc"
#include "NFA.cc"
#include "Parser.cpp"
#include "Iterator.cpp"
#endif
#include "regex.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

using namespace Regex;

extern "C" {

// Whole-string acceptance (the only iterator the reference ships, regex.h:150-165).
// Returns 1/0, or -1 if the parser threw (Parser.cpp:36,155).
int ref_accepts(const char* pattern, const char* text) {
    try {
        RRegex r(pattern);
        auto it = r.get_acceptance_iter(const_cast<char*>(text))++;
        return (*it).has_value() ? 1 : 0;
    } catch (const std::runtime_error&) {
        return -1;
    }
}

// Same scan with the reference's own stopwatch placement (main.cpp:25-31): compile outside,
// iterator construction + ++ + * inside.  `text` must be NUL-terminated.  Returns seconds.
double ref_time_acceptance(const char* pattern, const char* text, int repeats, int* accepted) {
    RRegex r(pattern);
    double best = 1e300;
    int acc = 0;
    for (int k = 0; k < repeats; k++) {
        auto t0 = std::chrono::high_resolution_clock::now();
        auto it = r.get_acceptance_iter(const_cast<char*>(text))++;
        acc = (*it).has_value() ? 1 : 0;
        auto t1 = std::chrono::high_resolution_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        if (s < best) best = s;
    }
    if (accepted) *accepted = acc;
    return best;
}

// Dump the compiled NFA with the reference's own NFA::print() (NFA.cc:14-41) to stdout.
int ref_dump(const char* pattern) {
    try {
        RRegex r(pattern);
        char empty[1] = {0};
        auto it = r.get_acceptance_iter(empty);
        it.print();
        std::fflush(stdout);
        return 0;
    } catch (const std::runtime_error&) {
        return -1;
    }
}

}  // extern "C"

#ifdef REF_TOOL_MAIN
// ref_tool dump <pattern>          -> NFA dump
// ref_tool accept <pattern> <text> -> prints 1/0
int main(int argc, char** argv) {
    if (argc >= 3 && !std::strcmp(argv[1], "dump")) return ref_dump(argv[2]) ? 2 : 0;
    if (argc >= 4 && !std::strcmp(argv[1], "accept")) {
        int a = ref_accepts(argv[2], argv[3]);
        std::printf("%d\n", a);
        return a < 0 ? 2 : 0;
    }
    std::fprintf(stderr, "usage: ref_tool dump <pattern> | accept <pattern> <text>\n");
    return 1;
}
#endif
