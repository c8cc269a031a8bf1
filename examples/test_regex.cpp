// A small command-line front end for the drop-in header: reads a text line and a pattern line from stdin (the same
// input the reference's interactive driver takes, RoaringRegex src/test/main.cpp — that file itself compiles
// unchanged against include/compat/regex.h, see tests/test_cpp_dropin.py) and prints the automaton summary, the
// whole-string answer, every lazy match, and the non-overlapping view.  Written for this repository.
//   g++ -std=c++17 -I include examples/test_regex.cpp -L roaringregex_b200 -lrrx -Wl,-rpath,$PWD/roaringregex_b200 -o test_regex
#include <chrono>
#include <iostream>
#include <string>

#include "rrx_regex.hpp"

namespace {
void list_matches(Regex::IteratorWrapper it, const char* label, char* text) {
    for (it++; (*it).has_value(); it++) {
        Regex::Match m = **it;
        std::cout << label << " [" << (m.start - text) << "," << (m.end - text) << "): " << m.str() << std::endl;
    }
}
}  // namespace

int main() {
    std::string text, pattern;
    if (!std::getline(std::cin, text) || !std::getline(std::cin, pattern)) return -1;
    try {
        Regex::RRegex regex(pattern.c_str());
        char* t = text.data();
        const auto t0 = std::chrono::steady_clock::now();
        Regex::IteratorWrapper acc = regex.get_acceptance_iter(t);
        acc++;
        const bool whole = (*acc).has_value();
        const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        acc.print();
        std::cout << "is match? " << whole << std::endl << "time: " << ms << "ms" << std::endl;
        list_matches(regex.get_lazy_iter(t), "match", t);
        list_matches(regex.get_nonoverlap_iter(t), "non-overlapping", t);
    } catch (const std::runtime_error& e) {
        std::cout << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
