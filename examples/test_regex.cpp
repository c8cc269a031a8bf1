// The reference's interactive driver (RoaringRegex src/test/main.cpp:13-35), against the drop-in
// header: line 1 = text, line 2 = pattern; prints the automaton summary, the whole-string answer
// with the reference's stopwatch placement, and — what its README promises — all lazy matches.
//   g++ -std=c++17 -I include examples/test_regex.cpp -L roaringregex_b200 -lrrx -Wl,-rpath,$PWD/roaringregex_b200 -o test_regex
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "rrx_regex.hpp"

using namespace Regex;
int main() {
    char* text = NULL;
    char* pattern = NULL;
    size_t len;
    if (getline(&text, &len, stdin) == -1 || getline(&pattern, &len, stdin) == -1) return -1;
    text[strlen(text) - 1] = '\0';
    pattern[strlen(pattern) - 1] = '\0';
    try {
        RRegex r(pattern);
        auto start_time = std::chrono::high_resolution_clock::now();
        auto acceptance_iter = r.get_acceptance_iter(text)++;
        bool is_match = (*acceptance_iter).has_value();
        auto end_time = std::chrono::high_resolution_clock::now();
        acceptance_iter.print();
        std::cout << "is match? " << is_match << std::endl;
        std::cout << "time: " << std::chrono::duration_cast<std::chrono::milliseconds>(end_time - start_time).count() << "ms\n";
        auto it = r.get_lazy_iter(text);
        for (it++; (*it).has_value(); it++) {
            Match m = **it;
            std::cout << "match [" << (m.start - text) << "," << (m.end - text) << "): " << m.str() << std::endl;
        }
    } catch (const std::runtime_error& e) {
        std::cout << "error: " << e.what() << std::endl;
        free(text);
        free(pattern);
        return 2;
    }
    free(text);
    free(pattern);
    return 0;
}
