// Lets code written against RoaringRegex (`#include "regex.h"`, src/inc/regex.h) build against rrx:
//   g++ -std=c++17 -I include/compat <your file>.cpp -L roaringregex_b200 -lrrx
#pragma once
#include "../rrx_regex.hpp"
