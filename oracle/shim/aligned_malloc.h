/* TEST INFRASTRUCTURE ONLY — force-included (-include) when building oracle/_ref.
 * The reference stores BitSet<4> rows with aligned AVX2 stores (BitSet.cc:18,32,54,177) into a
 * table obtained from plain malloc (regex.h:27), which is only 16-byte aligned on glibc, so its
 * 129..256-state path segfaults as shipped.  Handing it 64-byte-aligned storage is a build
 * recipe matter; the reference sources stay untouched. */
#pragma once
#include <stdlib.h>
#ifdef __cplusplus
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include <iostream>
#include <stack>
#include <optional>
#endif
#include <immintrin.h>
#define malloc(x) aligned_alloc(64, ((((size_t)(x)) + 63) / 64) * 64)
