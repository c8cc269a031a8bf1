// Run-time specialisation of the scan kernel: the regex becomes machine code.
// For a one-word automaton the host generates a `ProgJit` policy whose follow()/rfollow() are
// straight-line code with the automaton's masks, shifts and multipliers as immediates, compiles
// rrx_kernel.cuh with it through NVRTC for sm_100a, and launches the result through the driver API.
// The pre-compiled generic kernels (rrx_scan.cu) remain the path for wider automata, for short
// inputs (where a ~1 s compile does not pay) and when NVRTC cannot be loaded.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "rrx_compile.h"

namespace rrx {
struct ScanParams;
struct JitKernel;   // opaque: a loaded module + function

// Generated CUDA source of the ProgJit policy (exposed for tests / inspection).
std::string jit_policy_source(const Program& prog);
// True if this program can be specialised (one-word sets, no row tables).
bool jit_applicable(const Program& prog);
// NVRTC only (needs no GPU): the sm_100a cubin of the specialised kernel.
bool jit_compile_cubin(const Program& prog, std::vector<char>* cubin, std::string* err, int variant = 0);
// Compile (or fetch from the in-process cache).  Returns nullptr and fills *err on failure.
// variant (thread-owned programs): 0 = lean kernel (a run alive at its row's end is handed to the next row through the
// row's exit), 1 = straddle kernel (the run's lane follows it into the next row: tiles keep their fast tail when matches
// straddle rows; ~5 % more code, which costs the lean case its instruction-cache fit).  See rrx_capi.cu: a scanner
// starts lean and switches by what the previous scan saw.
JitKernel* jit_get(const Program& prog, std::string* err, int variant = 0);
cudaError_t jit_launch(JitKernel* k, const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out);
}  // namespace rrx
