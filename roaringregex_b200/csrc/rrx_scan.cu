// Pre-compiled instantiations of the rrx scan kernel (sm_100a) and their launcher.
// The kernel body is in rrx_kernel.cuh; the run-time specialised variant in rrx_jit.cc.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "rrx_kernel.cuh"
#include "rrx_scan.h"

namespace rrx {

// warps per CTA by state width: registers per thread grow with NW
template <int NW, int NT> struct Warps { static constexpr int N = (NW == 1 && NT >= 0) ? 32 : NW <= 2 ? 16 : NW <= 8 ? 8 : 4; };

template <int NW, int NT, int NB>
__global__ void __launch_bounds__(Warps<NW, NT>::N * 32, 1) rrx_scan_kernel(const __grid_constant__ ScanParams sp) {
    rrx_scan_body<NW, Prog<NW, NT, NB>, Warps<NW, NT>::N>(sp);
}

template <int NW, int NT, int NB>
static cudaError_t launch_one(const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out) {
    using K = KCfg<NW, Warps<NW, NT>::N>;
    auto kern = rrx_scan_kernel<NW, NT, NB>;
    const size_t smem = K::SMEM;
    static bool configured = false;     // per instantiation
    static int occ = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::THREADS, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        configured = true;
    }
    long long grid = (long long)sm_count * occ;
    const long long need = ((long long)sp.num_tiles + K::WARPS - 1) / K::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    kern<<<(unsigned)grid, K::THREADS, smem, st>>>(sp);
    return cudaGetLastError();
}

cudaError_t launch_scan(const ScanParams& sp, const ProgHeader& h, int sm_count, cudaStream_t st, int* grid_out) {
    if (h.NW == 1) {
        const uint32_t nt = h.n_terms > h.n_rterms ? h.n_terms : h.n_rterms;
        const uint32_t nb = h.n_bicl > h.n_rbicl ? h.n_bicl : h.n_rbicl;
        if (!h.use_rows && !h.use_rrows) {
#define RRX_TRY(NT_, NB_) if (nt <= NT_ && nb <= NB_) return launch_one<1, NT_, NB_>(sp, sm_count, st, grid_out);
            RRX_TRY(1, 0) RRX_TRY(2, 0) RRX_TRY(3, 0) RRX_TRY(4, 0) RRX_TRY(1, 1) RRX_TRY(2, 1) RRX_TRY(3, 1) RRX_TRY(4, 1)
            RRX_TRY(2, 2) RRX_TRY(4, 2)
#undef RRX_TRY
        }
        return launch_one<1, -1, 0>(sp, sm_count, st, grid_out);
    }
    switch (h.NW) {
        case 2: return launch_one<2, -1, 0>(sp, sm_count, st, grid_out);
        case 4: return launch_one<4, -1, 0>(sp, sm_count, st, grid_out);
        case 8: return launch_one<8, -1, 0>(sp, sm_count, st, grid_out);
        case 16: return launch_one<16, -1, 0>(sp, sm_count, st, grid_out);
        case 32: return launch_one<32, -1, 0>(sp, sm_count, st, grid_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace rrx
