// Explicit instantiations of the pre-compiled rrx scan kernels.  Compiled once per RRX_INST_GROUP (0..7) into
// separate objects so that the build runs in parallel (roaringregex_b200/_build.py); see rrx_scan.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <mutex>

#include "rrx_kernel.cuh"
#include "rrx_scan.h"
#include "rrx_scan_inst.h"

namespace rrx {

// warps per CTA by state width: registers per thread grow with NW
template <int NW, int NT> struct Warps { static constexpr int N = (NW == 1 && NT >= 0) ? 32 : NW <= 2 ? 16 : NW <= 8 ? 8 : 4; };

template <int NW, int NT, int NB>
__global__ void __launch_bounds__(Warps<NW, NT>::N * 32, 1) rrx_scan_kernel(const __grid_constant__ ScanParams sp) {
    rrx_scan_body<NW, NW, Prog<NW, NT, NB>, Warps<NW, NT>::N>(sp);
}

template <int NW, int NT, int NB>
cudaError_t launch_one(const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out) {
    using K = KCfg<NW, NW, Warps<NW, NT>::N>;
    auto kern = rrx_scan_kernel<NW, NT, NB>;
    const size_t smem = K::SMEM;
    // per instantiation AND per device: the shared-memory opt-in and the occupancy belong to the device
    constexpr int MAX_DEV = 64;
    static std::mutex mu;
    static int occ_of[MAX_DEV] = {0};
    int dev = 0, occ = 1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= MAX_DEV || occ_of[dev] == 0) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::THREADS, smem);
            if (e != cudaSuccess) return e;
            if (occ < 1) occ = 1;
            if (dev >= 0 && dev < MAX_DEV) occ_of[dev] = occ;
        } else {
            occ = occ_of[dev];
        }
    }
    long long grid = (long long)sm_count * occ;
    const long long need = ((long long)sp.num_tiles + K::WARPS - 1) / K::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    kern<<<(unsigned)grid, K::THREADS, smem, st>>>(sp);
    return cudaGetLastError();
}


#define RRX_INST(NW_, NT_, NB_) template cudaError_t launch_one<NW_, NT_, NB_>(const ScanParams&, int, cudaStream_t, int*);
#ifndef RRX_INST_GROUP
#error "compile with -DRRX_INST_GROUP=0..7"
#endif
#if RRX_INST_GROUP == 0
RRX_INST(1, 1, 0) RRX_INST(1, 2, 0) RRX_INST(1, 3, 0)
#elif RRX_INST_GROUP == 1
RRX_INST(1, 4, 0) RRX_INST(1, 1, 1) RRX_INST(1, 2, 1)
#elif RRX_INST_GROUP == 2
RRX_INST(1, 3, 1) RRX_INST(1, 4, 1)
#elif RRX_INST_GROUP == 3
RRX_INST(1, 2, 2) RRX_INST(1, 4, 2)
#elif RRX_INST_GROUP == 4
RRX_INST(1, -1, 0) RRX_INST(2, -1, 0)
#elif RRX_INST_GROUP == 5
RRX_INST(4, -1, 0) RRX_INST(8, -1, 0)
#elif RRX_INST_GROUP == 6
RRX_INST(16, -1, 0)
#elif RRX_INST_GROUP == 7
RRX_INST(32, -1, 0)
#endif

}  // namespace rrx
