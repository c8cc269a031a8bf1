// Launcher of the pre-compiled rrx scan kernels (sm_100a): picks the instantiation for a program.
// The kernel body is in rrx_kernel.cuh, the instantiations in rrx_scan_inst.cu, the run-time specialised
// variant in rrx_jit.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rrx_prog.h"
#include "rrx_scan.h"
#include "rrx_scan_inst.h"

namespace rrx {

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
