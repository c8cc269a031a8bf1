// launch_one<NW, NT, NB>: configure (once per device) and launch one pre-compiled instantiation of the scan
// kernel.  Defined and explicitly instantiated in rrx_scan_inst.cu (one object per RRX_INST_GROUP).
#pragma once
#include <cuda_runtime.h>

namespace rrx {
struct ScanParams;
template <int NW, int NT, int NB>
cudaError_t launch_one(const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out);
}  // namespace rrx
