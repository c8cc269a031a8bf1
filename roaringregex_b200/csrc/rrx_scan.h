// Internal interface between the C ABI (rrx_capi.cu) and the kernels (rrx_scan.cu, rrx_corpus.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rrx_compile.h"

namespace rrx {
struct ScanParams;
cudaError_t launch_scan(const ScanParams& sp, const ProgHeader& h, int sm_count, cudaStream_t st, int* grid_out);
cudaError_t launch_order(const uint32_t* tile_cnt, const unsigned long long* tile_seg, uint32_t num_tiles,
                         unsigned long long* chunk_scratch, const long long* staging, uint64_t staging_cap, long long* out, uint64_t cap,
                         unsigned long long* n_total, cudaStream_t st);
cudaError_t launch_continue_starts(uint32_t NW, const uint32_t* d_prog, const uint8_t* d_text, uint64_t len, uint64_t left_ctx,
                                   uint64_t base, void* d_recs, uint64_t n, long long* d_start_out, cudaStream_t st);
cudaError_t launch_accepts(uint32_t NW, const uint32_t* d_prog, const uint8_t* d_text, uint64_t len, int* d_out, cudaStream_t st);
size_t no_overlap_workspace_bytes(uint64_t n);
cudaError_t launch_no_overlap(const longlong2* in, uint64_t n, longlong2* out, unsigned long long* d_n_out, void* workspace,
                              size_t workspace_bytes, cudaStream_t st);
cudaError_t filter_start_ge(const longlong2* in, uint64_t n, long long carry, longlong2** out, uint64_t* n_out, cudaStream_t st);
void fold_entry_host(uint32_t NW, const uint32_t* summaries, uint32_t n, uint32_t rank, uint32_t* entry_out, int* exact);
void apply_summary_host(uint32_t NW, const uint32_t* s, const uint32_t* entry, uint32_t* out);
cudaError_t launch_digest(const long long* d_pairs, uint64_t n, int64_t end_limit, unsigned long long* d_out4, cudaStream_t st);
cudaError_t launch_corpus_iid(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, const uint8_t* alpha256, cudaStream_t st);
cudaError_t launch_corpus_lines(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, cudaStream_t st);
cudaError_t launch_corpus_records(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, cudaStream_t st);
}  // namespace rrx
