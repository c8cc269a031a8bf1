// Non-overlapping view of a LAZY-v1 match list (the `finditer`-style reading
// of "all matches"):   p = 0;  for (s,e) in list:  if s >= p { emit; p = e }
// The greedy selection is a chain: after match i the next selected match is
//   nxt(i) = first j > i with start_j >= end_i.
// Ends ascend and start_k < end_k, so every start at or before i lies below end_i; hence
// nxt(i) = first j with PM[j] >= end_i, where PM is the running maximum of the starts — a
// binary search on a monotone array, whatever the overlap structure.  The chain from match 0 is then
// marked by pointer doubling (ceil(log2 M) rounds) and the marked matches are compacted in order.
// Everything runs on the device; cost O(M log M) on 16-byte records, small next to the scan.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "rrx_scan.h"

namespace rrx {
namespace {

__global__ void nov_starts_kernel(const longlong2* m, uint32_t n, long long* starts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) starts[i] = m[i].x;
}
__global__ void nov_next_kernel(const longlong2* m, const long long* pm, uint32_t n, uint32_t* nxt, uint8_t* mark) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long e = m[i].y;
    uint32_t lo = i + 1, hi = n;            // first j in (i, n) with pm[j] >= e
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (pm[mid] >= e) hi = mid; else lo = mid + 1;
    }
    nxt[i] = lo;
    mark[i] = i == 0 ? 1 : 0;               // p starts at the text's base offset: the first match is always taken
}
__global__ void nov_jump_kernel(const uint32_t* jump, uint32_t* jump2, uint8_t* mark, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = jump[i];
    if (j < n) {
        if (mark[i]) mark[j] = 1;
        jump2[i] = jump[j];
    } else {
        jump2[i] = n;
    }
}
struct MaxOp {
    __host__ __device__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};

}  // namespace

static size_t arrays_bytes(uint64_t n) { return (((size_t)n * 25) + 255) & ~(size_t)255; }   // starts 8 | pm 8 | jump 4 | jump2 4 | mark 1

size_t no_overlap_workspace_bytes(uint64_t n) {
    size_t t1 = 0, t2 = 0;
    cub::DeviceScan::InclusiveScan((void*)nullptr, t1, (const long long*)nullptr, (long long*)nullptr, MaxOp(), (int)n);
    cub::DeviceSelect::Flagged((void*)nullptr, t2, (const longlong2*)nullptr, (const uint8_t*)nullptr, (longlong2*)nullptr,
                               (unsigned long long*)nullptr, (int)n);
    return arrays_bytes(n) + (t1 > t2 ? t1 : t2) + 256;
}

// in: n matches sorted by end.  out: the selected ones (may not alias in); *d_n_out their number.
cudaError_t launch_no_overlap(const longlong2* in, uint64_t n, longlong2* out, unsigned long long* d_n_out, void* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
    if (n == 0) return cudaMemsetAsync(d_n_out, 0, 8, st);
    if (n >= 0x7fffffffull || workspace_bytes < no_overlap_workspace_bytes(n)) return cudaErrorInvalidValue;
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    long long* starts = reinterpret_cast<long long*>(w);
    long long* pm = starts + n;
    uint32_t* jump = reinterpret_cast<uint32_t*>(pm + n);
    uint32_t* jump2 = jump + n;
    uint8_t* mark = reinterpret_cast<uint8_t*>(jump2 + n);
    void* tmp = w + arrays_bytes(n);
    size_t tmp_bytes = workspace_bytes - arrays_bytes(n);
    const uint32_t N = (uint32_t)n, T = 256, G = (N + T - 1) / T;
    nov_starts_kernel<<<G, T, 0, st>>>(in, N, starts);
    cudaError_t e = cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, starts, pm, MaxOp(), (int)N, st);
    if (e != cudaSuccess) return e;
    nov_next_kernel<<<G, T, 0, st>>>(in, pm, N, jump, mark);
    for (uint32_t span = 1; span < N; span <<= 1) {
        nov_jump_kernel<<<G, T, 0, st>>>(jump, jump2, mark, N);
        uint32_t* t = jump; jump = jump2; jump2 = t;
    }
    e = cub::DeviceSelect::Flagged(tmp, tmp_bytes, in, mark, out, d_n_out, (int)N, st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

// Matches with start >= carry, in list order (the first step of the non-overlapping view of a chunk or shard that
// inherits the "last selected end" of the text before it).  Allocates *out (cudaFree); synchronises `st`.
struct StartGE {
    long long c;
    __host__ __device__ bool operator()(const longlong2& m) const { return m.x >= c; }
};
cudaError_t filter_start_ge(const longlong2* in, uint64_t n, long long carry, longlong2** out, uint64_t* n_out, cudaStream_t st) {
    *out = nullptr;
    *n_out = 0;
    if (n == 0) return cudaSuccess;
    if (n >= 0x7fffffffull) return cudaErrorInvalidValue;
    size_t tb = 0;
    unsigned long long* d_cnt = nullptr;
    void* tmp = nullptr;
    longlong2* d_out = nullptr;
    cub::DeviceSelect::If((void*)nullptr, tb, in, d_out, d_cnt, (int)n, StartGE{carry}, st);
    cudaError_t e = cudaMalloc(&tmp, tb + 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 8);
    if (e == cudaSuccess) e = cub::DeviceSelect::If(tmp, tb, in, d_out, d_cnt, (int)n, StartGE{carry}, st);
    unsigned long long cnt = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    cudaFree(d_cnt);
    if (e != cudaSuccess) { cudaFree(d_out); return e; }
    *out = d_out;
    *n_out = cnt;
    return cudaSuccess;
}

}  // namespace rrx
