// Digest of a device match list: what the parity checks at full corpus size compare (SURVEY.md §8d
// "Parity check at scale": count + xor/sum-of-splitmix hash over all (start,end) pairs) — computed where the
// list lives, so that a 4 GB list never has to travel to the host.
//
//   digest[0]  number of matches with end <= end_limit
//   digest[1]  sum  (mod 2^64) of h(start,end) over those matches      h = sm64(sm64(start) + end)
//   digest[2]  xor              of h(start,end) over those matches
//   digest[3]  order violations: matches i>0 (of the whole list) with end[i] <= end[i-1] or start >= end
//
// Sum and xor do not depend on the order, so the digests of the shards of a sharded scan add up / xor up
// to the digest of the whole list; the order is checked separately (digest[3] must be 0).  The same digest
// of a host list is oracle/api.py::digest (numpy); tests/test_gpu_parity.py compares the two.
#include "rrx_scan.h"

namespace rrx {
namespace {

__host__ __device__ __forceinline__ unsigned long long sm64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) digest_kernel(const longlong2* __restrict__ m, unsigned long long n, long long end_limit,
                                                    unsigned long long* out) {
    unsigned long long cnt = 0, sum = 0, x = 0, bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const longlong2 v = m[i];
        if (v.x >= v.y) bad++;
        if (i > 0 && m[i - 1].y >= v.y) bad++;
        if (v.y <= end_limit) {
            const unsigned long long h = sm64(sm64((unsigned long long)v.x) + (unsigned long long)v.y);
            cnt++; sum += h; x ^= h;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        x ^= __shfl_xor_sync(0xffffffffu, x, d);
        bad += __shfl_xor_sync(0xffffffffu, bad, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(out + 0, cnt);
        if (sum) atomicAdd(out + 1, sum);
        if (x) atomicXor(out + 2, x);
        if (bad) atomicAdd(out + 3, bad);
    }
}

}  // namespace

cudaError_t launch_digest(const long long* d_pairs, uint64_t n, int64_t end_limit, unsigned long long* d_out4, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(d_out4, 0, 32, st);
    if (e != cudaSuccess || n == 0) return e;
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    digest_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const longlong2*>(d_pairs), n, end_limit, d_out4);
    return cudaGetLastError();
}

}  // namespace rrx
