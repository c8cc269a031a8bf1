// Ordering pass: the scan kernel leaves every tile's matches in a contiguous segment of the
// staging list, segments in tile-COMPLETION order.  Three small launches put them in tile order
// (= ascending end offset, the order the reference's iterator would produce them in):
//   chunk_sums   per chunk of 1024 tiles: number of matches
//   chunk_scan   exclusive scan of the chunk sums (one CTA)
//   gather       per chunk: block scan of the tile counts, then one warp per tile copies its segment
// Cost: one extra read + write of the match list (16 B per match each way), usually served from L2.
#include "rrx_scan.h"
#include "rrx_tile.cuh"

namespace rrx {
namespace {

constexpr int CHUNK = 1024;

__device__ __forceinline__ unsigned long long block_sum_1024(unsigned long long v, unsigned long long* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    unsigned long long t = threadIdx.x < 32 ? sh[threadIdx.x] : 0ull;
    if (warp == 0) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// exclusive scan over the 1024 threads of the CTA; also returns the CTA total in *total
__device__ __forceinline__ unsigned long long block_excl_1024(unsigned long long v, unsigned long long* sh, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = sh[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += u;
        }
        sh[lane] = winc - w;
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    const unsigned long long res = sh[warp] + inc - v;
    if (total) *total = sh[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(CHUNK) chunk_sums_kernel(const uint32_t* tile_cnt, uint32_t num_tiles, unsigned long long* chunk_sum) {
    __shared__ unsigned long long sh[33];
    const uint32_t t = blockIdx.x * CHUNK + threadIdx.x;
    const unsigned long long v = t < num_tiles ? tile_cnt[t] : 0u;
    const unsigned long long s = block_sum_1024(v, sh);
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(CHUNK) chunk_scan_kernel(unsigned long long* chunk_sum, uint32_t nchunks, unsigned long long* n_total) {
    __shared__ unsigned long long sh[33];
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < nchunks; base += CHUNK) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < nchunks ? chunk_sum[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_excl_1024(v, sh, &tot);
        if (i < nchunks) chunk_sum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && n_total) *n_total = carry;       // number of matches of the scan
}

__global__ void __launch_bounds__(CHUNK) gather_kernel(const uint32_t* tile_cnt, const unsigned long long* tile_seg, unsigned long long* tile_seg_out, uint32_t num_tiles,
                                                       const unsigned long long* chunk_off, const longlong2* staging,
                                                       unsigned long long staging_cap, longlong2* out, unsigned long long cap) {
    __shared__ unsigned long long sh[33];
    __shared__ unsigned long long off[CHUNK];
    const uint32_t t = blockIdx.x * CHUNK + threadIdx.x;
    const unsigned long long v = t < num_tiles ? tile_cnt[t] : 0u;
    off[threadIdx.x] = chunk_off[blockIdx.x] + block_excl_1024(v, sh, nullptr);
    __syncthreads();
    // Every warp moves the segments of its 32 consecutive tiles.  The copy is flattened over the
    // matches of those tiles — element e of the warp's output range belongs to the tile j with
    // pre[j] <= e < pre[j+1], found by a 5-step search over the lanes' prefix sums — so that all 32
    // lanes copy in every iteration, the stores are contiguous and the loads independent.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tt = blockIdx.x * CHUNK + warp * 32 + lane;
    const uint32_t n = tt < num_tiles ? tile_cnt[tt] : 0u;
    const unsigned long long src = tt < num_tiles ? tile_seg[tt] : 0ull;
    uint32_t inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    const uint32_t pre = inc - n, E = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned long long dst0 = off[warp * 32];
    for (uint32_t e0 = 0; e0 < E; e0 += 32) {
        const uint32_t e = e0 + lane;
        int j = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const uint32_t pc = __shfl_sync(0xffffffffu, pre, j + s);
            if (pc <= e) j += s;
        }
        const unsigned long long sj = __shfl_sync(0xffffffffu, src, j);
        const uint32_t pj = __shfl_sync(0xffffffffu, pre, j);
        const unsigned long long from = sj + (e - pj), to = dst0 + e;
        if (e < E && from < staging_cap && to < cap) out[to] = staging[from];
    }
    // leave the tile's offset in the ORDERED list behind (pending start scans refer to it)
    if (tt < num_tiles) tile_seg_out[tt] = off[threadIdx.x];
}

}  // namespace

cudaError_t launch_order(const uint32_t* tile_cnt, const unsigned long long* tile_seg, uint32_t num_tiles,
                         unsigned long long* chunk_scratch, const long long* staging, uint64_t staging_cap, long long* out, uint64_t cap,
                         unsigned long long* n_total, cudaStream_t st) {
    if (!num_tiles) return cudaSuccess;
    const uint32_t nchunks = (num_tiles + CHUNK - 1) / CHUNK;
    chunk_sums_kernel<<<nchunks, CHUNK, 0, st>>>(tile_cnt, num_tiles, chunk_scratch);
    chunk_scan_kernel<<<1, CHUNK, 0, st>>>(chunk_scratch, nchunks, n_total);
    gather_kernel<<<nchunks, CHUNK, 0, st>>>(tile_cnt, tile_seg, const_cast<unsigned long long*>(tile_seg), num_tiles, chunk_scratch,
                                             reinterpret_cast<const longlong2*>(staging), staging_cap, reinterpret_cast<longlong2*>(out), cap);
    return cudaGetLastError();
}

}  // namespace rrx
