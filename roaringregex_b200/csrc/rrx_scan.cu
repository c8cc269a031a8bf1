// rrx scan kernel (sm_100a) and its launcher.
//
// One persistent CTA per SM slot walks tiles in ticket order.  Per tile:
//   load   128-bit coalesced loads -> shared memory, rows of 128 B, XOR-swizzled 16-byte chunks
//   A      every thread scans its row from the empty set (register state word(s)), finds match ends
//   S      row transfer functions composed; tile aggregate published; decoupled look-back over
//          predecessor tiles yields the tile's exact entry set
//   B      rows replay their prefix from the exact entry (dies within a few bytes unless the
//          automaton has an absorbing loop)
//   C      match counts -> second look-back (output base) -> lazy start scans -> (start,end) writes
// The phase bodies live in rrx_tile.cuh (shared with the host-side test emulator).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "rrx_scan.h"
#include "rrx_tile.cuh"

namespace rrx {

// warps per CTA by state width: registers per thread grow with NW
template <int NW, int NT> struct KCfg {
    static constexpr int WARPS = (NW == 1 && NT >= 0) ? 32 : NW <= 2 ? 16 : NW <= 8 ? 8 : 4;
    static constexpr int THREADS = WARPS * 32;
    static constexpr uint32_t BT_BYTES = (NW == 1 ? 256u * 128u : 256u * (uint32_t)NW * 4u);
    static constexpr size_t REGION = (TileState<NW>::bytes() + 127) & ~(size_t)127;
};

template <int NW, int NT, int NB>
__global__ void __launch_bounds__(KCfg<NW, NT>::THREADS, 1) rrx_scan_kernel(const __grid_constant__ ScanParams sp) {
    using K = KCfg<NW, NT>;
    using TS = TileState<NW>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // shared memory: [byte table | per-warp tile regions | small program part]
    uint32_t* Btab = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* regions = smem_raw + K::BT_BYTES;
    uint32_t* progsm = reinterpret_cast<uint32_t*>(regions + (size_t)K::WARPS * K::REGION);

    const ProgHeader* gh = reinterpret_cast<const ProgHeader*>(sp.prog);
    const uint32_t n_small = gh->off_B, off_B = gh->off_B;
    for (uint32_t i = tid; i < n_small; i += K::THREADS) progsm[i] = sp.prog[i];
    if (NW == 1) {
        // one private copy of the table per lane: entry (c, lane) lives in bank `lane`
        for (uint32_t i = tid; i < K::BT_BYTES / 4; i += K::THREADS) Btab[i] = sp.prog[off_B + (i >> 5)];
    } else {
        for (uint32_t i = tid; i < K::BT_BYTES / 4; i += K::THREADS) Btab[i] = sp.prog[off_B + i];
    }
    __syncthreads();                      // the only block-level barrier of the kernel
    ProgView pv;
    pv.sm = progsm;
    pv.glob = sp.prog;
    pv.Btab = reinterpret_cast<const uint8_t*>(Btab);
    pv.lane_off = (uint32_t)lane * 4u;
    Prog<NW, NT, NB> fw;
    fw.init(pv);
    TS ts;
    ts.carve(regions + (size_t)warp * K::REGION);
    const int64_t len = (int64_t)sp.len;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(sp.ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= sp.num_tiles) break;
        ts.row0 = (uint64_t)tile * TILE_ROWS;

        // ---- stage the tile: 8 x 128-bit loads per lane, fully coalesced, XOR-swizzled store ----
        {
            const int64_t g0 = ts.lo();
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int q = k * 32 + lane;              // 16-byte chunk index inside the tile
                const int s = q >> 3, j = q & 7;
                const int64_t g = g0 + (int64_t)q * 16;
                uint8_t* dst = ts.text + (s << ROW_SHIFT) + ((j ^ (s & 7)) << 4);
                if (g + 16 <= len) {
                    *reinterpret_cast<uint4*>(dst) = __ldcs(reinterpret_cast<const uint4*>(sp.text + g));
                } else if (g < len) {
                    for (int b = 0; b < 16 && g + b < len; b++) dst[b] = sp.text[g + b];
                }
            }
        }
        __syncwarp();

        phase_a<NW>(sp, pv, fw, ts, lane);
        const uint32_t my_kind = ts.sum_kind[lane];
        const bool any_open = __any_sync(0xffffffffu, my_kind == K_RANK1 || my_kind == K_HARD) != 0;
        __syncwarp();
        if (lane == 0) phase_s<NW>(sp, pv, fw, ts, tile, any_open);
        __syncwarp();
        const uint32_t c = phase_b<NW>(sp, pv, fw, ts, lane, any_open);
        // exclusive offsets of the rows' matches inside the tile
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t offset = inc - c;
        unsigned long long base = 0;
        if (lane == 0) base = phase_c0(sp, tile, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (total) {
            uint32_t hit[4];
#pragma unroll
            for (int q = 0; q < 4; q++) hit[q] = ts.hits[q * TILE_ROWS + lane];
            __syncwarp();                 // every lane is done with the row summaries: reuse as hit list
            if (total <= (uint32_t)TS::LIST_CAP) {
                phase_c1_list<NW>(ts, lane, offset, hit);
                __syncwarp();
                phase_c2_balanced<NW>(sp, pv, fw, ts, lane, total, base);
            } else {
                phase_c2_rows<NW>(sp, pv, fw, ts, lane, offset, hit, base);
            }
        }
        __syncwarp();
    }
}

template <int NW, int NT> static size_t smem_bytes_for() {
    using K = KCfg<NW, NT>;
    // small program part: header + 3 sets + 2 decompositions
    const size_t small = sizeof(ProgHeader) + 4u * (3 * NW + 2 * (RRX_MAX_TERMS * (1 + NW) + RRX_MAX_BICL * 2 * NW + NW));
    return K::BT_BYTES + (size_t)K::WARPS * K::REGION + small + 64;
}

template <int NW, int NT, int NB>
static cudaError_t launch_one(const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out) {
    using K = KCfg<NW, NT>;
    auto kern = rrx_scan_kernel<NW, NT, NB>;
    const size_t smem = smem_bytes_for<NW, NT>();
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
