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

template <int NW> struct KCfg { static constexpr int ROWS = NW <= 8 ? 256 : 128; };

template <int NW, int NT, int NB>
__global__ void __launch_bounds__(KCfg<NW>::ROWS) rrx_scan_kernel(const __grid_constant__ ScanParams sp) {
    constexpr int ROWS = KCfg<NW>::ROWS;
    using TS = TileState<NW, ROWS>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;

    // shared memory: [tile state | byte table | small program part]
    TS ts;
    ts.carve(smem_raw);
    uint32_t* Btab = reinterpret_cast<uint32_t*>(smem_raw + ((TS::bytes() + 15) & ~(size_t)15));
    constexpr uint32_t BT_WORDS = NW == 1 ? 256u * 32u : 256u * (uint32_t)NW;
    uint32_t* progsm = Btab + BT_WORDS;

    const ProgHeader* gh = reinterpret_cast<const ProgHeader*>(sp.prog);
    const uint32_t n_small = gh->off_B, off_B = gh->off_B;
    for (uint32_t i = tid; i < n_small; i += ROWS) progsm[i] = sp.prog[i];
    if (NW == 1) {
        // one private copy of the table per lane: entry (c, lane) lives in bank `lane`
        for (uint32_t i = tid; i < BT_WORDS; i += ROWS) Btab[i] = sp.prog[off_B + (i >> 5)];
    } else {
        for (uint32_t i = tid; i < BT_WORDS; i += ROWS) Btab[i] = sp.prog[off_B + i];
    }
    __syncthreads();
    ProgView pv;
    pv.sm = progsm;
    pv.glob = sp.prog;
    pv.Btab = Btab;
    pv.btab_stride = NW == 1 ? 32u : (uint32_t)NW;
    pv.btab_lane = NW == 1 ? (uint32_t)(tid & 31) : 0u;
    Fwd<NW, NT, NB> fw;
    fw.init(pv);

    for (;;) {
        if (tid == 0) ts.misc[TS::M_TILE] = atomicAdd(sp.ticket, 1u);
        __syncthreads();
        const uint32_t tile = ts.misc[TS::M_TILE];
        if (tile >= sp.num_tiles) break;
        ts.row0 = (uint64_t)tile * ROWS;
        const bool have_halo = tile > 0 || sp.left_ctx >= ROW_BYTES;
        ts.smem_lo = have_halo ? ((int64_t)ts.row0 - 1) * ROW_BYTES : 0;

        // ---- stage the tile (+ halo row) ------------------------------------------------------
        {
            const int64_t g0 = ((int64_t)ts.row0 - 1) * ROW_BYTES;
            const int64_t len = (int64_t)sp.len;
            for (int q = tid; q < (ROWS + 1) * 8; q += ROWS) {
                const int s = q >> 3, j = q & 7;
                const int64_t g = g0 + (int64_t)s * ROW_BYTES + j * 16;
                uint8_t* dst = ts.text + s * ROW_BYTES + ((j ^ (s & 7)) << 4);
                if (g < ts.smem_lo || g >= len) continue;
                if (g + 16 <= len) {
                    *reinterpret_cast<uint4*>(dst) = __ldcs(reinterpret_cast<const uint4*>(sp.text + g));
                } else {
                    for (int k = 0; k < 16 && g + k < len; k++) dst[k] = sp.text[g + k];
                }
            }
        }
        __syncthreads();

        phase_a<NW, ROWS>(sp, pv, fw, ts, tid);
        const uint32_t my_kind = ts.sum_kind[tid];
        const bool any_open = __syncthreads_or(my_kind == K_RANK1 || my_kind == K_HARD) != 0;
        if (tid == 0) phase_s<NW, ROWS>(sp, pv, fw, ts, tile, any_open);
        __syncthreads();
        phase_b<NW, ROWS>(sp, pv, fw, ts, tid, any_open);
        __syncthreads();
        if (tid == 0) phase_c0<NW, ROWS>(sp, ts, tile);
        __syncthreads();
        phase_c<NW, ROWS>(sp, pv, ts, tid);
        __syncthreads();
    }
}

template <int NW> static size_t smem_bytes_for() {
    using TS = TileState<NW, KCfg<NW>::ROWS>;
    const size_t bt = (NW == 1 ? 256u * 32u : 256u * (uint32_t)NW) * 4u;
    // small program part: header + 3 sets + 2 decompositions
    const size_t small = sizeof(ProgHeader) + 4u * (3 * NW + 2 * (RRX_MAX_TERMS * (1 + NW) + RRX_MAX_BICL * 2 * NW + NW));
    return ((TS::bytes() + 15) & ~(size_t)15) + bt + small + 64;
}

template <int NW, int NT, int NB>
static cudaError_t launch_one(const ScanParams& sp, int sm_count, cudaStream_t st, int* grid_out) {
    auto kern = rrx_scan_kernel<NW, NT, NB>;
    const size_t smem = smem_bytes_for<NW>();
    static bool configured = false;     // per instantiation
    static int occ = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, KCfg<NW>::ROWS, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        configured = true;
    }
    long long grid = (long long)sm_count * occ;
    if (grid > (long long)sp.num_tiles) grid = sp.num_tiles;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    kern<<<(unsigned)grid, KCfg<NW>::ROWS, smem, st>>>(sp);
    return cudaGetLastError();
}

int rows_per_tile(uint32_t NW) { return NW <= 8 ? 256 : 128; }

cudaError_t launch_scan(const ScanParams& sp, const ProgHeader& h, int sm_count, cudaStream_t st, int* grid_out) {
    const uint32_t nt = h.n_terms, nb = h.n_bicl;
    if (h.NW == 1) {
        bool rot_only = !h.use_rows;
        if (rot_only) {
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
