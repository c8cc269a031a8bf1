// The rrx scan kernel body (sm_100a), shared by the pre-compiled instantiations (rrx_scan.cu) and
// the run-time specialised ones (rrx_jit.cu feeds this header to NVRTC together with a generated
// `ProgJit` policy that has the automaton's masks and shifts baked in as immediates).
//
// One persistent CTA per SM; every WARP walks tiles (32 rows of 128 B) in ticket order:
//   load   8 x 128-bit coalesced loads per lane -> shared memory rows (pitch RRX_ROW_PITCH: 144 bytes for the
//          thread-owned kernels — 16-byte aligned rows, 128-bit row loads —, 132 otherwise); or already there,
//          copied asynchronously (16-byte cp.async) during the previous tile's fast tail
//   A      thread-owned variant (program header `donly`: a single-byte cut whose dominated states are all that is
//          ever carried): candidate mask (+ stop-byte mask for skip-state walks), candidate-driven match ends with
//          verified lazy starts, row kinds decided lazily (lean kernel: eagerly, no run followed into the next row);
//          row-bounded variant (cut, but entry sets outside the dominated set): exit sets first, aggregate published
//          at once when all rows are closed, then the candidates; dense variant: every lane scans its row from the
//          empty set (register state word(s))
//   S      row transfer functions composed; tile aggregate published; two-level decoupled
//          look-back over predecessor tiles yields the tile's exact entry set
//   B      rows replay their prefix from the exact entry (dies within a few bytes unless the
//          automaton has an absorbing loop)
//   C      the tile's matches go to its fixed staging slots (or an overflow segment); balanced lazy start
//          scans for the matches whose start is not known yet; (start,end) writes.  Fast tail: B and C are
//          skipped when there is nothing to replay or resolve.  rrx_order.cu puts the segments in text order.
// The phase bodies live in rrx_tile.cuh (shared with the host-side test emulator).
#pragma once
#include "rrx_tile.cuh"

namespace rrx {

// shared-memory geometry of a CTA of WARPS warps
template <int NW, int CW, int WARPS_> struct KCfg {
    static constexpr int WARPS = WARPS_;
    static constexpr int THREADS = WARPS * 32;
    static constexpr uint32_t BT_BYTES = (NW == 1 ? 256u * (uint32_t)BT_COPIES * 4u : 256u * (uint32_t)NW * 4u);
    // per-warp region: the tile state, then 16 bytes for the warp's mbarrier (bulk-copy staging)
    static constexpr size_t REGION = (TileState<NW, CW>::bytes() + 16 + 15) & ~(size_t)15;
    // small program part: header + 3 sets + 2 decompositions
    static constexpr size_t SMALL = sizeof(ProgHeader) + 4u * (3 * NW + 2 * (RRX_MAX_TERMS * (1 + NW) + RRX_MAX_BICL * 2 * NW + NW) + NW + RRX_MAX_GROUPS * (2 * NW + 1) + 2 * NW);
    static constexpr size_t SMEM = BT_BYTES + (size_t)WARPS * REGION + SMALL + 64;
};

// =====================================================================================================
// Two-level decoupled look-back.
//
// Every warp owns one tile at a time; tiles are handed out in ticket order, so every tile with a
// smaller index is finished or in flight on a resident warp.  With ~4700 tiles in flight and all of
// them reaching the same phase at about the same time, a flat look-back would walk thousands of
// aggregates one by one.  Here
//   level 0: each tile publishes its aggregate; the LAST tile of a 32-tile block to arrive folds the
//            32 aggregates into the block aggregate (level 1);
//   a tile's exclusive prefix = fold(aggregates of the tiles before it in its block)   [1 round,
//            32 lanes read 32 entries in parallel]  combined with  fold(previous blocks' aggregates
//            back to the nearest published block-inclusive value)   [32 blocks = 1024 tiles per round];
//   the warp holding the last tile of a block publishes the block-inclusive value.
// Every wait is on a value that its producer publishes BEFORE the producer itself waits for anything,
// so the scheme cannot deadlock (exception: K_WAIT aggregates, whose consumers wait for the producer's
// inclusive value — the producer only ever waits on smaller tile indices).
// =====================================================================================================
constexpr unsigned FULL = 0xffffffffu;
// One lane's atomic add as ONE instruction: atomicAdd() would wrap it in warp-aggregation code
// (vote, leader election, popc, shuffle) although only lane 0 ever calls it here.
__device__ __forceinline__ uint32_t atom_add_u32(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long atom_add_u64(unsigned long long* p, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long CT_READY = 1ull << 63;

#ifdef RRX_BULK
// Tile staging with bulk asynchronous copies (the TMA engine's 1-D mode): each lane asks for ITS 128-byte row with one
// cp.async.bulk, completion is counted in bytes on the warp's mbarrier — no registers, no per-word instructions.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}
// Ask for the 32 rows of the tile at text offset g0 (whole tile inside the text).  Every lane has finished reading the
// staging rows (the caller's __syncwarp); generic-proxy accesses are ordered before the async-proxy writes by the fence.
__device__ __forceinline__ void bulk_stage_tile(const uint8_t* text, int64_t g0, uint32_t rows_smem, uint32_t mbar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect_tx(mbar, (uint32_t)TILE_BYTES);
    __syncwarp();
    bulk_g2s(rows_smem + (uint32_t)lane * (uint32_t)ROW_PITCH, text + g0 + ((int64_t)lane << ROW_SHIFT), (uint32_t)ROW_BYTES, mbar);
}
#endif

template <int NW> __device__ __forceinline__ void bcast_comp(Comp<NW>& c, int src) {
    c.kind = __shfl_sync(FULL, c.kind, src);
#pragma unroll
    for (int w = 0; w < NW; w++) {
        c.b[w] = __shfl_sync(FULL, c.b[w], src);
        c.U[w] = __shfl_sync(FULL, c.U[w], src);
        c.P[w] = __shfl_sync(FULL, c.P[w], src);
    }
}
// Publishing / awaiting a composite.  NW == 1: the composite and its ready bit form ONE aligned
// 16-byte word {kind|READY, b, U, P}, stored and loaded with single 128-bit accesses (not torn on
// this hardware; the same assumption CUB's decoupled look-back makes for its 16-byte tile
// descriptors), so no fence is needed.  NW > 1: payload, fence, then a separate flag word.
constexpr uint32_t K_READY = 0x80000000u;
template <int NW> __device__ __forceinline__ void publish_comp(uint32_t* pay, volatile uint32_t* flag, size_t idx, const Comp<NW>& c) {
    if (NW == 1) {
        const uint4 v = make_uint4(c.kind | K_READY, c.b[0], c.U[0], c.P[0]);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pay + idx * 4), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    } else {
        store_comp_global<NW>(pay + idx * (1 + 3 * NW), c);
        __threadfence();
        flag[idx] = 1u;
    }
}
template <int NW> __device__ __forceinline__ bool try_comp(const uint32_t* pay, volatile uint32_t* flag, size_t idx, Comp<NW>& c) {
    if (NW == 1) {
        uint4 v;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(pay + idx * 4) : "memory");
        if (!(v.x & K_READY)) return false;
        c.kind = v.x & ~K_READY; c.b[0] = v.y; c.U[0] = v.z; c.P[0] = v.w;
        return true;
    } else {
        if (flag[idx] == 0u) return false;
        __threadfence();
        load_comp_global<NW>(pay + idx * (1 + 3 * NW), c);
        return true;
    }
}
template <int NW> __device__ __forceinline__ void wait_comp(const uint32_t* pay, volatile uint32_t* flag, size_t idx, Comp<NW>& c) {
    while (!try_comp<NW>(pay, flag, idx, c)) rrx_backoff();
}

// state composites (ordered): all lanes pass the same `agg`; `acc` = composite of everything before
// the tile, relative to the shard entry (same value in all lanes).
// lookback_publish: make the tile's aggregate visible (and fold the block aggregate if this is the
// last tile of its block to arrive).  lookback_wait: composite of everything before the tile.
// A tile whose aggregate is known early (all rows closed: the exit of its last row) publishes
// before it looks for matches, so that its successors rarely wait.
template <int NW>
__device__ __forceinline__ uint32_t lookback_publish_issue(const ScanParams& sp, uint32_t tile, const Comp<NW>& agg, int lane) {
    uint32_t old = 0;
    if (lane == 0) {
        publish_comp<NW>(sp.st_a0, sp.f_a0, tile, agg);
        if (NW > 1) __threadfence();
        old = atom_add_u32(sp.st_arrive + (tile >> 5), 1u);
    }
    return old;        // lane 0: tiles of the block that had published before; consumed by lookback_publish_fold
}
// The LAST tile of a 32-tile block to arrive folds the block aggregate.  `old` is the value
// lookback_publish_issue returned: keeping the two apart lets the atomic's latency hide behind
// whatever the caller does in between.
template <int NW>
__device__ __forceinline__ void lookback_publish_fold(const ScanParams& sp, uint32_t tile, uint32_t old, int lane) {
    const uint32_t b = tile >> 5;
    const uint32_t rest = sp.num_tiles - (b << 5);
    const uint32_t nblk = rest < 32u ? rest : 32u;
    Comp<NW> mine, x, nxt;
    old = __shfl_sync(FULL, old, 0);
    if (RRX_UNLIKELY(old == nblk - 1)) {                    // fold the block aggregate, in tile order (1 tile in 32)
        comp_ident(mine);
        if ((uint32_t)lane < nblk) wait_comp<NW>(sp.st_a0, sp.f_a0, (size_t)(b << 5) + lane, mine);
        Comp<NW> fold;
        comp_ident(fold);
        bool has_wait = false;
        for (uint32_t l = 0; l < nblk; l++) {
            x = mine;
            bcast_comp(x, (int)l);
            if ((x.kind & K_TYPE) == K_WAIT) { has_wait = true; continue; }
            comp_compose(fold, x, nxt);
            fold = nxt;
        }
        if (has_wait) fold.kind = K_WAIT;
        if (lane == 0) publish_comp<NW>(sp.st_a1, sp.f_a1, b, fold);
    }
}
template <int NW>
__device__ __forceinline__ void lookback_publish(const ScanParams& sp, uint32_t tile, const Comp<NW>& agg, int lane) {
    const uint32_t old = lookback_publish_issue<NW>(sp, tile, agg, lane);
    lookback_publish_fold<NW>(sp, tile, old, lane);
}
template <int NW>
__device__ __forceinline__ void lookback_wait(const ScanParams& sp, uint32_t tile, int lane, Comp<NW>& acc) {
    const uint32_t b = tile >> 5, k = tile & 31;
    Comp<NW> mine, x, nxt;
    // (0) shortcut: if the tile right before this one is closed, nothing earlier matters
    comp_ident(acc);
    if (tile == 0) return;
    wait_comp<NW>(sp.st_a0, sp.f_a0, (size_t)tile - 1, x);
    if (!RRX_UNLIKELY((x.kind & K_TYPE) != K_CLOSED)) { acc = x; return; }
    // (a) the tiles before this one inside the block, nearest first
    comp_ident(mine);
    if ((uint32_t)lane < k) wait_comp<NW>(sp.st_a0, sp.f_a0, (size_t)(b << 5) + lane, mine);
    comp_ident(acc);
    bool done = false;
    for (int l = (int)k - 1; l >= 0 && !done; l--) {
        x = mine;
        bcast_comp(x, l);
        if ((x.kind & K_TYPE) == K_WAIT) {                   // opaque tile: take its inclusive state
            wait_comp<NW>(sp.st_i0, sp.f_i0, (size_t)(b << 5) + l, x);
            done = true;
        }
        comp_compose(x, acc, nxt);
        acc = nxt;
        if ((acc.kind & K_TYPE) == K_CLOSED) done = true;
    }
    // (b) previous blocks, 32 per round, nearest first
    for (int64_t jb = (int64_t)b - 1; jb >= 0 && !done; jb -= 32) {
        const int64_t my = jb - lane;
        uint32_t st = 0;                                    // 1 aggregate, 2 inclusive
        comp_ident(mine);
        if (my >= 0) {
            for (;;) {
                if (try_comp<NW>(sp.st_i1, sp.f_i1, (size_t)my, mine)) { st = 2; break; }
                if (try_comp<NW>(sp.st_a1, sp.f_a1, (size_t)my, mine)) {
                    st = 1;
                    if ((mine.kind & K_TYPE) == K_WAIT) { wait_comp<NW>(sp.st_i1, sp.f_i1, (size_t)my, mine); st = 2; }
                    break;
                }
                rrx_backoff();
            }
        }
        for (int l = 0; l < 32 && !done; l++) {
            const uint32_t sl = __shfl_sync(FULL, st, l);
            if (sl == 0) { done = true; break; }            // ran past the first block
            x = mine;
            bcast_comp(x, l);
            comp_compose(x, acc, nxt);
            acc = nxt;
            if (sl == 2 || (acc.kind & K_TYPE) == K_CLOSED) done = true;
        }
    }
}

// First staging slot for a tile with `total` matches (lane 0 only; see ScanParams::seg_stride).
__device__ __forceinline__ unsigned long long reserve_segment(const ScanParams& sp, uint32_t tile, uint32_t total) {
    if (total == 0) return 0;
    if (total <= sp.seg_stride) return (unsigned long long)tile * sp.seg_stride;
    const unsigned long long b = sp.ovf_base + atom_add_u64(sp.total, (unsigned long long)total);
    if (b + total > sp.out_cap) atomicOr(sp.err, ERR_STAGING_OVERFLOW);
    return b;
}

#ifdef RRX_PHASE_TIMING
#define RRX_T(i) { const long long now_ = clock64(); if (lane == 0 && sp.phase_cycles) atomicAdd(sp.phase_cycles + (i), (unsigned long long)(now_ - tprev_)); tprev_ = now_; }
#else
#define RRX_T(i)
#endif

// The slow tail of a tile — rows that received carried states, matches whose lazy start is not known yet: replay,
// count, reserve, hit list, start walks — as ONE out-of-line function with by-value arguments.  It runs for a small
// minority of the tiles; out of line it stays out of the instruction-cache lines the hot path lives in
// (RRX_SLOW_INLINE puts it back, A/B).
#ifdef RRX_SLOW_INLINE
#define RRX_SLOW __device__ __forceinline__
#else
#define RRX_SLOW __device__ __noinline__
#endif
template <int NW, int CW, class F>
RRX_SLOW void slow_tail(const ScanParams& sp, ProgView pv, F fw, TileState<NW, CW> ts, int lane, bool any_open, uint32_t tile) {
    using TS = TileState<NW, CW>;
        const uint32_t c = phase_b<NW>(sp, pv, fw, ts, lane, any_open);
        // exclusive offsets of the rows' matches inside the tile
        uint32_t inc_c = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, inc_c, d);
            if (lane >= d) inc_c += v;
        }
        const uint32_t total = __shfl_sync(FULL, inc_c, 31);
        const uint32_t offset = inc_c - c;
        
        // reserve the tile's segment of the staging list; the hit list is built while the atomic is in flight
        unsigned long long base = 0;
        if (lane == 0) base = reserve_segment(sp, tile, total);
        
        if (total) {
            uint32_t hit[4];
#pragma unroll
            for (int q = 0; q < 4; q++) hit[q] = ts.hits[q * TILE_ROWS + lane];
            __syncwarp();                 // every lane is done with the row summaries: reuse as hit list
            const bool fits = total <= (uint32_t)TS::LIST_CAP;
            if (fits) phase_c1_list<NW>(ts, lane, offset, hit);
            __syncwarp();
            if (lane == 0) {
                sp.tile_seg[tile] = base;
                sp.tile_cnt[tile] = total;
            }
            base = __shfl_sync(FULL, base, 0);
            if (fits) {
                phase_c2_balanced<NW>(sp, pv, fw, ts, lane, total, base);
            } else {
                phase_c2_rows<NW>(sp, pv, fw, ts, lane, offset, hit, base);
            }
        } else if (lane == 0) {
            sp.tile_seg[tile] = 0;
            sp.tile_cnt[tile] = 0;
        }
}

// NW: words of a state set.  CW: words of a carried set (composites, row exits, look-back payloads); NW unless the
// kernel is specialised for a thread-owned program with fewer dominated words (rrx_prog.h CW).
template <int NW, int CW, class F, int WARPS>
__device__ __forceinline__ void rrx_scan_body(const ScanParams& sp) {
    using K = KCfg<NW, CW, WARPS>;
    using TS = TileState<NW, CW>;
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
        for (uint32_t i = tid; i < K::BT_BYTES / 4; i += K::THREADS) Btab[i] = sp.prog[off_B + i / (uint32_t)BT_COPIES];
    } else {
        for (uint32_t i = tid; i < K::BT_BYTES / 4; i += K::THREADS) Btab[i] = sp.prog[off_B + i];
    }
#ifdef RRX_BULK
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(regions + (size_t)warp * K::REGION + K::REGION - 16);
    const uint32_t rows_smem = (uint32_t)__cvta_generic_to_shared(regions + (size_t)warp * K::REGION);
    uint32_t mphase = 0;
    if (lane == 0) {
        mbar_init(mbar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();                      // the only block-level barrier of the kernel
    ProgView pv;
    pv.sm = progsm;
    pv.glob = sp.prog;
    pv.Blane = reinterpret_cast<const uint8_t*>(Btab) + (NW == 1 ? (lane % BT_COPIES) * 4 : 0);
    F fw;
    fw.init(pv);
    TS ts;
    ts.carve(regions + (size_t)warp * K::REGION);
    // pad bytes behind every staged row: a byte that no position accepts (see phase_a_cand)
    *reinterpret_cast<uint32_t*>(ts.text + lane * ROW_PITCH + ROW_BYTES) = 0xffffffffu;
    if (lane == 0) ts.misc[TS::M_ZVALID] = 0u;
    const int64_t len = (int64_t)sp.len;
#ifdef RRX_PHASE_TIMING
    long long tprev_ = clock64();
#endif

    // De-phase the warps of the SM: if all of them load, scan, look back and emit at the same
    // moments, nobody computes while everybody waits for memory.  Tickets are taken in arrival
    // order afterwards, so the offsets persist.
    if (sp.stagger_ns) {
        const unsigned long long t0 = globaltimer_ns();
        const unsigned long long wait = (unsigned long long)(warp & 7) * sp.stagger_ns;
        while (globaltimer_ns() - t0 < wait) __nanosleep(200);
    }

    // Tickets.  A tile that publishes its aggregate early (all rows closed) takes the ticket of the
    // warp's NEXT tile right then: the atomic's latency hides behind the match search, and the next
    // tile's 32 lines are prefetched into L2 as soon as the ticket arrives.  Taking tickets any
    // earlier makes successors of a not-yet-started tile wait for its holder (head-of-line blocking
    // whenever tile times vary), so all other tiles take theirs when they are done.
    const uint32_t warps_in_flight = gridDim.x * (uint32_t)K::WARPS;
    int dense_left = 0;
    // ticket_ahead == 3: no tickets at all — warp g of the grid takes tiles g, g + W, g + 2W, ...
    // (W = warps in flight).  Every wait is still on a smaller tile index owned by a resident warp
    // that reaches it before any larger one, so the look-back cannot deadlock; what is lost is
    // dynamic balancing, what is won is one contended global atomic per tile.
    const bool static_tiles = sp.ticket_ahead == 3;
    // ticket_ahead == 5: a tile on the fast tail (its text is dead by then) takes the next ticket right
    // after its look-back and, before it writes its matches, starts asynchronous copies (cp.async,
    // no registers involved) of the next tile straight into its staging rows: the ticket and the
    // HBM/L2 latency hide behind the tail instead of heading the next tile.
    bool preloaded = false;
    uint32_t tile = 0;
    if (static_tiles) {
        tile = blockIdx.x * (uint32_t)K::WARPS + (uint32_t)warp;
    } else {
        if (lane == 0) tile = atom_add_u32(sp.ticket, 1u);
        tile = __shfl_sync(FULL, tile, 0);
    }
    for (;;) {
        if (tile >= sp.num_tiles) break;
        uint32_t next_raw = 0;
        bool have_next = false, late_ticket = false;
        if (static_tiles) {
            next_raw = tile + warps_in_flight;
            have_next = true;
            const int64_t g = (((int64_t)next_raw) * TILE_ROWS + lane) << ROW_SHIFT;
            if (g < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp.text + g));
        } else if (sp.ticket_ahead == 2) {
            if (lane == 0) next_raw = atom_add_u32(sp.ticket, 1u);
            have_next = true;
        } else {
            // guess which tile some warp will pick up about one round from now, and ask L2 for it
            const int64_t g = (((int64_t)tile + warps_in_flight) * TILE_ROWS + lane) << ROW_SHIFT;
            if (g < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp.text + g));
        }
        ts.row0 = (uint64_t)tile * TILE_ROWS;
        RRX_T(0)
        // ---- stage the tile: 8 x 128-bit loads per lane (fully coalesced), stored as 4 words each
        //      into the rows (pitch ROW_PITCH: one 128-bit store per chunk when the pitch is 16-byte aligned) -
        {
            const int64_t g0 = ts.lo();
            uint4 v[8];
#ifdef RRX_BULK
            if (!preloaded && g0 + TILE_BYTES <= len) {
                __syncwarp();
                bulk_stage_tile(sp.text, g0, rows_smem, mbar, lane);
                preloaded = true;
            }
            if (preloaded) {                              // bulk copies in flight (started just now, or during the previous tile's tail)
                mbar_wait(mbar, mphase);
                mphase ^= 1u;
            } else if (g0 + TILE_BYTES <= len) {
#else
            if (preloaded) {                              // copies started during the previous tile's tail
                asm volatile("cp.async.wait_all;" ::: "memory");
            } else if (g0 + TILE_BYTES <= len) {          // whole tile inside the text: no per-load checks
#endif
                const uint4* src = reinterpret_cast<const uint4*>(sp.text + g0) + lane;
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = __ldcs(src + k * 32);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int64_t g = g0 + (int64_t)(k * 32 + lane) * 16;
                    if (g + 16 <= len) {
                        v[k] = __ldcs(reinterpret_cast<const uint4*>(sp.text + g));
                    } else {
                        // bytes behind the end of the text: 0xff, a byte no position accepts (the thread-owned scan
                        // relies on it; the other scans never look at them)
                        uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                        for (int b = 0; b < 16 && g + b < len; b++) w[b >> 2] = (w[b >> 2] & ~(0xffu << (8 * (b & 3)))) | ((uint32_t)sp.text[g + b] << (8 * (b & 3)));
                        v[k] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            if (!preloaded) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    uint32_t* dst = ts.chunk_ptr(k * 32 + lane);
#if RRX_ROW_PITCH % 16 == 0
                    *reinterpret_cast<uint4*>(dst) = v[k];
#else
                    dst[0] = v[k].x; dst[1] = v[k].y; dst[2] = v[k].z; dst[3] = v[k].w;
#endif
                }
            }
            preloaded = false;
        }
        if (lane == 0) ts.misc[TS::M_RQCNT] = 0;
        __syncwarp();
        RRX_T(1)
        // Candidate-driven row scan where the automaton has a cut.  Exit sets first: if every row is
        // closed the tile aggregate is the last row's exit and is published at once.  If any row of the
        // tile needs the dense scan the whole tile takes it (the lanes run in lock-step anyway), and
        // the next few tiles of this warp do not try again.
        bool dense = true, published = false, agg_done = false;
        Comp<CW> agg, acc, inc;
        SparseTail tail;
        tail.kq = tail.ks = tail.nacc = tail.nruns = 0;
        tail.rt = 0ull;
        tail.drained = 1;
        if (fw.donly(pv) && fw.rs_n(pv) != 0u) {
            // The program's backward walks have a skip state: note where its stop bytes are — per row in shared memory
            // (walks inside the tile), and the tile's last one in the stop-byte table (walks that arrive from later tiles).
            uint32_t zm[4];
            phase_z_mask<NW>(ts, lane, fw.rs_bytes(pv), fw.rs_n(pv), zm);
            const uint32_t rowhas = __ballot_sync(FULL, (zm[0] | zm[1] | zm[2] | zm[3]) != 0u);
            if (lane == 0) { ts.misc[TS::M_ROWHAS] = rowhas; ts.misc[TS::M_ZVALID] = 1u; }
            if (rowhas == 0u ? lane == 0 : lane == 31 - __clz((int)rowhas)) {
                uint32_t v = 1u;
                if (rowhas) {
                    const int q = zm[3] ? 3 : zm[2] ? 2 : zm[1] ? 1 : 0;
                    const uint32_t m = q == 3 ? zm[3] : q == 2 ? zm[2] : q == 1 ? zm[1] : zm[0];
                    v = 2u + (uint32_t)lane * ROW_BYTES + (uint32_t)q * 32u + (31u - (uint32_t)__clz((int)m));
                }
                sp.tile_stop[tile] = v;
            }
            __syncwarp();
        }
        if (fw.donly(pv)) {
            // Thread-owned candidate scan: every row — full, partial or empty — goes through it; there is no dense
            // fallback.  The rows' exits (dominated states only) are known once the candidates have been followed;
            // what a row does to carried states is decided by phase_d_kind.  All rows closed: the tile's aggregate
            // is the last row's exit.
            // What a row does to CARRIED states (phase_d_kind) matters for the last row (the tile's aggregate: a closed last
            // row makes the tile closed, whatever precedes it), for row 0 (it may receive the tile's entry) and for rows that
            // receive a state from the row before — known once the candidates have been followed, and rare.
            // Usually the last row's exit is provably empty before any candidate has been followed (no candidate byte
            // behind the point where nothing carried survives): the tile then publishes FIRST, and its successors find
            // its aggregate ready when they look back.
#ifdef RRX_EAGER_KINDS
            // (lean kernel: every row's kind up front, one inlined copy of phase_d_kind)
            uint32_t cm[4];
            int qfree;
            phase_a_mask<NW>(ts, lane, fw.sp_cand(pv), cm);
            const uint32_t kind = phase_d_kind<NW>(sp, pv, fw, ts, lane, &qfree);
            const uint32_t behind = qfree >= ROW_BYTES ? 0u : ((qfree >> 5) == 0 ? (cm[0] >> (qfree & 31)) | cm[1] | cm[2] | cm[3]
                                                             : (qfree >> 5) == 1 ? (cm[1] >> (qfree & 31)) | cm[2] | cm[3]
                                                             : (qfree >> 5) == 2 ? (cm[2] >> (qfree & 31)) | cm[3] : (cm[3] >> (qfree & 31)));
            const bool all_closed = __all_sync(FULL, kind == K_CLOSED) != 0;
            const bool early = all_closed && __shfl_sync(FULL, behind, TILE_ROWS - 1) == 0u;
            uint32_t arrived = 0;
            dense = false;
            if (early) {
                uint32_t bl[CW];
#pragma unroll
                for (int w = 0; w < CW; w++) bl[w] = 0u;
                comp_closed(agg, bl, 0);
                arrived = lookback_publish_issue<CW>(sp, tile, agg, lane);
                published = true;
            }
            phase_a_cand<NW, F, true>(sp, pv, fw, ts, lane, cm, tail);
            __syncwarp();
            if (early) {
                lookback_publish_fold<CW>(sp, tile, arrived, lane);       // the arrival count has long come back
            } else if (all_closed) {
                uint32_t bl[CW];
#pragma unroll
                for (int w = 0; w < CW; w++) bl[w] = ts.sum_b[w * TILE_ROWS + (TILE_ROWS - 1)];
                comp_closed(agg, bl, 0);
                lookback_publish<CW>(sp, tile, agg, lane);
                published = true;
            }
#else
            uint32_t cm[4];
            int qfree = 0;
            phase_a_mask<NW>(ts, lane, fw.sp_cand(pv), cm);
            const bool edge = lane == 0 || lane == TILE_ROWS - 1;
            uint32_t kind = K_CLOSED;
            if (edge) kind = phase_d_kind<NW>(sp, pv, fw, ts, lane, &qfree);
            const uint32_t behind = qfree >= ROW_BYTES ? 0u : ((qfree >> 5) == 0 ? (cm[0] >> (qfree & 31)) | cm[1] | cm[2] | cm[3]
                                                             : (qfree >> 5) == 1 ? (cm[1] >> (qfree & 31)) | cm[2] | cm[3]
                                                             : (qfree >> 5) == 2 ? (cm[2] >> (qfree & 31)) | cm[3] : (cm[3] >> (qfree & 31)));
            const bool early = __all_sync(FULL, kind == K_CLOSED) != 0 && __shfl_sync(FULL, behind, TILE_ROWS - 1) == 0u;
            uint32_t arrived = 0;
            dense = false;
            if (early) {
                uint32_t bl[CW];
#pragma unroll
                for (int w = 0; w < CW; w++) bl[w] = 0u;
                comp_closed(agg, bl, 0);
                arrived = lookback_publish_issue<CW>(sp, tile, agg, lane);
            }
            phase_a_cand<NW, F, true>(sp, pv, fw, ts, lane, cm, tail);
            __syncwarp();
            uint32_t carried_in = 0;
            if (!edge) {
#pragma unroll
                for (int w = 0; w < CW; w++) carried_in |= ts.sum_b[w * TILE_ROWS + lane - 1];
            }
            // (ONE more inlined copy of phase_d_kind, entered at most twice: rows that receive a state; then, if some row that
            // matters is open or the tile is partial, every remaining row — the general composition wants all kinds)
            bool rows_ok = true, have_kind = edge;
#ifdef RRX_KIND_PASSES
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const bool mine = !have_kind && (pass == 1 || carried_in != 0u);
                if (RRX_UNLIKELY(__any_sync(FULL, mine) != 0)) {
                    if (mine) { kind = phase_d_kind<NW>(sp, pv, fw, ts, lane, &qfree); have_kind = true; }
                }
                if (pass == 0) {
                    rows_ok = __all_sync(FULL, kind == K_CLOSED) != 0;
                    if (!RRX_UNLIKELY(!rows_ok)) break;
                }
            }
#else
            if (RRX_UNLIKELY(__any_sync(FULL, carried_in != 0u) != 0)) {
                if (carried_in != 0u) kind = phase_d_kind<NW>(sp, pv, fw, ts, lane, &qfree);
            }
            rows_ok = __all_sync(FULL, kind == K_CLOSED) != 0;
            if (RRX_UNLIKELY(!rows_ok)) {
                if (!edge && carried_in == 0u) phase_d_kind<NW>(sp, pv, fw, ts, lane, &qfree);
            }
#endif
            if (early) lookback_publish_fold<CW>(sp, tile, arrived, lane);       // the arrival count has long come back
            if (!RRX_UNLIKELY(!rows_ok)) {
                if (RRX_UNLIKELY(!early)) {
                    uint32_t bl[CW];
#pragma unroll
                    for (int w = 0; w < CW; w++) bl[w] = ts.sum_b[w * TILE_ROWS + (TILE_ROWS - 1)];
                    comp_closed(agg, bl, 0);
                    lookback_publish<CW>(sp, tile, agg, lane);
                }
                published = true;
            } else {
                // an aggregate published early stands (it is what the general composition will find)
                agg_done = early;
                __syncwarp();
            }
#endif
        } else if (sp.sparse) {
            if (dense_left == 0) {
                const bool closed = phase_a_exit<NW>(sp, pv, fw, ts, lane);
                if (__all_sync(FULL, closed)) {
                    __syncwarp();
                    uint32_t bl[CW];
#pragma unroll
                    for (int w = 0; w < CW; w++) bl[w] = ts.sum_b[w * TILE_ROWS + (TILE_ROWS - 1)];
                    comp_closed(agg, bl, 0);
                    const uint32_t arrived = lookback_publish_issue<CW>(sp, tile, agg, lane);
                    published = true;
                    if (sp.ticket_ahead == 1) {
                        if (lane == 0) next_raw = atom_add_u32(sp.ticket, 1u);
                        have_next = true;
                    }
                    uint32_t cm[4];
                    phase_a_mask<NW>(ts, lane, fw.sp_cand(pv), cm);
                    const bool ok = phase_a_cand<NW, F, false>(sp, pv, fw, ts, lane, cm, tail);
                    lookback_publish_fold<CW>(sp, tile, arrived, lane);       // the arrival count has long come back
                    dense = __any_sync(FULL, !ok) != 0;
                }
                if (dense) dense_left = 4;
            } else {
                dense_left--;
            }
        }
        if (dense) {
            __syncwarp();
            phase_a<NW>(sp, pv, fw, ts, lane);
            __syncwarp();
            phase_a2<NW>(pv, fw, ts, lane);
        }
        uint32_t next_tile = 0;
        if (static_tiles) {
            next_tile = next_raw;
        } else if (have_next) {        // the next tile's ticket has arrived by now: ask L2 for its lines
            next_tile = __shfl_sync(FULL, next_raw, 0);
            const int64_t g = (((int64_t)next_tile) * TILE_ROWS + lane) << ROW_SHIFT;
            if (g < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp.text + g));
        }
        // A tile that published early has 32 closed rows: its aggregate (= its inclusive composite, a
        // closed map absorbs whatever precedes it) is the last row's exit, and row t's entry is row
        // t-1's exit — phases S1/S2 have nothing to compute beyond the tile's entry set.
        bool any_open = false;
        SLocal<CW> L;
        if (RRX_UNLIKELY(!published)) {
            const uint32_t my_kind = ts.sum_kind[lane];
            any_open = __any_sync(FULL, my_kind == K_RANK1 || my_kind == K_HARD) != 0;
            __syncwarp();
            if (lane == 0) { phase_s1<NW>(sp, pv, fw, ts, any_open, L); agg = L.run; } else { comp_ident(agg); }
            bcast_comp(agg, 0);
            if (!agg_done) lookback_publish<CW>(sp, tile, agg, lane);
        }
        RRX_T(2)
        lookback_wait<CW>(sp, tile, lane, acc);
        if (((sp.ticket_ahead == 4) || (sp.ticket_ahead == 5 && published && !dense)) && !have_next) {       // the tile's tail is short: take the next ticket now
            if (lane == 0) next_raw = atom_add_u32(sp.ticket, 1u);
            have_next = true;
            late_ticket = true;
        }
        RRX_T(3)
        if (lane == 0) {
            if (published) {
                uint32_t E[CW];
                comp_apply<CW>(acc, sp.entry, E);
#pragma unroll
                for (int w = 0; w < CW; w++) ts.misc[TS::M_ENTRY + w] = E[w];
                inc = agg;
            } else {
                phase_s2<NW>(sp, pv, fw, ts, L, acc, inc);
            }
            publish_comp<CW>(sp.st_i0, sp.f_i0, tile, inc);
            const uint32_t rest = sp.num_tiles - (tile & ~31u);
            if ((tile & 31u) == (rest < 32u ? rest : 32u) - 1u) publish_comp<CW>(sp.st_i1, sp.f_i1, tile >> 5, inc);
            if (tile + 1 == sp.num_tiles) store_comp_external<NW, CW>(sp.shard_comp, inc);
        }
        __syncwarp();
        RRX_T(4)
        // Fast tail: the sparse scan kept every match of every row in its register queue with a known
        // start, and no row receives a carried state (its predecessor's exit / the tile entry is empty):
        // nothing to replay, nothing to resolve — count, reserve, write.
        bool fast = false;
        if (!dense) {
            uint32_t carried = 0;
#pragma unroll
            for (int w = 0; w < CW; w++) carried |= lane == 0 ? ts.misc[TS::M_ENTRY + w] : ts.sum_b[w * TILE_ROWS + lane - 1];
            // (only tiles that published early still have the rows' exits in sum_b: phase S1 rewrites them otherwise)
            fast = published && __all_sync(FULL, tail.simple() && carried == 0) != 0;
#if defined(RRX_NO_OVERRUN) && !defined(RRX_NO_COUNTER)
            // lean kernel: a tile that leaves the fast tail counts towards the switch to the straddle kernel
            // (ScanParams::straddle_tiles; on the slow path only — the fast path pays nothing for it)
            if (!fast && fw.donly(pv) && lane == 0 && sp.straddle_tiles) atom_add_u32(sp.straddle_tiles, 1u);
#endif
            if (RRX_UNLIKELY(!fast)) sparse_drain<NW>(ts, lane, tail);
            __syncwarp();
        }
        if (!RRX_UNLIKELY(!fast)) {
            uint32_t inc_c, total;
            if (__all_sync(FULL, tail.nacc <= 1u)) {       // the usual case: a vote and two popcounts
                const uint32_t have = __ballot_sync(FULL, tail.nacc != 0u);
                inc_c = (uint32_t)__popc(have & (0xffffffffu >> (31 - lane)));
                total = (uint32_t)__popc(have);
            } else {
                inc_c = tail.nacc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, inc_c, d);
                    if (lane >= d) inc_c += v;
                }
                total = __shfl_sync(FULL, inc_c, 31);
            }
            unsigned long long base = 0;
            if (lane == 0) {
                base = reserve_segment(sp, tile, total);
                sp.tile_seg[tile] = base;
                sp.tile_cnt[tile] = total;
            }
            base = __shfl_sync(FULL, base, 0);
            if (sp.ticket_ahead == 5 && late_ticket) {
                const uint32_t nt = __shfl_sync(FULL, next_raw, 0);
                const int64_t n0 = ((int64_t)nt * TILE_ROWS) << ROW_SHIFT;
                if (nt < sp.num_tiles && n0 + TILE_BYTES <= len) {
#ifdef RRX_BULK
                    __syncwarp();
                    bulk_stage_tile(sp.text, n0, rows_smem, mbar, lane);
#elif RRX_ROW_PITCH % 16 == 0
                    // 16-byte aligned rows: chunk q = 32k + lane of the tile -> row q/8, bytes 16*(q%8).. (8 copies per lane)
                    const uint8_t* src = sp.text + n0 + 16 * lane;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ts.text) + (uint32_t)(lane >> 3) * (uint32_t)ROW_PITCH + 16u * (uint32_t)(lane & 7);
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 4u * (uint32_t)ROW_PITCH * k), "l"(src + 512 * k) : "memory");
#else
                    // word j = 32k + lane of the tile -> staged row k, word `lane` (rows are 33 words apart)
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(sp.text + n0) + lane;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ts.text) + 4u * (uint32_t)lane;
#pragma unroll
                    for (int k = 0; k < 32; k++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)ROW_PITCH * k), "l"(src + 32 * k) : "memory");
#endif
                    preloaded = true;
                }
            }
            phase_tail_fast<NW>(sp, ts, lane, tail, base + (inc_c - tail.nacc));
        } else {
            slow_tail<NW, CW, F>(sp, pv, fw, ts, lane, any_open, tile);
        }
        __syncwarp();
        RRX_T(7)
        if (!have_next) {
            if (lane == 0) next_raw = atom_add_u32(sp.ticket, 1u);
            next_tile = __shfl_sync(FULL, next_raw, 0);
        } else if (late_ticket) {
            next_tile = __shfl_sync(FULL, next_raw, 0);
        }
        tile = next_tile;
    }
}

}  // namespace rrx
