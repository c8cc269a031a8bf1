// TEST INFRASTRUCTURE ONLY — host-side emulator of the rrx scan kernel.
//
// Compiles the very same phase functions the GPU kernel runs (roaringregex_b200/csrc/rrx_tile.cuh,
// with -DRRX_HOST_EMU) and steps them tile by tile, thread by thread, in the order the kernel's
// __syncthreads() barriers impose.  Lets the CPU-only test tier check the chunk-stitching logic
// (row transfer functions, look-back, prefix replay, lazy start scans) against the oracle.
// librrx.so does not contain this file and the product never runs it.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#define RRX_HOST_EMU 1
#include "../../roaringregex_b200/csrc/rrx_compile.h"
#include "../../roaringregex_b200/csrc/rrx_tile.cuh"

using namespace rrx;

namespace {

// 0: dense row scans only; 1: candidate-driven row scan with dense fallback, as the kernel does;
// 2: run both on every tile and report (err bit 0x80) any row where they disagree
int g_sparse_mode = 1;
unsigned long long g_sparse_rows = 0, g_dense_rows = 0, g_fast_tiles = 0;

// test-only policy: the generic shared-memory program with a narrower carried width, standing in for the run-time
// specialised kernels of thread-owned programs (rrx_prog.h CW)
template <int NW, int CW_> struct ProgNarrow : Prog<NW, -1, 0> {};

template <int NW, int CW, class PROG>
int run(const Program& prog, const uint8_t* text, uint64_t left_ctx, uint64_t len, int64_t base,
        const uint32_t* entry, int64_t* out, uint64_t cap, uint64_t* counts, uint32_t* err, uint32_t* shard_comp,
        PendingStart* pending, uint32_t pending_cap) {
    constexpr int ROWS = TILE_ROWS;
    using TS = TileState<NW, CW>;
    const ProgHeader& h = prog.hdr();
    const uint32_t num_tiles = (uint32_t)((len + TILE_BYTES - 1) / TILE_BYTES);

    std::vector<uint8_t> smem(TS::bytes() + 64, 0xAA);
    std::vector<uint32_t> btab(NW == 1 ? 256 * 32 : 256 * NW);
    for (uint32_t i = 0; i < btab.size(); i++) btab[i] = prog.blob[h.off_B + (NW == 1 ? (i >> 5) : i)];
    std::vector<uint32_t> incl((size_t)(num_tiles + 1) * (1 + 3 * CW), 0);   // inclusive composite per tile
    unsigned long long matches_so_far = 0;
    std::vector<unsigned long long> tile_base(num_tiles + 1, 0);
    std::vector<uint32_t> tile_stop(num_tiles + 1, 0);
    unsigned long long total[2] = {0, 0};
    uint32_t ticket = 0;
    *err = 0;

    ScanParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.text = text;
    sp.len = len;
    sp.left_ctx = left_ctx;
    sp.base = base;
    sp.prog = prog.blob.data();
    sp.num_tiles = num_tiles;
    sp.static_closed = (h.static_depth != 0xffffffffu && h.static_depth + 1 <= (uint32_t)ROW_BYTES) ? 1u : 0u;
    sp.sparse = h.sparse_ok && g_sparse_mode ? 1u : 0u;
    sp.donly = h.donly;
    sp.tile_stop = (h.donly && h.rskip_n && g_sparse_mode) ? tile_stop.data() : nullptr;
    sp.ticket = &ticket;
    sp.out = reinterpret_cast<long long*>(out);
    sp.out_cap = cap;
    sp.total = total;
    sp.err = err;
    sp.shard_comp = shard_comp;
    sp.pending = pending;
    sp.pending_cap = pending_cap;
    if (entry) for (int w = 0; w < NW; w++) sp.entry[w] = entry[w];

    if (num_tiles == 0) { shard_comp[0] = K_IDENT; for (int i = 1; i < 1 + 3 * NW; i++) shard_comp[i] = 0; }

    auto view = [&](int lane) {
        ProgView pv;
        pv.sm = prog.blob.data();
        pv.glob = prog.blob.data();
        pv.Blane = reinterpret_cast<const uint8_t*>(btab.data()) + (NW == 1 ? lane * 4 : 0);
        return pv;
    };

    for (uint32_t tile = 0; tile < num_tiles; tile++) {
        TS ts;
        ts.carve(smem.data());
        ts.row0 = (uint64_t)tile * ROWS;
        std::memset(smem.data(), 0xAA, smem.size());
        const int64_t g0 = ts.lo();
        for (int q = 0; q < ROWS * 8; q++) {
            const int64_t g = g0 + (int64_t)q * 16;
            uint8_t* dst = reinterpret_cast<uint8_t*>(ts.chunk_ptr(q));
            for (int k = 0; k < 16 && g + k < (int64_t)len; k++) dst[k] = text[g + k];
        }
        bool any_open = false;
        ts.misc[TS::M_RQCNT] = 0;
        bool dense = true;
        bool sp_ok[ROWS];
        uint32_t sp_hits[ROWS][4], sp_b[ROWS][NW];
        uint32_t sp_kst[ROWS][2];
        SparseTail tails[ROWS];
        for (int t = 0; t < ROWS; t++) { tails[t].kq = tails[t].ks = tails[t].nacc = tails[t].nruns = 0; tails[t].rt = 0; tails[t].drained = 1; }
        bool published = false, all_closed_d = true;
        ts.misc[TS::M_ZVALID] = 0;
        if (h.donly && g_sparse_mode && h.rskip_n) {
            // stop-byte masks of the tile + its entry of the stop-byte table, as the kernel
            uint32_t rowhas = 0, v = 1;
            for (int t = 0; t < ROWS; t++) {
                uint32_t zm[4];
                phase_z_mask<NW>(ts, t, h.rskip_bytes, h.rskip_n, zm);
                if (zm[0] | zm[1] | zm[2] | zm[3]) {
                    rowhas |= 1u << t;
                    const int q = zm[3] ? 3 : zm[2] ? 2 : zm[1] ? 1 : 0;
                    v = 2u + (uint32_t)t * ROW_BYTES + (uint32_t)q * 32u + (31u - (uint32_t)__builtin_clz(zm[q]));
                }
            }
            ts.misc[TS::M_ROWHAS] = rowhas;
            ts.misc[TS::M_ZVALID] = 1;
            tile_stop[tile] = v;
        }
        if (h.donly && g_sparse_mode) {
            // thread-owned candidate scan, as the kernel: every row, no dense fallback.  Row kinds lazily: the two edge rows
            // first, rows that receive a carried state after the candidates have been followed, the rest only if the
            // general composition is needed.
            uint32_t kinds[ROWS];
            bool have_kind[ROWS];
            for (int t = 0; t < ROWS; t++) {
                ProgView pv = view(t);
                PROG fw;
                fw.init(pv);
                uint32_t cm[4];
                int qfree = 0;
                phase_a_mask<NW>(ts, t, fw.sp_cand(pv), cm);
                have_kind[t] = t == 0 || t == ROWS - 1;
                kinds[t] = have_kind[t] ? phase_d_kind<NW>(sp, pv, fw, ts, t, &qfree) : (uint32_t)K_CLOSED;
                phase_a_cand<NW, PROG, true>(sp, pv, fw, ts, t, cm, tails[t]);
                // the kernel's early-publish rule: no candidate at or behind qfree => the row's exit must be empty
                bool behind = false;
                for (int p = qfree; p < ROW_BYTES; p++) behind = behind || ((cm[p >> 5] >> (p & 31)) & 1u);
                if (have_kind[t] && kinds[t] == K_CLOSED && !behind)
                    for (int w = 0; w < CW; w++) if (ts.sum_b[w * ROWS + t]) *err |= 0x40u;
            }
            for (int t = 1; t < ROWS - 1; t++) {
                uint32_t carried_in = 0;
                for (int w = 0; w < CW; w++) carried_in |= ts.sum_b[w * ROWS + t - 1];
                if (carried_in) {
                    ProgView pv = view(t);
                    PROG fw;
                    fw.init(pv);
                    int qf;
                    kinds[t] = phase_d_kind<NW>(sp, pv, fw, ts, t, &qf);
                    have_kind[t] = true;
                }
            }
            for (int t = 0; t < ROWS; t++) if (kinds[t] != K_CLOSED) all_closed_d = false;
            if (!all_closed_d)
                for (int t = 0; t < ROWS; t++) if (!have_kind[t]) {
                    ProgView pv = view(t);
                    PROG fw;
                    fw.init(pv);
                    int qf;
                    phase_d_kind<NW>(sp, pv, fw, ts, t, &qf);
                }
            published = all_closed_d;
            dense = false;
            g_sparse_rows += ROWS;
        } else if (h.sparse_ok && g_sparse_mode) {
            // as the kernel: exit sets first; candidates only if every row of the tile is closed
            bool all_closed = true;
            for (int t = 0; t < ROWS; t++) {
                ProgView pv = view(t);
                PROG fw;
                fw.init(pv);
                sp_ok[t] = phase_a_exit<NW>(sp, pv, fw, ts, t);
                all_closed = all_closed && sp_ok[t];
            }
            dense = !all_closed;
            published = all_closed;
            if (all_closed) {
                for (int t = 0; t < ROWS; t++) {
                    ProgView pv = view(t);
                    PROG fw;
                    fw.init(pv);
                    uint32_t cm[4];
                    phase_a_mask<NW>(ts, t, fw.sp_cand(pv), cm);
                    sp_ok[t] = phase_a_cand<NW, PROG, false>(sp, pv, fw, ts, t, cm, tails[t]);
                    if (!sp_ok[t]) dense = true;
                    else if (g_sparse_mode == 2) {
                        sparse_drain<NW>(ts, t, tails[t]);      // materialise for the comparison
                        if (ts.sum_kind[t] != K_CLOSED) *err |= 0x80u;
                        for (int q = 0; q < 4; q++) sp_hits[t][q] = ts.hits[q * ROWS + t];
                        for (int w = 0; w < CW; w++) sp_b[t][w] = ts.sum_b[w * ROWS + t];
                        sp_kst[t][0] = ts.kst[2 * t]; sp_kst[t][1] = ts.kst[2 * t + 1];
                    }
                }
            } else {
                for (int t = 0; t < ROWS; t++) sp_ok[t] = false;
            }
            if (g_sparse_mode == 2) dense = true;
            (dense ? g_dense_rows : g_sparse_rows) += ROWS;
        }
        if (dense) {
            for (int t = 0; t < ROWS; t++) {
                ProgView pv = view(t);
                PROG fw;
                fw.init(pv);
                phase_a<NW>(sp, pv, fw, ts, t);
            }
            for (int t = 0; t < ROWS; t++) {
                ProgView pv = view(t);
                PROG fw;
                fw.init(pv);
                phase_a2<NW>(pv, fw, ts, t);
            }
            if (h.sparse_ok && g_sparse_mode == 2) {
                for (int t = 0; t < ROWS; t++) {
                    if (!sp_ok[t]) continue;
                    if (ts.sum_kind[t] != K_CLOSED) *err |= 0x80u;
                    for (int q = 0; q < 4; q++) if (sp_hits[t][q] != ts.hits[q * ROWS + t]) *err |= 0x80u;
                    for (int w = 0; w < CW; w++) if (sp_b[t][w] != ts.sum_b[w * ROWS + t]) *err |= 0x80u;
                    ts.kst[2 * t] = sp_kst[t][0]; ts.kst[2 * t + 1] = sp_kst[t][1];   // let phase C use them
                }
            }
        }
        for (int t = 0; t < ROWS; t++) any_open = any_open || ts.sum_kind[t] == K_RANK1 || ts.sum_kind[t] == K_HARD;
        {
            ProgView pv = view(0);
            PROG fw;
            fw.init(pv);
            // The emulator runs the tiles in order, so "look-back" is just the previous tile's
            // inclusive composite; the two-level GPU look-back (rrx_scan.cu) is exercised by -m gpu.
            SLocal<CW> L;
            Comp<CW> acc, inc;
            phase_s1<NW>(sp, pv, fw, ts, any_open, L);
            if (tile == 0) comp_ident(acc);
            else load_comp_global<CW>(incl.data() + (size_t)(tile - 1) * (1 + 3 * CW), acc);
            phase_s2<NW>(sp, pv, fw, ts, L, acc, inc);
            store_comp_global<CW>(incl.data() + (size_t)tile * (1 + 3 * CW), inc);
            if (tile + 1 == num_tiles) store_comp_external<NW, CW>(sp.shard_comp, inc);
        }
        uint32_t cnt[ROWS], off[ROWS], total_t = 0;
        // fast tail, as in the kernel: every row simple and without carried state
        bool fast = !dense && published;
        if (!dense) {
            for (int t = 0; t < ROWS; t++) {
                uint32_t carried = 0;
                for (int w = 0; w < CW; w++) carried |= t == 0 ? ts.misc[TS::M_ENTRY + w] : ts.sum_b[w * ROWS + t - 1];
                if (!tails[t].simple() || carried) fast = false;
            }
            if (!fast) for (int t = 0; t < ROWS; t++) sparse_drain<NW>(ts, t, tails[t]);
        }
        tile_base[tile] = matches_so_far;
        if (fast) {
            g_fast_tiles++;
            for (int t = 0; t < ROWS; t++) { off[t] = total_t; total_t += tails[t].nacc; }
            for (int t = 0; t < ROWS; t++) phase_tail_fast<NW>(sp, ts, t, tails[t], matches_so_far + off[t]);
            matches_so_far += total_t;
            if (tile + 1 == num_tiles) total[0] = matches_so_far;
            continue;
        }
        for (int t = 0; t < ROWS; t++) {
            ProgView pv = view(t);
            PROG fw;
            fw.init(pv);
            cnt[t] = phase_b<NW>(sp, pv, fw, ts, t, any_open);
        }
        for (int t = 0; t < ROWS; t++) { off[t] = total_t; total_t += cnt[t]; }
        const unsigned long long b0 = matches_so_far;
        matches_so_far += total_t;
        if (tile + 1 == num_tiles) total[0] = matches_so_far;
        if (total_t) {
            uint32_t hit[ROWS][4];
            for (int t = 0; t < ROWS; t++) for (int q = 0; q < 4; q++) hit[t][q] = ts.hits[q * ROWS + t];
            if (total_t <= (uint32_t)TS::LIST_CAP) {
                for (int t = 0; t < ROWS; t++) phase_c1_list<NW>(ts, t, off[t], hit[t]);
                for (int t = 0; t < ROWS; t++) {
                    ProgView pv = view(t);
                    PROG fw;
                    fw.init(pv);
                    phase_c2_balanced<NW>(sp, pv, fw, ts, t, total_t, b0);
                }
            } else {
                for (int t = 0; t < ROWS; t++) {
                    ProgView pv = view(t);
                    PROG fw;
                    fw.init(pv);
                    phase_c2_rows<NW>(sp, pv, fw, ts, t, off[t], hit[t], b0);
                }
            }
        }
    }
    // (tile, rank in tile) -> index in the ordered list, as rrx_scanner_copy_pending does
    for (unsigned long long i = 0; i < total[1] && i < pending_cap; i++) {
        const uint64_t tl = pending[i].match_index >> 32, rank = pending[i].match_index & 0xffffffffull;
        pending[i].match_index = tile_base[tl] + rank;
    }
    counts[0] = total[0];
    counts[1] = total[1];
    return 0;
}

}  // namespace

extern "C" {

// text_with_ctx points at the first byte of the left context; the shard starts left_ctx bytes in.
// reduce: bit 0 = merge equivalent positions, bit 1 = forbid the thread-owned candidate scan (rrx_prog.h donly).
// fast != 0: use the register-resident forward program when the automaton allows it (as the GPU does).
// Returns 0, or <0 on a parse error.  counts[0] = matches, counts[1] = pending start scans.
int emu_scan(const char* pattern, int reduce, int fast, const uint8_t* text_with_ctx, uint64_t left_ctx, uint64_t len,
             int64_t base, const uint32_t* entry, int64_t* out, uint64_t cap, uint64_t* counts, uint32_t* err,
             uint32_t* shard_comp, uint32_t* info /* NW, P, n_terms, n_bicl, use_rows, static_depth, sparse_ok, cand_byte, donly */,
             void* pending, uint32_t pending_cap) {
    Program prog;
    try {
        Automaton a = compile_ere(pattern);
        prog = build_program(a, pattern, (reduce & 1) != 0, (reduce & 2) == 0);
    } catch (const std::runtime_error&) {
        return -1;
    }
    const ProgHeader& h = prog.hdr();
    if (info) { info[0] = h.NW; info[1] = h.P; info[2] = h.n_terms; info[3] = h.n_bicl; info[4] = h.use_rows; info[5] = h.static_depth; info[6] = h.sparse_ok; info[7] = h.cand_byte; info[8] = h.donly; info[9] = h.CW; }
    const uint8_t* text = text_with_ctx + left_ctx;
    PendingStart* pend = reinterpret_cast<PendingStart*>(pending);
#define GO(NW_, NT_, NB_) return run<NW_, NW_, Prog<NW_, NT_, NB_>>(prog, text, left_ctx, len, base, entry, out, cap, counts, err, shard_comp, pend, pending_cap)
#define GON(NW_, CW_) return run<NW_, CW_, ProgNarrow<NW_, CW_>>(prog, text, left_ctx, len, base, entry, out, cap, counts, err, shard_comp, pend, pending_cap)
    if (h.NW == 1) {
        if (fast && !h.use_rows && !h.use_rrows) {
            const uint32_t nt = h.n_terms > h.n_rterms ? h.n_terms : h.n_rterms;
            const uint32_t nb = h.n_bicl > h.n_rbicl ? h.n_bicl : h.n_rbicl;
            if (nt <= 1 && nb <= 0) GO(1, 1, 0);
            if (nt <= 2 && nb <= 0) GO(1, 2, 0);
            if (nt <= 3 && nb <= 0) GO(1, 3, 0);
            if (nt <= 4 && nb <= 0) GO(1, 4, 0);
            if (nt <= 1 && nb <= 1) GO(1, 1, 1);
            if (nt <= 2 && nb <= 1) GO(1, 2, 1);
            if (nt <= 3 && nb <= 1) GO(1, 3, 1);
            if (nt <= 4 && nb <= 1) GO(1, 4, 1);
            if (nt <= 2 && nb <= 2) GO(1, 2, 2);
            if (nt <= 4 && nb <= 2) GO(1, 4, 2);
        }
        GO(1, -1, 0);
    }
    // `fast` on a thread-owned multi-word program: carry CW words only, as the run-time specialised kernels do
    if (fast && h.donly && h.CW < h.NW) {
        if (h.CW == 1) switch (h.NW) { case 2: GON(2, 1); case 4: GON(4, 1); case 8: GON(8, 1); case 16: GON(16, 1); case 32: GON(32, 1); }
        if (h.CW == 2) switch (h.NW) { case 4: GON(4, 2); case 8: GON(8, 2); case 16: GON(16, 2); case 32: GON(32, 2); }
        if (h.CW == 4) switch (h.NW) { case 8: GON(8, 4); case 16: GON(16, 4); case 32: GON(32, 4); }
    }
    switch (h.NW) {
        case 2: GO(2, -1, 0);
        case 4: GO(4, -1, 0);
        case 8: GO(8, -1, 0);
        case 16: GO(16, -1, 0);
        case 32: GO(32, -1, 0);
    }
    return -2;
}

uint32_t emu_pending_size(void) { return (uint32_t)sizeof(PendingStart); }
void emu_set_sparse(int mode) { g_sparse_mode = mode; }
// rows handled by the candidate-driven scan / by the dense scan since the last call (sparse-capable programs only)
void emu_sparse_stats(unsigned long long* out, int reset) {
    out[0] = g_sparse_rows; out[1] = g_dense_rows; out[2] = g_fast_tiles;
    if (reset) g_sparse_rows = g_dense_rows = g_fast_tiles = 0;
}

}  // extern "C"
