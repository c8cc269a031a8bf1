// Sharded scans: continuing lazy start scans that ran off a shard's left edge (rare; one thread per
// record, everything read from global memory), and the host-side fold of shard summaries.
#include "rrx_scan.h"
#include "rrx_tile.cuh"

namespace rrx {
namespace {

template <int NW>
__global__ void continue_starts_kernel(const uint32_t* prog, const uint8_t* text, long long len, long long left_ctx,
                                       long long base, PendingStart* recs, unsigned long long n, long long* start_out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ProgHeader& h = *reinterpret_cast<const ProgHeader*>(prog);
    const uint32_t* first = prog + h.off_first;
    PendingStart& rec = recs[i];
    uint32_t R[NW], x[NW], Bc[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) R[w] = rec.R[w];
    long long j = rec.next_offset - base;          // record carries a GLOBAL offset (validated by the host)
    long long found = -1;
    for (;;) {
        if (j < -left_ctx) break;
        const uint32_t c = text[j];
#pragma unroll
        for (int w = 0; w < NW; w++) Bc[w] = prog[h.off_B + c * NW + w];
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { x[w] = R[w] & Bc[w]; m |= x[w] & first[w]; }
        if (m) { found = base + j; break; }
        follow_terms<NW>(prog + h.off_rterms, h.n_rterms, prog + h.off_rbicl, h.n_rbicl, h.use_rrows,
                         prog + h.off_rrowsrc, prog + h.off_rrows, x, R);
        j--;
        if (!any_w<NW>(R)) { found = -2; break; }   // no start exists (inconsistent hand-over)
    }
    start_out[i] = found;
    if (found == -1) {
        rec.next_offset = base + j;
#pragma unroll
        for (int w = 0; w < NW; w++) rec.R[w] = R[w];
    }
}

// Whole-string acceptance (the reference's AcceptanceIterator, regex.h:150-165): text[0:len) is in
// the language iff, walking BACKWARDS from the accepting positions over the unpruned automaton, a
// first-position is alive after consuming byte 0.  One thread, sequential in len like the reference;
// kept for source compatibility of the drop-in API, not a throughput path.
template <int NW>
__global__ void accepts_kernel(const uint32_t* prog, const uint8_t* text, long long len, int* out) {
    const ProgHeader& h = *reinterpret_cast<const ProgHeader*>(prog);
    if (len == 0) { *out = (int)h.nullable; return; }
    uint32_t R[NW], x[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) R[w] = prog[h.off_accept + w];
    for (long long j = len - 1; j >= 0; j--) {
        const uint32_t c = text[j];
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { x[w] = R[w] & prog[h.off_B + c * NW + w]; any |= x[w]; }
        if (!any) { *out = 0; return; }
        if (j == 0) {
            uint32_t m = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) m |= x[w] & prog[h.off_first + w];
            *out = m != 0;
            return;
        }
#pragma unroll
        for (int w = 0; w < NW; w++) R[w] = 0;
        for (int w = 0; w < NW; w++) {
            uint32_t v = x[w];
            while (v) {
                const uint32_t k = (uint32_t)w * 32u + (uint32_t)(__ffs((int)v) - 1);
                v &= v - 1;
                for (int y = 0; y < NW; y++) R[y] |= prog[h.off_rfull + k * NW + y];
            }
        }
    }
}

}  // namespace

cudaError_t launch_accepts(uint32_t NW, const uint32_t* d_prog, const uint8_t* d_text, uint64_t len, int* d_out, cudaStream_t st) {
#define RRX_GO(NW_) accepts_kernel<NW_><<<1, 1, 0, st>>>(d_prog, d_text, (long long)len, d_out)
    switch (NW) {
        case 1: RRX_GO(1); break;
        case 2: RRX_GO(2); break;
        case 4: RRX_GO(4); break;
        case 8: RRX_GO(8); break;
        case 16: RRX_GO(16); break;
        case 32: RRX_GO(32); break;
        default: return cudaErrorInvalidValue;
    }
#undef RRX_GO
    return cudaGetLastError();
}

cudaError_t launch_continue_starts(uint32_t NW, const uint32_t* d_prog, const uint8_t* d_text, uint64_t len, uint64_t left_ctx,
                                   uint64_t base, void* d_recs, uint64_t n, long long* d_start_out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
#define RRX_GO(NW_) continue_starts_kernel<NW_><<<grid, 128, 0, st>>>(d_prog, d_text, (long long)len, (long long)left_ctx, (long long)base, (PendingStart*)d_recs, n, d_start_out)
    switch (NW) {
        case 1: RRX_GO(1); break;
        case 2: RRX_GO(2); break;
        case 4: RRX_GO(4); break;
        case 8: RRX_GO(8); break;
        case 16: RRX_GO(16); break;
        case 32: RRX_GO(32); break;
        default: return cudaErrorInvalidValue;
    }
#undef RRX_GO
    return cudaGetLastError();
}

// ---- host-side composite algebra on flat word arrays [kind, b[W], U[W], P[W]] -------------------------
template <int NW> static void fold_t(const uint32_t* summaries, uint32_t n, uint32_t rank, uint32_t* entry_out, int* exact) {
    Comp<NW> acc, x, nxt;
    comp_ident(acc);
    bool ok = true;
    for (uint32_t r = 0; r < rank && r < n; r++) {
        const uint32_t* s = summaries + (size_t)r * (1 + 3 * NW);
        x.kind = s[0];
        for (int w = 0; w < NW; w++) { x.b[w] = s[1 + w]; x.U[w] = s[1 + NW + w]; x.P[w] = s[1 + 2 * NW + w]; }
        // An opaque summary is only valid for the entry its scan was given; it stays harmless if
        // something closed (entry-independent) follows it, which comp_compose already expresses by
        // dropping the flag.
        comp_compose(acc, x, nxt);
        acc = nxt;
    }
    if (acc.kind & K_OPAQUE) ok = false;
    uint32_t zero[NW];
    for (int w = 0; w < NW; w++) zero[w] = 0;
    comp_apply<NW>(acc, zero, entry_out);
    *exact = ok ? 1 : 0;
}
void fold_entry_host(uint32_t NW, const uint32_t* summaries, uint32_t n, uint32_t rank, uint32_t* entry_out, int* exact) {
    switch (NW) {
        case 1: fold_t<1>(summaries, n, rank, entry_out, exact); break;
        case 2: fold_t<2>(summaries, n, rank, entry_out, exact); break;
        case 4: fold_t<4>(summaries, n, rank, entry_out, exact); break;
        case 8: fold_t<8>(summaries, n, rank, entry_out, exact); break;
        case 16: fold_t<16>(summaries, n, rank, entry_out, exact); break;
        default: fold_t<32>(summaries, n, rank, entry_out, exact); break;
    }
}
template <int NW> static void apply_t(const uint32_t* s, const uint32_t* entry, uint32_t* out) {
    Comp<NW> x;
    x.kind = s[0];
    for (int w = 0; w < NW; w++) { x.b[w] = s[1 + w]; x.U[w] = s[1 + NW + w]; x.P[w] = s[1 + 2 * NW + w]; }
    comp_apply<NW>(x, entry, out);
}
void apply_summary_host(uint32_t NW, const uint32_t* s, const uint32_t* entry, uint32_t* out) {
    switch (NW) {
        case 1: apply_t<1>(s, entry, out); break;
        case 2: apply_t<2>(s, entry, out); break;
        case 4: apply_t<4>(s, entry, out); break;
        case 8: apply_t<8>(s, entry, out); break;
        case 16: apply_t<16>(s, entry, out); break;
        default: apply_t<32>(s, entry, out); break;
    }
}

}  // namespace rrx
