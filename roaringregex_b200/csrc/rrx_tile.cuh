// Per-tile scan logic of the rrx kernels, written as phase functions over an explicit
// shared-memory state so that
//   * rrx_scan.cu runs them on the GPU — a TILE is the 32 rows of ONE WARP, the phases are
//     separated by __syncwarp() only, there is no block-level barrier on the scan path — and
//   * tests/emu/rrx_emu.cc compiles the SAME functions for the host (-DRRX_HOST_EMU) and steps the
//     phases lane by lane.  The emulator is test infrastructure only: it lets the CPU-only CI
//     exercise the chunk-stitching logic (summaries, look-back, fix-up, start scans) bit for bit;
//     librrx.so neither contains nor calls it.
//
// Hot path replaced (reference file:line):
//   NFA.cc:72-102   Processor::shift<fwd>      -> step: S' = (follow'(S) | first) & B[c]
//   NFA.cc:103-107  Processor::operator*       -> (S & accept) != 0 / (R & first) != 0
//   regex.h:156-162 AcceptanceIterator loop    -> row scans (one 128-byte row per lane)
//   Iterator.cpp    (empty: the lazy iterator) -> LAZY-v1 enumeration, SURVEY.md Appendix D
//
// Text geometry: the corpus is cut into ROWS of 128 bytes; a tile is 32 consecutive rows (4 KiB).
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#include "rrx_prog.h"

#if defined(__CUDACC__) && !defined(RRX_HOST_EMU)
#define RRX_DEV __device__ __forceinline__
#define RRX_HD __host__ __device__ __forceinline__
// Cold paths (carried-state replay, start walks, open-row analysis, ...).  The per-tile code of 32 de-phased warps has
// to fit the instruction cache: a 5 % larger kernel took sm__icc_request_hit_rate from 96 % to 84 % and cost config 2
// 12 % (profiles/README.md), so cold code is kept to ONE inlined copy per function.  Calling it out of line
// (RRX_NOINLINE_COLD, A/B) is worse still: the ABI forces the match queue and the tile state into local memory.
#ifdef RRX_NOINLINE_COLD
#define RRX_COLD __device__ __noinline__
#else
#define RRX_COLD __device__ __forceinline__
#endif
// warp-uniform loop control inside per-lane phase functions (all 32 lanes call them convergently)
#define RRX_WARP_ANY(x) (__any_sync(0xffffffffu, (x)) != 0)
#else
#define RRX_DEV inline
#define RRX_HD inline
#define RRX_COLD inline
#define RRX_WARP_ANY(x) (x)          // the emulator steps one lane at a time
#endif

// Branch hints: the compiler lays out what it is told is unlikely away from the hot instruction stream, and the
// per-tile code of 32 de-phased warps has to fit the instruction cache (see RRX_COLD).
#if defined(__CUDACC__) || defined(__GNUC__)
#define RRX_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define RRX_UNLIKELY(x) (x)
#endif

#ifdef RRX_HOST_EMU
struct uint4 { uint32_t x, y, z, w; };
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p |= v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
#endif

namespace rrx {

constexpr int ROW_BYTES = 128;
constexpr int ROW_SHIFT = 7;
constexpr int TILE_ROWS = 32;                         // one row per lane
constexpr int TILE_BYTES = TILE_ROWS * ROW_BYTES;     // 4 KiB
// Rows are staged in shared memory 132 bytes apart: lane t reading byte p of ITS row touches bank
// (33t + p/4) mod 32 = (t + p/4) mod 32, so the 32 lanes of a warp stepping through their rows in
// lock-step never collide, for byte, word or any other same-offset access.
// (The run-time specialised thread-owned kernels — and any kernel that stages tiles with bulk copies, which need 16-byte
// aligned destinations — are compiled with RRX_ROW_PITCH 144: 128-bit loads of the same 16-byte chunk of 8 consecutive
// rows are conflict-free, same-offset 32-bit accesses of all 32 lanes 4-way.  The candidate scan reads rows with 128-bit
// loads and bytes at scattered offsets, so it does not care; the dense row scan does, and keeps 132 — as do the
// pre-compiled kernels.)
#ifndef RRX_ROW_PITCH
#define RRX_ROW_PITCH 132
#endif
constexpr int ROW_PITCH = RRX_ROW_PITCH;
constexpr int ROW_GAP = ROW_PITCH - ROW_BYTES;
// NW == 1 byte table: RRX_BT_COPIES private copies, lane l uses copy l % RRX_BT_COPIES (32: conflict-free)
#ifndef RRX_BT_COPIES
#define RRX_BT_COPIES 32
#endif
constexpr int BT_COPIES = RRX_BT_COPIES;

// error bits reported by the kernel
// ERR_START_LOST: a lazy start walk died without finding a start and without reaching the left edge (only
// possible when the entry set handed to the scan was not the exact one)
// ERR_VERIFY_EDGE (soft): in the thread-owned scan a verification walk reached the left edge of the readable
// text still alive; the scan treated it as "nothing before the text".  Right for a whole text or its first shard;
// a caller that knows there IS text further left (sharded scans) must widen the halo or rescan without DONLY.
// ERR_VERIFY_BUDGET (hard): a verification walk exceeded its step budget; the scan gave up (results invalid),
// the host rescans with the row-bounded candidate scan.
enum : uint32_t { ERR_PENDING_OVERFLOW = 2u, ERR_STAGING_OVERFLOW = 4u, ERR_START_LOST = 8u, ERR_VERIFY_EDGE = 16u, ERR_VERIFY_BUDGET = 32u };

// Kinds of chunk transfer functions S_out = A*S_in | b:
//   K_CLOSED  A = 0: the exit does not depend on the entry
//   K_RANK1   A = U x P^T: exit = b | (entry meets P ? U : 0)
//   K_IDENT   empty chunk
//   K_HARD    (rows only) survivors reach the row end without ever collapsing to one position:
//             A is not representable here; such a row is replayed from its exact entry instead
//   K_WAIT    (tile aggregates only) the tile could not summarise itself: successors wait for
//             its inclusive state
// K_OPAQUE marks a closed composite whose b was computed for the entry set actually used (it is
// not valid as a function of other entries); only inclusive composites can carry it.
enum : uint32_t { K_CLOSED = 0, K_RANK1 = 1, K_IDENT = 2, K_HARD = 3, K_WAIT = 4, K_TYPE = 0xffu, K_OPAQUE = 0x100u };

#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
RRX_DEV uint32_t rotl32(uint32_t x, uint32_t r) { return __funnelshift_l(x, x, r); }
RRX_DEV uint32_t fshl(uint32_t lo, uint32_t hi, uint32_t r) { return __funnelshift_l(lo, hi, r); }   // (hi:lo << r) >> 32
RRX_DEV uint32_t fshr(uint32_t lo, uint32_t hi, uint32_t r) { return __funnelshift_r(lo, hi, r); }   // (hi:lo >> r) & 0xffffffff
RRX_DEV int popc32(uint32_t x) { return __popc(x); }
RRX_DEV int ctz32(uint32_t x) { return __ffs((int)x) - 1; }
RRX_DEV int clz32(uint32_t x) { return __clz((int)x); }
#else
RRX_DEV uint32_t rotl32(uint32_t x, uint32_t r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }
RRX_DEV uint32_t fshl(uint32_t lo, uint32_t hi, uint32_t r) { r &= 31; return r ? (hi << r) | (lo >> (32 - r)) : hi; }
RRX_DEV uint32_t fshr(uint32_t lo, uint32_t hi, uint32_t r) { r &= 31; return r ? (lo >> r) | (hi << (32 - r)) : lo; }
RRX_DEV int popc32(uint32_t x) { return __builtin_popcount(x); }
RRX_DEV int ctz32(uint32_t x) { return __builtin_ctz(x); }
RRX_DEV int clz32(uint32_t x) { return __builtin_clz(x); }
#endif

RRX_DEV void rrx_fence() {
#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
    __threadfence();
#endif
}
RRX_DEV void rrx_backoff() {
#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
    __nanosleep(100);
#endif
}

// ---- small fixed-width set helpers ---------------------------------------------------------------
template <int NW> RRX_DEV bool any_w(const uint32_t* a) {
    uint32_t o = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) o |= a[w];
    return o != 0;
}
template <int NW> RRX_HD bool meets(const uint32_t* a, const uint32_t* b) {
    uint32_t o = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) o |= a[w] & b[w];
    return o != 0;
}
template <int NW> RRX_DEV int popc_w(const uint32_t* a) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) c += popc32(a[w]);
    return c;
}

// ---- program view ----------------------------------------------------------------------------------
// The "small" part of the program blob (everything before the byte table) is copied to shared
// memory once per CTA; the byte table gets its own layout; the per-source row tables stay in
// global memory (only irregular automata use them).
//
// Byte table in shared memory.  NW == 1: one private copy per lane — entry (c, lane) is the
// 32-bit word at byte offset c*128 + lane*4, i.e. always in bank `lane`: conflict-free whatever
// the 32 lanes' bytes are.  NW > 1: entry c is NW consecutive words at c*NW*4.
struct ProgView {
    const uint32_t* sm;        // shared copy of blob[0 .. off_B)
    const uint32_t* glob;      // the blob in global memory
    const uint8_t* Blane;      // shared byte table; NW == 1: already offset by lane*4
    RRX_DEV const ProgHeader& h() const { return *reinterpret_cast<const ProgHeader*>(sm); }
};

// Shared-memory accesses of the scan loop, pinned to the instructions we want: a byte load
// (LSU pipe) instead of word load + PRMT (ALU pipe), and one IMAD (FMA pipe) for the table address.
// The ALU pipe is the bottleneck of the scan loop, the other two have room.
#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
template <int K> RRX_DEV uint32_t lds_u8(const uint8_t* p) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)), "n"(K));
    return v;
}
RRX_DEV uint32_t lds_tab1(const uint8_t* blane, uint32_t c) {
    uint32_t a, v;
    asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(a) : "r"(c), "r"((uint32_t)__cvta_generic_to_shared(blane)), "n"(BT_COPIES * 4));
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// (a | b) & c in one LOP3
RRX_DEV uint32_t or_and(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#else
template <int K> RRX_DEV uint32_t lds_u8(const uint8_t* p) { return p[K]; }
RRX_DEV uint32_t lds_tab1(const uint8_t* blane, uint32_t c) { return *reinterpret_cast<const uint32_t*>(blane + c * (uint32_t)(BT_COPIES * 4)); }
RRX_DEV uint32_t or_and(uint32_t a, uint32_t b, uint32_t c) { return (a | b) & c; }
#endif

// Byte reads at a run-time offset from a shared-memory row: the 32-bit shared address is formed once
// (a generic pointer would have it re-derived at every access).
struct SmemBytes {
#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
    uint32_t base;
    RRX_DEV explicit SmemBytes(const uint8_t* p) {
        // the volatile move keeps the compiler from re-deriving the address inside loops
        asm volatile("mov.u32 %0, %1;" : "=r"(base) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    }
    RRX_DEV uint32_t at(uint32_t i) const {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + i));
        return v;
    }
#else
    const uint8_t* base;
    explicit SmemBytes(const uint8_t* p) : base(p) {}
    uint32_t at(uint32_t i) const { return base[i]; }
#endif
};

template <int NW> RRX_DEV void load_B(const ProgView& pv, uint32_t c, uint32_t* out) {
    if (NW == 1) {
        out[0] = lds_tab1(pv.Blane, c);
    } else {
        const uint32_t* e = reinterpret_cast<const uint32_t*>(pv.Blane) + c * NW;
#if defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
        if (NW == 2) {                       // entries are NW*4 bytes and the table is 16-byte aligned: vector loads
            const uint2 v = *reinterpret_cast<const uint2*>(e);
            out[0] = v.x; out[1] = v.y;
        } else {
#pragma unroll
            for (int w = 0; w < NW; w += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(e + w);
                out[w] = v.x; out[w + 1] = v.y; out[w + 2] = v.z; out[w + 3] = v.w;
            }
        }
#else
#pragma unroll
        for (int w = 0; w < NW; w++) out[w] = e[w];
#endif
    }
}
// f = follow'(S) through a decomposition (shift terms, bicliques, rows) read from shared memory.
template <int NW>
RRX_DEV void follow_terms(const uint32_t* terms, uint32_t n_terms, const uint32_t* bicl, uint32_t n_bicl,
                          uint32_t use_rows, const uint32_t* rowsrc, const uint32_t* rows_g, const uint32_t* S,
                          uint32_t* f) {
#pragma unroll
    for (int w = 0; w < NW; w++) f[w] = 0;
    for (uint32_t t = 0; t < n_terms; t++) {
        const uint32_t* tp = terms + t * (1 + NW);
        const uint32_t kind = tp[0] >> 8, r = tp[0] & 0xff;
        if (NW == 1 || kind == TERM_ROT) {
            f[0] |= rotl32(S[0], r) & tp[1];
        } else if (kind == TERM_SELF) {
#pragma unroll
            for (int w = 0; w < NW; w++) f[w] |= S[w] & tp[1 + w];
        } else if (kind == TERM_SHL) {
#pragma unroll
            for (int w = 0; w < NW; w++) f[w] |= fshl(w ? S[w - 1] : 0u, S[w], r) & tp[1 + w];
        } else {
#pragma unroll
            for (int w = 0; w < NW; w++) f[w] |= fshr(S[w], w + 1 < NW ? S[w + 1] : 0u, r) & tp[1 + w];
        }
    }
    for (uint32_t b = 0; b < n_bicl; b++) {
        const uint32_t* bp = bicl + b * 2 * NW;
        if (meets<NW>(S, bp)) {
#pragma unroll
            for (int w = 0; w < NW; w++) f[w] |= bp[NW + w];
        }
    }
    if (use_rows) {
#pragma unroll 1
        for (int w = 0; w < NW; w++) {
            uint32_t v = S[w] & rowsrc[w];
            while (v) {
                const uint32_t k = (uint32_t)w * 32u + (uint32_t)ctz32(v);
                v &= v - 1;
                const uint32_t* rp = rows_g + k * NW;
#pragma unroll
                for (int x = 0; x < NW; x++) f[x] |= rp[x];
            }
        }
    }
}

// Parameters of the candidate-driven row scan (rrx_prog.h: cand_byte, D, groups) read from the
// shared copy of the program; the run-time specialised policy (rrx_jit.cu) overrides these with
// immediates.
template <int NW> struct SparseFromSmem {
    RRX_DEV bool donly(const ProgView& pv) const { return pv.h().donly != 0u; }
    RRX_DEV uint32_t sp_cand(const ProgView& pv) const { return pv.h().cand_byte; }
    RRX_DEV uint32_t sp_ngroups(const ProgView& pv) const { return pv.h().n_groups; }
    RRX_DEV uint32_t sp_D(const ProgView& pv, int w) const { return pv.sm[pv.h().off_sparse + w]; }
    RRX_DEV uint32_t sp_gk(const ProgView& pv, uint32_t g, int w) const { return pv.sm[pv.h().off_sparse + NW + g * (2 * NW + 1) + w]; }
    RRX_DEV uint32_t sp_gpred(const ProgView& pv, uint32_t g, int w) const { return pv.sm[pv.h().off_sparse + NW + g * (2 * NW + 1) + NW + w]; }
    RRX_DEV uint32_t sp_gflags(const ProgView& pv, uint32_t g) const { return pv.sm[pv.h().off_sparse + NW + g * (2 * NW + 1) + 2 * NW]; }
    // reverse-walk skip state (rrx_prog.h)
    RRX_DEV uint32_t rs_n(const ProgView& pv) const { return pv.h().rskip_n; }
    RRX_DEV uint32_t rs_bytes(const ProgView& pv) const { return pv.h().rskip_bytes; }
    RRX_DEV uint32_t rs_F(const ProgView& pv, int w) const { return pv.sm[pv.h().off_rskip + w]; }
    // forward skip state of the sparse row scan (rrx_prog.h)
    RRX_DEV uint32_t fs_n(const ProgView& pv) const { return pv.h().fskip_n; }
    RRX_DEV uint32_t fs_lo(const ProgView& pv) const { return pv.h().fskip_lo; }
    RRX_DEV uint32_t fs_hi(const ProgView& pv) const { return pv.h().fskip_hi; }
    RRX_DEV uint32_t fs_G(const ProgView& pv, int w) const { return pv.sm[pv.h().off_fskip + w]; }
};

// Follow policy.  NT >= 0: one-word automaton whose forward AND reverse decompositions (up to NT
// rotations, NB bicliques each) live in registers; NT < 0: anything, read from shared memory.
template <int NW, int NT, int NB> struct Prog : SparseFromSmem<NW> {
    static_assert(NW == 1, "register program is for one-word sets");
    uint32_t rot[NT > 0 ? NT : 1], msk[NT > 0 ? NT : 1], bsrc[NB > 0 ? NB : 1], bdst[NB > 0 ? NB : 1];
    uint32_t rrot[NT > 0 ? NT : 1], rmsk[NT > 0 ? NT : 1], rbsrc[NB > 0 ? NB : 1], rbdst[NB > 0 ? NB : 1];
    uint32_t first_, accept_, all_;
    RRX_DEV void init(const ProgView& pv) {
        const ProgHeader& h = pv.h();
        first_ = pv.sm[h.off_first];
        accept_ = pv.sm[h.off_accept];
        all_ = pv.sm[h.off_all];
#pragma unroll
        for (int t = 0; t < NT; t++) {
            const bool live = (uint32_t)t < h.n_terms, rlive = (uint32_t)t < h.n_rterms;
            rot[t] = live ? (pv.sm[h.off_terms + t * 2] & 0xff) : 0u;
            msk[t] = live ? pv.sm[h.off_terms + t * 2 + 1] : 0u;
            rrot[t] = rlive ? (pv.sm[h.off_rterms + t * 2] & 0xff) : 0u;
            rmsk[t] = rlive ? pv.sm[h.off_rterms + t * 2 + 1] : 0u;
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const bool live = (uint32_t)b < h.n_bicl, rlive = (uint32_t)b < h.n_rbicl;
            bsrc[b] = live ? pv.sm[h.off_bicl + b * 2] : 0u;
            bdst[b] = live ? pv.sm[h.off_bicl + b * 2 + 1] : 0u;
            rbsrc[b] = rlive ? pv.sm[h.off_rbicl + b * 2] : 0u;
            rbdst[b] = rlive ? pv.sm[h.off_rbicl + b * 2 + 1] : 0u;
        }
    }
    RRX_DEV uint32_t first(const ProgView&, int) const { return first_; }
    RRX_DEV uint32_t accept(const ProgView&, int) const { return accept_; }
    RRX_DEV uint32_t all(const ProgView&, int) const { return all_; }
    RRX_DEV void follow(const ProgView&, const uint32_t* S, uint32_t* f) const {
        uint32_t o = 0;
#pragma unroll
        for (int t = 0; t < NT; t++) o |= rotl32(S[0], rot[t]) & msk[t];
#pragma unroll
        for (int b = 0; b < NB; b++) o |= (S[0] & bsrc[b]) ? bdst[b] : 0u;
        f[0] = o;
    }
    RRX_DEV void rfollow(const ProgView&, const uint32_t* S, uint32_t* f) const {
        uint32_t o = 0;
#pragma unroll
        for (int t = 0; t < NT; t++) o |= rotl32(S[0], rrot[t]) & rmsk[t];
#pragma unroll
        for (int b = 0; b < NB; b++) o |= (S[0] & rbsrc[b]) ? rbdst[b] : 0u;
        f[0] = o;
    }
};
template <int NW, int NB> struct Prog<NW, -1, NB> : SparseFromSmem<NW> {
    RRX_DEV void init(const ProgView&) {}
    RRX_DEV uint32_t first(const ProgView& pv, int w) const { return pv.sm[pv.h().off_first + w]; }
    RRX_DEV uint32_t accept(const ProgView& pv, int w) const { return pv.sm[pv.h().off_accept + w]; }
    RRX_DEV uint32_t all(const ProgView& pv, int w) const { return pv.sm[pv.h().off_all + w]; }
    RRX_DEV void follow(const ProgView& pv, const uint32_t* S, uint32_t* f) const {
        const ProgHeader& h = pv.h();
        follow_terms<NW>(pv.sm + h.off_terms, h.n_terms, pv.sm + h.off_bicl, h.n_bicl, h.use_rows,
                         pv.sm + h.off_rowsrc, pv.glob + h.off_rows, S, f);
    }
    RRX_DEV void rfollow(const ProgView& pv, const uint32_t* S, uint32_t* f) const {
        const ProgHeader& h = pv.h();
        follow_terms<NW>(pv.sm + h.off_rterms, h.n_rterms, pv.sm + h.off_rbicl, h.n_rbicl, h.use_rrows,
                         pv.sm + h.off_rrowsrc, pv.glob + h.off_rrows, S, f);
    }
};

// ---- composites (rank<=1 chunk transfer functions) ---------------------------------------------------
template <int NW> struct Comp {
    uint32_t kind;
    uint32_t b[NW], U[NW], P[NW];   // exit = b | (entry meets P ? U : 0)   [K_RANK1]
};
template <int NW> RRX_HD void comp_ident(Comp<NW>& c) {
    c.kind = K_IDENT;
#pragma unroll
    for (int w = 0; w < NW; w++) c.b[w] = c.U[w] = c.P[w] = 0;
}
// r = "x first, then y".  Neither is K_HARD/K_WAIT here.
template <int NW> RRX_HD void comp_compose(const Comp<NW>& x, const Comp<NW>& y, Comp<NW>& r) {
    const uint32_t xt = x.kind & K_TYPE, yt = y.kind & K_TYPE;
    if (yt == K_CLOSED || xt == K_IDENT) { r = y; return; }
    if (yt == K_IDENT) { r = x; return; }
    // y is rank 1, x is closed or rank 1
    const bool hit_b = meets<NW>(x.b, y.P);
    const bool hit_U = xt == K_RANK1 && meets<NW>(x.U, y.P);
#pragma unroll
    for (int w = 0; w < NW; w++) {
        r.b[w] = y.b[w] | (hit_b ? y.U[w] : 0u);
        r.U[w] = hit_U ? y.U[w] : 0u;
        r.P[w] = hit_U ? x.P[w] : 0u;
    }
    r.kind = (hit_U ? K_RANK1 : K_CLOSED) | (x.kind & K_OPAQUE);
}
template <int NW> RRX_HD void comp_apply(const Comp<NW>& c, const uint32_t* entry, uint32_t* out) {
    const uint32_t ct = c.kind & K_TYPE;
    if (ct == K_IDENT) {
#pragma unroll
        for (int w = 0; w < NW; w++) out[w] = entry[w];
        return;
    }
    const bool hit = ct == K_RANK1 && meets<NW>(entry, c.P);
#pragma unroll
    for (int w = 0; w < NW; w++) out[w] = c.b[w] | (hit ? c.U[w] : 0u);
}
template <int NW> RRX_HD void comp_closed(Comp<NW>& c, const uint32_t* b, uint32_t flags) {
    c.kind = K_CLOSED | flags;
#pragma unroll
    for (int w = 0; w < NW; w++) { c.b[w] = b[w]; c.U[w] = 0; c.P[w] = 0; }
}

// ---- scan parameters (kernel argument) ------------------------------------------------------------
constexpr int RRX_NW_MAX = 32;
struct PendingStart {           // a start scan that ran off the left edge of the available text
    uint64_t match_index;       // (tile << 32) | rank of the match inside its tile; the host turns it into the
                                // index in the ordered list (rrx_scanner_copy_pending)
    int64_t next_offset;        // offset of the next byte to consume: shard-local (negative) as the kernel writes it,
                                // GLOBAL once rrx_scanner_copy_pending has handed it out (what rrx_continue_starts expects)
    uint32_t R[RRX_NW_MAX];     // backward set before consuming that byte
};
struct ScanParams {
    const uint8_t* text;        // shard bytes; text[-left_ctx .. len) is readable
    uint64_t len;
    uint64_t left_ctx;
    int64_t base;               // global offset of text[0]
    const uint32_t* prog;       // program blob (global memory)
    uint32_t num_tiles;
    uint32_t static_closed;     // 1: every row is closed by construction (acyclic E', depth < ROW_BYTES)
    uint32_t stagger_ns;        // warps of a CTA start (warp%8)*stagger_ns apart, so that they are not all in the same phase
    uint32_t backoff_ns;
    uint32_t sparse;            // 1: try the candidate-driven row scan first (program has a cut; see phase_a_exit/phase_a_cand)
    uint32_t ticket_ahead;      // when the next tile's ticket is taken: 0 after the tile, 1 once the tile has published early, 2 at its start
    unsigned long long* phase_cycles;   // optional [8]: per-phase warp cycles (RRX_PHASE_TIMING builds)
    uint32_t* ticket;
    // Two-level decoupled look-back (tiles, blocks of 32 tiles), see rrx_scan.cu.
    // state composites: payload arrays of 1+3*NW words per entry, ready flags apart (zeroed per scan)
    uint32_t* st_a0;            // [tiles]  tile aggregates
    uint32_t* st_i0;            // [tiles]  tile inclusive composites (relative to the shard entry)
    uint32_t* st_a1;            // [blocks] block aggregates
    uint32_t* st_i1;            // [blocks] block inclusive composites
    volatile uint32_t* f_a0;    // [tiles]
    volatile uint32_t* f_i0;    // [tiles]
    volatile uint32_t* f_a1;    // [blocks]
    volatile uint32_t* f_i1;    // [blocks]
    uint32_t* st_arrive;        // [blocks] tiles of the block that have published their aggregate
    // Output placement: a tile reserves a segment of the staging list with one atomicAdd (tiles land
    // in completion order, matches inside a tile in text order) and records it; the ordering pass
    // (rrx_order.cu) then moves the segments into tile order.  Nothing waits for match counts.
    unsigned long long* tile_seg;   // [tiles] first staging slot of the tile's matches
    uint32_t* tile_cnt;             // [tiles] matches of the tile
    long long* out;             // staging list of (start,end) pairs
    uint64_t out_cap;
    unsigned long long* total;  // [0] staging slots handed out by the allocator, [1] pending count
    // Staging slots.  A tile with at most seg_stride matches writes them at slot tile*seg_stride (no
    // atomic, no waiting); a fuller tile takes a segment of the overflow area [ovf_base, out_cap) with one
    // atomicAdd on total[0].  seg_stride == 0: every tile allocates.  The ordering pass compacts.
    uint32_t seg_stride;
    uint32_t donly;             // 1: thread-owned candidate scan (program header `donly`; see phase_a_cand<.., true>)
    // Stop-byte table (programs with a reverse skip state, thread-owned scan): tile_stop[tile] = 0 not known yet,
    // 1 the tile holds none of the skip state's stop bytes, else 2 + tile-relative offset of its LAST one.  A backward
    // walk in the skip state that leaves its tile hops over whole tiles with it (see back_walk).  Zeroed per scan.
    volatile uint32_t* tile_stop;
    // thread-owned scan: tiles that left the fast tail (lean kernel) / rows whose run its lane followed into the next row
    // (straddle kernel) — what the scanner picks its next kernel by
    uint32_t* straddle_tiles;
    unsigned long long ovf_base;
    uint32_t* err;
    uint32_t* shard_comp;       // [1+3*NW] composite of the whole shard (relative to an unknown entry)
    PendingStart* pending;
    uint32_t pending_cap;
    uint32_t entry[RRX_NW_MAX]; // forward set carried into text[0]
};

// ---- shared-memory state of one tile (= one warp) ----------------------------------------------------
// size of TileState<NW, CW> (also evaluated on the host for run-time compiled kernels)
RRX_HD constexpr size_t tile_state_bytes(int NW, int CW, int pitch = ROW_PITCH) {
    return (size_t)TILE_ROWS * pitch +
           sizeof(uint32_t) * ((size_t)3 * CW * TILE_ROWS + 8 * TILE_ROWS + (8 + CW) + (NW <= 2 ? 64 : NW <= 4 ? 32 : 16) * (1 + NW));
}
// NW: words of a state set; CW: words of a CARRIED set (row exits, composites, entry) — CW == NW unless the kernel is
// specialised for a thread-owned program whose dominated positions fit fewer words (rrx_prog.h CW).
template <int NW, int CW = NW> struct TileState {
    static constexpr int ROWS = TILE_ROWS;
    uint8_t* text;        // 32 rows of 128 B, ROW_PITCH bytes apart
    uint32_t* sum_b;      // [CW][32]   exit of the row scanned from the empty set
    uint32_t* sum_U;      // [CW][32]   (open rows) survivors at the row end / prefix composite U
    uint32_t* sum_P;      // [CW][32]   (open rows) entry states that feed the survivor
    uint32_t* sum_kind;   // [32]
    uint32_t* hits;       // [4][32]    one bit per byte of the row: a match ends after that byte
    uint32_t* cnt;        // [32]       matches per row -> exclusive offsets
    uint32_t* misc;       // see M_* below
    uint32_t* rq;         // replay queue: RQ_CAP items of {row<<5 | quad, state[NW]}
    uint32_t* zmask;      // [4][32], aliases rq
    uint16_t* list;       // hit list for the balanced start scans; aliases sum_b/sum_U/sum_P
    uint32_t* kst;        // [2][32]    known lazy starts of up to two matches per row: (end-1)<<16 | (start - row start + KS_BIAS)
    uint64_t row0;        // first row of the tile
    static constexpr int M_RQCNT = 0, M_ROWHAS = 1, M_ZVALID = 2, M_ENTRY = 8;
    // Stop-byte masks of the staged tile (zmask[q*32 + t]: bit p <=> byte 32q+p of row t is one of the reverse skip
    // state's stop bytes or >= 0x80), M_ROWHAS: rows that have any, M_ZVALID: masks valid for this tile.
    // They alias the replay queue, which only the dense row scan uses.
    static constexpr int MISC_WORDS = 8 + CW;
    static constexpr int RQ_CAP = NW <= 2 ? 64 : NW <= 4 ? 32 : 16;         // (>= 512 bytes: the stop-byte masks alias it)
    static constexpr int LIST_CAP = 3 * CW * ROWS * 2;
    static constexpr size_t bytes() { return tile_state_bytes(NW, CW); }
    RRX_DEV void carve(uint8_t* base) {
        text = base;
        uint32_t* p = reinterpret_cast<uint32_t*>(base + (size_t)ROWS * ROW_PITCH);
        sum_b = p; p += CW * ROWS;
        sum_U = p; p += CW * ROWS;
        sum_P = p; p += CW * ROWS;
        sum_kind = p; p += ROWS;
        hits = p; p += 4 * ROWS;
        cnt = p; p += ROWS;
        misc = p; p += MISC_WORDS;
        kst = p; p += 2 * ROWS;
        rq = p;
        zmask = p;
        list = reinterpret_cast<uint16_t*>(sum_b);
    }
    RRX_DEV int64_t lo() const { return (int64_t)(row0 << ROW_SHIFT); }
    // byte at tile-relative offset rel (0 .. TILE_BYTES)
    RRX_DEV uint32_t rel_byte(uint32_t rel) const { return text[rel + (uint32_t)ROW_GAP * (rel >> ROW_SHIFT)]; }
    // where the q-th 16-byte chunk of the tile goes (4-byte aligned)
    RRX_DEV uint32_t* chunk_ptr(int q) const {
        return reinterpret_cast<uint32_t*>(text + (q >> 3) * ROW_PITCH + ((q & 7) << 4));
    }
    RRX_DEV uint32_t row_byte(int t, int p) const { return rel_byte(((uint32_t)t << ROW_SHIFT) + (uint32_t)p); }
};

// Number of valid bytes of row `t` of the tile starting at row0.
RRX_DEV int row_nbytes(const ScanParams& sp, uint64_t row0, int t) {
    const uint64_t lo = (row0 + (uint64_t)t) << ROW_SHIFT;
    if (lo >= sp.len) return 0;
    const uint64_t rem = sp.len - lo;
    return rem >= ROW_BYTES ? ROW_BYTES : (int)rem;
}

// the same, clamped to TWO rows: 2 * ROW_BYTES <=> rows t and t+1 are both full rows of the text
RRX_DEV uint32_t row_nbytes2(const ScanParams& sp, uint64_t row0, int t) {
    const uint64_t lo = (row0 + (uint64_t)t) << ROW_SHIFT;
    if (lo >= sp.len) return 0u;
    const uint64_t rem = sp.len - lo;
    return rem >= 2u * ROW_BYTES ? 2u * ROW_BYTES : (uint32_t)rem;
}

// Known lazy starts travel as 16-bit values: start offset relative to the row start, biased by KS_BIAS (a start
// found by a verification walk may lie in an earlier row or tile); 0 = not known.
constexpr uint32_t KST_NONE = 0xffffffffu;
constexpr int KS_BIAS = 32768;
constexpr int VERIFY_BUDGET = 1 << 16;      // steps one verification walk may take outside its row

// ===== phase A: per-row scan from the empty set (+ openness analysis) ===============================
template <int NW, class F, int CW>
RRX_DEV void phase_a(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int t) {
    constexpr int ROWS = TILE_ROWS;
    using TS = TileState<NW, CW>;
    const int nb = row_nbytes(sp, ts.row0, t);
    uint32_t kind = nb ? K_CLOSED : K_IDENT;
    uint32_t Bc[NW], f[NW];

    // (1) survivor scan: U = all positions, no injection.  Dead before the row ends => the row's
    //     exit does not depend on its entry (closed).  Otherwise remember the first point where
    //     exactly one survivor is left: from there on the row acts as a rank-1 map.
    // (static_closed: a FULL row outlasts every path of an acyclic automaton of depth < ROW_BYTES; a partial
    // last row does not, so it is analysed like any other row)
    if (nb && !(sp.static_closed && nb == ROW_BYTES)) {
        uint32_t U[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) U[w] = fw.all(pv, w);
        int pc = -1;
        uint32_t J[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) J[w] = 0;
        int p = 0;
        for (; p < nb; p++) {
            load_B<NW>(pv, ts.row_byte(t, p), Bc);
            fw.follow(pv, U, f);
#pragma unroll
            for (int w = 0; w < NW; w++) U[w] = f[w] & Bc[w];
            if (!any_w<NW>(U)) break;
            if (pc < 0 && popc_w<NW>(U) == 1) {
                pc = p + 1;
#pragma unroll
                for (int w = 0; w < NW; w++) J[w] = U[w];
            }
        }
        if (p >= nb) {                       // survivors reach the end of the row: open
            kind = pc < 0 ? K_HARD : K_RANK1;
            uint32_t R[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) R[w] = J[w];
            if (pc >= 0) {
                // entry states that lead to the lone survivor: walk the prefix backwards
                for (int q = pc - 1; q >= 0; q--) {
                    load_B<NW>(pv, ts.row_byte(t, q), Bc);
                    uint32_t x[NW];
#pragma unroll
                    for (int w = 0; w < NW; w++) x[w] = R[w] & Bc[w];
                    fw.rfollow(pv, x, R);
                }
            }
#pragma unroll
            for (int w = 0; w < NW; w++) {
                if (w < CW) {
                    ts.sum_U[w * ROWS + t] = U[w];
                    ts.sum_P[w * ROWS + t] = R[w];
                }
            }
        }
    }

    // (2) the scan proper, from the empty set, with re-injection; record ends of matches.
    uint32_t S[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) S[w] = 0;
    uint32_t hit[4] = {0, 0, 0, 0};
    if (nb == ROW_BYTES) {
        const uint8_t* rp = ts.text + t * ROW_PITCH;
#pragma unroll 1
        for (int g = 0; g < 8; g++) {
            const uint8_t* gp = rp + g * 16;
            uint32_t Sq[4][NW], acc[4];
#define RRX_STEP(K_)                                                              \
            {                                                                          \
                load_B<NW>(pv, lds_u8<K_>(gp), Bc);                                    \
                fw.follow(pv, S, f);                                                   \
                _Pragma("unroll") for (int w = 0; w < NW; w++) {                       \
                    S[w] = or_and(f[w], fw.first(pv, w), Bc[w]);                       \
                    a |= S[w] & fw.accept(pv, w);                                      \
                }                                                                      \
            }
#define RRX_QUAD(Q_)                                                              \
            {                                                                          \
                _Pragma("unroll") for (int w = 0; w < NW; w++) Sq[Q_][w] = S[w];       \
                uint32_t a = 0;                                                        \
                RRX_STEP(Q_ * 4 + 0) RRX_STEP(Q_ * 4 + 1) RRX_STEP(Q_ * 4 + 2) RRX_STEP(Q_ * 4 + 3) \
                acc[Q_] = a;                                                           \
            }
            RRX_QUAD(0) RRX_QUAD(1) RRX_QUAD(2) RRX_QUAD(3)
#undef RRX_QUAD
#undef RRX_STEP
            if (acc[0] | acc[1] | acc[2] | acc[3]) {
                // A match ends somewhere in these 16 bytes.  Finding the exact byte means replaying
                // a 4-byte word from its saved start state; doing that here would make the whole
                // warp wait for the one or two lanes concerned (and that happens in most groups as
                // soon as matches are a few hundred bytes apart).  Queue the word instead: after the
                // row loop the warp replays all queued words of the tile together, one per lane.
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (acc[q]) {
                        const uint32_t slot = atomicAdd(ts.misc + TS::M_RQCNT, 1u);
                        if (slot < (uint32_t)TS::RQ_CAP) {
                            uint32_t* it = ts.rq + slot * (1 + NW);
                            it[0] = ((uint32_t)t << 5) | (uint32_t)(g * 4 + q);
#pragma unroll
                            for (int w = 0; w < NW; w++) it[1 + w] = Sq[q][w];
                        } else {            // queue full: replay on the spot
                            uint32_t T[NW];
#pragma unroll
                            for (int w = 0; w < NW; w++) T[w] = Sq[q][w];
#pragma unroll 1
                            for (int i = 0; i < 4; i++) {
                                load_B<NW>(pv, gp[q * 4 + i], Bc);
                                fw.follow(pv, T, f);
                                uint32_t a2 = 0;
#pragma unroll
                                for (int w = 0; w < NW; w++) {
                                    T[w] = (f[w] | fw.first(pv, w)) & Bc[w];
                                    a2 |= T[w] & fw.accept(pv, w);
                                }
                                if (a2) {
                                    const int p = g * 16 + q * 4 + i;
                                    if (p < 32) hit[0] |= 1u << p; else if (p < 64) hit[1] |= 1u << (p - 32);
                                    else if (p < 96) hit[2] |= 1u << (p - 64); else hit[3] |= 1u << (p - 96);
                                }
                            }
                        }
                    }
                }
            }
        }
    } else {
        for (int p = 0; p < nb; p++) {
            load_B<NW>(pv, ts.row_byte(t, p), Bc);
            fw.follow(pv, S, f);
            uint32_t a2 = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                S[w] = (f[w] | fw.first(pv, w)) & Bc[w];
                a2 |= S[w] & fw.accept(pv, w);
            }
            if (a2) hit[p >> 5] |= 1u << (p & 31);
        }
    }
#pragma unroll
    for (int w = 0; w < NW; w++) if (w < CW) ts.sum_b[w * ROWS + t] = S[w];
    ts.sum_kind[t] = kind;
#pragma unroll
    for (int q = 0; q < 4; q++) ts.hits[q * ROWS + t] = hit[q];
    ts.kst[2 * t] = ts.kst[2 * t + 1] = KST_NONE;      // no lazy start known in advance
}

// ===== backward walk over the reversed automaton =======================================================
// From the set R (positions that may consume the byte at shard-local offset e1, then e1-1, ...) until a
// first-position consumes a byte: that byte is the right-most (= lazy) start.  Used for the lazy start of a match
// (R = accept, e1 = end-1: phase C) and for the verification of a candidate byte in the thread-owned scan
// (R = predecessors of a cut group, e1 = candidate-1: "was a predecessor active, and since where?").
// Inside the tile the walk reads the staged rows (32-bit tile-relative offsets), before it global memory, down to
// the left edge of the readable text.  While R equals the program's skip state F the walk runs over the text
// word-wise looking only for the bytes that matter.
struct WalkResult {
    int64_t j;          // found: offset of the start byte; pend: offset of the next byte to consume (left of the edge)
    bool found, pend;   // neither: the walk died — no start exists
    bool over;          // the step budget ran out (verification walks only)
};
template <int NW, class F, int CW>
RRX_COLD WalkResult back_walk(const ScanParams& sp, const ProgView& pv, const F& fw, const TileState<NW, CW>& ts, uint32_t* R,
                             int64_t e1, int budget) {
    uint32_t Bc[NW], x[NW];
    const int64_t tile_lo = ts.lo(), left_edge = -(int64_t)sp.left_ctx;
    WalkResult wr;
    wr.found = wr.pend = wr.over = false;
    bool dead = false;
    // Skip state: while R equals F, bytes other than b0/b1 (and below 0x80) neither find a start nor
    // change R, so the walk just runs over them.
    const uint32_t rsn = fw.rs_n(pv), rsb = fw.rs_bytes(pv);
    const uint32_t b0 = rsb & 0xffu, b1 = (rsb >> 8) & 0xffu;
    // fast path: while the walk stays inside the tile (32-bit tile-relative offsets, shared memory)
    int rel = (e1 - tile_lo) < 0 ? -1 : (int)(e1 - tile_lo);
    for (; rel >= 0; rel--) {
        uint32_t c = ts.rel_byte((uint32_t)rel);
        if (rsn) {
            uint32_t d = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) d |= R[w] ^ fw.rs_F(pv, w);
#ifdef RRX_ZMASK_ALWAYS
            if (d == 0) {
#else
            if (d == 0 && ts.misc[TileState<NW, CW>::M_ZVALID]) {
#endif
                // the tile's stop-byte masks answer "last stop byte at or before rel" in a few instructions
                uint32_t row = (uint32_t)rel >> ROW_SHIFT, pos = (uint32_t)rel & (ROW_BYTES - 1), w = pos >> 5;
                uint32_t m = ts.zmask[w * TILE_ROWS + row] & (0xffffffffu >> (31u - (pos & 31u)));
                while (m == 0u && w > 0u) { w--; m = ts.zmask[w * TILE_ROWS + row]; }
                if (m == 0u) {
                    const uint32_t rh = ts.misc[TileState<NW, CW>::M_ROWHAS] & ((1u << row) - 1u);
                    if (rh == 0u) { rel = -1; break; }              // ran to the tile start inside F
                    row = 31u - (uint32_t)clz32(rh);
                    w = 3u;
                    m = ts.zmask[3 * TILE_ROWS + row];
                    while (m == 0u) { w--; m = ts.zmask[w * TILE_ROWS + row]; }
                }
                rel = (int)(row * ROW_BYTES + w * 32u + (31u - (uint32_t)clz32(m)));
                c = ts.rel_byte((uint32_t)rel);
            }
#ifndef RRX_ZMASK_ALWAYS        // (kernels that always have the masks do not carry the word-wise walk)
            else if (d == 0) {
                // bytes down to a word boundary, then whole words (SWAR: does the word hold b0, b1 or a
                // byte >= 0x80?), then the bytes of the word that does
                while (c != b0 && c != b1 && c < 128u && rel > 0 && (rel & 3) != 0) { rel--; c = ts.rel_byte((uint32_t)rel); }
                if (c != b0 && c != b1 && c < 128u && rel >= 4 && (rel & 3) == 0) {
                    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(ts.text);
                    const uint32_t B0 = b0 * 0x01010101u, B1 = b1 * 0x01010101u;
                    int wi = (rel >> 2) - 1;                 // last whole word before byte rel
                    for (; wi >= 0; wi--) {
                        const uint32_t w = t32[wi + (ROW_GAP / 4) * (wi >> 5)];          // ROW_PITCH / 4 words per staged row
                        const uint32_t x0 = w ^ B0, x1 = w ^ B1;
                        const uint32_t z = ((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | w;
                        if (z & 0x80808080u) break;
                    }
                    rel = wi * 4 + 3;                        // last byte of the word that stopped the run (or -1)
                    if (rel < 0) break;                      // ran to the tile start inside F
                    c = ts.rel_byte((uint32_t)rel);
                    while (c != b0 && c != b1 && c < 128u && rel > 0) { rel--; c = ts.rel_byte((uint32_t)rel); }
                } else {
                    while (c != b0 && c != b1 && c < 128u && rel > 0) { rel--; c = ts.rel_byte((uint32_t)rel); }
                }
                if (c != b0 && c != b1 && c < 128u) { rel = -1; break; }      // ran to the tile start inside F
            }
#endif
        }
        load_B<NW>(pv, c, Bc);
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { x[w] = R[w] & Bc[w]; m |= x[w] & fw.first(pv, w); }
        if (m) { wr.found = true; break; }
        fw.rfollow(pv, x, R);
        if (!any_w<NW>(R)) { dead = true; break; }
    }
    int64_t j = (e1 - tile_lo) < 0 ? e1 : tile_lo + rel;
    if (!wr.found && !dead) {
        // slow path: the walk left the tile; continue in global memory
        for (;;) {
            if (j < left_edge) { wr.pend = true; break; }
            uint32_t d = 1;
            if (rsn) {
                d = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) d |= R[w] ^ fw.rs_F(pv, w);
                if (d == 0 && sp.tile_stop && j >= 0 && (j & (TILE_BYTES - 1)) == TILE_BYTES - 1) {
                    // in the skip state on the last byte of an earlier tile of this shard: that tile's entry of the
                    // stop-byte table (written by its warp right after staging; smaller tile index, so it is in
                    // flight or done) says where its last stop byte is, or that it has none
                    const int64_t tt = j >> 12;
                    uint32_t v = sp.tile_stop[tt];
                    while (v == 0u) { rrx_backoff(); v = sp.tile_stop[tt]; }
                    if (v == 1u) {
                        j -= TILE_BYTES;
                        if (budget > 0 && --budget == 0) { wr.over = true; break; }
                        continue;
                    }
                    j = (tt << 12) + (int64_t)(v - 2u);
                }
            }
            uint32_t c = (uint32_t)sp.text[j];
            if (rsn) {
                if (d == 0) {
                    // bytes down to a word boundary, whole words (the text pointer is 16-byte aligned), then bytes
                    while (c != b0 && c != b1 && c < 128u && j > left_edge && (j & 3) != 0) { j--; c = (uint32_t)sp.text[j]; }
                    if (c != b0 && c != b1 && c < 128u && (j & 3) == 0 && j - 4 >= left_edge) {
                        const uint32_t B0 = b0 * 0x01010101u, B1 = b1 * 0x01010101u;
                        int64_t wj = j - 4;
                        for (; wj >= left_edge; wj -= 4) {
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(sp.text + wj);
                            const uint32_t x0 = w ^ B0, x1 = w ^ B1;
                            const uint32_t z = ((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | w;
                            if (z & 0x80808080u) break;
                            if (budget > 0 && --budget == 0) { wr.over = true; break; }
                        }
                        if (wr.over) break;
                        if (wj < left_edge) {
                            j = wj + 3;                  // bytes below the last whole word (if any), one by one
                            if (j < left_edge) { j = left_edge - 1; wr.pend = true; break; }
                        } else {
                            j = wj + 3;
                        }
                        c = (uint32_t)sp.text[j];
                    }
                    while (c != b0 && c != b1 && c < 128u && j > left_edge) { j--; c = (uint32_t)sp.text[j]; }
                    if (c != b0 && c != b1 && c < 128u) { j = left_edge - 1; wr.pend = true; break; }   // still in F at the edge
                }
            }
            load_B<NW>(pv, c, Bc);
            uint32_t m = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) { x[w] = R[w] & Bc[w]; m |= x[w] & fw.first(pv, w); }
            if (m) { wr.found = true; break; }
            fw.rfollow(pv, x, R);
            j--;
            if (!any_w<NW>(R)) break;
            if (budget > 0 && --budget == 0) { wr.over = true; break; }
        }
    }
    wr.j = j;
    return wr;
}

// ===== phase A, candidate-driven ("sparse") variant =================================================
// Same outputs as phase_a + phase_a2 for a CLOSED full row — hits[], sum_b (exit of the row scanned
// from the empty set with re-injection), sum_kind = K_CLOSED — computed without stepping the
// automaton over every byte.  Needs a cut (rrx_prog.h: cand_byte, D, groups).  Two functions:
//
// phase_a_exit — the exit set.  Walk back from the row end with the set of positions that can still
//   lie on a path reaching the row end; where it dies no earlier byte can influence the exit, and a
//   dense forward scan of that suffix gives sum_b.  If it survives to the row start the row is open
//   (returns false).  Runs FIRST, so that a tile whose rows are all closed can publish its
//   aggregate (= the last row's exit) before it looks for matches: successors then find it ready.
//
// phase_a_cand — the match ends.
//   (1) SWAR pass over the row's 32 words: bit mask of the positions holding cand_byte;
//   (2) only the dominated states H = S & D are tracked.  H is empty except after a candidate byte:
//       there a cut position k is injected iff it is a first-position or some predecessor of k was
//       active before the byte, which a backward walk over the reversed automaton decides (exactly the
//       lazy start scan of phase C, bounded by the row start since phase A assumes an empty entry).
//       While H is empty the lane jumps straight to its next candidate.  accept is inside D, so
//       H & accept gives the match ends exactly.
//   Known starts: a run of dominated states fed by ONE injection (no other candidate byte inside
//   the run) can only contain matches that cross the cut at that byte, so their lazy start is the
//   start the backward walk found (or the candidate itself for a first-position).  Up to two
//   (end, start) pairs per row are handed to phase C in kst[], which then skips its own backward scan.
//   Returns false when the row has to go through the dense scan (too many candidates, step budget
//   exhausted); the caller then runs phase_a/phase_a2 for the whole tile.
constexpr int SPARSE_MAX_CAND = 16;
constexpr int SPARSE_BUDGET = 96;

template <int NW, class F, int CW>
RRX_DEV bool phase_a_exit(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int t) {
    constexpr int ROWS = TILE_ROWS;
    if (row_nbytes(sp, ts.row0, t) != ROW_BYTES) return false;
    const uint8_t* rp = ts.text + t * ROW_PITCH;
    uint32_t S[NW], A[NW], x[NW], Bc[NW], f[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) A[w] = fw.all(pv, w);
    int q = ROW_BYTES - 1;
    for (;;) {
        load_B<NW>(pv, rp[q], Bc);
#pragma unroll
        for (int w = 0; w < NW; w++) x[w] = A[w] & Bc[w];
        if (!any_w<NW>(x)) break;          // nothing that consumes byte q reaches the row end
        if (q == 0) return false;          // some path spans the whole row: open
        fw.rfollow(pv, x, A);
        q--;
    }
#pragma unroll
    for (int w = 0; w < NW; w++) S[w] = 0;
    for (int i = q + 1; i < ROW_BYTES; i++) {
        load_B<NW>(pv, rp[i], Bc);
        fw.follow(pv, S, f);
#pragma unroll
        for (int w = 0; w < NW; w++) S[w] = (f[w] | fw.first(pv, w)) & Bc[w];
    }
#pragma unroll
    for (int w = 0; w < NW; w++) if (w < CW) ts.sum_b[w * ROWS + t] = S[w];
    ts.sum_kind[t] = K_CLOSED;
    return true;
}

// What phase_a_cand leaves in registers: up to four match ends of the row (kq: end+1 per byte, most recent in
// the low byte), each tagged with the run it belongs to (ks: tag per byte, 0 = lazy start not known), and the
// lazy starts of the last four tagged runs (rt: 16 bits each, most recent lowest: start - row start + KS_BIAS).
struct SparseTail {
    uint32_t kq, ks, nacc, drained, nruns;
    unsigned long long rt;
    // lazy start (biased, see KS_BIAS) of the run with tag g, 0 if it is not on record any more
    RRX_DEV uint32_t start_of(uint32_t g) const {
        const uint32_t back = nruns - g;
        return (g == 0u || back > 3u) ? 0u : (uint32_t)(rt >> (16u * back)) & 0xffffu;
    }
    // every match of the row is in the queue and has a known start: no zero tag among the nacc entries, and the
    // oldest entry's run (tags never decrease along the queue) is still among the four on record
    RRX_DEV bool simple() const {
        if (drained) return false;
        if (nacc == 0u) return true;
        const uint32_t m = 0xffffffffu >> (32u - 8u * nacc);
        const uint32_t v = ks | ~m;                              // bytes beyond the queue count as tagged
        if ((((v - 0x01010101u) & ~v) & 0x80808080u) != 0u) return false;   // a zero byte
        return nruns - ((ks >> (8u * (nacc - 1u))) & 0xffu) <= 3u;
    }
};

// Fast tail of a tile: the row's matches go from the queue straight to the output list, in text order
// (oldest queue entry first), at slots slot0, slot0+1, ...
template <int NW, int CW>
RRX_DEV void phase_tail_fast(const ScanParams& sp, const TileState<NW, CW>& ts, int t, const SparseTail& tail, unsigned long long slot0) {
    const long long row_lo = sp.base + ts.lo() + ((long long)t << ROW_SHIFT);
    for (uint32_t i = 0; i < tail.nacc; i++) {
        const uint32_t sh = 8u * (tail.nacc - 1u - i);
        const uint32_t e1 = (tail.kq >> sh) & 0xffu, s16 = tail.start_of((tail.ks >> sh) & 0xffu);
        const unsigned long long slot = slot0 + i;
        if (slot < sp.out_cap) {
            long long* o = sp.out + 2 * slot;
            o[0] = row_lo + (long long)s16 - KS_BIAS;
            o[1] = row_lo + (long long)e1;            // e1 = (end offset - 1) + 1
        }
    }
}

// Unpack the match-end queue of phase_a_cand into the row's hit mask (and its first two known starts).
template <int NW, int CW>
RRX_COLD void sparse_drain(TileState<NW, CW>& ts, int t, const SparseTail& q) {
    constexpr int ROWS = TILE_ROWS;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t e1 = (q.kq >> (8 * i)) & 0xffu;
        if (e1 && (uint32_t)i < q.nacc) {
            const uint32_t pos = e1 - 1, s16 = q.start_of((q.ks >> (8 * i)) & 0xffu);
#ifndef RRX_NO_OVERRUN
            if (pos >= (uint32_t)ROW_BYTES) {
                // an end found while overrunning into the next row: it belongs to that row's hit mask (another lane's
                // words: atomic); its start is left to the backward walk
                atomicOr(ts.hits + ((pos - ROW_BYTES) >> 5) * ROWS + t + 1, 1u << (pos & 31));
                continue;
            }
            atomicOr(ts.hits + (pos >> 5) * ROWS + t, 1u << (pos & 31));
#else
            ts.hits[(pos >> 5) * ROWS + t] |= 1u << (pos & 31);
#endif
            if (s16) {
                const uint32_t rec = (pos << 16) | s16;
                if (ts.kst[2 * t] == KST_NONE) ts.kst[2 * t] = rec;
                else if (ts.kst[2 * t + 1] == KST_NONE) ts.kst[2 * t + 1] = rec;
            }
        }
    }
}

// DONLY = false: the row-bounded scan described above (a closed full row, scanned from the empty set).
// DONLY = true:  the THREAD-OWNED scan.  The lane owns every activation thread that crosses the cut at a
//   candidate byte of ITS row, wherever that thread began: the verification walk continues before the row start
//   (earlier rows of the tile, then global memory, to the left edge of the readable text).  Only dominated
//   states are ever carried, so nothing outside D has to be known at the row start: no exit-set scan, no
//   dense fallback, no bail-outs.  The row may be partial (nb < 128 valid bytes; the staged bytes behind them are
//   >= 0x80).  sum_b[t] receives the dominated states alive after the row's last byte that stem from the row's
//   own candidates; carried states are phase B's business as before.
// (1) candidate mask of row t: bit p of cm[q] <=> byte 32q+p == cand
// word j of a staged row (rows that are 16-byte aligned are read with 128-bit loads)
RRX_DEV void row_words8(const uint32_t* rw, int q, uint32_t* w8) {
#if RRX_ROW_PITCH % 16 == 0 && defined(__CUDA_ARCH__) && !defined(RRX_HOST_EMU)
    const uint4 a = *reinterpret_cast<const uint4*>(rw + q * 8), b = *reinterpret_cast<const uint4*>(rw + q * 8 + 4);
    w8[0] = a.x; w8[1] = a.y; w8[2] = a.z; w8[3] = a.w; w8[4] = b.x; w8[5] = b.y; w8[6] = b.z; w8[7] = b.w;
#else
#pragma unroll
    for (int j = 0; j < 8; j++) w8[j] = rw[q * 8 + j];
#endif
}

template <int NW, int CW>
RRX_DEV void phase_a_mask(const TileState<NW, CW>& ts, int t, uint32_t cand, uint32_t* cm) {
    const uint32_t C4 = cand * 0x01010101u;
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(ts.text + t * ROW_PITCH);
    // (RRX_ROLL_MASKS: the quarter loop is NOT unrolled — a quarter of the static code; the per-tile code of 32 de-phased
    // warps has to fit the instruction cache, see RRX_COLD)
#ifdef RRX_ROLL_MASKS
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        uint32_t acc = 0, w8[8];
        row_words8(rw, q, w8);
#pragma unroll
        for (int j = 7; j >= 0; j--) {
            const uint32_t x = w8[j] ^ C4;
            const uint32_t y = (x | 0x80808080u) - 0x01010101u;
            const uint32_t f = ~(y | x) & 0x80808080u;
            acc = fshl(f * 0x00204081u, acc, 4);
        }
        c0 = q == 0 ? acc : c0; c1 = q == 1 ? acc : c1; c2 = q == 2 ? acc : c2; c3 = q == 3 ? acc : c3;
    }
    cm[0] = c0; cm[1] = c1; cm[2] = c2; cm[3] = c3;
#else
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t acc = 0, w8[8];
        row_words8(rw, q, w8);
#pragma unroll
        for (int j = 7; j >= 0; j--) {
            const uint32_t x = w8[j] ^ C4;
            const uint32_t y = (x | 0x80808080u) - 0x01010101u;     // no borrow between bytes
            const uint32_t f = ~(y | x) & 0x80808080u;              // 0x80 where the byte equals cand
            acc = fshl(f * 0x00204081u, acc, 4);                    // the 4 flags land in the top nibble
        }
        cm[q] = acc;
    }
#endif
}

// Stop-byte mask of row t for the reverse skip state (stop bytes b0/b1 of rrx_prog.h rskip, and every byte >= 0x80):
// bit p of zm[q] <=> byte 32q+p stops a backward walk that is in the skip state.  Also left in ts.zmask.
template <int NW, int CW>
RRX_DEV void phase_z_mask(TileState<NW, CW>& ts, int t, uint32_t rs_bytes, uint32_t rs_n, uint32_t* zm) {
    const uint32_t B0 = (rs_bytes & 0xffu) * 0x01010101u, B1 = ((rs_bytes >> 8) & 0xffu) * 0x01010101u;
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(ts.text + t * ROW_PITCH);
#ifdef RRX_ROLL_MASKS
    uint32_t z0 = 0, z1 = 0, z2 = 0, z3 = 0;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        uint32_t acc = 0, w8[8];
        row_words8(rw, q, w8);
#pragma unroll
        for (int j = 7; j >= 0; j--) {
            const uint32_t wv = w8[j];
            const uint32_t x0 = wv ^ B0;
            uint32_t f = ~(((x0 | 0x80808080u) - 0x01010101u) | x0);
            if (rs_n > 1u) {
                const uint32_t x1 = wv ^ B1;
                f |= ~(((x1 | 0x80808080u) - 0x01010101u) | x1);
            }
            f = (f | wv) & 0x80808080u;
            acc = fshl(f * 0x00204081u, acc, 4);
        }
        z0 = q == 0 ? acc : z0; z1 = q == 1 ? acc : z1; z2 = q == 2 ? acc : z2; z3 = q == 3 ? acc : z3;
        ts.zmask[q * TILE_ROWS + t] = acc;
    }
    zm[0] = z0; zm[1] = z1; zm[2] = z2; zm[3] = z3;
    return;
#endif
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t acc = 0, w8[8];
        row_words8(rw, q, w8);
#pragma unroll
        for (int j = 7; j >= 0; j--) {
            const uint32_t wv = w8[j];
            const uint32_t x0 = wv ^ B0;
            uint32_t f = ~(((x0 | 0x80808080u) - 0x01010101u) | x0);      // bit 7 of a byte set <=> the byte equals b0
            if (rs_n > 1u) {
                const uint32_t x1 = wv ^ B1;
                f |= ~(((x1 | 0x80808080u) - 0x01010101u) | x1);
            }
            f = (f | wv) & 0x80808080u;                                    // ... or is >= 0x80
            acc = fshl(f * 0x00204081u, acc, 4);
        }
        zm[q] = acc;
        ts.zmask[q * TILE_ROWS + t] = acc;
    }
}

template <int NW, class F, bool DONLY, int CW>
RRX_DEV bool phase_a_cand(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int t, const uint32_t* cm,
                          SparseTail& tail) {
    constexpr int ROWS = TILE_ROWS;
    const uint32_t ngr = fw.sp_ngroups(pv), cand = fw.sp_cand(pv);
    const uint8_t* rp = ts.text + t * ROW_PITCH;
#ifdef RRX_NO_OVERRUN
    const uint32_t rowend = DONLY ? (uint32_t)row_nbytes(sp, ts.row0, t) : (uint32_t)ROW_BYTES;
#else
    // bytes of the text from the start of row t on, clamped to two rows: 256 <=> row t AND row t+1 are full rows
    const uint32_t rem2 = DONLY ? row_nbytes2(sp, ts.row0, t) : 2u * ROW_BYTES;
    const uint32_t rowend = rem2 < (uint32_t)ROW_BYTES ? rem2 : (uint32_t)ROW_BYTES;
#endif
    int ncand = popc32(cm[0]) + popc32(cm[1]) + popc32(cm[2]) + popc32(cm[3]);

    // (2) runs of dominated states.  Two nested warp-uniform loops (every lane iterates until all
    // are done, finished lanes idle, so the lanes reconverge after every step):
    //   outer, once per run: a lane without live states jumps to its next candidate; a lane standing
    //          on a candidate byte decides the injection (backward verification) and takes that step;
    //   inner, once per byte: the plain step of a live lane — follow, mask, accept test — a dozen
    //          instructions; a lane that meets a candidate byte inside a run parks until the outer
    //          loop has handled it.
    // Match ends are pushed into a 4-entry byte queue (kq: end+1, ks: tag of the run or 0) and
    // unpacked after the loops.
    uint32_t H[NW], Bc[NW], f[NW], DA[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) { H[w] = 0; DA[w] = fw.sp_D(pv, w); }
    int budget = SPARSE_BUDGET;         // backward verification steps this row may spend (row-bounded scan)
    uint32_t p = 0;
    uint32_t rs1 = 0;                   // tag of the current run if the lazy start of its matches is known, else 0
    SparseTail q;
    q.kq = q.ks = q.nacc = q.drained = q.nruns = 0;
    q.rt = 0ull;
    uint32_t bail = (!DONLY && ncand > SPARSE_MAX_CAND) ? 1u : 0u;
    uint32_t live = 0;                  // OR of the words of H
    uint32_t work = bail ? 0u : (uint32_t)ncand;     // non-zero while the lane has something left to do
    const SmemBytes rb(rp);
    const bool fs_on = fw.fs_n(pv) != 0;
    const uint32_t fsHM = (127u - fw.fs_hi(pv)) * 0x01010101u, fsLO = fw.fs_lo(pv) * 0x01010101u;
    const uint32_t* rw32 = reinterpret_cast<const uint32_t*>(rp);
    uint32_t fs_tried = 0;
#pragma unroll
    for (int qd = 0; qd < 4; qd++) ts.hits[qd * ROWS + t] = 0;
    ts.kst[2 * t] = ts.kst[2 * t + 1] = KST_NONE;
    while (RRX_WARP_ANY(work != 0)) {
        // ---- outer: the queue is full, or the lane is at a candidate byte (or jumps to its next one) ----
        if (RRX_UNLIKELY(work != 0 && q.nacc >= 4)) {
            sparse_drain<NW>(ts, t, q);
            q.kq = q.ks = q.nacc = 0;
            q.drained = 1u;
        } else if (work != 0) {
            const uint32_t fresh = live == 0 ? 1u : 0u;
            if (fresh) {
                uint32_t w = p >> 5;
                uint32_t m = (w == 0 ? cm[0] : w == 1 ? cm[1] : w == 2 ? cm[2] : cm[3]) & (0xffffffffu << (p & 31));
                while (!m) { w++; m = w == 1 ? cm[1] : w == 2 ? cm[2] : cm[3]; }
                p = w * 32 + (uint32_t)ctz32(m);
            }
            // byte p is a candidate here: either the lane just jumped to it, or the inner loop parked on it
            load_B<NW>(pv, cand, Bc);
            fw.follow(pv, H, f);
            uint32_t nx[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) nx[w] = f[w] & DA[w];
            ncand--;
            int n_inj = 0, s_inj = -1;
#pragma unroll
            for (uint32_t g = 0; g < RRX_MAX_GROUPS; g++) {
                if (g >= ngr || bail) break;
                uint32_t need = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) need |= fw.sp_gk(pv, g, w) & ~f[w];
                if (!need) continue;                       // already entered from H
                bool inj = (fw.sp_gflags(pv, g) & 1u) != 0;
                int s = (int)p;
                if (!inj) {
                    uint32_t R[NW], Bq[NW], x[NW];
#pragma unroll
                    for (int w = 0; w < NW; w++) R[w] = fw.sp_gpred(pv, g, w);
                    int qb = (int)p - 1;
                    bool alive = true;
                    if (!(DONLY && fw.rs_n(pv) != 0u)) {       // (with a skip state the general walker below is the faster one)
                        for (; qb >= 0; qb--) {
                            load_B<NW>(pv, rb.at((uint32_t)qb), Bq);
                            uint32_t m = 0;
#pragma unroll
                            for (int w = 0; w < NW; w++) { x[w] = R[w] & Bq[w]; m |= x[w] & fw.first(pv, w); }
                            if (m) { inj = true; s = qb; break; }
                            fw.rfollow(pv, x, R);
                            if (!any_w<NW>(R)) { alive = false; break; }
                            if (!DONLY && --budget < 0) { bail = 1u; break; }
                        }
                    }
                    if (RRX_UNLIKELY(DONLY && !inj && alive)) {
                        // the walk goes on before the row start: earlier rows of the tile, then global memory
                        // (once a walk somewhere has run out of budget the scan is void: let the grid drain quickly)
                        const int64_t row_lo = ts.lo() + ((int64_t)t << ROW_SHIFT);
                        const bool over = (*reinterpret_cast<volatile uint32_t*>(sp.err) & ERR_VERIFY_BUDGET) != 0u;
                        const WalkResult wr = back_walk<NW>(sp, pv, fw, ts, R, row_lo + qb, over ? 1 : VERIFY_BUDGET);
                        if (wr.found) {
                            inj = true;
                            const int64_t d = wr.j - row_lo;
                            s = d < -(int64_t)(KS_BIAS - 1) ? -KS_BIAS : (int)d;      // -KS_BIAS: too far back to record
                        } else if (wr.over) {
                            atomicOr(sp.err, ERR_VERIFY_BUDGET);
                        } else if (wr.pend) {
                            atomicOr(sp.err, ERR_VERIFY_EDGE);      // soft: "nothing before the readable text"
                        }
                    }
                }
                if (inj) {
                    n_inj++;
                    s_inj = s;
#pragma unroll
                    for (int w = 0; w < NW; w++) nx[w] |= fw.sp_gk(pv, g, w);
                }
            }
            // a run fed by one injection: its matches share the start found.  A candidate inside a run: ambiguous from here on.
            rs1 = 0u;
            if (fresh && n_inj == 1 && s_inj > -KS_BIAS) {
                q.rt = (q.rt << 16) | (unsigned long long)(uint32_t)(s_inj + KS_BIAS);
                q.nruns++;
                rs1 = q.nruns & 0xffu;
            }
            uint32_t a = 0;
            live = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                H[w] = nx[w] & Bc[w];
                live |= H[w];
                a |= H[w] & fw.accept(pv, w);
            }
            p++;
            if (a) { q.kq = (q.kq << 8) | p; q.ks = (q.ks << 8) | rs1; q.nacc++; }
            if (!DONLY && (p == (uint32_t)ROW_BYTES || bail)) { live = 0; ncand = 0; }
            work = live | (uint32_t)ncand;
        }
        // ---- inner: plain steps while the run lives and no candidate byte comes up ----
        // (branch-free: every lane computes the step, lanes that are not in a run or stand on a
        // candidate byte just do not commit it)
        // Two steps per vote: the second one is harmless for lanes that have stopped.  Row-bounded scan: there is no
        // end-of-row test — the 4 pad bytes behind every staged row hold 0xff (set once per warp by the
        // kernel; the emulator's fill pattern is >= 0x80 too), a byte no position accepts, so a run that
        // reaches the row end dies there, and every candidate has been consumed by then.  Thread-owned scan: a run
        // that reaches the row end stops THERE with its state intact (the row's exit).
        uint32_t go = (q.nacc < 4 && (!DONLY || p < rowend)) ? live : 0u;
#define RRX_SPARSE_STEP()                                                                       \
        {                                                                                           \
            const uint32_t c = rb.at(p);                /* p == 128: the row's pad, a byte >= 0x80 */ \
            load_B<NW>(pv, c, Bc);                                                                  \
            fw.follow(pv, H, f);                                                                    \
            const bool step = go != 0 && c != cand;     /* c == cand: parked for the outer loop */  \
            uint32_t a = 0, lv = 0;                                                                 \
            _Pragma("unroll") for (int w = 0; w < NW; w++) {                                        \
                const uint32_t hn = f[w] & DA[w] & Bc[w];                                           \
                lv |= hn;                                                                           \
                a |= hn & fw.accept(pv, w);                                                         \
                H[w] = step ? hn : H[w];                                                            \
            }                                                                                       \
            p += step ? 1u : 0u;                                                                    \
            const bool take = step && a != 0;                                                       \
            q.kq = take ? ((q.kq << 8) | p) : q.kq;                                                 \
            q.ks = take ? ((q.ks << 8) | rs1) : q.ks;                                               \
            q.nacc += take ? 1u : 0u;                                                               \
            live = step ? lv : live;                    /* dies on the pad at the latest */          \
            go = (step && q.nacc < 4 && (!DONLY || p < rowend)) ? live : 0u;   /* queue full: parked until drained */ \
        }
        while (RRX_WARP_ANY(go != 0)) {
            // skip: a run in the forward skip state G advances four bytes at a time while they all lie
            // in [lo, hi] (they map G to itself, no match ends, no candidate among them).  A lane tries
            // once per stay in G; runs start together, so the whole warp skips in the same few rounds.
            if (fs_on) {
                uint32_t dG = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) dG |= H[w] ^ fw.fs_G(pv, w);
                uint32_t sk = (go != 0 && dG == 0 && !fs_tried) ? 1u : 0u;
                while (RRX_WARP_ANY(sk != 0)) {
                    if (sk) {
                        const uint32_t i = p >> 2;                       // p <= 128: word 32 is the pad (0xff bytes)
                        const uint32_t w0 = rw32[i], w1 = rw32[i < 32u ? i + 1 : i];
                        const uint32_t w = fshr(w0, w1, 8u * (p & 3u));   // the four bytes at p
                        // any byte >= 0x80, > hi or < lo?  (carries/borrows can only add flags, never hide one)
                        const uint32_t bad = (w | (w + fsHM) | (w - fsLO)) & 0x80808080u;
                        if (bad == 0) p += 4; else { sk = 0; fs_tried = 1u; }
                    }
                }
                if (dG != 0) fs_tried = 0u;
                if (DONLY && p >= rowend) go = 0u;
            }
            RRX_SPARSE_STEP()
            RRX_SPARSE_STEP()
        }
#undef RRX_SPARSE_STEP
        if (DONLY && p >= rowend) {         // the row is done; H (if the run still lives) is the row's exit
            ncand = 0;
            work = 0u;
        } else {
            work = live | (uint32_t)ncand;
        }
    }
    if (bail) return false;
    if (DONLY) {
        // ---- overrun: a run that is alive at the end of its row is followed INTO THE NEXT ROW by the lane that owns it
        // (rows t+1's staged bytes), as long as it meets no candidate byte there: its matches then still precede every
        // match of lane t+1 (whose runs start at ITS candidates), so the lanes' queues stay in text order, row t hands
        // no state to row t+1, and the tile keeps its fast tail.  If the run meets a candidate byte, outlives the queue or
        // the next row, the overrun is undone and the run's state at the row end becomes the row's exit as before.
        // (Not for the tile's last row: its exit travels through the look-back.)
#ifdef RRX_NO_OVERRUN
        const uint32_t over = 0u;
#else
        // Only into a FULL row of the text (rem2): behind the end of the text the staged bytes are fill bytes, the run would
        // seem to die there, and the state alive at the end of the text — the shard's exit, which the next shard enters
        // with — would be lost.
        const uint32_t over = (t < ROWS - 1 && rem2 == 2u * ROW_BYTES && p >= rowend && live != 0u && q.nacc < 4u) ? 1u : 0u;
#endif
        if (RRX_UNLIKELY(RRX_WARP_ANY(over != 0u))) {
            const uint32_t ov = over;
            uint32_t Hx[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) Hx[w] = H[w];
            const uint32_t kq0 = q.kq, ks0 = q.ks, na0 = q.nacc;
            const SmemBytes rb2(rp + ROW_PITCH);
            uint32_t p2 = 0, ovf = 0, go = ov ? live : 0u;
            while (RRX_WARP_ANY(go != 0u)) {
                const uint32_t c = rb2.at(p2);
                load_B<NW>(pv, c, Bc);
                fw.follow(pv, H, f);
                const bool on = go != 0u;
                const bool step = on && c != cand;
                ovf |= (on && c == cand) ? 1u : 0u;                 // another cut crossing: not this lane's business
                uint32_t a = 0, lv = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const uint32_t hn = f[w] & DA[w] & Bc[w];
                    lv |= hn;
                    a |= hn & fw.accept(pv, w);
                    H[w] = step ? hn : H[w];
                }
                p2 += step ? 1u : 0u;
                const bool take = step && a != 0u;
                q.kq = take ? ((q.kq << 8) | (p2 + (uint32_t)ROW_BYTES)) : q.kq;
                q.ks = take ? ((q.ks << 8) | rs1) : q.ks;
                q.nacc += take ? 1u : 0u;
                live = step ? lv : live;
                ovf |= (step && lv != 0u && (q.nacc >= 4u || p2 >= (uint32_t)ROW_BYTES - 1u)) ? 1u : 0u;   // outlives the queue / the next row
                go = (step && !ovf) ? live : 0u;
            }
            if (ov) {
                // (straddle kernel: rows whose run was followed into the next row — ScanParams::straddle_tiles)
                if (sp.straddle_tiles) atomicAdd(sp.straddle_tiles, 1u);
                if (ovf) {                      // undo: the classic hand-over through the row's exit
                    q.kq = kq0; q.ks = ks0; q.nacc = na0;
#pragma unroll
                    for (int w = 0; w < NW; w++) H[w] = Hx[w];
                } else {                        // the run died in the next row: nothing is handed over
#pragma unroll
                    for (int w = 0; w < NW; w++) H[w] = 0u;
                }
            }
        }
        // exit of the row: the run alive after the last byte (a run that ended earlier left H empty)
#pragma unroll
        for (int w = 0; w < NW; w++) if (w < CW) ts.sum_b[w * ROWS + t] = (p >= rowend && rowend != 0u) ? H[w] : 0u;
    }
    // The queue is handed to the caller: a tile whose rows all kept their matches in the queue, with
    // known starts, and whose rows receive no carried state writes them straight from registers
    // (phase_tail_fast); otherwise the caller drains it into the hit masks (sparse_drain).
    tail = q;
    return true;
}

// ===== thread-owned scan: what a row does to CARRIED dominated states =================================
// Row kind for the composites: K_IDENT (no bytes), K_CLOSED (no carried state survives the row — decided by
// walking back from the row end with the dominated positions that can still lie on a surviving path; on
// ordinary text that set is empty after a byte or two), else the survivor analysis of phase A restricted to D:
// K_RANK1 with sum_U / sum_P, or K_HARD.
// *qfree: a closed row's exit (sum_b, known only after the candidates have been followed) is certainly empty if the row
// has no candidate byte at or after this position — what lets a tile publish its aggregate before it scans.
template <int NW, class F, int CW>
RRX_COLD uint32_t phase_d_kind(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int t, int* qfree) {
    constexpr int ROWS = TILE_ROWS;
    const int nb = row_nbytes(sp, ts.row0, t);
    *qfree = 0;
    if (nb == 0) { ts.sum_kind[t] = K_IDENT; return K_IDENT; }
    const SmemBytes rb(ts.text + t * ROW_PITCH);
    uint32_t A[NW], x[NW], Bc[NW], f[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) A[w] = fw.sp_D(pv, w);
    bool open = true;
    for (int q = nb - 1; q >= 0; q--) {
        load_B<NW>(pv, rb.at((uint32_t)q), Bc);
#pragma unroll
        for (int w = 0; w < NW; w++) x[w] = A[w] & Bc[w];
        if (!any_w<NW>(x)) { open = false; *qfree = q + 1; break; }     // nothing alive after byte q survives
        fw.rfollow(pv, x, A);
#pragma unroll
        for (int w = 0; w < NW; w++) A[w] &= fw.sp_D(pv, w);
        if (!any_w<NW>(A)) { open = false; *qfree = q; break; }         // only a thread injected AT byte q could
    }
    if (!open) { ts.sum_kind[t] = K_CLOSED; return K_CLOSED; }
    // open: survivors of ALL dominated states, no injection
    uint32_t U[NW], J[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) { U[w] = fw.sp_D(pv, w); J[w] = 0; }
    int pc = -1;
    for (int p = 0; p < nb; p++) {
        load_B<NW>(pv, rb.at((uint32_t)p), Bc);
        fw.follow(pv, U, f);
#pragma unroll
        for (int w = 0; w < NW; w++) U[w] = f[w] & Bc[w] & fw.sp_D(pv, w);
        if (!any_w<NW>(U)) { ts.sum_kind[t] = K_CLOSED; return K_CLOSED; }       // (cannot happen: the walk above said open)
        if (pc < 0 && popc_w<NW>(U) == 1) {
            pc = p + 1;
#pragma unroll
            for (int w = 0; w < NW; w++) J[w] = U[w];
        }
    }
    const uint32_t kind = pc < 0 ? K_HARD : K_RANK1;
    uint32_t R[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) R[w] = J[w];
    if (pc >= 0) {
        for (int q = pc - 1; q >= 0; q--) {           // entry states that lead to the lone survivor
            load_B<NW>(pv, rb.at((uint32_t)q), Bc);
#pragma unroll
            for (int w = 0; w < NW; w++) x[w] = R[w] & Bc[w];
            fw.rfollow(pv, x, R);
#pragma unroll
            for (int w = 0; w < NW; w++) R[w] &= fw.sp_D(pv, w);
        }
    }
#pragma unroll
    for (int w = 0; w < NW; w++) {
        if (w < CW) {
            ts.sum_U[w * ROWS + t] = U[w];
            ts.sum_P[w * ROWS + t] = R[w];
        }
    }
    ts.sum_kind[t] = kind;
    return kind;
}

// ===== phase A2: the warp replays the queued 4-byte words, one per lane ==============================
// (after all lanes have finished phase A and stored their hit words)
template <int NW, class F, int CW>
RRX_DEV void phase_a2(const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int lane) {
    constexpr int ROWS = TILE_ROWS;
    using TS = TileState<NW, CW>;
    uint32_t n = ts.misc[TS::M_RQCNT];
    if (n > (uint32_t)TS::RQ_CAP) n = TS::RQ_CAP;
    for (uint32_t i = (uint32_t)lane; i < n; i += ROWS) {
        const uint32_t* it = ts.rq + i * (1 + NW);
        const uint32_t row = it[0] >> 5, quad = it[0] & 31u;
        uint32_t T[NW], Bc[NW], f[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) T[w] = it[1 + w];
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            load_B<NW>(pv, ts.row_byte((int)row, (int)(quad * 4 + k)), Bc);
            fw.follow(pv, T, f);
            uint32_t a2 = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                T[w] = (f[w] | fw.first(pv, w)) & Bc[w];
                a2 |= T[w] & fw.accept(pv, w);
            }
            if (a2) bits |= 1u << k;
        }
        atomicOr(ts.hits + (quad >> 3) * ROWS + row, bits << ((quad & 7u) * 4u));
    }
}

template <int NW, int CW> RRX_DEV void load_row_comp(const TileState<NW, CW>& ts, int t, Comp<CW>& c) {
    constexpr int ROWS = TILE_ROWS;
    c.kind = ts.sum_kind[t];
#pragma unroll
    for (int w = 0; w < CW; w++) {
        c.b[w] = ts.sum_b[w * ROWS + t];
        c.U[w] = (c.kind & K_TYPE) == K_RANK1 ? ts.sum_U[w * ROWS + t] : 0u;
        c.P[w] = (c.kind & K_TYPE) == K_RANK1 ? ts.sum_P[w * ROWS + t] : 0u;
    }
}
template <int NW, int CW> RRX_DEV void store_row_comp(TileState<NW, CW>& ts, int t, const Comp<CW>& c) {
    constexpr int ROWS = TILE_ROWS;
    ts.sum_kind[t] = c.kind;
#pragma unroll
    for (int w = 0; w < CW; w++) {
        ts.sum_b[w * ROWS + t] = c.b[w];
        ts.sum_U[w * ROWS + t] = c.U[w];
        ts.sum_P[w * ROWS + t] = c.P[w];
    }
}
template <int NW> RRX_DEV void store_comp_global(uint32_t* dst, const Comp<NW>& c) {
    dst[0] = c.kind;
#pragma unroll
    for (int w = 0; w < NW; w++) { dst[1 + w] = c.b[w]; dst[1 + NW + w] = c.U[w]; dst[1 + 2 * NW + w] = c.P[w]; }
}
template <int NW> RRX_DEV void load_comp_global(const uint32_t* src, Comp<NW>& c) {
    const volatile uint32_t* v = src;
    c.kind = v[0];
#pragma unroll
    for (int w = 0; w < NW; w++) { c.b[w] = v[1 + w]; c.U[w] = v[1 + NW + w]; c.P[w] = v[1 + 2 * NW + w]; }
}
// The shard composite as it leaves the scanner: always NW words per set (rrx.h), whatever the kernel's carried width.
template <int NW, int CW> RRX_DEV void store_comp_external(uint32_t* dst, const Comp<CW>& c) {
    dst[0] = c.kind;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        dst[1 + w] = w < CW ? c.b[w < CW ? w : 0] : 0u;
        dst[1 + NW + w] = w < CW ? c.U[w < CW ? w : 0] : 0u;
        dst[1 + 2 * NW + w] = w < CW ? c.P[w < CW ? w : 0] : 0u;
    }
}


// Exit of the no-injection replay of row t from the exact entry set X (in place).
template <int NW, class F, int CW>
RRX_DEV void row_replay_exit(const ScanParams& sp, const ProgView& pv, const F& fw, const TileState<NW, CW>& ts, int t,
                             uint32_t* X) {
    const int nb = row_nbytes(sp, ts.row0, t);
    const bool dn = fw.donly(pv);          // thread-owned scan: only dominated states are carried
    uint32_t Bc[NW], f[NW];
    for (int p = 0; p < nb && any_w<NW>(X); p++) {
        load_B<NW>(pv, ts.row_byte(t, p), Bc);
        fw.follow(pv, X, f);
#pragma unroll
        for (int w = 0; w < NW; w++) X[w] = f[w] & Bc[w] & (dn ? fw.sp_D(pv, w) : 0xffffffffu);
    }
}

// ===== phase S (one lane) ===============================================================================
// S1: compose the rows into the tile aggregate (no global traffic).
// [look-back: yields the composite of everything before the tile]
// S2: tile entry, blocked segment, inclusive composite.
// After S2 sum_* of row t hold the composite of rows [0,t) of the tile (what turns the tile entry
// into row t's entry) and misc[M_ENTRY..] holds the tile's exact entry set.
template <int CW> struct SLocal {
    Comp<CW> run;          // tile aggregate (kind K_WAIT if it cannot be summarised)
    Comp<CW> pre_block;    // composite of the rows before the blocked segment
    int blocked_from, resume_at;
    bool wait_agg;
};

// widen a carried set (CW words) to a full state set (NW words) and back
template <int NW, int CW> RRX_DEV void widen(const uint32_t* c, uint32_t* full) {
#pragma unroll
    for (int w = 0; w < NW; w++) full[w] = w < CW ? c[w < CW ? w : 0] : 0u;
}

template <int NW, class F, int CW>
RRX_COLD void phase_s1(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, bool any_open,
                      SLocal<CW>& L) {
    constexpr int ROWS = TILE_ROWS;
    Comp<CW> rowc, nxt;
    comp_ident(L.run);
    comp_ident(L.pre_block);
    comp_ident(nxt);
    L.blocked_from = -1;
    L.resume_at = ROWS;
    L.wait_agg = false;
    if (any_open) {
        for (int t = 0; t < ROWS; t++) {
            if (L.blocked_from >= 0 && t < L.resume_at) continue;     // inside the blocked segment
            load_row_comp(ts, t, rowc);
            if ((rowc.kind & K_TYPE) == K_HARD && (L.run.kind & K_TYPE) != K_CLOSED) {
                // A row we cannot summarise, reached with a state that still depends on the tile
                // entry.  Everything up to the next closed row has to wait for the look-back.
                L.blocked_from = t;
                L.pre_block = L.run;
                int c = -1;
                for (int u = t + 1; u < ROWS; u++) if ((ts.sum_kind[u] & K_TYPE) == K_CLOSED) { c = u; break; }
                if (c < 0) { L.wait_agg = true; L.resume_at = ROWS; break; }
                L.resume_at = c + 1;
                uint32_t bc[CW];
#pragma unroll
                for (int w = 0; w < CW; w++) bc[w] = ts.sum_b[w * ROWS + c];
                comp_closed(L.run, bc, 0);
                continue;
            }
            store_row_comp(ts, t, L.run);
            if ((rowc.kind & K_TYPE) == K_HARD) {                  // exact state: replay the row
                uint32_t X[NW];
                widen<NW, CW>(L.run.b, X);
                row_replay_exit<NW>(sp, pv, fw, ts, t, X);
#pragma unroll
                for (int w = 0; w < CW; w++) X[w] |= rowc.b[w];
                comp_closed(L.run, X, L.run.kind & K_OPAQUE);
            } else {
                comp_compose(L.run, rowc, nxt);
                L.run = nxt;
            }
        }
    } else {
        // every row closed (or empty, at the very end of the shard): row t's entry is the exit of
        // the row before it; nothing to rewrite, only the tile aggregate is needed.
        int last = -1;
        for (int t = ROWS - 1; t >= 0; t--) if (ts.sum_kind[t] != K_IDENT) { last = t; break; }
        if (last >= 0) load_row_comp(ts, last, L.run);
    }
    if (L.wait_agg) L.run.kind = K_WAIT;
}

// `acc` = composite of the whole shard before this tile.  Returns the tile's inclusive composite.
template <int NW, class F, int CW>
RRX_COLD void phase_s2(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, const SLocal<CW>& L,
                      const Comp<CW>& acc, Comp<CW>& inc) {
    using TS = TileState<NW, CW>;
    uint32_t E[CW];
    comp_apply<CW>(acc, sp.entry, E);
#pragma unroll
    for (int w = 0; w < CW; w++) ts.misc[TS::M_ENTRY + w] = E[w];
    if (L.blocked_from >= 0) {
        // second stage: walk the blocked segment with exact states
        uint32_t X[CW], Y[NW];
        comp_apply<CW>(L.pre_block, E, X);
        Comp<CW> ex, rowc;
        for (int t = L.blocked_from; t < L.resume_at; t++) {
            load_row_comp(ts, t, rowc);
            comp_closed(ex, X, 0);
            store_row_comp(ts, t, ex);
            const uint32_t rt = rowc.kind & K_TYPE;
            if (rt == K_HARD) {
                widen<NW, CW>(X, Y);
                row_replay_exit<NW>(sp, pv, fw, ts, t, Y);
#pragma unroll
                for (int w = 0; w < CW; w++) X[w] = Y[w] | rowc.b[w];
            } else {
                uint32_t Z[CW];
                comp_apply<CW>(rowc, X, Z);
#pragma unroll
                for (int w = 0; w < CW; w++) X[w] = Z[w];
            }
        }
        if (L.wait_agg) { comp_closed(inc, X, K_OPAQUE); return; }
    }
    comp_compose(acc, L.run, inc);   // inclusive composite (relative to the shard entry)
}

// ===== phase B: replay each row's prefix from its exact entry set (no injection) ====================
// Returns the number of matches of the row.
template <int NW, class F, int CW>
RRX_COLD uint32_t phase_b(const ScanParams& sp, const ProgView& pv, const F& fw, TileState<NW, CW>& ts, int t,
                         bool any_open) {
    constexpr int ROWS = TILE_ROWS;
    using TS = TileState<NW, CW>;
    const int nb = row_nbytes(sp, ts.row0, t);
    uint32_t X[NW], Bc[NW], f[NW], Xc[CW];
    const uint32_t* E = ts.misc + TS::M_ENTRY;
    if (any_open) {
        Comp<CW> pre;
        load_row_comp(ts, t, pre);
        comp_apply<CW>(pre, E, Xc);
    } else if (t == 0) {
#pragma unroll
        for (int w = 0; w < CW; w++) Xc[w] = E[w];
    } else {
#pragma unroll
        for (int w = 0; w < CW; w++) Xc[w] = ts.sum_b[w * ROWS + (t - 1)];
    }
    widen<NW, CW>(Xc, X);
    uint32_t hit[4] = {0, 0, 0, 0};
    bool touched = false;
    const bool dn = fw.donly(pv);          // thread-owned scan: only dominated states are carried
    for (int p = 0; p < nb && any_w<NW>(X); p++) {
        load_B<NW>(pv, ts.row_byte(t, p), Bc);
        fw.follow(pv, X, f);
        uint32_t a2 = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            X[w] = f[w] & Bc[w] & (dn ? fw.sp_D(pv, w) : 0xffffffffu);
            a2 |= X[w] & fw.accept(pv, w);
        }
        if (a2) { hit[p >> 5] |= 1u << (p & 31); touched = true; }
    }
    if (dn && touched) {
        // A match end the carried state reaches as well: the start the row's own scan recorded for it (which may
        // lie before the row) need not be the lazy one any more — let the backward walk decide.
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t k = ts.kst[2 * t + i];
            const uint32_t pos = k >> 16, wd = pos >> 5;
            const uint32_t hw = wd == 0u ? hit[0] : wd == 1u ? hit[1] : wd == 2u ? hit[2] : hit[3];
            if (k != KST_NONE && ((hw >> (pos & 31u)) & 1u)) ts.kst[2 * t + i] = KST_NONE;
        }
    }
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t v = ts.hits[q * ROWS + t];
        if (touched) { v |= hit[q]; ts.hits[q * ROWS + t] = v; }
        c += (uint32_t)popc32(v);
    }
    return c;
}

// ===== phase C: resolve the lazy start of every match and write (start,end) ========================
// Backward walk from the accepting positions, starting with the byte at shard-local offset e-1; stops at the
// first (= right-most) start that works: lazy.
template <int NW, class F, int CW>
RRX_COLD void resolve_and_write(const ScanParams& sp, const ProgView& pv, const F& fw, const TileState<NW, CW>& ts,
                               int64_t e, unsigned long long slot, uint32_t rank_in_tile) {
    uint32_t R[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) R[w] = fw.accept(pv, w);
    const WalkResult wr = back_walk<NW>(sp, pv, fw, ts, R, e - 1, 0);
    if (slot < sp.out_cap) {
        long long* o = sp.out + 2 * slot;
        o[0] = wr.found ? sp.base + wr.j : -1;        // -1: the start lies left of the available text
        o[1] = sp.base + e;
    }
    if (!wr.found && !wr.pend) atomicOr(sp.err, ERR_START_LOST);
    if (wr.pend) {
        const unsigned long long k = atomicAdd(sp.total + 1, 1ull);
        if (k < sp.pending_cap) {
            PendingStart& ps = sp.pending[k];
            ps.match_index = ((uint64_t)(ts.row0 / TILE_ROWS) << 32) | rank_in_tile;
            ps.next_offset = wr.j;
#pragma unroll
            for (int w = 0; w < NW; w++) ps.R[w] = R[w];
        } else {
            atomicOr(sp.err, ERR_PENDING_OVERFLOW);
        }
    }
}

// The match ending after tile-relative byte `rel`: if phase A already knows its lazy start, write it.
template <int NW, int CW>
RRX_DEV bool write_known(const ScanParams& sp, const TileState<NW, CW>& ts, uint32_t rel, unsigned long long slot) {
    const uint32_t row = rel >> ROW_SHIFT, p = rel & (ROW_BYTES - 1);
    const uint32_t k0 = ts.kst[2 * row], k1 = ts.kst[2 * row + 1];
    uint32_t s;
    if ((k0 >> 16) == p) s = k0 & 0xffffu;
    else if ((k1 >> 16) == p) s = k1 & 0xffffu;
    else return false;
    if (slot < sp.out_cap) {
        long long* o = sp.out + 2 * slot;
        const long long row_lo = sp.base + ts.lo() + ((long long)row << ROW_SHIFT);
        o[0] = row_lo + (long long)s - KS_BIAS;
        o[1] = row_lo + p + 1;
    }
    return true;
}

// C1 (per lane): append the row's match ends to the tile's hit list (tile-relative end-1).
// cnt[t] must hold the row's exclusive offset.  Only when the tile total fits the list.
template <int NW, int CW> RRX_COLD void phase_c1_list(TileState<NW, CW>& ts, int t, uint32_t offset, const uint32_t* hit) {
    uint32_t k = offset;
    for (int q = 0; q < 4; q++) {
        uint32_t v = hit[q];
        while (v) {
            const int p = q * 32 + ctz32(v);
            v &= v - 1;
            ts.list[k++] = (uint16_t)((t << ROW_SHIFT) + p);
        }
    }
}
// C2 (per lane): balanced — lane handles list entries lane, lane+32, ...
template <int NW, class F, int CW>
RRX_COLD void phase_c2_balanced(const ScanParams& sp, const ProgView& pv, const F& fw, const TileState<NW, CW>& ts, int lane,
                               uint32_t total, unsigned long long base) {
    for (uint32_t i = (uint32_t)lane; i < total; i += TILE_ROWS) {
        const uint32_t rel = ts.list[i];
        const int64_t e = ts.lo() + (int64_t)rel + 1;
        if (!write_known<NW>(sp, ts, rel, base + i)) resolve_and_write<NW>(sp, pv, fw, ts, e, base + i, i);
    }
}
// C2 fallback (per lane): every row resolves its own matches (tiles with more hits than the list holds).
template <int NW, class F, int CW>
RRX_COLD void phase_c2_rows(const ScanParams& sp, const ProgView& pv, const F& fw, const TileState<NW, CW>& ts, int t,
                           uint32_t offset, const uint32_t* hit, unsigned long long base) {
    unsigned long long slot = base + offset;
    const int64_t row_lo = ts.lo() + ((int64_t)t << ROW_SHIFT);
    for (int q = 0; q < 4; q++) {
        uint32_t v = hit[q];
        while (v) {
            const int p = q * 32 + ctz32(v);
            v &= v - 1;
            if (!write_known<NW>(sp, ts, ((uint32_t)t << ROW_SHIFT) + (uint32_t)p, slot)) resolve_and_write<NW>(sp, pv, fw, ts, row_lo + p + 1, slot, (uint32_t)(slot - base));
            slot++;
        }
    }
}

}  // namespace rrx
