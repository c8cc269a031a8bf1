// Device program layout of the rrx scanner — plain C structs shared by the host compiler
// (rrx_compile.cc), the kernels (rrx_tile.cuh) and the runtime-compiled specialisations.
// No STL here: this header is also fed to NVRTC.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

namespace rrx {

// The device automaton works on the pruned edge set E' = follow minus edges into `first`:
// with per-byte re-injection S' = (follow'(S) | first) & B[c], an edge into a first-position is
// subsumed by the injection, and the backward start scan stops at the first first-position, so it
// never traverses such an edge either (DESIGN.md "Pruned edges").
enum : uint32_t { RRX_PROG_MAGIC = 0x52525831u };   // "RRX1"
enum : uint32_t { TERM_ROT = 0, TERM_SHL = 1, TERM_SHR = 2, TERM_SELF = 3 };
constexpr uint32_t RRX_MAX_TERMS = 8;
constexpr uint32_t RRX_MAX_BICL = 8;

// All fields are 32-bit words so the blob can be copied to the device verbatim.
struct ProgHeader {
    uint32_t magic, P, NW, flags;
    uint32_t n_terms, n_bicl, use_rows, pad0;        // forward decomposition
    uint32_t n_rterms, n_rbicl, use_rrows, pad1;     // reverse decomposition
    // word offsets from the start of the blob
    uint32_t off_first, off_accept, off_all;
    uint32_t off_terms, off_bicl, off_rowsrc;        // terms: n*(1+NW) [kind<<8|r, mask]; bicl: n*2*NW
    uint32_t off_rterms, off_rbicl, off_rrowsrc;
    uint32_t off_B;                                   // 256*NW
    uint32_t off_follow, off_rfollow;                 // P*NW each: full pruned follow'/its transpose
    uint32_t off_rows, off_rrows;                     // P*NW each: edges NOT covered by terms/bicliques
    uint32_t total_words;
    uint32_t n_raw, ref_width_class, nullable;
    uint32_t static_depth;      // longest path (edges) in the pruned follow graph; 0xffffffff if cyclic
    uint32_t n_flat;            // atoms + START before reduction
    uint32_t off_rfull;         // P*NW: transpose of the UNPRUNED follow relation (whole-string acceptance)
    // Candidate-driven ("sparse") row scan, see rrx_tile.cuh phase_a_sparse.  A CUT is a set K of
    // positions that every activation chain first -> accept passes through and whose class is the
    // single byte `cand_byte`: a match can only end after a run of dominated states that starts at an
    // occurrence of that byte.  off_sparse: D[NW] (positions dominated by K), then n_groups records
    // {kmask[NW], pred[NW], flags} — the K positions grouped by predecessor set; flags bit 0: the
    // group's positions are first-positions (injected without looking back).
    uint32_t sparse_ok, cand_byte, n_groups, off_sparse;
    // Reverse-walk skip state (lazy start scans, rrx_tile.cuh resolve_and_write): a reverse state set F
    // (off_rskip, NW words) that every 7-bit byte except rskip_bytes[0..rskip_n) maps to itself without
    // reaching a first-position.  While the backward walk is in F it only has to look for those bytes.
    uint32_t rskip_n;           // 0: no such state; else 1 or 2
    uint32_t rskip_bytes;       // byte0 | byte1 << 8
    uint32_t off_rskip;
    // Forward skip state of the candidate-driven row scan: a set G of dominated states (off_fskip, NW
    // words) that every byte of the range [fskip_lo, fskip_hi] maps to itself without a match end
    // ("[a-z]+" inside a run).  While a run is in G it advances four bytes at a time as long as they
    // all lie in the range.  fskip_n = hi - lo + 1, 0: none.  The candidate byte is outside the range.
    uint32_t fskip_n, fskip_lo, fskip_hi, off_fskip;
    // "Thread-owned" candidate scan (rrx_tile.cuh, DONLY): only the dominated states D are ever carried from
    // row to row, tile to tile and shard to shard.  Whether a cut position is entered at a candidate byte is
    // decided by a backward walk that may leave the row, the tile (global memory) and reach back to the left
    // edge of the readable text, so no state outside D has to be known anywhere: no exit-set scans, no dense
    // fallback.  Chosen by the compiler when runs of dominated states die by themselves (no broad-class
    // loop inside D) and the verification walk is short or has a skip state; RRX_COMPILE_NO_DONLY forbids it.
    uint32_t donly;
    // Carried width: with `donly` only subsets of D travel between rows, tiles and shards.  The compiler numbers the
    // dominated positions first, so those sets fit the first CW words (CW a power of two <= NW); a kernel specialised
    // for the program keeps its composites, row summaries and look-back payloads CW words wide.  Generic kernels
    // use NW.  (Everything that leaves the scanner — shard summaries, entry sets — stays NW words wide.)
    uint32_t CW;
    // Anchors.  The reference compiles `^` and `$` to a transition on byte 0x00 (Parser.cpp:142-146) that its iterator
    // never feeds.  Here the text is framed by ONE virtual 0x00 on each side which only those positions consume:
    // off_anch = the positions whose class is exactly {0x00} (NW words), n_anchor their number.  The one-shot scan
    // calls turn the leading virtual byte into an entry set and the trailing one into a final virtual step (rrx_capi.cu).
    uint32_t n_anchor, off_anch;
};
constexpr uint32_t RRX_MAX_GROUPS = 4;

}  // namespace rrx
