/* rrx — C ABI of the B200 regex scanner that stands in for RoaringRegex's hot path.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in the signatures (a CUDA stream is
 * passed as void*).  Every entry point names the reference interface it replaces
 * (file:line relative to the RoaringRegex tree).  librrx.so has NO CPU scanning path: every
 * rrx_scan* call runs the sm_100a kernels or fails with RRX_ERR_CUDA.
 *
 * Anchors: `^` / `$` hold at the ends of the text handed to a one-shot call (rrx_scan_host / rrx_scan_device): the
 * text is framed by one virtual 0x00 byte on each side that only the anchor positions consume (the
 * reference compiles them to a transition on 0x00 that it never feeds, so there they never match).
 *
 * Match semantics: "LAZY-v1" (DESIGN.md / SURVEY.md Appendix D): for every end offset e such that
 * some text[s:e], s<e, is in the language, exactly one match (max such s, e); ascending e.
 * Offsets are 64-bit, half-open, global (base_offset + position in the buffer).
 */
#ifndef RRX_H_
#define RRX_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RRX_OK = 0,
    RRX_ERR_PARSE = 1,        /* the reference would throw std::runtime_error (Parser.cpp:36,155) or has UB */
    RRX_ERR_CUDA = 2,         /* no device / CUDA runtime failure; there is no CPU fallback */
    RRX_ERR_UNSUPPORTED = 3,  /* the request is outside what this build implements (e.g. an unknown scan mode) */
    RRX_ERR_ARG = 4
};

/* RRX_MODE_NO_OVERLAP: the finditer-style view of the LAZY-v1 list:
 *   p = base_offset; for (s,e) in LAZY-v1: if s >= p { emit; p = e }
 * computed on the device (csrc/rrx_overlap.cu) by the one-shot calls rrx_scan_host / rrx_scan_device;
 * the reusable scanner (rrx_scanner_scan) produces LAZY-v1 only. */
enum { RRX_MODE_LAZY_V1 = 0, RRX_MODE_NO_OVERLAP = 1 };
/* RRX_COMPILE_NO_DONLY: never choose the thread-owned candidate scan (rrx_info.donly) for this program. */
enum { RRX_COMPILE_DEFAULT = 0, RRX_COMPILE_NO_REDUCE = 1, RRX_COMPILE_NO_DONLY = 2 };

typedef struct rrx_program rrx_program;   /* compiled regex: replaces Regex::RRegex's NFA<S> + MemoryPool (regex.h:16-75,212-228) */
typedef struct rrx_scanner rrx_scanner;   /* reusable device-side scan context (one per CUDA stream) */
typedef struct rrx_matches rrx_matches;   /* a match list: replaces repeated IteratorWrapper::operator++ / operator* (regex.h:113-122) */

typedef struct rrx_info {
    uint32_t n_raw;          /* reference raw state count (Parser.cpp:163) */
    uint32_t ref_width_class;/* set type the reference would pick (Parser.cpp:165-168): 0 BitSet<1> 1 BitSet<2> 2 BitSet<4> 3 Roaring */
    uint32_t n_flat;         /* flattened states = atoms + START */
    uint32_t n_device;       /* positions in the device automaton (after merging equivalent positions) */
    uint32_t words;          /* 32-bit words per device state set: 1,2,4,8,16,32 */
    uint32_t nullable;       /* the empty string is in the language */
    uint32_t n_terms, n_bicliques, uses_rows;   /* forward follow decomposition */
    uint32_t static_closed;  /* no chunk's exit can depend on its entry (acyclic pruned follow, depth < 128) */
    uint32_t sparse_ok;      /* a single-byte cut exists: rows are scanned candidate-driven (DESIGN.md "Sparse row scan") */
    uint32_t cand_byte;      /* that byte */
    uint32_t donly;          /* thread-owned candidate scan: only the states dominated by the cut are carried between rows,
                              * tiles and shards; candidates are verified by backward walks (DESIGN.md "Thread-owned scan") */
    uint32_t carried_words;  /* 32-bit words that hold every carried set (dominated positions are numbered first) */
} rrx_info;

/* ---- compile: replaces RRegex::RRegex(const char*) (regex.h:224, Parser.cpp:161-170) ---------- */
int rrx_compile(const char* ere, rrx_program** out);
int rrx_compile_ex(const char* ere, uint32_t flags, rrx_program** out);
void rrx_program_free(rrx_program* p);
void rrx_program_info(const rrx_program* p, rrx_info* out);
/* The flattened automaton BEFORE device reduction, for inspection/tests: positions = n_flat-1.
 * which: 0 first, 1 accept (words32 = ceil(P/32) each), 2 follow matrix (P rows),
 *        3 byte classes (P rows of 4 words = 128 bits).  Returns words written or <0. */
int64_t rrx_program_flat_table(const rrx_program* p, int which, uint32_t* out, uint64_t cap_words);
/* The device program blob (ProgHeader + tables, see csrc/rrx_compile.h). */
const uint32_t* rrx_program_blob(const rrx_program* p, uint64_t* n_words);
/* Raw reference-layout forward row, rebuilt from the tracked atoms (NFA.cc:9-12 fwd rows):
 * targets of `state` on `byte` as a bitset of ceil(n_raw/32) words.  For fixture tests. */
int64_t rrx_program_raw_row(const rrx_program* p, uint32_t state, uint32_t byte, uint32_t* out, uint64_t cap_words);
int64_t rrx_program_raw_finals(const rrx_program* p, uint32_t* out, uint64_t cap_words);

/* ---- scan: replaces RRegex::get_acceptance_iter(char*) + IteratorWrapper::operator++/operator*
 *      (regex.h:118-119,156-162,225) and supplies the lazy iterator Iterator.cpp leaves empty ---- */
/* d_text: device pointer, 16-byte aligned.  stream: cudaStream_t (NULL = default stream). */
int rrx_scan_device(const rrx_program* p, const uint8_t* d_text, uint64_t len, uint64_t base_offset,
                    int mode, rrx_matches** out, void* stream);
/* text: host pointer; the copy to the device is part of the call (asynchronous when the buffer is
 * page-locked).  The program keeps the device text buffer, a scanner and pinned match buffers between
 * calls, so repeated host scans do not allocate.  The match list of a host scan lives in pinned host
 * memory: rrx_matches_data() returns it without a further copy; rrx_matches_device_data() is NULL. */
int rrx_scan_host(const rrx_program* p, const uint8_t* text, uint64_t len, uint64_t base_offset,
                  int mode, rrx_matches** out);

/* Whole-string acceptance, what the reference's AcceptanceIterator answers (regex.h:150-165):
 * *accepted = 1 iff text[0:len) is in the language.  Sequential in len (one GPU thread), like the
 * reference; provided for source compatibility, not for throughput. */
int rrx_accepts_host(const rrx_program* p, const uint8_t* text, uint64_t len, int* accepted);

uint64_t rrx_matches_count(const rrx_matches* m);
const int64_t* rrx_matches_data(rrx_matches* m);                /* host pairs (start,end); copied D2H on first use */
const int64_t* rrx_matches_device_data(const rrx_matches* m);   /* device pairs */
void rrx_matches_free(rrx_matches* m);

/* ---- reusable context (what the convenience calls above are built from; used by bench.py) ----- */
typedef struct rrx_scan_opts {
    uint64_t left_context;     /* bytes readable before d_text (for start scans crossing the left edge) */
    const uint32_t* entry;     /* forward set carried into d_text[0] (words32 per rrx_info.words), or NULL */
    int mode;
    uint32_t flags;            /* RRX_SCAN_* */
} rrx_scan_opts;
/* RRX_SCAN_NO_DONLY: run this scan with the row-bounded candidate scan even if the program has rrx_info.donly
 * (what a sharded driver asks for when a thread-owned scan reported RRX_SOFT_VERIFY_EDGE on a shard that has
 * text to its left).  A thread-owned program also takes this path by itself when `entry` holds states outside
 * the dominated set, and when a verification walk runs out of its step budget. */
enum { RRX_SCAN_NO_DONLY = 1 };

typedef struct rrx_scan_result {
    uint64_t n_matches;        /* total found (may exceed capacity: then only the first `capacity` were written) */
    uint64_t capacity;
    uint64_t n_pending;        /* start scans that ran off the left edge (sharded use) */
    uint32_t error_bits;       /* kernel-reported, 0 = fine: 2 pending-start list overflow, 4 staging overflow (reported as
                                * n_matches > capacity), 8 a start walk died (the entry set given was not exact) */
    uint32_t launches;         /* kernels launched by the scan */
    const int64_t* d_matches;  /* device pairs */
    uint32_t soft_bits;        /* RRX_SOFT_*: the list is valid under the stated assumption */
    uint32_t reserved;
} rrx_scan_result;
/* RRX_SOFT_VERIFY_EDGE: a backward verification walk of the thread-owned scan reached the left edge of the readable
 * text (d_text - left_context) still alive and assumed that nothing precedes it.  True for a whole text or its first
 * shard; a shard with text to its left needs more left context or RRX_SCAN_NO_DONLY. */
enum { RRX_SOFT_VERIFY_EDGE = 16 };

int rrx_scanner_create(const rrx_program* p, uint64_t max_len, uint64_t match_capacity, rrx_scanner** out);
/* flags: RRX_SCANNER_JIT / RRX_SCANNER_NO_JIT force / forbid the run-time specialised kernel (NVRTC);
 * default: specialise when max_len >= 32 MiB and the automaton fits one 32-bit word.
 * The environment variable RRX_JIT=0/1 overrides the default.  If the default policy wants the specialised
 * kernel but NVRTC is unavailable, the scanner falls back to the pre-compiled kernel (same results), prints one
 * warning line to stderr and leaves the text in rrx_last_error(); with RRX_SCANNER_JIT it fails instead. */
enum { RRX_SCANNER_DEFAULT = 0, RRX_SCANNER_JIT = 1, RRX_SCANNER_NO_JIT = 2 };
int rrx_scanner_create_ex(const rrx_program* p, uint64_t max_len, uint64_t match_capacity, uint32_t flags, rrx_scanner** out);
/* 1 if the scanner runs a run-time specialised kernel, 0 if a pre-compiled one. */
int rrx_scanner_is_jit(const rrx_scanner* s);
/* Specialised kernel the NEXT scan will use: 0 lean, 1 straddle (matches straddling 128-byte rows are followed by
 * their lane).  A scanner starts lean and switches by what each scan saw (rrx_scanner_result); RRX_JIT_VARIANT=0/1
 * pins it. */
int rrx_scanner_kernel_variant(const rrx_scanner* s);
/* 1 if the last scan ran on the fallback (row-bounded) scanner of a thread-owned program. */
int rrx_scanner_used_fallback(const rrx_scanner* s);
/* CUDA source of the specialised follow functions for this program (for inspection); returns the
 * length needed (including the NUL), copies at most cap bytes. */
uint64_t rrx_program_jit_source(const rrx_program* p, char* buf, uint64_t cap);
/* Compile the specialised kernel to an sm_100a cubin with NVRTC (no GPU needed) and write it to `path`
 * (NULL: discard).  Returns RRX_OK, or RRX_ERR_UNSUPPORTED / RRX_ERR_CUDA with rrx_last_error(). */
int rrx_program_jit_compile(const rrx_program* p, const char* path);
void rrx_scanner_free(rrx_scanner* s);
/* Enqueue one scan on `stream` (asynchronous). */
int rrx_scanner_scan(rrx_scanner* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset,
                     const rrx_scan_opts* opts, void* stream);
/* Wait for the last scan and fetch its counters. */
int rrx_scanner_result(rrx_scanner* s, rrx_scan_result* out, void* stream);
/* Copy the first n matches of the last scan to host memory (n*2 int64). */
int rrx_scanner_copy_matches(rrx_scanner* s, int64_t* host_out, uint64_t n, void* stream);
/* Composite transfer function of the last scanned shard: 1+3*words uint32 (kind, b, U, P); host copy. */
int rrx_scanner_shard_summary(rrx_scanner* s, uint32_t* out, uint64_t cap_words, void* stream);

/* Measurement: with timing enabled every rrx_scanner_scan records CUDA events on the launching stream
 * right before and after the scan kernel's launch; rrx_scanner_kernel_ms returns the duration of the
 * last one (the dominant kernel alone, without the memsets before and the ordering pass after it). */
int rrx_scanner_enable_timing(rrx_scanner* s, int on);
int rrx_scanner_kernel_ms(rrx_scanner* s, float* ms);

/* The same composite followed by the scan's counters, left in DEVICE memory (asynchronous): words32
 * [kind, b[W], U[W], P[W], 0, error_bits, allocator lo, hi, n_pending lo, hi, n_matches lo, hi].  This is the
 * message a rank contributes to the all-gather of a sharded scan: the collective can be enqueued on
 * the same stream right behind the scan, and ONE device-to-host copy of the gathered messages gives
 * every rank its own counters and everybody's composite (one synchronisation per step). */
int rrx_scanner_message_device(rrx_scanner* s, uint32_t* d_out, uint64_t cap_words, void* stream);

/* ---- sharded scans (one shard per GPU): host-side helpers, no device work ---------------------
 * A shard summary is the composite transfer function of the shard (rrx_scanner_shard_summary):
 * words32 [kind, b[W], U[W], P[W]].  Given the summaries of ALL shards in rank order (as returned
 * by an allgather), rrx_fold_entry computes the exact forward set entering shard `rank`.
 * *exact is 0 when some summary left of `rank` is opaque (only valid for the entry it was computed
 * with): the caller must then rescan that shard with its exact entry and fold again. */
int rrx_fold_entry(const rrx_program* p, const uint32_t* summaries, uint32_t n_shards, uint32_t rank,
                   uint32_t* entry_out, int* exact);
/* The same for every shard at once: entries_out[n_shards][W], exact_out[n_shards]. */
int rrx_fold_entries(const rrx_program* p, const uint32_t* summaries, uint32_t n_shards, uint32_t* entries_out, int* exact_out);
/* exit = summary applied to entry (both words32 per rrx_info.words). */
int rrx_apply_summary(const rrx_program* p, const uint32_t* summary, const uint32_t* entry, uint32_t* exit_out);
/* Pending start scans of the last scan (those that ran off the left edge of the readable text):
 * copies up to cap records {uint64 match_index; int64 next_offset; uint32 R[32]} to host.
 * next_offset is the GLOBAL offset (base_offset of that scan included) of the next byte to consume. */
int rrx_scanner_copy_pending(rrx_scanner* s, void* host_out, uint64_t cap, uint64_t* n_out, void* stream);
/* Continue start scans handed over by the shard to the right: for each record walk backwards from
 * record.next_offset, a GLOBAL offset that must lie inside this shard's readable text
 * [base_offset - left_context - 1, base_offset + len) (RRX_ERR_ARG otherwise).  On return start_out[i]
 * is the global start offset, -2 if no start exists (inconsistent record), or -1 if the walk ran off
 * this shard's left edge too (then the record is updated, still global, for the next shard to the left). */
int rrx_continue_starts(const rrx_program* p, const uint8_t* d_text, uint64_t len, uint64_t left_context,
                        uint64_t base_offset, void* records_host, uint64_t n, int64_t* start_out, void* stream);

/* Non-overlapping (finditer-style) view of the scanner's last list, with the carry that chains chunks and shards
 * (SURVEY.md Appendix D.2): *carry_io = global offset from which matches may start (0, or the end of the last match
 * selected in the text before this piece); on return the end of the last match selected here (unchanged if none).
 * The selected matches are a new device list (*out, rrx_matches_free).  Shards pass the carry from rank to rank. */
int rrx_scanner_no_overlap(rrx_scanner* s, int64_t* carry_io, rrx_matches** out, void* stream);

/* ---- sharded scans over NCCL (csrc/rrx_dist.cu): SURVEY.md §8b `rrx_scan_sharded`, §8e -------------------------
 * One contiguous byte range of the text per GPU, one process per GPU.  `nccl_comm` is the caller's ncclComm_t (passed
 * as void*; librrx loads libnccl.so.2 at run time and uses the library already in the process).  A step, per rank:
 * speculative scan of the shard, ONE ncclAllGather of [shard composite | counters | entry used] (a few hundred bytes
 * per rank over NVLink), a one-thread kernel that folds the composites of the ranks to the left into this rank's exact
 * entry set and flags, one small copy to pinned memory — all enqueued on `stream` by rrx_sharded_begin, no host
 * round trip.  rrx_sharded_finish waits for that copy; if no rank needs more (no state crosses a shard boundary, no
 * start walk leaves its halo) the speculative list is the result.  Otherwise all ranks together rescan with their
 * exact entries (rounds of all-gather + fold) and hand start walks to the left neighbour; finish returns when done.
 * Two buffer sets: begin(k+1) may be called before finish(k), so the host side of one step overlaps the next scan.
 * left_context: bytes readable before d_text (the halo: the tail of the shard to the left), for start and
 * verification walks that cross the shard's left edge.  Output stays sharded: rank order = text order. */
typedef struct rrx_sharded rrx_sharded;
typedef struct rrx_sharded_result {
    uint64_t n_matches;         /* exact matches of this rank's shard */
    const int64_t* d_matches;   /* device pairs (start,end), ascending end; valid until the buffer set is reused (two steps later) */
    uint32_t rescans;           /* whole-shard rescans this step needed on this rank (0 in the common case) */
    uint32_t patched_starts;    /* starts found by ranks to the left and patched into the list */
    uint32_t flags;             /* internal status bits of the last exchange */
    uint32_t reserved;
    rrx_scanner* scanner;       /* the scanner that holds the list (for rrx_scanner_digest / copy_matches) */
} rrx_sharded_result;
int rrx_sharded_create(const rrx_program* p, uint64_t max_len, uint64_t match_capacity, void* nccl_comm, int rank, int nranks,
                       rrx_sharded** out);
int rrx_sharded_begin(rrx_sharded* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset, uint64_t left_context, void* stream);
int rrx_sharded_finish(rrx_sharded* s, rrx_sharded_result* out);
/* begin + finish */
int rrx_scan_sharded(rrx_sharded* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset, uint64_t left_context, void* stream,
                     rrx_sharded_result* out);
int rrx_sharded_enable_timing(rrx_sharded* s, int on);
float rrx_sharded_kernel_ms(rrx_sharded* s);          /* scan kernel of the last finished step (see rrx_scanner_kernel_ms) */
void rrx_sharded_free(rrx_sharded* s);

/* ---- parity at scale: digest of a device match list (csrc/rrx_hash.cu) --------------------------
 * digest[0] = matches with end <= end_limit, [1] = sum and [2] = xor over those of
 * sm64(sm64(start) + end) (splitmix64 finaliser), [3] = order violations in the WHOLE list
 * (end not ascending, or start >= end; must be 0).  Sum/xor are order-free: shard digests add/xor up.
 * Synchronous (waits for `stream`). */
int rrx_digest_device(const int64_t* d_pairs, uint64_t n, int64_t end_limit, uint64_t digest[4], void* stream);
/* The same for the match list of the scanner's last scan. */
int rrx_scanner_digest(rrx_scanner* s, int64_t end_limit, uint64_t digest[4], void* stream);

/* ---- measurement harness: synthetic corpora of SURVEY.md §8d, generated in device memory ------ */
int rrx_corpus_iid_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, const uint8_t alpha[256], void* stream);
int rrx_corpus_records_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, void* stream);
/* config 2 variant: the same log lines unpadded (variable length, unaligned); oracle/corpus.c corpus_fill_lines */
int rrx_corpus_lines_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, void* stream);

const char* rrx_last_error(void);
const char* rrx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RRX_H_ */
