/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's hot path + the LAZY-v1 enumerator.
 *
 * Works on the reference's RAW transition tables (built by oracle/ref_nfa.py in the layout of
 * /root/reference/src/NFA.cc:9-12, full-width state index) and follows the reference's own
 * algorithm step by step:
 *   - step          NFA.cc:72-102  Processor::shift<fwd>: new = OR over active states i of table[idx(i,c,fwd)]
 *                   (BitSet path, NFA.cc:87-99; set-bit iteration as BitSet.cc:72-83)
 *   - flag word     NFA.cc:103-107 bit0 = set ∩ final ≠ ∅ (FLAG_ACCEPTING), bit1 = initial ∈ set (FLAG_INITIAL)
 *   - initial set   regex.h:135-143 {initial_state}
 *   - driver shape  regex.h:156-162 one step per text byte, Match = half-open [start,end)
 *   - enumeration   SURVEY.md Appendix D.1 ("LAZY-v1"): re-inject the initial state before every byte;
 *                   whenever FLAG_ACCEPTING is raised walk the BACKWARD tables (NFA.cc:11) from the
 *                   final set until FLAG_INITIAL is raised for the first time; emit (start,end).
 * The reference ships no match enumerator (src/Iterator.cpp:1-5 is empty): the enumeration rule is
 * PARITY-UNPINNED by the reference; its building blocks (construction, step, acceptance) are pinned
 * against the compiled reference in tests/test_oracle_pins.py.
 *
 * Bytes >= 0x80 match nothing (the reference indexes out of bounds there, README.md:50 "ASCII only").
 * Nothing under roaringregex_b200/ may link, load or call this file.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    uint32_t n;            /* raw states */
    uint32_t W;            /* uint64 words per set */
    uint32_t init;         /* initial state */
    const uint64_t* table; /* [128*2*n][W], row = state + n*(2*byte + dir) */
    const uint64_t* finals;/* [W] */
} raw_nfa;

static inline const uint64_t* row(const raw_nfa* a, uint32_t state, uint32_t c, int bwd) {
    return a->table + ((size_t)state + (size_t)a->n * (2u * c + (bwd ? 1u : 0u))) * a->W;
}

/* one Processor::shift: dst = OR_{i in src} table[idx(i,c,dir)] */
static inline void shift(const raw_nfa* a, const uint64_t* src, uint64_t* dst, uint32_t c, int bwd) {
    uint32_t W = a->W;
    for (uint32_t w = 0; w < W; w++) dst[w] = 0;
    if (c >= 128) return;
    for (uint32_t w = 0; w < W; w++) {
        uint64_t v = src[w];
        while (v) {
            uint32_t i = w * 64 + (uint32_t)__builtin_ctzll(v);
            v &= v - 1;
            const uint64_t* r = row(a, i, c, bwd);
            for (uint32_t k = 0; k < W; k++) dst[k] |= r[k];
        }
    }
}
static inline int intersects(const uint64_t* x, const uint64_t* y, uint32_t W) {
    for (uint32_t w = 0; w < W; w++) if (x[w] & y[w]) return 1;
    return 0;
}
static inline int contains(const uint64_t* x, uint32_t s) { return (int)((x[s >> 6] >> (s & 63)) & 1); }

/* AcceptanceIterator: anchored whole-buffer membership (no re-injection). */
int oracle_accepts(const raw_nfa* a, const uint8_t* text, uint64_t len) {
    uint32_t W = a->W;
    uint64_t* cur = calloc(2 * (size_t)W, 8);
    uint64_t* nxt = cur + W;
    cur[a->init >> 6] |= 1ull << (a->init & 63);
    for (uint64_t i = 0; i < len; i++) {
        shift(a, cur, nxt, text[i], 0);
        uint64_t* t = cur; cur = nxt; nxt = t;
    }
    int r = intersects(cur, a->finals, W);
    free(cur < nxt ? cur : nxt);
    return r;
}

/* LAZY-v1 over text[0,len).  `entry` (nullable, W words) is the forward set carried in from
 * text before offset 0 (sharded use); backward scans stop at offset 0 — a scan that would have
 * to continue to the left of the buffer is reported with start = -1 (callers that shard must
 * give enough left context).  Offsets written are base + local offset.
 * Writes up to cap (start,end) pairs into out; returns the total number of matches.
 * exit_state (nullable, W words) receives the forward set after the last byte. */
uint64_t oracle_lazy_scan(const raw_nfa* a, const uint8_t* text, uint64_t len, int64_t base,
                          const uint64_t* entry, int64_t* out, uint64_t cap, uint64_t* exit_state) {
    uint32_t W = a->W;
    uint64_t* buf = calloc(4 * (size_t)W, 8);
    uint64_t *S = buf, *T = buf + W, *B = buf + 2 * W, *B2 = buf + 3 * W;
    uint64_t count = 0;
    if (entry) memcpy(S, entry, 8 * (size_t)W);
    for (uint64_t i = 0; i < len; i++) {
        S[a->init >> 6] |= 1ull << (a->init & 63);       /* the one addition: unanchored search */
        shift(a, S, T, text[i], 0);
        { uint64_t* t = S; S = T; T = t; }
        if (intersects(S, a->finals, W)) {               /* FLAG_ACCEPTING */
            memcpy(B, a->finals, 8 * (size_t)W);
            int64_t j = (int64_t)i + 1;
            int found = 0;
            while (j > 0) {
                j--;
                shift(a, B, B2, text[j], 1);
                { uint64_t* t = B; B = B2; B2 = t; }
                if (contains(B, a->init)) { found = 1; break; }   /* FLAG_INITIAL: first hit = lazy */
            }
            if (count < cap) {
                out[2 * count] = found ? base + j : -1;
                out[2 * count + 1] = base + (int64_t)i + 1;
            }
            count++;
        }
    }
    if (exit_state) memcpy(exit_state, S, 8 * (size_t)W);
    free(buf);
    return count;
}

/* ---- timing helpers for bench.py's cpu_baseline / reference arm -------------------------------- */
typedef struct {
    const raw_nfa* a; const uint8_t* text; uint64_t len; uint64_t matches;
} job_t;
static void* job_run(void* p) {
    job_t* j = (job_t*)p;
    j->matches = oracle_lazy_scan(j->a, j->text, j->len, 0, NULL, NULL, 0, NULL);
    return NULL;
}
/* Scan `nthreads` independent contiguous pieces text[t*piece, (t+1)*piece) concurrently, one per
 * thread, each from an empty entry set (a throughput baseline, not a parity run).  Returns
 * seconds of wall time; *total_matches gets the sum. */
/* Non-overlapping view of a LAZY-v1 list: p = base; for (s,e): if s >= p { emit; p = e }.
 * Equivalent to resetting the forward set to the empty set after each reported match.  In place; returns the new length. */
uint64_t oracle_no_overlap(int64_t* pairs, uint64_t n, int64_t base) {
    uint64_t k = 0;
    int64_t p = base;
    for (uint64_t i = 0; i < n; i++) {
        if (pairs[2 * i] >= p) {
            pairs[2 * k] = pairs[2 * i];
            pairs[2 * k + 1] = pairs[2 * i + 1];
            p = pairs[2 * i + 1];
            k++;
        }
    }
    return k;
}

double oracle_time_parallel(const raw_nfa* a, const uint8_t* text, uint64_t piece, int nthreads,
                            uint64_t* total_matches) {
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = malloc(sizeof(job_t) * (size_t)nthreads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        jobs[t].a = a; jobs[t].text = text + (uint64_t)t * piece; jobs[t].len = piece; jobs[t].matches = 0;
        pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    uint64_t m = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); m += jobs[t].matches; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_matches) *total_matches = m;
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
