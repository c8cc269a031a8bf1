/* TEST / MEASUREMENT INFRASTRUCTURE — synthetic corpus generator, CPU side (see corpus.h). */
#include "corpus.h"
#include <string.h>

uint64_t corpus_sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t corpus_h(uint64_t seed, uint64_t w) { return corpus_sm64(corpus_sm64(seed) + w); }

void corpus_fill_iid(uint64_t seed, const uint8_t alpha[256], uint64_t offset, uint64_t len, uint8_t* out) {
    uint64_t s = corpus_sm64(seed);
    uint64_t o = offset, end = offset + len;
    while (o < end) {
        uint64_t hw = corpus_sm64(s + (o >> 3));
        do {
            *out++ = alpha[(hw >> (8 * (o & 7))) & 255];
            o++;
        } while (o < end && (o & 7));
    }
}

static int put_dec(uint8_t* p, uint64_t v, int digits) {
    for (int i = digits - 1; i >= 0; i--) { p[i] = (uint8_t)('0' + v % 10); v /= 10; }
    return digits;
}
static int put_str(uint8_t* p, const char* s) {
    int n = 0;
    while (s[n]) { p[n] = (uint8_t)s[n]; n++; }
    return n;
}
static int put_letters(uint8_t* p, uint64_t hl, uint64_t hc) {
    int n = 1 + (int)(hl % 12);
    for (int i = 0; i < n; i++) p[i] = (uint8_t)('a' + ((hc >> (5 * i)) & 31) % 26);
    return n;
}

/* One 128-byte record:
 *   "%010u INFO svc%02u user=" local "@" domain "." tld " rc=%03u", space-padded to 127, then '\n'
 * field k of record r is drawn from h(seed, 8r+k):
 *   k0 timestamp (low 32 bits)      k1 service id (% 100)
 *   k2 kind (&3): 0,1 valid address; 2 near miss (must NOT match); 3 no address ("user=-")
 *   k3 local-part length / k4 its letters      k5 domain length / k6 its letters
 *   k7 tld (%3: com org net), near-miss variant ((>>8)%5: con ogr ne. COM, or '#' for '@'),
 *      return code ((>>32)%1000) */
void corpus_record(uint64_t seed, uint64_t rec, uint8_t out[128]) {
    static const char* tld_ok[3] = {"com", "org", "net"};
    static const char* tld_bad[4] = {"con", "ogr", "ne.", "COM"};
    uint64_t s = corpus_sm64(seed), f[8];
    for (int k = 0; k < 8; k++) f[k] = corpus_sm64(s + 8 * rec + (uint64_t)k);
    memset(out, ' ', 127);
    out[127] = '\n';
    int p = 0;
    p += put_dec(out + p, f[0] & 0xFFFFFFFFull, 10);
    p += put_str(out + p, " INFO svc");
    p += put_dec(out + p, f[1] % 100, 2);
    p += put_str(out + p, " user=");
    unsigned kind = (unsigned)(f[2] & 3);
    if (kind == 3) {
        out[p++] = '-';
    } else {
        unsigned variant = (unsigned)((f[7] >> 8) % 5);
        p += put_letters(out + p, f[3], f[4]);
        out[p++] = (kind == 2 && variant == 4) ? '#' : '@';
        p += put_letters(out + p, f[5], f[6]);
        out[p++] = '.';
        if (kind == 2 && variant < 4) p += put_str(out + p, tld_bad[variant]);
        else p += put_str(out + p, tld_ok[f[7] % 3]);
    }
    p += put_str(out + p, " rc=");
    p += put_dec(out + p, (f[7] >> 32) % 1000, 3);
}

void corpus_fill_records(uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out) {
    uint8_t rec[128];
    uint64_t o = offset, end = offset + len;
    while (o < end) {
        uint64_t r = o >> 7, col = o & 127;
        uint64_t take = 128 - col;
        if (take > end - o) take = end - o;
        corpus_record(seed, r, rec);
        memcpy(out, rec + col, take);
        out += take;
        o += take;
    }
}

/* "lines" mode (config 2 variant, VERDICT W10): the same log lines WITHOUT padding — variable length
 * (39..254 bytes), not aligned to anything — so that addresses straddle the scanner's 128-byte rows.
 * Random access by 2048-byte blocks: block b holds lines drawn from h(seed, 1024*b + 16*i + k) for its
 * i-th line (k0..k7 as in record mode, k8 message length, k9 message letters); lines are laid out back
 * to back while at least 256 bytes are left, the rest of the block is spaces and a '\n'. */
static const char LINES_MSG[33] = "abcdefghijklmnopqrstuvwxyz 0123 ";
static int lines_make(uint64_t s, uint64_t blk, int i, uint8_t* out, int* valid) {
    static const char* tld_ok[3] = {"com", "org", "net"};
    static const char* tld_bad[4] = {"con", "ogr", "ne.", "COM"};
    uint64_t f[10];
    for (int k = 0; k < 10; k++) f[k] = corpus_sm64(s + 1024 * blk + 16 * (uint64_t)i + (uint64_t)k);
    int p = 0;
    p += put_dec(out + p, f[0] & 0xFFFFFFFFull, 10);
    p += put_str(out + p, " INFO svc");
    p += put_dec(out + p, f[1] % 100, 2);
    p += put_str(out + p, " user=");
    unsigned kind = (unsigned)(f[2] & 3);
    *valid = kind < 2;
    if (kind == 3) {
        out[p++] = '-';
    } else {
        unsigned variant = (unsigned)((f[7] >> 8) % 5);
        p += put_letters(out + p, f[3], f[4]);
        out[p++] = (kind == 2 && variant == 4) ? '#' : '@';
        p += put_letters(out + p, f[5], f[6]);
        out[p++] = '.';
        if (kind == 2 && variant < 4) p += put_str(out + p, tld_bad[variant]);
        else p += put_str(out + p, tld_ok[f[7] % 3]);
    }
    p += put_str(out + p, " rc=");
    p += put_dec(out + p, (f[7] >> 32) % 1000, 3);
    p += put_str(out + p, " m=");
    int L = (int)(f[8] % 188);
    for (int j = 0; j < L; j++) out[p++] = (uint8_t)LINES_MSG[(corpus_sm64(f[9] + (uint64_t)(j / 12)) >> (5 * (j % 12))) & 31];
    out[p++] = '\n';
    return p;
}
static uint64_t lines_block(uint64_t s, uint64_t blk, uint8_t out[2048]) {
    int pos = 0, i = 0, valid;
    uint64_t nvalid = 0;
    while (2048 - pos >= 256) { pos += lines_make(s, blk, i++, out + pos, &valid); nvalid += (uint64_t)valid; }
    if (pos < 2048) { memset(out + pos, ' ', (size_t)(2047 - pos)); out[2047] = '\n'; }
    return nvalid;
}
void corpus_fill_lines(uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out) {
    uint8_t blk[2048];
    uint64_t s = corpus_sm64(seed), o = offset, end = offset + len;
    while (o < end) {
        uint64_t b = o >> 11, col = o & 2047, take = 2048 - col;
        if (take > end - o) take = end - o;
        lines_block(s, b, blk);
        memcpy(out, blk + col, take);
        out += take;
        o += take;
    }
}
/* closed-form match count of `[a-z]+@[a-z]+\.(com|org|net)` over the first n_blocks blocks: one per valid line */
uint64_t corpus_lines_valid(uint64_t seed, uint64_t n_blocks) {
    uint8_t blk[2048];
    uint64_t s = corpus_sm64(seed), n = 0;
    for (uint64_t b = 0; b < n_blocks; b++) n += lines_block(s, b, blk);
    return n;
}
/* closed-form match count over the first n_records 128-byte records (record mode): kinds 0 and 1 */
uint64_t corpus_records_valid(uint64_t seed, uint64_t n_records) {
    uint64_t s = corpus_sm64(seed), n = 0;
    for (uint64_t r = 0; r < n_records; r++) n += (corpus_sm64(s + 8 * r + 2) & 3) < 2;
    return n;
}
