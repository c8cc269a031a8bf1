/* TEST / MEASUREMENT INFRASTRUCTURE — synthetic corpus generator, CPU side.
 * The device generator in roaringregex_b200/csrc/rrx_corpus.cu must produce the same bytes;
 * tests/test_gpu_parity.py compares them.  Random access: any byte range can be produced in place. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* splitmix64 finaliser and the two-level hash h(seed,w) = sm(sm(seed)+w) */
uint64_t corpus_sm64(uint64_t z);
uint64_t corpus_h(uint64_t seed, uint64_t w);

/* iid mode: byte at offset o = alpha[(h(seed, o>>3) >> 8*(o&7)) & 255] */
void corpus_fill_iid(uint64_t seed, const uint8_t alpha[256], uint64_t offset, uint64_t len, uint8_t* out);

/* record mode (config 2): fixed 128-byte log records, see corpus.c */
void corpus_fill_records(uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out);
void corpus_record(uint64_t seed, uint64_t rec, uint8_t out[128]);

/* lines mode (config 2 variant): the same log lines unpadded, variable length, see corpus.c */
void corpus_fill_lines(uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out);
uint64_t corpus_lines_valid(uint64_t seed, uint64_t n_blocks);
uint64_t corpus_records_valid(uint64_t seed, uint64_t n_records);

#ifdef __cplusplus
}
#endif
