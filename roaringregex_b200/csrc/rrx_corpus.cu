// Measurement harness: the synthetic corpora of SURVEY.md §8d generated directly in HBM.
// Must produce the same bytes as oracle/corpus.c (tests/test_gpu_parity.py::test_device_corpus_equals_cpu_corpus compares them).
#include "rrx_scan.h"

namespace rrx {
namespace {

__host__ __device__ inline uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Alpha { uint8_t a[256]; };

// iid mode: byte at offset o = alpha[(h(seed, o>>3) >> 8*(o&7)) & 255]; one 8-byte word per thread.
__global__ void corpus_iid_kernel(uint8_t* out, uint64_t offset, uint64_t len, uint64_t sseed, const __grid_constant__ Alpha al) {
    __shared__ uint8_t a[256];
    a[threadIdx.x & 255] = al.a[threadIdx.x & 255];
    __syncthreads();
    const uint64_t w0 = offset >> 3, w1 = (offset + len + 7) >> 3;
    for (uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t hw = sm64(sseed + w);
        uint64_t packed = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) packed |= (uint64_t)a[(hw >> (8 * k)) & 255] << (8 * k);
        const uint64_t o = w << 3;
        if (o >= offset && o + 8 <= offset + len && (((uintptr_t)(out + (o - offset))) & 7) == 0) {
            *reinterpret_cast<uint64_t*>(out + (o - offset)) = packed;
        } else {
            for (int k = 0; k < 8; k++) {
                const uint64_t ok = o + k;
                if (ok >= offset && ok < offset + len) out[ok - offset] = (uint8_t)(packed >> (8 * k));
            }
        }
    }
}

__device__ inline int put_dec(uint8_t* p, uint64_t v, int digits) {
    for (int i = digits - 1; i >= 0; i--) { p[i] = (uint8_t)('0' + v % 10); v /= 10; }
    return digits;
}
__device__ inline int put_str(uint8_t* p, const char* s) {
    int n = 0;
    while (s[n]) { p[n] = (uint8_t)s[n]; n++; }
    return n;
}
__device__ inline int put_letters(uint8_t* p, uint64_t hl, uint64_t hc) {
    const int n = 1 + (int)(hl % 12);
    for (int i = 0; i < n; i++) p[i] = (uint8_t)('a' + ((hc >> (5 * i)) & 31) % 26);
    return n;
}

// record mode (config 2): one 128-byte log record per thread; layout documented in oracle/corpus.c.
__global__ void corpus_records_kernel(uint8_t* out, uint64_t offset, uint64_t len, uint64_t sseed) {
    const uint64_t r0 = offset >> 7, r1 = (offset + len + 127) >> 7;
    for (uint64_t r = r0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t f[8];
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = sm64(sseed + 8 * r + (uint64_t)k);
        uint8_t rec[128];
        for (int i = 0; i < 127; i++) rec[i] = ' ';
        rec[127] = '\n';
        int p = 0;
        p += put_dec(rec + p, f[0] & 0xFFFFFFFFull, 10);
        p += put_str(rec + p, " INFO svc");
        p += put_dec(rec + p, f[1] % 100, 2);
        p += put_str(rec + p, " user=");
        const unsigned kind = (unsigned)(f[2] & 3);
        if (kind == 3) {
            rec[p++] = '-';
        } else {
            const unsigned variant = (unsigned)((f[7] >> 8) % 5);
            p += put_letters(rec + p, f[3], f[4]);
            rec[p++] = (kind == 2 && variant == 4) ? '#' : '@';
            p += put_letters(rec + p, f[5], f[6]);
            rec[p++] = '.';
            if (kind == 2 && variant < 4) {
                const char* bad = variant == 0 ? "con" : variant == 1 ? "ogr" : variant == 2 ? "ne." : "COM";
                p += put_str(rec + p, bad);
            } else {
                const unsigned t = (unsigned)(f[7] % 3);
                p += put_str(rec + p, t == 0 ? "com" : t == 1 ? "org" : "net");
            }
        }
        p += put_str(rec + p, " rc=");
        p += put_dec(rec + p, (f[7] >> 32) % 1000, 3);
        const uint64_t o = r << 7;
        for (int i = 0; i < 128; i++) {
            const uint64_t oi = o + i;
            if (oi >= offset && oi < offset + len) out[oi - offset] = rec[i];
        }
    }
}

// lines mode (config 2 variant): unpadded variable-length log lines, one 2048-byte block per thread; layout
// documented in oracle/corpus.c (corpus_fill_lines).
__global__ void corpus_lines_kernel(uint8_t* out, uint64_t offset, uint64_t len, uint64_t sseed) {
    const char* MSG = "abcdefghijklmnopqrstuvwxyz 0123 ";
    const uint64_t b0 = offset >> 11, b1 = (offset + len + 2047) >> 11;
    for (uint64_t b = b0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = b << 11;
        int pos = 0, i = 0;
        uint8_t rec[256];
        while (2048 - pos >= 256) {
            uint64_t f[10];
#pragma unroll
            for (int k = 0; k < 10; k++) f[k] = sm64(sseed + 1024 * b + 16 * (uint64_t)i + (uint64_t)k);
            i++;
            int p = 0;
            p += put_dec(rec + p, f[0] & 0xFFFFFFFFull, 10);
            p += put_str(rec + p, " INFO svc");
            p += put_dec(rec + p, f[1] % 100, 2);
            p += put_str(rec + p, " user=");
            const unsigned kind = (unsigned)(f[2] & 3);
            if (kind == 3) {
                rec[p++] = '-';
            } else {
                const unsigned variant = (unsigned)((f[7] >> 8) % 5);
                p += put_letters(rec + p, f[3], f[4]);
                rec[p++] = (kind == 2 && variant == 4) ? '#' : '@';
                p += put_letters(rec + p, f[5], f[6]);
                rec[p++] = '.';
                if (kind == 2 && variant < 4) {
                    const char* bad = variant == 0 ? "con" : variant == 1 ? "ogr" : variant == 2 ? "ne." : "COM";
                    p += put_str(rec + p, bad);
                } else {
                    const unsigned t = (unsigned)(f[7] % 3);
                    p += put_str(rec + p, t == 0 ? "com" : t == 1 ? "org" : "net");
                }
            }
            p += put_str(rec + p, " rc=");
            p += put_dec(rec + p, (f[7] >> 32) % 1000, 3);
            p += put_str(rec + p, " m=");
            const int L = (int)(f[8] % 188);
            for (int j = 0; j < L; j++) rec[p++] = (uint8_t)MSG[(sm64(f[9] + (uint64_t)(j / 12)) >> (5 * (j % 12))) & 31];
            rec[p++] = '\n';
            for (int k = 0; k < p; k++) {
                const uint64_t ok = o + (uint64_t)(pos + k);
                if (ok >= offset && ok < offset + len) out[ok - offset] = rec[k];
            }
            pos += p;
        }
        for (; pos < 2048; pos++) {
            const uint64_t ok = o + (uint64_t)pos;
            if (ok >= offset && ok < offset + len) out[ok - offset] = pos == 2047 ? (uint8_t)'\n' : (uint8_t)' ';
        }
    }
}

}  // namespace

cudaError_t launch_corpus_lines(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, cudaStream_t st) {
    if (!len) return cudaSuccess;
    corpus_lines_kernel<<<148 * 8, 64, 0, st>>>(d_out, offset, len, sm64(seed));
    return cudaGetLastError();
}
cudaError_t launch_corpus_iid(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, const uint8_t* alpha256, cudaStream_t st) {
    if (!len) return cudaSuccess;
    Alpha al;
    for (int i = 0; i < 256; i++) al.a[i] = alpha256[i];
    corpus_iid_kernel<<<148 * 8, 256, 0, st>>>(d_out, offset, len, sm64(seed), al);
    return cudaGetLastError();
}
cudaError_t launch_corpus_records(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, cudaStream_t st) {
    if (!len) return cudaSuccess;
    corpus_records_kernel<<<148 * 8, 128, 0, st>>>(d_out, offset, len, sm64(seed));
    return cudaGetLastError();
}

}  // namespace rrx
