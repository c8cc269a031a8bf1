// C ABI of librrx.so (include/rrx.h).  Host C++ around the sm_100a kernels; no CPU scan path.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rrx.h"
#include "rrx_compile.h"
#include "rrx_jit.h"
#include "rrx_scan.h"
#include "rrx_tile.cuh"

using namespace rrx;

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace
int rrx_dist_set_err(int code, const std::string& msg) { return set_err(code, msg); }      // for rrx_dist.cu
namespace {
int cuda_err(cudaError_t e, const char* what) {
    return set_err(RRX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define RRX_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_err(e_, #call); } while (0)

// Pending-start slots for scanners created by this thread, when larger than the default: the one-shot calls on anchored
// programs set it after a scan whose start walks reached the text start more often than the default list holds
// (`^.*b` on a text full of `b`), create the scanner (and, through it, its fallback scanner) again and rescan.
thread_local uint32_t t_pending_cap_want = 0;
struct PendingCapScope {
    explicit PendingCapScope(uint32_t want) { t_pending_cap_want = want; }
    ~PendingCapScope() { t_pending_cap_want = 0; }
};
constexpr uint64_t PENDING_CAP_MAX = 1ull << 24;

// cudaMemcpy from pageable host memory returns once the bytes sit in the driver's staging buffer: the DMA into device memory
// runs on the legacy default stream and may still be in flight.  The scans run on non-blocking streams, which do not wait
// for that stream — so an upload that a scan depends on is followed by a synchronisation of the legacy stream.
// (Seen as an empty match list from the first rrx_scan_host call of a program on a busy GPU: scripts/fuzz_gpu.py.)
static cudaError_t upload_sync(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e;
}

// Control block of a scan over T tiles (B = ceil(T/32) blocks); everything in it is zeroed before
// the scan: [-,err,total[2],n_matches | pad | ticket (its own 128-byte line: every tile of the grid adds to it, nothing
// else should wait behind those atomics) | f_a0[T] f_i0[T] f_a1[B] f_i1[B] st_arrive[B] tile_stop[T]].
// Returns the size; fills the five offsets.
size_t ctrl_layout(size_t T, size_t* off) {
    const size_t B = (T + 31) / 32;
    size_t o = 256, tmp[8];
    size_t* f = off ? off : tmp;
    f[0] = o; o += 4 * T;
    f[1] = o; o += 4 * T;
    f[2] = o; o += 4 * B;
    f[3] = o; o += 4 * B;
    f[4] = o; o += 4 * B;
    f[5] = o; o += 4 * T;        // stop-byte table (ScanParams::tile_stop)
    return (o + 15) & ~(size_t)15;
}
// Non-overlapping view (RRX_MODE_NO_OVERLAP): filter a device match list of n pairs into a newly
// allocated device list.  *out_pairs is owned by the caller (cudaFree).
int filter_no_overlap(const long long* d_in, uint64_t n, long long** out_pairs, uint64_t* out_n, cudaStream_t st) {
    *out_pairs = nullptr;
    *out_n = 0;
    if (n == 0) return RRX_OK;
    const size_t wsb = no_overlap_workspace_bytes(n);
    void* ws = nullptr;
    long long* d_out = nullptr;
    unsigned long long* d_cnt = nullptr;
    cudaError_t e = cudaMalloc(&ws, wsb);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 8);
    if (e == cudaSuccess) e = launch_no_overlap(reinterpret_cast<const longlong2*>(d_in), n, reinterpret_cast<longlong2*>(d_out), d_cnt, ws, wsb, st);
    unsigned long long cnt = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(ws);
    cudaFree(d_cnt);
    if (e != cudaSuccess) { cudaFree(d_out); return cuda_err(e, "non-overlap filter"); }
    *out_pairs = d_out;
    *out_n = cnt;
    return RRX_OK;
}
}  // namespace

// Device-side state kept between rrx_scan_host calls on one program: the text buffer, a scanner
// sized for it, a private stream and a pool of pinned match buffers.  A host scan is then
// H2D copy -> kernels -> D2H copy with no allocation on the way (after the first call).
struct HostCtx {
    std::mutex mu;
    int device = -1;
    cudaStream_t st = nullptr;
    uint8_t* d_text = nullptr;
    uint64_t text_cap = 0;
    rrx_scanner* sc = nullptr;          // freed by rrx_program_free (it refers to the program)
    uint64_t sc_len = 0, sc_cap = 0;
    struct Pinned { int64_t* p; uint64_t pairs; };
    std::vector<Pinned> pool;           // free pinned buffers
    // streaming ingestion (rrx_scan_host on long texts): a copy stream, one event per chunk, a second scanner sized for
    // one chunk, and pinned bounce buffers for pageable input
    cudaStream_t st_copy = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    rrx_scanner* sc_chunk = nullptr;
    uint64_t scc_len = 0, scc_cap = 0;
    uint8_t* bounce[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_bounce[3] = {nullptr, nullptr, nullptr};
    uint64_t bounce_bytes = 0;
    ~HostCtx() {
        cudaFree(d_text);
        for (auto& b : pool) cudaFreeHost(b.p);
        for (auto e : ev_chunk) cudaEventDestroy(e);
        for (int i = 0; i < 3; i++) { if (bounce[i]) cudaFreeHost(bounce[i]); if (ev_bounce[i]) cudaEventDestroy(ev_bounce[i]); }
        if (st_copy) cudaStreamDestroy(st_copy);
        if (st) cudaStreamDestroy(st);
    }
};

struct rrx_program {
    Automaton automaton;   // reference-equivalent view (atoms, flattened positions)
    Program prog;          // device program
    uint32_t flags = 0;    // rrx_compile_ex flags
    std::string pattern;
    // the same regex compiled WITHOUT the thread-owned candidate scan: what a scan falls back to when it is given
    // an entry set outside D, or when a verification walk runs out of budget (built on first use)
    std::mutex alt_mu;
    rrx_program* alt = nullptr;
    std::shared_ptr<HostCtx> host = std::make_shared<HostCtx>();
};

struct rrx_scanner {
    const rrx_program* p = nullptr;
    int device = 0, sm_count = 0;
    uint64_t max_len = 0, capacity = 0, max_tiles = 0;
    uint32_t NW = 1;
    uint32_t* d_prog = nullptr;
    uint8_t* d_ctrl = nullptr;      // header + look-back flags and counters, zeroed before every scan
    size_t ctrl_bytes = 0;
    uint32_t* d_tpay = nullptr;
    long long* d_out = nullptr;      // matches in text order
    long long* d_stage = nullptr;    // matches in tile-completion order (scan kernel output)
    unsigned long long* d_tile_seg = nullptr;
    uint32_t* d_tile_cnt = nullptr;
    unsigned long long* d_chunk = nullptr;
    uint32_t* d_shard = nullptr;
    PendingStart* d_pending = nullptr;
    uint32_t pending_cap = 4096;
    uint32_t launches = 0;
    int64_t last_base = 0;           // base_offset of the last scan (pending records are handed out with global offsets)
    // fallback scanner on p->alt (see rrx_program::alt), created on first use; `use_alt`: the last scan ran there
    rrx_scanner* alt = nullptr;
    bool use_alt = false;
    uint32_t create_flags = 0;
    struct { const uint8_t* text; uint64_t len, base, left_ctx; bool has_entry; uint32_t entry[RRX_NW_MAX]; void* stream; } last = {};
    bool owns_out = true;
    JitKernel* jit = nullptr;        // run-time specialised kernel (owned by the process-wide cache)
    // Thread-owned programs have two specialised kernels (rrx_jit.h): lean and straddle.  A scanner starts lean; every
    // rrx_scanner_result looks at how many tiles left the fast tail (lean) / how many rows had their run followed into the
    // next row (straddle) and picks the kernel of the NEXT scan: above 1/8 of the tiles -> straddle, below 1/64 -> lean.
    JitKernel* jit_v[2] = {nullptr, nullptr};
    int variant = 0;
    uint32_t last_tiles = 0;
    // pinned host mirror of [ticket, err, total0, total1] followed by the shard composite (1+3*NW words)
    uint64_t* h_ctr = nullptr;
    bool h_shard_valid = false;     // the composite in h_ctr belongs to the last scan
    // optional: CUDA events around the scan kernel launch alone (measurement)
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct rrx_matches {
    uint64_t count = 0;
    long long* d_pairs = nullptr;   // owned (device scans)
    std::vector<int64_t> host;
    bool host_valid = false;
    // host scans: the list lives in a pinned buffer borrowed from the program's pool
    std::shared_ptr<HostCtx> owner;
    HostCtx::Pinned pinned = {nullptr, 0};
};

extern "C" {

const char* rrx_last_error(void) { return g_err.c_str(); }
const char* rrx_version(void) { return "rrx 0.1 (sm_100a)"; }

int rrx_compile_ex(const char* ere, uint32_t flags, rrx_program** out) {
    if (!ere || !out) return set_err(RRX_ERR_ARG, "null argument");
    try {
        std::unique_ptr<rrx_program> p(new rrx_program);
        p->automaton = compile_ere(ere);
        p->prog = build_program(p->automaton, ere, !(flags & RRX_COMPILE_NO_REDUCE), !(flags & RRX_COMPILE_NO_DONLY));
        p->flags = flags;
        p->pattern = ere;
        *out = p.release();
        return RRX_OK;
    } catch (const std::runtime_error& e) {
        return set_err(RRX_ERR_PARSE, e.what());
    } catch (const std::bad_alloc&) {
        return set_err(RRX_ERR_ARG, "out of memory");
    }
}
int rrx_compile(const char* ere, rrx_program** out) { return rrx_compile_ex(ere, RRX_COMPILE_DEFAULT, out); }
void rrx_program_free(rrx_program* p) {
    if (!p) return;
    if (p->alt) { rrx_program_free(p->alt); p->alt = nullptr; }
    if (p->host && p->host->sc) { rrx_scanner_free(p->host->sc); p->host->sc = nullptr; }
    if (p->host && p->host->sc_chunk) { rrx_scanner_free(p->host->sc_chunk); p->host->sc_chunk = nullptr; }
    delete p;
}

void rrx_program_info(const rrx_program* p, rrx_info* o) {
    const ProgHeader& h = p->prog.hdr();
    o->n_raw = h.n_raw;
    o->ref_width_class = h.ref_width_class;
    o->n_flat = h.n_flat;
    o->n_device = h.P;
    o->words = h.NW;
    o->nullable = h.nullable;
    o->n_terms = h.n_terms;
    o->n_bicliques = h.n_bicl;
    o->uses_rows = h.use_rows;
    o->static_closed = (h.static_depth != 0xffffffffu && h.static_depth + 1 <= (uint32_t)ROW_BYTES) ? 1u : 0u;
    o->sparse_ok = h.sparse_ok;
    o->cand_byte = h.cand_byte;
    o->donly = h.donly;
    o->carried_words = h.donly ? h.CW : h.NW;
}

int64_t rrx_program_flat_table(const rrx_program* p, int which, uint32_t* out, uint64_t cap) {
    const Automaton& a = p->automaton;
    const uint32_t P = a.P, W = (P + 31) / 32;
    uint64_t need = which == 2 ? (uint64_t)P * W : which == 3 ? (uint64_t)P * 4 : W;
    if (which < 0 || which > 3) return -1;
    if (!out || cap < need) return -(int64_t)need;
    if (which == 0) for (uint32_t w = 0; w < W; w++) out[w] = a.first.word(w);
    if (which == 1) for (uint32_t w = 0; w < W; w++) out[w] = a.accept.word(w);
    if (which == 2) for (uint32_t k = 0; k < P; k++) for (uint32_t w = 0; w < W; w++) out[(uint64_t)k * W + w] = a.follow[k].word(w);
    if (which == 3) for (uint32_t k = 0; k < P; k++) {
        out[k * 4 + 0] = (uint32_t)a.atoms[k].cls[0]; out[k * 4 + 1] = (uint32_t)(a.atoms[k].cls[0] >> 32);
        out[k * 4 + 2] = (uint32_t)a.atoms[k].cls[1]; out[k * 4 + 3] = (uint32_t)(a.atoms[k].cls[1] >> 32);
    }
    return (int64_t)need;
}
const uint32_t* rrx_program_blob(const rrx_program* p, uint64_t* n_words) {
    if (n_words) *n_words = p->prog.blob.size();
    return p->prog.blob.data();
}
int64_t rrx_program_raw_row(const rrx_program* p, uint32_t state, uint32_t byte, uint32_t* out, uint64_t cap) {
    const Automaton& a = p->automaton;
    const uint32_t W = (a.n_raw + 31) / 32;
    if (!out || cap < W) return -(int64_t)W;
    for (uint32_t w = 0; w < W; w++) out[w] = 0;
    if (byte >= 128) return W;
    for (const Atom& at : a.atoms)
        if (at.pre.has(state) && ((at.cls[byte >> 6] >> (byte & 63)) & 1ull))
            for (uint32_t w = 0; w < W; w++) out[w] |= at.post.word(w);
    return W;
}
int64_t rrx_program_raw_finals(const rrx_program* p, uint32_t* out, uint64_t cap) {
    const Automaton& a = p->automaton;
    const uint32_t W = (a.n_raw + 31) / 32;
    if (!out || cap < W) return -(int64_t)W;
    for (uint32_t w = 0; w < W; w++) out[w] = a.raw_finals.word(w);
    return W;
}

// ---- scanner ---------------------------------------------------------------------------------------
static rrx_scanner* act(rrx_scanner* s) { return (s && s->use_alt && s->alt) ? s->alt : s; }

void rrx_scanner_free(rrx_scanner* s) {
    if (!s) return;
    if (s->alt) { rrx_scanner_free(s->alt); s->alt = nullptr; }
    cudaFree(s->d_prog); cudaFree(s->d_ctrl);          // (d_tpay lives inside d_ctrl's allocation)
    if (s->owns_out) cudaFree(s->d_out);
    cudaFree(s->d_stage); cudaFree(s->d_tile_seg); cudaFree(s->d_tile_cnt); cudaFree(s->d_chunk);
    cudaFree(s->d_shard); cudaFree(s->d_pending);
    if (s->h_ctr) cudaFreeHost(s->h_ctr);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

int rrx_scanner_create(const rrx_program* p, uint64_t max_len, uint64_t capacity, rrx_scanner** out) {
    return rrx_scanner_create_ex(p, max_len, capacity, RRX_SCANNER_DEFAULT, out);
}
int rrx_scanner_is_jit(const rrx_scanner* s) { return s && s->jit ? 1 : 0; }
int rrx_scanner_used_fallback(const rrx_scanner* s) { return s && s->use_alt ? 1 : 0; }
int rrx_scanner_kernel_variant(const rrx_scanner* s) { return s ? s->variant : 0; }
uint64_t rrx_program_jit_source(const rrx_program* p, char* buf, uint64_t cap) {
    const std::string src = jit_applicable(p->prog) ? jit_policy_source(p->prog) : std::string();
    if (buf && cap) { const size_t n = src.size() + 1 < cap ? src.size() + 1 : cap; std::memcpy(buf, src.c_str(), n); buf[n - 1] = 0; }
    return src.size() + 1;
}

int rrx_program_jit_compile(const rrx_program* p, const char* path) {
    if (!p) return set_err(RRX_ERR_ARG, "null argument");
    if (!jit_applicable(p->prog)) return set_err(RRX_ERR_UNSUPPORTED, "automata that need per-source row tables are not specialised");
    std::vector<char> cubin;
    std::string err;
    // (RRX_JIT_VARIANT=1: the straddle kernel of a thread-owned program instead of the lean one, as in rrx_scanner_create_ex)
    const char* v = getenv("RRX_JIT_VARIANT");
    const int variant = (v && v[0] == '1' && p->prog.hdr().donly) ? 1 : 0;
    if (!jit_compile_cubin(p->prog, &cubin, &err, variant)) return set_err(RRX_ERR_CUDA, err);
    if (path) {
        FILE* f = fopen(path, "wb");
        if (!f) return set_err(RRX_ERR_ARG, "cannot open output file");
        fwrite(cubin.data(), 1, cubin.size(), f);
        fclose(f);
    }
    return RRX_OK;
}

int rrx_scanner_create_ex(const rrx_program* p, uint64_t max_len, uint64_t capacity, uint32_t flags, rrx_scanner** out) {
    if (!p || !out) return set_err(RRX_ERR_ARG, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(RRX_ERR_CUDA, "no CUDA device: librrx has no CPU fallback");
    std::unique_ptr<rrx_scanner, void (*)(rrx_scanner*)> s(new rrx_scanner, rrx_scanner_free);
    s->p = p;
    s->create_flags = flags;
    RRX_CUDA(cudaGetDevice(&s->device));
    RRX_CUDA(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device));
    const ProgHeader& h = p->prog.hdr();
    s->NW = h.NW;
    const uint64_t tile_bytes = (uint64_t)TILE_BYTES;
    s->max_len = max_len;
    s->max_tiles = (max_len + tile_bytes - 1) / tile_bytes;
    if (s->max_tiles == 0) s->max_tiles = 1;
    s->capacity = capacity ? capacity : 1;
    RRX_CUDA(cudaMalloc(&s->d_prog, p->prog.blob.size() * 4));
    RRX_CUDA(upload_sync(s->d_prog, p->prog.blob.data(), p->prog.blob.size() * 4));
    s->ctrl_bytes = ctrl_layout(s->max_tiles, nullptr);
    // control block and look-back payloads in ONE allocation: a scan places the payloads right behind the part of the
    // control block it uses, so that one memset zeroes both (payloads with the ready bit inside need it)
    RRX_CUDA(cudaMalloc(&s->d_ctrl, s->ctrl_bytes + 256 + (2 * s->max_tiles + 2 * ((s->max_tiles + 31) / 32)) * (1 + 3 * (size_t)h.NW) * 4));
    RRX_CUDA(cudaMalloc(&s->d_out, s->capacity * 16));
    // staging list: fixed slots for tiles with few matches (at most capacity/2, see rrx_scanner_scan) in
    // front of an overflow area that can hold every match
    RRX_CUDA(cudaMalloc(&s->d_stage, (s->capacity + s->capacity / 2 + 64) * 16));
    RRX_CUDA(cudaMalloc(&s->d_tile_seg, s->max_tiles * 8));
    RRX_CUDA(cudaMalloc(&s->d_tile_cnt, s->max_tiles * 4));
    RRX_CUDA(cudaMalloc(&s->d_chunk, ((s->max_tiles + 1023) / 1024 + 1) * 8));
    RRX_CUDA(cudaMalloc(&s->d_shard, (1 + 3 * (size_t)h.NW) * 4));
    if (h.n_anchor) s->pending_cap = 1u << 16;      // a start-anchored thread that is carried far ends every one of its matches' start walks at the text start
    if (t_pending_cap_want > s->pending_cap) s->pending_cap = t_pending_cap_want;   // (one-shot calls that saw the list overflow)
    RRX_CUDA(cudaMalloc(&s->d_pending, sizeof(PendingStart) * s->pending_cap));
    RRX_CUDA(cudaMallocHost(&s->h_ctr, 32 + 4 * (1 + 3 * (size_t)RRX_NW_MAX)));
    {
        // run-time specialisation: worth its ~1 s compile only for long inputs
        bool want = (flags & RRX_SCANNER_JIT) || (!(flags & RRX_SCANNER_NO_JIT) && max_len >= (32ull << 20));
        if (const char* e = getenv("RRX_JIT")) want = e[0] == '1' && !(flags & RRX_SCANNER_NO_JIT);
        if (want && jit_applicable(p->prog)) {
            std::string err;
            s->jit = jit_get(p->prog, &err, 0);
            s->jit_v[0] = s->jit;
            if (const char* v = getenv("RRX_JIT_VARIANT")) {          // force a kernel variant (A/B measurements)
                if (v[0] == '1' && s->jit && p->prog.hdr().donly) {
                    JitKernel* k1 = jit_get(p->prog, &err, 1);
                    if (k1) { s->jit = s->jit_v[1] = k1; s->variant = 1; }
                }
            }
            if (!s->jit) {
                if (flags & RRX_SCANNER_JIT) return set_err(RRX_ERR_CUDA, "run-time specialisation failed: " + err);
                // not fatal (the pre-compiled kernel gives the same results, slower), but never silent: one line on
                // stderr, and the text stays readable through rrx_last_error() although the call returns RRX_OK
                g_err = "warning: run-time specialisation unavailable (" + err + "); using the pre-compiled kernel";
                fprintf(stderr, "rrx: %s\n", g_err.c_str());
            }
        }
    }
    *out = s.release();
    return RRX_OK;
}

// Run the scan recorded in s->last on the fallback scanner (created on first use).
static int scan_on_alt(rrx_scanner* s) {
    rrx_program* p = const_cast<rrx_program*>(s->p);
    {
        std::lock_guard<std::mutex> lk(p->alt_mu);
        if (!p->alt) {
            int rc = rrx_compile_ex(p->pattern.c_str(), p->flags | RRX_COMPILE_NO_DONLY, &p->alt);
            if (rc) return rc;
        }
    }
    if (!s->alt) {
        int rc = rrx_scanner_create_ex(p->alt, s->max_len, s->capacity, s->create_flags, &s->alt);
        if (rc) return rc;
        if (s->timing) rrx_scanner_enable_timing(s->alt, 1);
    }
    rrx_scan_opts o;
    std::memset(&o, 0, sizeof(o));
    o.left_context = s->last.left_ctx;
    o.entry = s->last.has_entry ? s->last.entry : nullptr;
    return rrx_scanner_scan(s->alt, s->last.text, s->last.len, s->last.base, &o, s->last.stream);
}

int rrx_scanner_scan(rrx_scanner* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset,
                     const rrx_scan_opts* opts, void* stream) {
    if (!s || (!d_text && len)) return set_err(RRX_ERR_ARG, "null argument");
    if (len > s->max_len) return set_err(RRX_ERR_ARG, "text longer than the scanner was created for");
    if (((uintptr_t)d_text) & 15) return set_err(RRX_ERR_ARG, "device text pointer must be 16-byte aligned");
    if (opts && opts->mode != RRX_MODE_LAZY_V1) return set_err(RRX_ERR_UNSUPPORTED, "only RRX_MODE_LAZY_V1 is implemented on the device");
    cudaStream_t st = (cudaStream_t)stream;
    const ProgHeader& h = s->p->prog.hdr();
    if (h.donly) {
        // The thread-owned candidate scan carries dominated states only.  An entry set with other states, or an
        // explicit request (RRX_SCAN_NO_DONLY), goes to the fallback scanner: the same regex, row-bounded scan.
        s->last.text = d_text; s->last.len = len; s->last.base = base_offset; s->last.left_ctx = opts ? opts->left_context : 0;
        s->last.has_entry = opts && opts->entry; s->last.stream = stream;
        bool outside = opts && (opts->flags & RRX_SCAN_NO_DONLY);
        for (uint32_t w = 0; w < h.NW; w++) {
            s->last.entry[w] = s->last.has_entry ? opts->entry[w] : 0u;
            if (s->last.entry[w] & ~s->p->prog.blob[h.off_sparse + w]) outside = true;
        }
        s->use_alt = false;
        if (outside) {
            int rc = scan_on_alt(s);
            if (rc) return rc;
            s->use_alt = true;
            return RRX_OK;
        }
    }
    const uint64_t tile_bytes = (uint64_t)TILE_BYTES;
    ScanParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.text = d_text;
    sp.len = len;
    sp.left_ctx = opts ? opts->left_context : 0;
    sp.base = (int64_t)base_offset;
    sp.prog = s->d_prog;
    sp.num_tiles = (uint32_t)((len + tile_bytes - 1) / tile_bytes);
    sp.static_closed = (h.static_depth != 0xffffffffu && h.static_depth + 1 <= (uint32_t)ROW_BYTES) ? 1u : 0u;
    {
        const char* e = getenv("RRX_STAGGER_NS");
        sp.stagger_ns = e ? (uint32_t)atoi(e) : 1500u;
        const char* sv = getenv("RRX_SPARSE");     // 0 forces the dense row scan (A/B measurements)
        sp.sparse = (h.sparse_ok && !(sv && sv[0] == '0')) ? 1u : 0u;
        sp.donly = h.donly;
        const char* ta = getenv("RRX_TICKET_AHEAD");
        sp.ticket_ahead = ta ? (uint32_t)atoi(ta) : 5u;
    }
#ifdef RRX_PHASE_TIMING
    {
        static unsigned long long* d_ph = nullptr;
        if (!d_ph) { cudaMalloc(&d_ph, 64); cudaMemset(d_ph, 0, 64); }
        sp.phase_cycles = d_ph;
        if (getenv("RRX_PHASE_DUMP")) {
            unsigned long long h[8];
            cudaMemcpy(h, d_ph, 64, cudaMemcpyDeviceToHost);
            fprintf(stderr, "phase cycles: ticket %llu load %llu A+S1 %llu lookback_state %llu S2+publish %llu B %llu lookback_count %llu C %llu\n",
                    h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
            cudaMemset(d_ph, 0, 64);
        }
    }
#endif
    uint8_t* c = s->d_ctrl;
    size_t off[10];
    const size_t used = ctrl_layout(sp.num_tiles, off);
    // look-back payload entries: 1+3*CW words, CW = the kernel's carried width (rrx_prog.h; the pre-compiled
    // kernels carry NW words); CW == 1: one 16-byte word per entry with the ready bit inside
    const size_t cw = (s->jit && h.donly) ? h.CW : h.NW;
    const size_t T = sp.num_tiles, B = (T + 31) / 32, PW = cw == 1 ? 4 : 1 + 3 * cw;
    sp.ticket = reinterpret_cast<uint32_t*>(c + 128);
    sp.err = reinterpret_cast<uint32_t*>(c + 4);
    sp.total = reinterpret_cast<unsigned long long*>(c + 8);
    sp.f_a0 = reinterpret_cast<volatile uint32_t*>(c + off[0]);
    sp.f_i0 = reinterpret_cast<volatile uint32_t*>(c + off[1]);
    sp.f_a1 = reinterpret_cast<volatile uint32_t*>(c + off[2]);
    sp.f_i1 = reinterpret_cast<volatile uint32_t*>(c + off[3]);
    sp.st_arrive = reinterpret_cast<uint32_t*>(c + off[4]);
    // (only kernels that WRITE the stop-byte table may wait on it: thread-owned programs with a reverse skip state)
    sp.straddle_tiles = reinterpret_cast<uint32_t*>(c);          // word 0 of the control block
    s->last_tiles = sp.num_tiles;
    sp.tile_stop = (h.donly && h.rskip_n) ? reinterpret_cast<volatile uint32_t*>(c + off[5]) : nullptr;
    sp.tile_seg = s->d_tile_seg;
    sp.tile_cnt = s->d_tile_cnt;
    s->d_tpay = reinterpret_cast<uint32_t*>(c + ((used + 255) & ~(size_t)255));
    sp.st_a0 = s->d_tpay;
    sp.st_i0 = sp.st_a0 + T * PW;
    sp.st_a1 = sp.st_i0 + T * PW;
    sp.st_i1 = sp.st_a1 + B * PW;
    sp.out = s->d_stage;
    sp.out_cap = s->capacity;
    sp.shard_comp = s->d_shard;
    sp.pending = s->d_pending;
    sp.pending_cap = s->pending_cap;
    {
        // fixed staging slots for tiles with few matches: at most half of the staging list, 8..64 per tile
        uint64_t stride = sp.num_tiles ? s->capacity / (2ull * sp.num_tiles) : 0;
        if (stride > 64) stride = 64;
        if (stride < 8 || getenv("RRX_NO_FIXED_SLOTS")) stride = 0;
        sp.seg_stride = (uint32_t)stride;
        sp.ovf_base = stride * sp.num_tiles;          // <= capacity / 2
        sp.out_cap = sp.ovf_base + s->capacity;       // staging slots
    }
    if (opts && opts->entry) for (uint32_t w = 0; w < h.NW; w++) sp.entry[w] = opts->entry[w];
    s->launches = 0;
    s->last_base = (int64_t)base_offset;
    s->h_shard_valid = false;
    // one memset: the control block, and — one-word composites keep the ready bit inside their 16-byte words — the
    // look-back payloads behind it
    RRX_CUDA(cudaMemsetAsync(s->d_ctrl, 0, cw == 1 ? ((used + 255) & ~(size_t)255) + (2 * T + 2 * B) * PW * 4 : used, st));
    if (sp.num_tiles == 0) {
        // empty text: identity transfer function, no matches
        uint32_t ident[1 + 3 * RRX_NW_MAX] = {K_IDENT};
        RRX_CUDA(cudaMemcpyAsync(s->d_shard, ident, (1 + 3 * (size_t)h.NW) * 4, cudaMemcpyHostToDevice, st));
        return RRX_OK;
    }
    int grid = 0;
    if (s->timing) RRX_CUDA(cudaEventRecord(s->ev0, st));
    cudaError_t e = s->jit ? jit_launch(s->jit, sp, s->sm_count, st, &grid) : launch_scan(sp, h, s->sm_count, st, &grid);
    if (e != cudaSuccess) return cuda_err(e, "rrx_scan_kernel launch");
    if (s->timing) RRX_CUDA(cudaEventRecord(s->ev1, st));
    e = launch_order(s->d_tile_cnt, s->d_tile_seg, sp.num_tiles, s->d_chunk, s->d_stage, sp.out_cap, s->d_out, s->capacity,
                     reinterpret_cast<unsigned long long*>(s->d_ctrl + 24), st);
    if (e != cudaSuccess) return cuda_err(e, "ordering pass launch");
    s->launches = 4;
    return RRX_OK;
}

int rrx_scanner_result(rrx_scanner* s, rrx_scan_result* out, void* stream) {
    if (!s || !out) return set_err(RRX_ERR_ARG, "null argument");
    if (s->use_alt && s->alt) return rrx_scanner_result(s->alt, out, stream);
    cudaStream_t st = (cudaStream_t)stream;
    // counters and the shard composite come back together: one synchronisation per scan
    RRX_CUDA(cudaMemcpyAsync(s->h_ctr, s->d_ctrl, 32, cudaMemcpyDeviceToHost, st));
    RRX_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(s->h_ctr) + 32, s->d_shard, (1 + 3 * (size_t)s->NW) * 4, cudaMemcpyDeviceToHost, st));
    RRX_CUDA(cudaStreamSynchronize(st));
    s->h_shard_valid = true;
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(s->h_ctr);
    if (s->jit && s->p->prog.hdr().donly && s->last_tiles >= 64 && !getenv("RRX_JIT_VARIANT")) {
        // pick the kernel of the next scan by what this one saw (see rrx_scanner::jit_v)
        const uint64_t straddle = w32[0], tiles = s->last_tiles;
        const int want = (s->variant == 0 && straddle * 8 > tiles) ? 1 : (s->variant == 1 && straddle * 64 < tiles) ? 0 : s->variant;
        if (want != s->variant) {
            if (!s->jit_v[want]) {
                std::string err;
                s->jit_v[want] = jit_get(s->p->prog, &err, want);
            }
            if (s->jit_v[want]) { s->jit = s->jit_v[want]; s->variant = want; }
        }
    }
    out->error_bits = w32[1] & ~(uint32_t)ERR_VERIFY_EDGE;
    out->soft_bits = w32[1] & (uint32_t)ERR_VERIFY_EDGE;
    out->reserved = 0;
    if ((w32[1] & ERR_VERIFY_BUDGET) && s->p->prog.hdr().donly) {
        // a verification walk of the thread-owned scan ran out of budget: the list is void.  Scan again with the
        // row-bounded candidate scan (same regex, same arguments) and report that result.
        int rc = scan_on_alt(s);
        if (rc) return rc;
        s->use_alt = true;
        return rrx_scanner_result(s->alt, out, stream);
    }
    out->n_matches = s->h_ctr[3];            // written by the ordering pass (sum of the tile counts)
    if (out->error_bits & ERR_STAGING_OVERFLOW) {
        // the overflow area of the staging list ran out: report it as "more matches than capacity" so that
        // callers take their resize-and-rescan path
        out->error_bits &= ~(uint32_t)ERR_STAGING_OVERFLOW;
        if (out->n_matches < 2 * s->capacity) out->n_matches = 2 * s->capacity;
    }
    out->n_pending = s->h_ctr[2];
    out->capacity = s->capacity;
    out->launches = s->launches;
    out->d_matches = reinterpret_cast<const int64_t*>(s->d_out);
    return RRX_OK;
}

int rrx_scanner_copy_matches(rrx_scanner* s, int64_t* host_out, uint64_t n, void* stream) {
    if (!s || (!host_out && n)) return set_err(RRX_ERR_ARG, "null argument");
    s = act(s);
    if (n > s->capacity) return set_err(RRX_ERR_ARG, "more matches requested than the scanner holds");
    cudaStream_t st = (cudaStream_t)stream;
    if (n) RRX_CUDA(cudaMemcpyAsync(host_out, s->d_out, n * 16, cudaMemcpyDeviceToHost, st));
    RRX_CUDA(cudaStreamSynchronize(st));
    return RRX_OK;
}

int rrx_fold_entry(const rrx_program* p, const uint32_t* summaries, uint32_t n_shards, uint32_t rank, uint32_t* entry_out, int* exact) {
    if (!p || !summaries || !entry_out || !exact) return set_err(RRX_ERR_ARG, "null argument");
    fold_entry_host(p->prog.hdr().NW, summaries, n_shards, rank, entry_out, exact);
    return RRX_OK;
}
int rrx_fold_entries(const rrx_program* p, const uint32_t* summaries, uint32_t n_shards, uint32_t* entries_out, int* exact_out) {
    if (!p || !summaries || !entries_out || !exact_out) return set_err(RRX_ERR_ARG, "null argument");
    const uint32_t W = p->prog.hdr().NW;
    for (uint32_t r = 0; r < n_shards; r++) fold_entry_host(W, summaries, n_shards, r, entries_out + (size_t)r * W, exact_out + r);
    return RRX_OK;
}
int rrx_apply_summary(const rrx_program* p, const uint32_t* summary, const uint32_t* entry, uint32_t* exit_out) {
    if (!p || !summary || !entry || !exit_out) return set_err(RRX_ERR_ARG, "null argument");
    apply_summary_host(p->prog.hdr().NW, summary, entry, exit_out);
    return RRX_OK;
}
int rrx_scanner_copy_pending(rrx_scanner* s, void* host_out, uint64_t cap, uint64_t* n_out, void* stream) {
    if (!s || !n_out) return set_err(RRX_ERR_ARG, "null argument");
    s = act(s);
    cudaStream_t st = (cudaStream_t)stream;
    RRX_CUDA(cudaMemcpyAsync(s->h_ctr, s->d_ctrl, 32, cudaMemcpyDeviceToHost, st));
    RRX_CUDA(cudaStreamSynchronize(st));
    uint64_t n = s->h_ctr[2];
    if (n > s->pending_cap) n = s->pending_cap;
    *n_out = n;
    if (n > cap) n = cap;
    if (n && host_out) {
        RRX_CUDA(cudaMemcpy(host_out, s->d_pending, n * sizeof(PendingStart), cudaMemcpyDeviceToHost));
        // (tile, rank in tile) -> index in the ordered match list: the ordering pass left every tile's
        // offset in d_tile_seg
        PendingStart* recs = reinterpret_cast<PendingStart*>(host_out);
        uint64_t t_lo = ~0ull, t_hi = 0;
        for (uint64_t i = 0; i < n; i++) { const uint64_t t = recs[i].match_index >> 32; t_lo = t < t_lo ? t : t_lo; t_hi = t > t_hi ? t : t_hi; }
        std::vector<unsigned long long> segs(t_hi - t_lo + 1);          // the tiles' offsets in the ordered list, one copy
        RRX_CUDA(cudaMemcpy(segs.data(), s->d_tile_seg + t_lo, segs.size() * 8, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t tile = recs[i].match_index >> 32, rank = recs[i].match_index & 0xffffffffull;
            const unsigned long long off = segs[tile - t_lo];
            recs[i].match_index = off + rank;
            recs[i].next_offset += s->last_base;          // shard-local -> global
        }
    }
    return RRX_OK;
}
int rrx_continue_starts(const rrx_program* p, const uint8_t* d_text, uint64_t len, uint64_t left_context, uint64_t base_offset,
                        void* records_host, uint64_t n, int64_t* start_out, void* stream) {
    if (!p || (!records_host && n) || (!start_out && n)) return set_err(RRX_ERR_ARG, "null argument");
    if (!n) return RRX_OK;
    {
        // every record must point into this shard's readable text (global offsets)
        const PendingStart* recs = reinterpret_cast<const PendingStart*>(records_host);
        const int64_t lo = (int64_t)base_offset - (int64_t)left_context, hi = (int64_t)base_offset + (int64_t)len;
        for (uint64_t i = 0; i < n; i++)
            if (recs[i].next_offset < lo - 1 || recs[i].next_offset >= hi)
                return set_err(RRX_ERR_ARG, "rrx_continue_starts: record offset outside this shard's text");
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* d_prog = nullptr;
    PendingStart* d_recs = nullptr;
    long long* d_out = nullptr;
    const size_t pb = p->prog.blob.size() * 4;
    RRX_CUDA(cudaMalloc(&d_prog, pb));
    RRX_CUDA(cudaMalloc(&d_recs, n * sizeof(PendingStart)));
    RRX_CUDA(cudaMalloc(&d_out, n * 8));
    cudaMemcpyAsync(d_prog, p->prog.blob.data(), pb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_recs, records_host, n * sizeof(PendingStart), cudaMemcpyHostToDevice, st);
    cudaError_t e = launch_continue_starts(p->prog.hdr().NW, d_prog, d_text, len, left_context, base_offset, d_recs, n, d_out, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(start_out, d_out, n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(records_host, d_recs, n * sizeof(PendingStart), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_prog); cudaFree(d_recs); cudaFree(d_out);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "rrx_continue_starts");
}

int rrx_scanner_shard_summary(rrx_scanner* s, uint32_t* out, uint64_t cap_words, void* stream) {
    s = act(s);
    const size_t n = 1 + 3 * (size_t)s->NW;
    if (!out || cap_words < n) return set_err(RRX_ERR_ARG, "summary buffer too small");
    if (s->h_shard_valid) {          // already fetched with the counters (rrx_scanner_result)
        std::memcpy(out, reinterpret_cast<const uint8_t*>(s->h_ctr) + 32, n * 4);
        return RRX_OK;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RRX_CUDA(cudaMemcpyAsync(out, s->d_shard, n * 4, cudaMemcpyDeviceToHost, st));
    RRX_CUDA(cudaStreamSynchronize(st));
    return RRX_OK;
}

int rrx_scanner_enable_timing(rrx_scanner* s, int on) {
    if (!s) return set_err(RRX_ERR_ARG, "null argument");
    if (on && !s->ev0) {
        RRX_CUDA(cudaEventCreate(&s->ev0));
        RRX_CUDA(cudaEventCreate(&s->ev1));
    }
    s->timing = on != 0;
    if (s->alt) rrx_scanner_enable_timing(s->alt, on);
    return RRX_OK;
}
int rrx_scanner_kernel_ms(rrx_scanner* s, float* ms) {
    s = act(s);
    if (!s || !ms || !s->ev0) return set_err(RRX_ERR_ARG, "timing was not enabled");
    RRX_CUDA(cudaEventSynchronize(s->ev1));
    RRX_CUDA(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return RRX_OK;
}

int rrx_scanner_message_device(rrx_scanner* s, uint32_t* d_out, uint64_t cap_words, void* stream) {
    if (!s || !d_out) return set_err(RRX_ERR_ARG, "null argument");
    s = act(s);
    const size_t n = 1 + 3 * (size_t)s->NW;
    if (cap_words < n + 8) return set_err(RRX_ERR_ARG, "message buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    RRX_CUDA(cudaMemcpyAsync(d_out, s->d_shard, n * 4, cudaMemcpyDeviceToDevice, st));
    RRX_CUDA(cudaMemcpyAsync(d_out + n, s->d_ctrl, 32, cudaMemcpyDeviceToDevice, st));
    return RRX_OK;
}

// device pointers of the last scan's shard composite (1+3*NW words) and counter block (8 words): for rrx_dist.cu, which
// packs its message with one kernel instead of copies
void rrx_scanner_message_sources(rrx_scanner* s, const uint32_t** d_comp, const uint32_t** d_ctr) {
    s = act(s);
    *d_comp = s->d_shard;
    *d_ctr = reinterpret_cast<const uint32_t*>(s->d_ctrl);
}
// where the last scan's ordered list lives and how many pairs fit — no synchronisation (rrx_dist.cu has the counts
// from its own status block)
void rrx_scanner_peek(rrx_scanner* s, const int64_t** d_matches, uint64_t* capacity) {
    s = act(s);
    *d_matches = reinterpret_cast<const int64_t*>(s->d_out);
    *capacity = s->capacity;
}

// (anchors: defined further down)
static bool anchor_entry(const rrx_program* p, uint32_t* entry);
static int anchor_fixups(const rrx_program* p, rrx_scanner* s, const uint8_t* d_text, uint64_t len, uint64_t base, const uint32_t* entry,
                         const rrx_scan_result& r, void* stream, int* extra, int64_t* extra_pair);
static int anchor_merge_extra(long long* d_pairs, uint64_t* count, const int64_t* ep);

// ---- one-shot convenience calls ----------------------------------------------------------------------
int rrx_scan_device(const rrx_program* p, const uint8_t* d_text, uint64_t len, uint64_t base_offset,
                    int mode, rrx_matches** out, void* stream) {
    if (!p || !out) return set_err(RRX_ERR_ARG, "null argument");
    if (mode != RRX_MODE_LAZY_V1 && mode != RRX_MODE_NO_OVERLAP) return set_err(RRX_ERR_ARG, "unknown mode");
    uint64_t cap = len / 64 + 1024;
    uint32_t aentry[RRX_NW_MAX];
    const bool anchored = p->prog.hdr().n_anchor != 0;
    const bool a_entry = anchor_entry(p, aentry);
    uint32_t pend_want = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        PendingCapScope pcs(pend_want);
        rrx_scanner* s = nullptr;
        int rc = rrx_scanner_create(p, len, cap + 1, &s);          // (+1: room for the match of a trailing anchor)
        if (rc) return rc;
        rrx_scan_opts o;
        std::memset(&o, 0, sizeof(o));
        if (a_entry) o.entry = aentry;
        rc = rrx_scanner_scan(s, d_text, len, base_offset, &o, stream);
        rrx_scan_result r;
        if (!rc) rc = rrx_scanner_result(s, &r, stream);
        if (rc) { rrx_scanner_free(s); return rc; }
        if (anchored && r.error_bits == ERR_PENDING_OVERFLOW && !pend_want && r.n_pending < PENDING_CAP_MAX) {
            pend_want = (uint32_t)r.n_pending + 1;      // the kernel counted them all: rescan with room for every one
            if (r.n_matches > cap) cap = r.n_matches;
            rrx_scanner_free(s);
            continue;
        }
        if (r.error_bits) {
            rrx_scanner_free(s);
            return set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(r.error_bits));
        }
        if (r.n_matches <= cap) {
            int extra = 0;
            int64_t ep[2];
            if (anchored) {
                rc = anchor_fixups(p, s, d_text, len, base_offset, aentry, r, stream, &extra, ep);
                if (rc) { rrx_scanner_free(s); return rc; }
            }
            rrx_matches* m = new rrx_matches;
            m->count = r.n_matches;
            rrx_scanner* owner = act(s);              // the list may live in the fallback scanner
            m->d_pairs = owner->d_out;
            owner->owns_out = false;
            rrx_scanner_free(s);
            if (extra) {
                rc = anchor_merge_extra(m->d_pairs, &m->count, ep);
                if (rc) { rrx_matches_free(m); return rc; }
            }
            if (mode == RRX_MODE_NO_OVERLAP) {
                long long* filtered = nullptr;
                uint64_t nf = 0;
                rc = filter_no_overlap(m->d_pairs, m->count, &filtered, &nf, (cudaStream_t)stream);
                if (rc) { rrx_matches_free(m); return rc; }
                cudaFree(m->d_pairs);
                m->d_pairs = filtered;
                m->count = nf;
            }
            *out = m;
            return RRX_OK;
        }
        cap = r.n_matches;          // exact size known now: scan again
        rrx_scanner_free(s);
    }
    return set_err(RRX_ERR_CUDA, "match buffer overflow after resize");
}

// ---- anchors ---------------------------------------------------------------------------------
// `^` / `$` are positions whose class is exactly {0x00} (the reference's construction, Parser.cpp:142-146).  The one-shot
// calls frame the text with one virtual 0x00 on each side that only those positions consume:
//   leading byte  = an entry set for the scan: first ∩ anch ("the state after the virtual byte");
//                   a lazy start walk that reaches the text start and could consume the virtual byte there started at
//                   offset 0 (pending records are resolved on the host);
//   trailing byte = one virtual step from the text's exit state; if it accepts, one more match ends at `len`, its lazy
//                   start found by the ordinary backward walk (rrx_continue_starts) from the byte before.
// Matches are reported in text coordinates (the virtual bytes are clipped); empty matches are not reported.
namespace {
struct AnchorView {
    const ProgHeader* h;
    const uint32_t* b;
    uint32_t NW;
    bool meets(const uint32_t* x, uint32_t off) const { uint32_t o = 0; for (uint32_t w = 0; w < NW; w++) o |= x[w] & b[off + w]; return o != 0; }
    void rows(const uint32_t* S, uint32_t off_rows, uint32_t* out) const {            // OR of the table rows of the set bits of S
        for (uint32_t w = 0; w < NW; w++) out[w] = 0;
        for (uint32_t k = 0; k < h->P; k++)
            if ((S[k >> 5] >> (k & 31)) & 1u) for (uint32_t w = 0; w < NW; w++) out[w] |= b[off_rows + k * NW + w];
    }
};
}  // namespace

// entry set that stands for the leading virtual byte (all zero if the program has no start anchor)
static bool anchor_entry(const rrx_program* p, uint32_t* entry) {
    const ProgHeader& h = p->prog.hdr();
    const uint32_t* b = p->prog.blob.data();
    bool any = false;
    for (uint32_t w = 0; w < RRX_NW_MAX; w++) entry[w] = 0;
    if (!h.n_anchor) return false;
    for (uint32_t w = 0; w < h.NW; w++) { entry[w] = b[h.off_first + w] & b[h.off_anch + w]; any = any || entry[w]; }
    return any;
}

// the matches whose start walk ended at the leading virtual byte start at the text start
__global__ void set_starts_kernel(long long* pairs, const unsigned long long* idx, uint64_t n, long long start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pairs[2 * idx[i]] = start;
}

// After the scan of a whole text [0,len) with anchor_entry() as entry: resolve walks that reached the text start,
// take the trailing virtual step.  *extra = 1 and extra_pair = (start,end) if one more match ends at base+len.
static int anchor_fixups(const rrx_program* p, rrx_scanner* s, const uint8_t* d_text, uint64_t len, uint64_t base, const uint32_t* entry,
                         const rrx_scan_result& r, void* stream, int* extra, int64_t* extra_pair) {
    const ProgHeader& h = p->prog.hdr();
    AnchorView av{&h, p->prog.blob.data(), h.NW};
    *extra = 0;
    uint32_t fa[RRX_NW_MAX];                            // first ∩ anch
    for (uint32_t w = 0; w < h.NW; w++) fa[w] = av.b[h.off_first + w] & av.b[h.off_anch + w];
    rrx_scanner* sa = act(s);
    // (a) start walks that ran off the text start
    if (r.n_pending) {
        std::vector<PendingStart> recs(r.n_pending);
        uint64_t n = 0;
        int rc = rrx_scanner_copy_pending(s, recs.data(), recs.size(), &n, stream);
        if (rc) return rc;
        if (n != r.n_pending) return set_err(RRX_ERR_CUDA, "too many start walks reach the text start");
        std::vector<unsigned long long> idx(n);
        for (uint64_t i = 0; i < n; i++) {
            uint32_t o = 0;
            for (uint32_t w = 0; w < h.NW; w++) o |= recs[i].R[w] & fa[w];
            if (!o) return set_err(RRX_ERR_CUDA, "a start walk left the text without a start anchor to end at");
            idx[i] = recs[i].match_index;
        }
        unsigned long long* d_idx = nullptr;
        RRX_CUDA(cudaMalloc(&d_idx, n * 8));
        cudaError_t e = cudaMemcpyAsync(d_idx, idx.data(), n * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e == cudaSuccess) {
            set_starts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sa->d_out, d_idx, n, (long long)base);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(d_idx);
        if (e != cudaSuccess) return cuda_err(e, "anchored starts");
    }
    if (len == 0) return RRX_OK;
    // (b) the trailing virtual byte
    std::vector<uint32_t> summary(1 + 3 * (size_t)h.NW);
    int rc = rrx_scanner_shard_summary(s, summary.data(), summary.size(), stream);
    if (rc) return rc;
    uint32_t S[RRX_NW_MAX], f[RRX_NW_MAX], V[RRX_NW_MAX], x[RRX_NW_MAX], R[RRX_NW_MAX];
    apply_summary_host(h.NW, summary.data(), entry, S);
    // the scan's own exit set may lack states outside the dominated set (thread-owned programs): those cannot reach an
    // accepting anchor position in one step (its predecessors are dominated), so they do not matter here
    av.rows(S, h.off_follow, f);
    bool acc = false, any_x = false, only_virtual = false;
    for (uint32_t w = 0; w < h.NW; w++) {
        V[w] = (f[w] | av.b[h.off_first + w]) & av.b[h.off_anch + w];
        if (V[w] & av.b[h.off_accept + w]) acc = true;
        x[w] = av.b[h.off_accept + w] & av.b[h.off_anch + w];
        any_x = any_x || x[w];
        if (x[w] & av.b[h.off_first + w]) only_virtual = true;      // the lazy (shortest) match is the virtual byte alone: empty
    }
    if (!acc || !any_x || only_virtual) return RRX_OK;
    av.rows(x, h.off_rfollow, R);
    PendingStart rec;
    std::memset(&rec, 0, sizeof(rec));
    rec.next_offset = (int64_t)(base + len) - 1;
    for (uint32_t w = 0; w < h.NW; w++) rec.R[w] = R[w];
    int64_t start = -1;
    rc = rrx_continue_starts(p, d_text, len, 0, base, &rec, 1, &start, stream);
    if (rc) return rc;
    if (start == -1) {                                   // reached the text start: the leading virtual byte
        uint32_t o = 0;
        for (uint32_t w = 0; w < h.NW; w++) o |= rec.R[w] & fa[w];
        if (!o) return RRX_OK;
        start = (int64_t)base;
    }
    if (start < 0 || start >= (int64_t)(base + len)) return RRX_OK;
    *extra = 1;
    extra_pair[0] = start;
    extra_pair[1] = (int64_t)(base + len);
    return RRX_OK;
}

// merge the match of the trailing virtual byte into the ordered device list (room for one more pair is guaranteed by the
// callers): LAZY-v1 keeps one match per end, the one with the larger start
static int anchor_merge_extra(long long* d_pairs, uint64_t* count, const int64_t* ep) {
    if (*count) {
        long long last[2];
        RRX_CUDA(cudaMemcpy(last, d_pairs + 2 * (*count - 1), 16, cudaMemcpyDeviceToHost));
        if (last[1] == ep[1]) {
            if (ep[0] > last[0]) RRX_CUDA(upload_sync(d_pairs + 2 * (*count - 1), ep, 8));
            return RRX_OK;
        }
    }
    RRX_CUDA(upload_sync(d_pairs + 2 * *count, ep, 16));
    (*count)++;
    return RRX_OK;
}

// ---- streaming host scan ------------------------------------------------------------------
// The text is copied to the device in chunks on hc->st_copy (all copies enqueued up front for page-locked input; for
// pageable input a producer thread moves each chunk through page-locked bounce buffers with a few memcpy threads, which
// is what takes the copy from the driver's ~11 GB/s staging to the host's memcpy rate).  As soon as a chunk's event fires
// it is scanned on hc->st with everything before it as left context and the previous chunk's exit state as entry; its
// matches go to the page-locked result buffer while the next chunk is still in flight.  The list is identical to a
// one-piece scan (chunk boundaries are shard boundaries with full left context).
static int scan_host_streaming(const rrx_program* p, HostCtx* hc, const uint8_t* text, uint64_t len, uint64_t base_offset,
                               uint64_t chunk, rrx_matches** out) {
    const uint64_t nchunks = (len + chunk - 1) / chunk;
    if (!hc->st_copy) RRX_CUDA(cudaStreamCreateWithFlags(&hc->st_copy, cudaStreamNonBlocking));
    while (hc->ev_chunk.size() < nchunks) {
        cudaEvent_t e;
        RRX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hc->ev_chunk.push_back(e);
    }
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, text) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    std::atomic<uint64_t> produced(0);
    std::atomic<int> prod_err(0);
    std::thread producer;
    if (!pageable) {
        for (uint64_t k = 0; k < nchunks; k++) {
            const uint64_t o = k * chunk, n = std::min(chunk, len - o);
            RRX_CUDA(cudaMemcpyAsync(hc->d_text + o, text + o, n, cudaMemcpyHostToDevice, hc->st_copy));
            RRX_CUDA(cudaEventRecord(hc->ev_chunk[k], hc->st_copy));
        }
        produced = nchunks;
    } else {
        const uint64_t piece = 16ull << 20;                  // bounce granularity
        if (hc->bounce_bytes < piece) {
            for (int i = 0; i < 3; i++) {
                if (hc->bounce[i]) cudaFreeHost(hc->bounce[i]);
                RRX_CUDA(cudaMallocHost(&hc->bounce[i], piece));
                if (!hc->ev_bounce[i]) RRX_CUDA(cudaEventCreateWithFlags(&hc->ev_bounce[i], cudaEventDisableTiming));
            }
            hc->bounce_bytes = piece;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        producer = std::thread([=, &produced, &prod_err] {
            cudaSetDevice(dev);
            const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
            uint64_t turn = 0;
            for (uint64_t k = 0; k < nchunks && !prod_err; k++) {
                const uint64_t o0 = k * chunk, n0 = std::min(chunk, len - o0);
                for (uint64_t q = 0; q < n0; q += piece, turn++) {
                    const int b = (int)(turn % 3);
                    const uint64_t n = std::min(piece, n0 - q);
                    if (turn >= 3 && cudaEventSynchronize(hc->ev_bounce[b]) != cudaSuccess) { prod_err = 1; break; }
                    std::vector<std::thread> th;
                    const uint64_t per = ((n + nt - 1) / nt + 4095) & ~4095ull;
                    for (unsigned t = 0; t < nt; t++) {
                        const uint64_t lo = t * per, hi = std::min(n, lo + per);
                        if (lo < hi) th.emplace_back([=] { std::memcpy(hc->bounce[b] + lo, text + o0 + q + lo, hi - lo); });
                    }
                    for (auto& x : th) x.join();
                    if (cudaMemcpyAsync(hc->d_text + o0 + q, hc->bounce[b], n, cudaMemcpyHostToDevice, hc->st_copy) != cudaSuccess ||
                        cudaEventRecord(hc->ev_bounce[b], hc->st_copy) != cudaSuccess) { prod_err = 1; break; }
                }
                if (cudaEventRecord(hc->ev_chunk[k], hc->st_copy) != cudaSuccess) prod_err = 1;
                produced = k + 1;
            }
        });
    }
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{producer};

    const ProgHeader& h = p->prog.hdr();
    uint64_t want_cap = chunk / 64 + 1024;
    std::unique_ptr<rrx_matches> m(new rrx_matches);
    m->owner = p->host;
    uint64_t total = 0;
    auto grow_result = [&](uint64_t need) -> int {         // page-locked result buffer with room for `need` pairs
        if (m->pinned.p && m->pinned.pairs >= need) return RRX_OK;
        HostCtx::Pinned nb = {nullptr, 0};
        size_t best = hc->pool.size();
        for (size_t i = 0; i < hc->pool.size(); i++)
            if (hc->pool[i].pairs >= need && (best == hc->pool.size() || hc->pool[i].pairs < hc->pool[best].pairs)) best = i;
        if (best != hc->pool.size()) { nb = hc->pool[best]; hc->pool.erase(hc->pool.begin() + (long)best); }
        else {
            const uint64_t pairs = need + need / 4 + 1024;
            RRX_CUDA(cudaMallocHost(&nb.p, pairs * 16));
            nb.pairs = pairs;
        }
        if (m->pinned.p) {
            RRX_CUDA(cudaStreamSynchronize(hc->st));          // copies of earlier chunks' matches may still be in flight
            std::memcpy(nb.p, m->pinned.p, total * 16);
            if (hc->pool.size() < 4) hc->pool.push_back(m->pinned); else cudaFreeHost(m->pinned.p);
        }
        m->pinned = nb;
        return RRX_OK;
    };
    int rc = grow_result(len / 200 + 4096);
    if (rc) return rc;
    uint32_t entry[RRX_NW_MAX];
    for (uint32_t w = 0; w < RRX_NW_MAX; w++) entry[w] = 0u;
    std::vector<uint32_t> summary(1 + 3 * (size_t)h.NW);
    // A thread-owned program whose scan of some chunk fell back to the row-bounded scan (verification budget) starts over
    // with the row-bounded scan for EVERY chunk: the two scans carry different things from chunk to chunk.
    bool no_donly = false;
    for (uint64_t k = 0; k < nchunks; k++) {
        const uint64_t o = k * chunk, n = std::min(chunk, len - o);
        while (produced.load() <= k && !prod_err.load()) std::this_thread::yield();
        if (prod_err.load()) return set_err(RRX_ERR_CUDA, "streaming host scan: host-to-device copy failed");
        RRX_CUDA(cudaStreamWaitEvent(hc->st, hc->ev_chunk[k], 0));
        for (int attempt = 0;; attempt++) {
            if (!hc->sc_chunk || hc->scc_len < chunk || hc->scc_cap < want_cap) {
                if (hc->sc_chunk) { rrx_scanner_free(hc->sc_chunk); hc->sc_chunk = nullptr; }
                rc = rrx_scanner_create(p, chunk, want_cap, &hc->sc_chunk);
                if (rc) return rc;
                hc->scc_len = chunk; hc->scc_cap = want_cap;
            }
            rrx_scan_opts so;
            std::memset(&so, 0, sizeof(so));
            so.left_context = o;                       // everything before the chunk is resident
            so.entry = entry;
            so.flags = no_donly ? RRX_SCAN_NO_DONLY : 0;
            rc = rrx_scanner_scan(hc->sc_chunk, hc->d_text + o, n, base_offset + o, &so, hc->st);
            rrx_scan_result r;
            if (!rc) rc = rrx_scanner_result(hc->sc_chunk, &r, hc->st);
            if (rc) return rc;
            if (r.error_bits) return set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(r.error_bits));
            if (r.n_matches > r.capacity) {
                if (attempt) return set_err(RRX_ERR_CUDA, "match buffer overflow after resize");
                want_cap = r.n_matches + 1024;
                continue;
            }
            if (!no_donly && rrx_scanner_used_fallback(hc->sc_chunk)) {
                no_donly = true;
                for (uint32_t w = 0; w < RRX_NW_MAX; w++) entry[w] = 0u;
                total = 0;
                k = (uint64_t)-1;                      // (the loop increment makes it 0)
                break;
            }
            if (r.n_pending) return set_err(RRX_ERR_CUDA, "streaming host scan: a start walk left the resident text");
            rc = grow_result(total + r.n_matches);
            if (rc) return rc;
            if (r.n_matches) RRX_CUDA(cudaMemcpyAsync(m->pinned.p + 2 * total, r.d_matches, r.n_matches * 16, cudaMemcpyDeviceToHost, hc->st));
            total += r.n_matches;
            // exit state of this chunk = entry of the next
            rc = rrx_scanner_shard_summary(hc->sc_chunk, summary.data(), summary.size(), hc->st);
            if (rc) return rc;
            uint32_t nxt[RRX_NW_MAX];
            apply_summary_host(h.NW, summary.data(), entry, nxt);
            for (uint32_t w = 0; w < h.NW; w++) entry[w] = nxt[w];
            break;
        }
    }
    RRX_CUDA(cudaStreamSynchronize(hc->st));
    m->count = total;
    m->host_valid = true;
    *out = m.release();
    return RRX_OK;
}

int rrx_scan_host(const rrx_program* p, const uint8_t* text, uint64_t len, uint64_t base_offset,
                  int mode, rrx_matches** out) {
    if (!p || !out || (!text && len)) return set_err(RRX_ERR_ARG, "null argument");
    if (mode != RRX_MODE_LAZY_V1 && mode != RRX_MODE_NO_OVERLAP) return set_err(RRX_ERR_ARG, "unknown mode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(RRX_ERR_CUDA, "no CUDA device: librrx has no CPU fallback");
    std::shared_ptr<HostCtx> hc = p->host;
    std::lock_guard<std::mutex> lk(hc->mu);
    int dev = 0;
    RRX_CUDA(cudaGetDevice(&dev));
    if (hc->device != dev) {            // first use, or the caller switched devices: start over
        if (hc->sc) { rrx_scanner_free(hc->sc); hc->sc = nullptr; }
        cudaFree(hc->d_text); hc->d_text = nullptr; hc->text_cap = 0;
        if (hc->st) { cudaStreamDestroy(hc->st); hc->st = nullptr; }
        RRX_CUDA(cudaStreamCreateWithFlags(&hc->st, cudaStreamNonBlocking));
        hc->device = dev;
    }
    if (len + 16 > hc->text_cap) {
        if (hc->sc) { rrx_scanner_free(hc->sc); hc->sc = nullptr; }
        cudaFree(hc->d_text); hc->d_text = nullptr; hc->text_cap = 0;
        const uint64_t cap = ((len + 16 + (1ull << 20) - 1) >> 20) << 20;
        RRX_CUDA(cudaMalloc(&hc->d_text, cap));
        hc->text_cap = cap;
    }
    // Long texts stream: the copy proceeds chunk by chunk on a stream of its own while the chunks that have arrived are
    // scanned (state carried from chunk to chunk) and their matches copied back — see scan_host_streaming.
    {
        const char* ce = getenv("RRX_HOST_CHUNK");
        const uint64_t chunk = ce ? (strtoull(ce, nullptr, 10) & ~4095ull) : (64ull << 20);
        if (mode == RRX_MODE_LAZY_V1 && chunk >= (1ull << 20) && len >= 3 * chunk && p->prog.hdr().n_anchor == 0)
            return scan_host_streaming(p, hc.get(), text, len, base_offset, chunk, out);
    }
    // the copy is enqueued first, so that whatever set-up follows overlaps it
    if (len) RRX_CUDA(cudaMemcpyAsync(hc->d_text, text, len, cudaMemcpyHostToDevice, hc->st));
    uint64_t want_cap = len / 64 + 1024;
    uint32_t aentry[RRX_NW_MAX];
    const bool anchored = p->prog.hdr().n_anchor != 0;
    const bool a_entry = anchor_entry(p, aentry);
    uint32_t pend_want = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        PendingCapScope pcs(pend_want);
        if (!hc->sc || hc->sc_len < len || hc->sc_cap < want_cap) {
            if (hc->sc) { rrx_scanner_free(hc->sc); hc->sc = nullptr; }
            const uint64_t sl = hc->text_cap - 16;
            int rc = rrx_scanner_create(p, sl, want_cap, &hc->sc);
            if (rc) return rc;
            hc->sc_len = sl; hc->sc_cap = want_cap;
        }
        rrx_scan_opts o;
        std::memset(&o, 0, sizeof(o));
        if (a_entry) o.entry = aentry;
        int rc = rrx_scanner_scan(hc->sc, hc->d_text, len, base_offset, &o, hc->st);
        rrx_scan_result r;
        if (!rc) rc = rrx_scanner_result(hc->sc, &r, hc->st);
        if (rc) return rc;
        if (anchored && r.error_bits == ERR_PENDING_OVERFLOW && !pend_want && r.n_pending < PENDING_CAP_MAX) {
            pend_want = (uint32_t)r.n_pending + 1;      // the kernel counted them all: new scanner with room for every one
            if (r.n_matches + 1 > want_cap) want_cap = r.n_matches + 1024;
            rrx_scanner_free(hc->sc);
            hc->sc = nullptr;
            continue;
        }
        if (r.error_bits) return set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(r.error_bits));
        if (r.n_matches + 1 > r.capacity) { want_cap = r.n_matches + 1024; continue; }    // exact size known now: scan again
        std::unique_ptr<rrx_matches> m(new rrx_matches);
        m->count = r.n_matches;
        m->owner = hc;
        if (anchored) {
            int extra = 0;
            int64_t ep[2];
            rc = anchor_fixups(p, hc->sc, hc->d_text, len, base_offset, aentry, r, hc->st, &extra, ep);
            if (rc) return rc;
            if (extra) {
                rc = anchor_merge_extra(act(hc->sc)->d_out, &m->count, ep);
                if (rc) return rc;
            }
        }
        long long* filtered = nullptr;
        if (mode == RRX_MODE_NO_OVERLAP) {
            uint64_t nf = 0;
            rc = filter_no_overlap(reinterpret_cast<const long long*>(r.d_matches), m->count, &filtered, &nf, hc->st);
            if (rc) return rc;
            m->count = nf;
        }
        std::unique_ptr<long long, void (*)(long long*)> filtered_guard(filtered, [](long long* q) { cudaFree(q); });
        const long long* d_src = filtered ? filtered : reinterpret_cast<const long long*>(r.d_matches);
        if (m->count) {
            size_t best = hc->pool.size();
            for (size_t i = 0; i < hc->pool.size(); i++)
                if (hc->pool[i].pairs >= m->count && (best == hc->pool.size() || hc->pool[i].pairs < hc->pool[best].pairs)) best = i;
            if (best != hc->pool.size()) {
                m->pinned = hc->pool[best];
                hc->pool.erase(hc->pool.begin() + (long)best);
            } else {
                const uint64_t pairs = m->count + m->count / 8 + 1024;
                RRX_CUDA(cudaMallocHost(&m->pinned.p, pairs * 16));
                m->pinned.pairs = pairs;
            }
            cudaError_t e = cudaMemcpyAsync(m->pinned.p, d_src, m->count * 16, cudaMemcpyDeviceToHost, hc->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(hc->st);
            if (e != cudaSuccess) { hc->pool.push_back(m->pinned); return cuda_err(e, "cudaMemcpy D2H"); }
        }
        m->host_valid = true;
        *out = m.release();
        return RRX_OK;
    }
    return set_err(RRX_ERR_CUDA, "match buffer overflow after resize");
}

int rrx_accepts_host(const rrx_program* p, const uint8_t* text, uint64_t len, int* accepted) {
    if (!p || !accepted || (!text && len)) return set_err(RRX_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(RRX_ERR_CUDA, "no CUDA device: librrx has no CPU fallback");
    uint8_t* d = nullptr;
    uint32_t* d_prog = nullptr;
    int* d_out = nullptr;
    const size_t pb = p->prog.blob.size() * 4;
    RRX_CUDA(cudaMalloc(&d, len + 16));
    RRX_CUDA(cudaMalloc(&d_prog, pb));
    RRX_CUDA(cudaMalloc(&d_out, 4));
    cudaMemcpy(d, text, len, cudaMemcpyHostToDevice);
    cudaMemcpy(d_prog, p->prog.blob.data(), pb, cudaMemcpyHostToDevice);
    cudaError_t e = launch_accepts(p->prog.hdr().NW, d_prog, d, len, d_out, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(accepted, d_out, 4, cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(d_prog); cudaFree(d_out);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "rrx_accepts_host");
}

uint64_t rrx_matches_count(const rrx_matches* m) { return m ? m->count : 0; }
const int64_t* rrx_matches_device_data(const rrx_matches* m) { return reinterpret_cast<const int64_t*>(m->d_pairs); }
const int64_t* rrx_matches_data(rrx_matches* m) {
    if (!m) return nullptr;
    if (m->owner) {
        static const int64_t none[2] = {0, 0};
        return m->pinned.p ? m->pinned.p : none;
    }
    if (!m->host_valid) {
        m->host.resize(2 * m->count + 2);
        if (m->count) {
            cudaError_t e = cudaMemcpy(m->host.data(), m->d_pairs, m->count * 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { cuda_err(e, "cudaMemcpy D2H"); return nullptr; }
        }
        m->host_valid = true;
    }
    return m->host.data();
}
void rrx_matches_free(rrx_matches* m) {
    if (!m) return;
    if (m->owner && m->pinned.p) {
        std::lock_guard<std::mutex> lk(m->owner->mu);
        if (m->owner->pool.size() < 4) m->owner->pool.push_back(m->pinned);
        else cudaFreeHost(m->pinned.p);
    }
    cudaFree(m->d_pairs);
    delete m;
}

int rrx_digest_device(const int64_t* d_pairs, uint64_t n, int64_t end_limit, uint64_t digest[4], void* stream) {
    if (!digest || (!d_pairs && n)) return set_err(RRX_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d = nullptr;
    RRX_CUDA(cudaMalloc(&d, 32));
    cudaError_t e = launch_digest(reinterpret_cast<const long long*>(d_pairs), n, end_limit, d, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(digest, d, 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "rrx_digest_device");
}
int rrx_scanner_digest(rrx_scanner* s, int64_t end_limit, uint64_t digest[4], void* stream) {
    if (!s) return set_err(RRX_ERR_ARG, "null argument");
    rrx_scan_result r;
    int rc = rrx_scanner_result(s, &r, stream);
    if (rc) return rc;
    const uint64_t n = r.n_matches < r.capacity ? r.n_matches : r.capacity;
    return rrx_digest_device(r.d_matches, n, end_limit, digest, stream);
}

// Non-overlapping view of the scanner's last list, with the carry that chains chunks and shards (one
// 8-byte "last selected end" passed left to right): keeps the matches with start >= *carry_io, in list order, each
// selected match moving the carry to its end.
int rrx_scanner_no_overlap(rrx_scanner* s, int64_t* carry_io, rrx_matches** out, void* stream) {
    if (!s || !carry_io || !out) return set_err(RRX_ERR_ARG, "null argument");
    rrx_scan_result r;
    int rc = rrx_scanner_result(s, &r, stream);
    if (rc) return rc;
    if (r.error_bits) return set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(r.error_bits));
    if (r.n_matches > r.capacity) return set_err(RRX_ERR_ARG, "match buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_ptr<rrx_matches> m(new rrx_matches);
    // matches that start before the carry cannot be selected and do not move it: drop them first (they form a
    // prefix-closed condition only per match, so this is a plain filter), then the usual chain from the first survivor
    longlong2* d_in2 = nullptr;
    uint64_t n_in = 0;
    cudaError_t ce = filter_start_ge(reinterpret_cast<const longlong2*>(r.d_matches), r.n_matches, *carry_io, &d_in2, &n_in, st);
    if (ce != cudaSuccess) return cuda_err(ce, "rrx_scanner_no_overlap");
    long long* d_in = reinterpret_cast<long long*>(d_in2);
    long long* filtered = nullptr;
    uint64_t nf = 0;
    rc = filter_no_overlap(d_in, n_in, &filtered, &nf, st);
    cudaFree(d_in);
    if (rc) return rc;
    m->d_pairs = filtered;
    m->count = nf;
    if (nf) {
        long long last[2];
        RRX_CUDA(cudaMemcpy(last, filtered + 2 * (nf - 1), 16, cudaMemcpyDeviceToHost));
        *carry_io = last[1];
    }
    *out = m.release();
    return RRX_OK;
}

int rrx_corpus_iid_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, const uint8_t alpha[256], void* stream) {
    cudaError_t e = launch_corpus_iid(d_out, offset, len, seed, alpha, (cudaStream_t)stream);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "corpus_iid");
}
int rrx_corpus_lines_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, void* stream) {
    cudaError_t e = launch_corpus_lines(d_out, offset, len, seed, (cudaStream_t)stream);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "corpus_lines");
}
int rrx_corpus_records_device(uint8_t* d_out, uint64_t offset, uint64_t len, uint64_t seed, void* stream) {
    cudaError_t e = launch_corpus_records(d_out, offset, len, seed, (cudaStream_t)stream);
    return e == cudaSuccess ? RRX_OK : cuda_err(e, "corpus_records");
}

}  // extern "C"
