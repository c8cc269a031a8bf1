// Sharded scans in the C ABI (include/rrx.h "sharded scans over NCCL"): one contiguous byte range per GPU, ONE
// exchange step.  Per rank and per step, all enqueued on one stream with no host involvement:
//
//   speculative scan of the shard (empty entry)  ->  [composite | counters] message (device copy)
//   ->  ncclAllGather of the messages (NVLink)   ->  fold kernel: composites of the ranks to the left give this
//   rank's exact entry set; flags say whether ANY rank needs more than that  ->  24-word status block to pinned memory
//
// rrx_sharded_finish waits for that one copy.  Common case — no state crosses any shard boundary, no start walk
// leaves its halo, the thread-owned scan held everywhere: the speculative list IS the result, and the next step may
// already be running (two buffer sets: begin k+1 before finish k).  Otherwise every rank sees the same flags and
// all take the slow path together: whole-shard rescans with the exact entry (rounds of all-gather + fold until
// no entry changes), then the hand-over of start walks that ran off a shard's left edge, one rank per hop.
//
// NCCL is loaded at run time (dlopen libnccl.so.2, the library the caller's communicator comes from); librrx has no
// link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rrx.h"
#include "rrx_scan.h"
#include "rrx_tile.cuh"

using namespace rrx;

int rrx_dist_set_err(int code, const std::string& msg);      // rrx_capi.cu (thread-local error text)
extern "C" void rrx_scanner_message_sources(rrx_scanner* s, const uint32_t** d_comp, const uint32_t** d_ctr);
extern "C" void rrx_scanner_peek(rrx_scanner* s, const int64_t** d_matches, uint64_t* capacity);

namespace {

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // the library already in the process first (the communicator was made by it), then the usual names
        const char* cands[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* c : cands) { n.h = dlopen(c, RTLD_NOW | RTLD_NOLOAD); if (n.h) break; }
        if (!n.h) for (const char* c : cands) { n.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
        if (!n.h) return;
        *(void**)(&n.AllGather) = dlsym(n.h, "ncclAllGather");
        *(void**)(&n.GetErrorString) = dlsym(n.h, "ncclGetErrorString");
        n.ok = n.AllGather && n.GetErrorString;
    });
    return n;
}

enum : uint32_t { ST_ENTRY = 0, ST_FLAGS = 32, ST_MATCHES_LO = 33, ST_MATCHES_HI = 34, ST_PENDING = 35, ST_ERR = 36, ST_ANY_PENDING = 37, ST_WORDS = 40 };
// FL_MORE: some rank's exact entry (fold of the current composites) differs from the entry its current list was scanned
// with — that rank rescans, everybody exchanges again.  FL_FALLBACK: the thread-owned scan did not hold on some rank.
enum : uint32_t { FL_MORE = 1u, FL_FALLBACK = 8u, FL_FATAL = 16u };

// One thread: fold the gathered messages — composite (1+3*NW words), the 8 counter words of rrx_scanner_message_device,
// then the entry set the rank's current list was scanned with (NW words) — into this rank's exact entry set, and
// summarise everybody's state.  Every rank computes the same flags from the same gathered data.
template <int NW>
__global__ void fold_status_kernel(const uint32_t* __restrict__ msgs, uint32_t msg_words, uint32_t nranks, uint32_t rank, uint32_t* status) {
    if (threadIdx.x || blockIdx.x) return;
    constexpr uint32_t PW = 1 + 3 * NW;
    Comp<NW> acc, x, nxt;
    uint32_t flags = 0, any_pending = 0;
    comp_ident(acc);                                   // composite of the ranks left of r
    for (uint32_t r = 0; r < nranks; r++) {
        const uint32_t* m = msgs + (size_t)r * msg_words;
        uint32_t zero[NW], e[NW];
        for (int w = 0; w < NW; w++) zero[w] = 0;
        comp_apply<NW>(acc, zero, e);                  // exact entry of rank r, given the composites as they are now
        for (int w = 0; w < NW; w++) if (e[w] != m[PW + 8 + w]) flags |= FL_MORE;
        if (r == rank) for (int w = 0; w < NW; w++) status[ST_ENTRY + w] = e[w];
        const uint32_t* c = m + PW;
        const uint32_t err = c[1];
        if (err & ~(uint32_t)(ERR_VERIFY_EDGE | ERR_VERIFY_BUDGET)) flags |= FL_FATAL;
        if ((err & ERR_VERIFY_BUDGET) || ((err & ERR_VERIFY_EDGE) && r > 0)) flags |= FL_FALLBACK;
        if (c[4] | c[5]) any_pending = 1;
        if (r == rank) {
            status[ST_MATCHES_LO] = c[6]; status[ST_MATCHES_HI] = c[7];
            status[ST_PENDING] = c[4];
            status[ST_ERR] = err;
        }
        x.kind = m[0];
        for (int w = 0; w < NW; w++) { x.b[w] = m[1 + w]; x.U[w] = m[1 + NW + w]; x.P[w] = m[1 + 2 * NW + w]; }
        comp_compose(acc, x, nxt);
        acc = nxt;
    }
    status[ST_FLAGS] = flags;
    status[ST_ANY_PENDING] = any_pending;
}

cudaError_t launch_fold_status(uint32_t NW, const uint32_t* msgs, uint32_t msg_words, uint32_t nranks, uint32_t rank, uint32_t* status, cudaStream_t st) {
    switch (NW) {
        case 1: fold_status_kernel<1><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
        case 2: fold_status_kernel<2><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
        case 4: fold_status_kernel<4><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
        case 8: fold_status_kernel<8><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
        case 16: fold_status_kernel<16><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
        default: fold_status_kernel<32><<<1, 32, 0, st>>>(msgs, msg_words, nranks, rank, status); break;
    }
    return cudaGetLastError();
}

// the rank's message [composite | 8 counter words | entry used] in one launch
struct UsedEntry { uint32_t w[RRX_NW_MAX]; };
__global__ void pack_message_kernel(const uint32_t* __restrict__ comp, const uint32_t* __restrict__ ctr, const __grid_constant__ UsedEntry used,
                                    uint32_t nw, uint32_t* __restrict__ msg) {
    const uint32_t pw = 1 + 3 * nw, i = threadIdx.x;
    if (i < pw) msg[i] = comp[i];
    else if (i < pw + 8) msg[i] = ctr[i - pw];
    else if (i < pw + 8 + nw) msg[i] = used.w[i - pw - 8];
}

__global__ void patch_starts_kernel(long long* pairs, const unsigned long long* idx, const long long* start, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pairs[2 * idx[i]] = start[i];
}

}  // namespace

struct rrx_sharded {
    const rrx_program* p = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    uint32_t NW = 1, msg_words = 0;
    uint64_t max_len = 0, capacity = 0;
    rrx_scanner* sc[2] = {nullptr, nullptr};        // two buffer sets: step k+1 may be enqueued before step k is collected
    uint32_t* d_msg[2] = {nullptr, nullptr};
    uint32_t* d_all[2] = {nullptr, nullptr};
    uint32_t* d_status[2] = {nullptr, nullptr};
    uint32_t* h_status[2] = {nullptr, nullptr};     // pinned
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // The exchange (pack, all-gather, fold, status copy) runs on a stream of its own, behind an event of the caller's
    // stream: with pipelined steps it overlaps the NEXT step's scan instead of standing in front of it.
    cudaStream_t xs = nullptr;
    cudaEvent_t ev_scan[2] = {nullptr, nullptr};
    int turn = 0;
    struct Step { const uint8_t* text; uint64_t len, base, left_ctx; void* stream; int slot; bool live; } step[2] = {};
    int oldest = 0;
    // result of the last finished step
    rrx_scanner* res_sc = nullptr;
    uint64_t res_n = 0;
    uint32_t res_rescans = 0, res_patched = 0;
};

#define RRX_DCUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return rrx_dist_set_err(RRX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define RRX_DNCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return rrx_dist_set_err(RRX_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r_)); } while (0)

extern "C" {

void rrx_sharded_free(rrx_sharded* s) {
    if (!s) return;
    for (int i = 0; i < 2; i++) {
        if (s->sc[i]) rrx_scanner_free(s->sc[i]);
        cudaFree(s->d_msg[i]); cudaFree(s->d_all[i]); cudaFree(s->d_status[i]);
        if (s->h_status[i]) cudaFreeHost(s->h_status[i]);
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
        if (s->ev_scan[i]) cudaEventDestroy(s->ev_scan[i]);
    }
    if (s->xs) cudaStreamDestroy(s->xs);
    delete s;
}

int rrx_sharded_create(const rrx_program* p, uint64_t max_len, uint64_t capacity, void* nccl_comm, int rank, int nranks, rrx_sharded** out) {
    if (!p || !out || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return rrx_dist_set_err(RRX_ERR_ARG, "bad argument");
    if (!nccl().ok) return rrx_dist_set_err(RRX_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
    rrx_info info;
    rrx_program_info(p, &info);
    rrx_sharded* s = new rrx_sharded;
    s->p = p; s->comm = (ncclComm_t)nccl_comm; s->rank = rank; s->nranks = nranks;
    s->NW = info.words; s->msg_words = 1 + 3 * info.words + 8 + info.words;
    s->max_len = max_len; s->capacity = capacity;
    for (int i = 0; i < 2; i++) {
        int rc = rrx_scanner_create(p, max_len, capacity, &s->sc[i]);
        if (rc) { rrx_sharded_free(s); return rc; }
        cudaError_t e = cudaMalloc(&s->d_msg[i], s->msg_words * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_all[i], (size_t)s->msg_words * 4 * nranks);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_status[i], ST_WORDS * 4);
        if (e == cudaSuccess) e = cudaMallocHost(&s->h_status[i], ST_WORDS * 4);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_scan[i], cudaEventDisableTiming);
        if (e == cudaSuccess && !s->xs) e = cudaStreamCreateWithFlags(&s->xs, cudaStreamNonBlocking);
        if (e != cudaSuccess) { rrx_sharded_free(s); return rrx_dist_set_err(RRX_ERR_CUDA, std::string("rrx_sharded_create: ") + cudaGetErrorString(e)); }
    }
    *out = s;
    return RRX_OK;
}

// `used`: the entry set the scanner's current list was scanned with (NULL: the empty set).  `st`: the stream the scan
// was enqueued on; the exchange follows it on the side stream.
static int enqueue_exchange(rrx_sharded* s, int slot, rrx_scanner* sc, const uint32_t* used, cudaStream_t st) {
    const uint32_t *d_comp = nullptr, *d_ctr = nullptr;
    rrx_scanner_message_sources(sc, &d_comp, &d_ctr);
    UsedEntry u;
    for (uint32_t w = 0; w < RRX_NW_MAX; w++) u.w[w] = (used && w < s->NW) ? used[w] : 0u;
    RRX_DCUDA(cudaEventRecord(s->ev_scan[slot], st));
    RRX_DCUDA(cudaStreamWaitEvent(s->xs, s->ev_scan[slot], 0));
    pack_message_kernel<<<1, 160, 0, s->xs>>>(d_comp, d_ctr, u, s->NW, s->d_msg[slot]);
    RRX_DCUDA(cudaGetLastError());
    RRX_DNCCL(nccl().AllGather(s->d_msg[slot], s->d_all[slot], s->msg_words, ncclUint32, s->comm, s->xs));
    RRX_DCUDA(launch_fold_status(s->NW, s->d_all[slot], s->msg_words, (uint32_t)s->nranks, (uint32_t)s->rank, s->d_status[slot], s->xs));
    RRX_DCUDA(cudaMemcpyAsync(s->h_status[slot], s->d_status[slot], ST_WORDS * 4, cudaMemcpyDeviceToHost, s->xs));
    RRX_DCUDA(cudaEventRecord(s->ev[slot], s->xs));
    return RRX_OK;
}

int rrx_sharded_begin(rrx_sharded* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset, uint64_t left_context, void* stream) {
    if (!s) return rrx_dist_set_err(RRX_ERR_ARG, "null argument");
    const int slot = s->turn;
    if (s->step[slot].live) return rrx_dist_set_err(RRX_ERR_ARG, "rrx_sharded_begin: two steps are already in flight (call rrx_sharded_finish)");
    cudaStream_t st = (cudaStream_t)stream;
    rrx_scan_opts o;
    std::memset(&o, 0, sizeof(o));
    o.left_context = left_context;
    int rc = rrx_scanner_scan(s->sc[slot], d_text, len, base_offset, &o, stream);
    if (rc) return rc;
    rc = enqueue_exchange(s, slot, s->sc[slot], nullptr, st);
    if (rc) return rc;
    s->step[slot] = {d_text, len, base_offset, left_context, stream, slot, true};
    if (!s->step[slot ^ 1].live) s->oldest = slot;
    s->turn ^= 1;
    return RRX_OK;
}

// slow path pieces -----------------------------------------------------------------------------------------------
static int rescan_round(rrx_sharded* s, const rrx_sharded::Step& sp, rrx_scanner* sc, const uint32_t* entry, bool no_donly, cudaStream_t st) {
    rrx_scan_opts o;
    std::memset(&o, 0, sizeof(o));
    o.left_context = sp.left_ctx;
    o.entry = entry;
    o.flags = no_donly ? RRX_SCAN_NO_DONLY : 0;
    return rrx_scanner_scan(sc, sp.text, sp.len, sp.base, &o, sp.stream);
}

static int hand_over_pending(rrx_sharded* s, const rrx_sharded::Step& sp, rrx_scanner* sc, cudaStream_t main_st, uint32_t* patched) {
    // (every collective of the communicator goes to the side stream; the caller's stream is idle here — the step's
    // event has been waited for — and is synchronised before the walks are continued on it)
    cudaStream_t st = s->xs;
    if (cudaStreamSynchronize(main_st) != cudaSuccess) return rrx_dist_set_err(RRX_ERR_CUDA, "pending hand-over: stream");
    // Start walks that ran off the left edge of a rank's readable text travel left, one rank per hop, until some rank
    // finds the start.  Fixed-size exchange: up to KMAX records per rank and hop.
    constexpr uint32_t KMAX = 64;
    struct Rec { uint64_t match_index; int64_t next_offset; uint32_t R[RRX_NW_MAX]; uint32_t owner, live; };
    static_assert(sizeof(Rec) % 4 == 0, "record size");
    const uint32_t RW = sizeof(Rec) / 4;
    std::vector<Rec> mine(KMAX), all((size_t)KMAX * s->nranks);
    std::memset(mine.data(), 0, mine.size() * sizeof(Rec));
    uint64_t n = 0;
    {
        std::vector<uint8_t> raw(KMAX * sizeof(PendingStart));
        int rc = rrx_scanner_copy_pending(sc, raw.data(), KMAX, &n, sp.stream);
        if (rc) return rc;
        if (n > KMAX) return rrx_dist_set_err(RRX_ERR_CUDA, "too many start walks cross the shard's left edge");
        const PendingStart* ps = reinterpret_cast<const PendingStart*>(raw.data());
        for (uint64_t i = 0; i < n; i++) {
            mine[i].match_index = ps[i].match_index; mine[i].next_offset = ps[i].next_offset;
            std::memcpy(mine[i].R, ps[i].R, sizeof(mine[i].R));
            mine[i].owner = (uint32_t)s->rank; mine[i].live = 1;
        }
    }
    Rec *d_mine = nullptr, *d_all = nullptr;
    RRX_DCUDA(cudaMalloc(&d_mine, KMAX * sizeof(Rec)));
    RRX_DCUDA(cudaMalloc(&d_all, (size_t)KMAX * sizeof(Rec) * s->nranks));
    struct Found { uint64_t match_index; int64_t start; uint32_t owner, live; };
    std::vector<Found> found;
    std::vector<Rec> flight = mine;
    int rc = RRX_OK;
    for (int hop = 1; hop < s->nranks && rc == RRX_OK; hop++) {
        cudaMemcpyAsync(d_mine, flight.data(), KMAX * sizeof(Rec), cudaMemcpyHostToDevice, st);
        if (nccl().AllGather(d_mine, d_all, (size_t)KMAX * RW, ncclUint32, s->comm, st) != ncclSuccess) { rc = rrx_dist_set_err(RRX_ERR_CUDA, "ncclAllGather (pending hand-over)"); break; }
        cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(Rec), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = rrx_dist_set_err(RRX_ERR_CUDA, "pending hand-over: stream"); break; }
        std::vector<Rec> todo;
        for (auto& r : all) if (r.live && (int)r.owner - hop == s->rank) todo.push_back(r);
        for (auto& f : flight) f.live = 0;
        if (!todo.empty()) {
            std::vector<uint8_t> raw(todo.size() * sizeof(PendingStart));
            PendingStart* ps = reinterpret_cast<PendingStart*>(raw.data());
            for (size_t i = 0; i < todo.size(); i++) { ps[i].match_index = todo[i].match_index; ps[i].next_offset = todo[i].next_offset; std::memcpy(ps[i].R, todo[i].R, sizeof(ps[i].R)); }
            std::vector<int64_t> starts(todo.size());
            rc = rrx_continue_starts(s->p, sp.text, sp.len, sp.left_ctx, sp.base, raw.data(), todo.size(), starts.data(), sp.stream);
            if (rc) break;
            size_t k = 0;
            for (size_t i = 0; i < todo.size(); i++) {
                if (starts[i] >= 0) found.push_back({todo[i].match_index, starts[i], todo[i].owner, 1});
                else if (starts[i] == -1 && k < KMAX) {
                    flight[k] = todo[i];
                    flight[k].next_offset = ps[i].next_offset;
                    std::memcpy(flight[k].R, ps[i].R, sizeof(ps[i].R));
                    flight[k].live = 1;
                    k++;
                }
            }
        }
    }
    // everybody learns what was found; owners patch their lists
    if (rc == RRX_OK) {
        const uint32_t FW = sizeof(Found) / 4;
        std::vector<Found> fm(KMAX), fa((size_t)KMAX * s->nranks);
        std::memset(fm.data(), 0, fm.size() * sizeof(Found));
        for (size_t i = 0; i < found.size() && i < KMAX; i++) fm[i] = found[i];
        cudaMemcpyAsync(d_mine, fm.data(), KMAX * sizeof(Found), cudaMemcpyHostToDevice, st);
        if (nccl().AllGather(d_mine, d_all, (size_t)KMAX * FW, ncclUint32, s->comm, st) != ncclSuccess) rc = rrx_dist_set_err(RRX_ERR_CUDA, "ncclAllGather (found starts)");
        cudaMemcpyAsync(fa.data(), d_all, fa.size() * sizeof(Found), cudaMemcpyDeviceToHost, st);
        if (rc == RRX_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = rrx_dist_set_err(RRX_ERR_CUDA, "found starts: stream");
        if (rc == RRX_OK) {
            std::vector<unsigned long long> idx;
            std::vector<long long> st_;
            for (auto& f : fa) if (f.live && (int)f.owner == s->rank) { idx.push_back(f.match_index); st_.push_back(f.start); }
            if (!idx.empty()) {
                rrx_scan_result r;
                rc = rrx_scanner_result(sc, &r, sp.stream);
                unsigned long long* d_i = nullptr; long long* d_s = nullptr;
                if (rc == RRX_OK && cudaMalloc(&d_i, idx.size() * 8) == cudaSuccess && cudaMalloc(&d_s, idx.size() * 8) == cudaSuccess) {
                    cudaMemcpyAsync(d_i, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice, st);
                    cudaMemcpyAsync(d_s, st_.data(), idx.size() * 8, cudaMemcpyHostToDevice, st);
                    patch_starts_kernel<<<(unsigned)((idx.size() + 127) / 128), 128, 0, st>>>(reinterpret_cast<long long*>(const_cast<int64_t*>(r.d_matches)), d_i, d_s, (uint32_t)idx.size());
                    cudaStreamSynchronize(st);
                }
                cudaFree(d_i); cudaFree(d_s);
                *patched = (uint32_t)idx.size();
            }
        }
    }
    cudaFree(d_mine); cudaFree(d_all);
    return rc;
}

int rrx_sharded_finish(rrx_sharded* s, rrx_sharded_result* out) {
    if (!s || !out) return rrx_dist_set_err(RRX_ERR_ARG, "null argument");
    const int slot = s->oldest;
    rrx_sharded::Step& sp = s->step[slot];
    if (!sp.live) return rrx_dist_set_err(RRX_ERR_ARG, "rrx_sharded_finish: no step in flight");
    cudaStream_t st = (cudaStream_t)sp.stream;
    RRX_DCUDA(cudaEventSynchronize(s->ev[slot]));
    sp.live = false;
    s->oldest = slot ^ 1;
    const uint32_t* hs = s->h_status[slot];
    uint32_t flags = hs[ST_FLAGS];
    rrx_scanner* sc = s->sc[slot];
    uint32_t rescans = 0, patched = 0;
    if (flags & FL_FATAL) return rrx_dist_set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(hs[ST_ERR]));
    if (flags & (FL_MORE | FL_FALLBACK)) {
        // Slow path, taken by ALL ranks together (the flags derive from the gathered messages, identical everywhere).
        // FL_FALLBACK first: everybody rescans without the thread-owned scan.  Then rounds of {ranks whose exact entry
        // differs from the one their list was scanned with rescan their shard with it; all-gather; fold} until every
        // rank's list belongs to its exact entry — at most one round per rank of a chain of dependent shards.
        const bool no_donly = (flags & FL_FALLBACK) != 0;
        uint32_t used[RRX_NW_MAX];
        for (uint32_t w = 0; w < RRX_NW_MAX; w++) used[w] = 0u;
        if (no_donly) {
            int rc = rescan_round(s, sp, sc, nullptr, true, st);
            if (rc) return rc;
            rescans++;
            rc = enqueue_exchange(s, slot, sc, used, st);
            if (rc) return rc;
            RRX_DCUDA(cudaEventSynchronize(s->ev[slot]));
            flags = hs[ST_FLAGS];
        }
        for (int round = 0; (flags & FL_MORE) && !(flags & FL_FATAL) && round <= s->nranks + 1; round++) {
            bool mine = false;
            for (uint32_t w = 0; w < s->NW; w++) if (hs[ST_ENTRY + w] != used[w]) mine = true;
            if (mine) {
                for (uint32_t w = 0; w < s->NW; w++) used[w] = hs[ST_ENTRY + w];
                int rc = rescan_round(s, sp, sc, used, no_donly, st);
                if (rc) return rc;
                rescans++;
            }
            int rc = enqueue_exchange(s, slot, sc, used, st);
            if (rc) return rc;
            RRX_DCUDA(cudaEventSynchronize(s->ev[slot]));
            flags = hs[ST_FLAGS];
        }
        if (flags & FL_FATAL) return rrx_dist_set_err(RRX_ERR_CUDA, "scan kernel reported error bits " + std::to_string(hs[ST_ERR]));
        if (flags & FL_MORE) return rrx_dist_set_err(RRX_ERR_CUDA, "sharded scan: entry sets did not settle");
    }
    if (hs[ST_ANY_PENDING]) {
        int rc = hand_over_pending(s, sp, sc, st, &patched);
        if (rc) return rc;
    }
    // counts come from the status block of the last exchange: no synchronisation with the caller's stream, on which the
    // next step's scan may already be running
    out->n_matches = (uint64_t)hs[ST_MATCHES_LO] | ((uint64_t)hs[ST_MATCHES_HI] << 32);
    out->rescans = rescans;
    out->patched_starts = patched;
    out->flags = flags;
    uint64_t capacity = 0;
    rrx_scanner_peek(sc, &out->d_matches, &capacity);
    if (out->n_matches > capacity) return rrx_dist_set_err(RRX_ERR_ARG, "match buffer too small for this shard");
    out->scanner = sc;
    s->res_sc = sc; s->res_n = out->n_matches; s->res_rescans = rescans; s->res_patched = patched;
    return RRX_OK;
}

int rrx_scan_sharded(rrx_sharded* s, const uint8_t* d_text, uint64_t len, uint64_t base_offset, uint64_t left_context, void* stream,
                     rrx_sharded_result* out) {
    int rc = rrx_sharded_begin(s, d_text, len, base_offset, left_context, stream);
    if (rc) return rc;
    return rrx_sharded_finish(s, out);
}

float rrx_sharded_kernel_ms(rrx_sharded* s) {
    float ms = 0.f;
    if (s && s->res_sc && rrx_scanner_kernel_ms(s->res_sc, &ms) == RRX_OK) return ms;
    return -1.f;
}
int rrx_sharded_enable_timing(rrx_sharded* s, int on) {
    if (!s) return rrx_dist_set_err(RRX_ERR_ARG, "null argument");
    for (int i = 0; i < 2; i++) { int rc = rrx_scanner_enable_timing(s->sc[i], on); if (rc) return rc; }
    return RRX_OK;
}

}  // extern "C"
