// Data-parallel gradient exchange: the all-reduce(mean) the reference constructs with DDP
// (ref src/src_t5/runner/DistributedRunner.py:26) but never triggers (it calls model.module(...), :63).
// One NCCL communicator per process over the NVSwitch domain; NCCL is resolved at run time with dlopen so that the
// library loads (and single-GPU paths run) without it and so that the process shares torch's libnccl.so.2.
//
// Gradients live in ONE flat fp32 buffer, so a bucket is just a [offset, count) range.  During the fused
// train step the engine hands over ranges as soon as backward has finished them (decoder first, then encoder
// blocks in reverse, embeddings last); they are reduced on a side stream and joined before the optimiser.
#include "engine.h"
#include <dlfcn.h>
#include <mutex>
#include <algorithm>
#include <string.h>

namespace p5 {

typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid_t*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);

static struct {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
} N;

static void nccl_load() {
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            N.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (N.lib) break;
        }
        if (!N.lib) return;
        N.get_uid = (fn_get_uid)dlsym(N.lib, "ncclGetUniqueId");
        N.init_rank = (fn_init_rank)dlsym(N.lib, "ncclCommInitRank");
        N.allreduce = (fn_allreduce)dlsym(N.lib, "ncclAllReduce");
        N.destroy = (fn_destroy)dlsym(N.lib, "ncclCommDestroy");
        N.errstr = (fn_errstr)dlsym(N.lib, "ncclGetErrorString");
    });
    P5_CHECK(N.lib && N.get_uid && N.init_rank && N.allreduce && N.destroy, "NCCL (libnccl.so.2) is not available");
}
static void nccl_check(int rc, const char* what) {
    if (rc != 0) {
        char b[256];
        snprintf(b, sizeof(b), "NCCL error in %s: %s", what, N.errstr ? N.errstr(rc) : "?");
        throw P5Error(5, b);
    }
}

struct CommState {
    nccl_comm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    int64_t reduced_lo = 0, reduced_hi = 0;   // [lo, hi) already handed to NCCL this step (hi grows downward...)
    std::vector<std::pair<int64_t, int64_t>> pending_done;  // ranges reduced this step
};

void comm_unique_id(void* id128) {
    nccl_load();
    nccl_uid_t id;
    nccl_check(N.get_uid(&id), "ncclGetUniqueId");
    memcpy(id128, &id, 128);
}

void comm_init(Engine* e, const void* id128, int rank, int world) {
    nccl_load();
    P5_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    P5_CUDA(cudaSetDevice(e->device));
    CommState* c = new CommState();
    nccl_uid_t id;
    memcpy(&id, id128, 128);
    // The all-reduce runs UNDER the backward's persistent tcgen05 kernels.  P5_COMM_CTAS=n gives it a fixed number of CTAs
    // and takes exactly those SMs out of the persistent grids (common.cuh sm_budget) instead of letting NCCL's default
    // evict GEMM CTAs at random; an NCCL_MAX_CTAS set by the user wins.  Measured at N=2 (profiles/r02_scale_n2.txt):
    // 0 (NCCL default, no reservation) 17.46 ms/step, 8: 17.55, 16: 17.82 -> default 0.  Also measured and dropped: bf16
    // gradients on the wire (cast -> ncclAllReduce(bf16) -> cast back on the comm stream): 17.27 ms/step against 17.12 with
    // fp32 on the same box — at N = 2 the two extra passes over the 892 MB buffer cost more than the halved NVLink bytes save.
    int ctas = 0;
    if (const char* ev = getenv("P5_COMM_CTAS")) ctas = atoi(ev);
    if (ctas > 0 && world > 1) {
        char b[16];
        snprintf(b, sizeof(b), "%d", ctas);
        setenv("NCCL_MAX_CTAS", b, 0);
        setenv("NCCL_MIN_CTAS", b, 0);
        int reserve = ctas;
        if (const char* ev = getenv("P5_GEMM_RESERVE")) reserve = atoi(ev);
        sm_reserve(reserve);
    }
    nccl_check(N.init_rank(&c->comm, world, id, rank), "ncclCommInitRank");
    P5_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    P5_CUDA(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
    P5_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    e->nccl_comm = c;
    e->world = world;
    e->rank = rank;
}

// reduce G[off, off+n) (mean over ranks) on the side stream once the main stream has produced it
void comm_allreduce_range(Engine* e, int64_t off, int64_t n) {
    CommState* c = (CommState*)e->nccl_comm;
    if (!c || e->world <= 1 || n <= 0) return;
    P5_CUDA(cudaEventRecord(c->ready, e->st));
    P5_CUDA(cudaStreamWaitEvent(c->stream, c->ready, 0));
    nccl_check(N.allreduce(e->G + off, e->G + off, (size_t)n, /*ncclFloat*/ 7, /*ncclAvg*/ 4, c->comm, c->stream),
               "ncclAllReduce");
    c->pending_done.push_back({off, off + n});
}

cudaStream_t comm_stream(Engine* e) {
    CommState* c = (CommState*)e->nccl_comm;
    return c ? c->stream : nullptr;
}

// finish: reduce every range not yet handed over, then make the main stream wait for the side stream
void comm_allreduce_grads(Engine* e) {
    CommState* c = (CommState*)e->nccl_comm;
    if (!c || e->world <= 1) return;
    P5_CUDA(cudaSetDevice(e->device));
    // complement of pending_done within [0, n_flat)
    std::vector<std::pair<int64_t, int64_t>> done = c->pending_done;
    std::sort(done.begin(), done.end());
    int64_t cur = 0;
    std::vector<std::pair<int64_t, int64_t>> todo;
    for (auto& r : done) {
        if (r.first > cur) todo.push_back({cur, r.first});
        if (r.second > cur) cur = r.second;
    }
    if (cur < e->n_flat) todo.push_back({cur, e->n_flat});
    for (auto& r : todo) {
        comm_allreduce_range(e, r.first, r.second - r.first);
        if (r.second > e->early_lo) e->early_norm_ready = false;   // a range whose early sum of squares was taken changes
    }
    c->pending_done.clear();
    P5_CUDA(cudaEventRecord(c->done, c->stream));
    P5_CUDA(cudaStreamWaitEvent(e->st, c->done, 0));
    e->norm_valid = false;
}

void comm_destroy(Engine* e) {
    CommState* c = (CommState*)e->nccl_comm;
    if (!c) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm && N.destroy) N.destroy(c->comm);
    cudaEventDestroy(c->ready);
    cudaEventDestroy(c->done);
    cudaStreamDestroy(c->stream);
    delete c;
    e->nccl_comm = nullptr;
}

}  // namespace p5
