// Database-sharded search across the GPUs of one box (BASELINE.json configs[4]).
//
// The reference is a single process (its only parallelism is multiprocessing.Pool over utterances,
// script/synth_halfphone.py:897-903), so there is no reference interface to mirror: this is the exchange step the
// north star adds.  Every rank holds a block of database rows in its own snk_db; queries are replicated.
//
//   snk_knn_sharded_dev : local certified top-k (global row ids)  ->  ONE ncclAllGather of the packed
//                         [dist f64 | id i64] results over NVLink  ->  k-way merge kernel (lowest global id on ties).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): inside a PyTorch process this binds to the copy torch has
// already loaded, in a plain C host to the system library; the engine has no link-time dependency on it.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <vector>

// the slice of nccl.h this file uses (the ABI of these entry points is stable across NCCL 2.x)
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}

constexpr int SNK_MAX_RANKS = 16;

namespace {

struct nccl_api {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
nccl_api g_nccl;

int load_nccl() {
    if (g_nccl.lib) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy already in the process (PyTorch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    SNK_CHECK(h, "libnccl.so.2 not found: %s", dlerror());
#define SYM(field, name)                                                          \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                         \
    SNK_CHECK(g_nccl.field, "libnccl lacks %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GetVersion, "ncclGetVersion");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
    g_nccl.lib = h;
    return 0;
}

#define SNK_NCCL(call)                                                                          \
    do {                                                                                        \
        ncclResult_t r_ = (call);                                                               \
        if (r_ != ncclSuccess) {                                                                \
            snk_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString(r_)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

}  // namespace

struct snk_comm_state {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // a sharded search whose certificates are still pending (snk_knn_sharded_finish)
    bool pending = false;
    int k = 0;
    int64_t nq = 0;
    double *d_dist = nullptr;
    int64_t *d_idx = nullptr;
    cudaStream_t st = nullptr;
    int *d_nfail = nullptr;    // [2]: local failures, sum over ranks
    // peer-memory exchange of the sharded greedy search: every rank owns one region
    //     flags [nranks] u32 (epoch of the last step rank r has delivered), padded to 256 B
    //     data  [2 step parities][nranks writers][P2P_CAP queries] x {dist f64, row i64, bound f64}
    // mapped into every other rank through CUDA IPC; peers[r] is rank r's region as seen from this process
    bool p2p = false;
    char *region = nullptr;
    char *peers[SNK_MAX_RANKS] = {};
    char **d_peers = nullptr;   // the same table on the device
    unsigned epoch = 0;
};
constexpr int P2P_CAP = 4096;                                   // queries per exchange step (more: NCCL path)
constexpr size_t P2P_FLAGS_BYTES = 256;
static size_t p2p_region_bytes(int nranks) { return P2P_FLAGS_BYTES + (size_t)2 * nranks * P2P_CAP * 24; }

void snk_comm_free(snk_db *db) {
    if (!db->comm) return;
    if (db->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(db->comm->comm);
    if (db->comm->d_nfail) cudaFree(db->comm->d_nfail);
    for (int r = 0; r < db->comm->nranks && r < SNK_MAX_RANKS; ++r)
        if (r != db->comm->rank && db->comm->peers[r]) cudaIpcCloseMemHandle(db->comm->peers[r]);
    if (db->comm->region) cudaFree(db->comm->region);
    if (db->comm->d_peers) cudaFree(db->comm->d_peers);
    delete db->comm;
    db->comm = nullptr;
}

namespace {

// exchange + merge of the local results sitting in ws_ag's send half
int exchange_and_merge(snk_db *db, cudaStream_t st) {
    snk_comm_state *c = db->comm;
    const size_t half = (size_t)c->nq * c->k * 8;            // bytes of one [nq, k] array
    char *send = (char *)db->ws_ag.p;
    char *recv = send + snk_round_up(2 * half, 256);
    {
        snk_prof_scope prof(db, SNK_PROF_ALLGATHER, (double)2 * half * (c->nranks - 1), st);   // bytes this rank receives
        SNK_NCCL(g_nccl.AllGather(send, recv, 2 * half, ncclInt8, c->comm, st));
    }
    {
        snk_prof_scope prof(db, SNK_PROF_MERGE, (double)2 * half * c->nranks, st);
        SNK_TRY(snk_topk_merge_launch((const double *)recv, (const int64_t *)(recv + half), (int64_t)(2 * half / 8),
                                      (int64_t)(2 * half / 8), c->nranks, c->nq, c->k, c->d_dist, c->d_idx, st));
    }
    db->counters[2] += 2;
    return 0;
}

}  // namespace

namespace {

// per query the best (distance, global row) pair over the R ranks' answers (ties: lowest global row), judged against every
// rank's bound on the rows it did not look at exactly: the answer stands iff it is not above any of them
__global__ void best_of_ranks_kernel(const double *__restrict__ dist_all, const int64_t *__restrict__ ix_all,
                                     const double *__restrict__ bound_all, int R, int n, double *__restrict__ dist,
                                     int64_t *__restrict__ ix, int *__restrict__ flags, int *__restrict__ count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    double bd = dist_all[b], lb = bound_all ? bound_all[b] : INFINITY;
    int64_t bi = ix_all[b];
    for (int r = 1; r < R; ++r) {
        const double d = dist_all[(size_t)r * n + b];
        const int64_t i = ix_all[(size_t)r * n + b];
        if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
        if (bound_all) lb = fmin(lb, bound_all[(size_t)r * n + b]);
    }
    dist[b] = bd;
    ix[b] = bi;
    if (!(bd <= lb)) {           // some shard may hide a closer row: every rank clears the same flag
        if (flags[b]) atomicAdd(count, 1);
        flags[b] = 0;
    }
}

}  // namespace

// One exchange step of the database-sharded greedy search: every rank's best
// (distance, global row) per utterance and its bound on the rows outside its shortlist -> one grouped ncclAllGather
// (n * 24 bytes per rank) -> arg-min over the ranks, written back in place, certificate flags cleared where the global
// best is not below every bound.  Collective: every rank calls it with the same n.
namespace {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The exchange step as ONE kernel over NVLink peer memory (one block): every thread stores its queries' (distance, row,
// bound) into the slot this rank owns in EVERY rank's region (plain 8-byte stores to IPC-mapped addresses), a system-scope
// fence and one release store per peer publish them under the step's epoch, then the block waits until every rank's epoch
// has arrived in its own region and takes the arg-min / judges the certificate from local memory.  No NCCL launch, no
// host involvement; two step parities because a fast rank may deliver step e + 1 while a slow one still reads step e
// (it cannot deliver e + 2 before every rank has delivered e + 1, i.e. has finished reading e).
__global__ void __launch_bounds__(1024) exchange_best_p2p_kernel(char *const *__restrict__ peers, int rank, int R, int n,
                                                                 unsigned epoch, double *__restrict__ dist,
                                                                 int64_t *__restrict__ ix, const double *__restrict__ bound,
                                                                 int *__restrict__ flags, int *__restrict__ count) {
    const size_t slot_bytes = (size_t)P2P_CAP * 24;
    const size_t mine = P2P_FLAGS_BYTES + ((size_t)(epoch & 1u) * R + rank) * slot_bytes;
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        const double d = dist[b], lb = bound ? bound[b] : INFINITY;
        const int64_t i = ix[b];
        for (int r = 0; r < R; ++r) {
            double *dst = reinterpret_cast<double *>(peers[r] + mine) + 3 * (size_t)b;
            dst[0] = d;
            reinterpret_cast<int64_t *>(dst)[1] = i;
            dst[2] = lb;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < R) st_release_sys(reinterpret_cast<unsigned *>(peers[threadIdx.x]) + rank, epoch);
    if (threadIdx.x < R) {
        const unsigned *f = reinterpret_cast<const unsigned *>(peers[rank]) + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0)
            if (clock64() - t0 > 20000000000LL) __trap();      // a peer that never arrives must not hang the GPU
    }
    __syncthreads();
    const char *base = peers[rank] + P2P_FLAGS_BYTES + (size_t)(epoch & 1u) * R * slot_bytes;
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        const double *e0 = reinterpret_cast<const double *>(base) + 3 * (size_t)b;     // written by remote GPUs: read past L1
        double bd = __ldcg(e0), lb = __ldcg(e0 + 2);
        int64_t bi = __ldcg(reinterpret_cast<const long long *>(e0) + 1);
        for (int r = 1; r < R; ++r) {
            const double *e = reinterpret_cast<const double *>(base + r * slot_bytes) + 3 * (size_t)b;
            const double d = __ldcg(e);
            const int64_t i = __ldcg(reinterpret_cast<const long long *>(e) + 1);
            if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
            lb = fmin(lb, __ldcg(e + 2));
        }
        dist[b] = bd;
        ix[b] = bi;
        if (!(bd <= lb)) {
            if (flags[b]) atomicAdd(count, 1);
            flags[b] = 0;
        }
    }
}

}  // namespace

int snk_comm_exchange_best(snk_db *db, double *d_dist, int64_t *d_ix, double *d_bound, int n, int *d_flags, int *d_count,
                           cudaStream_t st) {
    snk_comm_state *c = db->comm;
    SNK_CHECK(c, "snk_comm_init has not been called");
    if (n <= 0) return 0;
    if (c->p2p && c->nranks > 1 && n <= P2P_CAP) {
        ++c->epoch;
        snk_prof_scope prof(db, SNK_PROF_ALLGATHER, (double)n * 24 * (c->nranks - 1), st);
        exchange_best_p2p_kernel<<<1, 1024, 0, st>>>(c->d_peers, c->rank, c->nranks, n, c->epoch, d_dist, d_ix, d_bound, d_flags,
                                                     d_count);
        SNK_CUDA(cudaGetLastError());
        db->counters[2] += 1;
        return 0;
    }
    const size_t one = snk_round_up((size_t)n * 8 * c->nranks, 256);
    SNK_TRY(snk_buf_reserve(&db->ws_ag, 3 * one));
    double *rd = (double *)db->ws_ag.p;
    int64_t *ri = (int64_t *)((char *)db->ws_ag.p + one);
    double *rb = (double *)((char *)db->ws_ag.p + 2 * one);
    if (c->nranks > 1) {
        snk_prof_scope prof(db, SNK_PROF_ALLGATHER, (double)n * 24 * (c->nranks - 1), st);
        SNK_NCCL(g_nccl.GroupStart());
        SNK_NCCL(g_nccl.AllGather(d_dist, rd, (size_t)n, ncclFloat64, c->comm, st));
        SNK_NCCL(g_nccl.AllGather(d_ix, ri, (size_t)n, ncclInt64, c->comm, st));
        if (d_bound) SNK_NCCL(g_nccl.AllGather(d_bound, rb, (size_t)n, ncclFloat64, c->comm, st));
        SNK_NCCL(g_nccl.GroupEnd());
    } else {
        rd = d_dist; ri = d_ix; rb = d_bound;      // a communicator of one: judge the local answer against the local bound
    }
    if (!d_bound) rb = nullptr;          // exhaustive float64 stage: exact by construction, nothing to judge
    best_of_ranks_kernel<<<(n + 127) / 128, 128, 0, st>>>(rd, ri, rb, c->nranks, n, d_dist, d_ix, d_flags, d_count);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 2;
    return 0;
}

// element-wise minimum of an int array over the ranks, in place (certificate flags: 0 = some rank could not certify)
int snk_comm_allreduce_min(snk_db *db, int *d_buf, int n, cudaStream_t st) {
    snk_comm_state *c = db->comm;
    SNK_CHECK(c, "snk_comm_init has not been called");
    if (n <= 0 || c->nranks == 1) return 0;
    SNK_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclInt32, ncclMin, c->comm, st));
    return 0;
}

int snk_comm_nranks(const snk_db *db) { return db->comm ? db->comm->nranks : 1; }

bool snk_comm_has_p2p(const snk_db *db) { return db->comm && db->comm->p2p && db->comm->nranks > 1; }

int snk_comm_p2p_claim(snk_db *db, int nsteps, char *const **peers, int *rank, int *nranks, unsigned *epoch0, size_t *flags_bytes,
                       size_t *slot_bytes) {
    snk_comm_state *c = db->comm;
    *peers = nullptr;
    if (!c || !c->p2p || c->nranks < 2) return 0;
    *peers = c->d_peers;
    *rank = c->rank;
    *nranks = c->nranks;
    *epoch0 = c->epoch;
    c->epoch += (unsigned)nsteps;
    *flags_bytes = P2P_FLAGS_BYTES;
    *slot_bytes = (size_t)P2P_CAP * 24;
    return 0;
}

extern "C" {

int snk_comm_unique_id(void *id_out, int id_bytes) {
    SNK_CHECK(id_out && id_bytes >= (int)sizeof(ncclUniqueId), "unique id buffer must hold %d bytes", (int)sizeof(ncclUniqueId));
    SNK_TRY(load_nccl());
    ncclUniqueId id;
    SNK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int snk_comm_init(snk_db *db, const void *unique_id, int rank, int nranks) {
    SNK_CHECK(db && unique_id, "NULL argument");
    SNK_LOCK(db);
    SNK_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    SNK_TRY(load_nccl());
    SNK_CUDA(cudaSetDevice(db->device));
    snk_comm_free(db);
    snk_comm_state *c = new snk_comm_state();
    db->comm = c;
    c->rank = rank;
    c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    SNK_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    SNK_CUDA(cudaMalloc((void **)&c->d_nfail, 8));
    // Peer-memory exchange region (sharded greedy): allocate, publish the IPC handle through the communicator, map the
    // peers'.  Every step below that can fail on a given platform (no IPC between the processes, more ranks than the table
    // holds) only switches the exchange to the NCCL path -- on ALL ranks, agreed by an all-reduce.
    int ok = nranks > 1 && nranks <= SNK_MAX_RANKS && !getenv("SNK_COMM_NO_P2P");
    cudaIpcMemHandle_t *d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> handles((size_t)nranks);
    if (nranks > 1) {
        const size_t bytes = p2p_region_bytes(nranks);
        if (ok && cudaMalloc((void **)&c->region, bytes) != cudaSuccess) { ok = 0; cudaGetLastError(); }
        if (ok) SNK_CUDA(cudaMemset(c->region, 0, bytes));
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof(mine));
        if (ok && cudaIpcGetMemHandle(&mine, c->region) != cudaSuccess) { ok = 0; cudaGetLastError(); }
        SNK_CUDA(cudaMalloc((void **)&d_handles, sizeof(cudaIpcMemHandle_t) * (size_t)(nranks + 1)));
        SNK_CUDA(cudaMemcpy(d_handles + nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice));
        SNK_NCCL(g_nccl.AllGather(d_handles + nranks, d_handles, sizeof(mine), ncclInt8, c->comm, db->stream));
        SNK_CUDA(cudaStreamSynchronize(db->stream));
        SNK_CUDA(cudaMemcpy(handles.data(), d_handles, sizeof(mine) * (size_t)nranks, cudaMemcpyDeviceToHost));
        for (int r = 0; r < nranks && ok; ++r) {
            if (r == rank) { c->peers[r] = c->region; continue; }
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, handles[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            c->peers[r] = (char *)ptr;
        }
        // all or none
        SNK_CUDA(cudaMemcpy(c->d_nfail, &ok, 4, cudaMemcpyHostToDevice));
        SNK_NCCL(g_nccl.AllReduce(c->d_nfail, c->d_nfail, 1, ncclInt32, ncclMin, c->comm, db->stream));
        SNK_CUDA(cudaStreamSynchronize(db->stream));
        SNK_CUDA(cudaMemcpy(&ok, c->d_nfail, 4, cudaMemcpyDeviceToHost));
        cudaFree(d_handles);
        if (ok) {
            SNK_CUDA(cudaMalloc((void **)&c->d_peers, sizeof(char *) * SNK_MAX_RANKS));
            SNK_CUDA(cudaMemcpy(c->d_peers, c->peers, sizeof(char *) * SNK_MAX_RANKS, cudaMemcpyHostToDevice));
        }
    }
    c->p2p = ok != 0;
    return 0;
}

int snk_comm_peer_exchange(const snk_db *db) { return db && db->comm && db->comm->p2p ? 1 : 0; }

int snk_comm_info(snk_db *db, int *rank, int *nranks, int *nccl_version) {
    SNK_CHECK(db && db->comm, "snk_comm_init has not been called");
    if (rank) *rank = db->comm->rank;
    if (nranks) *nranks = db->comm->nranks;
    if (nccl_version) g_nccl.GetVersion(nccl_version);
    return 0;
}

int snk_knn_sharded_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                        int64_t id_offset, void *stream) {
    SNK_CHECK(db && db->comm, "snk_comm_init has not been called");
    SNK_LOCK(db);
    SNK_CHECK(space == SNK_SPACE_TARGET || space == SNK_SPACE_JOINT, "unknown search space %d", space);
    SNK_CHECK(nq >= 1 && k >= 1, "bad nq / k");
    SNK_CHECK(!db->comm->pending, "the previous sharded search has not been finished (snk_knn_sharded_finish)");
    SNK_CUDA(cudaSetDevice(db->device));
    cudaStream_t st = (cudaStream_t)stream;
    snk_comm_state *c = db->comm;
    const size_t half = (size_t)nq * k * 8;
    SNK_TRY(snk_buf_reserve(&db->ws_ag, snk_round_up(2 * half, 256) + 2 * half * c->nranks));
    char *send = (char *)db->ws_ag.p;
    // local top-k straight into the send buffer: [nq, k] distances, then [nq, k] global row ids
    SNK_TRY(snk_knn_enqueue(db, space, dQ, nq, k, (double *)send, (int64_t *)(send + half), k, id_offset, st));
    c->pending = true; c->k = k; c->nq = nq; c->d_dist = d_dist; c->d_idx = d_idx; c->st = st;
    return exchange_and_merge(db, st);
}

// Completes a sharded search: local certificate repair, then all ranks agree (one 4-byte all-reduce) whether any of
// them changed its local answer; only then is the exchange repeated.  Collective: every rank must call it.
int snk_knn_sharded_finish(snk_db *db) {
    SNK_CHECK(db && db->comm, "snk_comm_init has not been called");
    SNK_LOCK(db);
    snk_comm_state *c = db->comm;
    if (!c->pending) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    const int64_t before = db->counters[1];
    SNK_TRY(snk_knn_finish(db));                       // syncs; re-searches flagged local queries into the send buffer
    int local = (int)(db->counters[1] - before), total[2] = {0, 0};
    SNK_CUDA(cudaMemcpyAsync(c->d_nfail, &local, 4, cudaMemcpyHostToDevice, c->st));
    SNK_NCCL(g_nccl.AllReduce(c->d_nfail, c->d_nfail + 1, 1, ncclInt32, ncclSum, c->comm, c->st));
    SNK_CUDA(cudaMemcpyAsync(total, c->d_nfail, 8, cudaMemcpyDeviceToHost, c->st));
    SNK_CUDA(cudaStreamSynchronize(c->st));
    if (total[1] > 0) {
        SNK_TRY(exchange_and_merge(db, c->st));
        SNK_CUDA(cudaStreamSynchronize(c->st));
    }
    c->pending = false;
    return 0;
}

}  // extern "C"
