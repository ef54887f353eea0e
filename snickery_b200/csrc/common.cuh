// Shared declarations for the snk_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <vector>
#include <utility>
#include <mutex>
#include "snk_b200.h"

void snk_set_error(const char *fmt, ...);

#define SNK_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            snk_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                  \
                          cudaGetErrorString(e_));                                              \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define SNK_CHECK(cond, ...)                                                                    \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            snk_set_error(__VA_ARGS__);                                                         \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define SNK_TRY(expr)                                                                           \
    do {                                                                                        \
        int r_ = (expr);                                                                        \
        if (r_) return r_;                                                                      \
    } while (0)

static inline int64_t snk_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t snk_cdiv(int64_t x, int64_t m) { return (x + m - 1) / m; }

// grow-only device scratch buffer
struct snk_buf {
    void *p = nullptr;
    size_t cap = 0;
};
int snk_buf_reserve(snk_buf *b, size_t bytes);
void snk_buf_free(snk_buf *b);

// How one searchable row is assembled from the resident matrices.
// row u, dim d:  d <  dA : A[(u + a_row_off) * ldA + a_col + d]
//                d >= dA : B[u * ldB + (d - dA)]         (window of m consecutive frames is contiguous)
struct snk_space {
    int64_t rows;   // number of searchable rows
    int D;          // dA + dB
    int dA, dB;
    int a_row_off, a_col;
    int ldA_raw, ldB_raw;   // leading dims of the raw f32 matrices (Dj, Dt)
    int ldA32, ldB32;       // leading dims of the weighted f32 matrices
};

// One slot of the pinned staging ring: small host->device uploads of the `_dev` entry points (launch metadata)
// are copied here first, so cudaMemcpyAsync is truly asynchronous and the caller's buffer may die at once.
struct snk_stage_slot {
    void *host = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;   // recorded after the copy was enqueued; waited on (host side) before the slot is reused
};
constexpr int SNK_STAGE_SLOTS = 8;

struct snk_comm_state;     // comm.cu: NCCL communicator + exchange workspaces
struct snk_pending_state;  // search.cu: deferred exactness certificates of the last _dev search / greedy batch
struct snk_acoustic_job;   // api.cu: arguments of an snk_acoustic_viterbi_batch_dev awaiting its finish

struct snk_db {
    std::recursive_mutex mu;   // entry points on one handle are serialised; different handles run concurrently
    int device = 0;
    int sm_count = 148;
    int64_t N = 0, Np = 0;
    int Dt = 0, Dj = 0, m = 1;
    unsigned layout = 0;
    // greedy join-context geometry (see SNK_LAYOUT_*)
    int Djq = 0, prev_col = 0, prev_row_off = 0, cur_col = 0, cur_row_off = 0;
    int engine = SNK_ENGINE_AUTO;
    bool weights_set = false;
    int debug_fail_mod = 0;      // SNK_DEBUG_CERT_FAIL=n: pretend every n-th query failed its tensor-core certificate (tests)
    int debug_fail_mod2 = 0;     // SNK_DEBUG_CERT_FAIL2=n: ... its fp32 certificate too (exercises the exhaustive scan)
    bool tc_ok = false;          // fp16 operands of the current weighting are finite (no overflow)
    // database-sharded greedy search: the re-rank writes, per query, the distance below which no row outside its shortlist
    // can lie (instead of clearing certificate flags); the ranks judge the exchanged answer against all bounds (comm.cu)
    double *cert_bound_out = nullptr;
    // resident arrays
    float *F_raw = nullptr;   // [N, Dt]
    float *Jc_raw = nullptr;  // [N+1, Dj]
    double *wt = nullptr;     // [Dt]
    double *wj = nullptr;     // [Dj]
    // target standardisation (snk_db_set_standardisation): lets the search read un-normalised f32 speech
    double *std_mean = nullptr, *std_sd = nullptr;   // [Dt] each
    double uv_special = -1000.0, uv_scale = 20.0;
    bool std_set = false;
    int std_f32 = 0;          // SNK_STD_FLOAT32
    float *Fw32 = nullptr;    // [N, Dt]       weighted, rounded to f32
    float *Jw32 = nullptr;    // [N+1, ldJ32]  weighted, rounded to f32, zero padded
    int ldJ32 = 0;
    // fp16 operands of the tensor-core kernel
    __half *G16 = nullptr;    // [N + pad, ldG16] target frames
    __half *S16 = nullptr;    // [N+1 + pad, ldS16] join contexts
    int ldG16 = 0, ldS16 = 0;
    float *nrm_t16 = nullptr;   // [N]   ||fp16(Fw[u])||^2
    float *nrm_j16 = nullptr;   // [Np]  ||fp16(joint row u)||^2
    float *err_t16 = nullptr;   // [1]   max_u ||Fw[u] - fp16(Fw[u])||      (target space)
    float *err_j16 = nullptr;   // [1]   max_u ||joint row - fp16(joint row)||
    float *maxn_t16 = nullptr;  // [1]   max_u nrm_t16
    float *maxn_j16 = nullptr;  // [1]
    // power-of-two operand scale of the fused join-cost + Viterbi kernel (join_tc.cu), refreshed on first use after a re-weighting
    float *split_sc = nullptr;  // {s, 1/s, max |weighted join value|}
    bool jsplit_valid = false;
    unsigned long long jv_stats[2] = {0, 0};   // last snk_join_tiles: finite entries, entries recomputed by direct differences
    void *tc_state = nullptr;   // tensor maps etc. (knn_tc.cu)
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host entry points: uploads that overlap the first search steps
    // greedy batches may wait on upload events before given steps (step index, event), sorted by step
    std::vector<std::pair<int64_t, cudaEvent_t>> step_waits;
    std::vector<cudaEvent_t> upload_events;
    cudaEvent_t ev = nullptr;
    snk_buf ws_q, ws_dist, ws_list, ws_misc, ws_io, ws_io2, ws_tiles, ws_bp, ws_tc, ws_h0, ws_h1, ws_h2, ws_h3, ws_flags;
    snk_buf ws_kflags, ws_meta, ws_ag, ws_jv, ws_g1;
    snk_stage_slot stage[SNK_STAGE_SLOTS];
    int stage_next = 0;
    snk_comm_state *comm = nullptr;
    snk_pending_state *pending = nullptr;
    snk_acoustic_job *acoustic = nullptr;
    int g1_resident = 0;         // greedy_one.cu: 1 = the cooperative single-utterance kernel fits this device, -1 = it does not
    int64_t counters[4] = {0, 0, 0, 0};
    // optional kernel timing (snk_db_profile_*)
    bool prof_on = false;
    struct prof_rec { int which; cudaEvent_t e0, e1; double work; };
    std::vector<prof_rec> prof;
};

// bracket a hot-kernel launch with events when profiling is enabled
struct snk_prof_scope {
    snk_db *db; cudaStream_t st; int idx;
    snk_prof_scope(snk_db *d, int which, double work, cudaStream_t s) : db(d), st(s), idx(-1) {
        if (!db->prof_on) return;
        snk_db::prof_rec r{which, nullptr, nullptr, work};
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
        cudaEventRecord(r.e0, st);
        db->prof.push_back(r);
        idx = (int)db->prof.size() - 1;
    }
    ~snk_prof_scope() { if (idx >= 0) cudaEventRecord(db->prof[idx].e1, st); }
};

#define SNK_ENGINE_EXACT 3   // internal: exhaustive float64 scan (last stage of the certificate chain)

#define SNK_LOCK(db) std::lock_guard<std::recursive_mutex> snk_lock_guard_((db)->mu)

snk_space snk_make_space(const snk_db *db, int space);

// api.cu: enqueue a host->device copy through the pinned staging ring (never blocks on the GPU unless the ring wraps
// around work that is still pending eight uploads later)
int snk_upload_async(snk_db *db, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);

// search.cu / comm.cu: lifetime hooks called by snk_db_destroy
void snk_pending_destroy(snk_db *db);
void snk_comm_free(snk_db *db);
// comm.cu: exchange step of the database-sharded greedy search, flag agreement, communicator size (1 without one)
int snk_comm_exchange_best(snk_db *db, double *d_dist, int64_t *d_ix, double *d_bound, int n, int *d_flags, int *d_count,
                           cudaStream_t st);
int snk_comm_allreduce_min(snk_db *db, int *d_buf, int n, cudaStream_t st);
int snk_comm_nranks(const snk_db *db);

// ---- weights.cu
int snk_apply_weights(snk_db *db, cudaStream_t st);

// ---- join_viterbi.cu / join_tc.cu
struct snk_vit_meta {
    int64_t frame_off;   // first frame of the utterance in cand / tdist / paths
    int64_t tile_off;    // first lattice step (join tile) of the utterance
    int64_t T;
};
bool snk_join_tc_supported(const snk_db *db, int K);
void snk_join_tc_free(snk_db *db);
int snk_join_tc_launch(snk_db *db, const int64_t *d_cand, int K, const int *d_tile2frame, int64_t ntiles, float *d_tiles,
                       unsigned long long *d_stats, cudaStream_t st);

// ---- knn_simt.cu : fp32 direct-difference shortlist
// Q32 [nq, ldq] f32 queries.  Writes the KP smallest (dist^2 f32, row id i32) per query, unsorted.
int snk_shortlist_simt(snk_db *db, const snk_space &sp, const float *dQ32, int ldq, int64_t nq, int KP,
                       float *d_val, int *d_id, cudaStream_t st);
// generic row-wise top-KP scan (see knn_simt.cu)
// d_seg_cnt (optional): the row is n / seg_cap buffers of seg_cap slots, only the first d_seg_cnt[q, buffer] filled
int snk_topk_scan(snk_db *db, const float *d_vals, const int *d_ids, int64_t nq, int64_t n, int64_t ld,
                  int id_base, int KP, bool init, float *d_val, int *d_id, cudaStream_t st,
                  const int *d_seg_cnt = nullptr, int seg_cap = 0);

// ---- knn_tc.cu : tcgen05 shortlist
bool snk_tc_supported(const snk_db *db, const snk_space &sp, int KP);
int snk_tc_prepare(snk_db *db);   // (re)build tensor maps / K-block tables after create
void snk_tc_destroy(snk_db *db);
int snk_tc_query_ld(const snk_db *db, int space);           // leading dim of the fp16 query operand
const short *snk_tc_qmap(const snk_db *db, int space);      // device map: operand column -> query dim (or -1)
// dQ16 [round_up(nq,128), ldq16] fp16 queries in K-block order.  Writes the KP smallest approximate
// keys (||y~||^2 - 2 x~.y~, f32) with row ids per query (unsorted) and d_tau[q], a lower bound on the
// key of every row that was dropped before the final merge (+inf if none was).
// If `lists` is given and the per-chunk lists are small enough for the fused merge+re-rank kernel, the
// merge is skipped: lists->valid is set and the raw [nq, nlists, lsz] lists are returned instead.
struct snk_tc_lists { bool valid; const float *val; const int *id; int nlists, lsz; };
int snk_shortlist_tc(snk_db *db, int space, const __half *dQ16, int ldq16, int64_t nq, int k, int KP,
                     float *d_val, int *d_id, float *d_tau, snk_tc_lists *lists, cudaStream_t st);

// ---- rerank.cu
// Exact float64 distances of the shortlisted rows, sorted ascending (ties: lowest id).
// d_cert (optional, [nq + 1]): per query 1 = certified exact / 0 = needs SIMT re-search, from the
// approx threshold (largest shortlist key, optionally capped by d_tau_extra[q]) and the fp16
// perturbation bounds; d_cert[nq] counts failures.  d_qsel (optional): shortlist row i belongs
// to query d_qsel[i] (compact re-search of flagged queries).
int snk_rerank(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const float *d_val,
               const int *d_id, int KP, int k, double *d_dist, int64_t *d_idx, int64_t out_stride,
               int64_t id_offset, const float *d_qerr, const float *d_dberr, const float *d_qn,
               const float *d_maxn, const float *d_tau_extra, int *d_cert, int *d_nfail, int sticky,
               const int *d_qsel, float eps_rel, int cert_mode, cudaStream_t st);
// certificate arithmetic of snk_rerank (cert_mode): none / fp16 tensor-core keys / fp32 direct-difference keys
#define SNK_CERT_MODE_NONE 0
#define SNK_CERT_MODE_FP16 1
#define SNK_CERT_MODE_FP32 2
// relative fp32-accumulation slack of the tensor-core key for a D-column operand row (knn_tc.cu)
float snk_tc_eps_rel(const snk_db *db, int space);
int snk_tc_debug_keys(snk_db *db, int space, const __half *dQ16, int ldq16, int64_t nq, int64_t row0, int64_t nrows,
                      float *d_keys, int64_t ld, cudaStream_t st);

// exhaustive float64 search of the n queries h_qidx (host array of query indices): certificate of last resort
int snk_exact_search(snk_db *db, const snk_space &sp, const double *dQ, const int *h_qidx, int n, int k, double *d_dist,
                     int64_t *d_idx, int64_t out_stride, int64_t id_offset, cudaStream_t st);
// merge of R sorted per-shard lists (see rerank.cu)
int snk_topk_merge_launch(const double *d_dist_all, const int64_t *d_idx_all, int64_t stride_d, int64_t stride_i, int R,
                          int64_t nq, int k, double *d_dist, int64_t *d_idx, cudaStream_t st);

bool snk_merge_rerank_fits(int nlists, int lsz);
int snk_merge_rerank(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const float *d_lval,
                     const int *d_lid, int nlists, int lsz, int KP, int k, double *d_dist, int64_t *d_idx,
                     int64_t out_stride, int64_t id_offset, const float *d_qerr, const float *d_dberr,
                     const float *d_qn, const float *d_maxn, int *d_cert, int *d_nfail, int sticky, float eps_rel,
                     cudaStream_t st);

// ---- search.cu : k-NN driver shared by snk_knn and the greedy loop
// dQ float64 [nq, D] device; results device.
// d_sticky (optional, [nq] preset to 1, followed by one int failure counter): deferred certificate mode --
// uncertified queries only clear their flag / bump the counter, nothing is synchronised or re-searched;
// the caller inspects the flags later (greedy batches do, once per batch).
int snk_prepare_targets_dev(snk_db *db, const float *d_unnorm, int64_t rows, double *d_out, void *stream);
int snk_search_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist,
                   int64_t *d_idx, int64_t out_stride, int64_t id_offset, int *d_sticky, int *d_sticky_count,
                   cudaStream_t st);
// ---- greedy_one.cu : one utterance as one persistent kernel (meta: the utterance's greedy_meta, greedy_dev.cuh)
bool snk_greedy_one_supported(const snk_db *db);
// d_Jc_full != nullptr: database-sharded (rows [id_offset, id_offset + Np) here), exchange over peer memory inside the kernel
int snk_greedy_one_launch(snk_db *db, const void *meta, const double *d_targets, const float *d_unnorm, int64_t *d_paths,
                          double *d_step_dist, int *d_flags, int *d_count, float *d_keys, cudaStream_t st,
                          const float *d_Jc_full, int64_t id_offset);
// comm.cu: the peer-memory exchange regions for a kernel that runs `nsteps` exchange steps by itself (claims their epochs);
// *peers == nullptr if the communicator has no peer-memory path
int snk_comm_p2p_claim(snk_db *db, int nsteps, char *const **peers, int *rank, int *nranks, unsigned *epoch0, size_t *flags_bytes,
                       size_t *slot_bytes);
bool snk_comm_has_p2p(const snk_db *db);
// enqueue a search with deferred certificates; snk_knn_finish(db) completes it (see search.cu)
int snk_knn_enqueue(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                    int64_t out_stride, int64_t id_offset, cudaStream_t st);
