// k-NN search driver and the batched greedy joint search.
//
// snk_search_dev  : tree.query(X, k) replacement (reference script/synth_halfphone.py:1364,
//                   script/synth_simple.py:490) -- shortlist (tensor-core or SIMT) then float64 re-rank.
// greedy batch    : Synthesiser.greedy_joint_search (reference script/synth_simple.py:458-503)
//                   for B utterances at once; the chain over time steps stays sequential, the
//                   B queries of one step form one batched search over the whole database.
#include "greedy_dev.cuh"
#include <algorithm>
#include <vector>

snk_space snk_make_space(const snk_db *db, int space) {
    snk_space sp{};
    if (space == SNK_SPACE_TARGET) {
        sp.rows = db->N;
        sp.dA = 0;
        sp.dB = db->Dt;
    } else {
        sp.rows = db->Np;
        sp.dA = db->Djq;
        sp.dB = db->m * db->Dt;
    }
    sp.D = sp.dA + sp.dB;
    sp.a_row_off = db->prev_row_off;
    sp.a_col = db->prev_col;
    sp.ldA_raw = db->Dj;
    sp.ldB_raw = db->Dt;
    sp.ldA32 = db->ldJ32;
    sp.ldB32 = db->Dt;
    return sp;
}

namespace {

// f64 -> f32 query copy, optionally gathering rows through qsel
__global__ void cvt_q32_kernel(const double *__restrict__ Q, int D, int64_t nq, const int *__restrict__ qsel,
                               float *__restrict__ out) {
    const int64_t total = nq * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / D;
        const int c = (int)(i - r * D);
        const int64_t src = qsel ? qsel[r] : r;
        out[i] = (float)Q[src * D + c];
    }
}

// fp16 query operand in K-block order + squared norm of the rounded query + rounding error norm.
// One block of CVT_THREADS per (padded) query row: a row is only ~600 columns, so a warp per row would
// walk it in ~18 dependent gathers; four warps finish in five.
constexpr int CVT_THREADS = 128;
// block-wide sums of (n2, e2); thread 0 returns them.  Ends with a barrier so `red` can be reused at once.
__device__ __forceinline__ void cvt_reduce(float &n2, float &e2, float (&red)[2][CVT_THREADS / 32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        e2 += __shfl_xor_sync(0xffffffffu, e2, off);
    }
    if (lane == 0) { red[0][warp] = n2; red[1][warp] = e2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        n2 = e2 = 0.f;
        for (int w = 0; w < CVT_THREADS / 32; ++w) { n2 += red[0][w]; e2 += red[1][w]; }
    }
    __syncthreads();
}
__global__ void __launch_bounds__(CVT_THREADS)
cvt_q16_kernel(const double *__restrict__ Q, int D, int64_t nq, int64_t nq_pad, const short *__restrict__ qmap,
               int ld16, __half *__restrict__ out, float *__restrict__ qn, float *__restrict__ qerr) {
    __shared__ float red[2][CVT_THREADS / 32];
    for (int64_t q = blockIdx.x; q < nq_pad; q += gridDim.x) {
        float n2 = 0.f, e2 = 0.f;
        for (int c = threadIdx.x; c < ld16; c += CVT_THREADS) {
            __half h = __float2half_rn(0.f);
            const int d = qmap[c];
            if (q < nq && d == -2) h = __float2half_rn(-0.5f);   // multiplies the norm pieces embedded in the row
            if (q < nq && d >= 0) h = cvt_element(Q[q * D + d], n2, e2);
            out[q * ld16 + c] = h;
        }
        cvt_reduce(n2, e2, red);
        if (threadIdx.x == 0 && q < nq) {
            qn[q] = n2;
            qerr[q] = sqrtf(e2);
        }
    }
}

int shortlist_size(int k) {
    const int want = k + 8;
    if (want <= 32) return 32;
    if (want <= 64) return 64;
    if (want <= 128) return 128;
    if (want <= 256) return 256;
    return -1;
}

// relative rounding bound of an fp32 sum of D squared differences (each term: one rounded subtraction, one fma)
float simt_eps_rel(int D) { return 1.05f * (float)(D + 3) * 5.9604644775390625e-8f; }

// fp32 direct-difference shortlist + float64 re-rank.  With d_cert the answer is certified against the fp32 rounding
// of the keys (rerank.cu, SNK_CERT_MODE_FP32): uncertified queries clear their flag / bump the counter (sticky).
int search_simt(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const int *d_qsel, int k,
                double *d_dist, int64_t *d_idx, int64_t out_stride, int64_t id_offset, int *d_cert, int *d_nfail,
                cudaStream_t st) {
    const int KP = shortlist_size(k);
    SNK_CHECK(KP > 0, "k = %d too large (max 248)", k);
    SNK_TRY(snk_buf_reserve(&db->ws_q, (size_t)nq * sp.D * 4));
    SNK_TRY(snk_buf_reserve(&db->ws_list, (size_t)nq * KP * 8));
    float *q32 = (float *)db->ws_q.p;
    float *val = (float *)db->ws_list.p;
    int *id = (int *)(val + (size_t)nq * KP);
    cvt_q32_kernel<<<db->sm_count * 4, 256, 0, st>>>(dQ, sp.D, nq, d_qsel, q32);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    SNK_TRY(snk_shortlist_simt(db, sp, q32, sp.D, nq, KP, val, id, st));
    SNK_TRY(snk_rerank(db, sp, dQ, nq, val, id, KP, k, d_dist, d_idx, out_stride, id_offset, nullptr, nullptr,
                       nullptr, nullptr, nullptr, d_cert, d_nfail, 1, d_qsel, simt_eps_rel(sp.D),
                       d_cert ? SNK_CERT_MODE_FP32 : SNK_CERT_MODE_NONE, st));
    return 0;
}

}  // namespace

// A greedy step hands the search a recipe for its queries instead of finished rows (greedy_src, below): the
// tensor-core path then builds, converts and norms every row in ONE kernel; other paths assemble first.
static int greedy_launch_assemble(snk_db *db, const greedy_src *gs, int nact, double *Q, cudaStream_t st);
static int greedy_launch_assemble_cvt(snk_db *db, const greedy_src *gs, int64_t nq, int64_t qpad, int D, const short *qmap,
                                      int ld16, double *Q, __half *q16, float *qn, float *qerr, cudaStream_t st);
static int search_dev_impl(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                           int64_t out_stride, int64_t id_offset, int *d_sticky, int *d_sticky_count, cudaStream_t st,
                           const greedy_src *gs);

int snk_search_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                   int64_t out_stride, int64_t id_offset, int *d_sticky, int *d_sticky_count, cudaStream_t st) {
    return search_dev_impl(db, space, dQ, nq, k, d_dist, d_idx, out_stride, id_offset, d_sticky, d_sticky_count, st, nullptr);
}

// gs != nullptr: dQ [nq, D] is filled here (by the assemble kernels) before it is searched.
// Certificates are always DEFERRED: d_sticky [nq] is preset to 1 by the caller and cleared for every query whose
// answer could not be certified, d_sticky_count counts them; nothing here synchronises (the *_finish entry points
// look at the flags and re-search).  db->engine == SNK_ENGINE_EXACT searches exhaustively in float64 (no flags needed).
static int search_dev_impl(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                           int64_t out_stride, int64_t id_offset, int *d_sticky, int *d_sticky_count, cudaStream_t st,
                           const greedy_src *gs) {
    SNK_CHECK(db->weights_set, "snk_db_set_weights has not been called");
    SNK_CHECK(k >= 1, "k must be >= 1");
    if (nq <= 0) return 0;
    const snk_space sp = snk_make_space(db, space);
    SNK_CHECK(sp.rows > 0, "search space is empty (N=%lld, multiepoch=%d)", (long long)db->N, db->m);
    db->counters[0] += nq;
    const int KP = shortlist_size(k);
    SNK_CHECK(KP > 0, "k = %d too large (max 248)", k);
    if (db->engine == SNK_ENGINE_EXACT) {
        if (gs) SNK_TRY(greedy_launch_assemble(db, gs, (int)nq, const_cast<double *>(dQ), st));
        std::vector<int> all((size_t)nq);
        for (int64_t i = 0; i < nq; ++i) all[(size_t)i] = (int)i;
        return snk_exact_search(db, sp, dQ, all.data(), (int)nq, k, d_dist, d_idx, out_stride, id_offset, st);
    }
    const bool use_tc = db->engine != SNK_ENGINE_SIMT && snk_tc_supported(db, sp, KP);
    SNK_CHECK(use_tc || db->engine != SNK_ENGINE_TC, "tensor-core engine requested but this search shape is not supported by it");
    SNK_CHECK(d_sticky && d_sticky_count, "internal: a search needs certificate flags");
    const int64_t QB = std::max<int64_t>(4096, (int64_t)db->sm_count * 128 / 256 * 256);   // one 128-query tile per SM and batch (148 SMs: 18944)
    const bool fuse = gs && use_tc && nq <= QB;          // one batch: assemble + convert in one kernel
    if (gs && !fuse) SNK_TRY(greedy_launch_assemble(db, gs, (int)nq, const_cast<double *>(dQ), st));
    if (!use_tc)
        return search_simt(db, sp, dQ, nq, nullptr, k, d_dist, d_idx, out_stride, id_offset, d_sticky, d_sticky_count, st);

    // ---- tensor-core shortlist + float64 re-rank + certificate, in query batches
    const int ld16 = snk_tc_query_ld(db, space);
    const float eps_rel = snk_tc_eps_rel(db, space);
    for (int64_t qb = 0; qb < nq; qb += QB) {
        const int64_t qn_ = std::min(QB, nq - qb);
        const int64_t qpad = snk_round_up(qn_, 256);   // two query tiles: the tensor-core kernel may pair CTAs
        // ws_io2 layout: Q16 [qpad, ld16] | qn [qpad] | qerr [qpad] | tau [qpad]
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        const size_t o_q16 = take((size_t)qpad * ld16 * 2), o_qn = take(qpad * 4), o_qe = take(qpad * 4),
                     o_tau = take(qpad * 4);
        SNK_TRY(snk_buf_reserve(&db->ws_io2, off));
        char *base = (char *)db->ws_io2.p;
        __half *q16 = (__half *)(base + o_q16);
        float *qn = (float *)(base + o_qn), *qerr = (float *)(base + o_qe), *tau = (float *)(base + o_tau);
        SNK_TRY(snk_buf_reserve(&db->ws_list, (size_t)qn_ * KP * 8));
        float *val = (float *)db->ws_list.p;
        int *id = (int *)(val + (size_t)qn_ * KP);
        const double *Qb = dQ + qb * sp.D;

        if (fuse) {
            SNK_TRY(greedy_launch_assemble_cvt(db, gs, qn_, qpad, sp.D, snk_tc_qmap(db, space), ld16, const_cast<double *>(Qb),
                                               q16, qn, qerr, st));
        } else {
            cvt_q16_kernel<<<(unsigned)std::min<int64_t>(qpad, (int64_t)db->sm_count * 16), CVT_THREADS, 0, st>>>(
                Qb, sp.D, qn_, qpad, snk_tc_qmap(db, space), ld16, q16, qn, qerr);
            SNK_CUDA(cudaGetLastError());
            db->counters[2] += 1;
        }
        snk_tc_lists lists;
        SNK_TRY(snk_shortlist_tc(db, space, q16, ld16, qn_, k, KP, val, id, tau, &lists, st));
        const bool joint = space == SNK_SPACE_JOINT;
        // fused merge + re-rank: 16 exact distances are plenty for k <= 2 (the certificate still guards it)
        if (lists.valid)
            SNK_TRY(snk_merge_rerank(db, sp, Qb, qn_, lists.val, lists.id, lists.nlists, lists.lsz, k <= 2 ? 16 : KP, k,
                                     d_dist + qb * out_stride, d_idx + qb * out_stride, out_stride, id_offset, qerr,
                                     joint ? db->err_j16 : db->err_t16, qn, joint ? db->maxn_j16 : db->maxn_t16,
                                     d_sticky + qb, d_sticky_count, 1, eps_rel, st));
        else
            SNK_TRY(snk_rerank(db, sp, Qb, qn_, val, id, KP, k, d_dist + qb * out_stride, d_idx + qb * out_stride,
                               out_stride, id_offset, qerr, joint ? db->err_j16 : db->err_t16, qn,
                               joint ? db->maxn_j16 : db->maxn_t16, tau, d_sticky + qb, d_sticky_count, 1, nullptr,
                               eps_rel, SNK_CERT_MODE_FP16, st));
    }
    return 0;
}

// Test instrumentation (include/snk_b200.h: snk_debug_tc_keys): the tensor-core kernel's raw keys, the fp32 squared norm
// of the rounded query the certificate adds to them, and the slack the certificate allows.
extern "C" int snk_debug_tc_keys(snk_db *db, int space, const double *Q, int64_t nq, int64_t row0, int64_t nrows, float *keys,
                                 float *qnorm, float *eps_rel, float *maxnorm) {
    SNK_CHECK(db && Q && keys && db->weights_set, "NULL argument / weights not set");
    SNK_LOCK(db);
    SNK_CHECK(space == SNK_SPACE_TARGET || space == SNK_SPACE_JOINT, "unknown search space %d", space);
    SNK_CUDA(cudaSetDevice(db->device));
    const snk_space sp = snk_make_space(db, space);
    SNK_CHECK(snk_tc_supported(db, sp, 32), "tensor-core engine does not support this search space");
    SNK_CHECK(nq >= 1 && nq <= 4096 && row0 >= 0 && nrows >= 1 && row0 + nrows <= sp.rows, "bad query / row range");
    const int ld16 = snk_tc_query_ld(db, space);
    const int64_t qpad = snk_round_up(nq, 256), ld = snk_round_up(nrows, 128);
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)nq * sp.D * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_io2, (size_t)qpad * ld16 * 2 + (size_t)qpad * 8 + 512));
    SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)qpad * ld * 4));
    __half *q16 = (__half *)db->ws_io2.p;
    float *qn = (float *)((char *)db->ws_io2.p + snk_round_up((size_t)qpad * ld16 * 2, 256)), *qerr = qn + qpad;
    cudaStream_t st = db->stream;
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, Q, (size_t)nq * sp.D * 8, cudaMemcpyHostToDevice, st));
    cvt_q16_kernel<<<(unsigned)std::min<int64_t>(qpad, (int64_t)db->sm_count * 16), CVT_THREADS, 0, st>>>(
        (const double *)db->ws_h0.p, sp.D, nq, qpad, snk_tc_qmap(db, space), ld16, q16, qn, qerr);
    SNK_CUDA(cudaGetLastError());
    SNK_TRY(snk_tc_debug_keys(db, space, q16, ld16, nq, row0, nrows, (float *)db->ws_dist.p, ld, st));
    SNK_CUDA(cudaMemcpy2DAsync(keys, (size_t)nrows * 4, db->ws_dist.p, (size_t)ld * 4, (size_t)nrows * 4, (size_t)nq,
                               cudaMemcpyDeviceToHost, st));
    if (qnorm) SNK_CUDA(cudaMemcpyAsync(qnorm, qn, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (maxnorm)
        SNK_CUDA(cudaMemcpyAsync(maxnorm, space == SNK_SPACE_JOINT ? db->maxn_j16 : db->maxn_t16, 4, cudaMemcpyDeviceToHost, st));
    SNK_CUDA(cudaStreamSynchronize(st));
    if (eps_rel) *eps_rel = snk_tc_eps_rel(db, space);
    return 0;
}

// =====================================================================================
// greedy joint search
namespace {

__global__ void prepare_targets_kernel(const float *__restrict__ x, int64_t total, int Dt, std_params sp,
                                       double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = standardise_weight(x[i], (int)(i % Dt), sp);
}

// Half-phone targets (train_halfphone.py:959-1070, synth_halfphone.py:1510-1548): unit i is described by P frames of the
// utterance (first / middle / last frame of the half-phone, chosen from the state alignment), optionally followed by its
// normalised duration; the frames are standardised, the row is weighted.  Column p * dim + c of the row takes the
// statistics and the weight of target column p * dim + c, so the Dt-wide vectors of snk_db_set_standardisation hold the
// frame statistics repeated per point.  The duration column is not standardised here (get_norm_durations did that).
__global__ void halfphone_targets_kernel(const float *__restrict__ x, int64_t frames, int dim, const int64_t *__restrict__ pts,
                                         int64_t n, int P, const double *__restrict__ dur, int Dt, std_params sp,
                                         double *__restrict__ out) {
    const int64_t total = n * Dt;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / Dt;
        const int col = (int)(i - u * Dt);
        if (col >= P * dim) { out[i] = __dmul_rn(dur[u], sp.w[col]); continue; }
        int64_t f = pts[u * P + col / dim];
        if (f < 0) f += frames;                      // numpy's negative indexing
        out[i] = (f >= 0 && f < frames) ? standardise_weight(x[f * dim + col % dim], col, sp) : NAN;
    }
}

// float64 query rows only (SIMT engine, table-free paths, the last scatter-only step)
__global__ void greedy_assemble_kernel(const greedy_src g, int nact, double *__restrict__ Q) {
    const int b = blockIdx.x;
    const greedy_meta mt = g.meta[b];
    const int D = g.Djq + g.m * g.Dt;
    if (threadIdx.x == 0) greedy_scatter(g, mt, b);
    if (b >= nact) return;
    int64_t row;
    int col;
    greedy_prev(g, mt, b, row, col);
    double *q = Q + (int64_t)b * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) q[d] = greedy_value(g, mt, row, col, d);
}

// The same plus what cvt_q16_kernel does, in one pass over the operand columns: float64 row (for the re-rank),
// fp16 operand row in K-block order, squared norm of the rounded row and rounding-error norm.
__global__ void __launch_bounds__(CVT_THREADS)
greedy_assemble_cvt_kernel(const greedy_src g, int64_t nq, int64_t qpad, int D, const short *__restrict__ qmap, int ld16,
                           double *__restrict__ Q, __half *__restrict__ out, float *__restrict__ qn,
                           float *__restrict__ qerr) {
    __shared__ float red[2][CVT_THREADS / 32];
    const int64_t nrows = max(qpad, (int64_t)g.nact_prev);
    for (int64_t q = blockIdx.x; q < nrows; q += gridDim.x) {
        const bool live = q < nq || q < g.nact_prev;
        greedy_meta mt{};
        if (live) mt = g.meta[q];
        if (threadIdx.x == 0 && live) greedy_scatter(g, mt, q);
        if (q >= qpad) continue;                      // finished utterances beyond the padded batch: scatter only
        int64_t row = -1;
        int col = 0;
        if (q < nq) greedy_prev(g, mt, q, row, col);
        float n2 = 0.f, e2 = 0.f;
        for (int c = threadIdx.x; c < ld16; c += CVT_THREADS) {
            __half h = __float2half_rn(0.f);
            const int d = qmap[c];
            if (q < nq && d == -2) h = __float2half_rn(-0.5f);   // multiplies the norm pieces embedded in the row
            if (q < nq && d >= 0) {
                const double x = greedy_value(g, mt, row, col, d);
                Q[q * D + d] = x;
                h = cvt_element(x, n2, e2);
            }
            out[q * ld16 + c] = h;
        }
        cvt_reduce(n2, e2, red);
        if (threadIdx.x == 0 && q < nq) {
            qn[q] = n2;
            qerr[q] = sqrtf(e2);
        }
    }
}

}  // namespace

static int greedy_launch_assemble(snk_db *db, const greedy_src *gs, int nact, double *Q, cudaStream_t st) {
    const int grid = std::max(nact, gs->nact_prev);
    if (grid == 0) return 0;
    greedy_assemble_kernel<<<grid, 128, 0, st>>>(*gs, nact, Q);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

static int greedy_launch_assemble_cvt(snk_db *db, const greedy_src *gs, int64_t nq, int64_t qpad, int D, const short *qmap,
                                      int ld16, double *Q, __half *q16, float *qn, float *qerr, cudaStream_t st) {
    const int64_t nrows = std::max<int64_t>(qpad, gs->nact_prev);
    greedy_assemble_cvt_kernel<<<(unsigned)std::min<int64_t>(nrows, (int64_t)db->sm_count * 16), CVT_THREADS, 0, st>>>(
        *gs, nq, qpad, D, qmap, ld16, Q, q16, qn, qerr);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

namespace {

// One pass of the greedy chain over the utterances in `meta` (sorted longest first).  No step synchronises: the
// certificate flags [n] are preset to 1 and cleared by any step whose answer could not be certified.
// sh (database-sharded search): this handle holds a block of the joint rows; row ids become global
// by sh->id_offset, after every step the ranks exchange their best (distance, row) pairs, and the previous join vector is
// read from the replicated join contexts sh->Jc_full (current_join_rep[u] = Jw[u + m], synth_simple.py:213-214,501).
struct greedy_shard { const float *Jc_full; int64_t id_offset; };

int greedy_run(snk_db *db, const std::vector<greedy_meta> &meta, const double *d_targets, const float *d_unnorm,
               int64_t *d_paths, double *d_step_dist, int *d_flags, int *d_count, cudaStream_t st,
               const greedy_shard *sh = nullptr) {
    const int B = (int)meta.size();
    if (B == 1 && snk_greedy_one_supported(db) && (!sh || snk_comm_has_p2p(db) || snk_comm_nranks(db) == 1)) {
        // one utterance: the whole chain is one persistent kernel (greedy_one.cu), over a sharded database with the per-step
        // exchange through peer memory inside it; all its frames must have landed
        for (auto &w : db->step_waits) SNK_CUDA(cudaStreamWaitEvent(st, w.second, 0));
        return snk_greedy_one_launch(db, &meta[0], d_targets, d_unnorm, d_paths, d_step_dist, d_flags, d_count, nullptr, st,
                                     sh ? sh->Jc_full : nullptr, sh ? sh->id_offset : 0);
    }
    const std_params stp{db->std_mean, db->std_sd, db->wt, db->uv_special, db->uv_scale, db->std_f32};
    const int m = db->m;
    const snk_space sp = snk_make_space(db, SNK_SPACE_JOINT);
    int64_t maxsteps = 0;
    for (const greedy_meta &g : meta) maxsteps = std::max(maxsteps, g.nsteps);
    // ws_io: meta | Q [B, D] | ix [B] | dist [B] | bound [B]
    const size_t meta_bytes = snk_round_up(sizeof(greedy_meta) * B, 256);
    const size_t q_bytes = snk_round_up((size_t)B * sp.D * 8, 256);
    SNK_TRY(snk_buf_reserve(&db->ws_io, meta_bytes + q_bytes + (size_t)B * 24 + 1024));
    char *base = (char *)db->ws_io.p;
    greedy_meta *d_meta = (greedy_meta *)base;
    double *Q = (double *)(base + meta_bytes);
    int64_t *ix = (int64_t *)(base + meta_bytes + q_bytes);
    double *dist = (double *)(base + meta_bytes + q_bytes + snk_round_up((size_t)B * 8, 256));
    double *bound = (double *)(base + meta_bytes + q_bytes + 2 * snk_round_up((size_t)B * 8, 256));   // sharded search only
    SNK_TRY(snk_upload_async(db, d_meta, meta.data(), sizeof(greedy_meta) * B, st));   // pinned staging: no sync
    int nact_prev = 0;
    size_t next_wait = 0;
    for (int64_t t = 0; t <= maxsteps; ++t) {
        int nact = 0;
        while (nact < B && meta[nact].nsteps > t) ++nact;
        if (std::max(nact, nact_prev) == 0) break;
        // chunked uploads (host entry point): step t may start once its target frames have landed
        while (next_wait < db->step_waits.size() && db->step_waits[next_wait].first <= t)
            SNK_CUDA(cudaStreamWaitEvent(st, db->step_waits[next_wait++].second, 0));
        const greedy_src gs{d_meta, nact_prev, t, d_targets, d_unnorm, stp, db->Dt, m, sh ? sh->Jc_full : db->Jc_raw, db->wj, db->Dj, db->Djq,
                            db->prev_row_off, db->prev_col, db->cur_row_off, db->cur_col, ix, dist, d_paths, d_step_dist};
        if (nact > 0) {   // the search builds its queries from the recipe (one fused kernel on the tensor-core path)
            db->cert_bound_out = sh ? bound : nullptr;
            const int rcs = search_dev_impl(db, SNK_SPACE_JOINT, Q, nact, 1, dist, ix, 1, sh ? sh->id_offset : 0, d_flags, d_count, st, &gs);
            db->cert_bound_out = nullptr;
            SNK_TRY(rcs);
            if (sh) SNK_TRY(snk_comm_exchange_best(db, dist, ix, db->engine == SNK_ENGINE_EXACT ? nullptr : bound, nact, d_flags, d_count, st));
        } else {          // every utterance has finished: only the last choices remain to be written out
            SNK_TRY(greedy_launch_assemble(db, &gs, 0, Q, st));
        }
        nact_prev = nact;
    }
    return 0;
}

__global__ void fill_int_kernel(int *p, int64_t n, int v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// flags [n] preset to 1 followed by one int counter preset to 0
int launch_flag_reset(snk_db *db, int *flags, int64_t n, cudaStream_t st) {
    fill_int_kernel<<<64, 256, 0, st>>>(flags, n, 1);
    SNK_CUDA(cudaGetLastError());
    SNK_CUDA(cudaMemsetAsync(flags + n, 0, 4, st));
    db->counters[2] += 1;
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Deferred certificates.  A `_dev` search leaves one flag per query (or per utterance) and a failure counter on the
// device and records here what it would take to repair the answer; snk_knn_finish / snk_greedy_batch_finish wait for
// the stream, read the counter and, if it is non-zero, repeat the flagged work with the next engine of the chain
//     tensor-core fp16 shortlist  ->  fp32 direct-difference shortlist  ->  exhaustive float64 scan,
// each stage certified against its own rounding, the last one exact by construction.
struct snk_pending_state {
    struct flagbuf { int *p; size_t cap; };
    std::vector<flagbuf> pool;          // idle flag buffers
    struct knn_job {
        int space, k;
        const double *dQ;
        int64_t nq, out_stride, id_offset;
        double *d_dist;
        int64_t *d_idx;
        cudaStream_t st;
        flagbuf fb;
    };
    struct greedy_job {
        std::vector<greedy_meta> meta;
        const double *d_targets;
        const float *d_unnorm;
        int64_t *d_paths;
        double *d_step_dist;
        cudaStream_t st;
        flagbuf fb;
        bool sharded = false;
        const float *Jc_full = nullptr;
        int64_t id_offset = 0;
    };
    std::vector<knn_job> knn;
    std::vector<greedy_job> greedy;
};

static snk_pending_state *pending_of(snk_db *db) {
    if (!db->pending) db->pending = new snk_pending_state();
    return db->pending;
}

void snk_pending_destroy(snk_db *db) {
    if (!db->pending) return;
    for (auto &f : db->pending->pool) cudaFree(f.p);
    for (auto &j : db->pending->knn) cudaFree(j.fb.p);
    for (auto &j : db->pending->greedy) cudaFree(j.fb.p);
    delete db->pending;
    db->pending = nullptr;
}

static int take_flags(snk_db *db, int64_t n, snk_pending_state::flagbuf *out) {
    snk_pending_state *ps = pending_of(db);
    const size_t need = (size_t)(n + 1) * 4;
    for (size_t i = 0; i < ps->pool.size(); ++i)
        if (ps->pool[i].cap >= need) {
            *out = ps->pool[i];
            ps->pool.erase(ps->pool.begin() + i);
            return 0;
        }
    out->cap = need + need / 2 + 256;
    SNK_CUDA(cudaMalloc((void **)&out->p, out->cap));
    return 0;
}

int snk_knn_enqueue(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                    int64_t out_stride, int64_t id_offset, cudaStream_t st) {
    if (nq <= 0) return 0;
    snk_pending_state::knn_job job{space, k, dQ, nq, out_stride, id_offset, d_dist, d_idx, st, {nullptr, 0}};
    SNK_TRY(take_flags(db, nq, &job.fb));
    pending_of(db)->knn.push_back(job);
    SNK_TRY(launch_flag_reset(db, job.fb.p, nq, st));
    return search_dev_impl(db, space, dQ, nq, k, d_dist, d_idx, out_stride, id_offset, job.fb.p, job.fb.p + nq, st, nullptr);
}

// reads the failure counter of a finished job (the stream must have been synchronised)
static int read_count(const int *d_count, int *out) {
    SNK_CUDA(cudaMemcpy(out, d_count, 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int failed_indices(const int *d_flags, int64_t n, std::vector<int> *out) {
    std::vector<int> h((size_t)n);
    SNK_CUDA(cudaMemcpy(h.data(), d_flags, (size_t)n * 4, cudaMemcpyDeviceToHost));
    out->clear();
    for (int64_t i = 0; i < n; ++i)
        if (!h[(size_t)i]) out->push_back((int)i);
    return 0;
}

extern "C" int snk_knn_finish(snk_db *db) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    if (!db->pending || db->pending->knn.empty()) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    snk_pending_state *ps = db->pending;
    int rc = 0;
    for (auto &job : ps->knn) {
        if (!rc) rc = cudaStreamSynchronize(job.st) == cudaSuccess ? 0 : 1;
        int nfail = 0;
        if (!rc) rc = read_count(job.fb.p + job.nq, &nfail);
        const snk_space sp = snk_make_space(db, job.space);
        if (!rc && nfail > 0) {
            // stage 2: fp32 direct differences for the flagged queries, certified against fp32 rounding
            db->counters[1] += nfail;
            std::vector<int> idx;
            rc = failed_indices(job.fb.p, job.nq, &idx);
            if (!rc) rc = snk_buf_reserve(&db->ws_kflags, idx.size() * 4);
            int *qsel = (int *)db->ws_kflags.p;
            if (!rc) rc = snk_upload_async(db, qsel, idx.data(), idx.size() * 4, job.st);
            if (!rc) rc = launch_flag_reset(db, job.fb.p, job.nq, job.st);
            if (!rc) rc = search_simt(db, sp, job.dQ, (int64_t)idx.size(), qsel, job.k, job.d_dist, job.d_idx, job.out_stride,
                                      job.id_offset, job.fb.p, job.fb.p + job.nq, job.st);
            if (!rc) rc = cudaStreamSynchronize(job.st) == cudaSuccess ? 0 : 1;
            int nfail2 = 0;
            if (!rc) rc = read_count(job.fb.p + job.nq, &nfail2);
            if (!rc && nfail2 > 0) {
                // stage 3: exhaustive float64 scan
                db->counters[3] += nfail2;
                rc = failed_indices(job.fb.p, job.nq, &idx);
                if (!rc) rc = snk_exact_search(db, sp, job.dQ, idx.data(), (int)idx.size(), job.k, job.d_dist, job.d_idx,
                                               job.out_stride, job.id_offset, job.st);
                if (!rc) rc = cudaStreamSynchronize(job.st) == cudaSuccess ? 0 : 1;
            }
        }
        ps->pool.push_back(job.fb);
    }
    ps->knn.clear();
    if (rc) snk_set_error("snk_knn_finish failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

static int greedy_batch_core(snk_db *db, const double *d_targets, const float *d_unnorm, const int64_t *lens, int B,
                             const int64_t *start_state, int64_t *d_paths, double *d_step_dist, void *stream,
                             const float *d_Jc_full = nullptr, int64_t rows_full = 0, int64_t id_offset = 0);

int snk_greedy_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B,
                         const int64_t *start_state, int64_t *d_paths, double *d_step_dist, void *stream) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    return greedy_batch_core(db, d_targets, nullptr, lens, B, start_state, d_paths, d_step_dist, stream);
}

int snk_greedy_batch_unnorm_dev(snk_db *db, const float *d_unnorm, const int64_t *lens, int B,
                                const int64_t *start_state, int64_t *d_paths, double *d_step_dist, void *stream) {
    SNK_CHECK(db && db->std_set, "snk_db_set_standardisation has not been called");
    SNK_LOCK(db);
    return greedy_batch_core(db, nullptr, d_unnorm, lens, B, start_state, d_paths, d_step_dist, stream);
}

// Greedy search over a database whose joint rows are sharded by row block across the ranks of the communicator
// (snk_comm_init): this handle holds joint rows [id_offset, id_offset + Np) -- frames F[id_offset : id_offset + Np + m - 1],
// join contexts Jc[id_offset : id_offset + Np + m] -- and d_Jc_full is the replicated un-weighted join matrix
// [rows_full + m, Dj] every rank reads the previous join vector from.  One exchange per time step inside the library
// (grouped ncclAllGather of B * 24 bytes + arg-min / certificate kernel); paths hold GLOBAL row ids, identical on every rank.
// Collective: every rank calls it with the same targets, and snk_greedy_batch_finish afterwards.
int snk_greedy_sharded_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B, const int64_t *start_state,
                                 const float *d_Jc_full, int64_t rows_full, int64_t id_offset, int64_t *d_paths,
                                 double *d_step_dist, void *stream) {
    SNK_CHECK(db && d_Jc_full, "NULL argument");
    SNK_LOCK(db);
    SNK_CHECK(db->comm, "snk_comm_init has not been called");
    SNK_CHECK(id_offset >= 0 && id_offset + db->Np <= rows_full, "shard [%lld, %lld) outside the %lld joint rows",
              (long long)id_offset, (long long)(id_offset + db->Np), (long long)rows_full);
    return greedy_batch_core(db, d_targets, nullptr, lens, B, start_state, d_paths, d_step_dist, stream, d_Jc_full, rows_full,
                             id_offset);
}

int snk_prepare_targets_dev(snk_db *db, const float *d_unnorm, int64_t rows, double *d_out, void *stream) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_CHECK(db->std_set, "snk_db_set_standardisation has not been called");
    if (rows <= 0) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    const std_params stp{db->std_mean, db->std_sd, db->wt, db->uv_special, db->uv_scale, db->std_f32};
    const int64_t total = rows * db->Dt;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)db->sm_count * 8);
    prepare_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_unnorm, total, db->Dt, stp, d_out);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

int snk_halfphone_targets_dev(snk_db *db, const float *d_unnorm, int64_t frames, int dim, const int64_t *d_points, int64_t n,
                              int P, const double *d_dur, double *d_out, void *stream) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_CHECK(db->std_set, "snk_db_set_standardisation has not been called");
    SNK_CHECK(P >= 1 && P <= 3 && dim >= 1 && frames >= 1, "bad shape");
    SNK_CHECK(P * dim + (d_dur ? 1 : 0) == db->Dt, "%d points x %d dims%s do not make the %d target columns of this voice", P,
              dim, d_dur ? " + duration" : "", db->Dt);
    if (n <= 0) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    const std_params stp{db->std_mean, db->std_sd, db->wt, db->uv_special, db->uv_scale, db->std_f32};
    const int64_t total = n * db->Dt;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)db->sm_count * 8);
    halfphone_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_unnorm, frames, dim, d_points, n, P, d_dur, db->Dt, stp, d_out);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

static int greedy_batch_core(snk_db *db, const double *d_targets, const float *d_unnorm, const int64_t *lens, int B,
                             const int64_t *start_state, int64_t *d_paths, double *d_step_dist, void *stream,
                             const float *d_Jc_full, int64_t rows_full, int64_t id_offset) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    const bool sharded = d_Jc_full != nullptr;
    const int64_t rows_all = sharded ? rows_full : db->Np;
    SNK_CUDA(cudaSetDevice(db->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0) return 0;
    const int m = db->m;
    // utterance order: longest first so the active set of every step is a prefix
    snk_pending_state::greedy_job job;
    std::vector<greedy_meta> &meta = job.meta;
    meta.resize(B);
    int64_t toff = 0, poff = 0;
    for (int b = 0; b < B; ++b) {
        SNK_CHECK(lens[b] >= m, "utterance %d has %lld frames, fewer than multiepoch=%d "
                  "(the reference's segment_axis raises ValueError here)", b, (long long)lens[b], m);
        meta[b].tgt_off = toff;
        meta[b].path_off = poff;
        meta[b].nsteps = lens[b] / m;
        meta[b].start_state = start_state ? start_state[b] : -1;
        if (meta[b].start_state >= rows_all) {
            snk_set_error("start_state %lld out of range", (long long)meta[b].start_state);
            return 1;
        }
        toff += lens[b];
        poff += meta[b].nsteps;
    }
    std::stable_sort(meta.begin(), meta.end(),
                     [](const greedy_meta &a, const greedy_meta &b) { return a.nsteps > b.nsteps; });
    // deferred certificates: flags [B] preset to 1 + a failure counter, inspected by snk_greedy_batch_finish
    SNK_TRY(take_flags(db, B, &job.fb));
    job.d_targets = d_targets; job.d_unnorm = d_unnorm; job.d_paths = d_paths; job.d_step_dist = d_step_dist; job.st = st;
    job.sharded = sharded; job.Jc_full = d_Jc_full; job.id_offset = id_offset;
    const greedy_shard sh{d_Jc_full, id_offset};
    int *flags = job.fb.p, *count = flags + B;
    int rc = launch_flag_reset(db, flags, B, st);
    if (!rc) rc = greedy_run(db, meta, d_targets, d_unnorm, d_paths, d_step_dist, flags, count, st, sharded ? &sh : nullptr);
    pending_of(db)->greedy.push_back(std::move(job));
    return rc;
}

extern "C" int snk_greedy_batch_finish(snk_db *db) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    if (!db->pending || db->pending->greedy.empty()) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    snk_pending_state *ps = db->pending;
    int rc = 0;
    for (auto &job : ps->greedy) {
        const int B = (int)job.meta.size();
        // some step of some utterance was not certified: redo those utterances with the next engine (their later
        // steps depend on the doubtful choice, so the whole chain is repeated)
        const int chain[2] = {SNK_ENGINE_SIMT, SNK_ENGINE_EXACT};
        for (int stage = 0; stage < 2 && !rc; ++stage) {
            const int nflags = (int)job.meta.size();     // flags index the list of the run they belong to
            // sharded: an utterance is doubtful if ANY rank could not certify one of its steps -- all ranks agree on the
            // union of the cleared flags (element-wise minimum) and redo the same utterances in lockstep
            if (job.sharded) rc = snk_comm_allreduce_min(db, job.fb.p, nflags, job.st);
            if (!rc) rc = cudaStreamSynchronize(job.st) == cudaSuccess ? 0 : 1;
            std::vector<int> idx;
            if (!rc) rc = failed_indices(job.fb.p, nflags, &idx);
            if (rc || idx.empty()) break;
            std::vector<greedy_meta> redo;
            for (int b : idx) redo.push_back(job.meta[(size_t)b]);
            db->counters[stage == 0 ? 1 : 3] += (int64_t)redo.size();
            rc = launch_flag_reset(db, job.fb.p, B, job.st);
            const int saved = db->engine;
            db->engine = chain[stage];
            const greedy_shard sh{job.Jc_full, job.id_offset};
            if (!rc) rc = greedy_run(db, redo, job.d_targets, job.d_unnorm, job.d_paths, job.d_step_dist, job.fb.p, job.fb.p + B,
                                     job.st, job.sharded ? &sh : nullptr);
            db->engine = saved;
            job.meta.swap(redo);   // flags of the redo run index the redo list
            if (!rc) rc = cudaStreamSynchronize(job.st) == cudaSuccess ? 0 : 1;
        }
        ps->pool.push_back(job.fb);
    }
    ps->greedy.clear();
    return rc;
}
