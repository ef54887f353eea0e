// C ABI of the engine (include/snk_b200.h): handle lifetime and the host-pointer entry points.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

int snk_join_tiles_dev(snk_db *db, const int64_t *d_cand, const int64_t *lens, int B, int K, float *d_tiles,
                       cudaStream_t st);
int snk_candidate_distances_dev(snk_db *db, const int64_t *d_cand, const double *d_targets, int64_t T, int K,
                                double *d_dist, cudaStream_t st);
int snk_path_scores_dev(snk_db *db, const double *d_targets, int64_t T, const int64_t *d_path, int64_t P,
                        const int *d_tw, int nts, const int *d_jw, int njs, double *d_ts, double *d_js,
                        cudaStream_t st);

// ---- acoustic preselection + join + Viterbi in one call: candidates never leave the device -------------------
struct snk_acoustic_job {
    bool active = false;
    std::vector<int64_t> lens;
    int K = 0;
    unsigned flags = 0;
    int64_t *d_paths = nullptr, *d_plen = nullptr;
    double *d_pcost = nullptr, *d_tcost = nullptr, *d_jcost = nullptr;
    cudaStream_t st = nullptr;
};
static thread_local char g_err[1024] = "";

void snk_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int snk_buf_reserve(snk_buf *b, size_t bytes) {
    if (bytes <= b->cap) return 0;
    if (b->p) {
        SNK_CUDA(cudaFree(b->p));
        b->p = nullptr;
        b->cap = 0;
    }
    const size_t want = bytes + bytes / 4 + 256;
    SNK_CUDA(cudaMalloc(&b->p, want));
    b->cap = want;
    return 0;
}
void snk_buf_free(snk_buf *b) {
    if (b->p) cudaFree(b->p);
    b->p = nullptr;
    b->cap = 0;
}

int snk_upload_async(snk_db *db, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    snk_stage_slot &sl = db->stage[db->stage_next];
    db->stage_next = (db->stage_next + 1) % SNK_STAGE_SLOTS;
    if (!sl.ev) SNK_CUDA(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    else SNK_CUDA(cudaEventSynchronize(sl.ev));          // the copy that last used this slot (8 uploads ago) has run
    if (bytes > sl.cap) {
        if (sl.host) SNK_CUDA(cudaFreeHost(sl.host));
        sl.host = nullptr;
        sl.cap = 0;
        const size_t want = bytes + bytes / 2 + 4096;
        SNK_CUDA(cudaMallocHost(&sl.host, want));
        sl.cap = want;
    }
    memcpy(sl.host, h_src, bytes);
    SNK_CUDA(cudaMemcpyAsync(d_dst, sl.host, bytes, cudaMemcpyHostToDevice, st));
    SNK_CUDA(cudaEventRecord(sl.ev, st));
    return 0;
}

extern "C" {

const char *snk_last_error(void) { return g_err; }
int snk_version(void) { return 100; }

int snk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int snk_db_create(snk_db **out, int device_id, int64_t N, int Dt, int Dj, int multiepoch, const float *F,
                  const float *Jc, unsigned layout_flags) {
    SNK_CHECK(out, "out is NULL");
    *out = nullptr;
    SNK_CHECK(F && Jc, "F / Jc is NULL");
    SNK_CHECK(N >= 1 && Dt >= 1 && Dj >= 1, "bad database shape N=%lld Dt=%d Dj=%d", (long long)N, Dt, Dj);
    SNK_CHECK(N < (int64_t)2000000000, "N too large for 32-bit row ids");
    SNK_CHECK(multiepoch >= 1, "multiepoch must be >= 1");
    SNK_CHECK(layout_flags == SNK_LAYOUT_SIMPLE || layout_flags == SNK_LAYOUT_HALFPHONE_EPOCH, "unknown layout flag %u",
              layout_flags);
    SNK_CHECK(snk_device_count() > 0, "no CUDA device visible: this engine has no CPU fallback");
    SNK_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    SNK_CUDA(cudaGetDeviceProperties(&prop, device_id));
    SNK_CHECK(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device_id,
              prop.major, prop.minor);
    snk_db *db = new snk_db();
    db->device = device_id;
    if (const char *e = getenv("SNK_DEBUG_CERT_FAIL")) db->debug_fail_mod = atoi(e);
    if (const char *e = getenv("SNK_DEBUG_CERT_FAIL2")) db->debug_fail_mod2 = atoi(e);
    db->sm_count = prop.multiProcessorCount;
    db->N = N; db->Dt = Dt; db->Dj = Dj; db->m = multiepoch; db->layout = layout_flags;
    db->Np = N - (multiepoch - 1);
    if (db->Np < 0) db->Np = 0;
    if (layout_flags == SNK_LAYOUT_SIMPLE) {
        // prev_join_rep[u] = start[u] = Jw[u];  current_join_rep[u] = end[u + m - 1] = Jw[u + m]
        db->Djq = Dj; db->prev_row_off = 0; db->prev_col = 0; db->cur_row_off = multiepoch; db->cur_col = 0;
    } else {
        // prev = start[u][:Dj/2] = Jw[u][:h];  current = start[u + m - 1][h:]
        SNK_CHECK(Dj % 2 == 0, "halfphone-epoch layout needs an even join dimension");
        db->Djq = Dj / 2; db->prev_row_off = 0; db->prev_col = 0; db->cur_row_off = multiepoch - 1; db->cur_col = Dj / 2;
    }
    db->ldJ32 = (int)snk_round_up(Dj, 32);   // rows start on 128-byte lines: a 32-dim group of four rows is four lines, not seven (join_tc.cu)
    db->ldG16 = (int)snk_round_up(Dt, 64);
    db->ldS16 = (int)snk_round_up(db->Djq, 64);
#define ALLOC(ptr, bytes)                                                                      \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                                 \
        snk_set_error("cudaMalloc of %zu bytes failed: %s", (size_t)(bytes), cudaGetErrorString(cudaGetLastError())); \
        snk_db_destroy(db);                                                                    \
        return 1;                                                                              \
    }
    ALLOC(db->F_raw, (size_t)N * Dt * 4);
    ALLOC(db->Jc_raw, (size_t)(N + 1) * Dj * 4);
    ALLOC(db->wt, (size_t)Dt * 8);
    ALLOC(db->wj, (size_t)Dj * 8);
    ALLOC(db->Fw32, (size_t)N * Dt * 4);
    ALLOC(db->Jw32, (size_t)(N + 2) * db->ldJ32 * 4);      // row N + 1 stays zero: the row join_tc.cu loads for absent candidates
    cudaMemset(db->Jw32 + (size_t)(N + 1) * db->ldJ32, 0, (size_t)db->ldJ32 * 4);
    ALLOC(db->G16, (size_t)N * db->ldG16 * 2);
    ALLOC(db->S16, (size_t)(N + 1) * db->ldS16 * 2);
    ALLOC(db->nrm_t16, (size_t)N * 4);
    ALLOC(db->nrm_j16, (size_t)(db->Np > 0 ? db->Np : 1) * 4);
    ALLOC(db->err_t16, 256);
    db->err_j16 = db->err_t16 + 1;
    db->maxn_t16 = db->err_t16 + 2;
    db->maxn_j16 = db->err_t16 + 3;
#undef ALLOC
    if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&db->ev) != cudaSuccess) {
        snk_set_error("stream/event creation failed");
        snk_db_destroy(db);
        return 1;
    }
    if (cudaMemcpyAsync(db->F_raw, F, (size_t)N * Dt * 4, cudaMemcpyHostToDevice, db->stream) != cudaSuccess ||
        cudaMemcpyAsync(db->Jc_raw, Jc, (size_t)(N + 1) * Dj * 4, cudaMemcpyHostToDevice, db->stream) != cudaSuccess ||
        cudaStreamSynchronize(db->stream) != cudaSuccess) {
        snk_set_error("database upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        snk_db_destroy(db);
        return 1;
    }
    if (snk_tc_prepare(db)) {
        snk_db_destroy(db);
        return 1;
    }
    *out = db;
    return 0;
}

int snk_db_destroy(snk_db *db) {
    if (!db) return 0;
    cudaSetDevice(db->device);
    cudaDeviceSynchronize();
    snk_comm_free(db);
    snk_pending_destroy(db);
    delete db->acoustic;
    snk_tc_destroy(db);
    snk_join_tc_free(db);
    for (snk_stage_slot &sl : db->stage) {
        if (sl.ev) cudaEventDestroy(sl.ev);
        if (sl.host) cudaFreeHost(sl.host);
    }
    for (auto &r : db->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    cudaFree(db->F_raw); cudaFree(db->Jc_raw); cudaFree(db->wt); cudaFree(db->wj);
    cudaFree(db->std_mean); cudaFree(db->std_sd);
    cudaFree(db->Fw32); cudaFree(db->Jw32); cudaFree(db->G16); cudaFree(db->S16);
    cudaFree(db->nrm_t16); cudaFree(db->nrm_j16); cudaFree(db->err_t16);
    snk_buf *bufs[] = {&db->ws_q, &db->ws_dist, &db->ws_list, &db->ws_misc, &db->ws_io, &db->ws_io2, &db->ws_tiles,
                       &db->ws_bp, &db->ws_tc, &db->ws_h0, &db->ws_h1, &db->ws_h2, &db->ws_h3, &db->ws_flags,
                       &db->ws_kflags, &db->ws_meta, &db->ws_ag, &db->ws_jv, &db->ws_g1};
    for (snk_buf *b : bufs) snk_buf_free(b);
    if (db->ev) cudaEventDestroy(db->ev);
    for (cudaEvent_t e : db->upload_events) cudaEventDestroy(e);
    if (db->copy_stream) cudaStreamDestroy(db->copy_stream);
    if (db->stream) cudaStreamDestroy(db->stream);
    cudaGetLastError();
    delete db;
    return 0;
}

int snk_db_info(const snk_db *db, int64_t *N, int64_t *Nprime, int *Dt, int *Dj, int *multiepoch, int *joint_dim) {
    SNK_CHECK(db, "db is NULL");
    if (N) *N = db->N;
    if (Nprime) *Nprime = db->Np;
    if (Dt) *Dt = db->Dt;
    if (Dj) *Dj = db->Dj;
    if (multiepoch) *multiepoch = db->m;
    if (joint_dim) *joint_dim = db->Djq + db->m * db->Dt;
    return 0;
}

int snk_db_set_weights(snk_db *db, const double *wt, const double *wj) {
    SNK_CHECK(db && wt && wj, "NULL argument");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    SNK_CUDA(cudaMemcpyAsync(db->wt, wt, (size_t)db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(db->wj, wj, (size_t)db->Dj * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_apply_weights(db, db->stream));
    float stats[4] = {0, 0, 0, 0};   // err_t, err_j, maxn_t, maxn_j
    SNK_CUDA(cudaMemcpyAsync(stats, db->err_t16, sizeof(stats), cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    // fp16 range guard: weighted values (or their squared norms) beyond fp16 disable the tensor-core engine
    db->tc_ok = stats[2] < 6.0e4f && stats[3] < 6.0e4f && stats[0] == stats[0] && stats[1] == stats[1];
    db->weights_set = true;
    db->jsplit_valid = false;   // the operand scale of join_tc.cu follows the weights; refreshed by the next Viterbi search
    return 0;
}

int snk_db_set_engine(snk_db *db, int engine) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    SNK_CHECK(engine == SNK_ENGINE_AUTO || engine == SNK_ENGINE_SIMT || engine == SNK_ENGINE_TC, "unknown engine %d",
              engine);
    db->engine = engine;
    return 0;
}

int snk_db_counters(const snk_db *db, int64_t counters[4], int reset) {
    SNK_CHECK(db && counters, "NULL argument");
    memcpy(counters, db->counters, sizeof(db->counters));
    if (reset) memset(const_cast<snk_db *>(db)->counters, 0, sizeof(db->counters));
    return 0;
}

int snk_db_profile_enable(snk_db *db, int enable) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    db->prof_on = enable != 0;
    return 0;
}

int snk_db_profile_read(snk_db *db, int which, double *total_ms, int64_t *launches, double *work, int reset) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    SNK_CUDA(cudaDeviceSynchronize());
    double ms = 0.0, w = 0.0;
    int64_t n = 0;
    for (auto &r : db->prof) {
        if (r.which != which) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms += t; w += r.work; ++n; }
    }
    cudaGetLastError();
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (work) *work = w;
    if (reset) {
        std::vector<snk_db::prof_rec> keep;
        for (auto &r : db->prof) {
            if (r.which == which) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
            else keep.push_back(r);
        }
        db->prof.swap(keep);
    }
    return 0;
}

int snk_knn_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist, int64_t *d_idx,
                int64_t id_offset, void *stream) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    SNK_CHECK(space == SNK_SPACE_TARGET || space == SNK_SPACE_JOINT, "unknown search space %d", space);
    SNK_CHECK(nq >= 0 && k >= 1, "bad nq / k");
    SNK_CUDA(cudaSetDevice(db->device));
    return snk_knn_enqueue(db, space, dQ, nq, k, d_dist, d_idx, k, id_offset, (cudaStream_t)stream);
}

int snk_knn(snk_db *db, int space, const double *Q, int64_t nq, int k, double *dist, int64_t *idx) {
    SNK_CHECK(db && dist && idx, "NULL argument");
    SNK_LOCK(db);
    SNK_CHECK(space == SNK_SPACE_TARGET || space == SNK_SPACE_JOINT, "unknown search space %d", space);
    SNK_CHECK(nq >= 0 && k >= 1, "bad nq / k");
    if (nq == 0) return 0;
    SNK_CHECK(Q, "Q is NULL");
    SNK_CUDA(cudaSetDevice(db->device));
    const snk_space sp = snk_make_space(db, space);
    // queries go through in slabs so device staging stays bounded
    const int64_t slab = 65536;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)std::min(slab, nq) * sp.D * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, (size_t)std::min(slab, nq) * k * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)std::min(slab, nq) * k * 8));
    for (int64_t q0 = 0; q0 < nq; q0 += slab) {
        const int64_t n = std::min(slab, nq - q0);
        SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, Q + q0 * sp.D, (size_t)n * sp.D * 8, cudaMemcpyHostToDevice, db->stream));
        SNK_TRY(snk_knn_enqueue(db, space, (const double *)db->ws_h0.p, n, k, (double *)db->ws_h1.p, (int64_t *)db->ws_h2.p, k,
                                0, db->stream));
        SNK_TRY(snk_knn_finish(db));
        SNK_CUDA(cudaMemcpyAsync(dist + q0 * k, db->ws_h1.p, (size_t)n * k * 8, cudaMemcpyDeviceToHost, db->stream));
        SNK_CUDA(cudaMemcpyAsync(idx + q0 * k, db->ws_h2.p, (size_t)n * k * 8, cudaMemcpyDeviceToHost, db->stream));
        SNK_CUDA(cudaStreamSynchronize(db->stream));
    }
    return 0;
}

int snk_db_set_standardisation(snk_db *db, const double *mean, const double *std, double special_uv_value,
                               double uv_scaling_factor, unsigned flags) {
    SNK_CHECK(db && mean && std, "NULL argument");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    for (int c = 0; c < db->Dt; ++c) SNK_CHECK(std[c] != 0.0 && std[c] == std[c], "std[%d] is zero or NaN", c);
    if (!db->std_mean) {
        SNK_CUDA(cudaMalloc(&db->std_mean, (size_t)db->Dt * 8));
        SNK_CUDA(cudaMalloc(&db->std_sd, (size_t)db->Dt * 8));
    }
    SNK_CUDA(cudaMemcpyAsync(db->std_mean, mean, (size_t)db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(db->std_sd, std, (size_t)db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    db->uv_special = special_uv_value;
    db->uv_scale = uv_scaling_factor;
    db->std_f32 = (flags & SNK_STD_FLOAT32) ? 1 : 0;
    db->std_set = true;
    return 0;
}

int snk_prepare_targets(snk_db *db, const float *unnorm, int64_t rows, double *out) {
    SNK_CHECK(db && db->std_set, "snk_db_set_standardisation has not been called");
    SNK_LOCK(db);
    SNK_CHECK(rows >= 0, "bad row count");
    if (rows == 0) return 0;
    SNK_CHECK(unnorm && out, "NULL argument");
    SNK_CUDA(cudaSetDevice(db->device));
    const size_t n = (size_t)rows * db->Dt;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, n * 4));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, n * 8));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, unnorm, n * 4, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_prepare_targets_dev(db, (const float *)db->ws_h0.p, rows, (double *)db->ws_h1.p, db->stream));
    SNK_CUDA(cudaMemcpyAsync(out, db->ws_h1.p, n * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

int snk_halfphone_targets(snk_db *db, const float *unnorm, int64_t frames, int dim, const int64_t *points, int64_t n, int P,
                          const double *durations, double *out) {
    SNK_CHECK(db && db->std_set, "snk_db_set_standardisation has not been called");
    SNK_LOCK(db);
    SNK_CHECK(n >= 0 && frames >= 1 && dim >= 1 && P >= 1, "bad shape");
    if (n == 0) return 0;
    SNK_CHECK(unnorm && points && out, "NULL argument");
    for (int64_t i = 0; i < n * P; ++i)      // numpy raises IndexError for these
        SNK_CHECK(points[i] >= -frames && points[i] < frames, "frame index %lld out of bounds for %lld frames",
                  (long long)points[i], (long long)frames);
    SNK_CUDA(cudaSetDevice(db->device));
    const size_t xb = (size_t)frames * dim * 4, pb = (size_t)n * P * 8, db_ = durations ? (size_t)n * 8 : 0;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, xb));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, (size_t)n * db->Dt * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, pb + db_ + 16));
    int64_t *d_pts = (int64_t *)db->ws_h2.p;
    double *d_dur = durations ? (double *)((char *)db->ws_h2.p + pb) : nullptr;
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, unnorm, xb, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(d_pts, points, pb, cudaMemcpyHostToDevice, db->stream));
    if (durations) SNK_CUDA(cudaMemcpyAsync(d_dur, durations, db_, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_halfphone_targets_dev(db, (const float *)db->ws_h0.p, frames, dim, d_pts, n, P, d_dur, (double *)db->ws_h1.p,
                                      db->stream));
    SNK_CUDA(cudaMemcpyAsync(out, db->ws_h1.p, (size_t)n * db->Dt * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

// host entry point shared by the weighted-float64 and the un-normalised-float32 forms (elem = 8 / 4)
static int greedy_batch_host(snk_db *db, const void *targets, size_t elem, const int64_t *lens, int B,
                             const int64_t *start_state, int64_t *paths, double *step_dist) {
    SNK_CHECK(db && lens && paths, "NULL argument");
    SNK_LOCK(db);
    SNK_CHECK(B >= 0, "bad batch size");
    if (B == 0) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    int64_t frames = 0, steps = 0;
    for (int b = 0; b < B; ++b) {
        SNK_CHECK(lens[b] >= 0, "negative utterance length");
        frames += lens[b];
        steps += lens[b] / db->m;
    }
    SNK_CHECK(targets || frames == 0, "targets is NULL");
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)std::max<int64_t>(frames, 1) * db->Dt * elem));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, (size_t)std::max<int64_t>(steps, 1) * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)std::max<int64_t>(steps, 1) * 8));
    // Upload.  Equal-length utterances (the common batch) go up in time slices on a second stream: step t
    // only needs frames [t*m, (t+1)*m) of every utterance, so the search starts after the first slice and
    // the remaining host->device traffic hides behind it.  Ragged batches use one plain copy.
    bool equal = B > 1;
    for (int b = 1; b < B && equal; ++b) equal = lens[b] == lens[0];
    const int64_t T = lens[0], nsteps = T / db->m;
    const int NSLICE = 8;
    db->step_waits.clear();
    if (equal && nsteps >= 4 * NSLICE && frames * db->Dt >= ((int64_t)1 << 20)) {
        while ((int)db->upload_events.size() < NSLICE) {
            cudaEvent_t e;
            SNK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            db->upload_events.push_back(e);
        }
        const size_t pitch = (size_t)T * db->Dt * elem;
        // slice boundaries in steps: a short first slice, the rest even
        int64_t s0 = 0;
        for (int c = 0; c < NSLICE; ++c) {
            const int64_t s1 = c == NSLICE - 1 ? nsteps : std::max<int64_t>(2, (c + 1) * nsteps / NSLICE - nsteps / (2 * NSLICE));
            const int64_t f0 = s0 * db->m, f1 = c == NSLICE - 1 ? T : s1 * db->m;   // the last slice carries the cut remainder
            const size_t off = (size_t)f0 * db->Dt * elem, width = (size_t)(f1 - f0) * db->Dt * elem;
            SNK_CUDA(cudaMemcpy2DAsync((char *)db->ws_h0.p + off, pitch, (const char *)targets + off, pitch, width,
                                       (size_t)B, cudaMemcpyHostToDevice, db->copy_stream));
            SNK_CUDA(cudaEventRecord(db->upload_events[c], db->copy_stream));
            db->step_waits.push_back(std::make_pair(s0, db->upload_events[c]));
            s0 = s1;
        }
    } else if (frames) {
        SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, targets, (size_t)frames * db->Dt * elem, cudaMemcpyHostToDevice, db->stream));
    }
    int64_t *d_paths = (int64_t *)db->ws_h1.p;
    double *d_sd = step_dist ? (double *)db->ws_h2.p : nullptr;
    const int rc_greedy =
        elem == 8 ? snk_greedy_batch_dev(db, (const double *)db->ws_h0.p, lens, B, start_state, d_paths, d_sd, db->stream)
                  : snk_greedy_batch_unnorm_dev(db, (const float *)db->ws_h0.p, lens, B, start_state, d_paths, d_sd,
                                                db->stream);
    db->step_waits.clear();
    const int rc_fin = snk_greedy_batch_finish(db);    // waits for the chain; repairs uncertified utterances
    if (rc_greedy || rc_fin) {
        cudaStreamSynchronize(db->copy_stream);
        return 1;
    }
    if (steps) {
        SNK_CUDA(cudaMemcpyAsync(paths, db->ws_h1.p, (size_t)steps * 8, cudaMemcpyDeviceToHost, db->stream));
        if (step_dist)
            SNK_CUDA(cudaMemcpyAsync(step_dist, db->ws_h2.p, (size_t)steps * 8, cudaMemcpyDeviceToHost, db->stream));
    }
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

int snk_greedy_batch(snk_db *db, const double *targets, const int64_t *lens, int B, const int64_t *start_state,
                     int64_t *paths, double *step_dist) {
    return greedy_batch_host(db, targets, 8, lens, B, start_state, paths, step_dist);
}

int snk_greedy_batch_unnorm(snk_db *db, const float *unnorm, const int64_t *lens, int B, const int64_t *start_state,
                            int64_t *paths, double *step_dist) {
    SNK_CHECK(db && db->std_set, "snk_db_set_standardisation has not been called");
    return greedy_batch_host(db, unnorm, 4, lens, B, start_state, paths, step_dist);
}

int snk_candidate_distances(snk_db *db, const int64_t *cand, const double *targets, int64_t T, int K, double *dist) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(T >= 0 && K >= 1, "bad T / K");
    if (T == 0) return 0;
    SNK_CHECK(cand && targets && dist, "NULL argument");
    SNK_CUDA(cudaSetDevice(db->device));
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)T * K * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, (size_t)T * db->Dt * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)T * K * 8));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, cand, (size_t)T * K * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h1.p, targets, (size_t)T * db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_candidate_distances_dev(db, (const int64_t *)db->ws_h0.p, (const double *)db->ws_h1.p, T, K,
                                        (double *)db->ws_h2.p, db->stream));
    SNK_CUDA(cudaMemcpyAsync(dist, db->ws_h2.p, (size_t)T * K * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

static int count_frames(const int64_t *lens, int B, int64_t *frames, int64_t *tiles) {
    *frames = 0;
    *tiles = 0;
    for (int b = 0; b < B; ++b) {
        SNK_CHECK(lens[b] >= 0, "negative utterance length");
        *frames += lens[b];
        *tiles += lens[b] > 0 ? lens[b] - 1 : 0;
    }
    SNK_CHECK(*frames < (int64_t)2000000000, "too many frames in one batch");
    return 0;
}

int snk_join_tiles(snk_db *db, const int64_t *cand, const int64_t *lens, int B, int K, float *tiles) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(lens && B >= 0 && K >= 1, "bad arguments");
    SNK_CUDA(cudaSetDevice(db->device));
    int64_t frames, ntiles;
    SNK_TRY(count_frames(lens, B, &frames, &ntiles));
    if (ntiles == 0) return 0;
    SNK_CHECK(cand && tiles, "NULL argument");
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)frames * K * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_tiles, (size_t)ntiles * K * K * 4));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, cand, (size_t)frames * K * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_join_tiles_dev(db, (const int64_t *)db->ws_h0.p, lens, B, K, (float *)db->ws_tiles.p, db->stream));
    SNK_CUDA(cudaMemcpyAsync(tiles, db->ws_tiles.p, (size_t)ntiles * K * K * 4, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

int snk_join_stats(const snk_db *db, int64_t stats[2]) {
    SNK_CHECK(db && stats, "NULL argument");
    stats[0] = (int64_t)db->jv_stats[0];
    stats[1] = (int64_t)db->jv_stats[1];
    return 0;
}

int snk_join_viterbi_batch(snk_db *db, const int64_t *cand, const double *tdist, const int64_t *lens, int B, int K,
                           unsigned flags, int64_t *paths, int64_t *path_len, double *path_cost, double *tcost,
                           double *jcost) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(lens && B >= 0 && K >= 1, "bad arguments");
    if (B == 0) return 0;
    SNK_CHECK(path_len && path_cost, "NULL output");
    SNK_CUDA(cudaSetDevice(db->device));
    int64_t frames, ntiles;
    SNK_TRY(count_frames(lens, B, &frames, &ntiles));
    SNK_CHECK(frames == 0 || (cand && tdist && paths), "NULL argument");
    const size_t fK = (size_t)std::max<int64_t>(frames, 1) * K;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, fK * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, fK * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)std::max<int64_t>(frames, 1) * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h3, (size_t)B * 8 * 4));
    int64_t *d_plen = (int64_t *)db->ws_h3.p;
    double *d_pc = (double *)db->ws_h3.p + B, *d_tc = d_pc + B, *d_jc = d_tc + B;
    if (frames) {
        SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, cand, fK * 8, cudaMemcpyHostToDevice, db->stream));
        SNK_CUDA(cudaMemcpyAsync(db->ws_h1.p, tdist, fK * 8, cudaMemcpyHostToDevice, db->stream));
    }
    SNK_TRY(snk_join_viterbi_batch_dev(db, (const int64_t *)db->ws_h0.p, (const double *)db->ws_h1.p, lens, B, K, flags,
                                       (int64_t *)db->ws_h2.p, d_plen, d_pc, d_tc, d_jc, db->stream));
    if (frames) SNK_CUDA(cudaMemcpyAsync(paths, db->ws_h2.p, (size_t)frames * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaMemcpyAsync(path_len, d_plen, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaMemcpyAsync(path_cost, d_pc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    if (tcost) SNK_CUDA(cudaMemcpyAsync(tcost, d_tc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    if (jcost) SNK_CUDA(cudaMemcpyAsync(jcost, d_jc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

static snk_acoustic_job *acoustic_job(snk_db *db) {
    if (!db->acoustic) db->acoustic = new snk_acoustic_job();
    return db->acoustic;
}

int snk_acoustic_viterbi_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B, int K, unsigned flags,
                                   int64_t *d_paths, int64_t *d_path_len, double *d_path_cost, double *d_tcost,
                                   double *d_jcost, void *stream) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(lens && B >= 0 && K >= 1, "bad arguments");
    if (B == 0) return 0;
    SNK_CUDA(cudaSetDevice(db->device));
    int64_t frames, ntiles;
    SNK_TRY(count_frames(lens, B, &frames, &ntiles));
    SNK_CHECK(frames == 0 || (d_targets && d_paths), "NULL argument");
    snk_acoustic_job *job = acoustic_job(db);
    SNK_CHECK(!job->active, "the previous snk_acoustic_viterbi_batch_dev has not been finished");
    const size_t fK = (size_t)std::max<int64_t>(frames, 1) * K;
    SNK_TRY(snk_buf_reserve(&db->ws_jv, fK * 16));
    double *d_dist = (double *)db->ws_jv.p;
    int64_t *d_cand = (int64_t *)(d_dist + fK);
    cudaStream_t st = (cudaStream_t)stream;
    // preselect_units_acoustic (synth_halfphone.py:1359-1366) for every target of every utterance in one search ...
    if (frames) SNK_TRY(snk_knn_enqueue(db, SNK_SPACE_TARGET, d_targets, frames, K, d_dist, d_cand, K, 0, st));
    // ... and viterbi_search (synth_halfphone.py:1399-1436) on the device-resident candidate lists
    SNK_TRY(snk_join_viterbi_batch_dev(db, d_cand, d_dist, lens, B, K, flags, d_paths, d_path_len, d_path_cost, d_tcost,
                                       d_jcost, stream));
    job->active = true;
    job->lens.assign(lens, lens + B);
    job->K = K; job->flags = flags; job->d_paths = d_paths; job->d_plen = d_path_len; job->d_pcost = d_path_cost;
    job->d_tcost = d_tcost; job->d_jcost = d_jcost; job->st = st;
    return 0;
}

int snk_acoustic_viterbi_finish(snk_db *db) {
    SNK_CHECK(db, "db is NULL");
    SNK_LOCK(db);
    snk_acoustic_job *job = acoustic_job(db);
    if (!job->active) return 0;
    job->active = false;
    const int64_t before = db->counters[1];
    SNK_TRY(snk_knn_finish(db));                 // waits; repairs uncertified candidate lists
    if (db->counters[1] != before) {             // the lattice changed under the search: run it again (rare)
        int64_t frames = 0;
        for (int64_t l : job->lens) frames += l;
        const size_t fK = (size_t)std::max<int64_t>(frames, 1) * job->K;
        double *d_dist = (double *)db->ws_jv.p;
        int64_t *d_cand = (int64_t *)(d_dist + fK);
        SNK_TRY(snk_join_viterbi_batch_dev(db, d_cand, d_dist, job->lens.data(), (int)job->lens.size(), job->K, job->flags,
                                           job->d_paths, job->d_plen, job->d_pcost, job->d_tcost, job->d_jcost, job->st));
        SNK_CUDA(cudaStreamSynchronize(job->st));
    }
    return 0;
}

int snk_acoustic_viterbi_batch(snk_db *db, const double *targets, const int64_t *lens, int B, int K, unsigned flags,
                               int64_t *paths, int64_t *path_len, double *path_cost, double *tcost, double *jcost) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(lens && B >= 0 && K >= 1, "bad arguments");
    if (B == 0) return 0;
    SNK_CHECK(path_len && path_cost, "NULL output");
    SNK_CUDA(cudaSetDevice(db->device));
    int64_t frames, ntiles;
    SNK_TRY(count_frames(lens, B, &frames, &ntiles));
    SNK_CHECK(frames == 0 || (targets && paths), "NULL argument");
    const size_t tb = (size_t)std::max<int64_t>(frames, 1) * db->Dt * 8;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, tb));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)std::max<int64_t>(frames, 1) * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h3, (size_t)B * 8 * 4));
    int64_t *d_plen = (int64_t *)db->ws_h3.p;
    double *d_pc = (double *)db->ws_h3.p + B, *d_tc = d_pc + B, *d_jc = d_tc + B;
    // Upload in slices of one search batch on the copy stream: the k-NN of slice i runs while slice i + 1 crosses PCIe
    // (the targets are 8 * Dt bytes per frame -- 120 MB for 1024 x 80 half-phone targets -- and every target is an
    // independent query), then the lattice search runs over all candidate lists.
    const int64_t slice = std::max<int64_t>(4096, (int64_t)db->sm_count * 128 / 256 * 256);
    if (frames > slice) {
        snk_acoustic_job *job = acoustic_job(db);
        SNK_CHECK(!job->active, "the previous snk_acoustic_viterbi_batch_dev has not been finished");
        const size_t fK = (size_t)frames * K;
        SNK_TRY(snk_buf_reserve(&db->ws_jv, fK * 16));
        double *d_dist = (double *)db->ws_jv.p;
        int64_t *d_cand = (int64_t *)(d_dist + fK);
        const int nsl = (int)snk_cdiv(frames, slice);
        while ((int)db->upload_events.size() < nsl) {
            cudaEvent_t e;
            SNK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            db->upload_events.push_back(e);
        }
        auto upload = [&](int c) -> int {
            const int64_t q0 = c * slice, qn = std::min<int64_t>(slice, frames - q0);
            const size_t off = (size_t)q0 * db->Dt * 8;
            SNK_CUDA(cudaMemcpyAsync((char *)db->ws_h0.p + off, (const char *)targets + off, (size_t)qn * db->Dt * 8,
                                     cudaMemcpyHostToDevice, db->copy_stream));
            SNK_CUDA(cudaEventRecord(db->upload_events[c], db->copy_stream));
            return 0;
        };
        int rc = upload(0);
        for (int c = 0; c < nsl && !rc; ++c) {
            // the next slice is queued before this slice's search: with pageable host memory the copy call itself waits
            // for the staging, and it should do so while the GPU is busy
            if (c + 1 < nsl) rc = upload(c + 1);
            if (rc) break;
            const int64_t q0 = c * slice, qn = std::min<int64_t>(slice, frames - q0);
            rc = cudaStreamWaitEvent(db->stream, db->upload_events[c], 0) == cudaSuccess ? 0 : 1;
            if (!rc) rc = snk_knn_enqueue(db, SNK_SPACE_TARGET, (const double *)db->ws_h0.p + q0 * db->Dt, qn, K, d_dist + q0 * K,
                                          d_cand + q0 * K, K, 0, db->stream);
        }
        if (!rc) rc = snk_join_viterbi_batch_dev(db, d_cand, d_dist, lens, B, K, flags, (int64_t *)db->ws_h2.p, d_plen, d_pc, d_tc,
                                                 d_jc, db->stream);
        if (rc) {
            cudaStreamSynchronize(db->copy_stream);
            cudaStreamSynchronize(db->stream);
            return 1;
        }
        job->active = true;
        job->lens.assign(lens, lens + B);
        job->K = K; job->flags = flags; job->d_paths = (int64_t *)db->ws_h2.p; job->d_plen = d_plen; job->d_pcost = d_pc;
        job->d_tcost = d_tc; job->d_jcost = d_jc; job->st = db->stream;
    } else {
        if (frames) SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, targets, (size_t)frames * db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
        SNK_TRY(snk_acoustic_viterbi_batch_dev(db, (const double *)db->ws_h0.p, lens, B, K, flags, (int64_t *)db->ws_h2.p, d_plen,
                                               d_pc, d_tc, d_jc, db->stream));
    }
    SNK_TRY(snk_acoustic_viterbi_finish(db));
    if (frames) SNK_CUDA(cudaMemcpyAsync(paths, db->ws_h2.p, (size_t)frames * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaMemcpyAsync(path_len, d_plen, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaMemcpyAsync(path_cost, d_pc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    if (tcost) SNK_CUDA(cudaMemcpyAsync(tcost, d_tc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    if (jcost) SNK_CUDA(cudaMemcpyAsync(jcost, d_jc, (size_t)B * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

int snk_greedy_path_scores(snk_db *db, const double *targets, int64_t T, const int64_t *path, int64_t P,
                           const int *twidths, int n_tstreams, const int *jwidths, int n_jstreams, double *tscores,
                           double *jscores) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(targets && path && twidths && jwidths && tscores && (jscores || P < 2), "NULL argument");
    SNK_CHECK(P >= 1 && P * db->m <= T, "path length %lld does not fit %lld target frames", (long long)P, (long long)T);
    int ts = 0, js = 0;
    for (int i = 0; i < n_tstreams; ++i) ts += twidths[i];
    for (int i = 0; i < n_jstreams; ++i) js += jwidths[i];
    SNK_CHECK(ts == db->Dt, "target stream widths sum to %d, expected %d", ts, db->Dt);
    SNK_CHECK(js == db->Djq, "join stream widths sum to %d, expected %d", js, db->Djq);
    for (int64_t p = 0; p < P; ++p)   // the kernel gathers rows path[p] .. path[p] + m: reject ids numpy would reject
        SNK_CHECK(path[p] >= 0 && path[p] < db->Np, "path[%lld] = %lld is not a searchable row (0 <= id < %lld)",
                  (long long)p, (long long)path[p], (long long)db->Np);
    SNK_CUDA(cudaSetDevice(db->device));
    SNK_TRY(snk_buf_reserve(&db->ws_h0, (size_t)T * db->Dt * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, (size_t)P * 8));
    SNK_TRY(snk_buf_reserve(&db->ws_h2, (size_t)(n_tstreams + n_jstreams) * 4));
    SNK_TRY(snk_buf_reserve(&db->ws_h3, (size_t)P * (n_tstreams + n_jstreams) * 8));
    int *d_tw = (int *)db->ws_h2.p, *d_jw = d_tw + n_tstreams;
    double *d_ts = (double *)db->ws_h3.p, *d_js = d_ts + P * n_tstreams;
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, targets, (size_t)T * db->Dt * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h1.p, path, (size_t)P * 8, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(d_tw, twidths, (size_t)n_tstreams * 4, cudaMemcpyHostToDevice, db->stream));
    SNK_CUDA(cudaMemcpyAsync(d_jw, jwidths, (size_t)n_jstreams * 4, cudaMemcpyHostToDevice, db->stream));
    SNK_TRY(snk_path_scores_dev(db, (const double *)db->ws_h0.p, T, (const int64_t *)db->ws_h1.p, P, d_tw, n_tstreams,
                                d_jw, n_jstreams, d_ts, d_js, db->stream));
    SNK_CUDA(cudaMemcpyAsync(tscores, d_ts, (size_t)P * n_tstreams * 8, cudaMemcpyDeviceToHost, db->stream));
    if (P > 1)
        SNK_CUDA(cudaMemcpyAsync(jscores, d_js, (size_t)(P - 1) * n_jstreams * 8, cudaMemcpyDeviceToHost, db->stream));
    SNK_CUDA(cudaStreamSynchronize(db->stream));
    return 0;
}

}  // extern "C"
