// Exact float64 re-rank of a shortlist + exactness certificate (+ optional fused list merge).
//
// The reference evaluates distances in float64 on rows weight(f32 data, f64 weights)
// (script/speech_manip.py:209-213) -- both inside the KD-tree (script/synth_simple.py:229,490)
// and in explicit numpy code (script/synth_halfphone.py:1346-1351).  This kernel recomputes
// exactly those float64 values for the shortlisted rows from the RAW float32 matrices and the
// float64 weight vectors, so the final ordering is decided on the reference's own numbers.
//
// kMerge = true: the input is the tensor-core kernel's per-(chunk, column-half) lists
// [nq, nlists, lsz] (ascending keys); the block first selects the KP smallest by rank counting and
// derives tau = min over lists of the list's largest key, then proceeds as below.  This fuses what
// used to be three launches (chunk_tau, topk_scan merge, rerank) of the greedy step.
#include "rerank_dev.cuh"

namespace {

constexpr int SNK_CERT_FP16 = 1, SNK_CERT_FP32 = 2;   // = SNK_CERT_MODE_* of common.cuh (0: no certificate)
constexpr int RR_THREADS = 256;      // block size for few queries (latency matters: more warps per query)
constexpr int RR_THREADS_SMALL = 128;   // block size for many queries: 12 blocks per SM, so 1024 queries are one wave
constexpr int RR_MAX_MERGE = 2048;   // most list entries a block merges in shared memory

template <bool kMerge, int THREADS>
__global__ void __launch_bounds__(THREADS)
rerank_kernel(rr_space sp, const double *__restrict__ Q, const float *__restrict__ val, const int *__restrict__ id,
              int KP, int nlists, int lsz, int k, double *__restrict__ odist, int64_t *__restrict__ oidx,
              int64_t ostride, int64_t id_offset, int64_t nrows, const float *__restrict__ qerr,
              const float *__restrict__ dberr, const float *__restrict__ qn, const float *__restrict__ maxn,
              const float *__restrict__ tau_extra, int *__restrict__ cert, int *__restrict__ nfail, int sticky,
              const int *__restrict__ qsel, int debug_fail_mod, float eps_rel, int cert_mode,
              double *__restrict__ bound_out) {
    extern __shared__ double sm[];
    double *q_s = sm;                       // [D]
    double *wA_s = q_s + sp.D;              // [dA]
    double *wB_s = wA_s + sp.dA;            // [dB] target weights repeated over the window
    double *d2 = wB_s + sp.dB;              // [KP]
    int *ids = reinterpret_cast<int *>(d2 + KP);          // [KP]
    float *sval = reinterpret_cast<float *>(ids + KP);    // [KP] approximate keys of the selected rows
    // kMerge: [n] packed list entries, then [nwarp * KP] phase-1 winners
    unsigned long long *mkey = reinterpret_cast<unsigned long long *>(sval + KP);
    __shared__ float s_wtau[THREADS / 32];
    __shared__ double s_qn2[THREADS / 32];   // SNK_CERT_FP32: squared norm of the float64 query

    const int64_t ql = blockIdx.x;                    // index into the (compact) shortlist arrays
    const int64_t q = qsel ? qsel[ql] : ql;           // index into queries / outputs / per-query bounds
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = THREADS >> 5;

    double qn2 = 0.0;
    for (int d = tid; d < sp.D; d += THREADS) {
        const double x = Q[q * (int64_t)sp.D + d];
        q_s[d] = x;
        qn2 += x * x;
    }
    if (cert_mode == SNK_CERT_FP32) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) qn2 += __shfl_xor_sync(0xffffffffu, qn2, off);
        if (lane == 0) s_qn2[warp] = qn2;
    }
    for (int d = tid; d < sp.dA; d += THREADS) wA_s[d] = sp.wA[sp.a_col + d];
    for (int d = tid; d < sp.dB; d += THREADS) wB_s[d] = sp.wB[d % sp.Dt];

    if (kMerge) {
        const int n = nlists * lsz;
        const float *gv = val + ql * (int64_t)n;
        const int *gi = id + ql * (int64_t)n;
        for (int t = tid; t < KP; t += THREADS) { sval[t] = INFINITY; ids[t] = INT_MAX; }   // n may be < KP
        // a row dropped inside a list has a key >= that list's largest kept key
        float tl = INFINITY;
        for (int t = tid; t < n; t += THREADS) {
            const int i = gi[t];
            const float v = gv[t];
            mkey[t] = i >= 0 ? pack_key(v, i) : pad_key(t);
            if (t % lsz == lsz - 1) tl = fminf(tl, i >= 0 ? v : INFINITY);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tl = fminf(tl, __shfl_xor_sync(0xffffffffu, tl, off));
        if (lane == 0) s_wtau[warp] = tl;
        __syncthreads();
        // selection by rank counting: the packed entries are distinct, so ranks are a permutation
        if (n <= 256) {
            for (int t = tid; t < n; t += THREADS) {
                const unsigned long long me = mkey[t];
                int rank = 0;
#pragma unroll 4
                for (int j = 0; j < n; ++j) rank += mkey[j] < me ? 1 : 0;
                if (rank < KP) unpack_key(me, sval[rank], ids[rank]);
            }
        } else {
            // two phases: every warp ranks its own slice and keeps the slice's KP smallest, then the
            // nwarp * KP winners are ranked.  An entry dropped in phase 1 is >= its slice's KP-th smallest,
            // hence >= the final KP-th smallest, so the final list's maximum still bounds every drop.
            unsigned long long *wk = mkey + n;          // [nwarp * KP] winners
            for (int t = tid; t < nwarp * KP; t += THREADS) wk[t] = pad_key(n + t);
            __syncthreads();
            const int ns = (n + nwarp - 1) / nwarp, s0 = warp * ns, s1 = min(n, s0 + ns);
            for (int t = s0 + lane; t < s1; t += 32) {
                const unsigned long long me = mkey[t];
                int rank = 0;
#pragma unroll 4
                for (int j = s0; j < s1; ++j) rank += mkey[j] < me ? 1 : 0;
                if (rank < KP) wk[warp * KP + rank] = me;
            }
            __syncthreads();
            const int nw = nwarp * KP;
            for (int t = tid; t < nw; t += THREADS) {
                const unsigned long long me = wk[t];
                int rank = 0;
#pragma unroll 4
                for (int j = 0; j < nw; ++j) rank += wk[j] < me ? 1 : 0;
                if (rank < KP) unpack_key(me, sval[rank], ids[rank]);
            }
        }
    } else {
        for (int t = tid; t < KP; t += THREADS) {
            const int i = id[ql * KP + t];
            sval[t] = i >= 0 ? val[ql * KP + t] : INFINITY;
            ids[t] = i >= 0 ? i : INT_MAX;
        }
    }
    __syncthreads();

    // exact distances, four shortlisted rows per warp pass
    for (int c = 4 * warp; c < KP; c += 4 * nwarp) {
        int u[4];
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (c + i < KP && ids[c + i] != INT_MAX) ? ids[c + i] : -1;
        rows_dist<4>(sp, q_s, wA_s, wB_s, u, lane, r);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (c + i < KP) d2[c + i] = r[i];
        }
    }
    __syncthreads();

    for (int t = tid; t < KP; t += THREADS) {
        const double v = d2[t];
        const int i = ids[t];
        int rank = 0;
        for (int j = 0; j < KP; ++j)   // equal pairs (only the empty slots) keep their list order
            rank += (dpair_lt(d2[j], ids[j], v, i) || (d2[j] == v && ids[j] == i && j < t)) ? 1 : 0;
        if (rank < k) {
            const bool ok = i != INT_MAX;
            odist[q * ostride + rank] = ok ? sqrt(v) : INFINITY;
            oidx[q * ostride + rank] = ok ? (int64_t)i + id_offset : nrows;
            if (cert && rank == k - 1) {
                // tau: smallest approximate key any row OUTSIDE the shortlist can have
                float mx = -INFINITY;
                bool full = true;
                for (int j = 0; j < KP; ++j) {
                    if (ids[j] == INT_MAX) full = false;
                    else mx = fmaxf(mx, sval[j]);
                }
                if (!full) mx = INFINITY;                       // the final merge dropped nothing
                if (tau_extra) mx = fminf(mx, tau_extra[q]);    // ... but an earlier stage may have
                if (kMerge)
                    for (int w = 0; w < THREADS / 32; ++w) mx = fminf(mx, s_wtau[w]);
                int good = 1;
                double bound = INFINITY;     // every row outside the shortlist is at least this far from the query
                const double dk = ok ? sqrt(v) : INFINITY;
                if (mx < INFINITY && cert_mode == SNK_CERT_FP32) {
                    // fp32 direct-difference keys (knn_simt.cu) of float32-rounded operands.  A dropped row y that truly
                    // beat the k-th answer (d(x,y) < dk) would have ||y|| < ||x|| + dk, hence a rounded-operand distance
                    // below U = dk + 2^-24 (2||x|| + dk) and an fp32 key below U^2 (1 + eps_rel), eps_rel >= (D + 3) 2^-24
                    // bounding the rounding of D exact-sign terms.  Every dropped key is >= mx, so U^2 (1 + eps_rel) < mx
                    // certifies the answer.
                    double xn2 = 0.0;
                    for (int w = 0; w < THREADS / 32; ++w) xn2 += s_qn2[w];
                    const double U = dk + 5.9604644775390625e-8 * (2.0 * sqrt(xn2) + dk);
                    good = (U * U * (1.0 + (double)eps_rel) < (double)mx) ? 1 : 0;
                    // the same inequality solved for dk: any dk below this value would have been certified
                    bound = (sqrt(fmax((double)mx, 0.0) / (1.0 + (double)eps_rel)) - 1.1920928955078125e-7 * sqrt(xn2)) /
                            (1.0 + 5.9604644775390625e-8) * (1.0 - 1e-12);
                } else if (mx < INFINITY) {
                    // fp16 tensor-core keys (rerank_dev.cuh)
                    good = cert_fp16(mx, qn ? qn[q] : 0.f, maxn ? *maxn : 0.f, eps_rel, qerr ? qerr[q] : 0.f,
                                     dberr ? *dberr : 0.f, dk, bound);
                }
                if (debug_fail_mod > 0 && q % debug_fail_mod == 0) { good = 0; bound = -INFINITY; }   // test hook: exercise the re-search paths
                if (bound_out) {
                    // database-sharded search: this shard's answer is only one candidate for the global one.  Export the
                    // bound instead of judging here: after the exchange the global best d* is certified iff d* <= bound
                    // on every rank (a rank whose own best is certified has d* <= its best <= its bound).
                    bound_out[q] = bound;
                    good = 1;
                }
                // sticky: the flag array is preset to 1 by the caller and only ever cleared (a greedy batch
                // inspects it once, after the last step, instead of synchronising every step)
                if (!sticky) cert[q] = good;
                else if (!good) cert[q] = 0;
                if (!good) atomicAdd(nfail, 1);
            }
        }
    }
}

size_t rr_smem(const rr_space &rs, int KP, int nmerge) {
    return (size_t)(rs.D + rs.dA + rs.dB + KP) * 8 + (size_t)KP * 8 + (size_t)nmerge * 8 +
           (nmerge ? (size_t)(RR_THREADS / 32) * KP * 8 : 0) + 16;
}

// many queries: small blocks (one wave of 12 per SM); few queries: large blocks (shorter chain per query)
bool rr_small_blocks(const snk_db *db, int64_t nq) { return nq > (int64_t)6 * db->sm_count; }

}  // namespace

int snk_rerank(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const float *d_val,
               const int *d_id, int KP, int k, double *d_dist, int64_t *d_idx, int64_t out_stride,
               int64_t id_offset, const float *d_qerr, const float *d_dberr, const float *d_qn,
               const float *d_maxn, const float *d_tau_extra, int *d_cert, int *d_nfail, int sticky,
               const int *d_qsel, float eps_rel, int cert_mode, cudaStream_t st) {
    if (nq <= 0) return 0;
    SNK_CHECK(k <= KP, "rerank: k (%d) exceeds shortlist (%d)", k, KP);
    const rr_space rs = make_rr(db, sp);
    const size_t smem = rr_smem(rs, KP, 0);
    const int dbg = !d_cert ? 0 : (cert_mode == SNK_CERT_FP32 ? db->debug_fail_mod2 : db->debug_fail_mod);
    snk_prof_scope prof(db, SNK_PROF_RERANK, (double)nq * KP * sp.D * 4, st);      // work = row bytes gathered
    if (rr_small_blocks(db, nq))
        rerank_kernel<false, RR_THREADS_SMALL><<<(unsigned)nq, RR_THREADS_SMALL, smem, st>>>(
            rs, dQ, d_val, d_id, KP, 0, 0, k, d_dist, d_idx, out_stride, id_offset, sp.rows, d_qerr, d_dberr, d_qn,
            d_maxn, d_tau_extra, d_cert, d_nfail, sticky, d_qsel, dbg, eps_rel, cert_mode, d_cert ? db->cert_bound_out : nullptr);
    else
        rerank_kernel<false, RR_THREADS><<<(unsigned)nq, RR_THREADS, smem, st>>>(
            rs, dQ, d_val, d_id, KP, 0, 0, k, d_dist, d_idx, out_stride, id_offset, sp.rows, d_qerr, d_dberr, d_qn,
            d_maxn, d_tau_extra, d_cert, d_nfail, sticky, d_qsel, dbg, eps_rel, cert_mode, d_cert ? db->cert_bound_out : nullptr);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

bool snk_merge_rerank_fits(int nlists, int lsz) { return nlists * lsz <= RR_MAX_MERGE; }

int snk_merge_rerank(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const float *d_lval,
                     const int *d_lid, int nlists, int lsz, int KP, int k, double *d_dist, int64_t *d_idx,
                     int64_t out_stride, int64_t id_offset, const float *d_qerr, const float *d_dberr,
                     const float *d_qn, const float *d_maxn, int *d_cert, int *d_nfail, int sticky, float eps_rel,
                     cudaStream_t st) {
    if (nq <= 0) return 0;
    SNK_CHECK(k <= KP, "rerank: k (%d) exceeds shortlist (%d)", k, KP);
    SNK_CHECK(snk_merge_rerank_fits(nlists, lsz), "merge_rerank: %d list entries exceed the in-block merge", nlists * lsz);
    const rr_space rs = make_rr(db, sp);
    const size_t smem = rr_smem(rs, KP, nlists * lsz);
    const int dbg = d_cert ? db->debug_fail_mod : 0;
    snk_prof_scope prof(db, SNK_PROF_RERANK, (double)nq * KP * sp.D * 4, st);
    if (rr_small_blocks(db, nq))
        rerank_kernel<true, RR_THREADS_SMALL><<<(unsigned)nq, RR_THREADS_SMALL, smem, st>>>(
            rs, dQ, d_lval, d_lid, KP, nlists, lsz, k, d_dist, d_idx, out_stride, id_offset, sp.rows, d_qerr, d_dberr,
            d_qn, d_maxn, nullptr, d_cert, d_nfail, sticky, nullptr, dbg, eps_rel, SNK_CERT_FP16, d_cert ? db->cert_bound_out : nullptr);
    else
        rerank_kernel<true, RR_THREADS><<<(unsigned)nq, RR_THREADS, smem, st>>>(
            rs, dQ, d_lval, d_lid, KP, nlists, lsz, k, d_dist, d_idx, out_stride, id_offset, sp.rows, d_qerr, d_dberr,
            d_qn, d_maxn, nullptr, d_cert, d_nfail, sticky, nullptr, dbg, eps_rel, SNK_CERT_FP16, d_cert ? db->cert_bound_out : nullptr);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Exhaustive float64 search: the last resort for a query that neither the tensor-core nor the fp32 shortlist could
// certify (more than a shortlist of near-ties).  Same arithmetic as the re-rank, over EVERY row: the result is the
// float64 brute-force answer by construction.  Slow (one query at a time) and rare.
namespace {

__global__ void exact_dist_kernel(rr_space sp, const double *__restrict__ Q, int64_t q, int64_t rows, double *__restrict__ d2) {
    extern __shared__ double sm[];
    double *q_s = sm, *wA_s = q_s + sp.D, *wB_s = wA_s + sp.dA;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int d = tid; d < sp.D; d += blockDim.x) q_s[d] = Q[q * (int64_t)sp.D + d];
    for (int d = tid; d < sp.dA; d += blockDim.x) wA_s[d] = sp.wA[sp.a_col + d];
    for (int d = tid; d < sp.dB; d += blockDim.x) wB_s[d] = sp.wB[d % sp.Dt];
    __syncthreads();
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + tid) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = 4 * warp; u < rows; u += 4 * nwarp) {
        int uu[4];
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) uu[i] = u + i < rows ? (int)(u + i) : -1;
        rows_dist<4>(sp, q_s, wA_s, wB_s, uu, lane, r);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (u + i < rows) d2[u + i] = r[i];
        }
    }
}

// k rounds of "smallest (d2, id) pair after the previous one"; one block
__global__ void __launch_bounds__(1024)
exact_select_kernel(const double *__restrict__ d2, int64_t rows, int k, double *__restrict__ odist, int64_t *__restrict__ oidx,
                    int64_t id_offset) {
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ double s_lastv;
    __shared__ int s_lasti;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_lastv = -1.0; s_lasti = -1; }
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        const double lv = s_lastv;
        const int li = s_lasti;
        double bv = INFINITY;
        int bi = INT_MAX;
        for (int64_t u = tid; u < rows; u += blockDim.x) {
            const double v = d2[u];
            if (dpair_lt(lv, li, v, (int)u) && dpair_lt(v, (int)u, bv, bi)) { bv = v; bi = (int)u; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (dpair_lt(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = lane < (int)(blockDim.x >> 5) ? s_v[lane] : INFINITY;
            bi = lane < (int)(blockDim.x >> 5) ? s_i[lane] : INT_MAX;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (dpair_lt(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                const bool ok = bi != INT_MAX;
                odist[r] = ok ? sqrt(bv) : INFINITY;
                oidx[r] = ok ? (int64_t)bi + id_offset : rows;
                s_lastv = ok ? bv : INFINITY;
                s_lasti = bi;
            }
        }
        __syncthreads();
    }
}

}  // namespace

int snk_exact_search(snk_db *db, const snk_space &sp, const double *dQ, const int *h_qidx, int n, int k, double *d_dist,
                     int64_t *d_idx, int64_t out_stride, int64_t id_offset, cudaStream_t st) {
    if (n <= 0) return 0;
    const rr_space rs = make_rr(db, sp);
    SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)sp.rows * 8));
    double *d2 = (double *)db->ws_dist.p;
    const size_t smem = (size_t)(rs.D + rs.dA + rs.dB) * 8;
    for (int i = 0; i < n; ++i) {
        const int64_t q = h_qidx[i];
        exact_dist_kernel<<<db->sm_count * 4, 256, smem, st>>>(rs, dQ, q, sp.rows, d2);
        SNK_CUDA(cudaGetLastError());
        exact_select_kernel<<<1, 1024, 0, st>>>(d2, sp.rows, k, d_dist + q * out_stride, d_idx + q * out_stride, id_offset);
        SNK_CUDA(cudaGetLastError());
        db->counters[2] += 2;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// k-way merge of per-shard results after the all-gather of a database-sharded search
// (SURVEY.md section 8e): R ascending lists of k (distance, global row id) pairs per query -> the k smallest overall;
// ties: lowest global row id.  One block per query: the R * k pairs are staged in shared memory; the output rank of a
// pair is its own position plus, for every other list, the number of pairs that precede it there (binary search), so the
// work per query is R k (R log k) instead of (R k)^2.
// Layout: dist_all + r * rank_stride_d, idx_all + r * rank_stride_i are [nq, k] arrays.
namespace {
__device__ __forceinline__ bool mpair_lt(double v1, int64_t i1, double v2, int64_t i2) { return v1 < v2 || (v1 == v2 && i1 < i2); }

__global__ void topk_merge_kernel(const double *__restrict__ dist_all, const int64_t *__restrict__ idx_all, int64_t stride_d,
                                  int64_t stride_i, int R, int64_t nq, int k, double *__restrict__ odist,
                                  int64_t *__restrict__ oidx) {
    extern __shared__ double msm[];
    double *sv = msm;                                         // [R * k]
    int64_t *si = reinterpret_cast<int64_t *>(sv + R * k);    // [R * k]
    const int64_t q = blockIdx.x;
    const int n = R * k;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int r = t / k, j = t % k;
        sv[t] = dist_all[r * stride_d + q * k + j];
        si[t] = idx_all[r * stride_i + q * k + j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int r = t / k, j = t % k;
        const double v = sv[t];
        const int64_t i = si[t];
        int rank = j;
        for (int r2 = 0; r2 < R; ++r2) {
            if (r2 == r) continue;
            // pairs of list r2 that come before (v, i); equal pairs (duplicates across shards) order by shard
            int lo = 0, hi = k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const double v2 = sv[r2 * k + mid];
                const int64_t i2 = si[r2 * k + mid];
                const bool before = mpair_lt(v2, i2, v, i) || (v2 == v && i2 == i && r2 < r);
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            odist[q * k + rank] = v;
            oidx[q * k + rank] = i;
        }
    }
}
}  // namespace

int snk_topk_merge_launch(const double *d_dist_all, const int64_t *d_idx_all, int64_t stride_d, int64_t stride_i, int R,
                          int64_t nq, int k, double *d_dist, int64_t *d_idx, cudaStream_t st) {
    if (nq == 0) return 0;
    const size_t smem = (size_t)R * k * 16;
    SNK_CHECK(smem <= 200 * 1024, "merge of %d lists of %d does not fit shared memory", R, k);
    if (smem > 48 * 1024)
        SNK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = R * k >= 256 ? 256 : (R * k >= 64 ? 128 : 32);
    topk_merge_kernel<<<(unsigned)nq, threads, smem, st>>>(d_dist_all, d_idx_all, stride_d, stride_i, R, nq, k, d_dist, d_idx);
    SNK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int snk_topk_merge_dev(int device_id, const double *d_dist_all, const int64_t *d_idx_all, int R, int64_t nq,
                                  int k, double *d_dist, int64_t *d_idx, void *stream) {
    SNK_CHECK(R >= 1 && k >= 1 && nq >= 0, "bad merge shape");
    SNK_CUDA(cudaSetDevice(device_id));
    return snk_topk_merge_launch(d_dist_all, d_idx_all, nq * k, nq * k, R, nq, k, d_dist, d_idx, (cudaStream_t)stream);
}
