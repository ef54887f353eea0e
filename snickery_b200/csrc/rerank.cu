// Exact float64 re-rank of a shortlist + exactness certificate.
//
// The reference evaluates distances in float64 on rows weight(f32 data, f64 weights)
// (script/speech_manip.py:209-213) -- both inside the KD-tree (script/synth_simple.py:229,490)
// and in explicit numpy code (script/synth_halfphone.py:1346-1351).  This kernel recomputes
// exactly those float64 values for the shortlisted rows from the RAW float32 matrices and the
// float64 weight vectors, so the final ordering is decided on the reference's own numbers.
#include "common.cuh"
#include <limits.h>

namespace {

struct rr_space {
    const float *A;    // Jc_raw
    const float *B;    // F_raw
    const double *wA;  // wj
    const double *wB;  // wt
    int dA, D, a_row_off, a_col, ldA, ldB, periodB;
};

__device__ __forceinline__ double rr_elem(const rr_space &sp, int64_t u, int d) {
    if (d < sp.dA) {
        const int c = sp.a_col + d;
        return (double)__ldg(sp.A + (u + sp.a_row_off) * (int64_t)sp.ldA + c) * __ldg(sp.wA + c);
    }
    const int c = d - sp.dA;
    return (double)__ldg(sp.B + u * (int64_t)sp.ldB + c) * __ldg(sp.wB + (c % sp.periodB));
}

__device__ __forceinline__ bool dpair_lt(double v1, int i1, double v2, int i2) {
    return v1 < v2 || (v1 == v2 && i1 < i2);
}

__global__ void __launch_bounds__(128)
rerank_kernel(rr_space sp, const double *__restrict__ Q, const float *__restrict__ val,
              const int *__restrict__ id, int KP, int k, double *__restrict__ odist,
              int64_t *__restrict__ oidx, int64_t ostride, int64_t id_offset, int64_t nrows,
              const float *__restrict__ qerr, const float *__restrict__ dberr,
              const float *__restrict__ qn, const float *__restrict__ maxn,
              const float *__restrict__ tau_extra, int *__restrict__ cert, int *__restrict__ nfail,
              const int *__restrict__ qsel) {
    extern __shared__ double sm[];
    double *d2 = sm;
    int *ids = reinterpret_cast<int *>(sm + KP);
    const int64_t ql = blockIdx.x;                    // index into the (compact) shortlist arrays
    const int64_t q = qsel ? qsel[ql] : ql;           // index into queries / outputs / per-query bounds
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double *qrow = Q + q * (int64_t)sp.D;
    for (int c = warp; c < KP; c += nwarp) {
        const int u = id[ql * KP + c];
        double acc = INFINITY;
        if (u >= 0) {
            acc = 0.0;
            for (int d = lane; d < sp.D; d += 32) {
                const double df = __dsub_rn(qrow[d], rr_elem(sp, u, d));
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        }
        if (lane == 0) {
            d2[c] = acc;
            ids[c] = u >= 0 ? u : INT_MAX;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP; t += blockDim.x) {
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
                // every row outside the shortlist has approx distance >= tau; its true distance is
                // >= tau - (||dx|| + ||dy||) by the triangle inequality (fp16 rounding of both sides)
                // tau: smallest approximate key any row OUTSIDE the shortlist can have
                float mx = -INFINITY;
                bool full = true;
                for (int j = 0; j < KP; ++j) {
                    const float a = val[ql * KP + j];
                    if (id[ql * KP + j] < 0) full = false;
                    else mx = fmaxf(mx, a);
                }
                if (!full) mx = INFINITY;                       // the final merge dropped nothing
                if (tau_extra) mx = fminf(mx, tau_extra[q]);    // ... but an earlier stage may have
                int good = 1;
                if (mx < INFINITY) {
                    const float qq = qn ? qn[q] : 0.f;
                    const float eps = 4e-6f * (qq + (maxn ? *maxn : 0.f));
                    const double tau2 = (double)mx + (double)qq - (double)eps;
                    const double tau = tau2 > 0.0 ? sqrt(tau2) : 0.0;
                    const double delta = (double)(qerr ? qerr[q] : 0.f) + (double)(dberr ? *dberr : 0.f);
                    const double dk = ok ? sqrt(v) : INFINITY;
                    good = (dk + delta <= tau) ? 1 : 0;
                }
                cert[q] = good;
                if (!good) atomicAdd(nfail, 1);
            }
        }
    }
}

}  // namespace

int snk_rerank(snk_db *db, const snk_space &sp, const double *dQ, int64_t nq, const float *d_val,
               const int *d_id, int KP, int k, double *d_dist, int64_t *d_idx, int64_t out_stride,
               int64_t id_offset, const float *d_qerr, const float *d_dberr, const float *d_qn,
               const float *d_maxn, const float *d_tau_extra, int *d_cert, const int *d_qsel,
               cudaStream_t st) {
    if (nq <= 0) return 0;
    SNK_CHECK(k <= KP, "rerank: k (%d) exceeds shortlist (%d)", k, KP);
    rr_space rs;
    rs.A = db->Jc_raw; rs.B = db->F_raw; rs.wA = db->wj; rs.wB = db->wt;
    rs.dA = sp.dA; rs.D = sp.D; rs.a_row_off = sp.a_row_off; rs.a_col = sp.a_col;
    rs.ldA = sp.ldA_raw; rs.ldB = sp.ldB_raw; rs.periodB = db->Dt;
    const size_t smem = (size_t)KP * (sizeof(double) + sizeof(int));
    rerank_kernel<<<(unsigned)nq, 128, smem, st>>>(rs, dQ, d_val, d_id, KP, k, d_dist, d_idx, out_stride,
                                                    id_offset, sp.rows, d_qerr, d_dberr, d_qn, d_maxn,
                                                    d_tau_extra, d_cert, d_cert ? d_cert + nq : nullptr,
                                                    d_qsel);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// k-way merge of per-shard results after the all-gather of a database-sharded search
// (SURVEY.md section 8e): [R, nq, k] ascending lists -> [nq, k]; ties: lowest global row id.
namespace {
__global__ void topk_merge_kernel(const double *__restrict__ dist_all, const int64_t *__restrict__ idx_all, int R,
                                  int64_t nq, int k, double *__restrict__ odist, int64_t *__restrict__ oidx) {
    const int64_t q = blockIdx.x;
    const int n = R * k;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int r = t / k, j = t % k;
        const double v = dist_all[((int64_t)r * nq + q) * k + j];
        const int64_t i = idx_all[((int64_t)r * nq + q) * k + j];
        int rank = 0;
        for (int s = 0; s < n; ++s) {
            const int rs = s / k, js = s % k;
            const double v2 = dist_all[((int64_t)rs * nq + q) * k + js];
            const int64_t i2 = idx_all[((int64_t)rs * nq + q) * k + js];
            rank += (v2 < v || (v2 == v && (i2 < i || (i2 == i && s < t)))) ? 1 : 0;
        }
        if (rank < k) {
            odist[q * k + rank] = v;
            oidx[q * k + rank] = i;
        }
    }
}
}  // namespace

extern "C" int snk_topk_merge_dev(int device_id, const double *d_dist_all, const int64_t *d_idx_all, int R, int64_t nq,
                                  int k, double *d_dist, int64_t *d_idx, void *stream) {
    SNK_CHECK(R >= 1 && k >= 1 && nq >= 0, "bad merge shape");
    if (nq == 0) return 0;
    SNK_CUDA(cudaSetDevice(device_id));
    topk_merge_kernel<<<(unsigned)nq, 128, 0, (cudaStream_t)stream>>>(d_dist_all, d_idx_all, R, nq, k, d_dist, d_idx);
    SNK_CUDA(cudaGetLastError());
    return 0;
}
