// Device helpers shared by the re-rank kernels (rerank.cu) and the single-utterance greedy kernel (greedy_one.cu):
// packed (key, id) ordering, the float64 row arithmetic of the reference (script/speech_manip.py:209-213) and the
// exactness certificate of the fp16 keys.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace {

// (key, row id) packed so that one 64-bit compare orders by key, then id.  Padding entries (no row)
// sort after every real one and stay distinct through their position.
__device__ __forceinline__ unsigned long long pack_key(float v, int id) {
    unsigned u = __float_as_uint(v + 0.0f);            // -0 -> +0
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);    // order-preserving map of IEEE floats to unsigned
    return ((unsigned long long)u << 32) | (unsigned)id;
}
__device__ __forceinline__ unsigned long long pad_key(int pos) {
    return 0xFFFFFFFF00000000ull | (0x80000000u + (unsigned)pos);
}
__device__ __forceinline__ void unpack_key(unsigned long long k, float &v, int &id) {
    const unsigned hi = (unsigned)(k >> 32);
    if (hi == 0xFFFFFFFFu) { v = INFINITY; id = INT_MAX; return; }
    v = __uint_as_float((hi & 0x80000000u) ? (hi & 0x7FFFFFFFu) : ~hi);
    id = (int)(unsigned)k;
}

struct rr_space {
    const float *A;    // Jc_raw
    const float *B;    // F_raw
    const double *wA;  // wj
    const double *wB;  // wt
    int dA, dB, D, a_row_off, a_col, ldA, ldB, Dt, m;
};

__device__ __forceinline__ bool dpair_lt(double v1, int i1, double v2, int i2) {
    return v1 < v2 || (v1 == v2 && i1 < i2);
}

// sum over a contiguous segment of (q - f32 row * f64 weight)^2 for R rows at once.  The row value is the reference's own:
// the float32 voice value times the float64 weight, ROUNDED to float64 (speech_manip.py:209-213); the difference to the query
// is rounded, its square is fused into the sum.  A row's arithmetic does not depend on R, so every caller gets the same bits.
// Loads are issued four deep before they are consumed so the gathers overlap; the query and the weights are read from
// shared memory once for all R rows.
template <int R>
__device__ __forceinline__ void seg_rows(const float *const (&r)[R], const double *__restrict__ w_s,
                                         const double *__restrict__ q_s, int n, int lane, double (&a)[R]) {
    int d = lane;
    for (; d + 96 < n; d += 128) {
        float y[R][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < R; ++i) y[i][u] = __ldg(r[i] + d + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double w = w_s[d + 32 * u], x = q_s[d + 32 * u];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const double e = __dsub_rn(x, __dmul_rn((double)y[i][u], w));
                a[i] = __fma_rn(e, e, a[i]);
            }
        }
    }
    for (; d < n; d += 32) {
        const double w = w_s[d], x = q_s[d];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const double e = __dsub_rn(x, __dmul_rn((double)__ldg(r[i] + d), w));
            a[i] = __fma_rn(e, e, a[i]);
        }
    }
}

// squared float64 distances of the rows u[0..R) (u[i] < 0: skipped, +inf) to the query held in shared memory.
// wB_s holds the target weights repeated over the m frames of the window (the window is contiguous in F_raw).
template <int R>
__device__ __forceinline__ void rows_dist(const rr_space &sp, const double *__restrict__ q_s, const double *__restrict__ wA_s,
                                          const double *__restrict__ wB_s, const int (&u)[R], int lane, double (&out)[R]) {
    double a[R];
    int v[R];
    int first = -1;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        a[i] = 0.0;
        if (first < 0 && u[i] >= 0) first = u[i];
    }
    if (first < 0) {
#pragma unroll
        for (int i = 0; i < R; ++i) out[i] = INFINITY;
        return;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = u[i] >= 0 ? u[i] : first;       // absent rows recompute a present one
    if (sp.dA > 0) {
        const float *ra[R];
#pragma unroll
        for (int i = 0; i < R; ++i) ra[i] = sp.A + ((int64_t)v[i] + sp.a_row_off) * sp.ldA + sp.a_col;
        seg_rows<R>(ra, wA_s, q_s, sp.dA, lane, a);
    }
    const float *rb[R];
#pragma unroll
    for (int i = 0; i < R; ++i) rb[i] = sp.B + (int64_t)v[i] * sp.ldB;
    seg_rows<R>(rb, wB_s, q_s + sp.dA, sp.dB, lane, a);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < R; ++i) a[i] = __dadd_rn(a[i], __shfl_xor_sync(0xffffffffu, a[i], off));
#pragma unroll
    for (int i = 0; i < R; ++i) out[i] = u[i] >= 0 ? a[i] : INFINITY;
}


// Certificate of an answer found through fp16 keys (tensor-core or mma shortlist): key + ||x~||^2 is the squared distance
// between the ROUNDED vectors up to eps = eps_rel (||x~||^2 + 2 max||y~||^2), the fp32 accumulation bound derived in
// DESIGN.md section 2; the triangle inequality adds the rounding of the vectors.  mx: smallest key any row outside the
// shortlist can have; dk: the k-th exact distance.  bound: every row outside the shortlist is at least this far away.
__device__ __forceinline__ int cert_fp16(float mx, float qq, float maxn, float eps_rel, float qerr, float dberr, double dk,
                                         double &bound) {
    const float eps = eps_rel * (qq + 2.f * maxn);
    const double tau2 = (double)mx + (double)qq - (double)eps;
    const double tau = tau2 > 0.0 ? sqrt(tau2) : 0.0;
    const double delta = (double)qerr + (double)dberr;
    bound = tau - delta;
    return (dk + delta <= tau) ? 1 : 0;
}

inline rr_space make_rr(const snk_db *db, const snk_space &sp) {
    rr_space rs;
    rs.A = db->Jc_raw; rs.B = db->F_raw; rs.wA = db->wj; rs.wB = db->wt;
    rs.dA = sp.dA; rs.dB = sp.dB; rs.D = sp.D; rs.a_row_off = sp.a_row_off; rs.a_col = sp.a_col;
    rs.ldA = sp.ldA_raw; rs.ldB = sp.ldB_raw; rs.Dt = db->Dt; rs.m = sp.dB / db->Dt;
    return rs;
}

}  // namespace
