// fp32 direct-difference distance tiles on CUDA cores + warp top-k scan.
//
// This is the engine's exact-arithmetic shortlist path: sum_d (x_d - y_d)^2 is
// evaluated without the norm expansion, so it has full relative accuracy for
// near neighbours.  It serves (a) searches whose shape the tensor-core kernel
// does not take, (b) the re-search of queries whose tensor-core shortlist
// fails the exactness certificate (rerank.cu).
//
// Replaces (with rerank.cu): cKDTree.query -- reference script/synth_simple.py:490,
// script/synth_halfphone.py:1364.
#include "common.cuh"
#include <limits.h>
#include <algorithm>

namespace {

constexpr int TQ = 64;   // queries per CTA tile
constexpr int TR = 64;   // rows per CTA tile
constexpr int BK = 16;   // dims per smem stage
constexpr int PADW = 4;

struct space_dev {
    const float *A;   // weighted join matrix (may be null when dA == 0)
    const float *B;   // weighted target matrix
    int dA, D, a_row_off, a_col, ldA, ldB;
};

__device__ __forceinline__ float row_elem(const space_dev &sp, int64_t u, int d) {
    if (d < sp.dA) return __ldg(sp.A + (u + sp.a_row_off) * (int64_t)sp.ldA + sp.a_col + d);
    return __ldg(sp.B + u * (int64_t)sp.ldB + (d - sp.dA));
}

// out[q, r - row0] = sum_d (Q[q,d] - row[r,d])^2   (+inf for r >= row_end)
__global__ void __launch_bounds__(256)
dist_f32_kernel(space_dev sp, const float *__restrict__ Q, int ldq, int64_t nq, int64_t row0,
                int64_t row_end, float *__restrict__ out, int64_t ldo) {
    __shared__ float Qs[BK][TQ + PADW];
    __shared__ float Rs[BK][TR + PADW];
    const int tid = threadIdx.x;
    const int tr = tid % 16, tq = tid / 16;
    const int64_t q0 = (int64_t)blockIdx.y * TQ;
    const int64_t r0 = row0 + (int64_t)blockIdx.x * TR;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lk = tid % BK, lr = tid / BK;   // loader mapping: 16 dims x 16 rows per pass
    for (int k0 = 0; k0 < sp.D; k0 += BK) {
        const int d = k0 + lk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int qq = lr + 16 * i;
            const int64_t q = q0 + qq;
            float v = 0.f;
            if (q < nq && d < sp.D) v = __ldg(Q + q * (int64_t)ldq + d);
            Qs[lk][qq] = v;
            const int64_t r = r0 + qq;
            float w = 0.f;
            if (r < row_end && d < sp.D) w = row_elem(sp, r, d);
            Rs[lk][qq] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&Qs[k][tq * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Rs[k][tr * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float df = av[i] - bv[j];
                    acc[i][j] = fmaf(df, df, acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t q = q0 + tq * 4 + i;
        if (q >= nq) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r0 + tr * 4 + j;
            const int64_t col = r - row0;
            if (col < ldo) out[q * ldo + col] = (r < row_end) ? acc[i][j] : INFINITY;
        }
    }
}

// ---------------------------------------------------------------------------------
// Warp top-KP scan: one warp keeps the KP = 32*E smallest (value, id) pairs of a row
// segment in registers.  Ties: lower id wins.
__device__ __forceinline__ bool pair_lt(float v1, int i1, float v2, int i2) {
    return v1 < v2 || (v1 == v2 && i1 < i2);
}

// The list is kept SORTED ascending across the warp (position p = e * 32 + lane): inserting a pair is a
// lane-parallel shift -- every entry larger than the new pair takes its left neighbour (or the new pair at
// the insertion point) -- a handful of shuffles instead of a reduction per insertion.
template <int E>
struct warp_list {
    float lv[E];
    int li[E];
    float tv;   // current worst (largest) pair in the list = entry KP-1, warp-uniform
    int ti;

    __device__ __forceinline__ void refresh() {
        tv = __shfl_sync(0xffffffffu, lv[E - 1], 31);
        ti = __shfl_sync(0xffffffffu, li[E - 1], 31);
    }
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < E; ++e) { lv[e] = INFINITY; li[e] = INT_MAX; }
        tv = INFINITY; ti = INT_MAX;
    }
    // cv, ci warp-uniform
    __device__ __forceinline__ void offer(float cv, int ci) {
        if (!pair_lt(cv, ci, tv, ti)) return;
        const int lane = threadIdx.x & 31;
        float carry_v = -INFINITY;   // entry to the left of this row's lane 0 (last entry of the previous row)
        int carry_i = INT_MIN;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            float left_v = __shfl_up_sync(0xffffffffu, lv[e], 1);
            int left_i = __shfl_up_sync(0xffffffffu, li[e], 1);
            if (lane == 0) { left_v = carry_v; left_i = carry_i; }
            carry_v = __shfl_sync(0xffffffffu, lv[e], 31);      // before this row changes
            carry_i = __shfl_sync(0xffffffffu, li[e], 31);
            if (pair_lt(cv, ci, lv[e], li[e])) {                 // this entry moves right
                const bool left_moves = pair_lt(cv, ci, left_v, left_i);
                lv[e] = left_moves ? left_v : cv;
                li[e] = left_moves ? left_i : ci;
            }
        }
        refresh();
    }
    // per-lane candidate; all lanes must call
    __device__ __forceinline__ void offer_lanes(float v, int id) {
        unsigned mask = __ballot_sync(0xffffffffu, v < INFINITY && pair_lt(v, id, tv, ti));
        while (mask) {
            const int src = __ffs(mask) - 1;
            const float cv = __shfl_sync(0xffffffffu, v, src);
            const int ci = __shfl_sync(0xffffffffu, id, src);
            offer(cv, ci);
            mask &= mask - 1;
        }
    }
};

// grid (nsplit, nq), block 32.  vals [nq, ld] (first n valid); ids optional (same layout).
// out [nq, nsplit, KP].  If init == false and nsplit == 1 the existing out list is merged in.
// Segmented rows (seg_cnt != nullptr; ids required): the row is n / seg_cap buffers of seg_cap slots of which only
// the first seg_cnt[q, buffer] are filled (the emit epilogue of knn_tc.cu); the unfilled tails are never read.
template <int E>
__global__ void __launch_bounds__(32)
topk_scan_kernel(const float *__restrict__ vals, const int *__restrict__ ids, int64_t n, int64_t ld,
                 int id_base, bool init, float *__restrict__ oval, int *__restrict__ oid,
                 const int *__restrict__ seg_cnt, int seg_cap) {
    constexpr int KP = 32 * E;
    const int lane = threadIdx.x;
    const int64_t q = blockIdx.y;
    const int nsplit = gridDim.x, s = blockIdx.x;
    const int64_t len = (n + nsplit - 1) / nsplit;
    const int64_t beg = s * len, end = min(n, beg + len);
    warp_list<E> L;
    L.init();
    float *ov = oval + (q * nsplit + s) * KP;
    int *oi = oid + (q * nsplit + s) * KP;
    if (!init) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            L.lv[e] = ov[e * 32 + lane];
            const int id = oi[e * 32 + lane];
            L.li[e] = id < 0 ? INT_MAX : id;
        }
        L.refresh();   // lists written by this kernel are sorted ascending
    }
    const float *row = vals + q * ld;
    const int *irow = ids ? ids + q * ld : nullptr;
    auto consume = [&](int64_t from, int64_t to) {
        for (int64_t base = from; base < to; base += 128) {
            float v[4];
            int id[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t p = base + j * 32 + lane;
                v[j] = INFINITY;
                id[j] = INT_MAX;
                if (p < to) {
                    v[j] = __ldg(row + p);
                    id[j] = irow ? __ldg(irow + p) : (int)(id_base + p);
                    if (id[j] < 0) v[j] = INFINITY;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) L.offer_lanes(v[j], id[j]);
        }
    };
    if (seg_cnt) {
        const int nseg = (int)(n / seg_cap);
        const int per = (nseg + nsplit - 1) / nsplit;
        const int s1 = min(nseg, (s + 1) * per);
        for (int seg = s * per; seg < s1; ++seg) {
            const int64_t from = (int64_t)seg * seg_cap;
            consume(from, from + min(seg_cap, seg_cnt[q * nseg + seg]));
        }
    } else {
        consume(beg, end);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        ov[e * 32 + lane] = L.lv[e];
        oi[e * 32 + lane] = (L.lv[e] < INFINITY) ? L.li[e] : -1;
    }
}

template <int E>
void launch_scan(const float *vals, const int *ids, int64_t nq, int64_t n, int64_t ld, int id_base,
                 int nsplit, bool init, float *oval, int *oid, const int *seg_cnt, int seg_cap, cudaStream_t st) {
    dim3 grid(nsplit, (unsigned)nq);
    topk_scan_kernel<E><<<grid, 32, 0, st>>>(vals, ids, n, ld, id_base, init, oval, oid, seg_cnt, seg_cap);
}

int scan_dispatch(int KP, const float *vals, const int *ids, int64_t nq, int64_t n, int64_t ld, int id_base,
                  int nsplit, bool init, float *oval, int *oid, const int *seg_cnt, int seg_cap, cudaStream_t st) {
    switch (KP) {
    case 32: launch_scan<1>(vals, ids, nq, n, ld, id_base, nsplit, init, oval, oid, seg_cnt, seg_cap, st); break;
    case 64: launch_scan<2>(vals, ids, nq, n, ld, id_base, nsplit, init, oval, oid, seg_cnt, seg_cap, st); break;
    case 128: launch_scan<4>(vals, ids, nq, n, ld, id_base, nsplit, init, oval, oid, seg_cnt, seg_cap, st); break;
    case 256: launch_scan<8>(vals, ids, nq, n, ld, id_base, nsplit, init, oval, oid, seg_cnt, seg_cap, st); break;
    default: snk_set_error("shortlist size %d not supported (32/64/128/256)", KP); return 1;
    }
    SNK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int snk_topk_scan(snk_db *db, const float *d_vals, const int *d_ids, int64_t nq, int64_t n, int64_t ld,
                  int id_base, int KP, bool init, float *d_val, int *d_id, cudaStream_t st, const int *d_seg_cnt,
                  int seg_cap) {
    if (nq <= 0) return 0;
    SNK_CHECK(nq <= 65535, "topk scan: too many queries per launch (%lld)", (long long)nq);
    SNK_CHECK(!d_seg_cnt || (d_ids && seg_cap > 0 && n % seg_cap == 0), "topk scan: bad segment description");
    // split long rows over several warps when there are few queries
    int nsplit = 1;
    const int64_t want_warps = (int64_t)db->sm_count * 16;
    if (nq < want_warps && n > 4096) {
        nsplit = (int)std::min<int64_t>(snk_cdiv(want_warps, nq), snk_cdiv(n, 2048));
        if (d_seg_cnt) nsplit = (int)std::min<int64_t>(nsplit, n / seg_cap);
        if (nsplit < 1) nsplit = 1;
    }
    if (nsplit == 1) {
        db->counters[2] += 1;
        return scan_dispatch(KP, d_vals, d_ids, nq, n, ld, id_base, 1, init, d_val, d_id, d_seg_cnt, seg_cap, st);
    }
    // two-level: nsplit partial lists per query, then one more scan over them that also merges the
    // carried-in list (init == false)
    SNK_TRY(snk_buf_reserve(&db->ws_misc, (size_t)nq * nsplit * KP * 8));
    float *pv = (float *)db->ws_misc.p;
    int *pi = (int *)(pv + (size_t)nq * nsplit * KP);
    db->counters[2] += 2;
    SNK_TRY(scan_dispatch(KP, d_vals, d_ids, nq, n, ld, id_base, nsplit, true, pv, pi, d_seg_cnt, seg_cap, st));
    SNK_TRY(scan_dispatch(KP, pv, pi, nq, (int64_t)nsplit * KP, (int64_t)nsplit * KP, 0, 1, init, d_val, d_id, nullptr, 0,
                          st));
    return 0;
}

int snk_shortlist_simt(snk_db *db, const snk_space &sp, const float *dQ32, int ldq, int64_t nq, int KP,
                       float *d_val, int *d_id, cudaStream_t st) {
    if (nq <= 0) return 0;
    space_dev sd;
    sd.A = db->Jw32; sd.B = db->Fw32;
    sd.dA = sp.dA; sd.D = sp.D; sd.a_row_off = sp.a_row_off; sd.a_col = sp.a_col;
    sd.ldA = sp.ldA32; sd.ldB = sp.ldB32;
    // distance workspace: at most ~256 MiB, chunked over rows and over queries
    const size_t WS = (size_t)256 << 20;
    const int64_t qchunk = std::min<int64_t>(nq, 8192);
    int64_t rchunk = (int64_t)(WS / 4 / qchunk);
    rchunk = std::max<int64_t>(TR, rchunk / TR * TR);
    rchunk = std::min<int64_t>(rchunk, snk_round_up(sp.rows, TR));
    SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)qchunk * rchunk * 4));
    float *dist = (float *)db->ws_dist.p;
    for (int64_t qb = 0; qb < nq; qb += qchunk) {
        const int64_t qn = std::min<int64_t>(qchunk, nq - qb);
        bool first = true;
        for (int64_t rb = 0; rb < sp.rows; rb += rchunk) {
            const int64_t rn = std::min<int64_t>(rchunk, sp.rows - rb);
            dim3 grid((unsigned)snk_cdiv(rn, TR), (unsigned)snk_cdiv(qn, TQ));
            {
                snk_prof_scope prof(db, SNK_PROF_KNN, 2.0 * (double)qn * (double)rn * sp.D, st);
                dist_f32_kernel<<<grid, 256, 0, st>>>(sd, dQ32 + qb * ldq, ldq, qn, rb, rb + rn, dist, rchunk);
            }
            SNK_CUDA(cudaGetLastError());
            db->counters[2] += 1;
            SNK_TRY(snk_topk_scan(db, dist, nullptr, qn, rn, rchunk, (int)rb, KP, first, d_val + qb * KP,
                                  d_id + qb * KP, st));
            first = false;
        }
    }
    return 0;
}
