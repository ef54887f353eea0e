// Per-stream cost report along a greedy path.
// Replaces get_target_scores_per_stream / get_join_scores_per_stream /
// aggregate_squared_errors_by_stream (reference script/synth_halfphone.py:1964-1981,2977-3008):
// squared-error sums per stream, float64, no square root.  For multiepoch > 1 a step's target
// score sums the stream's columns over all m frames of the step (the reference's code only
// type-checks for m == 1; this is its natural extension and equals it there).
#include "common.cuh"

namespace {

__global__ void path_scores_kernel(const float *__restrict__ F_raw, const double *__restrict__ wt, int Dt, int m,
                                   const float *__restrict__ Jc_raw, const double *__restrict__ wj, int Dj,
                                   int prev_row_off, int prev_col, int cur_row_off, int cur_col,
                                   const double *__restrict__ targets, const int64_t *__restrict__ path, int64_t P,
                                   const int *__restrict__ tw, int nts, const int *__restrict__ jw, int njs,
                                   double *__restrict__ ts, double *__restrict__ js) {
    const int64_t p = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t u = path[p];
    for (int s = warp; s < nts + njs; s += nwarp) {
        double acc = 0.0;
        if (s < nts) {
            int c0 = 0;
            for (int i = 0; i < s; ++i) c0 += tw[i];
            const int w = tw[s];
            for (int e = lane; e < w * m; e += 32) {
                const int j = e / w, c = c0 + e % w;
                const double y = (double)F_raw[(u + j) * Dt + c] * wt[c];
                const double df = __dsub_rn(y, targets[(p * m + j) * Dt + c]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
        } else if (p + 1 < P) {
            const int sj = s - nts;
            int c0 = 0;
            for (int i = 0; i < sj; ++i) c0 += jw[i];
            const int w = jw[sj];
            const int64_t un = path[p + 1];
            for (int e = lane; e < w; e += 32) {
                const int c = c0 + e;
                const double a = (double)Jc_raw[(un + prev_row_off) * Dj + prev_col + c] * wj[prev_col + c];
                const double b = (double)Jc_raw[(u + cur_row_off) * Dj + cur_col + c] * wj[cur_col + c];
                const double df = __dsub_rn(a, b);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        if (lane == 0) {
            if (s < nts) ts[p * nts + s] = acc;
            else if (p + 1 < P) js[p * njs + (s - nts)] = acc;
        }
    }
}

}  // namespace

int snk_path_scores_dev(snk_db *db, const double *d_targets, int64_t T, const int64_t *d_path, int64_t P,
                        const int *d_tw, int nts, const int *d_jw, int njs, double *d_ts, double *d_js,
                        cudaStream_t st) {
    (void)T;
    if (P <= 0) return 0;
    path_scores_kernel<<<(unsigned)P, 128, 0, st>>>(db->F_raw, db->wt, db->Dt, db->m, db->Jc_raw, db->wj, db->Dj,
                                                    db->prev_row_off, db->prev_col, db->cur_row_off, db->cur_col,
                                                    d_targets, d_path, P, d_tw, nts, d_jw, njs, d_ts, d_js);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}
