// Device-side recipe of a greedy step's queries, shared by the batched greedy loop (search.cu) and the
// single-utterance kernel (greedy_one.cu): where the previous choice and the target frames are, and the reference's
// float64 arithmetic that turns them into a query row (script/synth_simple.py:371-391, 467-470, 488, 501).
#pragma once
#include "common.cuh"

namespace {

// one query value -> fp16 operand element; accumulates the squared rounded value and the squared rounding error
__device__ __forceinline__ __half cvt_element(double x, float &n2, float &e2) {
    const __half h = __double2half(x);
    const float hf = __half2float(h);
    n2 = fmaf(hf, hf, n2);
    const float df = (float)(x - (double)hf);
    e2 = fmaf(df, df, e2);
    return h;
}

struct greedy_meta {
    int64_t tgt_off;    // first frame of the utterance in the concatenated targets
    int64_t path_off;   // first step of the utterance in the concatenated paths
    int64_t nsteps;
    int64_t start_state;
};

// standardise() then weight() of one un-normalised target value, in the reference's float64 arithmetic
// (data_manipulation.py:162-186: (x - mean) / std, unvoiced marker -> std * -1.0 * uv_scaling_factor;
// speech_manip.py:209-213: * weight).  x is the float32 value compose_speech produced.
// f32: the statistics are float32 (as read from the voice file, train_simple.py:94-97) and numpy keeps the
// whole standardisation in float32; otherwise float64 statistics promote it to float64.
struct std_params {
    const double *mean, *sd, *w;
    double uv_special, uv_scale;
    int f32;
};
__device__ __forceinline__ double standardise_weight(float x, int c, const std_params &sp) {
    double v;
    if (sp.f32) {
        const float sd = (float)sp.sd[c];
        v = (double)(x == (float)sp.uv_special ? __fmul_rn(__fmul_rn(sd, -1.0f), (float)sp.uv_scale)
                                               : __fdiv_rn(__fsub_rn(x, (float)sp.mean[c]), sd));
    } else {
        const double sd = sp.sd[c];
        v = (double)x == sp.uv_special ? __dmul_rn(__dmul_rn(sd, -1.0), sp.uv_scale)
                                       : __ddiv_rn(__dsub_rn((double)x, sp.mean[c]), sd);
    }
    return __dmul_rn(v, sp.w[c]);
}

// Query b of step t = [ prev_join_vector || m consecutive target frames ]   (synth_simple.py:467-470,488,501).
// The recipe: where the previous choices and the target frames are, and where finished steps go.
// targets: weighted float64 frames, or (targets32 != nullptr) un-normalised float32 frames that are
// standardised and weighted on the fly (synth_simple.py:371-391).
struct greedy_src {
    const greedy_meta *meta;
    int nact_prev;
    int64_t t;
    const double *targets;
    const float *targets32;
    std_params stp;
    int Dt, m;
    const float *Jc_raw;
    const double *wj;
    int Dj, Djq, prev_row_off, prev_col, cur_row_off, cur_col;
    const int64_t *ix_prev;
    const double *dist_prev;
    int64_t *paths;
    double *step_dist;
};

// the previous step's result of utterance b goes to its place in the output path
__device__ __forceinline__ void greedy_scatter(const greedy_src &g, const greedy_meta &mt, int64_t b) {
    if (g.t > 0 && b < g.nact_prev) {
        g.paths[mt.path_off + g.t - 1] = g.ix_prev[b];
        if (g.step_dist) g.step_dist[mt.path_off + g.t - 1] = g.dist_prev[b];
    }
}
// row / column of the join vector that precedes step t of utterance b (row < 0: none, zeros)
__device__ __forceinline__ void greedy_prev(const greedy_src &g, const greedy_meta &mt, int64_t b, int64_t &row, int &col) {
    row = -1;
    col = 0;
    if (g.t == 0) {
        if (mt.start_state >= 0) { row = mt.start_state + g.prev_row_off; col = g.prev_col; }
    } else {
        row = g.ix_prev[b] + g.cur_row_off;
        col = g.cur_col;
    }
}
__device__ __forceinline__ double greedy_value(const greedy_src &g, const greedy_meta &mt, int64_t row, int col, int d) {
    if (d < g.Djq) return row >= 0 ? (double)g.Jc_raw[row * g.Dj + col + d] * g.wj[col + d] : 0.0;
    const int64_t i = (mt.tgt_off + g.t * g.m) * g.Dt + (d - g.Djq);
    return g.targets32 ? standardise_weight(g.targets32[i], (d - g.Djq) % g.Dt, g.stp) : g.targets[i];
}

}  // namespace
