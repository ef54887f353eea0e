// Re-weighting of the resident database: the GPU replacement for
// set_target_weights / set_join_weights / get_tree_for_greedy_search
// (reference script/synth_simple.py:234-274,190-230; script/synth_halfphone.py:682-737).
// The reference multiplies the float32 matrices by float64 weight vectors and rebuilds a
// KD-tree; here one pass over the raw matrices refreshes the f32 / fp16 operand copies,
// the squared norms and the fp16 perturbation bounds used by the exactness certificate.
#include "common.cuh"

namespace {

// out32[r, c] = f32(f64(raw[r, c]) * w[c]), zero padded to ld32
__global__ void weight_f32_kernel(const float *__restrict__ raw, const double *__restrict__ w, int64_t rows,
                                  int D, float *__restrict__ out, int ld) {
    const int64_t total = rows * ld;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld;
        const int c = (int)(i - r * ld);
        out[i] = c < D ? (float)((double)raw[r * D + c] * w[c]) : 0.f;
    }
}

// fp16 operand rows: out16[r, c] = half(f64(raw[r + row_off, col0 + c]) * w[col0 + c]) for c < Dv, else 0.
// Also per-row squared norm of the rounded values and squared rounding error.
// If embed: columns Dv, Dv+1, Dv+2 receive the squared norm split into three fp16 pieces (hi + mid + lo
// = the fp32 value to 33 bits).  A query that carries -0.5 in those columns makes the tensor-core
// accumulator x.y - 0.5 ||y||^2, so the kernel's key ||y||^2 - 2 x.y is just -2 * accumulator and no
// norm array has to be staged in the epilogue.
// One warp per row.
__global__ void weight_f16_kernel(const float *__restrict__ raw, const double *__restrict__ w, int64_t rows_out,
                                  int64_t rows_raw, int Draw, int row_off, int col0, int Dv,
                                  __half *__restrict__ out, int ld, float *__restrict__ nrm,
                                  float *__restrict__ err2, int embed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows_out; r += nwarp) {
        const int64_t rr = r + row_off;
        float n2 = 0.f, e2 = 0.f;
        for (int c = lane; c < ld; c += 32) {
            __half h = __float2half_rn(0.f);
            if (c < Dv && rr < rows_raw) {
                const double x = (double)raw[rr * Draw + col0 + c] * w[col0 + c];
                h = __double2half(x);
                const float hf = __half2float(h);
                n2 = fmaf(hf, hf, n2);
                const float df = (float)(x - (double)hf);
                e2 = fmaf(df, df, e2);
            }
            out[r * ld + c] = h;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            n2 += __shfl_xor_sync(0xffffffffu, n2, off);
            e2 += __shfl_xor_sync(0xffffffffu, e2, off);
        }
        if (lane == 0) {
            nrm[r] = n2;
            err2[r] = e2;
        }
        if (embed) {
            __syncwarp();
            if (lane == 0) {
                const __half hi = __float2half_rn(n2);
                const float r1 = n2 - __half2float(hi);
                const __half mid = __float2half_rn(r1);
                const float r2 = r1 - __half2float(mid);
                out[r * ld + Dv] = hi;
                out[r * ld + Dv + 1] = mid;
                out[r * ld + Dv + 2] = __float2half_rn(r2);
            }
        }
    }
}

// joint rows: nrm_j[u] = nA[u] + sum_{j<m} nB[u+j]; same for err2; track maxima
__global__ void joint_norm_kernel(const float *__restrict__ nA, const float *__restrict__ eA,
                                  const float *__restrict__ nB, const float *__restrict__ eB, int64_t Np, int m,
                                  float *__restrict__ nrm_j, float *__restrict__ max_n, float *__restrict__ max_e) {
    float mn = 0.f, me = 0.f;
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < Np; u += (int64_t)gridDim.x * blockDim.x) {
        float n = nA ? nA[u] : 0.f, e = eA ? eA[u] : 0.f;
        for (int j = 0; j < m; ++j) {
            n += nB[u + j];
            e += eB[u + j];
        }
        nrm_j[u] = n;
        mn = fmaxf(mn, n);
        me = fmaxf(me, e);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmaxf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<int *>(max_n), __float_as_int(mn));   // non-negative floats order as ints
        atomicMax(reinterpret_cast<int *>(max_e), __float_as_int(sqrtf(me)));
    }
}

}  // namespace

int snk_apply_weights(snk_db *db, cudaStream_t st) {
    const int blocks = db->sm_count * 8;
    weight_f32_kernel<<<blocks, 256, 0, st>>>(db->F_raw, db->wt, db->N, db->Dt, db->Fw32, db->Dt);
    SNK_CUDA(cudaGetLastError());
    weight_f32_kernel<<<blocks, 256, 0, st>>>(db->Jc_raw, db->wj, db->N + 1, db->Dj, db->Jw32, db->ldJ32);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 2;

    // fp16 operands for the tensor-core kernel
    SNK_TRY(snk_buf_reserve(&db->ws_misc, (size_t)(db->N + 1) * 4 * 4));
    float *nG = (float *)db->ws_misc.p, *eG = nG + (db->N + 1), *nS = eG + (db->N + 1), *eS = nS + (db->N + 1);
    weight_f16_kernel<<<blocks, 256, 0, st>>>(db->F_raw, db->wt, db->N, db->N, db->Dt, 0, 0, db->Dt, db->G16,
                                              db->ldG16, nG, eG, db->Dt + 3 <= db->ldG16);
    SNK_CUDA(cudaGetLastError());
    // S16 row u holds prev_join_rep[u] (row u + prev_row_off, columns prev_col ..)
    weight_f16_kernel<<<blocks, 256, 0, st>>>(db->Jc_raw, db->wj, db->N + 1, db->N + 1, db->Dj, db->prev_row_off,
                                              db->prev_col, db->Djq, db->S16, db->ldS16, nS, eS,
                                              db->Djq + 3 <= db->ldS16);
    SNK_CUDA(cudaGetLastError());
    SNK_CUDA(cudaMemsetAsync(db->maxn_t16, 0, 4, st));
    SNK_CUDA(cudaMemsetAsync(db->maxn_j16, 0, 4, st));
    SNK_CUDA(cudaMemsetAsync(db->err_t16, 0, 4, st));
    SNK_CUDA(cudaMemsetAsync(db->err_j16, 0, 4, st));
    joint_norm_kernel<<<blocks, 256, 0, st>>>(nullptr, nullptr, nG, eG, db->N, 1, db->nrm_t16, db->maxn_t16,
                                              db->err_t16);
    SNK_CUDA(cudaGetLastError());
    joint_norm_kernel<<<blocks, 256, 0, st>>>(nS, eS, nG, eG, db->Np, db->m, db->nrm_j16, db->maxn_j16,
                                              db->err_j16);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 4;
    return 0;
}
