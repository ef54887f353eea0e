// Join-cost tiles and batched Viterbi.
//
// join_tile_kernel : K x K tiles || end[a] - start[c] ||_2 between consecutive candidate sets.
//   Replaces get_natural_distance_vectorised over the pair lists built by
//   make_on_the_fly_join_lattice_BLOCK_DIRECT (reference script/synth_halfphone.py:2942-2951,
//   3206-3301), including its admissibility rules (:3238-3268).  Candidate rows are staged
//   into shared memory by the bulk async-copy engine (cp.async.bulk, one 608-byte row per
//   copy, completion on an mbarrier), the arithmetic is direct differences in packed fp32.
// viterbi_kernel   : min-plus relaxation over time with backpointers in HBM.
//   Replaces make_target_sausage_lattice + cost_cache_to_compiled_fst + compose +
//   shortestpath (reference script/fst_functions_wrapped.py:28-58,172-217,368,387-408): the
//   composed machine is a trellis whose arc t carries D[t,a] (+) join(a,c); weights are
//   float32 and are accumulated arc by arc as OpenFst's TropicalWeight<float> does.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int SUB = 5;              // each thread owns a SUB x SUB block of the tile
constexpr int SMEM_ROW_PAD = 4;     // floats; keeps 16-byte alignment and spreads banks

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// packed fp32 pair arithmetic (sm_100 FADD2/FFMA2): exact IEEE per element
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "sub.rn.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

// admissible units: synth_halfphone.py:3238-3240,3261-3268  (1 <= u < N-1; -1 is padding)
__device__ __forceinline__ bool admissible(int64_t u, int64_t N) { return u >= 1 && u < N - 1; }

// One CTA per tile.  tile2frame[i] = global frame index f of the tile's "first" candidate row
// (the "second" row is f + 1, same utterance).
__global__ void join_tile_kernel(const float *__restrict__ Jw, int ldJ, int64_t N, const int64_t *__restrict__ cand,
                                 int K, const int *__restrict__ tile2frame, float *__restrict__ tiles) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int nb = (K + SUB - 1) / SUB;
    const int KPAD = nb * SUB;
    const int sld = ldJ + SMEM_ROW_PAD;
    float *Es = reinterpret_cast<float *>(smraw);
    float *Ss = Es + (size_t)KPAD * sld;
    uint64_t *bar = reinterpret_cast<uint64_t *>(Ss + (size_t)KPAD * sld);
    int *okE = reinterpret_cast<int *>(bar + 1);
    int *okS = okE + KPAD;

    const int tile = blockIdx.x;
    const int64_t f = tile2frame[tile];
    const int tid = threadIdx.x;
    const uint32_t row_bytes = (uint32_t)ldJ * 4u;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the pad rows so they never feed NaNs into live lanes' registers
    for (int i = tid; i < (KPAD - K) * sld; i += blockDim.x) {
        Es[(size_t)K * sld + i] = 0.f;
        Ss[(size_t)K * sld + i] = 0.f;
    }
    int nvalid = 0;
    for (int j = tid; j < KPAD; j += blockDim.x) {
        const int64_t a = j < K ? cand[f * K + j] : -1;
        const int64_t c = j < K ? cand[(f + 1) * K + j] : -1;
        const int va = admissible(a, N), vc = admissible(c, N);
        okE[j] = va;
        okS[j] = vc;
        nvalid += va + vc;
        if (!va && j < K) for (int i = 0; i < ldJ; ++i) Es[(size_t)j * sld + i] = 0.f;
        if (!vc && j < K) for (int i = 0; i < ldJ; ++i) Ss[(size_t)j * sld + i] = 0.f;
    }
    // count valid rows block-wide
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (nvalid) atomicAdd(&s_cnt, nvalid);
    __syncthreads();
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)s_cnt * row_bytes);
    __syncthreads();
    for (int j = tid; j < K; j += blockDim.x) {
        if (okE[j]) {
            const int64_t a = cand[f * K + j];
            bulk_g2s(Es + (size_t)j * sld, Jw + (a + 1) * (int64_t)ldJ, row_bytes, bar);   // end[a] = Jw[a+1]
        }
        if (okS[j]) {
            const int64_t c = cand[(f + 1) * K + j];
            bulk_g2s(Ss + (size_t)j * sld, Jw + c * (int64_t)ldJ, row_bytes, bar);         // start[c] = Jw[c]
        }
    }
    mbar_wait(bar, 0);

    if (tid < nb * nb) {
        const int bi = tid / nb, bj = tid % nb;
        float2 acc[SUB][SUB];
#pragma unroll
        for (int i = 0; i < SUB; ++i)
#pragma unroll
            for (int j = 0; j < SUB; ++j) acc[i][j] = make_float2(0.f, 0.f);
        const float *e0 = Es + (size_t)(bi * SUB) * sld;
        const float *s0 = Ss + (size_t)(bj * SUB) * sld;
        for (int d = 0; d < ldJ; d += 4) {
            float4 ev[SUB], sv[SUB];
#pragma unroll
            for (int i = 0; i < SUB; ++i) ev[i] = *reinterpret_cast<const float4 *>(e0 + (size_t)i * sld + d);
#pragma unroll
            for (int j = 0; j < SUB; ++j) sv[j] = *reinterpret_cast<const float4 *>(s0 + (size_t)j * sld + d);
#pragma unroll
            for (int i = 0; i < SUB; ++i)
#pragma unroll
                for (int j = 0; j < SUB; ++j) {
                    const float2 d0 = sub2(make_float2(ev[i].x, ev[i].y), make_float2(sv[j].x, sv[j].y));
                    const float2 d1 = sub2(make_float2(ev[i].z, ev[i].w), make_float2(sv[j].z, sv[j].w));
                    acc[i][j] = fma2(d0, d0, acc[i][j]);
                    acc[i][j] = fma2(d1, d1, acc[i][j]);
                }
        }
        float *out = tiles + (size_t)tile * K * K;
#pragma unroll
        for (int i = 0; i < SUB; ++i) {
            const int a = bi * SUB + i;
            if (a >= K) continue;
#pragma unroll
            for (int j = 0; j < SUB; ++j) {
                const int c = bj * SUB + j;
                if (c >= K) continue;
                const float v = sqrtf(acc[i][j].x + acc[i][j].y);
                out[a * K + c] = (okE[a] && okS[c]) ? v : INFINITY;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
using vit_meta = snk_vit_meta;

// cp.async helpers (LDGSTS): tiles are prefetched into a shared-memory ring ahead of the DP front
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// One CTA per utterance, thread c < K owns "next" candidate c.  The K x K join tiles do not depend on
// the DP state, so they stream through a 3-deep cp.async ring two steps ahead of the relaxation; the
// chain over t only ever waits on shared memory.
template <int VIT_STAGES>
__global__ void viterbi_kernel(const vit_meta *__restrict__ meta, const int64_t *__restrict__ cand,
                               const double *__restrict__ tdist, const float *__restrict__ tiles, int K, int64_t N,
                               unsigned flags, short *__restrict__ bp, int64_t *__restrict__ paths,
                               int64_t *__restrict__ path_len, double *__restrict__ path_cost,
                               double *__restrict__ tcost, double *__restrict__ jcost, int bp_smem_frames) {
    extern __shared__ __align__(16) float vsm[];
    const int KK = K * K;
    const int KKp = (KK + 3) & ~3;
    const int Kp2 = (K + 1) & ~1;
    float *tile_s = vsm;                       // [VIT_STAGES][KKp]
    // per state a: {cost of reaching a (arcs 0..t-1 consumed), target cost D[t, a] as float32}, interleaved so
    // that the relaxation reads both (for two states) with one 16-byte broadcast load
    float2 *dcur = reinterpret_cast<float2 *>(vsm + VIT_STAGES * KKp);   // [Kp2]
    float2 *dnext = dcur + Kp2;                                          // [Kp2]
    // shared-memory mirror of the backpointers (the HBM copy is still written): the back-trace is a chain of
    // T dependent reads, far cheaper from shared memory.  Used when the utterance fits (bp_smem_frames).
    unsigned char *bp_s = reinterpret_cast<unsigned char *>(dnext + Kp2);       // [bp_smem_frames][K]
    int *col_s = reinterpret_cast<int *>(bp_s + (((size_t)bp_smem_frames * K + 3) & ~(size_t)3));   // [bp_smem_frames]
    __shared__ double s_red[2][32];
    __shared__ float s_best;
    __shared__ int s_arg;
    const vit_meta mt = meta[blockIdx.x];
    const int c = threadIdx.x, nthr = blockDim.x;
    const int64_t f0 = mt.frame_off;
    const int64_t T = mt.T;
    const bool beam1 = flags & 1u;
    const bool bp_local = T <= bp_smem_frames;

    auto fail = [&]() {
        if (c == 0) {
            path_len[blockIdx.x] = 0;
            path_cost[blockIdx.x] = INFINITY;
            if (tcost) tcost[blockIdx.x] = INFINITY;
            if (jcost) jcost[blockIdx.x] = INFINITY;
        }
        for (int64_t t = c; t < T; t += nthr) paths[f0 + t] = -1;
    };
    if (T < 2) { fail(); return; }   // empty J => empty composition (fst_functions_wrapped.py:172-217)

    const bool vec = (KK & 3) == 0;   // tile base addresses are 16-byte aligned iff K*K*4 is a multiple of 16
    auto prefetch = [&](int64_t t) {
        if (t < T - 1) {
            const float *src = tiles + (size_t)(mt.tile_off + t) * KK;
            float *dst = tile_s + (size_t)(t % VIT_STAGES) * KKp;
            if (vec) for (int i = c * 4; i < KK; i += nthr * 4) cp_async16(dst + i, src + i);
            else for (int i = c; i < KK; i += nthr) cp_async4(dst + i, src + i);
        }
        cp_async_commit();            // always commit so group counting stays uniform
    };
    for (int s = 0; s < VIT_STAGES - 1; ++s) prefetch(s);

    if (c < K) dcur[c].x = admissible(cand[f0 * K + c], N) ? 0.f : INFINITY;
    float d_next_row = c < K ? (float)tdist[f0 * K + c] : 0.f;
    for (int64_t t = 0; t < T - 1; ++t) {
        prefetch(t + VIT_STAGES - 1);
        if (c < K) {
            dcur[c].y = d_next_row;
            d_next_row = (float)tdist[(f0 + t + 1) * K + c];   // next step's target costs, a step early
        }
        cp_async_wait<VIT_STAGES - 1>();                        // tile t has landed (for this thread's copies)
        __syncthreads();                                        // ... and for everyone's; dcur visible
        if (c < K) {
            const float *tile = tile_s + (size_t)(t % VIT_STAGES) * KKp;
            float best = INFINITY;
            int arg = -1;
            int a = 0;
#pragma unroll 5
            for (; a + 1 < K; a += 2) {
                const float4 q = *reinterpret_cast<const float4 *>(dcur + a);   // {d[a], D[a], d[a+1], D[a+1]}
                const float arc0 = q.y + tile[a * K + c];       // Times(target arc, join arc)
                const float v0 = q.x + arc0;                    // Times(distance so far, arc)
                if (v0 < best) { best = v0; arg = a; }
                const float arc1 = q.w + tile[(a + 1) * K + c];
                const float v1 = q.z + arc1;
                if (v1 < best) { best = v1; arg = a + 1; }
            }
            if (a < K) {
                const float2 q = dcur[a];
                const float v = q.x + (q.y + tile[a * K + c]);
                if (v < best) { best = v; arg = a; }
            }
            dnext[c].x = best;
            bp[(f0 + t + 1) * K + c] = (short)arg;
            if (bp_local) bp_s[(t + 1) * K + c] = (unsigned char)(arg < 0 ? 255 : arg);
        }
        __syncthreads();
        if (beam1) {
            // greedy over candidates: only the cheapest state survives (lowest column on ties)
            if (c < 32) {
                float bv = INFINITY;
                int bi = -1;
                for (int a = c; a < K; a += 32)
                    if (dnext[a].x < bv) { bv = dnext[a].x; bi = a; }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ov < bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
                }
                if (c == 0) s_arg = bi;
            }
            __syncthreads();
            if (c < K && c != s_arg) dnext[c].x = INFINITY;
            __syncthreads();
        }
        float2 *tmp = dcur; dcur = dnext; dnext = tmp;
    }
    cp_async_wait<0>();
    // final arc: D[T-1, a] + 0 (exit arcs carry no weight, fst_functions_wrapped.py:206-208)
    if (c < K) dcur[c].y = d_next_row;
    __syncthreads();
    if (c < 32) {
        float bv = INFINITY;
        int bi = -1;
        for (int a = c; a < K; a += 32) {
            const float v = dcur[a].x + (dcur[a].y + 0.f);
            if (v < bv) { bv = v; bi = a; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov < bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
        }
        if (c == 0) { s_best = bv; s_arg = bi; }
    }
    __syncthreads();
    if (s_arg < 0) { fail(); return; }
    if (bp_local) {
        // walk the chain in shared memory, then gather units and costs with all threads
        if (c == 0) {
            int col = s_arg;
            for (int64_t t = T - 1; t >= 0; --t) {
                col_s[t] = col;
                if (t > 0) col = bp_s[t * K + col];
            }
        }
        __syncthreads();
        double tc = 0.0, jc = 0.0;
        for (int64_t t = c; t < T; t += nthr) {
            const int col = col_s[t];
            paths[f0 + t] = cand[(f0 + t) * K + col];
            tc += tdist[(f0 + t) * K + col];
            if (t > 0) jc += (double)tiles[(size_t)(mt.tile_off + t - 1) * K * K + col_s[t - 1] * K + col];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            tc += __shfl_xor_sync(0xffffffffu, tc, off);
            jc += __shfl_xor_sync(0xffffffffu, jc, off);
        }
        if ((c & 31) == 0) { s_red[0][c >> 5] = tc; s_red[1][c >> 5] = jc; }
        __syncthreads();
        if (c == 0) {
            double tcs = 0.0, jcs = 0.0;
            for (int w = 0; w < (nthr + 31) / 32; ++w) { tcs += s_red[0][w]; jcs += s_red[1][w]; }
            path_len[blockIdx.x] = T;
            path_cost[blockIdx.x] = (double)s_best;
            if (tcost) tcost[blockIdx.x] = tcs;
            if (jcost) jcost[blockIdx.x] = jcs;
        }
        return;
    }
    if (c == 0) {
        int col = s_arg;
        double tc = 0.0, jc = 0.0;
        for (int64_t t = T - 1; t >= 0; --t) {
            paths[f0 + t] = cand[(f0 + t) * K + col];
            tc += tdist[(f0 + t) * K + col];
            if (t > 0) {
                const int pcol = bp[(f0 + t) * K + col];
                jc += (double)tiles[(size_t)(mt.tile_off + t - 1) * K * K + pcol * K + col];
                col = pcol;
            }
        }
        path_len[blockIdx.x] = T;
        path_cost[blockIdx.x] = (double)s_best;
        if (tcost) tcost[blockIdx.x] = tc;
        if (jcost) jcost[blockIdx.x] = jc;
    }
}

// candidate target distances (synth_halfphone.py:1346-1351), float64, one warp per (t, j)
__global__ void cand_dist_kernel(const float *__restrict__ F_raw, const double *__restrict__ wt, int Dt, int64_t N,
                                 const int64_t *__restrict__ cand, const double *__restrict__ targets, int64_t T,
                                 int K, double *__restrict__ dist) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < T * K; p += nwarp) {
        const int64_t t = p / K;
        int64_t u = cand[p];
        if (u < 0) u += N;   // numpy negative indexing: -1 is the last unit
        if (u < 0 || u >= N) {   // numpy would raise IndexError; report NaN
            if (lane == 0) dist[p] = NAN;
            continue;
        }
        double acc = 0.0;
        for (int d = lane; d < Dt; d += 32) {
            const double y = (double)F_raw[u * Dt + d] * wt[d];
            const double df = __dsub_rn(y, targets[t * Dt + d]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        if (lane == 0) dist[p] = sqrt(acc);
    }
}

int build_meta(const int64_t *lens, int B, std::vector<vit_meta> &meta, std::vector<int> &tile2frame) {
    int64_t f = 0, tl = 0;
    meta.resize(B);
    for (int b = 0; b < B; ++b) {
        SNK_CHECK(lens[b] >= 0, "negative utterance length");
        meta[b].frame_off = f;
        meta[b].tile_off = tl;
        meta[b].T = lens[b];
        for (int64_t t = 0; t + 1 < lens[b]; ++t) tile2frame.push_back((int)(f + t));
        f += lens[b];
        tl += lens[b] > 0 ? lens[b] - 1 : 0;
    }
    return 0;
}

size_t tile_smem_bytes(int K, int ldJ) {
    const int nb = (K + SUB - 1) / SUB, KPAD = nb * SUB;
    return (size_t)2 * KPAD * (ldJ + SMEM_ROW_PAD) * 4 + 8 + (size_t)2 * KPAD * 4;
}

int launch_tiles(snk_db *db, const int64_t *d_cand, int K, const int *d_tile2frame, int64_t ntiles, float *d_tiles,
                 cudaStream_t st, unsigned long long *d_stats = nullptr) {
    if (ntiles <= 0) return 0;
    if (snk_join_tc_supported(db, K))   // n_candidates <= 64: centred norm expansion on the tensor cores (join_tc.cu)
        return snk_join_tc_launch(db, d_cand, K, d_tile2frame, ntiles, d_tiles, d_stats, st);
    const int nb = (K + SUB - 1) / SUB;
    const int threads = (int)snk_round_up(nb * nb, 32);
    SNK_CHECK(threads <= 1024, "n_candidates = %d too large for the join-tile kernel (max 160)", K);
    const size_t smem = tile_smem_bytes(K, db->ldJ32);
    SNK_CHECK(smem <= 227 * 1024, "n_candidates = %d needs %zu bytes of shared memory", K, smem);
    SNK_CUDA(cudaFuncSetAttribute(join_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {   // algorithmic bytes per tile: 2*K*Dj*4 gathered + K*K*4 written + 2*K*8 ids
        snk_prof_scope prof(db, SNK_PROF_JOIN, (double)ntiles * (2.0 * K * db->Dj * 4 + (double)K * K * 4 + 2.0 * K * 8), st);
        join_tile_kernel<<<(unsigned)ntiles, threads, smem, st>>>(db->Jw32, db->ldJ32, db->N, d_cand, K, d_tile2frame,
                                                                  d_tiles);
    }
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

}  // namespace

int snk_join_viterbi_batch_dev(snk_db *db, const int64_t *d_cand, const double *d_tdist, const int64_t *lens,
                               int B, int K, unsigned flags, int64_t *d_paths, int64_t *d_path_len,
                               double *d_path_cost, double *d_tcost, double *d_jcost, void *stream) {
    SNK_CHECK(db && db->weights_set, "snk_db_set_weights has not been called");
    SNK_LOCK(db);
    SNK_CHECK(K >= 1 && K <= 160, "n_candidates must be in [1, 160] (got %d)", K);
    SNK_CUDA(cudaSetDevice(db->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0) return 0;
    std::vector<vit_meta> meta;
    std::vector<int> t2f;
    SNK_TRY(build_meta(lens, B, meta, t2f));
    const int64_t ntiles = (int64_t)t2f.size();
    const int64_t nframes = meta[B - 1].frame_off + meta[B - 1].T;
    const size_t meta_bytes = snk_round_up(sizeof(vit_meta) * B, 256);
    const size_t t2f_bytes = snk_round_up(sizeof(int) * (size_t)std::max<int64_t>(ntiles, 1), 256);
    SNK_TRY(snk_buf_reserve(&db->ws_io, meta_bytes + t2f_bytes));
    SNK_TRY(snk_buf_reserve(&db->ws_bp, (size_t)std::max<int64_t>(nframes, 1) * K * 2));
    vit_meta *d_meta = (vit_meta *)db->ws_io.p;
    int *d_t2f = (int *)((char *)db->ws_io.p + meta_bytes);
    // launch metadata goes through the pinned staging ring: nothing here waits for the GPU
    SNK_TRY(snk_upload_async(db, d_meta, meta.data(), sizeof(vit_meta) * B, st));
    if (ntiles) SNK_TRY(snk_upload_async(db, d_t2f, t2f.data(), sizeof(int) * ntiles, st));
    SNK_TRY(snk_buf_reserve(&db->ws_tiles, (size_t)std::max<int64_t>(ntiles, 1) * K * K * 4));
    SNK_TRY(launch_tiles(db, d_cand, K, d_t2f, ntiles, (float *)db->ws_tiles.p, st));
    const int threads = (int)snk_round_up(K, 32);
    {   // per (utt, t): K*K*4 tile read + K*8 target costs + K*2 backpointers
        snk_prof_scope prof(db, SNK_PROF_VITERBI, (double)ntiles * ((double)K * K * 4 + K * 8.0 + K * 2.0), st);
        const size_t per_stage = (size_t)((K * K + 3) & ~3) * sizeof(float);
        int nst = 3;                                      // tiles in flight ahead of the DP front
        if (const char *e = getenv("SNK_VIT_STAGES")) nst = atoi(e);
        while (nst > 2 && nst * per_stage + 4 * ((K + 1) & ~1) * sizeof(float) > 200 * 1024) --nst;
        nst = std::max(2, std::min(nst, 5));
        // shared-memory backpointer mirror for utterances up to bp_frames frames (K <= 254: one byte each)
        int64_t maxT = 0;
        for (int b = 0; b < B; ++b) maxT = std::max(maxT, lens[b]);
        int bp_frames = 0;
        if (K <= 254 && maxT * K + maxT * 4 <= 24 * 1024 && !getenv("SNK_VIT_NOBPSMEM")) bp_frames = (int)maxT;
        size_t bp_bytes = bp_frames ? (((size_t)bp_frames * K + 3) & ~(size_t)3) + (size_t)bp_frames * 4 : 0;
        if (2 * per_stage + 4 * ((K + 1) & ~1) * sizeof(float) + bp_bytes > 224 * 1024) { bp_frames = 0; bp_bytes = 0; }   // widest tiles: no room
        // a third tile stage only if all B utterances stay resident in a single wave with it (measured:
        // a second wave costs far more than the deeper prefetch gains)
        if (!getenv("SNK_VIT_STAGES") && nst == 3) {
            const size_t s3 = 3 * per_stage + 4 * ((K + 1) & ~1) * sizeof(float) + bp_bytes + 1024;
            const int64_t resident = (int64_t)std::min<size_t>(32, (227 * 1024) / s3) * db->sm_count;
            if (resident < B) nst = 2;
        }
        const size_t vsmem = nst * per_stage + 4 * ((K + 1) & ~1) * sizeof(float) + bp_bytes;
#define SNK_LAUNCH_VIT(NST_)                                                                                              \
    do {                                                                                                                \
        SNK_CUDA(cudaFuncSetAttribute(viterbi_kernel<NST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));     \
        viterbi_kernel<NST_><<<B, threads, vsmem, st>>>(d_meta, d_cand, d_tdist, (const float *)db->ws_tiles.p, K, db->N,  \
                                                     flags, (short *)db->ws_bp.p, d_paths, d_path_len, d_path_cost,     \
                                                     d_tcost, d_jcost, bp_frames);                                      \
    } while (0)
        switch (nst) {
        case 2: SNK_LAUNCH_VIT(2); break;
        case 3: SNK_LAUNCH_VIT(3); break;
        case 4: SNK_LAUNCH_VIT(4); break;
        default: SNK_LAUNCH_VIT(5); break;
        }
#undef SNK_LAUNCH_VIT
    }
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}

int snk_join_tiles_dev(snk_db *db, const int64_t *d_cand, const int64_t *lens, int B, int K, float *d_tiles,
                       cudaStream_t st) {
    std::vector<vit_meta> meta;
    std::vector<int> t2f;
    SNK_TRY(build_meta(lens, B, meta, t2f));
    const int64_t ntiles = (int64_t)t2f.size();
    if (!ntiles) return 0;
    const size_t t2f_bytes = snk_round_up(sizeof(int) * ntiles, 16);
    SNK_TRY(snk_buf_reserve(&db->ws_io, t2f_bytes + 16));
    SNK_TRY(snk_upload_async(db, db->ws_io.p, t2f.data(), sizeof(int) * ntiles, st));
    unsigned long long *d_stats = (unsigned long long *)((char *)db->ws_io.p + t2f_bytes);
    SNK_CUDA(cudaMemsetAsync(d_stats, 0, 16, st));
    SNK_TRY(launch_tiles(db, d_cand, K, (const int *)db->ws_io.p, ntiles, d_tiles, st, d_stats));
    SNK_CUDA(cudaMemcpyAsync(db->jv_stats, d_stats, 16, cudaMemcpyDeviceToHost, st));   // read by snk_join_stats after the caller's sync
    return 0;
}

int snk_candidate_distances_dev(snk_db *db, const int64_t *d_cand, const double *d_targets, int64_t T, int K,
                                double *d_dist, cudaStream_t st) {
    if (T * K <= 0) return 0;
    cand_dist_kernel<<<db->sm_count * 4, 256, 0, st>>>(db->F_raw, db->wt, db->Dt, db->N, d_cand, d_targets, T, K,
                                                       d_dist);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}
