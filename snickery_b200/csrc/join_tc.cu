// Join-cost tiles on the tensor cores: K x K tiles || end[a] - start[c] ||_2 between consecutive candidate sets.
//
// Replaces get_natural_distance_vectorised over the pair lists of make_on_the_fly_join_lattice_BLOCK_DIRECT
// (reference script/synth_halfphone.py:2942-2951, 3206-3301, admissibility :3238-3268) for n_candidates <= 64; the
// direct-difference kernel of join_viterbi.cu (FP32-pipe bound at 0.28 of the HBM roofline) keeps the wider lattices.
//
// Arithmetic.  The candidates of two consecutive targets are acoustically alike (on the config-3 lattices half of all
// pairs have |e - s|^2 < 0.15 (|e|^2 + |s|^2)), so a plain norm expansion would cancel badly.  Per tile the kernel
//  * subtracts a common centre mu (the mean start row of the tile) from all 2K rows: x = e - mu, y = s - mu,
//    |e - s|^2 = |x|^2 + |y|^2 - 2 x.y, and now |x|^2 + |y|^2 is of the order of |e - s|^2 itself;
//  * splits the centred rows into fp16 hi/lo pairs (hi = fp16(s x), lo = fp16(s x - hi), s a power of two per voice) and
//    lets tcgen05 accumulate hi.hi + hi.lo + lo.hi in fp32 (the dropped lo.lo term is 2^-22 of |x||y|);
//  * decides natural joins by INDEX: end[a] and start[c] are the same row of the join matrix iff c == a + 1
//    (synth_halfphone.py:693-707), and cost exactly 0;
//  * recomputes the entries that still cancel, d^2 < theta (|x|^2 + |y|^2), theta = 3/32, from the raw float32 rows by
//    direct differences (exact subtraction of nearby float32 values; 1.3 % of the entries on the config-3 lattices).
//    The error of an entry that is NOT recomputed is that of the fp32 accumulation, <= 1.4e-6 (|x|^2 + |y|^2) on d^2 over
//    every entry measured, i.e. <= 1.4e-6 / (2 theta) = 7.5e-6 relative on the cost: inside the 1e-5 the costs are held to.
// Measured against the float64 formula: tests/test_gpu_join_tc.py and DESIGN.md section 4.4.
//
// Persistent CTAs (128 threads, seven per SM) take tiles round robin.  Per tile and 32-dim group: every thread loads
// eight float4 of the weighted join rows, the column means come from a shuffle + shared-memory reduction, the centred
// hi/lo halves go into the 128-byte-swizzled K-major layout UMMA reads ([hi 32 | lo 32] = one swizzle row per
// candidate), one thread issues the six UTCHMMAs of the group (M = 64 next candidates x N = 8 ceil(K/8) current
// candidates) while the loads of the next group are already in flight.  The accumulator comes back through tcgen05.ld
// into the same shared memory, is turned into costs, repaired where needed, and written out as the [a, c] tile the
// Viterbi kernel streams.  (A variant that also ran the min-plus relaxation in this kernel, one CTA per utterance, was
// measured at 2.3 ms against 3.4 ms for the old pair of kernels: with seven utterances per SM the chain
// load -> convert -> UTCHMMA -> tcgen05.ld of a single lattice step is latency bound.  Tiles are independent, the
// relaxation is not: this kernel keeps every SM busy with independent tiles and leaves the chain to viterbi_kernel.)
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <stdlib.h>

namespace {

using namespace snk_ptx;

constexpr int JT_THREADS = 128;
constexpr int JT_CTAS_PER_SM = 6;
constexpr int JT_M = 64;        // UMMA M: rows = "next" candidates; TMEM lanes 32 w + (0..15) hold rows 16 w + (0..15)
constexpr int JT_A_BYTES = JT_M * 128;
constexpr int JT_LDJ = 68;      // floats per row of the shared cost tile: bank = (4 a + c) mod 32
constexpr uint32_t JT_TMEM_COLS = 64;

struct jt_params {
    const float *Jw;         // [N + 1, ldJ] weighted join rows, float32, zero padded
    int ldJ;
    const float *sc;         // {s, 1 / s}: power-of-two operand scale of the current weighting
    const float *Jc_raw;     // [N + 1, Dj]
    const double *wj;        // [Dj]
    int Dj, G;
    int64_t N;
    const int64_t *cand;     // [frames, K]
    const int *tile2frame;   // [ntiles] frame of the tile's current candidates (the next ones are frame + 1)
    int64_t ntiles;
    int K, N8, HA;
    float theta;
    float *tiles;                   // [ntiles, K, K]
    unsigned long long *stats;      // optional {finite entries, entries recomputed by direct differences}
};

// generic-proxy writes (st.shared) -> visible to the async proxy the tensor core reads through
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {   // {lo, hi} -> one 32-bit word, round to nearest
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_half2(uint32_t h) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&h));
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

__device__ __forceinline__ bool admissible(int64_t u, int64_t N) { return u >= 1 && u < N - 1; }

// direct-difference join cost of join-matrix rows re (an end row) and rs (a start row) from the RAW float32 voice values:
// the subtraction of two nearby float32 numbers is exact, the weight multiplies the difference (one rounding) and the 151
// squares are summed in float32 over a shuffle tree: ~1e-7 relative however close the rows are -- and exactly 0 for equal
// rows.  One warp, all lanes return the cost.
__device__ __forceinline__ float direct_join(const jt_params &p, int re, int rs, int lane) {
    const float *e = p.Jc_raw + (size_t)re * p.Dj, *s = p.Jc_raw + (size_t)rs * p.Dj;
    float acc = 0.f;
    for (int d0 = 0; d0 < p.Dj; d0 += 160) {           // ten loads in flight per lane: one L2 round trip per 160 dims
        float ev[5], sv[5], wv[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int d = d0 + lane + 32 * i;
            const bool in = d < p.Dj;
            ev[i] = in ? __ldg(e + d) : 0.f;
            sv[i] = in ? __ldg(s + d) : 0.f;
            wv[i] = in ? (float)p.wj[d] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float x = (ev[i] - sv[i]) * wv[i];
            acc = fmaf(x, x, acc);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    return sqrtf(acc);
}

// centre, scale and split four consecutive dims of one row; store hi / lo into the swizzled operand row; return |x s|^2.
// addr = the thread's 8-byte slot in the hi half of the row; the lo half sits four 16-byte chunks on, and because the
// 128-byte swizzle is an XOR on the chunk index that is the same slot with address bit 6 flipped.
__device__ __forceinline__ float split_store(const float4 v, const float4 mus, float s, uint32_t addr) {
    const float x0 = fmaf(v.x, s, -mus.x), x1 = fmaf(v.y, s, -mus.y), x2 = fmaf(v.z, s, -mus.z), x3 = fmaf(v.w, s, -mus.w);
    const uint32_t h01 = pack_half2(x0, x1), h23 = pack_half2(x2, x3);
    const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
    const uint32_t l01 = pack_half2(x0 - f01.x, x1 - f01.y), l23 = pack_half2(x2 - f23.x, x3 - f23.y);
    sts64(addr, h01, h23);
    sts64(addr ^ 64u, l01, l23);
    return fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, x3 * x3)));
}

constexpr int JT_PLIST = 512;          // entries of a tile queued for direct recomputation (more: sentinel scan)

__global__ void __launch_bounds__(JT_THREADS, JT_CTAS_PER_SM) join_tile_tc_kernel(const jt_params p) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t sbase = smem_u32(sm);
    if (sbase & 1023u) __trap();
    const int K = p.K, N8 = p.N8, HA = p.HA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = max(JT_A_BYTES + N8 * 128, N8 * JT_LDJ * 4);
    float *Jbuf = reinterpret_cast<float *>(sm);                       // aliases the operand stage
    unsigned char *q = sm + ((stage_bytes + 127) & ~127);
    float4 *musum = reinterpret_cast<float4 *>(q); q += 4 * 8 * 16;    // per warp: column sums of the group (8 float4)
    float2 *ea = reinterpret_cast<float2 *>(q); q += 64 * 8;           // current candidate a: {|s (end row - mu)|^2, end row id as bits}
    float *nS = reinterpret_cast<float *>(q); q += 64 * 4;             // |s (start row - mu)|^2 of the next candidates
    int *ua = reinterpret_cast<int *>(q); q += 64 * 4;                 // join-matrix END row of current candidate a (unit + 1), -1 = inadmissible
    int *uc = reinterpret_cast<int *>(q); q += 64 * 4;                 // join-matrix START row of next candidate c (unit), -1 = inadmissible
    unsigned short *plist = reinterpret_cast<unsigned short *>(q); q += JT_PLIST * 2;
    uint64_t *bar_p = reinterpret_cast<uint64_t *>(q); q += 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(q); q += 4;
    int *s_cnt = reinterpret_cast<int *>(q);                           // [0], [1] valid next candidates per warp, [2] queued entries
    const uint32_t bar = smem_u32(bar_p);

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), JT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float scale = p.sc[0], inv_s = p.sc[1];
    uint32_t phase = 0;
    bool mma_pending = false;

    // cost role: thread pair (c, h): next candidate c, current candidates [a_lo, a_hi)
    const int c = 16 * warp + (lane & 15), h = lane >> 4;
    const int a_lo = h ? min(HA, K) : 0, a_hi = h ? K : min(HA, K);
    // conversion role: float4 j of a 32-dim group, rows rq + 16 i of the A (i < 4) and the B operand.  A warp's four rows
    // are {0, 1, 4, 5} + 2 (w & 1) + 8 (w >> 1), and each HALF warp holds two rows that differ in bit 2 ({0, 4} and {1, 5}):
    // an 8-byte store is served per half warp, and under the 128-byte swizzle rows r and r ^ 4 put their hi (and their lo)
    // halves into different halves of the banks -- two wavefronts per store instead of four (ncu: the stores were at twice
    // their ideal wavefront count, and the kernel at 74 % of the L1 wavefront peak).
    const int j = lane & 7;
    const int rq = 4 * ((lane >> 3) & 1) + (lane >> 4) + 2 * (warp & 1) + 8 * (warp >> 1);
    // 16-byte chunk ch of operand row r sits at chunk ch ^ (r & 7) (SWIZZLE_128B); r & 7 = rq & 7 for all four rows
    uint32_t st_base = sbase + rq * 128 + ((((uint32_t)(j >> 1)) ^ (uint32_t)(rq & 7)) << 4) + ((j & 1) << 3);
    st_base = __shfl_sync(0xffffffffu, st_base, lane);   // identity; ptxas otherwise recomputes the address from the thread id at all eight stores of a group
    unsigned n_fin = 0, n_patch = 0;

    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int64_t f = p.tile2frame[tile];
        if (tid < 64) {
            const int64_t ue = tid < K ? p.cand[f * K + tid] : -1;
            const int64_t us = tid < K ? p.cand[(f + 1) * K + tid] : -1;
            const bool oks = admissible(us, p.N);
            ua[tid] = admissible(ue, p.N) ? (int)ue + 1 : -1;
            uc[tid] = oks ? (int)us : -1;
            const unsigned m = __ballot_sync(0xffffffffu, oks);
            if (lane == 0) s_cnt[warp] = __popc(m);
            if (tid == 0) s_cnt[2] = 0;
        }
        __syncthreads();   // ua / uc visible; the previous tile has left Jbuf
        const float inv_na = 1.f / fmaxf((float)(s_cnt[0] + s_cnt[1]), 1.f);
        // element offsets of this thread's float4 in the four A (start) and four B (end) rows; absent rows read the zero
        // row that follows the join matrix (and are not stored)
        const uint32_t zoff = (uint32_t)(p.N + 1) * (uint32_t)p.ldJ;
        uint32_t offA[4], offB[4];
        unsigned valid = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ra = uc[rq + 16 * i], rb = ua[rq + 16 * i];
            offA[i] = ra >= 0 ? (uint32_t)ra * (uint32_t)p.ldJ + 4u * j : zoff;
            offB[i] = rb >= 0 ? (uint32_t)rb * (uint32_t)p.ldJ + 4u * j : zoff;
            valid |= (ra >= 0 ? 1u : 0u) << i | (rb >= 0 ? 16u : 0u) << i;
        }
        float nA[4] = {0.f, 0.f, 0.f, 0.f}, nB[4] = {0.f, 0.f, 0.f, 0.f};
        // ---- Gram tile D[c, a] = (start[c] - mu) . (end[a] - mu), one 32-dim group at a time through one operand stage
        float4 vA[4], vB[4];
        auto load_group = [&](int g) {
            const bool in = g * 32 + 4 * j < p.ldJ;            // the last group ends inside the (padded) row
            const float *base = p.Jw + g * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                vA[i] = __ldg(reinterpret_cast<const float4 *>(in ? base + offA[i] : p.Jw + zoff));
                vB[i] = __ldg(reinterpret_cast<const float4 *>(in ? base + offB[i] : p.Jw + zoff));
            }
        };
        load_group(0);
        for (int g = 0; g < p.G; ++g) {
            // column sums of the start rows: own rows, then the warp's four row groups, then the four warps
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sum.x += vA[i].x; sum.y += vA[i].y; sum.z += vA[i].z; sum.w += vA[i].w; }
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, off);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, off);
                sum.z += __shfl_xor_sync(0xffffffffu, sum.z, off);
                sum.w += __shfl_xor_sync(0xffffffffu, sum.w, off);
            }
            if (mma_pending) {            // the previous group's UTCHMMAs still read the stage
                mbar_wait(bar, phase);
                phase ^= 1;
                mma_pending = false;
            }
            if (lane < 8) musum[warp * 8 + lane] = sum;
            __syncthreads();
            float4 mus;
            {
                const float4 m0 = musum[j], m1 = musum[8 + j], m2 = musum[16 + j], m3 = musum[24 + j];
                const float k = inv_na * scale;
                mus = make_float4((m0.x + m1.x + m2.x + m3.x) * k, (m0.y + m1.y + m2.y + m3.y) * k,
                                  (m0.z + m1.z + m2.z + m3.z) * k, (m0.w + m1.w + m2.w + m3.w) * k);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (valid & (1u << i)) nA[i] += split_store(vA[i], mus, scale, st_base + i * 2048);
                if (valid & (16u << i)) nB[i] += split_store(vB[i], mus, scale, st_base + JT_A_BYTES + i * 2048);
            }
            if (g + 1 < p.G) load_group(g + 1);      // in flight while the tensor core works on this group
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                const uint32_t idesc = (1u << 4) | ((uint32_t)(N8 >> 3) << 17) | ((uint32_t)(JT_M >> 4) << 24);
                const uint64_t ad = DESC_HI | (uint64_t)((sbase >> 4) & 0x3FFFu);
                const uint64_t bd = DESC_HI | (uint64_t)(((sbase + JT_A_BYTES) >> 4) & 0x3FFFu);
                // 16-element K steps sit 32 B (= 2 descriptor units) apart: hi = steps 0, 1; lo = steps 2, 3
                umma_f16(tmem_base, ad + 0, bd + 0, idesc, g != 0);   // hi . hi
                umma_f16(tmem_base, ad + 2, bd + 2, idesc, 1u);
                umma_f16(tmem_base, ad + 0, bd + 4, idesc, 1u);       // hi . lo
                umma_f16(tmem_base, ad + 2, bd + 6, idesc, 1u);
                umma_f16(tmem_base, ad + 4, bd + 0, idesc, 1u);       // lo . hi
                umma_f16(tmem_base, ad + 6, bd + 2, idesc, 1u);
                umma_commit(bar);
            }
            mma_pending = true;
        }
        // squared norms of the centred, scaled rows: eight lanes per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int off = 1; off <= 4; off <<= 1) {
                nA[i] += __shfl_xor_sync(0xffffffffu, nA[i], off);
                nB[i] += __shfl_xor_sync(0xffffffffu, nB[i], off);
            }
            if (j == 0) {
                nS[rq + 16 * i] = nA[i];
                ea[rq + 16 * i] = make_float2(nB[i], __int_as_float(ua[rq + 16 * i]));
            }
        }
        mbar_wait(bar, phase);    // the accumulator is complete, the stage is free
        phase ^= 1;
        mma_pending = false;
        tc_fence_after();
        // ---- accumulator -> shared memory (lanes 0..15 of every warp hold rows 16 w + lane; N8 columns, Jbuf has N8 rows)
        for (int c0 = 0; c0 < N8; c0 += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 8; ++i) Jbuf[(c0 + i) * JT_LDJ + 16 * warp + lane] = v[i];
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- accumulators -> costs
        if (c < K) {
            const int rc = uc[c];
            float *col = Jbuf + c;
            if (rc < 0) {
                for (int a = a_lo; a < a_hi; ++a) col[a * JT_LDJ] = INFINITY;
            } else {
                const float nsc = nS[c];
                const float2 *e_p = ea + a_lo;
                float *col_p = col + a_lo * JT_LDJ;
#pragma unroll 4
                for (int a = a_lo; a < a_hi; ++a, ++e_p, col_p += JT_LDJ) {
                    const float2 e = *e_p;
                    const int ra = __float_as_int(e.y);
                    const float nn = e.x + nsc;
                    const float d2 = fmaf(-2.f, *col_p, nn);
                    float v = sqrt_approx(fmaxf(d2, 0.f)) * inv_s;
                    if (ra >= 0 && ra != rc && d2 < p.theta * nn) {   // the rows nearly coincide: queue for direct differences
                        const int slot = atomicAdd(&s_cnt[2], 1);
                        if (slot < JT_PLIST) plist[slot] = (unsigned short)((a << 8) | c);
                        v = -1.f;
                    }
                    v = ra == rc ? 0.f : v;                           // natural join: the same row of the join matrix
                    *col_p = ra < 0 ? INFINITY : v;
                    n_fin += ra >= 0;
                }
            }
        }
        __syncthreads();
        // ---- queued entries: direct differences of the raw rows, one warp per entry
        const int nq = s_cnt[2];
        if (nq > 0) {
            for (int i = warp; i < min(nq, JT_PLIST); i += JT_THREADS / 32) {
                const int a = plist[i] >> 8, cc = plist[i] & 255;
                const float d = direct_join(p, ua[a], uc[cc], lane);
                if (lane == 0) Jbuf[a * JT_LDJ + cc] = d;
            }
            if (nq > JT_PLIST) {      // more than the queue holds (a lattice full of duplicates): find the marked entries
                for (int i = warp; i < K * K; i += JT_THREADS / 32) {
                    const int a = i / K, cc = i - a * K;
                    if (Jbuf[a * JT_LDJ + cc] < 0.f) {
                        const float d = direct_join(p, ua[a], uc[cc], lane);
                        if (lane == 0) Jbuf[a * JT_LDJ + cc] = d;
                    }
                }
            }
            if (tid == 0) n_patch += (unsigned)nq;
            __syncthreads();
        }
        // ---- tile [a, c] -> HBM
        float *out = p.tiles + (size_t)tile * K * K;
        for (int a = warp; a < K; a += JT_THREADS / 32) {
            if (lane < K) out[a * K + lane] = Jbuf[a * JT_LDJ + lane];
            if (lane + 32 < K) out[a * K + lane + 32] = Jbuf[a * JT_LDJ + lane + 32];
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, JT_TMEM_COLS);
    if (p.stats) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            n_fin += __shfl_xor_sync(0xffffffffu, n_fin, off);
            n_patch += __shfl_xor_sync(0xffffffffu, n_patch, off);
        }
        if (lane == 0) {
            atomicAdd(p.stats, (unsigned long long)n_fin);
            atomicAdd(p.stats + 1, (unsigned long long)n_patch);
        }
    }
}

// ---------------------------------------------------------------- operand scale (once per re-weighting)
__global__ void absmax_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));   // non-negative floats order as ints
}

// sc[0] = s = the power of two that puts the largest |weighted value| in [2^13, 2^14) (a centred value is at most twice
// that: below the fp16 maximum); sc[1] = 1 / s
__global__ void scale_kernel(const float *__restrict__ absmax, float *__restrict__ sc) {
    int e = 0;
    const float m = *absmax;
    if (m > 0.f && isfinite(m)) frexpf(m, &e);     // m = f 2^e, f in [0.5, 1)
    int k = 14 - e;
    k = max(-60, min(60, k));
    sc[0] = ldexpf(1.f, k);
    sc[1] = ldexpf(1.f, -k);
}

int ensure_scale(snk_db *db, cudaStream_t st) {
    if (db->jsplit_valid) return 0;
    if (!db->split_sc) SNK_CUDA(cudaMalloc((void **)&db->split_sc, 4 * sizeof(float)));
    SNK_CUDA(cudaMemsetAsync(db->split_sc, 0, 4 * sizeof(float), st));
    absmax_kernel<<<db->sm_count * 8, 256, 0, st>>>(db->Jw32, (int64_t)(db->N + 1) * db->ldJ32, db->split_sc + 2);
    SNK_CUDA(cudaGetLastError());
    scale_kernel<<<1, 1, 0, st>>>(db->split_sc + 2, db->split_sc);
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 2;
    db->jsplit_valid = true;
    return 0;
}

float join_theta() {
    if (const char *e = getenv("SNK_JOIN_THETA")) return (float)atof(e);
    return 0.09375f;
}

}  // namespace

bool snk_join_tc_supported(const snk_db *db, int K) {
    return K >= 1 && K <= JT_M && db->N + 1 < (int64_t)1 << 31 && !getenv("SNK_JOIN_NOTC");
}

void snk_join_tc_free(snk_db *db) {
    cudaFree(db->split_sc);
    db->split_sc = nullptr;
}

// d_tile2frame [ntiles] device; d_tiles [ntiles, K, K]; d_stats optional {finite entries, recomputed entries}
int snk_join_tc_launch(snk_db *db, const int64_t *d_cand, int K, const int *d_tile2frame, int64_t ntiles, float *d_tiles,
                       unsigned long long *d_stats, cudaStream_t st) {
    if (ntiles <= 0) return 0;
    SNK_TRY(ensure_scale(db, st));
    jt_params p;
    p.Jw = db->Jw32; p.ldJ = db->ldJ32; p.sc = db->split_sc;
    p.Jc_raw = db->Jc_raw; p.wj = db->wj; p.Dj = db->Dj; p.G = (db->Dj + 31) / 32;
    p.N = db->N; p.cand = d_cand; p.tile2frame = d_tile2frame; p.ntiles = ntiles;
    p.K = K; p.N8 = (int)snk_round_up(K, 8);
    int ha = 4;
    while (ha < (K + 1) / 2) ha += 8;      // 4 * HA = 16 (mod 32): the two halves of a warp hit disjoint banks
    p.HA = ha;
    p.theta = join_theta();
    p.tiles = d_tiles; p.stats = d_stats;
    const size_t stage = (size_t)std::max(JT_A_BYTES + p.N8 * 128, p.N8 * JT_LDJ * 4);
    const size_t smem = ((stage + 127) & ~(size_t)127) + 4 * 8 * 16 + 64 * 8 + 3 * 64 * 4 + JT_PLIST * 2 + 32;
    SNK_CUDA(cudaFuncSetAttribute(join_tile_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)db->sm_count * JT_CTAS_PER_SM);
    {   // algorithmic bytes per tile: 2*K*Dj*4 gathered + K*K*4 written + 2*K*8 ids
        snk_prof_scope prof(db, SNK_PROF_JOIN, (double)ntiles * (2.0 * K * db->Dj * 4 + (double)K * K * 4 + 2.0 * K * 8), st);
        join_tile_tc_kernel<<<grid, JT_THREADS, smem, st>>>(p);
    }
    SNK_CUDA(cudaGetLastError());
    db->counters[2] += 1;
    return 0;
}
