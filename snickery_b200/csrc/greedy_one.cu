// Greedy joint search of ONE utterance as ONE persistent kernel.
//
// This is the literal call the reference makes, Synthesiser.greedy_joint_search(unit_features) for a single utterance
// (script/synth_simple.py:413, 458-503): a chain of dependent nearest-neighbour searches, one per multiepoch window.  With one
// query per step the search is a matrix-VECTOR product: every step streams the fp16 operand rows of the whole database
// (S16 join contexts + G16 target frames, weights.cu) from HBM once, and that stream is the floor of the step.  The batched
// path (search.cu) spends three launches per step and multiplies a 128-row tcgen05 query tile that holds one query; here one
// cooperative launch runs the whole utterance, one CTA per SM:
//
//   per step   scan      every warp walks a contiguous run of 16-row tiles of its CTA's row slice.  The keys
//                        ||y~||^2 - 2 x~.y~ come from mma.sync.m16n8k16 (fp16 x fp16 -> fp32, database rows are the M dimension)
//                        on the SAME operand rows, embedded norm pieces included, as the tcgen05 kernel: 34 K = 16 steps per
//                        key, so the certificate's error model (knn_tc.cu::snk_tc_eps_rel) carries over.  Operand rows go
//                        from global memory straight into the A fragments: the K order of a dot product is free, so lane
//                        (g, t) takes the 16 bytes at offset 16 t of each 64-byte chunk of row g / g + 8 and the query
//                        fragment is permuted to match.  The N dimension carries the m frames of the query window: one pass
//                        over a frame row yields its products with all m window positions, and a row's key is the join
//                        product plus a diagonal sum over m consecutive frame rows (shuffles).  That is what makes the
//                        scan HBM-bound: reading every frame row m times (from L1) left it bound by L1 wavefronts -- eight
//                        128-byte lines per load instruction -- at 0.74 of the HBM rate (ncu, profiles/).
//                        Join rows are requested one tile ahead (L1 bypassed), frame rows two.  Two smallest keys per thread.
//              merge     per CTA the 8 smallest of its threads' lists + tau (no dropped row has a smaller key) -> global.
//              barrier   one grid-wide arrive/wait (monotonic counter, two list parities); meanwhile the frame part of the
//                        next query, which does not depend on this step's answer, is assembled.
//              re-rank   EVERY CTA merges the CTA lists (a list whose head is not among the 16 smallest heads holds none of
//                        the 16 smallest entries), recomputes the 16 best rows in the reference's float64 arithmetic
//                        (rerank_dev.cuh; one row per warp, all loads of a row in flight at once), judges the certificate
//                        and assembles the join part of the next query from the chosen row -- the same deterministic code
//                        on the same data, so no second barrier and no broadcast is needed.
//
// An uncertified step clears the utterance's flag; snk_greedy_batch_finish then repeats the utterance with the next
// engine of the chain, exactly as for a batch.
#include "greedy_dev.cuh"
#include "rerank_dev.cuh"
#include <algorithm>

namespace {

constexpr int G1_THREADS = 512, G1_NW = G1_THREADS / 32;   // 16 warps at <= 128 registers: the scan lives on loads in flight
constexpr int G1_LC = 8;       // entries a CTA contributes per step
constexpr int G1_KP = 16;      // rows re-ranked in float64 per step
constexpr int LDS_H = 192, LDG_H = 64;   // halfs per operand row of the shapes the kernel is instantiated for (snk_greedy_one_supported)
constexpr int G1_ENT = G1_NW * 8 * 2;   // thread-list entries of a CTA: 8 result lanes per warp, two entries each

struct g1_params {
    greedy_src g;
    greedy_meta meta;
    rr_space rs;
    const __half *S16, *G16;
    int ldS, ldG;                 // operand row lengths in halfs
    const short *qmap;            // operand column -> query dim (knn_tc.cu)
    int ld16;
    int64_t Np;
    int64_t ngroups, groups_per_cta;
    unsigned long long *lists;    // [2][grid][G1_LC] packed (key, row)
    float *taus;                  // [2][grid]
    unsigned *durs;               // [2][grid] nanoseconds every CTA's scan took (row-slice balancing)
    int balance;                  // 1: the CTAs' row slices follow their measured scan rates
    unsigned *bar;                // arrive counter, zero at launch
    const float *dberr, *maxn;
    float eps_rel;
    int debug_fail_mod;
    int *flags, *count;
    float *dbg_keys;              // test instrumentation: keys of step 0, [Np]
    // database-sharded search: this rank scans joint rows [id_offset, id_offset + Np); after
    // the local re-rank the ranks exchange (distance, global row, bound) through their IPC-mapped regions (comm.cu)
    char *const *peers;           // peers[r] = rank r's exchange region as mapped here; nullptr: single GPU
    int rank, R;
    unsigned epoch0;              // epoch of step s is epoch0 + 1 + s
    int64_t id_offset;
    size_t xflags_bytes, xslot_bytes;   // layout of a region: flags, then [parity][writer rank] slots of xslot_bytes
    unsigned long long *dbg_cta;     // SNK_G1_TIMING: every CTA's globaltimer when its scan ends, [steps <= 256][grid]
    unsigned long long *dbg_times;   // SNK_G1_TIMING: CTA 0's globaltimer at up to 16 points of every step, [steps][16]
};

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {     // join contexts: read once per step, L1 bypassed
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_keep(const uint4 *p) {       // frames: m reads per step, L1 serves the repeats
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void mma16816(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// one 32-column chunk: a / b = the 16 bytes this lane holds of rows g / g + 8, q = the matching query halves
__device__ __forceinline__ void chunk_mma(float (&c)[4], const uint4 &a, const uint4 &b, const uint4 &q) {
    mma16816(c, a.x, b.x, a.y, b.y, q.x, q.y);
    mma16816(c, a.z, b.z, a.w, b.w, q.z, q.w);
}
// query halves from shared memory; volatile so that the 4 x NCH registers of a hoisted copy are not spent on it
__device__ __forceinline__ uint4 ld_query(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Squared float64 distance of ONE row to the query: the arithmetic of rerank_dev.cuh::rows_dist (per lane the dims lane,
// lane + 32, ... in order, join part then frames, one fma chain; xor tree), so the bits equal every other engine's -- but all
// of the row's loads are issued before the first is consumed: one memory round trip instead of one per 128 dims.
// NA, NB: compile-time bounds of dA / 32, dB / 32.
template <int NA, int NB>
__device__ __forceinline__ double row_dist_one(const rr_space &sp, const double *__restrict__ q_s, const double *__restrict__ wA_s,
                                               const double *__restrict__ wB_s, int u, int lane) {
    const float *ra = sp.A + ((int64_t)u + sp.a_row_off) * sp.ldA + sp.a_col;
    const float *rb = sp.B + (int64_t)u * sp.ldB;
    float ya[NA], yb[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) ya[i] = lane + 32 * i < sp.dA ? __ldg(ra + lane + 32 * i) : 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) yb[i] = lane + 32 * i < sp.dB ? __ldg(rb + lane + 32 * i) : 0.f;
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const int d = lane + 32 * i;
        if (d < sp.dA) {
            const double e = __dsub_rn(q_s[d], __dmul_rn((double)ya[i], wA_s[d]));
            a = __fma_rn(e, e, a);
        }
    }
    const double *qb = q_s + sp.dA;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int d = lane + 32 * i;
        if (d < sp.dB) {
            const double e = __dsub_rn(qb[d], __dmul_rn((double)yb[i], wB_s[d]));
            a = __fma_rn(e, e, a);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a = __dadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
    return a;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int M, int NCS, int NCG>
__global__ void __launch_bounds__(G1_THREADS, 1) greedy_one_kernel(const g1_params p) {
    extern __shared__ __align__(16) unsigned char g1_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int grid = gridDim.x, cta = blockIdx.x;
    const int nmerge = grid * G1_LC;
    const rr_space &rs = p.rs;

    // ---- shared memory carve
    double *q_s = reinterpret_cast<double *>(g1_smem);         // [2][D]: the query being searched and the next one
    double *wA_s = q_s + 2 * rs.D;                              // [dA]
    double *wB_s = wA_s + rs.dA;                                // [dB]
    double *d2 = wB_s + rs.dB;                                  // [KP]
    unsigned long long *mkey = reinterpret_cast<unsigned long long *>(d2 + G1_KP);   // [nmerge]
    unsigned long long *wk = mkey + nmerge;                     // [NW * KP]
    unsigned long long *ent = wk + G1_NW * G1_KP;               // [G1_ENT]
    unsigned long long *clist = ent + G1_ENT;                   // [LC]
    int *ids = reinterpret_cast<int *>(clist + G1_LC);          // [KP]
    float *sval = reinterpret_cast<float *>(ids + G1_KP);       // [KP]
    float *red = sval + G1_KP;                                  // [3][NW]
    __half *q16_s = reinterpret_cast<__half *>(red + 3 * G1_NW + 2);   // [ld16] (16-byte aligned: see g1_smem_bytes)
    q16_s = reinterpret_cast<__half *>((reinterpret_cast<uintptr_t>(q16_s) + 15) & ~(uintptr_t)15);
    short *qmap_s = reinterpret_cast<short *>(q16_s + p.ld16 + 8);  // [ld16]; q16_s[ld16 .. ld16 + 8) stays zero
    float *speed = reinterpret_cast<float *>(qmap_s + p.ld16);      // [grid] tiles per nanosecond of every CTA's scan, smoothed
    int *bounds = reinterpret_cast<int *>(speed + grid);            // [grid + 1] first tile of every CTA's slice
    __shared__ unsigned s_dur;
    __shared__ int64_t s_ix;
    __shared__ float s_qn, s_qerr;
    __shared__ double s_best;
    __shared__ float s_mx;
    __shared__ int s_sel[G1_KP];

    for (int d = tid; d < rs.dA; d += G1_THREADS) wA_s[d] = rs.wA[rs.a_col + d];
    for (int d = tid; d < rs.dB; d += G1_THREADS) wB_s[d] = rs.wB[d % rs.Dt];
    for (int c = tid; c < p.ld16; c += G1_THREADS) {
        const short d = p.qmap[c];
        qmap_s[c] = d;
        q16_s[c] = __float2half_rn(d == -2 ? -0.5f : 0.f);      // -0.5 multiplies the norm pieces embedded in the rows
    }
    if (tid < 8) q16_s[p.ld16 + tid] = __float2half_rn(0.f);
    __syncthreads();

    greedy_src gl = p.g;
    const greedy_meta mt = p.meta;
    const int64_t nsteps = mt.nsteps;
    float n2 = 0.f, e2 = 0.f;      // this thread's share of ||x~||^2 and ||x - x~||^2 of the query being assembled

    // query columns whose dim lies in [d_lo, d_hi): float64 value, fp16 operand element, norm shares
    auto fill = [&](double *q_dst, int64_t row, int col, int d_lo, int d_hi) {
        for (int c = tid; c < p.ld16; c += G1_THREADS) {
            const int d = qmap_s[c];
            if (d < d_lo || d >= d_hi) continue;
            const double x = greedy_value(gl, mt, row, col, d);
            q_dst[d] = x;
            q16_s[c] = cvt_element(x, n2, e2);
        }
    };
    // block-wide sums of the norm shares -> s_qn, s_qerr; ends with a barrier (q_s / q16_s complete as well)
    auto finish_query = [&]() {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            n2 += __shfl_xor_sync(0xffffffffu, n2, off);
            e2 += __shfl_xor_sync(0xffffffffu, e2, off);
        }
        if (lane == 0) { red[warp] = n2; red[G1_NW + warp] = e2; }
        __syncthreads();
        if (tid == 0) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < G1_NW; ++w) { a += red[w]; b += red[G1_NW + w]; }
            s_qn = a;
            s_qerr = sqrtf(b);
        }
        n2 = e2 = 0.f;
        __syncthreads();
    };

    // ---- query of step 0
    gl.t = 0;
    fill(q_s, -1, 0, gl.Djq, rs.D);
    {
        int64_t row = -1;
        int col = 0;
        if (mt.start_state >= 0) { row = mt.start_state + gl.prev_row_off; col = gl.prev_col; }
        fill(q_s, row, col, 0, gl.Djq);
    }
    finish_query();

    // row slices: equal at first; with p.balance they follow the scan rates measured in the previous steps (below)
    for (int c = tid; c <= grid; c += G1_THREADS) bounds[c] = (int)(p.ngroups * c / grid);
    for (int c = tid; c < grid; c += G1_THREADS) speed[c] = 0.f;
    if (tid == 0) s_dur = 0u;
    __syncthreads();
    const int64_t last = p.Np - 1, glast = p.Np + M - 2;      // last searchable row, last frame row
    const uint4 *S4 = reinterpret_cast<const uint4 *>(p.S16) + t4, *G4 = reinterpret_cast<const uint4 *>(p.G16) + t4;

    for (int64_t step = 0; step < nsteps; ++step) {
        const int par = (int)(step & 1);
        const double *q_cur = q_s + par * rs.D;
        double *q_nxt = q_s + (par ^ 1) * rs.D;
        const int64_t grp_lo = bounds[cta], grp_hi = bounds[cta + 1];
        // the warp's run of tiles: the CTA's slice in NW nearly equal parts
        const int64_t cta_groups = max((int64_t)0, grp_hi - grp_lo), w_base = cta_groups / G1_NW, w_rem = cta_groups % G1_NW;
        const int64_t w_lo = grp_lo + warp * w_base + min((int64_t)warp, w_rem), w_hi = w_lo + w_base + (warp < w_rem ? 1 : 0);
        const unsigned long long t_scan = p.balance ? global_ns() : 0ull;
        // ---- scan.  A warp walks a contiguous run of 16-row tiles.  Per tile k:
        //   J(k)   = join part, NCS chunks: every N column of the mma carries the same query, column 0 is read back;
        //   C(k+1) = frame products of the NEXT tile, NCG chunks: N column j carries frame j of the query window, so ONE pass
        //            over a frame row yields its products with all m window positions (C[f][j] = G[f] . x_j), instead of m
        //            passes over the row;
        //   key(u) = -2 (J[u] + sum_j C[u + j][j]): the diagonal sum takes rows of tile k and the first m - 1 rows of
        //            tile k + 1 from the lanes that hold them (shuffles), then the four lanes of a row add their columns.
        const unsigned qj_base = (unsigned)__cvta_generic_to_shared(q16_s + 8 * t4);
        const unsigned qt_base = (unsigned)__cvta_generic_to_shared(q16_s + LDS_H + min(g, M - 1) * LDG_H + 8 * t4);
        float k0 = INFINITY, k1 = INFINITY;
        int i0 = -1, i1 = -1;
        if (w_lo < w_hi) {
            // operand rows in 16-byte units: row pitch LDS_H / 8 (join contexts), LDG_H / 8 (frames), 4 units per 32-column chunk
            uint4 sA[NCS], sB[NCS], gA[NCG], gB[NCG];
            float cp[4];          // C(k): (row g, column 2 t4), (row g, 2 t4 + 1), (row g + 8, 2 t4), (row g + 8, 2 t4 + 1)
            {
                const int64_t ra = min(w_lo * 16 + g, last), rb = min(w_lo * 16 + g + 8, last);
                const uint4 *sa = S4 + ra * (LDS_H / 8), *sb = S4 + rb * (LDS_H / 8);
                const uint4 *ga = G4 + min(w_lo * 16 + g, glast) * (LDG_H / 8), *gb = G4 + min(w_lo * 16 + g + 8, glast) * (LDG_H / 8);
#pragma unroll
                for (int c = 0; c < NCS; ++c) { sA[c] = ld_stream(sa + 4 * c); sB[c] = ld_stream(sb + 4 * c); }
#pragma unroll
                for (int h = 0; h < NCG; ++h) { gA[h] = ld_keep(ga + 4 * h); gB[h] = ld_keep(gb + 4 * h); }
                cp[0] = cp[1] = cp[2] = cp[3] = 0.f;
                const int64_t na = min(w_lo * 16 + 16 + g, glast), nb = min(w_lo * 16 + 24 + g, glast);
                const uint4 *nga = G4 + na * (LDG_H / 8), *ngb = G4 + nb * (LDG_H / 8);
#pragma unroll
                for (int h = 0; h < NCG; ++h) {
                    chunk_mma(cp, gA[h], gB[h], ld_query(qt_base + 64 * h));
                    gA[h] = ld_keep(nga + 4 * h);
                    gB[h] = ld_keep(ngb + 4 * h);
                }
            }
            // shuffle sources of the diagonal sum: the lane that holds row g + j (mod 8) in the same column pair
            const int j0 = 2 * t4, j1 = 2 * t4 + 1;
            const int src0 = (((g + j0) & 7) << 2) | t4, src1 = (((g + j1) & 7) << 2) | t4;
            const bool lo0 = g + j0 < 8, lo1 = g + j1 < 8, on0 = j0 < M, on1 = j1 < M;
            for (int64_t grp = w_lo; grp < w_hi; ++grp) {
                // requested while this tile is multiplied: join rows of tile k + 1 (after the last tile: this one again,
                // in cache -- no branch around the loads), frame rows of tile k + 2
                const int64_t nxt = grp + 1 < w_hi ? grp + 1 : grp;
                const int64_t na = min(nxt * 16 + g, last), nb = min(nxt * 16 + g + 8, last);
                const uint4 *nsa = S4 + na * (LDS_H / 8), *nsb = S4 + nb * (LDS_H / 8);
                const int64_t fa = min(grp * 16 + 32 + g, glast), fb = min(grp * 16 + 40 + g, glast);
                const uint4 *nga = G4 + fa * (LDG_H / 8), *ngb = G4 + fb * (LDG_H / 8);
                float cn[4] = {0.f, 0.f, 0.f, 0.f}, cj[4] = {0.f, 0.f, 0.f, 0.f};
                uint4 qc = ld_query(qt_base);
#pragma unroll
                for (int h = 0; h < NCG; ++h) {
                    uint4 qn = qc;
                    if (h + 1 < NCG) qn = ld_query(qt_base + 64 * (h + 1));
                    else qn = ld_query(qj_base);
                    chunk_mma(cn, gA[h], gB[h], qc);
                    gA[h] = ld_keep(nga + 4 * h);
                    gB[h] = ld_keep(ngb + 4 * h);
                    qc = qn;
                }
#pragma unroll
                for (int c = 0; c < NCS; ++c) {
                    uint4 qn = qc;
                    if (c + 1 < NCS) qn = ld_query(qj_base + 64 * (c + 1));
                    chunk_mma(cj, sA[c], sB[c], qc);
                    sA[c] = ld_stream(nsa + 4 * c);
                    sB[c] = ld_stream(nsb + 4 * c);
                    qc = qn;
                }
                // diagonal sum.  Row g + j sits in the row-g half of C(k) while g + j < 8, else in its row-(g + 8) half; row
                // g + 8 + j in the row-(g + 8) half of C(k) while g + j < 8, else in the row-g half of C(k + 1).
                float ta, tb;
                {
                    const float a0 = __shfl_sync(0xffffffffu, cp[0], src0), b0 = __shfl_sync(0xffffffffu, cp[2], src0),
                                n0 = __shfl_sync(0xffffffffu, cn[0], src0);
                    const float a1 = __shfl_sync(0xffffffffu, cp[1], src1), b1 = __shfl_sync(0xffffffffu, cp[3], src1),
                                n1 = __shfl_sync(0xffffffffu, cn[1], src1);
                    const float ea = on0 ? (lo0 ? a0 : b0) : 0.f, eb = on0 ? (lo0 ? b0 : n0) : 0.f;
                    const float oa = on1 ? (lo1 ? a1 : b1) : 0.f, ob = on1 ? (lo1 ? b1 : n1) : 0.f;
                    ta = ea + oa;
                    tb = eb + ob;
                    ta += __shfl_xor_sync(0xffffffffu, ta, 1);
                    tb += __shfl_xor_sync(0xffffffffu, tb, 1);
                    ta += __shfl_xor_sync(0xffffffffu, ta, 2);
                    tb += __shfl_xor_sync(0xffffffffu, tb, 2);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) cp[i] = cn[i];
                if (t4 == 0) {
                    // key = ||y~||^2 - 2 x~.y~ = -2 * accumulator (the norms ride in the operand rows)
                    const int64_t r0 = grp * 16 + g, r1 = r0 + 8;
                    const float ka = r0 <= last ? -2.f * (cj[0] + ta) : INFINITY, kb = r1 <= last ? -2.f * (cj[2] + tb) : INFINITY;
                    if (p.dbg_keys && step == 0) {
                        if (cta == 0 && tid == 0) p.dbg_keys[p.Np] = s_qn;
                        if (r0 <= last) p.dbg_keys[r0] = ka;
                        if (r1 <= last) p.dbg_keys[r1] = kb;
                    }
                    if (ka < k1) {
                        if (ka < k0) { k1 = k0; i1 = i0; k0 = ka; i0 = (int)r0; }
                        else { k1 = ka; i1 = (int)r0; }
                    }
                    if (kb < k1) {
                        if (kb < k0) { k1 = k0; i1 = i0; k0 = kb; i0 = (int)r1; }
                        else { k1 = kb; i1 = (int)r1; }
                    }
                }
            }
        }

        if (p.balance && lane == 0) atomicMax(&s_dur, (unsigned)(global_ns() - t_scan));
        const bool timing = p.dbg_times && cta == 0 && tid == 0 && step < 4096;
        if (timing) p.dbg_times[step * 16 + 0] = global_ns();
        if (p.dbg_cta && tid == 0 && step < 256) p.dbg_cta[step * grid + cta] = global_ns();
        const bool has_next = step + 1 < nsteps;
        // ---- CTA merge: the LC smallest of the thread lists; tau = no row dropped so far has a smaller key
        {
            float tl = INFINITY;
            if (t4 == 0) {
                const int slot = (warp * 8 + g) * 2;
                ent[slot] = i0 >= 0 ? pack_key(k0, i0) : pad_key(slot);
                ent[slot + 1] = i1 >= 0 ? pack_key(k1, i1) : pad_key(slot + 1);
                tl = k1;                         // a thread drops only keys >= its second smallest (+inf while it has < 2 rows)
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) tl = fminf(tl, __shfl_xor_sync(0xffffffffu, tl, off));
            if (lane == 0) red[2 * G1_NW + warp] = tl;
            if (tid < G1_LC) clist[tid] = pad_key(tid);
            __syncwarp();
            // phase 1: the LC smallest of the warp's 16 entries (a dropped entry is >= the warp's LC-th, hence >= the CTA's)
            if (lane < 16) {
                const unsigned long long me = ent[warp * 16 + lane];
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) rank += ent[warp * 16 + j] < me ? 1 : 0;
                if (rank < G1_LC) wk[warp * G1_LC + rank] = me;
            }
            __syncthreads();
            if (timing) p.dbg_times[step * 16 + 1] = global_ns();
            // phase 2: the LC smallest of the warps' winners
            if (tid < G1_NW * G1_LC) {
                const unsigned long long me = wk[tid];
                int rank = 0;
#pragma unroll 8
                for (int j = 0; j < G1_NW * G1_LC; ++j) rank += wk[j] < me ? 1 : 0;
                if (rank < G1_LC) clist[rank] = me;
            }
            __syncthreads();
            if (timing) p.dbg_times[step * 16 + 9] = global_ns();
            if (tid < G1_LC) __stcg(p.lists + ((size_t)par * grid + cta) * G1_LC + tid, clist[tid]);
            if (tid == 0) {
                float tau = INFINITY;
                for (int w = 0; w < G1_NW; ++w) tau = fminf(tau, red[2 * G1_NW + w]);
                float v;
                int id;
                unpack_key(clist[G1_LC - 1], v, id);     // entries beyond the LC kept are >= the last kept (pad: +inf)
                tau = fminf(tau, v);
                __stcg(p.taus + (size_t)par * grid + cta, tau);
                if (p.balance) { __stcg(p.durs + (size_t)par * grid + cta, s_dur); s_dur = 0u; }
            }
            __syncthreads();
            if (tid == 0) {
                if (timing) p.dbg_times[step * 16 + 10] = global_ns();
                __threadfence();
                atomicAdd(p.bar, 1u);
                if (timing) p.dbg_times[step * 16 + 2] = global_ns();
            }
        }

        // ---- while the other CTAs finish: the target part of the next query does not depend on this step's answer
        if (has_next) {
            gl.t = step + 1;
            fill(q_nxt, -1, 0, gl.Djq, rs.D);
        }

        // ---- grid barrier: every CTA's list of this step has arrived
        if (tid == 0) {
            const unsigned want = (unsigned)(step + 1) * (unsigned)grid;
            const long long t_start = clock64();
            while ((int)(ld_acquire(p.bar) - want) < 0) {
                if (clock64() - t_start > 8000000000ll) __trap();   // seconds: a CTA died; fail instead of hanging the GPU
            }
            if (timing) p.dbg_times[step * 16 + 3] = global_ns();
        }
        __syncthreads();

        // ---- every CTA: merge the CTA lists -> KP rows, tau
        {
            float tl = INFINITY;
            for (int i = tid; i < grid; i += G1_THREADS) tl = fminf(tl, __ldcg(p.taus + (size_t)par * grid + i));
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) tl = fminf(tl, __shfl_xor_sync(0xffffffffu, tl, off));
            if (lane == 0) red[2 * G1_NW + warp] = tl;
            for (int i = tid; i < nmerge; i += G1_THREADS) {
                const unsigned long long e = __ldcg(p.lists + (size_t)par * nmerge + i);
                mkey[i] = (unsigned)(e >> 32) == 0xFFFFFFFFu ? pad_key(i) : e;     // pads made distinct across CTAs
            }
            if (tid < G1_KP) { sval[tid] = INFINITY; ids[tid] = INT_MAX; s_sel[tid] = -1; }
            // row-slice balancing: every CTA's scan rate of this step (tiles per nanosecond), smoothed over the steps
            if (p.balance) {
                for (int c = tid; c < grid; c += G1_THREADS) {
                    const unsigned d = __ldcg(p.durs + (size_t)par * grid + c);
                    const int gc = bounds[c + 1] - bounds[c];
                    if (d > 0u && gc > 0) {
                        const float r = (float)gc / (float)d;
                        speed[c] = speed[c] > 0.f ? 0.5f * speed[c] + 0.5f * r : r;
                    }
                }
            }
            __syncthreads();
            if (timing) p.dbg_times[step * 16 + 6] = global_ns();
            // A CTA's list is ascending, so a list whose head is not among the KP smallest heads holds none of the KP smallest
            // entries (KP smaller heads precede everything in it): rank the heads, then only the entries of those KP lists.
            // two threads per head, four per candidate entry: the counting loops are the serial part of the step
            for (int l0 = 0; l0 < grid; l0 += G1_THREADS / 2) {         // warp-uniform trip count (the shuffle below)
                const int l = l0 + (tid >> 1);
                const unsigned long long me = l < grid ? mkey[l * G1_LC] : 0ull;
                int rank = 0;
                if (l < grid) {
#pragma unroll 4
                    for (int j = tid & 1; j < grid; j += 2) rank += mkey[j * G1_LC] < me ? 1 : 0;
                }
                rank += __shfl_xor_sync(0xffffffffu, rank, 1);
                if (l < grid && (tid & 1) == 0 && rank < G1_KP) s_sel[rank] = l;
            }
            __syncthreads();
            if (tid < G1_KP * G1_LC) {
                const int l = s_sel[tid / G1_LC];
                wk[tid] = l >= 0 ? mkey[l * G1_LC + tid % G1_LC] : pad_key(nmerge + tid);
            }
            __syncthreads();
            static_assert(G1_KP * G1_LC * 4 == G1_THREADS, "four threads per candidate entry");
            {
                const unsigned long long me = wk[tid >> 2];
                int rank = 0;
#pragma unroll 8
                for (int j = tid & 3; j < G1_KP * G1_LC; j += 4) rank += wk[j] < me ? 1 : 0;
                rank += __shfl_xor_sync(0xffffffffu, rank, 1);
                rank += __shfl_xor_sync(0xffffffffu, rank, 2);
                if ((tid & 3) == 0 && rank < G1_KP) unpack_key(me, sval[rank], ids[rank]);
            }
            __syncthreads();
            if (timing) p.dbg_times[step * 16 + 7] = global_ns();
        }

        // ---- float64 distances of the KP rows, one per warp; the join context that would follow each of them is requested
        // into L2 meanwhile (one of them becomes the next query's join part)
        static_assert(G1_KP <= G1_NW && G1_KP * G1_LC <= G1_NW * G1_KP && G1_KP <= 32 && G1_NW <= 32, "one shortlisted row per warp; the winners fit wk");
        if (warp < G1_KP) {
            const int u = ids[warp];
            double r = INFINITY;
            if (u != INT_MAX) {
                if (has_next && lane * 32 < gl.Djq)
                    prefetch_l2_line(gl.Jc_raw + ((int64_t)u + gl.cur_row_off) * gl.Dj + gl.cur_col + lane * 32);
                r = row_dist_one<NCS, M * NCG>(rs, q_cur, wA_s, wB_s, u, lane);
            }
            if (lane == 0) d2[warp] = r;
        }
        __syncthreads();
        if (timing) p.dbg_times[step * 16 + 8] = global_ns();

        // ---- the nearest row (ties: lowest id): lexicographic minimum over the KP lanes of warp 0; the same warp works out
        // what the certificate needs (tau: no row outside the shortlist has a smaller key)
        if (warp == 0) {
            double v = lane < G1_KP ? d2[lane] : INFINITY;
            int i = lane < G1_KP ? ids[lane] : INT_MAX;
            float mx = lane < G1_KP && ids[lane] != INT_MAX ? sval[lane] : -INFINITY;
            const bool full = __all_sync(0xffffffffu, lane >= G1_KP || ids[lane] != INT_MAX);
            float tl = lane < G1_NW ? red[2 * G1_NW + lane] : INFINITY;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, off);
                const int oi = __shfl_xor_sync(0xffffffffu, i, off);
                if (dpair_lt(ov, oi, v, i)) { v = ov; i = oi; }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                tl = fminf(tl, __shfl_xor_sync(0xffffffffu, tl, off));
            }
            if (lane == 0) {
                s_ix = i != INT_MAX ? (int64_t)i + p.id_offset : -1;
                s_best = v;
                s_mx = fminf(full ? mx : INFINITY, tl);   // the final merge dropped nothing unless it was full; earlier stages may have
            }
        }
        __syncthreads();
        if (timing) p.dbg_times[step * 16 + 4] = global_ns();

        if (p.peers) {
            // ---- database-sharded: this rank's best is one candidate.  CTA 0 stores (distance, global row, bound) into its
            // slot in EVERY rank's region and publishes the step's epoch (system-scope release); every CTA of every rank
            // waits until all ranks' epochs have arrived in the local region and takes the arg-min.  The answer stands iff it
            // is not above any rank's bound (a shard without a close row cannot certify its own best, and need not).
            const unsigned epoch = p.epoch0 + 1u + (unsigned)step;
            const size_t slot0 = p.xflags_bytes + (size_t)(epoch & 1u) * p.R * p.xslot_bytes;
            if (cta == 0 && warp == 0) {
                double dk = INFINITY, bound = INFINITY;
                if (lane == 0) {
                    dk = s_ix >= 0 ? sqrt(s_best) : INFINITY;
                    if (s_mx < INFINITY) cert_fp16(s_mx, s_qn, *p.maxn, p.eps_rel, s_qerr, *p.dberr, dk, bound);
                    if (p.debug_fail_mod > 0) bound = -INFINITY;
                }
                dk = __shfl_sync(0xffffffffu, dk, 0);
                bound = __shfl_sync(0xffffffffu, bound, 0);
                if (lane < p.R) {
                    double *dst = reinterpret_cast<double *>(p.peers[lane] + slot0 + (size_t)p.rank * p.xslot_bytes);
                    dst[0] = dk;
                    reinterpret_cast<int64_t *>(dst)[1] = s_ix >= 0 ? s_ix : (int64_t)0x7fffffffffffffffll;
                    dst[2] = bound;
                    __threadfence_system();
                    st_release_sys(reinterpret_cast<unsigned *>(p.peers[lane]) + p.rank, epoch);
                }
            }
            if (warp == 1) {
                if (lane < p.R) {
                    const unsigned *f = reinterpret_cast<const unsigned *>(p.peers[p.rank]) + lane;
                    const long long t_start = clock64();
                    while ((int)(ld_acquire_sys(f) - epoch) < 0)
                        if (clock64() - t_start > 20000000000ll) __trap();   // a peer that never arrives must not hang the GPU
                }
                __syncwarp();
                double bd = INFINITY, lb = INFINITY;
                int64_t bi = 0x7fffffffffffffffll;
                if (lane < p.R) {      // written by remote GPUs: read past L1
                    const double *e = reinterpret_cast<const double *>(p.peers[p.rank] + slot0 + (size_t)lane * p.xslot_bytes);
                    bd = __ldcg(e);
                    bi = __ldcg(reinterpret_cast<const long long *>(e) + 1);
                    lb = __ldcg(e + 2);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    lb = fmin(lb, __shfl_xor_sync(0xffffffffu, lb, off));
                }
                if (lane == 0) {
                    s_ix = bi;
                    if (cta == 0) {
                        gl.paths[mt.path_off + step] = bi;
                        if (gl.step_dist) gl.step_dist[mt.path_off + step] = bd;
                        if (!(bd <= lb)) {
                            if (p.flags[0]) atomicAdd(p.count, 1);
                            p.flags[0] = 0;
                        }
                    }
                }
            }
            __syncthreads();
        } else if (warp == G1_NW - 1 && lane == 0) {
            // ---- single GPU: the last warp judges the certificate and writes the outputs while the others assemble the join
            // part of the next query (s_qn / s_qerr of THIS query are read before finish_query's barrier replaces them)
            const bool ok = s_ix >= 0;
            const double dk = ok ? sqrt(s_best) : INFINITY;
            int good = 1;
            double bound = INFINITY;
            if (s_mx < INFINITY) good = cert_fp16(s_mx, s_qn, *p.maxn, p.eps_rel, s_qerr, *p.dberr, dk, bound);
            if (p.debug_fail_mod > 0) good = 0;               // test hook (query 0 of a batch of one)
            if (cta == 0) {
                gl.paths[mt.path_off + step] = ok ? s_ix : p.Np;
                if (gl.step_dist) gl.step_dist[mt.path_off + step] = dk;
                if (!good) {
                    p.flags[0] = 0;
                    atomicAdd(p.count, 1);
                }
            }
        }
        // One warp re-cuts the slices in proportion to the rates (the same arithmetic on the same numbers in every CTA,
        // so all CTAs agree); the HBM rate an SM sees is tied to the SM (measured: the same few CTAs finish 3 us late
        // in every step), and the grid barrier waits for the slowest.  This warp is idle here (others assemble the next query /
        // judge the certificate); nobody reads `bounds` before the next scan starts.
        if (p.balance && has_next && warp == G1_NW - 2) {
            const int per = (grid + 31) / 32, c0 = lane * per, c1 = min(grid, c0 + per);
            float mine = 0.f, known = 0.f;
            int nknown = 0;
            for (int c = c0; c < c1; ++c)
                if (speed[c] > 0.f) { known += speed[c]; ++nknown; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                known += __shfl_xor_sync(0xffffffffu, known, off);
                nknown += __shfl_xor_sync(0xffffffffu, nknown, off);
            }
            if (nknown > 0) {
                const float fill_in = __fdividef(known, (float)nknown);   // CTAs without a measurement yet: the mean rate
                for (int c = c0; c < c1; ++c) mine += speed[c] > 0.f ? speed[c] : fill_in;
                float incl = mine;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const float total = __shfl_sync(0xffffffffu, incl, 31);
                const float scale = (float)p.ngroups / total;          // one division; the products below are monotone in cum
                float cum = incl - mine;
                for (int c = c0; c < c1; ++c) {
                    bounds[c] = min((int)p.ngroups, (int)(cum * scale));
                    cum += speed[c] > 0.f ? speed[c] : fill_in;
                }
                if (lane == 0) { bounds[0] = 0; bounds[grid] = (int)p.ngroups; }
            }
        }
        if (has_next) {
            fill(q_nxt, (s_ix >= 0 ? s_ix : p.Np) + gl.cur_row_off, gl.cur_col, 0, gl.Djq);
            finish_query();
        } else {
            __syncthreads();      // s_ix / s_best are rewritten by the next step only after every reader is done
        }
        if (timing) p.dbg_times[step * 16 + 5] = global_ns();
    }
}

size_t g1_smem_bytes(const rr_space &rs, int grid, int ld16) {
    return (size_t)(2 * rs.D + rs.dA + rs.dB + G1_KP) * 8 + (size_t)(grid * G1_LC + G1_NW * G1_KP + G1_ENT + G1_LC) * 8 +
           (size_t)G1_KP * 8 + (size_t)(3 * G1_NW + 2) * 4 + 16 + (size_t)ld16 * 4 + 16 + 16 + (size_t)(2 * grid + 1) * 4 + 8;
}

typedef void (*g1_fn)(const g1_params);
g1_fn g1_pick(int m) {
    switch (m) {
    case 1: return greedy_one_kernel<1, 5, 2>;
    case 2: return greedy_one_kernel<2, 5, 2>;
    case 3: return greedy_one_kernel<3, 5, 2>;
    case 4: return greedy_one_kernel<4, 5, 2>;
    case 5: return greedy_one_kernel<5, 5, 2>;
    case 6: return greedy_one_kernel<6, 5, 2>;
    case 7: return greedy_one_kernel<7, 5, 2>;     // the window's frames ride in the eight N columns of the mma (m <= 8); the
                                                   // query operand layout it shares with knn_tc.cu holds ten K-blocks (m <= 7)
    default: return nullptr;
    }
}

}  // namespace

// shapes the kernel is instantiated for: the shipped epoch voices (151-dim join contexts, 61-dim frames) whose operand rows
// carry their norms, multiepoch 1 - 7
bool snk_greedy_one_supported(const snk_db *db) {
    if (getenv("SNK_GREEDY_NO_ONE")) return false;
    if (!db->tc_ok || !db->tc_state || db->engine == SNK_ENGINE_SIMT || db->engine == SNK_ENGINE_EXACT) return false;
    if (db->ldS16 != 192 || db->Djq + 3 > 160 || db->Djq + 3 <= 128) return false;
    if (db->ldG16 != 64 || db->Dt + 3 > 64 || db->Dt + 3 <= 32) return false;
    if (db->Np < 1 || db->Np >= (int64_t)INT_MAX - 64) return false;
    g1_fn fn = g1_pick(db->m);
    if (!fn) return false;
    if (!snk_tc_supported(db, snk_make_space(db, SNK_SPACE_JOINT), G1_KP)) return false;   // the operand layout / query map it reads
    if (db->g1_resident == 0) {
        // once per handle: one CTA per SM must be resident at the same time (cooperative launch), and the device must allow it
        const snk_space sp = snk_make_space(db, SNK_SPACE_JOINT);
        const size_t smem = g1_smem_bytes(make_rr(db, sp), db->sm_count, db->ldS16 + db->m * db->ldG16);
        int coop = 0, nb = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, db->device);
        const bool ok = coop && cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)fn, G1_THREADS, smem) == cudaSuccess && nb >= 1;
        cudaGetLastError();
        const_cast<snk_db *>(db)->g1_resident = ok ? 1 : -1;
    }
    return db->g1_resident == 1;
}

// d_keys (optional, [Np] floats): the kernel also stores the keys of step 0 (certificate tests)
int snk_greedy_one_launch(snk_db *db, const void *meta_, const double *d_targets, const float *d_unnorm, int64_t *d_paths,
                          double *d_step_dist, int *d_flags, int *d_count, float *d_keys, cudaStream_t st,
                          const float *d_Jc_full, int64_t id_offset) {
    const greedy_meta &meta = *(const greedy_meta *)meta_;
    const snk_space sp = snk_make_space(db, SNK_SPACE_JOINT);
    g1_params p;
    memset(&p, 0, sizeof(p));
    const std_params stp{db->std_mean, db->std_sd, db->wt, db->uv_special, db->uv_scale, db->std_f32};
    p.g = greedy_src{nullptr, 0, 0, d_targets, d_unnorm, stp, db->Dt, db->m, d_Jc_full ? d_Jc_full : db->Jc_raw, db->wj, db->Dj, db->Djq,
                     db->prev_row_off, db->prev_col, db->cur_row_off, db->cur_col, nullptr, nullptr, d_paths, d_step_dist};
    p.meta = meta;
    p.rs = make_rr(db, sp);
    p.S16 = db->S16; p.G16 = db->G16; p.ldS = db->ldS16; p.ldG = db->ldG16;
    p.qmap = snk_tc_qmap(db, SNK_SPACE_JOINT);
    p.ld16 = snk_tc_query_ld(db, SNK_SPACE_JOINT);
    SNK_CHECK(p.ld16 == db->ldS16 + db->m * db->ldG16 && db->ldS16 == LDS_H && db->ldG16 == LDG_H, "internal: query operand layout changed");
    p.Np = db->Np;
    const int grid = db->sm_count;
    p.ngroups = snk_cdiv(db->Np, 16);
    p.groups_per_cta = snk_cdiv(p.ngroups, grid);
    // exchange area: lists [2][grid][LC] | taus [2][grid] | arrive counter
    const size_t lists_bytes = (size_t)2 * grid * G1_LC * 8, taus_bytes = snk_round_up((size_t)4 * grid * 4, 16);   // taus + durs
    const bool timing = getenv("SNK_G1_TIMING") != nullptr;
    SNK_TRY(snk_buf_reserve(&db->ws_g1, lists_bytes + taus_bytes + 16 + (timing ? 4096 * 128 + (size_t)256 * grid * 8 : 0)));
    p.lists = (unsigned long long *)db->ws_g1.p;
    p.taus = (float *)((char *)db->ws_g1.p + lists_bytes);
    p.durs = (unsigned *)(p.taus + 2 * grid);
    p.bar = (unsigned *)((char *)db->ws_g1.p + lists_bytes + taus_bytes);
    SNK_CUDA(cudaMemsetAsync(p.bar, 0, 4, st));
    p.dbg_times = timing ? (unsigned long long *)((char *)db->ws_g1.p + lists_bytes + taus_bytes + 16) : nullptr;
    p.dbg_cta = timing ? p.dbg_times + 4096 * 16 : nullptr;
    p.dberr = db->err_j16; p.maxn = db->maxn_j16;
    p.eps_rel = snk_tc_eps_rel(db, SNK_SPACE_JOINT);
    p.debug_fail_mod = db->debug_fail_mod;
    p.flags = d_flags; p.count = d_count;
    p.dbg_keys = d_keys;
    p.id_offset = id_offset;
    p.balance = getenv("SNK_G1_NO_BALANCE") ? 0 : 1;
    if (d_Jc_full && snk_comm_nranks(db) > 1) {   // database-sharded: the exchange regions of the communicator; this launch owns the next nsteps epochs
        SNK_TRY(snk_comm_p2p_claim(db, (int)meta.nsteps, &p.peers, &p.rank, &p.R, &p.epoch0, &p.xflags_bytes, &p.xslot_bytes));
        SNK_CHECK(p.peers && p.R >= 2 && p.R <= 32, "internal: single-utterance sharded search needs the peer-memory exchange");
    }
    g1_fn fn = g1_pick(db->m);
    SNK_CHECK(fn, "internal: no single-utterance kernel for multiepoch %d", db->m);
    const size_t smem = g1_smem_bytes(p.rs, grid, p.ld16);
    SNK_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&p};
    // cooperative launch: all CTAs are resident, the grid barrier cannot deadlock
    cudaLaunchAttribute attrs[1];
    memset(attrs, 0, sizeof(attrs));
    attrs[0].id = cudaLaunchAttributeCooperative;
    attrs[0].val.cooperative = 1;
    const int nattr = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(G1_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attrs;
    cfg.numAttrs = nattr;
    SNK_CUDA(cudaLaunchKernelExC(&cfg, (const void *)fn, args));
    db->counters[0] += meta.nsteps;
    db->counters[2] += 1;
    return 0;
}

// Test instrumentation (include/snk_b200.h): the keys the single-utterance kernel computes for the first step of an
// utterance whose first window is `targets` [multiepoch, Dt] (weighted float64, host), the fp32 ||x~||^2 it adds to them
// and the slack its certificate allows.
extern "C" int snk_debug_greedy_one_keys(snk_db *db, const double *targets, int64_t start_state, float *keys, float *qnorm,
                                         float *eps_rel, float *maxnorm) {
    SNK_CHECK(db && targets && keys && db->weights_set, "NULL argument / weights not set");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    SNK_CHECK(snk_greedy_one_supported(db), "the single-utterance kernel does not support this voice / engine setting");
    SNK_CHECK(start_state < db->Np, "start_state out of range");
    cudaStream_t st = db->stream;
    const size_t tbytes = (size_t)db->m * db->Dt * 8;
    SNK_TRY(snk_buf_reserve(&db->ws_h0, tbytes));
    SNK_TRY(snk_buf_reserve(&db->ws_h1, 64));
    SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)(db->Np + 1) * 4));
    SNK_CUDA(cudaMemcpyAsync(db->ws_h0.p, targets, tbytes, cudaMemcpyHostToDevice, st));
    int64_t *path = (int64_t *)db->ws_h1.p;
    int *flags = (int *)(path + 2);
    SNK_CUDA(cudaMemsetAsync(flags, 0, 8, st));
    const greedy_meta meta{0, 0, 1, start_state};
    SNK_TRY(snk_greedy_one_launch(db, &meta, (const double *)db->ws_h0.p, nullptr, path, nullptr, flags, flags + 1,
                                  (float *)db->ws_dist.p, st, nullptr, 0));
    SNK_CUDA(cudaMemcpyAsync(keys, db->ws_dist.p, (size_t)db->Np * 4, cudaMemcpyDeviceToHost, st));
    if (qnorm) SNK_CUDA(cudaMemcpyAsync(qnorm, (float *)db->ws_dist.p + db->Np, 4, cudaMemcpyDeviceToHost, st));
    if (maxnorm) SNK_CUDA(cudaMemcpyAsync(maxnorm, db->maxn_j16, 4, cudaMemcpyDeviceToHost, st));
    SNK_CUDA(cudaStreamSynchronize(st));
    if (eps_rel) *eps_rel = snk_tc_eps_rel(db, SNK_SPACE_JOINT);
    return 0;
}

// Diagnostic (SNK_G1_TIMING=1): CTA 0's timestamps of the last single-utterance launch, [steps][16] nanoseconds (slots 0-5 used):
// own scan finished, all warps' scans finished, list published, barrier passed, nearest row known, next query ready.
extern "C" int snk_debug_greedy_one_cta_times(snk_db *db, unsigned long long *out, int steps) {
    SNK_CHECK(db && out && steps >= 1 && steps <= 256, "bad argument");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    const int grid = db->sm_count;
    const size_t off = (size_t)2 * grid * G1_LC * 8 + snk_round_up((size_t)4 * grid * 4, 16) + 16 + (size_t)4096 * 128;
    SNK_CHECK(db->ws_g1.p && db->ws_g1.cap >= off + (size_t)256 * grid * 8, "no timed launch yet (SNK_G1_TIMING=1)");
    SNK_CUDA(cudaDeviceSynchronize());
    SNK_CUDA(cudaMemcpy(out, (char *)db->ws_g1.p + off, (size_t)steps * grid * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int snk_debug_greedy_one_times(snk_db *db, unsigned long long *out, int steps) {
    SNK_CHECK(db && out && steps >= 1 && steps <= 4096, "bad argument");
    SNK_LOCK(db);
    SNK_CUDA(cudaSetDevice(db->device));
    const int grid = db->sm_count;
    const size_t off = (size_t)2 * grid * G1_LC * 8 + snk_round_up((size_t)4 * grid * 4, 16) + 16;
    SNK_CHECK(db->ws_g1.p && db->ws_g1.cap >= off + (size_t)steps * 128, "no timed launch yet (SNK_G1_TIMING=1)");
    SNK_CUDA(cudaDeviceSynchronize());
    SNK_CUDA(cudaMemcpy(out, (char *)db->ws_g1.p + off, (size_t)steps * 128, cudaMemcpyDeviceToHost));
    return 0;
}
