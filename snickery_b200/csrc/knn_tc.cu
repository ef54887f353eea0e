// Tensor-core shortlist kernel: brute-force target/joint cost as a dense contraction on tcgen05.
//
// Replaces the KD-tree lookup of the reference (cKDTree.query at script/synth_simple.py:490 and
// script/synth_halfphone.py:1364) with an exact scan: for a tile of 128 queries x 128 database
// rows the kernel accumulates x~ . y~ in TMEM (fp16 operands, fp32 accumulate), turns it into the
// key ||y~||^2 - 2 x~.y~ in the epilogue and keeps, per query, the smallest keys it has seen.
//
//  * operands are fp16 copies of the weighted rows (weights.cu); their squared norms are taken
//    from the ROUNDED values, so key + ||x~||^2 is the squared distance between the rounded
//    vectors and |sqrt(key + ||x~||^2) - true distance| <= ||x - x~|| + ||y - y~|| (triangle
//    inequality).  rerank.cu uses that bound to certify the float64 re-ranked answer exact.
//  * the multiepoch joint row [start_join(u) || F(u) .. F(u+m-1)] is never materialised: one TMA
//    box of 136 rows of the frame matrix G16 (a "slab") serves all m window offsets -- K-block j
//    reads the slab from row j on (descriptor start address + j * 128 B).
//  * the squared norms ride in three spare columns of every operand row (hi/mid/lo fp16 pieces,
//    the query holds -0.5 there), so the accumulator already is x.y - ||y||^2 / 2.
//  * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM allocator,
//    warps 2.. = epilogue (one thread per query / TMEM lane, two or four warps per lane quarter).
//    The 128-query operand tile stays resident in shared memory; database tiles stream through an
//    mbarrier ring; two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
//  * for the shapes the reference ships the ring and the K-block sequence are compile-time
//    schedules (sched_traits); other shapes run a table-driven variant of the same kernel.
//  * multiepoch 6 keeps the query operand in TENSOR memory (written once per CTA with tcgen05.st): the UTCHMMAs then
//    read only the database operand from shared memory, whose bandwidth is what two smem operands saturate.
//  * epilogues: MODE_LIST (register lists of the 4 / 8 best per query, k <= 4), MODE_STORE (keys to HBM, small
//    databases), MODE_MINS (the smallest key of every 32-row group: sampling pass of larger k), MODE_EMIT (append every
//    row at or below a per-query bound; larger k).
//  * each CTA scans one (query tile, database chunk) pair; per-chunk results are merged by
//    rerank.cu (in-block) or snk_topk_scan.  Optionally two CTAs form a cluster and share every
//    database tile by TMA multicast (SNK_TC_CLUSTER=1).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#pragma nv_diag_suppress 177   // traits members a given schedule does not use
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int BM = 128;        // queries per tile (UMMA M)
constexpr int BN = 128;        // database rows per tile (UMMA N)
constexpr int BK = 64;         // fp16 elements per K-block = one 128-byte swizzle row
constexpr int MAX_STAGES = 12;
constexpr int MAXKB = 10;      // resident query K-blocks
constexpr int MAXLOAD = 10;    // TMA loads per database tile
constexpr int MAXSUB = 16;     // MMA K-blocks per database tile
constexpr int SLAB_EXTRA = 8;  // extra rows of a frame slab: serves window offsets 0..8
// epilogue warps per TMEM lane quarter (each takes BN / split columns of a tile): 2 by default, 4 for the
// half-phone target shape (K = 192: a tile is only 768 tensor-pipe cycles, the epilogue is the longer stage and
// more warps hide its latency; measured 22.7 -> 20.1 ms on the k = 50 pipeline).  The one-frame joint shape
// (multiepoch 1) does not gain: it is bound by the 64 KB per tile it pulls from L2.
__host__ __device__ constexpr int epi_split_of(int sched) { return (sched == 2 || sched == 12) ? 4 : 2; }
constexpr int TILE_BYTES = BM * BK * 2;                    // 16 KiB query K-block
constexpr int SLOT_BYTES = (BN + SLAB_EXTRA) * BK * 2;     // 17 KiB ring slot (plain tile or frame slab)
__host__ __device__ constexpr int num_threads_of(int sched) { return 64 + 128 * epi_split_of(sched); }

// One TMA load per ring slot; it feeds nsub MMA K-blocks.  A frame slab (BN + 8 rows of G16) feeds the
// m window offsets of the multiepoch row: K-block j reads the slab from row j on, so the sliding
// window is never materialised and each frame row is fetched once per tile instead of m times.
struct tc_load { int map, rowoff, col, bytes, sub0, nsub; };
struct tc_sub { int a_blk, b_off, base_off, ksteps; };
struct tc_params {
    int nkb;                // resident query K-blocks
    int nload, nsub, stages;
    int embed;              // squared norms ride in the operands: key = -2 * accumulator
    int cluster;            // CTAs per cluster sharing every database tile by TMA multicast (1 or 2)
    int b_bytes;            // bytes of the database-tile ring
    tc_load load[MAXLOAD];  // map: 0 join-context tile (S16), 1 frame tile (G16), 2 frame slab (G16, BN+8 rows)
    tc_sub sub[MAXSUB];
    int64_t row_lo, row_hi; // rows scanned by this launch
    int64_t chunk_rows;     // multiple of BN
    int nchunks;
    int64_t nq;
    const __half *q16;      // the query operand rows (SCHED 26 reads the frame part straight into tensor memory)
    int ldq16;
    const float *nrm;       // [rows] squared norms of the fp16 rows
    float *oval;            // fused : [nq_pad, nchunks * split, LSZ] keys (ascending)
    int *oid;               //         row ids
    float *odist;           // store : [nq, ldo] keys of the scanned tiles, compacted (tile_stride > 1 = sample)
    int64_t ldo;
    int tile_stride;        // scan every tile_stride-th tile of a chunk (1 = all)
    int mins_group, mins_per_chunk;   // MINS: tiles per minimum group, groups per chunk; odist [nq, nchunks * mins_per_chunk * split * 32]
    // emit mode: every row whose key is <= thr[q] is appended to the (query, chunk, half) buffer
    const float *thr;       // [nq]
    float *bufv;            // [nq_pad, nchunks * split, cap]
    int *bufi;
    int *bufn;              // [nq_pad, nchunks * split] filled slots of every buffer
    int cap;
    float *tau;             // [nq] preset to thr; a buffer overflow writes -inf (certificate must fail)
};
constexpr int MODE_LIST = 0, MODE_STORE = 1, MODE_EMIT = 2, MODE_MINS = 3;   // MINS: the smallest key of every 32-column group
// half-box tensor maps of the multicast path: each CTA of a pair fetches half of every database tile
struct tc_maps_mc { CUtensorMap S_h, G_h, Gslab72; };
constexpr int MC_ROWS0 = 72;   // rows of a frame slab fetched by rank 0 (9 swizzle atoms); rank 1 takes the other 64

using namespace snk_ptx;

// Shared-memory operand descriptors (K-major, 128-byte swizzle: rows of 64 fp16 = 128 B, 8-row groups
// 1024 B apart) are assembled in the MMA warp: start address >> 4 in bits 0-13, stride byte offset
// 1024 >> 4 in bits 32-45, version 1 in bits 46-47, SWIZZLE_128B (2) in bits 61-63.  The hardware
// applies the swizzle to absolute address bits, so an operand may start on any row of a slab
// (base offset field stays 0) and advance along K by +32 B per UMMA_K step.
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=128
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------- static schedules
// For the shapes the reference ships (config/*.cfg: 151-dim join + 61-dim target frames with
// multiepoch 6, 4, 3 or 1; 184-dim half-phone targets) the per-tile load / K-block sequence is a compile-time
// constant, so the issue thread's descriptors are "uniform base + immediate" and it sustains one
// UTCHMMA per tensor-pipe slot.  Every other shape runs the table-driven path (SCHED 0).
//   NS   : K-blocks of the join part (one plain tile load each), KSL: UMMA_K steps of its last block
//   TB   : 64-column blocks per target frame (one slab load each), KTL: UMMA_K steps of the last one
//   M    : frames per row (window offsets served by one slab)
//   GR   : slots per target-block load (2 = the next tile's frames land while this tile's are multiplied);
//          join-part loads keep one slot each: they are short, so their slot is free long before reuse
// Static schedules rely on the squared norms embedded in the operands (weights.cu), so their epilogue
// needs no norm staging.
//   SR   : slots per join-part load.  One is enough when a tile keeps the tensor pipe busy for longer than a load
//          takes to arrive (multiepoch 4, 6); shorter tiles (multiepoch 1, 3) need the second slot, and have the
//          shared memory for it because their query tile is smaller
//   ATM  : the query operand lives in TENSOR MEMORY instead of shared memory (SCHED 20 + m, SCHED 12): all frame
//          blocks plus the first AS join-part blocks, next to the two accumulators (512 columns in all).  A 128x128x16
//          UTCHMMA with both operands in shared memory reads 8 KB per 64 cycles -- all of an SM's shared-memory
//          bandwidth, with TMA writing into the same banks -- and ncu showed the tensor pipe waiting on operand fetch
//          ~20 % of the time.  With A in tensor memory the UTCHMMA reads only B from shared memory.  (Spending the freed
//          shared memory on a deeper database ring was measured slower: 10.1 vs 10.35 M frames/s.)
template <int SCHED> struct sched_traits {
    static constexpr int NS = 0, KSL = 0, TB = 0, KTL = 0, M = 1, GR = 1, SR = 1, AS = 0;
    static constexpr bool ATM = false;
};
template <int M_> struct joint_traits {   // joint 151 | M x 61
    static constexpr int NS = 3, KSL = 2, TB = 1, KTL = 4, M = M_, GR = 2, SR = M_ <= 3 ? 2 : 1, AS = 0;
    static constexpr bool ATM = false;
};
template <int M_> struct joint_traits_tm {   // the same with the query operand in tensor memory
    static constexpr int NS = 3, KSL = 2, TB = 1, KTL = 4, M = M_, GR = 2, SR = M_ <= 3 ? 2 : 1;   // same ring as the smem variants
    static constexpr int AS = M_ * 32 + 3 * 32 <= 256 ? 3 : 2;     // multiepoch 6: 192 + 64 columns, the last join block stays in smem
    static constexpr bool ATM = true;
};
template <> struct sched_traits<2> {   // target 184
    static constexpr int NS = 0, KSL = 0, TB = 3, KTL = 4, M = 1, GR = 2, SR = 1, AS = 0;
    static constexpr bool ATM = false;
};
template <> struct sched_traits<12> {   // target 184, query operand in tensor memory
    static constexpr int NS = 0, KSL = 0, TB = 3, KTL = 4, M = 1, GR = 2, SR = 1, AS = 0;
    static constexpr bool ATM = true;
};
// SCHED 10 + m: the joint space at the multiepoch values the reference's configs use (config/*.cfg: 6, 1, 4, 3)
template <> struct sched_traits<11> : joint_traits<1> {};
template <> struct sched_traits<13> : joint_traits<3> {};
template <> struct sched_traits<14> : joint_traits<4> {};
template <> struct sched_traits<16> : joint_traits<6> {};
template <> struct sched_traits<21> : joint_traits_tm<1> {};
template <> struct sched_traits<23> : joint_traits_tm<3> {};
template <> struct sched_traits<24> : joint_traits_tm<4> {};
template <> struct sched_traits<26> : joint_traits_tm<6> {};
template <int SCHED> struct sched_layout {
    using S = sched_traits<SCHED>;
    static constexpr int GBYTES = S::M > 1 ? SLOT_BYTES : TILE_BYTES;
    static constexpr int NSLOT = S::NS * S::SR + S::TB * S::GR;
    static constexpr int B_BYTES = S::NS * S::SR * TILE_BYTES + S::TB * S::GR * GBYTES;
    __host__ __device__ static constexpr int s_slot(int l, int r) { return l * S::SR + r; }
    __host__ __device__ static constexpr uint32_t s_off(int l, int r) { return (uint32_t)(l * S::SR + r) * TILE_BYTES; }
    __host__ __device__ static constexpr int g_slot(int b, int r) { return S::NS * S::SR + b * S::GR + r; }
    __host__ __device__ static constexpr uint32_t g_off(int b, int r) {
        return (uint32_t)(S::NS * S::SR) * TILE_BYTES + (uint32_t)(b * S::GR + r) * GBYTES;
    }
};

// ---------------------------------------------------------------- kernel
template <int MODE, int LSZ, int SCHED>
__global__ void __launch_bounds__(num_threads_of(SCHED), 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapS,
              const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapGslab,
              const __grid_constant__ tc_maps_mc mc, const tc_params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);                        // SWIZZLE_128B tiles need 1024 B alignment
    if (sbase & 1023u) __trap();
    uint8_t *gbase = smem_raw;
    constexpr bool ATM = sched_traits<SCHED>::ATM;
    constexpr int AS = sched_traits<SCHED>::AS;                       // join-part blocks of the query held in tensor memory
    const int NKB = ATM ? sched_traits<SCHED>::NS - AS : p.nkb;       // query K-blocks resident in shared memory
    // accumulators in flight: four for the static schedules that keep the query in shared memory (all 512 tensor-memory
    // columns; the issuing thread runs up to three tiles ahead of the epilogue, which otherwise waits a barrier and a
    // tcgen05.ld round trip per tile: the K = 192 tile is only 768 tensor-pipe cycles), two when the query operand takes
    // 256 columns and for the table-driven path (its norm staging is double-buffered)
    constexpr int NACC = (SCHED != 0 && !ATM) ? 4 : 2;
    constexpr uint32_t TCOLS = ATM ? 512u : (uint32_t)(NACC * BN);
    constexpr uint32_t ATM_COL = 2 * BN;                              // tensor-memory columns of A: [AS join blocks][frame blocks], 32 each
    const uint32_t sA = sbase;
    const uint32_t sB = sA + (uint32_t)NKB * TILE_BYTES;
    const int STAGES = p.stages;
    const uint32_t sBar = sB + (uint32_t)p.b_bytes;
    // barriers: full[MAX_STAGES] empty[MAX_STAGES] a_full tmem_full[2] tmem_empty[2]
    const uint32_t bar_full = sBar, bar_empty = sBar + 8 * MAX_STAGES, bar_a = sBar + 16 * MAX_STAGES;
    const uint32_t bar_tfull = bar_a + 8, bar_tempty = bar_tfull + 32;      // room for four accumulators each
    const uint32_t s_tmem_ptr = bar_tempty + 32;
    const uint32_t bar_atm = s_tmem_ptr + 8;                          // A frames have been written to tensor memory
    uint8_t *g_after = gbase + (size_t)NKB * TILE_BYTES + (size_t)p.b_bytes + 16 * MAX_STAGES + 8 + 64;
    volatile uint32_t *tmem_ptr_g = reinterpret_cast<volatile uint32_t *>(g_after);
    float *nrm_s = reinterpret_cast<float *>(g_after + 24);           // [2][BN], 16-byte aligned (table-driven path only)
    uint4 *sub_s = reinterpret_cast<uint4 *>(g_after + 24 + 2 * BN * 4);   // [MAXSUB] {a start addr >> 4, b byte offset, ksteps, -}
    (void)STAGES;

    constexpr int EPI_SPLIT = epi_split_of(SCHED);
    constexpr int NUM_EPI_THREADS = 128 * EPI_SPLIT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a cluster (pair) scans one chunk with two query tiles: every database tile is read from L2 once for both
    const int crank = p.cluster > 1 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / p.cluster;
    const int qt = (cid / p.nchunks) * p.cluster + crank, chunk = cid % p.nchunks;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    const int64_t row_beg = p.row_lo + (int64_t)chunk * p.chunk_rows;
    const int64_t row_end = min(p.row_hi, row_beg + p.chunk_rows);
    const int64_t tstep = (int64_t)BN * p.tile_stride;
    const int ntiles = row_end > row_beg ? (int)((row_end - row_beg + tstep - 1) / tstep) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, (uint32_t)p.cluster);   // one commit per CTA sharing the slot
        }
        mbar_init(bar_a, 1);
        if (ATM) mbar_init(bar_atm, 128 * epi_split_of(SCHED));
        for (int a = 0; a < NACC; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, NUM_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(s_tmem_ptr, TCOLS);
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();     // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_g;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            tma_prefetch_desc(&mapQ);
            tma_prefetch_desc(&mapS);
            tma_prefetch_desc(&mapG);
            tma_prefetch_desc(&mapGslab);
            mbar_expect_tx(bar_a, (uint32_t)NKB * TILE_BYTES);
            for (int kb = 0; kb < NKB; ++kb)      // ATM: only the join blocks that did not fit in tensor memory
                tma_load_2d(sA + kb * TILE_BYTES, &mapQ, (kb + (ATM ? AS : 0)) * BK, qt * BM, bar_a);
        }
        __syncwarp();
        if constexpr (SCHED != 0) {
            using S = sched_traits<SCHED>;
            using L = sched_layout<SCHED>;
            for (int t = 0; t < ntiles; ++t) {
                const int r0 = (int)(row_beg + (int64_t)t * tstep);
                const int sr = t % S::SR;
                const uint32_t ph_s = (uint32_t)(t / S::SR) & 1u;
                const int gr = t % S::GR;
                const uint32_t ph_g = (uint32_t)(t / S::GR) & 1u;
#pragma unroll
                for (int l = 0; l < S::NS; ++l) {
                    const uint32_t bar = 8 * (uint32_t)L::s_slot(l, sr);
                    mbar_wait(bar_empty + bar, ph_s ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + bar, TILE_BYTES);
                        if (p.cluster > 1)   // this CTA's half of the rows, delivered to both CTAs
                            tma_load_2d_mc(sB + L::s_off(l, sr) + crank * (TILE_BYTES / 2), &mc.S_h, l * BK, r0 + crank * (BN / 2),
                                           bar_full + bar, cmask);
                        else
                            tma_load_2d(sB + L::s_off(l, sr), &mapS, l * BK, r0, bar_full + bar);
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int b = 0; b < S::TB; ++b) {
                    const uint32_t bar = 8 * (uint32_t)L::g_slot(b, gr);
                    mbar_wait(bar_empty + bar, ph_g ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + bar, L::GBYTES);
                        if (p.cluster > 1) {
                            if (S::M > 1) {   // slab of BN + 8 rows: 72 rows (9 swizzle atoms) from rank 0, 64 from rank 1
                                if (crank == 0)
                                    tma_load_2d_mc(sB + L::g_off(b, gr), &mc.Gslab72, b * BK, r0, bar_full + bar, cmask);
                                else
                                    tma_load_2d_mc(sB + L::g_off(b, gr) + MC_ROWS0 * BK * 2, &mc.G_h, b * BK, r0 + MC_ROWS0,
                                                   bar_full + bar, cmask);
                            } else {
                                tma_load_2d_mc(sB + L::g_off(b, gr) + crank * (TILE_BYTES / 2), &mc.G_h, b * BK,
                                               r0 + crank * (BN / 2), bar_full + bar, cmask);
                            }
                        } else {
                            tma_load_2d(sB + L::g_off(b, gr), S::M > 1 ? &mapGslab : &mapG, b * BK, r0, bar_full + bar);
                        }
                    }
                    __syncwarp();
                }
            }
        } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int64_t r0 = row_beg + (int64_t)t * tstep;
            for (int l = 0; l < p.nload; ++l) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    const tc_load ld = p.load[l];
                    mbar_expect_tx(bar_full + 8 * stage, (uint32_t)ld.bytes);
                    const CUtensorMap *map = ld.map == 0 ? &mapS : (ld.map == 1 ? &mapG : &mapGslab);
                    tma_load_2d(sB + stage * SLOT_BYTES, map, ld.col, (int)(r0 + ld.rowoff), bar_full + 8 * stage);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the pipeline, one elected lane issues) ==========
        if constexpr (SCHED == 0) {   // descriptor table of the table-driven path (not allocated otherwise)
            if (lane < p.nsub) {
                const tc_sub su = p.sub[lane];
                sub_s[lane] = make_uint4((sA + (uint32_t)su.a_blk * TILE_BYTES) >> 4, (uint32_t)su.b_off, (uint32_t)su.ksteps, 0u);
            }
            __syncwarp();
        }
        mbar_wait(bar_a, 0);
        if (ATM) mbar_wait(bar_atm, 0);
        tc_fence_after();
        constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        if constexpr (SCHED != 0) {
            using S = sched_traits<SCHED>;
            using L = sched_layout<SCHED>;
            if (elect_one()) {
                const uint32_t a0 = sA >> 4, b0 = sB >> 4;       // start-address fields; every offset below is an immediate
                for (int t = 0; t < ntiles; ++t) {
                    const int acc = t % NACC;
                    const uint32_t acc_phase = (uint32_t)(t / NACC) & 1;
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);   // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                    const int sr = t % S::SR;
                    const uint32_t ph_s = (uint32_t)(t / S::SR) & 1u;
                    const int gr = t % S::GR;
                    const uint32_t ph_g = (uint32_t)(t / S::GR) & 1u;
#pragma unroll
                    for (int l = 0; l < S::NS; ++l) {
                        mbar_wait(bar_full + 8 * (uint32_t)L::s_slot(l, sr), ph_s);
                        tc_fence_after();
                        const uint32_t al = a0 + (l - (ATM ? AS : 0)) * (TILE_BYTES >> 4), bl = b0 + (L::s_off(l, sr) >> 4);
#pragma unroll
                        for (int ks = 0; ks < (l == S::NS - 1 ? S::KSL : 4); ++ks) {
                            if (ATM && l < AS)
                                umma_f16_ts(d_tmem, tmem_base + ATM_COL + (uint32_t)(l * (BK / 2) + ks * 8),
                                            DESC_HI | (uint64_t)(bl + 2 * ks), IDESC, (l | ks) != 0 ? 1u : 0u);
                            else
                                umma_f16(d_tmem, DESC_HI | (uint64_t)(al + 2 * ks), DESC_HI | (uint64_t)(bl + 2 * ks), IDESC,
                                         (l | ks) != 0 ? 1u : 0u);
                        }
                        if (p.cluster > 1) umma_commit_mc(bar_empty + 8 * (uint32_t)L::s_slot(l, sr), cmask);
                        else umma_commit(bar_empty + 8 * (uint32_t)L::s_slot(l, sr));
                    }
#pragma unroll
                    for (int b = 0; b < S::TB; ++b) {
                        const uint32_t bar = 8 * (uint32_t)L::g_slot(b, gr);
                        mbar_wait(bar_full + bar, ph_g);
                        tc_fence_after();
                        const uint32_t bl = b0 + (L::g_off(b, gr) >> 4);
#pragma unroll
                        for (int j = 0; j < S::M; ++j) {
                            const uint32_t al = a0 + (S::NS + j * S::TB + b) * (TILE_BYTES >> 4);
#pragma unroll
                            for (int ks = 0; ks < (b == S::TB - 1 ? S::KTL : 4); ++ks) {   // window offset j = +j rows = +8 in the field
                                if constexpr (ATM)   // A block (j, b): 32 columns of tensor memory, 8 per UMMA_K step
                                    umma_f16_ts(d_tmem, tmem_base + ATM_COL + (uint32_t)((AS + j * S::TB + b) * (BK / 2) + ks * 8),
                                                DESC_HI | (uint64_t)(bl + 8 * j + 2 * ks), IDESC, (S::NS + b + j + ks) != 0 ? 1u : 0u);
                                else
                                    umma_f16(d_tmem, DESC_HI | (uint64_t)(al + 2 * ks), DESC_HI | (uint64_t)(bl + 8 * j + 2 * ks),
                                             IDESC, (S::NS + b + j + ks) != 0 ? 1u : 0u);
                            }
                        }
                        if (p.cluster > 1) umma_commit_mc(bar_empty + bar, cmask);
                        else umma_commit(bar_empty + bar);
                    }
                    umma_commit(bar_tfull + 8 * acc);
                }
            }
            __syncwarp();
        } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int acc = t % NACC;
            const uint32_t acc_phase = (uint32_t)(t / NACC) & 1;
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);       // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
            int sb = 0;
            for (int l = 0; l < p.nload; ++l) {
                const int sb_end = sb + p.load[l].nsub;
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t bbase = sB + (uint32_t)stage * SLOT_BYTES;
                    for (int i = sb; i < sb_end; ++i) {
                        const uint4 su = sub_s[i];
                        const uint64_t a_desc = DESC_HI | (uint64_t)(su.x & 0x3FFFu);
                        const uint64_t b_desc = DESC_HI | (uint64_t)(((bbase + su.y) >> 4) & 0x3FFFu);
                        // +32 B along K inside the 128 B swizzle row = +2 in the start-address field
                        umma_f16(d_tmem, a_desc, b_desc, IDESC, (uint32_t)(l | i) != 0u);
                        if (su.z > 1) umma_f16(d_tmem, a_desc + 2, b_desc + 2, IDESC, 1u);
                        if (su.z > 2) umma_f16(d_tmem, a_desc + 4, b_desc + 4, IDESC, 1u);
                        if (su.z > 3) umma_f16(d_tmem, a_desc + 6, b_desc + 6, IDESC, 1u);
                    }
                    umma_commit(bar_empty + 8 * stage);           // smem slot reusable once these MMAs retire
                    if (l == p.nload - 1) umma_commit(bar_tfull + 8 * acc);   // accumulator complete
                }
                __syncwarp();
                sb = sb_end;
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        }
    } else {
        // ===================== epilogue: one thread per query =====================
        const int quarter = warp & 3;                 // TMEM lanes this warp may read
        const int half = (warp - 2) >> 2;             // which BN / EPI_SPLIT column range of every tile
        const int ql = quarter * 32 + lane;           // query row inside the tile == TMEM lane
        const int64_t q = (int64_t)qt * BM + ql;
        const int et = (warp - 2) * 32 + lane;        // 0 .. NUM_EPI_THREADS-1 among the epilogue threads
        constexpr int COLS = BN / EPI_SPLIT;
        float lv[LSZ];
        int li[LSZ];
#pragma unroll
        for (int i = 0; i < LSZ; ++i) { lv[i] = INFINITY; li[i] = -1; }
        if constexpr (ATM) {
            // This thread's query row goes to its tensor-memory lane: the first AS join blocks and every frame block of
            // the operand row (K-block order), 8 columns (16 fp16) per store, the chunks dealt out over the EPI_SPLIT
            // threads that share the lane.
            using S = sched_traits<SCHED>;
            constexpr int CH_S = AS * (BK / 16), CH = CH_S + S::M * S::TB * (BK / 16);     // 8-column chunks
            const uint4 *row = reinterpret_cast<const uint4 *>(p.q16 + q * (int64_t)p.ldq16);   // 8 fp16 per uint4
            const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + ATM_COL;
            for (int c = half; c < CH; c += EPI_SPLIT) {
                const int src = c < CH_S ? c : c - CH_S + S::NS * (BK / 16);     // chunk index in the operand row
                const uint4 u0 = __ldg(row + 2 * src), u1 = __ldg(row + 2 * src + 1);
                tmem_st8(lane_base + (uint32_t)c * 8, u0, u1);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_atm);
        }
        const bool embed = SCHED != 0 || p.embed != 0;
        // Static schedules always embed the norms, so key = -2 * accumulator.  Lists, votes and thresholds then work on
        // HALF keys (-accumulator, a sign flip the compiler folds into the comparisons; the group minimum is a max3 tree
        // over the raw accumulators) and the exact factor 2 is applied only to the few values that leave the kernel:
        // 32 multiplies per 32-column chunk and thread less, same bits out.
        constexpr bool HALFKEY = SCHED != 0 && MODE != MODE_STORE;
        constexpr float OSCALE = HALFKEY ? 2.f : 1.f;
        float nrm_next = (!embed && et < BN && ntiles > 0 && row_beg + et < row_end) ? __ldg(p.nrm + row_beg + et) : INFINITY;
        // emit mode state
        const float thr = (MODE == MODE_EMIT && q < p.nq) ? p.thr[q] * (HALFKEY ? 0.5f : 1.f) : -INFINITY;
        int ecnt = 0;
        float rmin[MODE == MODE_MINS ? 32 : 1];          // MINS: running per-column extremum of the current tile group
#pragma unroll
        for (int j = 0; j < (MODE == MODE_MINS ? 32 : 1); ++j) rmin[j] = HALFKEY ? -INFINITY : INFINITY;
        float *ebv = MODE == MODE_EMIT ? p.bufv + (((size_t)q * p.nchunks + chunk) * EPI_SPLIT + half) * p.cap : nullptr;
        int *ebi = MODE == MODE_EMIT ? p.bufi + (((size_t)q * p.nchunks + chunk) * EPI_SPLIT + half) * p.cap : nullptr;
        auto flush_lists = [&]() {
            const size_t slot = ((size_t)q * p.nchunks + chunk) * EPI_SPLIT + half;
            float *ov = p.oval + slot * LSZ;
            int *oi = p.oid + slot * LSZ;
#pragma unroll
            for (int i = 0; i < LSZ; ++i) { ov[i] = OSCALE * lv[i]; oi[i] = li[i]; lv[i] = INFINITY; li[i] = -1; }
        };
        for (int t = 0; t < ntiles; ++t) {
            const int acc = t % NACC;
            const uint32_t acc_phase = (uint32_t)(t / NACC) & 1;
            const int64_t r0 = row_beg + (int64_t)t * tstep;
            if (!embed) {
                if (et < BN) {
                    nrm_s[acc * BN + et] = nrm_next;
                    nrm_next = (t + 1 < ntiles && r0 + tstep + et < row_end) ? __ldg(p.nrm + r0 + tstep + et) : INFINITY;   // next tile's norm, a tile early
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
            }
            const bool partial = r0 + BN > row_end;   // rows past the chunk end (zero-filled by TMA) must not compete
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * BN;
#pragma unroll 1
            for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += 32) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c0, v);
                if (c0 + 32 == (half + 1) * COLS) {   // this thread's part of the accumulator is read: hand it back
                    tc_fence_before();
                    mbar_arrive(bar_tempty + 8 * acc);
                }
                // key = ||y~||^2 - 2 x~.y~ : with embedded norms the accumulator already holds x.y - ||y||^2 / 2
                float key[32];
                if constexpr (HALFKEY) {
                    if (partial) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (r0 + c0 + j >= row_end) v[j] = -INFINITY;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) key[j] = -v[j];     // folded into the consumers as a source modifier
                } else {
                    if (embed) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) key[j] = -2.f * v[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) key[j] = fmaf(-2.f, v[j], nrm_s[acc * BN + c0 + j]);
                    }
                    if (partial) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (r0 + c0 + j >= row_end) key[j] = INFINITY;
                    }
                }
                if (MODE == MODE_EMIT) {
                    // a few rows per thousand pass the bound: maxima of 8-column groups first (three-input max tree), so
                    // that the per-column test and its predicated stores run for the rare group that holds a hit
                    float g[4];
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        if constexpr (HALFKEY) {
                            const float *a8 = v + 8 * gi;
                            g[gi] = -max3(max3(a8[0], a8[1], a8[2]), max3(a8[3], a8[4], a8[5]), fmaxf(a8[6], a8[7]));
                        } else {
                            const float *k8 = key + 8 * gi;
                            g[gi] = min3(min3(k8[0], k8[1], k8[2]), min3(k8[3], k8[4], k8[5]), fminf(k8[6], k8[7]));
                        }
                    }
                    if (fminf(fminf(g[0], g[1]), fminf(g[2], g[3])) <= thr) {
#pragma unroll
                        for (int gi = 0; gi < 4; ++gi) {
                            if (g[gi] <= thr) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    if (key[gi * 8 + e] <= thr) {
                                        if (ecnt < p.cap) { ebv[ecnt] = OSCALE * key[gi * 8 + e]; ebi[ecnt] = (int)(r0 + c0 + gi * 8 + e); }
                                        ++ecnt;
                                    }
                                }
                            }
                        }
                    }
                } else if (MODE == MODE_MINS) {
                    // sampling pass of the larger-k search: running minimum per column over mins_group sampled tiles.  The
                    // 32 rows behind one minimum are whole tiles apart, so neighbouring (similar) rows of a recording do not
                    // share a group and the k-th smallest minimum stays close to the k-th smallest key of the sample.
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if constexpr (HALFKEY) rmin[j] = fmaxf(rmin[j], v[j]);     // half key = -accumulator
                        else rmin[j] = fminf(rmin[j], key[j]);
                    }
                    if (c0 + 32 == (half + 1) * COLS && ((t + 1) % p.mins_group == 0 || t + 1 == ntiles)) {
                        if (q < p.nq) {
                            float *dst = p.odist + q * p.ldo +
                                         (((size_t)chunk * p.mins_per_chunk + t / p.mins_group) * EPI_SPLIT + half) * 32;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if constexpr (HALFKEY)
                                    *reinterpret_cast<float4 *>(dst + j) = make_float4(-OSCALE * rmin[j], -OSCALE * rmin[j + 1],
                                                                                       -OSCALE * rmin[j + 2], -OSCALE * rmin[j + 3]);
                                else
                                    *reinterpret_cast<float4 *>(dst + j) = make_float4(rmin[j], rmin[j + 1], rmin[j + 2], rmin[j + 3]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) rmin[j] = HALFKEY ? -INFINITY : INFINITY;
                    }
                } else if (MODE == MODE_STORE) {
                    if (q < p.nq) {
                        float *dst = p.odist + q * p.ldo + (r0 - p.row_lo) / p.tile_stride + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(key[j], key[j + 1], key[j + 2], key[j + 3]);
                    }
                } else {
                    // keys of this 32-column chunk, then a min tree: almost every chunk holds nothing that
                    // beats any lane's current list, and one vote dismisses it.  Otherwise votes narrow
                    // down to the 8-column group and the columns that matter (all votes of a level are
                    // independent instructions, so they pipeline) before any lane touches its list.
                    float g[4];
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        if constexpr (HALFKEY) {     // min of the half keys = -(max of the accumulators)
                            const float *a8 = v + 8 * gi;
                            g[gi] = -max3(max3(a8[0], a8[1], a8[2]), max3(a8[3], a8[4], a8[5]), fmaxf(a8[6], a8[7]));
                        } else {
                            const float *k8 = key + 8 * gi;
                            g[gi] = min3(min3(k8[0], k8[1], k8[2]), min3(k8[3], k8[4], k8[5]), fminf(k8[6], k8[7]));
                        }
                    }
                    const float cmin = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
                    if (__any_sync(0xffffffffu, cmin < lv[LSZ - 1])) {
                        unsigned gm = 0;
#pragma unroll
                        for (int gi = 0; gi < 4; ++gi)
                            gm |= (__ballot_sync(0xffffffffu, g[gi] < lv[LSZ - 1]) != 0u ? 1u : 0u) << gi;
#pragma unroll
                        for (int gi = 0; gi < 4; ++gi) {
                            if (gm & (1u << gi)) {
                                unsigned em = 0;
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    em |= (__ballot_sync(0xffffffffu, key[gi * 8 + e] < lv[LSZ - 1]) != 0u ? 1u : 0u) << e;
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    if (em & (1u << e)) {
                                        float kk = key[gi * 8 + e];
                                        if (kk < lv[LSZ - 1]) {
                                            int id = (int)(r0 + c0 + gi * 8 + e);
#pragma unroll
                                            for (int i = 0; i < LSZ; ++i) {   // sink into the ascending list; ties keep the lower row
                                                const bool c = kk < lv[i];
                                                const float lo = fminf(kk, lv[i]), hi = fmaxf(kk, lv[i]);
                                                const int ilo = c ? id : li[i], ihi = c ? li[i] : id;
                                                lv[i] = lo; li[i] = ilo; kk = hi; id = ihi;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (MODE == MODE_EMIT) {
            p.bufn[((size_t)q * p.nchunks + chunk) * EPI_SPLIT + half] = min(ecnt, p.cap);
            if (ecnt > p.cap && q < p.nq) p.tau[q] = -INFINITY;   // rows were lost: never certify
        }
        if (MODE == MODE_LIST) flush_lists();
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();     // the peer may still multicast into this CTA's ring / arrive on its barriers
    if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

// per query: tau = min over chunks of the chunk list's largest key (a row dropped inside a chunk has a
// key >= that chunk's LSZ-th smallest)
__global__ void chunk_tau_kernel(const float *__restrict__ oval, int64_t nq, int nlists, int lsz, float *__restrict__ tau) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        float t = INFINITY;
        for (int c = 0; c < nlists; ++c) t = fminf(t, oval[((size_t)q * nlists + c) * lsz + lsz - 1]);
        tau[q] = t;
    }
}
// ---- order statistics by bisection on the key bits: one warp per query, no lists, no sorting
// monotone map float -> uint32 (negative keys are possible: a key is ||y||^2 - 2 x.y)
__device__ __forceinline__ uint32_t key_bits(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// smallest t with #{u <= t} >= K over the n staged keys (n >= K >= 1)
__device__ __forceinline__ uint32_t warp_kth(const uint32_t *keys, int n, int K, int lane) {
    uint32_t lo = 0u, hi = 0xffffffffu;
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += keys[i] <= mid;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= K) hi = mid; else lo = mid + 1u;
    }
    return lo;
}
// the same with the keys in registers (n <= 32 PER).  Returns t such that the entries below t, then those equal to t in
// order, make up the K smallest: the exact order statistic, or mid + 1 of a round that counted exactly K at or below mid.
template <int PER>
__device__ __forceinline__ uint32_t warp_kth_exact_regs(const uint32_t *keys, int n, int K, int lane) {
    uint32_t key[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) key[j] = lane + 32 * j < n ? keys[lane + 32 * j] : 0xffffffffu;
    uint32_t lo = 0u, hi = 0xffffffffu;
#pragma unroll 1
    for (int it = 0; it < 32 && lo < hi; ++it) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) cnt += key[j] <= mid ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt == K && mid != 0xffffffffu) return mid + 1u;
        if (cnt >= K) hi = mid; else lo = mid + 1u;
    }
    return lo;
}
constexpr int SEL_WARPS = 4;
constexpr int SEL_CAP = 1024;      // keys staged in shared memory per query; longer rows bisect over global memory

// An upper bound on the k-th smallest of row q of vals [nq, ld] (n <= ld values; +inf if fewer than k are finite):
// thr[q] = tau[q] = a value with at least k and at most k + slack of the row's values at or below it.  The bound only has to
// be the key of SOME count >= k of actual rows (knn search: every order statistic of actual keys bounds the k-th key from
// above), so the bisection on the key bits stops as soon as the count lands in [k, k + slack] instead of running all 32
// rounds down to the exact order statistic; the keys sit in registers (PER per lane), one warp per query.
template <int PER>
__device__ __forceinline__ uint32_t warp_kth_regs(const float *__restrict__ row, int n, int K, int slack, int lane) {
    uint32_t key[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) key[j] = lane + 32 * j < n ? key_bits(__ldg(row + lane + 32 * j)) : 0xffffffffu;
    uint32_t lo = 0u, hi = 0xffffffffu;
#pragma unroll 1
    for (int it = 0; it < 32 && lo < hi; ++it) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) cnt += key[j] <= mid ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= K) {
            hi = mid;
            if (cnt <= K + slack) break;
        } else {
            lo = mid + 1u;
        }
    }
    return hi;        // count(keys <= hi) >= K always holds for hi
}

__global__ void __launch_bounds__(SEL_WARPS * 32) kth_of_rows_kernel(const float *__restrict__ vals, int64_t nq, int n,
                                                                     int64_t ld, int k, int slack, float *__restrict__ thr,
                                                                     float *__restrict__ tau) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * SEL_WARPS + w;
    if (q >= nq) return;
    const float *row = vals + q * ld;
    float out = INFINITY;
    if (n >= k) {
        uint32_t t;
        if (n <= 256) t = warp_kth_regs<8>(row, n, k, slack, lane);
        else if (n <= 512) t = warp_kth_regs<16>(row, n, k, slack, lane);
        else if (n <= 768) t = warp_kth_regs<24>(row, n, k, slack, lane);
        else if (n <= 1024) t = warp_kth_regs<32>(row, n, k, slack, lane);
        else {
            uint32_t lo = 0u, hi = 0xffffffffu;
#pragma unroll 1
            for (int it = 0; it < 32; ++it) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                int cnt = 0;
                for (int i = lane; i < n; i += 32) cnt += key_bits(__ldg(row + i)) <= mid;
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (cnt >= k) hi = mid; else lo = mid + 1u;
            }
            t = lo;
        }
        out = key_value(t);
        if (out != out) out = INFINITY;
    }
    if (lane == 0) { thr[q] = out; tau[q] = out; }
}

// The KP smallest (key, id) pairs of the emit buffers of query q (nseg buffers of cap slots, cnt[q, seg] filled), unsorted;
// missing entries: key +inf, id -1.  Ties at the KP-th key: first come.
__global__ void __launch_bounds__(SEL_WARPS * 32) select_segments_kernel(const float *__restrict__ bufv, const int *__restrict__ bufi,
                                                                         const int *__restrict__ bufn, int64_t nq, int nseg,
                                                                         int cap, int KP, float *__restrict__ oval,
                                                                         int *__restrict__ oid) {
    __shared__ uint32_t sk[SEL_WARPS][SEL_CAP];
    __shared__ int si[SEL_WARPS][SEL_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * SEL_WARPS + w;
    if (q >= nq) return;
    float *ov = oval + q * KP;
    int *oi = oid + q * KP;
    int tot = 0;
    for (int sg = 0; sg < nseg; ++sg) tot += min(cap, bufn[q * nseg + sg]);
    int nout = 0;
    if (tot <= SEL_CAP) {
        int base = 0;
        for (int sg = 0; sg < nseg; ++sg) {
            const int c = min(cap, bufn[q * nseg + sg]);
            const size_t off = ((size_t)q * nseg + sg) * cap;
            for (int i = lane; i < c; i += 32) {
                sk[w][base + i] = key_bits(__ldg(bufv + off + i));
                si[w][base + i] = __ldg(bufi + off + i);
            }
            base += c;
        }
        __syncwarp();
        if (tot <= KP) {
            for (int i = lane; i < tot; i += 32) { ov[i] = key_value(sk[w][i]); oi[i] = si[w][i]; }
            nout = tot;
        } else {
            // the keys go to registers for the bisection (no shared-memory read per round); a round whose count is exactly
            // KP ends it: everything at or below that value IS the answer
            uint32_t t;
            if (tot <= 256) t = warp_kth_exact_regs<8>(sk[w], tot, KP, lane);
            else if (tot <= 512) t = warp_kth_exact_regs<16>(sk[w], tot, KP, lane);
            else if (tot <= 768) t = warp_kth_exact_regs<24>(sk[w], tot, KP, lane);
            else t = warp_kth_exact_regs<32>(sk[w], tot, KP, lane);
            // everything below t, then keys equal to t until KP entries are out
            for (int pass = 0; pass < 2; ++pass) {
                for (int i0 = 0; i0 < tot && nout < KP; i0 += 32) {
                    const int i = i0 + lane;
                    const bool take = i < tot && (pass == 0 ? sk[w][i] < t : sk[w][i] == t);
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    const int pos = nout + __popc(m & ((1u << lane) - 1u));
                    if (take && pos < KP) { ov[pos] = key_value(sk[w][i]); oi[pos] = si[w][i]; }
                    nout = min(KP, nout + __popc(m));
                }
            }
        }
    } else {
        // rare: more rows at or below the bound than fit in shared memory -- bisect over the buffers in global memory
        uint32_t lo = 0u, hi = 0xffffffffu;
#pragma unroll 1
        for (int it = 0; it < 32; ++it) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int cnt = 0;
            for (int sg = 0; sg < nseg; ++sg) {
                const int c = min(cap, bufn[q * nseg + sg]);
                const size_t off = ((size_t)q * nseg + sg) * cap;
                for (int i = lane; i < c; i += 32) cnt += key_bits(__ldg(bufv + off + i)) <= mid;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt >= KP) hi = mid; else lo = mid + 1u;
        }
        const uint32_t t = lo;
        for (int pass = 0; pass < 2; ++pass)
            for (int sg = 0; sg < nseg; ++sg) {
                const int c = min(cap, bufn[q * nseg + sg]);
                const size_t off = ((size_t)q * nseg + sg) * cap;
                for (int i0 = 0; i0 < c && nout < KP; i0 += 32) {
                    const int i = i0 + lane;
                    const uint32_t u = i < c ? key_bits(__ldg(bufv + off + i)) : 0xffffffffu;
                    const bool take = i < c && (pass == 0 ? u < t : u == t);
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    const int pos = nout + __popc(m & ((1u << lane) - 1u));
                    if (take && pos < KP) { ov[pos] = key_value(u); oi[pos] = __ldg(bufi + off + i); }
                    nout = min(KP, nout + __popc(m));
                }
            }
    }
    for (int i = nout + lane; i < KP; i += 32) { ov[i] = INFINITY; oi[i] = -1; }
}
__global__ void fill_kernel(float *p, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------- host state
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tc_space_host {
    int nkb = 0, nload = 0, nsub = 0, stages = 0, sched = 0, b_bytes = 0;
    bool embed = false;
    size_t smem = 0;
    tc_load load[MAXLOAD];
    tc_sub sub[MAXSUB];
    short *d_qmap = nullptr;
    int ldq = 0;
    bool ok = false;
};
struct tc_state {
    encode_fn encode = nullptr;
    CUtensorMap mapS, mapG, mapGslab;
    tc_maps_mc mc;
    tc_space_host sp[2];
    size_t smem[2] = {0, 0};
};

int make_map(encode_fn enc, CUtensorMap *map, const void *base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
             uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SNK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

size_t aux_bytes(int sched) {   // barriers + tmem pointer (+ norm staging and descriptor table of the table-driven path)
    return 16 * MAX_STAGES + 8 + 64 + 24 + (sched == 0 ? 2 * BN * 4 + MAXSUB * 16 : 0) + 16;
}

int build_space(snk_db *db, int space, tc_space_host *h) {
    const int Dt = db->Dt;
    const int tblocks = (Dt + BK - 1) / BK;       // K-blocks per target frame
    const int m = space == SNK_SPACE_JOINT ? db->m : 1;
    const bool slab = m > 1 && m <= SLAB_EXTRA + 1 && !getenv("SNK_TC_NOSLAB");
    // squared norms embedded in three spare operand columns (weights.cu) when every row has them
    const bool embedG = Dt + 3 <= db->ldG16, embedS = db->Djq + 3 <= db->ldS16;
    h->embed = (space == SNK_SPACE_JOINT ? (embedG && embedS) : embedG) && !getenv("SNK_TC_NOEMBED");
    std::vector<short> qmap;
    h->nkb = h->nload = h->nsub = 0;
    bool overflow = false;
    // next resident query K-block: operand columns -> query dims; `valid` data columns, then (embedded
    // norms) three columns the query fills with -0.5 (marker -2)
    auto add_ablock = [&](int valid, int dim0, bool norm_cols) {
        for (int c = 0; c < BK; ++c)
            qmap.push_back(c < valid ? (short)(dim0 + c) : (norm_cols && c < valid + 3 ? (short)-2 : (short)-1));
        return h->nkb++;
    };
    auto add_load = [&](int map, int rowoff, int col, int bytes) {
        if (h->nload >= MAXLOAD) { overflow = true; return; }
        h->load[h->nload++] = tc_load{map, rowoff, col, bytes, h->nsub, 0};
    };
    auto add_sub = [&](int a_blk, int b_off, int base_off, int cols) {
        if (h->nsub >= MAXSUB || h->nload == 0) { overflow = true; return; }
        h->sub[h->nsub++] = tc_sub{a_blk, b_off, base_off, (cols + 15) / 16};
        h->load[h->nload - 1].nsub++;
    };
    if (space == SNK_SPACE_JOINT) {
        for (int c = 0; c < db->Djq; c += BK) {
            const int valid = std::min(BK, db->Djq - c);
            const bool last = c + BK >= db->Djq;
            add_load(0, 0, c, TILE_BYTES);
            add_sub(add_ablock(valid, c, h->embed && last), 0, 0, valid + (h->embed && last ? 3 : 0));
        }
    }
    const int dim0 = space == SNK_SPACE_JOINT ? db->Djq : 0;
    // query K-block order of the target part: frame-major, then column block
    std::vector<int> ablk((size_t)m * tblocks);
    for (int j = 0; j < m; ++j)
        for (int b = 0; b < tblocks; ++b)
            ablk[(size_t)j * tblocks + b] = add_ablock(std::min(BK, Dt - b * BK), dim0 + j * Dt + b * BK,
                                                       h->embed && b == tblocks - 1);
    for (int b = 0; b < tblocks; ++b) {
        const int cols = std::min(BK, Dt - b * BK) + (h->embed && b == tblocks - 1 ? 3 : 0);
        if (slab) {
            // K-block j starts j rows (j * 128 B) into the slab.  The swizzle atoms themselves stay
            // 1024-byte aligned (TMA wrote them), so the descriptor's base offset stays 0.
            add_load(2, 0, b * BK, SLOT_BYTES);
            for (int j = 0; j < m; ++j) add_sub(ablk[(size_t)j * tblocks + b], j * BK * 2, 0, cols);
        } else {
            for (int j = 0; j < m; ++j) {
                add_load(1, j, b * BK, TILE_BYTES);
                add_sub(ablk[(size_t)j * tblocks + b], 0, 0, cols);
            }
        }
    }
    h->ok = !overflow && h->nkb >= 1 && h->nkb <= MAXKB;
    if (!h->ok) return 0;
    h->ldq = h->nkb * BK;
    // statically scheduled variants (sched_traits): shapes of the shipped configs
    h->sched = 0;
    if (!getenv("SNK_TC_NOSCHED") && h->embed) {
        if (space == SNK_SPACE_JOINT && (slab || m == 1) && (m == 1 || m == 3 || m == 4 || m == 6) && tblocks == 1 &&
            Dt + 3 > 48 && db->Djq + 3 > 144 && db->Djq + 3 <= 160)
            // query operand in tensor memory: on by default where it was measured faster (multiepoch 6: 1387-1403 ->
            // 1492-1497 TFLOP/s; multiepoch 4 and 1 neutral, multiepoch 3 and the half-phone target shape slower --
            // their short tiles are bound by tensor-memory reads, which the A operand then competes for)
            h->sched = ((m == 6 && !getenv("SNK_TC_NOATMEM")) || getenv("SNK_TC_ATMEM") ? 20 : 10) + m;
        else if (space == SNK_SPACE_TARGET && tblocks == 3 && Dt + 3 > 176)
            h->sched = getenv("SNK_TC_ATMEM") ? 12 : 2;
    }
    const size_t budget = 227 * 1024;
    switch (h->sched) {
    case 2: h->b_bytes = sched_layout<2>::B_BYTES; h->stages = sched_layout<2>::NSLOT; break;
    case 11: h->b_bytes = sched_layout<11>::B_BYTES; h->stages = sched_layout<11>::NSLOT; break;
    case 13: h->b_bytes = sched_layout<13>::B_BYTES; h->stages = sched_layout<13>::NSLOT; break;
    case 14: h->b_bytes = sched_layout<14>::B_BYTES; h->stages = sched_layout<14>::NSLOT; break;
    case 16: h->b_bytes = sched_layout<16>::B_BYTES; h->stages = sched_layout<16>::NSLOT; break;
    case 12: h->b_bytes = sched_layout<12>::B_BYTES; h->stages = sched_layout<12>::NSLOT; break;
    case 21: h->b_bytes = sched_layout<21>::B_BYTES; h->stages = sched_layout<21>::NSLOT; break;
    case 23: h->b_bytes = sched_layout<23>::B_BYTES; h->stages = sched_layout<23>::NSLOT; break;
    case 24: h->b_bytes = sched_layout<24>::B_BYTES; h->stages = sched_layout<24>::NSLOT; break;
    case 26: h->b_bytes = sched_layout<26>::B_BYTES; h->stages = sched_layout<26>::NSLOT; break;

    default: break;
    }
    // query K-blocks that stay in shared memory (tensor-memory variants keep at most the last join block there)
    const int nkb_smem = h->sched == 26 ? 1 : (h->sched == 12 || h->sched > 20) ? 0 : h->nkb;
    if (h->sched != 0 && (size_t)nkb_smem * TILE_BYTES + h->b_bytes + aux_bytes(h->sched) > budget) h->sched = 0;
    if (h->sched == 0) {
        const size_t fixed = (size_t)h->nkb * TILE_BYTES + aux_bytes(0);
        if (fixed + 2 * SLOT_BYTES > budget) { h->ok = false; return 0; }
        h->stages = (int)std::min<size_t>(4, (budget - fixed) / SLOT_BYTES);
        h->b_bytes = h->stages * SLOT_BYTES;
    }
    h->smem = (size_t)(h->sched != 0 ? nkb_smem : h->nkb) * TILE_BYTES + h->b_bytes + aux_bytes(h->sched);
    SNK_CUDA(cudaMalloc((void **)&h->d_qmap, qmap.size() * sizeof(short)));
    SNK_CUDA(cudaMemcpy(h->d_qmap, qmap.data(), qmap.size() * sizeof(short), cudaMemcpyHostToDevice));
    return 0;
}

typedef void (*tc_kernel_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const tc_maps_mc,
                             const tc_params);

template <int MODE, int LSZ>
tc_kernel_fn pick_sched(int sched) {
    switch (sched) {
    case 2: return knn_tc_kernel<MODE, LSZ, 2>;
    case 11: return knn_tc_kernel<MODE, LSZ, 11>;
    case 13: return knn_tc_kernel<MODE, LSZ, 13>;
    case 14: return knn_tc_kernel<MODE, LSZ, 14>;
    case 16: return knn_tc_kernel<MODE, LSZ, 16>;
    case 12: return knn_tc_kernel<MODE, LSZ, 12>;
    case 21: return knn_tc_kernel<MODE, LSZ, 21>;
    case 23: return knn_tc_kernel<MODE, LSZ, 23>;
    case 24: return knn_tc_kernel<MODE, LSZ, 24>;
    case 26: return knn_tc_kernel<MODE, LSZ, 26>;

    default: return knn_tc_kernel<MODE, LSZ, 0>;
    }
}
tc_kernel_fn pick_kernel(int mode, int lsz, int sched) {
    if (mode == MODE_STORE) return pick_sched<MODE_STORE, 4>(sched);
    if (mode == MODE_EMIT) return pick_sched<MODE_EMIT, 4>(sched);
    if (mode == MODE_MINS) return pick_sched<MODE_MINS, 4>(sched);
    return lsz == 4 ? pick_sched<MODE_LIST, 4>(sched) : pick_sched<MODE_LIST, 8>(sched);
}

// launches with a thread-block cluster of p.cluster CTAs (consecutive blockIdx.x) when the multicast path is on
int launch_tc(tc_kernel_fn fn, int grid, int threads, size_t smem, cudaStream_t st, const CUtensorMap &mapQ,
              const tc_state *s, const tc_params &p) {
    if (p.cluster <= 1) {
        fn<<<grid, threads, smem, st>>>(mapQ, s->mapS, s->mapG, s->mapGslab, s->mc, p);
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)p.cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    void *args[] = {(void *)&mapQ, (void *)&s->mapS, (void *)&s->mapG, (void *)&s->mapGslab, (void *)&s->mc, (void *)&p};
    SNK_CUDA(cudaLaunchKernelExC(&cfg, (const void *)fn, args));
    return 0;
}

// work split of one launch: query tiles (padded to the cluster size) x database chunks
struct tc_split { int cluster, nqt_pad, nchunks; };
tc_split make_split(const snk_db *db, const tc_space_host &h, int nqt, int64_t tiles) {
    tc_split sp;
    // Pairs of CTAs sharing every database tile by TMA multicast halve the L2 reads.  Measured neutral on B200
    // (m = 6: 1418 vs 1394 TFLOP/s, m = 4: 1278 vs 1329): these kernels are not L2-bound, so it is opt-in.
    sp.cluster = (h.sched != 0 && h.sched != 12 && h.sched < 20 && nqt >= 2 && getenv("SNK_TC_CLUSTER")) ? 2 : 1;
    sp.nqt_pad = (int)snk_round_up(nqt, sp.cluster);
    sp.nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(db->sm_count / std::max(sp.nqt_pad, 1), tiles));
    if (sp.nqt_pad > db->sm_count) sp.nchunks = 1;
    return sp;
}

}  // namespace

int snk_tc_prepare(snk_db *db) {
    tc_state *s = new tc_state();
    db->tc_state = s;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        snk_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    s->encode = (encode_fn)fn;
    SNK_TRY(make_map(s->encode, &s->mapS, db->S16, (uint64_t)db->ldS16, (uint64_t)db->N + 1, (uint64_t)db->ldS16, BN));
    SNK_TRY(make_map(s->encode, &s->mapG, db->G16, (uint64_t)db->ldG16, (uint64_t)db->N, (uint64_t)db->ldG16, BN));
    SNK_TRY(make_map(s->encode, &s->mapGslab, db->G16, (uint64_t)db->ldG16, (uint64_t)db->N, (uint64_t)db->ldG16,
                     BN + SLAB_EXTRA));
    // half boxes of the multicast (cluster of two) path
    SNK_TRY(make_map(s->encode, &s->mc.S_h, db->S16, (uint64_t)db->ldS16, (uint64_t)db->N + 1, (uint64_t)db->ldS16, BN / 2));
    SNK_TRY(make_map(s->encode, &s->mc.G_h, db->G16, (uint64_t)db->ldG16, (uint64_t)db->N, (uint64_t)db->ldG16, BN / 2));
    SNK_TRY(make_map(s->encode, &s->mc.Gslab72, db->G16, (uint64_t)db->ldG16, (uint64_t)db->N, (uint64_t)db->ldG16, MC_ROWS0));
    for (int sp = 0; sp < 2; ++sp) {
        SNK_TRY(build_space(db, sp, &s->sp[sp]));
        if (s->sp[sp].ok) s->smem[sp] = s->sp[sp].smem;
    }
    for (int sched : {0, 2, 11, 13, 14, 16, 12, 21, 23, 24, 26})
        for (int v = 0; v < 5; ++v)
            SNK_CUDA(cudaFuncSetAttribute((const void *)pick_kernel(v == 2 ? MODE_STORE : v == 3 ? MODE_EMIT : v == 4 ? MODE_MINS : MODE_LIST,
                                                                    v == 1 ? 8 : 4, sched),
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return 0;
}

void snk_tc_destroy(snk_db *db) {
    tc_state *s = (tc_state *)db->tc_state;
    if (!s) return;
    for (int sp = 0; sp < 2; ++sp)
        if (s->sp[sp].d_qmap) cudaFree(s->sp[sp].d_qmap);
    delete s;
    db->tc_state = nullptr;
}

bool snk_tc_supported(const snk_db *db, const snk_space &sp, int KP) {
    const tc_state *s = (const tc_state *)db->tc_state;
    if (!s) return false;
    const int space = sp.dA > 0 ? SNK_SPACE_JOINT : SNK_SPACE_TARGET;
    (void)KP;
    return db->tc_ok && s->sp[space].ok && sp.rows >= 1;
}

// Relative slack eps_rel of the certificate: |(qn + key) - ||x~ - y~||^2| <= eps_rel (||x~||^2 + 2 max||y~||^2).
//  * key = -2 acc, acc = sum of the exact fp16 products (x~.y~ and the embedded -0.5 ||y~||^2 pieces) accumulated in fp32
//    over `nsteps` UMMA K = 16 instructions.  Assumption (checked by tests/test_gpu_certificate.py, which measures the
//    constant): one instruction adds at most two roundings of 2^-23 (truncation) relative to S = sum |x_i y_i| + ||y~||^2 / 2
//    <= (||x~||^2 + ||y~||^2) / 2 + ||y~||^2 / 2, so |key error| <= 2 (nsteps + 1) 2^-23 (||x~||^2 + 2 ||y~||^2).
//  * the fp32 squared norms (weights.cu, search.cu: one fma chain of <= D/32 terms per lane, shuffle tree, window sum):
//    relative error <= (D/32 + 16) 2^-24 each.
float snk_tc_eps_rel(const snk_db *db, int space) {
    const tc_space_host &h = ((const tc_state *)db->tc_state)->sp[space];
    int nsteps = 0;
    for (int i = 0; i < h.nsub; ++i) nsteps += h.sub[i].ksteps;
    const double g_acc = 2.0 * (nsteps + 1) * 1.1920928955078125e-7;
    const double g_norm = (h.ldq / 32.0 + 16.0) * 5.9604644775390625e-8;
    return (float)(1.02 * (g_acc + g_norm));
}

// Raw tensor-core keys (||y~||^2 - 2 x~.y~, fp32) of every query against rows [row0, row0 + nrows): the store epilogue
// without any selection.  Test instrumentation for the certificate's error bound (snk_debug_tc_keys).
int snk_tc_debug_keys(snk_db *db, int space, const __half *dQ16, int ldq16, int64_t nq, int64_t row0, int64_t nrows,
                      float *d_keys, int64_t ld, cudaStream_t st) {
    tc_state *s = (tc_state *)db->tc_state;
    const tc_space_host &h = s->sp[space];
    SNK_CHECK(h.ok, "tensor-core engine does not support this search space");
    SNK_CHECK(ld % BN == 0 && ld >= snk_round_up(nrows, BN), "key matrix leading dimension must be a multiple of %d", BN);
    const int64_t nq_pad = snk_round_up(nq, 2 * BM);
    const int nqt = (int)snk_cdiv(nq, BM);
    CUtensorMap mapQ;
    SNK_TRY(make_map(s->encode, &mapQ, dQ16, (uint64_t)ldq16, (uint64_t)nq_pad, (uint64_t)ldq16, BM));
    tc_params p;
    memset(&p, 0, sizeof(p));
    p.nkb = h.nkb; p.nload = h.nload; p.nsub = h.nsub; p.stages = h.stages;
    p.embed = h.embed ? 1 : 0; p.b_bytes = h.b_bytes; p.cluster = 1;
    memcpy(p.load, h.load, sizeof(h.load));
    memcpy(p.sub, h.sub, sizeof(h.sub));
    p.nq = nq; p.q16 = dQ16; p.ldq16 = ldq16; p.tile_stride = 1;
    p.nrm = space == SNK_SPACE_JOINT ? db->nrm_j16 : db->nrm_t16;
    const int64_t tiles = snk_cdiv(nrows, BN);
    const tc_split ws = make_split(db, h, nqt, tiles);
    p.row_lo = row0; p.row_hi = row0 + nrows; p.nchunks = ws.nchunks;
    p.chunk_rows = snk_cdiv(tiles, ws.nchunks) * BN;
    p.odist = d_keys; p.ldo = ld;
    SNK_TRY(launch_tc(pick_kernel(MODE_STORE, 4, h.sched), ws.nqt_pad * ws.nchunks, num_threads_of(h.sched), s->smem[space], st,
                      mapQ, s, p));
    SNK_CUDA(cudaGetLastError());
    return 0;
}

int snk_tc_query_ld(const snk_db *db, int space) { return ((const tc_state *)db->tc_state)->sp[space].ldq; }
const short *snk_tc_qmap(const snk_db *db, int space) { return ((const tc_state *)db->tc_state)->sp[space].d_qmap; }

int snk_shortlist_tc(snk_db *db, int space, const __half *dQ16, int ldq16, int64_t nq, int k, int KP, float *d_val,
                     int *d_id, float *d_tau, snk_tc_lists *lists, cudaStream_t st) {
    if (lists) lists->valid = false;
    tc_state *s = (tc_state *)db->tc_state;
    const tc_space_host &h = s->sp[space];
    SNK_CHECK(h.ok, "tensor-core engine does not support this search space");
    const snk_space sp = snk_make_space(db, space);
    const int64_t nq_pad = snk_round_up(nq, 2 * BM);      // room for the padded second query tile of a cluster
    const int nqt = (int)snk_cdiv(nq, BM);
    CUtensorMap mapQ;
    SNK_TRY(make_map(s->encode, &mapQ, dQ16, (uint64_t)ldq16, (uint64_t)nq_pad, (uint64_t)ldq16, BM));
    tc_params p;
    memset(&p, 0, sizeof(p));
    p.nkb = h.nkb; p.nload = h.nload; p.nsub = h.nsub; p.stages = h.stages;
    p.embed = h.embed ? 1 : 0; p.b_bytes = h.b_bytes; p.cluster = 1;
    memcpy(p.load, h.load, sizeof(h.load));
    memcpy(p.sub, h.sub, sizeof(h.sub));
    p.nq = nq;
    p.q16 = dQ16; p.ldq16 = ldq16;
    p.tile_stride = 1;
    p.nrm = space == SNK_SPACE_JOINT ? db->nrm_j16 : db->nrm_t16;
    const size_t smem = s->smem[space];
    const int64_t row_tiles = snk_cdiv(sp.rows, BN);
    // Fused lists hold the LSZ best keys per (query, chunk, column half); they cover the k best with slack
    // (anything they miss is caught by the certificate).  Larger k goes through the key scan below.
    const bool fused = k <= 4;
    if (fused) {
        const int lsz = k <= 2 ? 4 : 8;
        const tc_split ws = make_split(db, h, nqt, row_tiles);
        const int nchunks = ws.nchunks;
        p.cluster = ws.cluster;
        p.row_lo = 0; p.row_hi = sp.rows; p.nchunks = nchunks;
        p.chunk_rows = snk_cdiv(row_tiles, nchunks) * BN;
        const int nlists = nchunks * epi_split_of(h.sched);
        const size_t nlist = (size_t)nq_pad * nlists * lsz;
        SNK_TRY(snk_buf_reserve(&db->ws_tc, nlist * 8));
        p.oval = (float *)db->ws_tc.p;
        p.oid = (int *)(p.oval + nlist);
        {
            snk_prof_scope prof(db, SNK_PROF_KNN, 2.0 * (double)nq * (double)sp.rows * sp.D, st);
            SNK_TRY(launch_tc(pick_kernel(MODE_LIST, lsz, h.sched), ws.nqt_pad * nchunks, num_threads_of(h.sched), smem, st,
                              mapQ, s, p));
        }
        SNK_CUDA(cudaGetLastError());
        if (lists && snk_merge_rerank_fits(nlists, lsz)) {   // merge + tau happen inside the re-rank kernel
            lists->valid = true; lists->val = p.oval; lists->id = p.oid; lists->nlists = nlists; lists->lsz = lsz;
            db->counters[2] += 1;
            return 0;
        }
        chunk_tau_kernel<<<64, 256, 0, st>>>(p.oval, nq, nlists, lsz, d_tau);
        SNK_CUDA(cudaGetLastError());
        db->counters[2] += 2;
        for (int64_t q0 = 0; q0 < nq; q0 += 32768) {
            const int64_t n = std::min<int64_t>(32768, nq - q0);
            SNK_TRY(snk_topk_scan(db, p.oval + (size_t)q0 * nlists * lsz, p.oid + (size_t)q0 * nlists * lsz, n,
                                  (int64_t)nlists * lsz, (int64_t)nlists * lsz, 0, KP, true, d_val + q0 * KP,
                                  d_id + q0 * KP, st));
        }
        return 0;
    }
    // ---- larger k.  store_scan: keys of a row span (every `stride`-th tile of it) go to HBM, the warp scan
    // keeps the KPs smallest per query.
    auto store_scan = [&](int stride, int KPs, float *o_val, int *o_id) -> int {
        const size_t WS = (size_t)512 << 20;
        int64_t span_cols = (int64_t)(WS / 4 / (size_t)nq_pad) / BN * BN;
        span_cols = std::max<int64_t>(BN, std::min<int64_t>(span_cols, snk_cdiv(row_tiles, stride) * BN));
        const int64_t span_rows = span_cols * stride;
        SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)nq_pad * span_cols * 4));
        p.odist = (float *)db->ws_dist.p;
        p.ldo = span_cols;
        p.tile_stride = stride;
        bool first = true;
        for (int64_t rb = 0; rb < sp.rows; rb += span_rows) {
            const int64_t rn = std::min<int64_t>(span_rows, sp.rows - rb);
            const int64_t tiles = snk_cdiv(rn, (int64_t)BN * stride);      // sampled tiles of this span
            const tc_split ws = make_split(db, h, nqt, tiles);
            const int nchunks = ws.nchunks;
            p.cluster = ws.cluster;
            p.row_lo = rb; p.row_hi = rb + rn; p.nchunks = nchunks;
            p.chunk_rows = snk_cdiv(tiles, nchunks) * BN * stride;
            {
                snk_prof_scope prof(db, SNK_PROF_KNN, 2.0 * (double)nq * (double)tiles * BN * sp.D, st);
                SNK_TRY(launch_tc(pick_kernel(MODE_STORE, 4, h.sched), ws.nqt_pad * nchunks, num_threads_of(h.sched), smem, st,
                                  mapQ, s, p));
            }
            SNK_CUDA(cudaGetLastError());
            db->counters[2] += 1;
            for (int64_t q0 = 0; q0 < nq; q0 += 32768) {
                const int64_t n = std::min<int64_t>(32768, nq - q0);
                SNK_TRY(snk_topk_scan(db, p.odist + q0 * span_cols, nullptr, n, tiles * BN, span_cols, (int)rb, KPs, first,
                                      o_val + q0 * KPs, o_id + q0 * KPs, st));
            }
            first = false;
        }
        p.tile_stride = 1;
        return 0;
    };

    // Sampled threshold + emit: the k-th smallest key of a sample of rows bounds the k-th smallest key overall
    // from above; a second pass over the whole database appends every row at or below that bound to
    // per-(query, chunk, half) buffers (a few k rows per query), and only those are scanned and re-ranked.
    // Rows above the bound cannot be among the k nearest, so the bound itself is the certificate's tau.
    // The sampling pass is a max tree per 32 keys, so it runs at the speed of the contraction: every 4th tile costs a
    // quarter of a pass and halves the emitted rows against every 8th (SNK_TC_SAMPLE overrides).
    // never every tile: a bound that hugs the k-th key leaves the certificate no room for the fp16 rounding of the keys
    int SAMPLE = getenv("SNK_TC_SAMPLE") ? std::max(2, atoi(getenv("SNK_TC_SAMPLE"))) : 4;
    while (SAMPLE > 2 && snk_cdiv(row_tiles, SAMPLE) * BN < 4 * (2 * k + 16)) SAMPLE /= 2;
    if (!getenv("SNK_TC_NOEMIT") && snk_cdiv(row_tiles, SAMPLE) * BN >= 2 * (2 * k + 16) && sp.rows >= (int64_t)64 * k) {
        {
            // Sampling pass: column-wise minima over groups of sampled tiles (every SAMPLE-th).  Each of them is the key of an
            // actual row, and the k-th smallest of ANY set of actual rows bounds the k-th smallest overall from above.
            const int64_t stiles = snk_cdiv(row_tiles, SAMPLE);
            const tc_split ss = make_split(db, h, nqt, stiles);
            const int64_t chunk_tiles = snk_cdiv(stiles, ss.nchunks);
            const int split = epi_split_of(h.sched);
            // tiles per group so that a query ends up with about 768 minima (at least 2k + 16)
            int group = (int)std::max<int64_t>(1, std::min<int64_t>(64, stiles * split * 32 / 768));
            while (group > 1 && snk_cdiv(chunk_tiles, group) * ss.nchunks * split * 32 < 2 * k + 16) --group;
            p.mins_group = group;
            p.mins_per_chunk = (int)snk_cdiv(chunk_tiles, group);
            const int64_t nmin = (int64_t)ss.nchunks * p.mins_per_chunk * split * 32;   // minima per query (groups no tile reaches: +inf)
            p.cluster = ss.cluster;
            p.tile_stride = SAMPLE;
            p.row_lo = 0; p.row_hi = sp.rows; p.nchunks = ss.nchunks;
            p.chunk_rows = chunk_tiles * BN * SAMPLE;
            SNK_TRY(snk_buf_reserve(&db->ws_dist, (size_t)nq_pad * nmin * 4));
            p.odist = (float *)db->ws_dist.p; p.ldo = nmin;
            fill_kernel<<<db->sm_count * 4, 256, 0, st>>>(p.odist, nq * nmin, INFINITY);   // groups no tile covers
            SNK_CUDA(cudaGetLastError());
            {
                snk_prof_scope prof(db, SNK_PROF_KNN, 2.0 * (double)nq * (double)stiles * BN * sp.D, st);
                SNK_TRY(launch_tc(pick_kernel(MODE_MINS, 4, h.sched), ss.nqt_pad * ss.nchunks, num_threads_of(h.sched), smem,
                                  st, mapQ, s, p));
            }
            SNK_CUDA(cudaGetLastError());
            db->counters[2] += 2;
            SNK_TRY(snk_buf_reserve(&db->ws_misc, (size_t)nq_pad * 4));
            // The bound is the (2k)-th smallest minimum, not the k-th: the certificate needs the k-th exact distance to stay
            // below the bound by the fp16 rounding of the keys, and a bound that hugs the k-th key fails it for a few
            // queries in 10^5 (measured; each failure costs a re-search).  SNK_TC_KMARGIN overrides the factor.
            const int margin = getenv("SNK_TC_KMARGIN") ? std::max(1, atoi(getenv("SNK_TC_KMARGIN"))) : 2;
            const int kth = (int)std::min<int64_t>(nmin, (int64_t)k * margin);
            // a bound a few ranks above the (2k)-th minimum emits a few rows more and saves a third of the bisection rounds
            const int slack = getenv("SNK_TC_KSLACK") ? std::max(0, atoi(getenv("SNK_TC_KSLACK"))) : std::max(1, kth / 8);
            kth_of_rows_kernel<<<(unsigned)snk_cdiv(nq, SEL_WARPS), SEL_WARPS * 32, 0, st>>>(p.odist, nq, (int)nmin, nmin, kth, slack,
                                                                                          (float *)db->ws_misc.p, d_tau);
            SNK_CUDA(cudaGetLastError());
            db->counters[2] += 1;
            p.tile_stride = 1; p.odist = nullptr; p.ldo = 0;
        }
        float *thr = (float *)db->ws_misc.p;
        const tc_split ws = make_split(db, h, nqt, row_tiles);
        const int nchunks = ws.nchunks;
        p.cluster = ws.cluster;
        const int nlists = nchunks * epi_split_of(h.sched);
        // neighbours cluster on a few consecutive rows (trajectories), so one list may take most of the
        // expected rows and an unsampled tile may hide a whole cluster: size every list for 20k rows (measured:
        // 10k overflowed for 0.17 % of the queries, 20k for 0.001 %)
        const int cap = getenv("SNK_TC_EMIT_CAP") ? atoi(getenv("SNK_TC_EMIT_CAP")) :
                        (int)snk_round_up(std::min<int64_t>(2048, std::max<int64_t>(256, (int64_t)20 * k)), 32);
        const size_t nent = (size_t)nq_pad * nlists * cap;
        SNK_TRY(snk_buf_reserve(&db->ws_tc, nent * 8 + (size_t)nq_pad * nlists * 4));
        p.bufv = (float *)db->ws_tc.p;
        p.bufi = (int *)(p.bufv + nent);
        p.bufn = p.bufi + nent;
        p.cap = cap; p.thr = thr; p.tau = d_tau;
        p.row_lo = 0; p.row_hi = sp.rows; p.nchunks = nchunks;
        p.chunk_rows = snk_cdiv(row_tiles, nchunks) * BN;
        // every (query, chunk, part) buffer gets its fill count from the kernel: no clearing, and the scan
        // below touches only filled slots
        {
            snk_prof_scope prof(db, SNK_PROF_KNN, 2.0 * (double)nq * (double)sp.rows * sp.D, st);
            SNK_TRY(launch_tc(pick_kernel(MODE_EMIT, 4, h.sched), ws.nqt_pad * nchunks, num_threads_of(h.sched), smem, st, mapQ,
                              s, p));
        }
        SNK_CUDA(cudaGetLastError());
        db->counters[2] += 3;
        select_segments_kernel<<<(unsigned)snk_cdiv(nq, SEL_WARPS), SEL_WARPS * 32, 0, st>>>(p.bufv, p.bufi, p.bufn, nq, nlists, cap,
                                                                                           KP, d_val, d_id);
        SNK_CUDA(cudaGetLastError());
        db->counters[2] += 1;
        return 0;
    }
    // small databases: plain store + scan of every key
    fill_kernel<<<64, 256, 0, st>>>(d_tau, nq, INFINITY);
    SNK_CUDA(cudaGetLastError());
    SNK_TRY(store_scan(1, KP, d_val, d_id));
    return 0;
}
