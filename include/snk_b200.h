/*
 * snk_b200.h -- C ABI of the B200-native unit-selection search engine.
 *
 * The reference (CSTR-Edinburgh/snickery) has no FFI: its seam for this path is a
 * set of Python calls on scipy/sklearn KD-trees, numpy and OpenFst objects
 * (SURVEY.md section 8b).  Each entry point below names the reference call site it
 * replaces (paths under /root/reference/script/).  INTEGRATION.md shows the ctypes
 * stub a maintainer would add on the reference side.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; snk_last_error()
 *    returns a thread-local, NUL-terminated description of the last failure.
 *  - no exceptions cross the ABI; no torch types appear in signatures.
 *  - "host" entry points take plain host pointers (pageable or pinned), perform
 *    H2D / D2H copies themselves and return when results are in host memory.
 *  - "_dev" entry points take device pointers on the handle's device and enqueue
 *    on the given cudaStream_t (passed as void*); they NEVER synchronise: launch
 *    metadata travels through a pinned staging ring, exactness certificates stay on
 *    the device.  The searches (snk_knn_dev, snk_greedy_batch*_dev, snk_knn_sharded_dev)
 *    are completed by the matching *_finish call, which waits for the stream, reads the
 *    certificate flags and repairs the (rare) uncertified answers; results are final
 *    only after it.  Input and output buffers must stay alive until then.  All _dev
 *    calls on one handle share its workspaces: issue them on ONE stream at a time.
 *    The one exception to "never": the handle's grow-only device workspaces are
 *    (re)allocated when a call needs more than any call before it (cudaFree /
 *    cudaMalloc wait for the device), i.e. during the first calls of a given size.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with an error.
 *  - thread-safe per handle: calls on one handle are serialised by an internal lock,
 *    different handles run concurrently.  A handle must not be shared across fork().
 */
#ifndef SNK_B200_H
#define SNK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snk_db snk_db;

/* join-context layouts (what prev_join_rep / current_join_rep mean for greedy search) */
#define SNK_LAYOUT_SIMPLE 0u      /* synth_simple.py:194-195,213-214: prev=start[u]=Jw[u], cur=end[u+m-1]=Jw[u+m]          */
#define SNK_LAYOUT_HALFPHONE_EPOCH 1u /* synth_halfphone.py:552-553,580-581: prev=Jw[u][:Dj/2], cur=Jw[u+m-1][Dj/2:]        */

/* search spaces for snk_knn */
#define SNK_SPACE_TARGET 0 /* rows = weighted train_unit_features [N, Dt]      (synth_halfphone.py:379,1364) */
#define SNK_SPACE_JOINT 1  /* rows = [prev_join_rep || multiepoch window] [N', Djq+m*Dt] (synth_simple.py:224-229) */

/* shortlist engines (SNK_ENGINE_AUTO picks the tensor-core kernel when shapes allow) */
#define SNK_ENGINE_AUTO 0
#define SNK_ENGINE_SIMT 1 /* fp32 direct-difference distances on CUDA cores + top-k scan              */
#define SNK_ENGINE_TC 2   /* fp16 tcgen05 distance GEMM with fused per-query top-k in the epilogue    */

const char *snk_last_error(void);
int snk_version(void);
/* number of visible CUDA devices (0 if none; never an error) */
int snk_device_count(void);

/* ---- database lifetime ---------------------------------------------------------------
 * Replaces: Synthesiser.__init__ array loads (synth_simple.py:76-106) + the cKDTree
 * constructors (synth_simple.py:229; synth_halfphone.py:379,605).  F and Jc are the
 * standardised, UNWEIGHTED float32 arrays exactly as stored in the voice HDF5
 * (train_simple.py:137,142,278-289): F [N, Dt], Jc [N+1, Dj].  They are copied.      */
int snk_db_create(snk_db **out, int device_id, int64_t N, int Dt, int Dj, int multiepoch,
                  const float *F, const float *Jc, unsigned layout_flags);
int snk_db_destroy(snk_db *db);
int snk_db_info(const snk_db *db, int64_t *N, int64_t *Nprime, int *Dt, int *Dj, int *multiepoch,
                int *joint_dim);
/* Replaces set_target_weights + set_join_weights + get_tree_for_greedy_search
 * (synth_simple.py:234-274,190-230): per-coefficient float64 weight vectors wt [Dt],
 * wj [Dj].  No tree is built: the "index" is the weighted matrices themselves.       */
int snk_db_set_weights(snk_db *db, const double *wt, const double *wj);
/* choose the shortlist engine for subsequent searches (default AUTO) */
int snk_db_set_engine(snk_db *db, int engine);
/* counters since creation: [0] queries searched, [1] queries (greedy: utterances) whose tensor-core
 * shortlist failed the exactness certificate and were re-searched with the fp32 SIMT engine,
 * [2] kernels launched, [3] queries (utterances) that also failed the fp32 certificate and were
 * answered by the exhaustive float64 scan */
int snk_db_counters(const snk_db *db, int64_t counters[4], int reset);

/* ---- kernel timing for bench.py -----------------------------------------------------------
 * When enabled, CUDA events bracket every launch of the three hot kernels on the launching
 * stream.  snk_db_profile_read synchronises, then returns the summed device time, the number
 * of launches and the algorithmic work (SNK_PROF_KNN: FLOPs 2*nq*rows*D; SNK_PROF_JOIN /
 * SNK_PROF_VITERBI: bytes) since the last reset.                      */
#define SNK_PROF_KNN 0
#define SNK_PROF_JOIN 1
#define SNK_PROF_VITERBI 2
#define SNK_PROF_ALLGATHER 3 /* work = bytes received over NVLink by this rank */
#define SNK_PROF_MERGE 4     /* work = bytes merged */
#define SNK_PROF_JOIN_VITERBI 5 /* reserved (a fused join-cost + Viterbi kernel was measured and not kept) */
#define SNK_PROF_RERANK 6    /* float64 re-rank of the shortlists; work = gathered bytes */
int snk_db_profile_enable(snk_db *db, int enable);
int snk_db_profile_read(snk_db *db, int which, double *total_ms, int64_t *launches, double *work, int reset);

/* ---- test instrumentation of the exactness certificate -----------------------------------------
 * The tensor-core kernel's raw keys ||y~||^2 - 2 x~.y~ (fp32) for queries Q [nq, D] (float64, host)
 * against rows [row0, row0 + nrows) -> keys [nq, nrows]; qnorm [nq] = the fp32 ||x~||^2 the
 * certificate adds; *eps_rel = the relative slack it allows, *maxnorm = max_u ||y~_u||^2.
 * tests/test_gpu_certificate.py compares key + qnorm with the float64 distance of the rounded
 * operands and checks the measured error against eps_rel (||x~||^2 + 2 maxnorm).           */
int snk_debug_tc_keys(snk_db *db, int space, const double *Q, int64_t nq, int64_t row0, int64_t nrows,
                      float *keys, float *qnorm, float *eps_rel, float *maxnorm);
/* The same for the single-utterance greedy kernel (greedy_one.cu, mma.sync keys): keys [N'] of the FIRST step of an
 * utterance whose first window is targets [multiepoch, Dt] (weighted float64, host) after start_state (-1: none);
 * *qnorm = the fp32 ||x~||^2 of that query.                                                     */
int snk_debug_greedy_one_keys(snk_db *db, const double *targets, int64_t start_state, float *keys, float *qnorm,
                              float *eps_rel, float *maxnorm);
/* Diagnostic: with SNK_G1_TIMING=1 in the environment the single-utterance kernel records CTA 0's %globaltimer at up to 16
 * points of every step; out [steps][16] nanoseconds of the last launch (tests/multigpu/probe_single.py prints the split). */
int snk_debug_greedy_one_times(snk_db *db, unsigned long long *out, int steps);
/* ... and every CTA's timestamp at the end of its scan, out [steps <= 256][SM count]: the skew the grid barrier waits for */
int snk_debug_greedy_one_cta_times(snk_db *db, unsigned long long *out, int steps);

/* ---- k-NN: tree.query(X, k) ------------------------------------------------------------
 * Replaces cKDTree.query / sklearn KDTree.query (synth_halfphone.py:1364,1384;
 * synth_simple.py:490; StashableKDTree.py).  Q is float64 [nq, D] already weighted like
 * the reference's queries.  Outputs: Euclidean (sqrt) float64 distances ascending and
 * int64 row ids, [nq, k].  If k exceeds the number of rows the tail is (inf, nrows) as
 * scipy does.  Ties: lowest row id first.                                              */
int snk_knn(snk_db *db, int space, const double *Q, int64_t nq, int k, double *dist, int64_t *idx);
int snk_knn_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist,
                int64_t *d_idx, int64_t id_offset, void *stream);
/* completes every snk_knn_dev enqueued on this handle since the last call (see "Conventions") */
int snk_knn_finish(snk_db *db);
/* k-way merge of R per-shard results [R, nq, k] (ascending) into [nq, k]; used after the
 * NCCL all-gather of the sharded-database search.               */
int snk_topk_merge_dev(int device_id, const double *d_dist_all, const int64_t *d_idx_all, int R,
                       int64_t nq, int k, double *d_dist, int64_t *d_idx, void *stream);

/* ---- database sharded over the GPUs of one box --------------------------
 * No reference call site: the reference is one process.  Every rank (one process per GPU) holds a
 * block of database rows in its own handle and calls these collectively.
 * snk_comm_unique_id: rank 0 obtains the 128-byte NCCL id and hands it to the other ranks through
 * the host's own channel (MPI, a file, torch.distributed...).  snk_comm_init attaches an NCCL
 * communicator to the handle.  NCCL is resolved with dlopen("libnccl.so.2") at that moment.
 * snk_knn_sharded_dev: replicated queries dQ [nq, D]; local certified top-k with global ids
 * (id_offset = first global row of this rank's block), ONE ncclAllGather of nq*k*16 bytes per rank
 * over NVLink, k-way merge on the device (ties: lowest global id); every rank gets the global
 * [nq, k] answer.  snk_knn_sharded_finish (collective) completes it like snk_knn_finish.      */
#define SNK_UNIQUE_ID_BYTES 128
int snk_comm_unique_id(void *id_out, int id_bytes);
int snk_comm_init(snk_db *db, const void *unique_id, int rank, int nranks);
int snk_comm_info(snk_db *db, int *rank, int *nranks, int *nccl_version);
/* 1 if the per-step exchange of snk_greedy_sharded_batch_dev runs over NVLink peer memory (every rank's exchange region
 * is mapped into the others through CUDA IPC at snk_comm_init; one kernel per step stores into all peers, publishes an
 * epoch and waits for the others'), 0 if it goes through ncclAllGather (a platform without IPC between the processes,
 * SNK_COMM_NO_P2P=1, or more than 4096 utterances per step).                                                          */
int snk_comm_peer_exchange(const snk_db *db);
int snk_knn_sharded_dev(snk_db *db, int space, const double *dQ, int64_t nq, int k, double *d_dist,
                        int64_t *d_idx, int64_t id_offset, void *stream);
int snk_knn_sharded_finish(snk_db *db);

/* ---- greedy joint search ------------------------------------------------------------------
 * Replaces Synthesiser.greedy_joint_search (synth_simple.py:458-503;
 * synth_halfphone.py:1900-1945) for a batch of B utterances.  targets: float64, the
 * utterances' weighted unit_features [T_b, Dt] concatenated; lens[b] = T_b;
 * start_state[b] = -1 or a unit id (may be NULL = all -1).  Outputs: paths, the
 * T_b // multiepoch selected row ids per utterance concatenated; step_dist (optional,
 * may be NULL) the joint Euclidean distance of each step.
 * B = 1 -- the reference's own call, one utterance at a time (synth_simple.py:413) -- runs as ONE
 * persistent cooperative kernel for the shipped epoch-voice shapes (greedy_one.cu: every step
 * streams the operand rows once, grid barrier, float64 re-rank and next query inside the kernel);
 * the results are bit-identical to the batched path (SNK_GREEDY_NO_ONE=1 forces that one).  */
int snk_greedy_batch(snk_db *db, const double *targets, const int64_t *lens, int B,
                     const int64_t *start_state, int64_t *paths, double *step_dist);
/* device variant: d_targets as above in device memory; lens / start_state are HOST arrays
 * (they only size the launch); outputs in device memory.                               */
int snk_greedy_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B,
                         const int64_t *start_state, int64_t *d_paths, double *d_step_dist,
                         void *stream);
/* completes every snk_greedy_batch_dev / snk_greedy_batch_unnorm_dev enqueued since the last call */
int snk_greedy_batch_finish(snk_db *db);
/* Greedy search over a database whose joint rows are sharded by row block across the ranks of the communicator
 * (snk_comm_init; SURVEY.md section 8e, row "greedy chain with sharded DB" -- the reference has no counterpart, its chain
 * is script/synth_simple.py:487-501 in one process).  This handle holds joint rows [id_offset, id_offset + rows) of the
 * rows_full searchable rows, i.e. frames F[id_offset : id_offset + rows + m - 1] and join contexts
 * Jc[id_offset : id_offset + rows + m]; d_Jc_full is the replicated un-weighted float32 join matrix [rows_full + m, Dj] on
 * this device, from which every rank reads the previous join vector (current_join_rep[u] = Jw[u + m]).  After every time
 * step the ranks exchange their best (distance, global row) per utterance and the distance below which no row outside
 * their shortlist can lie -- one grouped ncclAllGather of B * 24 bytes plus an arg-min kernel, enqueued by the library --
 * so the paths (GLOBAL row ids) are identical on every rank, and an answer is certified when it is not above any rank's
 * bound (a shard that holds no close row cannot certify its own best, and does not need to).
 * B = 1 with mapped peers: the single-utterance persistent kernel scans the shard and performs that exchange itself
 * (24-byte stores into every peer's region + epoch flag, once per step) -- one launch per utterance on every rank.
 * Collective: every rank calls it with the same targets, then snk_greedy_batch_finish (which also agrees, with one
 * all-reduce, on the utterances whose certificates failed on any rank and repeats them in lockstep).               */
int snk_greedy_sharded_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B, const int64_t *start_state,
                                 const float *d_Jc_full, int64_t rows_full, int64_t id_offset, int64_t *d_paths,
                                 double *d_step_dist, void *stream);

/* ---- target preparation -----------------------------------------------------
 * Replaces the per-utterance host numpy between compose_speech and the search
 * (synth_simple.py:371-391): standardise (data_manipulation.py:162-186) then weight
 * (speech_manip.py:209-213), in the reference's float64 arithmetic on its float32 input:
 *   y[r,c] = (x[r,c] == special_uv_value ? std[c] * -1.0 * uv_scaling_factor
 *                                        : (x[r,c] - mean[c]) / std[c]) * target_weight[c]
 * snk_db_set_standardisation stores mean_vec_target / std_vec_target ([Dt]) and the two
 * constants of const.py:12-14 (-1000.0, 20.0).  The weights are those of snk_db_set_weights,
 * so a re-weighting needs no second call here.  flags: SNK_STD_FLOAT32 when the statistics
 * are float32 values (the voice file stores them as 'f', train_simple.py:94-97): numpy then
 * keeps subtraction, division and the unvoiced value in float32, and so does the kernel;
 * without the flag the arithmetic is float64 (float64 statistics promote the float32 speech). */
#define SNK_STD_FLOAT32 1u
int snk_db_set_standardisation(snk_db *db, const double *mean, const double *std,
                               double special_uv_value, double uv_scaling_factor,
                               unsigned flags);
/* unnorm float32 [rows, Dt] (compose_speech output) -> out float64 [rows, Dt], host buffers */
int snk_prepare_targets(snk_db *db, const float *unnorm, int64_t rows, double *out);
/* Half-phone target preparation: what get_halfphone_stats + standardise + weight produce
 * between compose_speech and the search (reference script/train_halfphone.py:959-1070 as called from
 * script/synth_halfphone.py:1510-1548).  unnorm: float32 [frames, dim] un-normalised speech of one utterance;
 * points: int64 [n, P], P = 1 / 2 / 3 frame indices per half-phone (onepoint: middle; twopoint: first, last; threepoint:
 * first, middle, last -- taken from the 5-state alignment, see snickery_b200/synth.py::halfphone_unit_points);
 * durations: float64 [n] normalised durations appended as the last column, or NULL.  P * dim (+ 1) must equal Dt.
 * out float64 [n, Dt] = weight(hstack(standardise(unnorm)[points], durations)): the statistics of
 * snk_db_set_standardisation are Dt wide, i.e. the frame statistics repeated per point (any value for the duration
 * column, which is only weighted).  Bit-identical to the numpy expression in both arithmetic modes.                  */
int snk_halfphone_targets(snk_db *db, const float *unnorm, int64_t frames, int dim, const int64_t *points, int64_t n, int P,
                          const double *durations, double *out);
int snk_halfphone_targets_dev(snk_db *db, const float *d_unnorm, int64_t frames, int dim, const int64_t *d_points, int64_t n,
                              int P, const double *d_durations, double *d_out, void *stream);
/* snk_greedy_batch on un-normalised float32 speech: the standardise+weight step is fused
 * into the query assembly, half the host->device bytes of the float64 form.  Results are
 * identical to snk_greedy_batch(snk_prepare_targets(unnorm)).                              */
int snk_greedy_batch_unnorm(snk_db *db, const float *unnorm, const int64_t *lens, int B,
                            const int64_t *start_state, int64_t *paths, double *step_dist);
int snk_greedy_batch_unnorm_dev(snk_db *db, const float *d_unnorm, const int64_t *lens, int B,
                                const int64_t *start_state, int64_t *d_paths,
                                double *d_step_dist, void *stream);

/* ---- candidate target distances --------------------------------------------------------
 * Replaces the distance half of preselect_units_quinphone (synth_halfphone.py:1346-1351):
 * dist[t,j] = || Fw[cand[t,j]] - targets[t] ||_2 ; cand == -1 indexes the last unit as
 * numpy does.  cand int64 [T,K], targets float64 [T,Dt], dist float64 [T,K].           */
int snk_candidate_distances(snk_db *db, const int64_t *cand, const double *targets, int64_t T, int K,
                            double *dist);

/* ---- join-cost tiles ------------------------------------------------------------------------
 * Replaces get_natural_distance_vectorised over the pair lists of
 * make_on_the_fly_join_lattice_BLOCK_DIRECT (synth_halfphone.py:2942-2951,3206-3301):
 * for each utterance b and step t < T_b-1 the K x K tile
 *   tile[a][c] = || end[cand[t,a]] - start[cand[t+1,c]] ||_2   (float32)
 * with +inf where either unit is inadmissible (-1, < 1, >= N-1: :3238-3268).
 * tiles: float32 [sum_b (T_b - 1), K, K].                                               */
int snk_join_tiles(snk_db *db, const int64_t *cand, const int64_t *lens, int B, int K, float *tiles);
/* statistics of the last snk_join_tiles call when the tensor-core path took it (n_candidates <= 64):
 * stats[0] = finite tile entries, stats[1] = entries whose rows nearly coincide and were therefore recomputed by direct
 * float64 differences instead of the norm expansion (csrc/join_tc.cu).                                               */
int snk_join_stats(const snk_db *db, int64_t stats[2]);

/* ---- join + Viterbi --------------------------------------------------------------------------
 * Replaces Synthesiser.viterbi_search (synth_halfphone.py:1399-1436) =
 * make_target_sausage_lattice + join lattice + compose + shortestpath
 * (fst_functions_wrapped.py:28-58,172-217,368,387-408) for B utterances.
 * cand int64 [sum T_b, K] (-1 padded), tdist float64 [sum T_b, K].
 * paths int64 [sum T_b]: unit ids; when an utterance has no admissible path
 * (T_b < 2, or every path blocked) path_len[b] = 0 as the reference returns [].
 * path_cost[b]: float32-accumulated cost in OpenFst's arc order, widened to double;
 * tcost / jcost (optional, may be NULL): float64 sums of the target and join parts.
 * flags: bit 0 = greedy over candidates (beam 1: keep only the best state per step).    */
int snk_join_viterbi_batch(snk_db *db, const int64_t *cand, const double *tdist, const int64_t *lens,
                           int B, int K, unsigned flags, int64_t *paths, int64_t *path_len,
                           double *path_cost, double *tcost, double *jcost);
int snk_join_viterbi_batch_dev(snk_db *db, const int64_t *d_cand, const double *d_tdist,
                               const int64_t *lens, int B, int K, unsigned flags, int64_t *d_paths,
                               int64_t *d_path_len, double *d_path_cost, double *d_tcost,
                               double *d_jcost, void *stream);

/* ---- acoustic preselection + join + Viterbi in one call ----------------------------------------
 * Replaces the pair preselect_units_acoustic -> viterbi_search of synth_utt (synth_halfphone.py:1359-1366,
 * 1399-1436, call sites :1611-1625) for B utterances: targets float64 [sum T_b, Dt] weighted unit
 * features; the k-NN candidate lists (K = n_candidates per target) stay on the device and feed the
 * join / Viterbi stage directly; only the paths come back.  Outputs as snk_join_viterbi_batch.    */
int snk_acoustic_viterbi_batch(snk_db *db, const double *targets, const int64_t *lens, int B, int K,
                               unsigned flags, int64_t *paths, int64_t *path_len, double *path_cost,
                               double *tcost, double *jcost);
int snk_acoustic_viterbi_batch_dev(snk_db *db, const double *d_targets, const int64_t *lens, int B, int K,
                                   unsigned flags, int64_t *d_paths, int64_t *d_path_len,
                                   double *d_path_cost, double *d_tcost, double *d_jcost, void *stream);
int snk_acoustic_viterbi_finish(snk_db *db);

/* ---- per-stream cost report ------------------------------------------------------------------
 * Replaces get_target_scores_per_stream / get_join_scores_per_stream
 * (synth_halfphone.py:1964-1981,2977-3008) for one greedy path: squared-error sums per
 * stream.  stream widths are given as arrays; tscores [P, n_tstreams], jscores
 * [P-1, n_jstreams] (float64).  targets: the utterance's weighted unit_features [T, Dt]. */
int snk_greedy_path_scores(snk_db *db, const double *targets, int64_t T, const int64_t *path,
                           int64_t P, const int *twidths, int n_tstreams, const int *jwidths,
                           int n_jstreams, double *tscores, double *jscores);

/* ---- row N2: MagPhase epoch concatenation -------------------------------------------------------
 * Replaces the array part of concatenateMagPhaseEpoch_sep_files + retrieve_magphase_frag
 * (synth_simple.py:677-747,538-652; matrix_operations.py:1-30).  The frame store holds the
 * per-sentence full-band MagPhase files concatenated: mag / real / imag float32 [nframes, width]
 * (width = FFTHALFLEN = 1025, const.py:20), the interpolated f0 and the voicing flag, float64 [nframes]
 * (speech_manip.lin_interp_f0, synth_simple.py:562).  Per unit: the global index of its first
 * frame (unit_index_within_sentence + sentence offset) and the frame bounds [lo, hi) of its
 * sentence (fragments are zero padded outside them).                                           */
typedef struct snk_frames snk_frames;
int snk_frames_create(snk_frames **out, int device_id, int64_t nframes, int width, const float *mag,
                      const float *real, const float *imag, const double *f0_interp, const double *vuv,
                      int64_t nunits, const int64_t *unit_frame, const int64_t *sent_lo, const int64_t *sent_hi);
int snk_frames_destroy(snk_frames *fr);
/* path: P selected units; outputs float64 [P * multiepoch, width] (mag, real, imag) and
 * [P * multiepoch] (fz, vuv) -- the arrays the reference hands to magphase.synthesis_from_lossless.
 * taper_in: the in-taper np.hanning(((overlap + 1) * 2) + 1)[1:overlap + 1] (float64, [overlap]);
 * has_fzero != 0 means the caller imposes the target f0 afterwards (no unvoiced zeroing).
 * kernel_ms (optional): device time of the gather kernel.                                      */
int snk_concat_magphase_epoch(snk_frames *fr, const int64_t *path, int64_t P, int multiepoch, int overlap,
                              const double *taper_in, int has_fzero, double *mag, double *real, double *imag,
                              double *fz, double *vuv, double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* SNK_B200_H */
