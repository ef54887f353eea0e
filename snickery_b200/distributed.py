"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

Two shardings exist on this path:

* greedy / join+Viterbi shard by UTTERANCE with the database replicated -- no data-path collective,
  only a final gather of the (tiny) integer paths to rank 0;
* k-NN over a database too large or too slow for one GPU shards the database ROWS: every rank
  searches its block with the replicated queries, the per-shard top-k (float64 distance, int64
  global row id) are all-gathered (nq*k*16 bytes per rank -- latency bound) and k-way merged on the
  GPU with the lowest-global-id tie rule.  On GPUs the whole exchange lives INSIDE the library
  (snk_comm_init + snk_knn_sharded_dev: ncclAllGather + merge kernel on the caller's stream);
  torch.distributed only carries the 128-byte NCCL id to the ranks.  CPU tensors (gloo tests of the
  host logic) take the torch all-gather + merge_topk_reference path.

The reference has no distributed code at all (only multiprocessing.Pool over utterances,
script/synth_halfphone.py:897-903), so there is no reference interface to mirror here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine


def shard_rows(n_rows, rank, world):
    """Contiguous block partition [lo, hi) of database rows; blocks differ by at most one row."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_utterances(n_utts, rank, world):
    """Round-robin utterance assignment (keeps long/short utterances balanced after a length sort)."""
    return list(range(rank, int(n_utts), int(world)))


def merge_topk_reference(dist_all, idx_all):
    """Pure-torch k-way merge of [R, nq, k] ascending lists (ties: lowest id).  Used for CPU tensors in
    the gloo tests and as the checker of the CUDA merge kernel; the GPU path uses snk_topk_merge_dev."""
    import torch
    R, nq, k = dist_all.shape
    d = dist_all.permute(1, 0, 2).reshape(nq, R * k)
    i = idx_all.permute(1, 0, 2).reshape(nq, R * k)
    order = torch.argsort(i, dim=1, stable=True)
    d, i = torch.gather(d, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(d, dim=1, stable=True)[:, :k]
    return torch.gather(d, 1, order), torch.gather(i, 1, order)


def allgather_merge_topk(dist_local, idx_local, group=None):
    """All-gather per-shard [nq, k] results and merge them; every rank returns the global top-k."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq, k = dist_local.shape
    d_all = torch.empty((world * nq, k), dtype=dist_local.dtype, device=dist_local.device)
    i_all = torch.empty((world * nq, k), dtype=idx_local.dtype, device=idx_local.device)
    dist.all_gather_into_tensor(d_all, dist_local.contiguous(), group=group)
    dist.all_gather_into_tensor(i_all, idx_local.contiguous(), group=group)
    d_all, i_all = d_all.view(world, nq, k), i_all.view(world, nq, k)
    if not dist_local.is_cuda:
        return merge_topk_reference(d_all, i_all)
    out_d = torch.empty_like(dist_local)
    out_i = torch.empty_like(idx_local)
    lib = engine.load_library()
    stream = torch.cuda.current_stream(dist_local.device)
    rc = lib.snk_topk_merge_dev(dist_local.device.index, C.c_void_p(d_all.data_ptr()), C.c_void_p(i_all.data_ptr()),
                                world, nq, k, C.c_void_p(out_d.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                C.c_void_p(stream.cuda_stream))
    if rc:
        raise engine.EngineError(lib.snk_last_error().decode())
    return out_d, out_i


def init_comm(db, group=None):
    """Attach an NCCL communicator to a UnitDatabase: rank 0 draws the id, torch.distributed broadcasts it."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    db.comm_init(box[0], rank, world)


class ShardedKnn:
    """Database rows block-partitioned over the ranks of a process group; `.query` returns the same
    (dist, idx) on every rank as a single-GPU search of the whole matrix would."""

    def __init__(self, F_full_or_shard, weights, n_rows_total, rank, world, device, is_shard=False, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = shard_rows(n_rows_total, rank, world)
        F = F_full_or_shard if is_shard else F_full_or_shard[self.lo:self.hi]
        F = np.ascontiguousarray(F, dtype=np.float32)
        self.db = engine.UnitDatabase(F, np.zeros((F.shape[0] + 1, 1), np.float32), multiepoch=1, device=device)
        self.db.set_weights(np.asarray(weights, dtype=np.float64), np.ones(1))
        self.device = device
        if world > 1:
            init_comm(self.db, group)

    @classmethod
    def from_epoch_db(cls, F, Jc, multiepoch, wt, wj, rank, world, device, group=None):
        """Joint-space (greedy) search rows [prev_join || m-frame window] sharded by row block: rank r holds
        joint rows [lo, hi) and therefore frames [lo, hi+m-1) and join contexts [lo, hi+m]."""
        self = cls.__new__(cls)
        self.rank, self.world, self.group, self.device = rank, world, group, device
        n_joint = F.shape[0] - (multiepoch - 1)
        self.lo, self.hi = shard_rows(n_joint, rank, world)
        Fs = np.ascontiguousarray(F[self.lo:self.hi + multiepoch - 1], dtype=np.float32)
        Js = np.ascontiguousarray(Jc[self.lo:self.hi + multiepoch], dtype=np.float32)
        self.db = engine.UnitDatabase(Fs, Js, multiepoch=multiepoch, device=device)
        self.db.set_weights(np.asarray(wt, dtype=np.float64), np.asarray(wj, dtype=np.float64))
        self.space = engine.SPACE_JOINT
        if world > 1:
            init_comm(self.db, group)
        return self

    def query_local_dev(self, q_dev, k):
        """q_dev: torch float64 CUDA tensor [nq, D]; returns this shard's certified top-k with GLOBAL row ids."""
        import torch
        nq = q_dev.shape[0]
        d = torch.empty((nq, k), dtype=torch.float64, device=q_dev.device)
        i = torch.empty((nq, k), dtype=torch.int64, device=q_dev.device)
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        self.db.knn_dev(q_dev.data_ptr(), nq, k, d.data_ptr(), i.data_ptr(), getattr(self, "space", engine.SPACE_TARGET),
                        self.lo, stream)
        self.db.knn_finish()
        return d, i

    def query(self, q_dev, k):
        """Global top-k on every rank: local search + ncclAllGather + merge, all enqueued by the library."""
        import torch
        if self.world == 1:
            return self.query_local_dev(q_dev, k)
        nq = q_dev.shape[0]
        d = torch.empty((nq, k), dtype=torch.float64, device=q_dev.device)
        i = torch.empty((nq, k), dtype=torch.int64, device=q_dev.device)
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        self.db.knn_sharded_dev(q_dev.data_ptr(), nq, k, d.data_ptr(), i.data_ptr(), getattr(self, "space", engine.SPACE_TARGET),
                                self.lo, stream)
        self.db.knn_sharded_finish()
        return d, i


class ShardedGreedy:
    """Greedy joint search over a database whose joint rows are sharded by row block (SURVEY.md section 8e,
    row "greedy chain with sharded DB"): one exchange PER TIME STEP, enqueued by the library
    (snk_greedy_sharded_batch_dev: grouped ncclAllGather of the per-shard best (distance, global row, bound) triples, B * 24
    bytes per rank, and an arg-min kernel).  The next step's previous-join vector is read from the replicated join
    contexts (current_join_rep[u] = Jw[u + m], reference script/synth_simple.py:213-214,501), so no second exchange
    is needed.  Latency bound: the step is the shard's search plus one small collective.  A batch of ONE utterance runs as
    one persistent kernel per rank (csrc/greedy_one.cu): shard scan, float64 re-rank, the exchange (24-byte stores into every
    peer's IPC-mapped region + an epoch flag) and the next query inside a single launch -- 30 us per step on 8 GPUs against
    89 us through the batched path."""

    def __init__(self, F, Jc, multiepoch, wt, wj, rank, world, device, group=None):
        import torch
        self.m = int(multiepoch)
        self.knn = ShardedKnn.from_epoch_db(F, Jc, multiepoch, wt, wj, rank, world, device, group)
        if world == 1:                             # a communicator of one: the same entry point on a single GPU
            self.knn.db.comm_init(engine.comm_unique_id(), 0, 1)
        self.rows_full = F.shape[0] - (self.m - 1)
        # replicated un-weighted float32 join contexts, as the voice file holds them
        self.Jc_full = torch.from_numpy(np.ascontiguousarray(Jc, dtype=np.float32)).to(torch.device("cuda", device))
        self.Dt = F.shape[1]

    def search(self, targets, start_states=None, return_dists=False):
        """targets: torch float64 CUDA tensor [B, T, Dt] (weighted, equal lengths); returns int64 [B, T // m] of GLOBAL
        row ids, identical on every rank."""
        import torch
        B, T, Dt = targets.shape
        steps = T // self.m
        if steps == 0:
            raise ValueError("Not enough data points to segment array in 'cut' mode")
        targets = targets.contiguous()
        paths = torch.empty((B, steps), dtype=torch.int64, device=targets.device)
        dists = torch.empty((B, steps), dtype=torch.float64, device=targets.device)
        stream = torch.cuda.current_stream(targets.device).cuda_stream
        db = self.knn.db
        db.greedy_sharded_batch_dev(targets.data_ptr(), np.full(B, T, dtype=np.int64), self.Jc_full.data_ptr(), self.rows_full,
                                    self.knn.lo, paths.data_ptr(), dists.data_ptr(), start_states, stream)
        db.greedy_batch_finish()
        return (paths, dists) if return_dists else paths


def gather_paths(paths_local, utt_ids_local, n_utts, group=None):
    """Collect utterance-sharded path lists on every rank in utterance order (object all-gather: the
    payload is a few KB of integers)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    box = [None] * world
    dist.all_gather_object(box, (list(utt_ids_local), list(paths_local)), group=group)
    out = [None] * n_utts
    for ids, paths in box:
        for u, p in zip(ids, paths):
            out[u] = p
    return out
