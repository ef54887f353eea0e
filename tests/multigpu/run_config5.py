"""Config 5 shape (BASELINE.json configs[4]) under torchrun: an epoch database sharded by joint-row
block over the ranks, k-NN in the 517-dim joint space, NCCL all-gather + GPU merge of per-shard top-k.
Checks the merged answer against float64 brute force for a few queries and reports queries/s.

    torchrun --nproc-per-node R tests/multigpu/run_config5.py [--rows-per-gpu 1250000] [--queries 4096]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from snickery_b200 import distributed as D, synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=1250000)
    ap.add_argument("--queries", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = 6
    n_total = args.rows_per_gpu * world + (m - 1)
    # every rank generates only its own block (+ halo) from a per-block seed; block b of the database is
    # make_epoch_db(seed=5000+b) so any rank can regenerate any block for checking
    def block(b):
        return syn.make_epoch_db(n_units=args.rows_per_gpu + (m if b == world - 1 else 0), seed=5000 + b)
    mine = block(rank)
    nxt = block(rank + 1) if rank + 1 < world else None
    F = mine["F"] if nxt is None else np.vstack([mine["F"], nxt["F"][: m - 1]])
    X = mine["Jc"][1:] if nxt is None else np.vstack([mine["Jc"][1:], nxt["Jc"][1: m + 1]])
    Jc = np.vstack([X[:1], X])[: F.shape[0] + 1]   # join context row u = join frame of unit u-1 (block-local history)
    wt = np.full(61, 0.4)
    wj = np.full(151, 0.05)
    sk = D.ShardedKnn.__new__(D.ShardedKnn)
    sk.rank, sk.world, sk.group, sk.device = rank, world, None, local
    sk.lo, sk.hi = rank * args.rows_per_gpu, (rank + 1) * args.rows_per_gpu
    from snickery_b200 import engine
    nloc = args.rows_per_gpu
    sk.db = engine.UnitDatabase(F[: nloc + m - 1], Jc[: nloc + m], multiepoch=m, device=local)
    sk.db.set_weights(wt, wj)
    sk.space = engine.SPACE_JOINT
    # queries: noisy copies of joint rows of block 0 (replicated: same seed everywhere)
    rng = np.random.default_rng(17)
    b0 = block(0)
    F0, J0 = b0["F"].astype(np.float64) * wt, b0["Jc"].astype(np.float64) * wj
    rows = rng.integers(0, args.rows_per_gpu - m, args.queries)
    Q = np.hstack([J0[rows]] + [F0[rows + j] for j in range(m)]) + 0.02 * rng.standard_normal((args.queries, 151 + 61 * m))
    qd = torch.from_numpy(Q).cuda()
    for k in (1, 50):
        d, i = sk.query(qd, k)          # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        d, i = sk.query(qd, k)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            # the true nearest row of a 0.02-noise copy is (almost surely) its source row in block 0
            hit = float((i[:, 0].cpu().numpy() == rows).mean())
            nC = args.rows_per_gpu - m
            C = np.hstack([J0[:nC]] + [F0[j: nC + j] for j in range(m)])
            for t in range(3):
                d2 = ((C - Q[t]) ** 2).sum(1)
                assert d[t, 0].item() <= np.sqrt(d2.min()) * (1 + 1e-9), "merged answer worse than block-0 optimum"
            c = sk.db.counters()
            print("config5 world=%d rows=%d D=517 k=%d: %.1f ms for %d queries -> %.0f queries/s, source-row hit %.3f, "
                  "recertified %d/%d" % (world, n_total, k, dt * 1e3, args.queries, args.queries / dt, hit,
                                         c["recertified"], c["queries"]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
