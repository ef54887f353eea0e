"""Multi-GPU check (run under torchrun on a GPU box, not collected by pytest):
database-sharded k-NN with NCCL all-gather + GPU merge must equal the single-GPU search."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from snickery_b200 import GpuKDTree, distributed as D, synthetic as syn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hp = syn.make_halfphone_db(n_units=120000, seed=1237)
    w = np.full(184, 0.4)
    w[-1] = 0.5
    rng = np.random.default_rng(3)
    q = hp["F"][rng.integers(0, 120000, 300)].astype(np.float64) * w + 0.05 * rng.standard_normal((300, 184))
    qd = torch.from_numpy(q).cuda()
    for k in (1, 50):
        sk = D.ShardedKnn(hp["F"], w, hp["F"].shape[0], rank, world, local)
        d, i = sk.query(qd, k)
        torch.cuda.synchronize()
        if rank == 0:
            full = GpuKDTree.from_weighted(hp["F"], w, device=local)
            rd, ri = full.query(q, k=k)
            rd, ri = np.asarray(rd).reshape(300, -1), np.asarray(ri).reshape(300, -1)
            assert np.array_equal(i.cpu().numpy(), ri), "sharded ids differ from single-GPU ids (k=%d)" % k
            assert np.allclose(d.cpu().numpy(), rd, rtol=1e-12)
            # spot check against float64 brute force
            Y = hp["F"].astype(np.float64) * w
            for t in range(5):
                d2 = ((Y - q[t]) ** 2).sum(1)
                assert ri[t, 0] == int(np.argmin(d2))
            print("sharded knn k=%d world=%d OK" % (k, world), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
