"""Multi-GPU check (torchrun): greedy search over a row-sharded database with a per-step NCCL exchange
must select exactly the path the single-GPU engine selects."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import epoch_config  # noqa: E402
from snickery_b200 import Synthesiser, distributed as D, synthetic as syn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    db = syn.make_epoch_db(n_units=150000, seed=321)
    cfg = epoch_config(tsw=(0.5, 0.5))
    ref = Synthesiser(cfg, db["F"], db["Jc"], device=local)     # replicated database, single-GPU search
    wt, wj = ref.target_weight_vector, ref.join_weight_vector
    B, T = 16, 72
    utts = [x.astype(np.float64) * wt for x in syn.make_targets(db["F"], B, T, seed=9)]
    sg = D.ShardedGreedy(db["F"], db["Jc"], 6, wt, wj, rank, world, local)
    tg = torch.from_numpy(np.stack(utts)).cuda()
    starts = [-1] * (B - 1) + [777]
    paths = sg.search(tg, starts).cpu().numpy()
    want = ref.greedy_joint_search_batch(utts, starts)
    ok = all(paths[b].tolist() == want[b] for b in range(B))
    # the same with the fp16 certificate forced to fail for every 3rd query: all ranks must agree on the utterances to
    # repeat (one all-reduce in snk_greedy_batch_finish) and repeat them in lockstep with the fp32 engine
    os.environ["SNK_DEBUG_CERT_FAIL"] = "3"
    sg2 = D.ShardedGreedy(db["F"], db["Jc"], 6, wt, wj, rank, world, local)
    del os.environ["SNK_DEBUG_CERT_FAIL"]
    paths2 = sg2.search(tg, starts).cpu().numpy()
    ok = ok and all(paths2[b].tolist() == want[b] for b in range(B)) and sg2.knn.db.counters()["recertified"] > 0
    # one utterance at a time: the persistent single-utterance kernel with the exchange through peer memory inside it
    # (greedy_one.cu) -- paths AND distances must equal the replicated search's, with and without a start state, and with
    # the certificate forced to fail (repair through the batched fp32 engine, all ranks in lockstep)
    ok1 = True
    for b in (0, 5, B - 1):
        sg.knn.db.counters(reset=True)
        p1, d1 = sg.search(tg[b:b + 1], [starts[b]], return_dists=True)
        launches = sg.knn.db.counters()["launches"]
        w1, wd1 = ref.greedy_joint_search_batch([utts[b]], [starts[b]], return_dists=True)
        same = p1[0].cpu().numpy().tolist() == w1[0] and np.array_equal(d1[0].cpu().numpy(), np.asarray(wd1[0]))
        used_one = launches <= 3 if (world > 1 and sg.knn.db.comm_info()["peer_exchange"]) else True
        ok1 = ok1 and same and used_one
        if rank == 0:
            print("single utterance %d: same=%s launches=%d" % (b, same, launches), flush=True)
    p2 = sg2.search(tg[2:3], [starts[2]]).cpu().numpy()
    ok1 = ok1 and p2[0].tolist() == want[2]
    ok = ok and ok1
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        assert flag.item() == 1, "sharded greedy paths differ from the single-GPU paths"
        print("sharded greedy world=%d B=%d steps=%d OK (recertified %d)" % (world, B, T // 6, sg.knn.db.counters()["recertified"]),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
