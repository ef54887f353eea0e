"""bench.py's sharded_greedy block alone (torchrun): step latency of the database-sharded greedy search at B = 1024 and
B = 1 on configs[1]'s database, paths compared with the replicated single-GPU search."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from snickery_b200 import Synthesiser  # noqa: E402

D = bench.Dist()
cfg = bench.workload_config()
db = bench.make_database(bench.DB_UNITS)
probe = Synthesiser(cfg, db["F"][:2000], db["Jc"][:2001], device=D.local)      # only for the weight vectors
wt, wj = probe.target_weight_vector, probe.join_weight_vector
probe.db.close()
res = bench.block_sharded_greedy(D, db, wt, wj, cfg)
if D.rank == 0:
    print(json.dumps(res), flush=True)
D.close()
