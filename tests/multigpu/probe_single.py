"""Single-utterance greedy search at configs[1] size: the persistent kernel (greedy_one.cu) against the batched path
(SNK_GREEDY_NO_ONE=1), same utterance, paths compared.  python tests/multigpu/probe_single.py [units]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from snickery_b200 import Synthesiser  # noqa: E402

units = int(sys.argv[1]) if len(sys.argv) > 1 else bench.DB_UNITS
cfg = bench.workload_config()
db = bench.make_database(units)
syn = Synthesiser(cfg, db["F"], db["Jc"])
wt = syn.target_weight_vector
dev = torch.device("cuda:0")
steps = bench.UTT_FRAMES // bench.MULTIEPOCH
res = {}
if os.environ.get("PROBE_NCU"):
    # two launches of the persistent kernel for an ncu capture, nothing else
    uf = bench.make_batch(db["F"], wt, 1, bench.UTT_FRAMES, seed=4711)
    d_t = torch.from_numpy(uf).to(dev)
    d_p = torch.empty(steps, dtype=torch.int64, device=dev)
    lens = np.array([bench.UTT_FRAMES], dtype=np.int64)
    for _ in range(2):
        syn.db.greedy_batch_dev(d_t.data_ptr(), lens, d_p.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    syn.db.greedy_batch_finish()
    sys.exit(0)
for seed in (4711, 12):
    uf = bench.make_batch(db["F"], wt, 1, bench.UTT_FRAMES, seed=seed)
    d_t = torch.from_numpy(uf).to(dev)
    d_p = torch.empty(steps, dtype=torch.int64, device=dev)
    lens = np.array([bench.UTT_FRAMES], dtype=np.int64)
    stream = torch.cuda.current_stream()
    for mode in ("one", "batched"):
        if mode == "batched":
            os.environ["SNK_GREEDY_NO_ONE"] = "1"
        else:
            os.environ.pop("SNK_GREEDY_NO_ONE", None)
        for _ in range(2):
            syn.db.greedy_batch_dev(d_t.data_ptr(), lens, d_p.data_ptr(), stream=stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            syn.db.greedy_batch_dev(d_t.data_ptr(), lens, d_p.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        syn.db.greedy_batch_finish()
        res[(seed, mode)] = (e0.elapsed_time(e1) / 5 * 1e3 / steps, d_p.cpu().numpy().copy())
        print(seed, mode, "us/step %.2f" % res[(seed, mode)][0], syn.db.counters(reset=True), flush=True)
    same = np.array_equal(res[(seed, "one")][1], res[(seed, "batched")][1])
    print("paths equal:", same, flush=True)
    assert same
os.environ.pop("SNK_GREEDY_NO_ONE", None)
# phase split of the persistent kernel (CTA 0's clock)
import ctypes as C
from snickery_b200.engine import load_library
os.environ["SNK_G1_TIMING"] = "1"
lib = load_library()
lib.snk_debug_greedy_one_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for mb in [0.0]:
    for _ in range(2):
        syn.db.greedy_batch_dev(d_t.data_ptr(), lens, d_p.data_ptr(), stream=stream.cuda_stream)
    syn.db.greedy_batch_finish()
    tt = np.zeros((steps, 16), dtype=np.uint64)
    assert lib.snk_debug_greedy_one_times(syn.db._h, tt.ctypes.data, steps) == 0
    tt = tt.astype(np.int64)[2:]
    med = lambda a, b: np.median(tt[:, a] - tt[:, b]) / 1e3
    print("phases (%d) us: scan %.1f | warps %.1f cta-merge %.1f publish %.1f fence %.1f | wait %.1f | lists %.1f select %.1f rerank %.1f "
          "cert %.1f query %.1f | step %.1f" % (
              mb, np.median(tt[1:, 0] - tt[:-1, 5]) / 1e3, med(1, 0), med(9, 1), med(10, 9), med(2, 10), med(3, 2), med(6, 3),
              med(7, 6), med(8, 7), med(4, 8), med(5, 4), np.median(tt[1:, 5] - tt[:-1, 5]) / 1e3), flush=True)
    assert np.array_equal(d_p.cpu().numpy(), res[(seed, "one")][1])
# per-CTA scan end times: is the skew the barrier waits for tied to the SM (systematic) or random?
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
ct = np.zeros((min(steps, 256), n_sm), dtype=np.uint64)
lib.snk_debug_greedy_one_cta_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
if lib.snk_debug_greedy_one_cta_times(syn.db._h, ct.ctypes.data, min(steps, 256)) == 0:
    ct = ct.astype(np.int64)[2:]
    rel = (ct - np.median(ct, axis=1, keepdims=True)) / 1e3            # us after the median CTA, per step
    per_cta = rel.mean(axis=0)
    print("scan-end skew: last CTA %.1f us after the median (mean over steps); per-CTA mean offset std %.2f us, "
          "per-step residual std %.2f us; slowest CTAs %s" % (
              (rel.max(axis=1)).mean(), per_cta.std(), (rel - per_cta[None, :]).std(),
              np.argsort(per_cta)[-6:].tolist()), flush=True)
    print("per-CTA mean offsets (us):", np.round(per_cta, 1).tolist(), flush=True)
del os.environ["SNK_G1_TIMING"]
t0 = time.perf_counter()
syn.greedy_joint_search(uf)
print("host call ms %.2f" % ((time.perf_counter() - t0) * 1e3))
