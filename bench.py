#!/usr/bin/env python
"""bench.py -- target frames/sec through the unit-selection search hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement)
    python bench.py --workload halfphone [--impl reference]  # configs[2] as the headline instead (same contract)

Headline workload (BASELINE.json configs[1], IS2018_nick_simplified.cfg): a full SLT-Arctic-sized epoch
database (700k units, 61-dim target / 151-dim join streams, multiepoch 6 => 517-dim joint rows),
exact greedy joint search (search_epsilon = 0).  One "step" = one pass of the hot path over one
batch of B synthetic target utterances of 648 frames (108 greedy steps each) per GPU.  With N
GPUs the database is replicated and utterances are sharded (weak scaling, no data-path
collective).  Data are synthetic magphase-shaped features
(snickery_b200/synthetic.py); there are no published reference numbers.

The same JSON line carries, for every N, the other configs of BASELINE.json as blocks:
  halfphone        configs[2]: acoustic k-NN (k = 50) -> join costs -> Viterbi, 90k half-phones (roofline of the k = 50
                   search, host-buffer e2e, parity sample against the oracle);
  viterbi_sharded  configs[3]: 1024 utterances x 80 targets x 50 candidates PER GPU, sharded by utterance;
  sharded_knn      configs[4]: 1.25 M joint rows (517-dim) PER GPU, queries replicated, NCCL all-gather + merge of the
                   per-shard top-k over NVLink inside the library (10 M rows at N = 8);
and at N = 1 `parity` (the GPU path against the oracle on the cpu_baseline utterance), `single_utterance`
(the literal drop-in call, one utterance at a time) and `cpu_baseline`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DB_UNITS = 700_000
UTT_FRAMES = 648
MULTIEPOCH = 6
METRIC = "target_frames_per_sec"
UNIT = "frames/s"
HP_UNITS, HP_UTTS, HP_T, HP_K = 90_000, 1024, 80, 50
SHARD_ROWS = 1_250_000


def workload_config():
    from conftest import epoch_config
    # config/IS2018_nick_simplified.cfg:77-78,97,103: equal stream weights, jcw 0.2, multiepoch 6
    return epoch_config(multiepoch=MULTIEPOCH, jcw=0.2, tsw=(0.5, 0.5), jsw=(0.25, 0.25, 0.25, 0.25))


def make_database(units):
    from snickery_b200 import synthetic as syn
    return syn.make_epoch_db(n_units=units, seed=1234 + 2)


def make_batch(F, wt, n_utts, frames, seed):
    """Weighted float64 target utterances concatenated [n_utts * frames, Dt] (what the reference
    hands to greedy_joint_search, synth_simple.py:381-390,413)."""
    rng = np.random.default_rng(seed)
    n = F.shape[0]
    starts = rng.integers(0, n - frames, size=n_utts)
    idx = (starts[:, None] + np.arange(frames)[None, :]).reshape(-1)
    seg = F[idx].astype(np.float64)
    uv = seg == -20.0
    seg += 0.3 * rng.standard_normal(seg.shape)
    seg[uv] = -20.0
    seg = seg.astype(np.float32).astype(np.float64)     # targets come from float32 files
    return seg * wt[None, :]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel):
    """dram bytes per launch of a kernel from the ncu capture committed under profiles/ (None if absent)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        return json.load(open(tpath)).get(kernel)
    return None


# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(db, cfg, wt, cores, budget_s, seed):
    """The reference's CPU path: scipy cKDTree(leafsize=100, balanced_tree=False) over the joint
    rows + the sequential greedy loop (synth_simple.py:229,458-503), eps = 0.  Utterance-parallel
    over `cores` forked workers like the reference's Pool (synth_halfphone.py:897-903).
    Returns a timed callable (its .paths / .dists hold the last result), the frames it covers and a description."""
    from oracle import snickery_oracle as O
    t0 = time.time()
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    build_s = time.time() - t0
    # calibrate: a handful of steps of one utterance
    probe = make_batch(db["F"], wt, 1, MULTIEPOCH * 4, seed)
    t0 = time.time()
    o.greedy_joint_search(probe)
    per_step = (time.time() - t0) / 4
    steps_per_utt = int(max(4, min(UTT_FRAMES // MULTIEPOCH, budget_s / max(per_step, 1e-4))))
    frames = steps_per_utt * MULTIEPOCH
    utts = make_batch(db["F"], wt, cores, frames, seed + 1).reshape(cores, frames, -1)
    global _CPU_ORACLE
    _CPU_ORACLE = o

    def run():
        t0 = time.time()
        if cores == 1:
            run.paths, run.dists = zip(*[o.greedy_joint_search(utts[0], return_dists=True)])
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_cpu_worker, [utts[i] for i in range(cores)])
        return time.time() - t0

    run.utts = utts
    run.oracle = o
    return run, cores * frames, {"tree_build_s": round(build_s, 2), "steps_per_utt": steps_per_utt,
                                 "utts": cores, "probe_s_per_query": round(per_step, 4)}


_CPU_ORACLE = None


def _cpu_worker(utt):
    return _CPU_ORACLE.greedy_joint_search(utt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload == "halfphone":
        return run_reference_halfphone(args)
    cfg = workload_config()
    db = make_database(args.db_units)
    from oracle import snickery_oracle as O
    wt = O.per_coeff_weights(np.array(cfg["target_stream_weights"]) * (1 - cfg["join_cost_weight"]),
                             cfg["stream_list_target"], cfg["datadims_target"])
    cores = os.cpu_count() or 1
    run, frames, info = cpu_reference_rate(db, cfg, wt, cores, budget_s=4.0, seed=99)
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    total = sum(times)
    value = frames * args.steps / total
    sample = "%d utts x %d greedy steps (of %d) per step, one forked worker per core; tree build %.1fs excluded" % (
        info["utts"], info["steps_per_utt"], UTT_FRAMES // MULTIEPOCH, info["tree_build_s"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_description(args, args.utts),   # the same workload; cpu_baseline.sample says how much of it was timed
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Python 2 + pywrapfst and cannot run here; this is oracle/ (its Python 3 restatement, "
                "pinned against the reference's own code by tests/test_reference_exec.py, driving the reference's own "
                "scipy cKDTree engine) on the host cores",
    }
    emit_line(line)
    return 0


def workload_description(args, utts):
    return {"workload": "IS2018_nick_simplified greedy joint search (configs[1])", "db_units": args.db_units,
            "target_dim": 61, "join_dim": 151, "multiepoch": MULTIEPOCH, "joint_dim": 151 + 61 * MULTIEPOCH,
            "utts_per_gpu": utts, "frames_per_utt": UTT_FRAMES, "search_epsilon": 0.0,
            "parallelism": "utterance-sharded, database replicated",
            "l2": "database operands (S16+G16, 0.36 GB) exceed the 126 MB L2 and are re-streamed every greedy step"}


# ---------------------------------------------------------------------------------------------- configs[2]: halfphone
def halfphone_description():
    return {"workload": "hybrid_halfphone_default acoustic preselection + join + Viterbi (configs[2])", "db_units": HP_UNITS,
            "target_dim": 184, "join_dim": 151, "n_candidates": HP_K, "utts_per_gpu": HP_UTTS, "targets_per_utt": HP_T,
            "search_epsilon": 0.0, "parallelism": "utterance-sharded, database replicated",
            "l2": "k-NN operand (G16, 35 MB) is L2 resident; every step writes 65 MB of candidates and gathers 5 GB of join rows"}


def _hp_worker(uf):
    o = _CPU_ORACLE
    from oracle import snickery_oracle as O
    cand, dist = o.preselect_units_acoustic(uf)
    return O.viterbi_search_numpy(o, cand, dist)


def run_reference_halfphone(args):
    """configs[2] on the host cores: cKDTree(leafsize=100, compact_nodes=False, balanced_tree=False).query(k=50)
    (synth_halfphone.py:379,1364) + numpy join costs (:2942-2951) + min-plus DP, one forked worker per utterance."""
    from conftest import halfphone_config
    from oracle import snickery_oracle as O
    from snickery_b200 import synthetic as syn
    global _CPU_ORACLE
    hp = syn.make_halfphone_db(n_units=args.hp_units, seed=1237)
    cfg = halfphone_config(n_candidates=HP_K, preselection="acoustic")
    t0 = time.time()
    o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
    o.build_acoustic_tree()
    build_s = time.time() - t0
    _CPU_ORACLE = o
    cores = os.cpu_count() or 1
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(hp["F"], cores, HP_T, seed=3)]

    def run():
        import multiprocessing as mp
        t0 = time.time()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_hp_worker, utts)
        return time.time() - t0

    for _ in range(min(args.warmup, 1)):
        run()
    steps = max(1, min(args.steps, 3))
    total = sum(run() for _ in range(steps))
    value = cores * HP_T * steps / total
    sample = "%d utts x %d targets per step (k = %d), one forked worker per core; tree build %.1fs excluded" % (
        cores, HP_T, HP_K, build_s)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": halfphone_description(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "oracle/ restatement (pinned by tests/test_reference_exec.py) driving the reference's own scipy cKDTree; "
                    "OpenFst replaced by the min-plus DP"}
    emit_line(line)
    return 0


class Dist:
    """rank / world / collectives that degrade to no-ops at world 1."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, reps, warm=1):
        """CUDA-event time of `reps` calls of fn on the current stream, barrier + sync on both sides, max over ranks (ms/call)."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.current_stream()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        self.barrier()
        return self.max(e0.elapsed_time(e1)) / reps

    def timed_median(self, fn, reps, warm=1):
        """Like timed(), one event pair per call, median over the calls: for calls that end in a host-side wait (the
        *_finish step), where a single slow host iteration would otherwise be averaged into the device figure."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        stream = torch.cuda.current_stream()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in ev:
            e0.record(stream)
            fn()
            e1.record(stream)
        self.barrier()
        return self.max(float(np.median([e0.elapsed_time(e1) for e0, e1 in ev])))

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from snickery_b200 import Synthesiser, engine

    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    if args.workload == "halfphone":
        out = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "data": "synthetic", "config": halfphone_description(), "steps": args.steps, "warmup": args.warmup}
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        hp = block_halfphone(D, args, headline=True)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            out.update(hp)
            out["clocks"] = clocks
            emit_line(out)
        D.close()
        return 0

    cfg = workload_config()
    db = make_database(args.db_units)
    syn = Synthesiser(cfg, db["F"], db["Jc"], device=local)
    syn.get_tree_for_greedy_search()
    if args.engine:
        syn.db.set_engine({"auto": 0, "simt": 1, "tc": 2}[args.engine])
    wt = syn.target_weight_vector
    B, T = args.utts, UTT_FRAMES
    lens = np.full(B, T, dtype=np.int64)
    steps_per_utt = T // MULTIEPOCH

    # per-step batches: distinct inputs per step, resident in HBM for `value`, pinned on the host for `e2e`
    nbatch = min(args.steps + args.warmup, 4)
    host_batches, dev_batches = [], []
    for i in range(nbatch):
        cat = make_batch(db["F"], wt, B, T, seed=1000 + 17 * rank + i)
        pinned = torch.empty(cat.shape, dtype=torch.float64).pin_memory()
        pinned.numpy()[...] = cat
        host_batches.append(pinned)
        dev_batches.append(pinned.to(dev))
    d_paths = torch.empty(B * steps_per_utt, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step_dev(i):
        # enqueue only (no synchronisation inside), then complete: certificates are read once per step
        syn.db.greedy_batch_dev(dev_batches[i % nbatch].data_ptr(), lens, d_paths.data_ptr(), stream=stream.cuda_stream)
        syn.db.greedy_batch_finish()

    for i in range(args.warmup):
        step_dev(i)
    D.barrier()
    c0 = syn.db.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # every launch of the distance GEMM inside the timed region is bracketed by CUDA events on its own stream
    # (snk_db_profile_*): roofline.achieved is the live average over exactly the launches that make up `value`
    syn.db.profile_enable(True)
    D.barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    e1.record(stream)
    D.barrier()
    ms = e0.elapsed_time(e1)
    prof = syn.db.profile_read(engine.PROF_KNN)
    syn.db.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    c1 = syn.db.counters()
    ms_max = D.max(ms)
    frames_per_step_all = world * B * T
    value = frames_per_step_all * args.steps / (ms_max / 1000.0)

    # ---- e2e: host (pinned) arrays in, host lists out, through the reference-facing Python API
    e2e_steps = max(1, min(args.steps, 3))

    def step_e2e(i):
        # unit ids come back as one int64 array per utterance (as_arrays): the call a batch user makes; the per-utterance
        # Python lists of the reference's single-utterance API are timed in `single_utterance.host_call_ms`
        return syn.db.greedy_batch_cat(host_batches[i % nbatch].numpy(), lens, as_arrays=True)

    step_e2e(0)
    D.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(i + 1)
    torch.cuda.synchronize()
    e2e_value = frames_per_step_all * e2e_steps / D.max(time.perf_counter() - t0)
    h2d = int(B * T * 61 * 8)
    d2h = int(B * steps_per_utt * 8)

    out = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        tf = prof["work"] / (prof["ms"] / 1e3) / 1e12 if prof["ms"] > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 tensor-core shortlist (fp32 accumulate) + f64 re-rank", "data": "synthetic",
            "config": workload_description(args, B),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(c1["launches"] - c0["launches"]),
            "roofline": {"bound": "tensor", "kernel": "knn_tc_kernel<0,4,26> (tcgen05 distance GEMM, query operand in tensor memory, fused per-query top-k lists)",
                         "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak if peak else None,
                         "traffic": ncu_traffic("knn_tc_kernel_bytes_per_launch"),
                         "traffic_source": "ncu --set full capture of this kernel committed under profiles/ (dram__bytes_read+write per launch)",
                         "peak_kind": peak_kind + " (bf16 sustained; fp16 runs at the same rate)",
                         "launches": prof["launches"], "avg_launch_ms": prof["ms"] / max(prof["launches"], 1),
                         "flops_per_launch": prof["work"] / max(prof["launches"], 1),
                         "kernel_share_of_step": prof["ms"] / ms_max},
            "exactness": {"queries": int(c1["queries"] - c0["queries"]),
                          "recertified_by_simt": int(c1["recertified"] - c0["recertified"]),
                          "exhaustive_f64": int(c1["exhaustive"] - c0["exhaustive"])},
            "rates": {"search_steps_per_s": value / MULTIEPOCH, "utterances_per_s": value / T,
                      "frames_per_utterance": int(T), "multiepoch": MULTIEPOCH},
        }

    # ---- same call from un-normalised float32 speech (row N4), N = 1 only
    if world == 1 and not args.no_secondary:
        out["greedy_e2e_from_unnormalised_f32"] = block_unnorm(syn, db, wt, host_batches[0].numpy(), lens)
        out["single_utterance"] = block_single_utterance(D, syn, db, wt)
    del dev_batches, host_batches

    # ---- configs[2] / [3] / [4] as blocks, every N
    if not args.no_secondary:
        hp = block_halfphone(D, args, headline=False)
        sk = block_sharded_knn(D, args)
        sg = block_sharded_greedy(D, db, wt, syn.join_weight_vector, cfg)
        if rank == 0:
            out["halfphone"] = hp["halfphone"]
            out["viterbi_sharded"] = hp["viterbi_sharded"]
            out["sharded_knn"] = sk
            out["sharded_greedy"] = sg
    # ---- CPU baseline + parity of the GPU path against it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        run, frames, info = cpu_reference_rate(db, cfg, wt, 1, budget_s=12.0, seed=99)
        t = run()
        out["cpu_baseline"] = {"value": frames / t, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "1 utt x %d greedy steps, scipy cKDTree eps=0 single thread as the reference "
                                         "runs (synth_simple.py:279-284); tree build %.1fs excluded" %
                                         (info["steps_per_utt"], info["tree_build_s"])}
        out["parity"] = greedy_parity(syn, run)
    if rank == 0:
        emit_line(out)
    D.close()
    return 0


def greedy_parity(syn, run):
    """The GPU path on the cpu_baseline utterance against the oracle's (cKDTree) path for the same input."""
    uf = run.utts[0]
    ref_path, ref_d = list(run.paths[0]), np.asarray(run.dists[0])
    paths, dists = syn.greedy_joint_search_batch([uf], return_dists=True)
    p, d = paths[0], np.asarray(dists[0])
    n = len(ref_path)
    first = next((t for t in range(n) if p[t] != ref_path[t]), None)
    same = n if first is None else first          # after a divergence the chains see different inputs
    rel = np.abs(d[:same] - ref_d[:same]) / np.maximum(ref_d[:same], 1e-300)
    outside = 0
    if first is not None and abs(d[first] - ref_d[first]) > 1e-6 * ref_d[first]:
        outside = 1
    return {"what": "greedy path of the cpu_baseline utterance: CUDA engine vs oracle (scipy cKDTree, eps = 0)",
            "steps": n, "mismatches": 0 if first is None else 1, "mismatches_outside_1e-6_tie": outside,
            "first_divergence_step": first, "max_rel_cost_err": float(rel.max()) if same else None}


def block_unnorm(syn, db, wt, weighted, lens):
    rng = np.random.default_rng(5)
    import torch
    Dt = db["F"].shape[1]
    mean, std = rng.normal(size=Dt), rng.uniform(0.5, 2.0, size=Dt)
    x = weighted / np.where(wt != 0, wt, 1.0)
    raw = torch.empty(x.shape, dtype=torch.float32).pin_memory()
    raw.numpy()[...] = x * std + mean
    del x
    syn.set_standardisation(mean, std)
    syn.db.greedy_batch_cat(raw.numpy(), lens, unnorm=True, as_arrays=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        syn.db.greedy_batch_cat(raw.numpy(), lens, unnorm=True, as_arrays=True)
    torch.cuda.synchronize()
    return {"value": lens.sum() * 2 / (time.perf_counter() - t0), "unit": UNIT, "steps": 2,
            "h2d_bytes_per_step": int(lens.sum() * Dt * 4),
            "note": "snk_greedy_batch_unnorm: compose_speech-style float32 in, host lists out"}


def block_single_utterance(D, syn, db, wt):
    """The literal drop-in call: greedy_joint_search(one utterance) (synth_simple.py:413), 108 dependent steps.  One
    persistent cooperative kernel (greedy_one.cu); the batched path (three launches per step) is timed beside it."""
    import torch
    peaks, kind = measured_peaks()
    uf = make_batch(db["F"], wt, 1, UTT_FRAMES, seed=4711)
    d_t = torch.from_numpy(uf).to(D.dev)
    d_p = torch.empty(UTT_FRAMES // MULTIEPOCH, dtype=torch.int64, device=D.dev)
    lens = np.array([UTT_FRAMES], dtype=np.int64)
    stream = torch.cuda.current_stream()

    def one():
        syn.db.greedy_batch_dev(d_t.data_ptr(), lens, d_p.data_ptr(), stream=stream.cuda_stream)

    syn.db.counters(reset=True)
    ms = D.timed(one, reps=5, warm=2)
    syn.db.greedy_batch_finish()
    path_one = d_p.cpu().numpy().copy()
    launches = syn.db.counters(reset=True)["launches"]
    os.environ["SNK_GREEDY_NO_ONE"] = "1"            # read by the library at call time: the batched path for one utterance
    try:
        ms_batched = D.timed(one, reps=3, warm=1)
        syn.db.greedy_batch_finish()
    finally:
        del os.environ["SNK_GREEDY_NO_ONE"]
    same = bool(np.array_equal(path_one, d_p.cpu().numpy()))
    syn.db.counters(reset=True)
    steps = UTT_FRAMES // MULTIEPOCH
    nprime = db["F"].shape[0] - (MULTIEPOCH - 1)
    # algorithmic bytes of a step: every searchable row's 160 join + 64 frame operand columns (fp16) once
    row_bytes = (160 + 64) * 2
    floor_us = nprime * row_bytes / (float(peaks["hbm_gbs"]) * 1e9) * 1e6
    t0 = time.perf_counter()
    syn.greedy_joint_search(uf)
    host_ms = (time.perf_counter() - t0) * 1e3
    return {"what": "one 648-frame utterance, B = 1, 108 dependent steps, inputs in HBM; one persistent cooperative kernel",
            "steps": steps, "us_per_step": ms * 1e3 / steps, "hbm_floor_us_per_step": floor_us,
            "frac_of_hbm_floor": floor_us / (ms * 1e3 / steps), "algorithmic_bytes_per_step": nprime * row_bytes,
            "achieved_gbs": nprime * row_bytes / (ms * 1e-3 / steps) / 1e9, "peak_kind": kind,
            "roofline": {"bound": "hbm", "kernel": "greedy_one_kernel (one launch = the whole utterance)",
                         "achieved": nprime * row_bytes / (ms * 1e-3 / steps) / 1e9, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                         "frac": floor_us / (ms * 1e3 / steps), "traffic": ncu_traffic("greedy_one_kernel_bytes_per_step"),
                         "traffic_source": "ncu --set full capture committed under profiles/ (dram bytes per step)",
                         "algorithmic_bytes": nprime * row_bytes},
            "kernel_launches_per_utterance": launches // 7,
            "batched_path_us_per_step": ms_batched * 1e3 / steps, "same_path_as_batched": same,
            "host_call_ms": host_ms, "host_call_frames_per_s": UTT_FRAMES / (host_ms / 1e3)}


def block_sharded_greedy(D, db, wt, wj, cfg):
    """greedy chain over a row-sharded joint database, one exchange per time step inside the
    library (snk_greedy_sharded_batch_dev).  configs[1] database split over the ranks; step latency at B = 1024 and B = 1;
    the paths of a few utterances are compared with the replicated single-GPU search on rank 0's own database."""
    import torch
    from snickery_b200 import Synthesiser, distributed as dd, engine
    rank, world, dev = D.rank, D.world, D.dev
    F, Jc = db["F"], db["Jc"]
    sg = dd.ShardedGreedy(F, Jc, MULTIEPOCH, wt, wj, rank, world, D.local)
    steps = 12
    res = {"what": "configs[1] joint rows sharded by row block over the ranks, join contexts replicated; one grouped "
                   "ncclAllGather (distance, global row, bound: B x 24 bytes per rank) + arg-min / certificate kernel per "
                   "time step, enqueued by the library; device time, max over ranks",
           "n_gpus": world, "rows_per_gpu": int(sg.knn.hi - sg.knn.lo), "steps": steps,
           "exchange": "one kernel over NVLink peer memory (CUDA IPC): stores into every peer, epoch flag, wait, arg-min"
                       if sg.knn.db.comm_info()["peer_exchange"] else
                       ("grouped ncclAllGather + arg-min kernel" if world > 1 else "none (one rank)"),
           "cases": []}
    for B in (1024, 1):
        tg = torch.from_numpy(make_batch(F, wt, B, steps * MULTIEPOCH, seed=31 + B).reshape(B, steps * MULTIEPOCH, -1)).to(dev)
        paths = [None]

        def run():
            paths[0] = sg.search(tg)

        run()                                  # NCCL connects lazily on the first collective: keep that out of the numbers
        sg.knn.db.profile_enable(True)
        ms = D.timed(run, reps=3, warm=0)
        ag = sg.knn.db.profile_read(engine.PROF_ALLGATHER)
        sg.knn.db.profile_enable(False)
        case = {"utterances": B, "us_per_step": ms * 1e3 / steps,
                # event time around the exchange on this rank: includes waiting for the slowest rank to arrive
                "us_per_step_in_allgather": (ag["ms"] * 1e3 / ag["launches"]) if ag["launches"] else 0.0,
                "nvlink_bytes_received_per_rank_per_step": B * 24 * (world - 1)}
        if B == 1024:       # parity sample against the replicated database
            case["paths_hash"] = int(paths[0].sum().item() % 1000003)
        else:               # one utterance: the persistent kernel with the exchange inside it (greedy_one.cu) when peers are mapped
            case["path"] = "one persistent kernel per utterance, exchange through peer memory inside it" \
                if (world > 1 and sg.knn.db.comm_info()["peer_exchange"]) else \
                ("one persistent kernel per utterance (one rank: nothing to exchange)" if world == 1 else
                 "batched path (three launches + exchange per step)")
            keep1 = paths[0][0].cpu().numpy()
        res["cases"].append(case)
        keep = paths[0][:4].cpu().numpy() if B == 1024 else None
        if B == 1024 and rank == 0:
            ref = Synthesiser(cfg, F, Jc, device=D.local)
            utts = [tg[b].cpu().numpy() for b in range(4)]
            want = ref.greedy_joint_search_batch(utts)
            res["first_4_paths_equal_replicated_search"] = bool(all(keep[b].tolist() == want[b] for b in range(4)))
            ref.db.close()
        if B == 1 and rank == 0:
            ref = Synthesiser(cfg, F, Jc, device=D.local)
            res["single_utterance_path_equals_replicated_search"] = bool(
                keep1.tolist() == ref.greedy_joint_search_batch([tg[0].cpu().numpy()])[0])
            ref.db.close()
    c = sg.knn.db.counters()
    res["exactness"] = {"utterances_recertified": int(c["recertified"]), "exhaustive_f64": int(c["exhaustive"])}
    sg.knn.db.close()
    return res


def block_halfphone(D, args, headline):
    """hybrid_halfphone_default.cfg shape: 90k halfphones (replicated), HP_UTTS utterances x 80 targets x 50 candidates per GPU.
    (a) acoustic pipeline on device-resident targets: k-NN (k = 50) -> join costs -> Viterbi, one library call;
    (b) the same through host buffers (pinned float64 targets in, paths out);
    (c) join + Viterbi alone on given candidate lattices (quinphone-style preselection hands the ids over);
    (d) parity of (a) for two utterances against the oracle (cKDTree k = 50, numpy join costs, min-plus DP)."""
    import torch
    from conftest import halfphone_config
    from snickery_b200 import Synthesiser, engine, synthetic as syn
    rank, world, dev = D.rank, D.world, D.dev
    hp = syn.make_halfphone_db(n_units=args.hp_units, seed=1237)
    cfg = halfphone_config(n_candidates=HP_K, preselection="acoustic")
    g = Synthesiser(cfg, hp["F"], hp["Jc"], device=D.local)
    B, T, K = HP_UTTS, HP_T, HP_K
    peaks, kind = measured_peaks()
    hbm = float(peaks["hbm_gbs"])
    tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    uf = np.vstack(syn.make_targets(hp["F"], B, T, seed=3 + 101 * rank)).astype(np.float64) * g.target_weight_vector
    pinned = torch.empty(uf.shape, dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = uf
    d_q = pinned.to(dev)
    d_dist = torch.empty((B * T, K), dtype=torch.float64, device=dev)
    d_idx = torch.empty((B * T, K), dtype=torch.int64, device=dev)
    d_paths = torch.empty(B * T, dtype=torch.int64, device=dev)
    d_plen = torch.empty(B, dtype=torch.int64, device=dev)
    d_cost = torch.empty((3, B), dtype=torch.float64, device=dev)
    lens = np.full(B, T, dtype=np.int64)
    stream = torch.cuda.current_stream()
    lib = engine.load_library()

    def pipeline():
        rc = lib.snk_acoustic_viterbi_batch_dev(g.db.handle, C.c_void_p(d_q.data_ptr()), lens.ctypes.data_as(C.POINTER(C.c_int64)),
                                                B, K, 0, C.c_void_p(d_paths.data_ptr()), C.c_void_p(d_plen.data_ptr()),
                                                C.c_void_p(d_cost[0].data_ptr()), C.c_void_p(d_cost[1].data_ptr()),
                                                C.c_void_p(d_cost[2].data_ptr()), C.c_void_p(stream.cuda_stream))
        rc = rc or lib.snk_acoustic_viterbi_finish(g.db.handle)
        if rc:
            raise RuntimeError(lib.snk_last_error().decode())

    def knn_only():
        g.db.knn_dev(d_q.data_ptr(), B * T, K, d_dist.data_ptr(), d_idx.data_ptr(), stream=stream.cuda_stream)
        g.db.knn_finish()

    def jv_only():
        g.db.join_viterbi_batch_dev(d_idx.data_ptr(), d_dist.data_ptr(), lens, K, d_paths.data_ptr(), d_plen.data_ptr(),
                                    d_cost[0].data_ptr(), d_cost[1].data_ptr(), d_cost[2].data_ptr(), stream=stream.cuda_stream)

    reps = max(2, min(args.steps, 5))
    if headline:
        # this block is the first GPU work of the process when it is the headline: a fresh process starts at idle clocks and
        # the first tens of milliseconds run slow (measured: 7.4 ms vs 12-21 ms for the same call) -- bring the clocks up first
        t_heat = time.perf_counter()
        while time.perf_counter() - t_heat < 0.5:
            knn_only()
        torch.cuda.synchronize()
    g.db.counters(reset=True)
    # as the headline (--workload halfphone) the contract's timing: K calls between one pair of events; as a block of the
    # default run the median of five calls (each ends in a host-side wait, one slow host iteration should not be averaged in)
    ms_pipe = D.timed(pipeline, reps, warm=2) if headline else D.timed_median(pipeline, 5, warm=2)
    # every call ends in a host-side wait; as the headline the mean over K calls now and then carries one slow host iteration
    # (7.4 ms vs 12-21 ms for the same device work), so the median of five calls is reported beside it
    ms_pipe_median = D.timed_median(pipeline, 5, warm=0) if headline else ms_pipe
    c = g.db.counters()
    found = int((d_plen > 0).sum().item())
    g.db.profile_enable(True)
    ms_knn = D.timed(knn_only, reps, warm=1)
    pk = g.db.profile_read(engine.PROF_KNN)
    pr = g.db.profile_read(engine.PROF_RERANK)
    ms_jv = D.timed(jv_only, reps, warm=1)
    pj = g.db.profile_read(engine.PROF_JOIN)
    pv = g.db.profile_read(engine.PROF_VITERBI)
    g.db.profile_enable(False)
    # (b) host buffers
    g.db.acoustic_viterbi_batch_cat(pinned.numpy(), lens, K, as_arrays=True)
    D.barrier()
    t0 = time.perf_counter()
    e2e_reps = 2
    for _ in range(e2e_reps):
        host_paths, _, _, _ = g.db.acoustic_viterbi_batch_cat(pinned.numpy(), lens, K, as_arrays=True)
    wall = D.max(time.perf_counter() - t0) / e2e_reps
    # (b') the reference's own call pattern: one utterance at a time (synth_halfphone.py synth_utt)
    lens1 = lens[:1].copy()

    def pipeline1():
        rc = lib.snk_acoustic_viterbi_batch_dev(g.db.handle, C.c_void_p(d_q.data_ptr()), lens1.ctypes.data_as(C.POINTER(C.c_int64)),
                                                1, K, 0, C.c_void_p(d_paths.data_ptr()), C.c_void_p(d_plen.data_ptr()),
                                                C.c_void_p(d_cost[0].data_ptr()), C.c_void_p(d_cost[1].data_ptr()),
                                                C.c_void_p(d_cost[2].data_ptr()), C.c_void_p(stream.cuda_stream))
        rc = rc or lib.snk_acoustic_viterbi_finish(g.db.handle)
        if rc:
            raise RuntimeError(lib.snk_last_error().decode())

    g.db.counters(reset=True)
    ms_one = D.timed_median(pipeline1, 7, warm=2)
    launches_one = g.db.counters()["launches"] // 9
    pipeline()                                   # leave the full batch's paths / costs in place for the parity sample
    frames_all = world * B * T
    knn_flops = 2.0 * B * T * args.hp_units * 184
    knn_tf = knn_flops / (ms_knn / 1e3) / 1e12
    nlaunch = max(reps + 1, 1)
    res = {}
    halfphone = {
        "workload": halfphone_description()["workload"],
        "value": frames_all / (ms_pipe / 1e3), "unit": UNIT, "ms_per_step": ms_pipe, "ms_per_step_median_of_5": ms_pipe_median, "steps": reps,
        "what": "k-NN (k=50, 184-dim) -> join costs -> Viterbi in one library call (snk_acoustic_viterbi_batch_dev), targets in HBM",
        "e2e": {"value": frames_all / wall, "unit": UNIT, "h2d_bytes_per_step": int(uf.nbytes), "d2h_bytes_per_step": int(B * T * 8 + B * 32),
                "steps": e2e_reps, "what": "pinned float64 targets in, path lists out (snk_acoustic_viterbi_batch); candidates stay on the device"},
        "paths_found": found, "utterances": B,
        "exactness": {"queries": c["queries"], "recertified_by_simt": c["recertified"], "exhaustive_f64": c["exhaustive"]},
        "knn_roofline": {"bound": "tensor", "kernel": "k = 50 search: knn_tc_kernel passes + shortlist scan + float64 re-rank (all launches)",
                         "achieved": knn_tf, "peak": tpeak, "unit": "TFLOP/s", "frac": knn_tf / tpeak,
                         "flops": knn_flops, "ms": ms_knn, "gemm_launch_ms": pk["ms"] / nlaunch,
                         "gemm_launches_per_search": pk["launches"] / nlaunch,
                         # the contraction alone (sampling pass over every 4th tile + emit pass over all of them), 184-dim shape
                         "gemm_only_tflops": pk["work"] / nlaunch / (pk["ms"] / nlaunch / 1e3) / 1e12 if pk["ms"] else None,
                         "gemm_only_frac_of_peak": (pk["work"] / nlaunch / (pk["ms"] / nlaunch / 1e3) / 1e12 / tpeak) if pk["ms"] else None,
                         "rerank_launch_ms": pr["ms"] / nlaunch,       # float64 re-rank + certificate of the shortlists
                         "selection_and_conversion_ms": ms_knn - (pk["ms"] + pr["ms"]) / nlaunch,   # bound, shortlist selection, query conversion, gaps
                         "peak_kind": kind},
        "single_utterance": {"ms": ms_one, "targets": T, "launches": launches_one,
                             "what": "the same call for ONE utterance of 80 targets (the reference synthesises one at a time), "
                                     "device-resident, including the host-side finish"},
        "join_viterbi": {"ms": ms_jv, "frames_per_s": world * B * T / (ms_jv / 1e3),
                         "what": "join costs + Viterbi on given candidate lattices (snk_join_viterbi_batch_dev)"},
    }
    for name, p in (("join_tiles", pj), ("viterbi", pv)):
        if p["launches"]:
            msl = p["ms"] / p["launches"]
            gbs = p["work"] / p["launches"] / (msl / 1e3) / 1e9
            halfphone["join_viterbi"][name] = {"ms": msl, "achieved_GBps": gbs, "peak_GBps": hbm, "frac": gbs / hbm,
                                               "algorithmic_bytes": p["work"] / p["launches"], "peak_kind": kind}
    # (d) parity sample on rank 0
    if rank == 0:
        from oracle import snickery_oracle as O
        o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
        o.build_acoustic_tree()
        mism = outside = 0
        max_rel = 0.0
        costs = d_cost[0].cpu().numpy()
        for b in (0, B - 1):
            ufb = uf[b * T:(b + 1) * T]
            rc, rd = o.preselect_units_acoustic(ufb)
            ref_path, ref_cost = O.viterbi_search_numpy(o, rc, rd, return_cost=True)
            if host_paths[b].tolist() != ref_path:
                mism += 1
                _, _, tot = o.path_costs(rc, rd, host_paths[b].tolist())
                outside += int(abs(tot - ref_cost) > 1e-6 * ref_cost)
            max_rel = max(max_rel, abs(costs[b] - ref_cost) / ref_cost)
        halfphone["parity"] = {"what": "2 utterances x 80 targets: CUDA pipeline vs oracle (cKDTree k=50 + numpy join + DP)",
                               "utterances": 2, "mismatches": mism, "mismatches_outside_1e-6_tie": outside,
                               "max_rel_cost_err": max_rel}
    res["halfphone"] = halfphone
    res["viterbi_sharded"] = {
        "what": "configs[3]: %d utterances x %d targets x %d candidates PER GPU, sharded by utterance, database replicated, "
                "no data-path collective; device time, max over ranks" % (B, T, K),
        "n_gpus": world, "utterances_total": world * B,
        "acoustic_pipeline": {"ms": ms_pipe, "frames_per_s": frames_all / (ms_pipe / 1e3)},
        "join_viterbi_from_candidates": {"ms": ms_jv, "frames_per_s": frames_all / (ms_jv / 1e3)},
        "knn_k50": {"ms": ms_knn, "frames_per_s": frames_all / (ms_knn / 1e3)},
    }
    if headline:
        res = {"value": halfphone["value"], "ms_per_step": ms_pipe, "dtype": "fp16 tensor-core shortlist + f64 re-rank; f32 join costs / Viterbi",
               "e2e": halfphone["e2e"], "gpu_launches": int(c["launches"]),
               "roofline": dict(halfphone["knn_roofline"], traffic=None), "halfphone": halfphone,
               "viterbi_sharded": res["viterbi_sharded"]}
    g.db.close()
    return res


def block_sharded_knn(D, args):
    """configs[4]: the database sharded by row block over the ranks (SHARD_ROWS joint rows of 517 dims per GPU; 10 M at
    N = 8), queries replicated, local certified top-k -> ncclAllGather -> merge, all inside the library
    (snk_knn_sharded_dev).  Device time per search = max over ranks; the split comes from CUDA events around the GEMM
    launches, the all-gather and the merge kernel.  The answer is checked at a reduced size against single-GPU searches
    of every block on rank 0 merged in numpy, and at full size against a numpy merge of the gathered local lists."""
    import torch
    from snickery_b200 import distributed as SD, engine, synthetic as syn
    rank, world, dev = D.rank, D.world, D.dev
    m = MULTIEPOCH
    wt, wj = np.full(61, 0.4), np.full(151, 0.05)
    rows = args.shard_rows

    def block(b, n):        # block b of the database: an independent set of recordings (no window straddles two blocks)
        return syn.make_epoch_db(n_units=n + m - 1, seed=5000 + b)

    def make_queries(F, Jc, nq, seed):
        rng = np.random.default_rng(seed)
        r = rng.integers(0, F.shape[0] - m, nq)
        Fw, Jw = F.astype(np.float64) * wt, Jc.astype(np.float64) * wj
        return np.hstack([Jw[r]] + [Fw[r + j] for j in range(m)]) + 0.02 * rng.standard_normal((nq, 151 + 61 * m))

    def search(sk, qd, k):
        d = torch.empty((qd.shape[0], k), dtype=torch.float64, device=dev)
        i = torch.empty((qd.shape[0], k), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def fn():
            if world > 1:
                sk.db.knn_sharded_dev(qd.data_ptr(), qd.shape[0], k, d.data_ptr(), i.data_ptr(), engine.SPACE_JOINT, sk.lo, st)
                sk.db.knn_sharded_finish()
            else:
                sk.db.knn_dev(qd.data_ptr(), qd.shape[0], k, d.data_ptr(), i.data_ptr(), engine.SPACE_JOINT, sk.lo, st)
                sk.db.knn_finish()
        return fn, d, i

    def make_shard(n):
        b = block(rank, n)
        sk = SD.ShardedKnn.__new__(SD.ShardedKnn)
        sk.rank, sk.world, sk.group, sk.device = rank, world, None, D.local
        sk.lo, sk.hi = rank * n, (rank + 1) * n
        sk.db = engine.UnitDatabase(b["F"], b["Jc"], multiepoch=m, device=D.local)
        sk.db.set_weights(wt, wj)
        sk.space = engine.SPACE_JOINT
        if world > 1:
            SD.init_comm(sk.db)
        return sk, b

    # ---- correctness at reduced size: distributed answer == numpy merge of single-GPU searches of every block (rank 0)
    n_small = 30000
    sk, b_own = make_shard(n_small)
    b0 = b_own if rank == 0 else block(0, n_small)
    Qs = make_queries(b0["F"], b0["Jc"], 512, seed=17)
    qd = torch.from_numpy(Qs).to(dev)
    check = {"rows_per_gpu": n_small, "queries": 512, "ok": True}
    for k in (1, 50):
        fn, d, i = search(sk, qd, k)
        fn()
        torch.cuda.synchronize()
        if rank == 0:
            dd, ii = [], []
            for b in range(world):
                bb = b_own if b == 0 else block(b, n_small)
                one = engine.UnitDatabase(bb["F"], bb["Jc"], multiepoch=m, device=D.local)
                one.set_weights(wt, wj)
                db_, ib_ = one.knn(Qs, k, engine.SPACE_JOINT)
                dd.append(db_)
                ii.append(ib_ + b * n_small)
                one.close()
            dd, ii = np.hstack(dd), np.hstack(ii)
            order = np.lexsort((ii, dd), axis=1)[:, :k]
            ok = np.array_equal(np.take_along_axis(ii, order, 1), i.cpu().numpy()) and \
                np.array_equal(np.take_along_axis(dd, order, 1), d.cpu().numpy())
            check["k%d_equals_single_gpu_merge" % k] = bool(ok)
            check["ok"] = check["ok"] and bool(ok)
    sk.db.close()
    del sk
    D.barrier()

    # ---- timed at full size
    sk, b_own = make_shard(rows)
    b0 = b_own if rank == 0 else None
    cases = []
    nq_max = max(args.shard_queries)
    if rank == 0:
        Q = make_queries(b_own["F"], b_own["Jc"], nq_max, seed=23)
        qall = torch.from_numpy(Q).to(dev)
    else:
        qall = torch.empty((nq_max, 151 + 61 * m), dtype=torch.float64, device=dev)
    if world > 1:
        D.dist.broadcast(qall, src=0)          # queries replicated
    peaks, kind = measured_peaks()
    tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    for nq in args.shard_queries:
        for k in (1, 50):
            qd = qall[:nq].contiguous()
            fn, d, i = search(sk, qd, k)
            fn()
            sk.db.profile_enable(True)
            reps = 3
            ms = D.timed(fn, reps, warm=1)
            pg = sk.db.profile_read(engine.PROF_KNN)
            pa = sk.db.profile_read(engine.PROF_ALLGATHER)
            pm = sk.db.profile_read(engine.PROF_MERGE)
            sk.db.profile_enable(False)
            n_calls = reps + 1
            ms_gemm, ms_ag, ms_merge = (D.max(p["ms"] / n_calls) for p in (pg, pa, pm))
            # full-size property: the merged answer equals a numpy merge of the gathered per-rank lists
            prop = None
            if world > 1:
                ld = torch.empty((nq, k), dtype=torch.float64, device=dev)
                li = torch.empty((nq, k), dtype=torch.int64, device=dev)
                st = torch.cuda.current_stream().cuda_stream
                sk.db.knn_dev(qd.data_ptr(), nq, k, ld.data_ptr(), li.data_ptr(), engine.SPACE_JOINT, sk.lo, st)
                sk.db.knn_finish()
                sample = slice(0, min(nq, 256))
                gd = [torch.empty_like(ld[sample]) for _ in range(world)]
                gi = [torch.empty_like(li[sample]) for _ in range(world)]
                D.dist.all_gather(gd, ld[sample].contiguous())
                D.dist.all_gather(gi, li[sample].contiguous())
                if rank == 0:
                    dd, ii = torch.cat(gd, 1).cpu().numpy(), torch.cat(gi, 1).cpu().numpy()
                    order = np.lexsort((ii, dd), axis=1)[:, :k]
                    prop = bool(np.array_equal(np.take_along_axis(ii, order, 1), i[sample].cpu().numpy()))
            flops = 2.0 * nq * rows * (151 + 61 * m)
            cases.append({"queries": nq, "k": k, "ms": ms, "queries_per_s": nq / (ms / 1e3),
                          "ms_gemm": ms_gemm, "ms_allgather": ms_ag, "ms_merge": ms_merge,
                          "ms_other": max(ms - ms_gemm - ms_ag - ms_merge, 0.0),
                          "frac_in_collective": (ms_ag + ms_merge) / ms if ms > 0 else None,
                          "nvlink_bytes_received_per_rank": int(nq * k * 16 * (world - 1)),
                          "allgather_GBps_per_rank": (nq * k * 16 * (world - 1)) / (ms_ag / 1e3) / 1e9 if ms_ag > 0 else None,
                          "gemm_tflops_per_gpu": flops / (ms_gemm / 1e3) / 1e12 if ms_gemm > 0 else None,
                          "search_tflops_per_gpu": flops / (ms / 1e3) / 1e12, "frac_of_tensor_peak": flops / (ms / 1e3) / 1e12 / tpeak,
                          "merged_equals_numpy_merge_of_local_lists": prop})
    cnt = sk.db.counters()
    info = sk.db.comm_info() if world > 1 else {"nranks": 1, "nccl_version": None}
    sk.db.close()
    lim = None
    if world > 1 and cases:
        worst = max(cases, key=lambda c_: c_["frac_in_collective"] or 0)
        lim = "ncclAllGather" if worst["ms_allgather"] >= worst["ms_merge"] else "merge kernel"
    return {"what": "configs[4]: %d joint rows (517-dim) per GPU, %d rows in all; queries replicated; local top-k -> "
                    "ncclAllGather -> merge inside the library; device time, max over ranks" % (rows, rows * world),
            "rows_per_gpu": rows, "rows_total": rows * world, "dim": 151 + 61 * m, "n_gpus": world,
            "nccl_ranks": info["nranks"], "nccl_version": info["nccl_version"], "cases": cases, "check": check,
            "limiting_collective": lim, "peak_kind": kind,
            "exactness": {"queries": cnt["queries"], "recertified_by_simt": cnt["recertified"], "exhaustive_f64": cnt["exhaustive"]}}


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    communicator when NCCL_DEBUG asks for it), so file descriptor 1 is pointed at stderr for the whole run and the line
    goes out through a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="greedy", choices=["greedy", "halfphone"])
    ap.add_argument("--utts", type=int, default=1024, help="target utterances per GPU per step")
    ap.add_argument("--db-units", type=int, default=DB_UNITS)
    ap.add_argument("--hp-units", type=int, default=HP_UNITS)
    ap.add_argument("--shard-rows", type=int, default=SHARD_ROWS)
    ap.add_argument("--shard-queries", type=int, nargs="+", default=[1024, 65536])
    ap.add_argument("--engine", default="", choices=["", "auto", "simt", "tc"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
