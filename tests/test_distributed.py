"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard maps, all-gather + k-way merge
ordering, utterance-sharded path gathering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snickery_b200 import distributed as D


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            spans = [D.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_utterances_partition():
    for n in (0, 5, 1024):
        for world in (1, 2, 4, 8):
            got = sorted(sum((D.shard_utterances(n, r, world) for r in range(world)), []))
            assert got == list(range(n))


def test_merge_reference_tie_rule():
    d = torch.tensor([[[0.5, 1.0, 3.0]], [[0.5, 1.0, 2.0]]], dtype=torch.float64)      # [R=2, nq=1, k=3]
    i = torch.tensor([[[10, 4, 7]], [[3, 12, 9]]], dtype=torch.int64)
    md, mi = D.merge_topk_reference(d, i)
    assert md.tolist() == [[0.5, 0.5, 1.0]] and mi.tolist() == [[3, 10, 4]]           # ties: lowest global id


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- sharded k-NN: each rank searches its row block (numpy stands in for the GPU search)
        rng = np.random.default_rng(0)
        data = rng.standard_normal((203, 9))
        q = rng.standard_normal((11, 9))
        k = 5
        lo, hi = D.shard_rows(data.shape[0], rank, world)
        d2 = ((data[lo:hi][None, :, :] - q[:, None, :]) ** 2).sum(-1)
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        dl = torch.from_numpy(np.sqrt(np.take_along_axis(d2, order, 1)))
        il = torch.from_numpy(order + lo)
        md, mi = D.allgather_merge_topk(dl, il)
        full = ((data[None, :, :] - q[:, None, :]) ** 2).sum(-1)
        want = np.argsort(full, axis=1, kind="stable")[:, :k]
        assert np.array_equal(mi.numpy(), want)
        assert np.allclose(md.numpy(), np.sqrt(np.take_along_axis(full, want, 1)))
        # --- utterance sharding: each rank "searches" its utterances, everyone sees all paths in order
        n_utts = 7
        mine = D.shard_utterances(n_utts, rank, world)
        paths = [[u * 100 + t for t in range(u + 1)] for u in mine]
        allp = D.gather_paths(paths, mine, n_utts)
        assert allp == [[u * 100 + t for t in range(u + 1)] for u in range(n_utts)]
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allgather_merge_and_path_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
