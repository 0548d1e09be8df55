"""world_size-2 gloo tests (CPU) of the batch-sharding host logic in papc_b200/dist.py
(SURVEY.md 8e): shard ranges, the one all-gather of per-shard features, and the SyncBN exchange
(all-reduce of per-layer [2,C] fp64 sums) that makes sharded == unsharded batch statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from papc_b200 import dist as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import layers_np
        from papc_b200 import synth

        B, N, S, K = 4, 128, 32, 8
        xyz = synth.clouds(B, N, seed=0)                      # [B,3,N]
        start = synth.fps_start(B, N, seed=1)
        params = synth.mlp_params(3, [16, 32], seed=2)

        def make():
            sa = layers_np.PointNetSetAbstraction(S, 0.4, K, 3, [16, 32], False)
            for l, p in enumerate(params):
                sa.mlp_convs[l].weight = p["weight"].reshape(*p["weight"].shape, 1, 1)
                sa.mlp_convs[l].bias = p["bias"]
            for bn in sa.mlp_bns:
                bn.training = False                           # running statistics: no cross-shard coupling
            return sa

        full_xyz, full_pts = make()(xyz, None, start_idx=start)
        lo, hi = pdist.shard_range(B, rank, world)
        assert (lo, hi) == (rank * B // world, (rank + 1) * B // world)
        shard = pdist.shard_batch(torch.from_numpy(xyz)).numpy()
        assert np.array_equal(shard, xyz[lo:hi])
        _, loc_pts = make()(shard, None, start_idx=start[lo:hi])
        gathered = pdist.all_gather_features(torch.from_numpy(np.ascontiguousarray(loc_pts)))
        ok_gather = np.array_equal(gathered.numpy(), full_pts)

        # SyncBN exchange: per-shard sums of y, y^2 -> all-reduce -> statistics of the whole batch
        rng = np.random.default_rng(7)
        y = rng.standard_normal((B * 64, 24)).astype(np.float32)  # rows of all shards
        rows = y.reshape(B, 64, 24)[lo:hi].reshape(-1, 24).astype(np.float64)
        sums = torch.from_numpy(np.stack([rows.sum(0), (rows * rows).sum(0)]))
        pdist.all_reduce_sums_(sums)
        cnt = float(y.shape[0])
        mean = sums[0].numpy() / cnt
        var = sums[1].numpy() / cnt - mean * mean
        y64 = y.astype(np.float64)
        ok_bn = np.allclose(mean, y64.mean(0), rtol=0, atol=1e-12) and np.allclose(var, y64.var(0), rtol=0, atol=1e-12)
        q.put((rank, bool(ok_gather), bool(ok_bn), tuple(gathered.shape)))
    finally:
        dist.destroy_process_group()


def test_shard_range_rejects_ragged():
    with pytest.raises(ValueError):
        pdist.shard_range(5, 0, 2)
    assert pdist.shard_range(256, 3, 8) == (96, 128)


def test_single_process_is_identity():
    x = torch.arange(12.0).reshape(4, 3)
    assert pdist.all_gather_features(x) is x
    s = torch.ones(2, 3, dtype=torch.float64)
    assert pdist.all_reduce_sums_(s) is s


@pytest.mark.timeout(300)
def test_world2_gloo_gather_and_syncbn():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, ok_bn, shape in res:
        assert ok_gather, f"rank {rank}: all-gathered shard outputs differ from the unsharded run"
        assert ok_bn, f"rank {rank}: all-reduced BatchNorm sums differ from whole-batch statistics"
        assert shape[0] == 4
