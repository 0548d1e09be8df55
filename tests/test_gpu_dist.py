"""NCCL world-size-2 GPU test of the batch-sharded path: with SyncBN on
(``papc_b200.dist.set_sync_bn``) every rank's shard of the SSG SetAbstraction stack equals the matching slice of
the UNSHARDED forward (3e-5 at the third level of the chain, see tests/test_gpu_fullsize.py; train-mode BatchNorm over the whole batch, reference layers.py:214-219), the
running statistics agree, and one all-gather rebuilds the full feature tensor.  Without SyncBN the shards use
per-shard statistics (the bench's weak-scaling mode) and must differ.

Needs 2 GPUs: skipped on a 1-GPU box (run it with ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu``)."""
import os
import socket
import traceback

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    from papc_b200 import sa_stack, synth
    model = sa_stack.SSGSetAbstractionStack().to(dev)
    cfg = [(3, [64, 64, 128]), (131, [128, 128, 256]), (259, [256, 512, 1024])]
    for i, sa in enumerate(model.layers_()):
        sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(cfg[i][0], cfg[i][1], seed=30 + i))
    return model


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from papc_b200 import dist as pdist
        from papc_b200 import synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        B, N = 8, 1024
        xyz = torch.from_numpy(synth.clouds(B, N, seed=5)).to(dev)
        st1 = torch.from_numpy(synth.fps_start(B, N, seed=6)).to(dev)
        st2 = torch.zeros(B, dtype=torch.int64, device=dev)
        lo, hi = pdist.shard_range(B, rank, world)

        full = _build(dev)                                    # unsharded: the whole batch on this GPU
        _, want = full(xyz, None, start_idx=(st1, st2))
        want = want.reshape(B, 1024)

        synced = pdist.set_sync_bn(_build(dev))               # sharded, statistics all-reduced per layer
        _, got = synced(xyz[lo:hi].contiguous(), None, start_idx=(st1[lo:hi].contiguous(), st2[lo:hi].contiguous()))
        got = got.reshape(hi - lo, 1024)
        err = float((got - want[lo:hi]).abs().max())
        gathered = pdist.all_gather_features(got)
        err_g = float((gathered - want).abs().max())
        # running statistics of a middle layer: the same update as the unsharded model's
        rm_f = full.sa2.mlp_bns[1]._mean
        rm_s = synced.sa2.mlp_bns[1]._mean
        err_rm = float((rm_f - rm_s).abs().max())

        local = _build(dev)                                   # sharded, per-shard statistics (bench default)
        _, loc = local(xyz[lo:hi].contiguous(), None, start_idx=(st1[lo:hi].contiguous(), st2[lo:hi].contiguous()))
        err_local = float((loc.reshape(hi - lo, 1024) - want[lo:hi]).abs().max())
        # the same exchange over peer memory (PeerAllGather) == the NCCL all-gather, bit for bit, over reuse
        p2p = "unavailable"
        try:
            ag = pdist.PeerAllGather(hi - lo, 1024, dev)
        except Exception as e:  # noqa: BLE001  (symmetric memory not supported on this box: reported, not failed)
            ag, p2p = None, f"unavailable: {type(e).__name__}: {e}"
        if ag is not None:
            ok = True
            for k in range(3):
                loc_k = (got + float(k)).contiguous()
                ag.pre()
                full = ag.gather(loc_k)
                ok = ok and bool(torch.equal(full, pdist.all_gather_features(loc_k)))
            p2p = "equal" if ok else "DIFFERENT"
        torch.cuda.synchronize()
        dist.barrier()
        q.put((rank, "ok", err, err_g, err_rm, err_local, p2p))
        q.close()
        q.join_thread()   # the result is in the pipe before the process leaves
        os._exit(0)       # symmetric-memory handles + NCCL teardown order is not worth a hang in a test process
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc(), 0, 0, 0, ""))


def test_syncbn_sharded_equals_unsharded_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL world size 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, err, err_g, err_rm, err_local, p2p in sorted(res):
        assert status == "ok", err
        print(f"rank {rank}: synced shard vs unsharded {err:.3e}, gathered {err_g:.3e}, running mean {err_rm:.3e}, "
              f"per-shard statistics differ by {err_local:.3e}; peer-memory all-gather vs NCCL: {p2p}")
        assert p2p == "equal" or p2p.startswith("unavailable"), p2p
        # l3 features (|values| up to ~8) after three chained levels computed by two different kernel paths
        # (fused vs step-wise): the chain bound of tests/test_gpu_fullsize.py; measured 1.9e-5 on B200
        assert err <= 3e-5 and err_g <= 3e-5 and err_rm <= 1e-6
        assert err_local > 1e-4   # per-shard BatchNorm is a different function: SyncBN is what closes the gap
