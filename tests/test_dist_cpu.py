"""World-size-2 gloo test of the N>1 host logic: batch sharding + the single all-reduce of shared-parameter gradients
reproduces the single-process gradient (computed with the CPU oracle; no GPU involved)."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    import workloads as wl
    from nvdiffrast_b200.distributed import shard_range, allreduce_shared_grads
    from util import oracle_chain
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, res = 3, 40
    c = wl.config_c2(n=n_total, res=res)
    dy = np.random.default_rng(0).normal(size=(n_total, res, res, 3)).astype(np.float32)
    a, b = shard_range(n_total, rank, world)
    cs = dict(c); cs["pos"] = c["pos"][a:b]
    _, _, _, g = oracle_chain(cs, b - a, False, dy[a:b])
    g_attr = torch.from_numpy(g["attr"].copy())                  # shared parameter: vertex colours
    g_pos_sum = torch.from_numpy(g["pos"].sum(0).copy())         # stand-in for a shared-mesh gradient
    allreduce_shared_grads([g_attr, g_pos_sum])
    if rank == 0:
        q.put((g_attr.numpy(), g_pos_sum.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_match_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import workloads as wl
    from util import oracle_chain
    from nvdiffrast_b200.distributed import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(256, r, 8) for r in range(8)] == [(32 * r, 32 * r + 32) for r in range(8)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_attr, got_pos = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    n_total, res = 3, 40
    c = wl.config_c2(n=n_total, res=res)
    dy = np.random.default_rng(0).normal(size=(n_total, res, res, 3)).astype(np.float32)
    _, _, _, g = oracle_chain(c, n_total, False, dy)
    assert np.allclose(got_attr, g["attr"], rtol=1e-5, atol=1e-5)
    assert np.allclose(got_pos, g["pos"].sum(0), rtol=1e-4, atol=1e-4)
