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
    # bucketed reducer (the path bench.py uses): hooks on the leaves, the large slot goes out early (async), the small tail in
    # one call at finish(); a parameter that received no gradient this step contributes zeros
    from nvdiffrast_b200.distributed import SharedGradReducer
    big = torch.arange(40, dtype=torch.float32).reshape(2, 20).requires_grad_(True)
    small = torch.ones(3, 4).requires_grad_(True)
    unused = torch.ones(5).requires_grad_(True)
    red = SharedGradReducer([("small", small.shape), ("big", big.shape), ("unused", unused.shape)], "cpu", early_bytes=64)
    assert red.early == {"big"} and red.bytes_per_step == (40 + 12 + 5) * 4
    for step in range(2):                                        # reused across steps
        red.begin({"big": big, "small": small, "unused": unused})
        big.grad = None; small.grad = None
        ((big * (rank + 1 + step)).sum() + (small * small * (rank + 2)).sum()).backward()
        out = red.finish()
    if rank == 0:
        q.put((g_attr.numpy(), g_pos_sum.numpy(), {k: v.clone().numpy() for k, v in out.items()}))
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
    got_attr, got_pos, red = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    n_total, res = 3, 40
    c = wl.config_c2(n=n_total, res=res)
    dy = np.random.default_rng(0).normal(size=(n_total, res, res, 3)).astype(np.float32)
    _, _, _, g = oracle_chain(c, n_total, False, dy)
    assert np.allclose(got_attr, g["attr"], rtol=1e-5, atol=1e-5)
    assert np.allclose(got_pos, g["pos"].sum(0), rtol=1e-4, atol=1e-4)
    # reducer, second step (step = 1): d/d big = (rank + 2) summed over ranks 0, 1 = 5; d/d small = 2 (rank + 2) summed = 10; unused = 0
    assert np.array_equal(red["big"], np.full((2, 20), 5.0, np.float32))
    assert np.array_equal(red["small"], np.full((3, 4), 10.0, np.float32))
    assert np.array_equal(red["unused"], np.zeros(5, np.float32))
