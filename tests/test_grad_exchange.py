"""N > 1 host logic on CPU: the [shared | domain] gradient exchange over a world_size-2 gloo group must
equal the reference's dense all-reduce over the full parameter vector (train_multi.py:579,779-781)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rank_domains, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hma_b200.train import exchange_gradients

    shared, dom_sizes = 64, {"a": 24, "b": 16, "c": 20}
    dom_range, off = {}, shared
    for k, n in dom_sizes.items():
        dom_range[k] = (off, n)
        off += n
    total = off
    max_dom = max(dom_sizes.values())
    g = torch.Generator().manual_seed(100 + rank)
    dom = rank_domains[rank]
    # dense gradient of this rank: shared part + its own domain block, zeros elsewhere
    dense = torch.zeros(total)
    dense[:shared] = torch.randn(shared, generator=g)
    lo, n = dom_range[dom]
    dense[lo:lo + n] = torch.randn(n, generator=g)
    flat = torch.zeros(shared + max_dom)
    flat[:shared] = dense[:shared]
    flat[shared:shared + n] = dense[lo:lo + n]
    gathered = torch.zeros(world, max_dom)
    updates = exchange_gradients(flat, shared, max_dom, dom_range, rank_domains, gathered)
    mine = torch.zeros(total)
    mine[:shared] = flat[:shared]
    for lo2, n2, t in updates:
        mine[lo2:lo2 + n2] = t
    ref = dense.clone()
    dist.all_reduce(ref)  # what DDP's all-reduce over every parameter would produce (before the /world)
    q.put((rank, torch.allclose(mine, ref, atol=1e-6), sorted(lo2 for lo2, _, _ in updates)))
    dist.destroy_process_group()


@pytest.mark.parametrize("rank_domains", [["a", "b"], ["c", "c"]])
def test_exchange_equals_dense_allreduce(rank_domains):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rank_domains, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in results), results
    assert results[0][2] == results[1][2]  # every rank applies the same set of domain updates
