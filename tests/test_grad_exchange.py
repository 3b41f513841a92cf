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


def test_shared_segment_ranges_cover_the_shared_range_once_in_backward_order():
    """The overlapped exchange all-reduces the shared gradient range in pieces, each as soon as the backward has finished
    the layers it belongs to: the pieces must tile the range exactly once; readout first, embeddings last."""
    from hma_b200.train import shared_segment_ranges

    L = 8
    names = ["pos_embed_TSC", "token_embed.mask_token_embed", "token_embed.factored_embeds.0.weight"]
    names += [f"decoder.layers.{i}.mlp.fc1.weight" for i in range(L)] + ["out_x_proj.weight"]
    names += [f"decoder.layers.{i}.mlp.fc1.bias" for i in range(L)] + ["out_x_proj.bias"]  # [decayed | not decayed] layout
    sizes = [12, 4, 8] + [16] * L + [20] + [4] * L + [4]
    for segs in (1, 2, 4, 8, 16):
        lows, ranges = shared_segment_ranges(names, sizes, L, segs)
        assert lows[-1] == 0 and lows == sorted(lows, reverse=True) and len(ranges) == len(lows)
        covered = sorted((o, n) for r in ranges for o, n in r)
        pos = 0
        for o, n in covered:
            assert o == pos
            pos += n
        assert pos == sum(sizes)
        offs = {k: sum(sizes[:i]) for i, k in enumerate(names)}
        seg_of = lambda k: next(s for s, r in enumerate(ranges) if any(o <= offs[k] < o + n for o, n in r))  # noqa: E731
        assert seg_of("out_x_proj.weight") == 0 and seg_of("out_x_proj.bias") == 0
        assert seg_of("pos_embed_TSC") == len(ranges) - 1
        assert seg_of(f"decoder.layers.{L - 1}.mlp.fc1.weight") == 0 and seg_of("decoder.layers.0.mlp.fc1.bias") == len(ranges) - 1
        for i in range(L - 1):  # a later layer never lands in a later segment than an earlier one
            assert seg_of(f"decoder.layers.{i + 1}.mlp.fc1.weight") <= seg_of(f"decoder.layers.{i}.mlp.fc1.weight")
