"""N=2 data-parallel parity on real GPUs. Run with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/ddp_parity_2gpu.py

(`gpurun --gpus 2`; tests/test_ddp_gpu.py launches it when two devices are visible). Each rank runs forward + loss +
backward of the tiny golden model on ITS OWN batch; then

  * TrainStep's exchange (all-reduce of the shared gradient range + all-gather of the per-domain range, train.py) must equal
    the reference's DDP semantics — a dense all-reduce(sum) over EVERY parameter gradient of the autograd path
    (train_multi.py:779-781: accelerate's DDP averages; the 1/world factor is applied by the optimizer here) — for both the
    case where the two ranks train different action domains and the case where they train the same one;
  * after one full step (exchange + clip + AdamW) the parameters of the two ranks are bit-identical (replicas stay in sync)
    and equal to a single-process step on the concatenated gradients within fp32 rounding.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from hma_b200.train import TrainStep, exchange_gradients
    from tests._util import build_cuda_model, golden

    rec, cfg, sd = golden()
    doms = rec["domains"]
    for case, rank_doms in (("different domains", [doms[0], doms[1]]), ("same domain", [doms[0], doms[0]])):
        dom = rank_doms[rank]
        r = rec[dom]
        g = torch.Generator().manual_seed(100 + rank)  # each rank its own masks over the fixture's tokens
        labels = r["labels"].clone()
        mask = torch.rand(labels.shape, generator=g) < 0.5
        mask.view(2, cfg.T, -1)[:, 0] = False
        ids = torch.where(mask, torch.full_like(labels, cfg.mask_token_id), labels).to(dev)
        labels, acts = labels.to(dev), (r["actions"] + 0.1 * rank).to(dev)
        # (a) autograd path + dense all-reduce over every parameter (what DDP does)
        ref = build_cuda_model(rec, sd, device=dev)
        out = ref(ids, labels, action_ids=acts, domain=[dom, dom])
        out.loss.backward()
        dense = {}
        for k, p in ref.named_parameters():
            gk = p.grad.detach().clone().float() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gk)
            dense[k] = gk
        # (b) TrainStep: forward + loss + backward into the flat buffer, then the two-collective exchange
        model = build_cuda_model(rec, sd, device=dev)
        step = TrainStep(model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, overlap_segments=1)  # exchange called by hand below
        p = step._params()
        d = step.engine.dims(2, cfg.T, cfg.S, True)
        step._fwd_bwd(p, ids.reshape(2, cfg.T, -1).contiguous(), labels.reshape(2, -1).contiguous(), acts, dom, d)
        updates = exchange_gradients(step.grad, step.arena.shared_size, step.arena.max_dom_size, step.arena.dom_range, rank_doms,
                                     step.gathered, None)
        named = dict(model.named_parameters())
        checked = 0
        worst = 0.0

        def compare(k, got):
            nonlocal checked, worst
            want = dense[k]
            scale = want.abs().max().item()
            err = (got.reshape(want.shape) - want).abs().max().item()
            # the two paths run the same kernels on the same data: they differ by the order of fp32 atomics / reductions only
            assert err <= 2e-3 * scale + 1e-7, (case, k, err, scale)
            worst = max(worst, err / max(scale, 1e-12))
            checked += 1

        off = 0
        for k in step.engine.shared_param_names(named, d):
            n = named[k].numel()
            compare(k, step.grad[off:off + n])
            off += step.engine.padded_numel(named[k])
        for lo, n_r, gbuf in updates:
            dname = [dn for dn, (l0, _) in step.arena.dom_range.items() if l0 == lo][0]
            o2 = 0
            for k in step.engine.domain_param_names(named, d, dname, True):
                compare(k, gbuf[o2:o2 + named[k].numel()])
                o2 += step.engine.padded_numel(named[k])
        assert len(updates) == len(set(rank_doms))
        # parameters nobody trained this step have zero dense gradient (the reference all-reduces those zeros)
        touched = set(step.engine.shared_param_names(named, d))
        for dn in set(rank_doms):
            touched |= set(step.engine.domain_param_names(named, d, dn, True))
        for k, gk in dense.items():
            if k not in touched:
                assert gk.abs().max().item() == 0.0, k
        # (c) full optimisation steps keep the replicas bit-identical, and the OVERLAPPED exchange (shared range all-reduced
        # segment by segment during the backward; eager launches and one CUDA graph per segment) gives the parameters of the
        # single all-reduce after the backward (a two-rank sum does not depend on how the range is cut)
        finals = {}
        for tag, kw2 in (("single all-reduce", dict(overlap_segments=1)), ("overlapped eager", dict(overlap_segments=2)),
                         ("overlapped graphs", dict(overlap_segments=2, cuda_graphs=True))):
            model2 = build_cuda_model(rec, sd, device=dev)
            step2 = TrainStep(model2, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, **kw2)
            for _ in range(3):  # graphs: eager warm-up, capture + replay, replay
                step2(ids, labels, acts, [dom, dom], rank_domains=rank_doms)
            flat = step2.arena.flat.detach().clone()
            other = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(other, flat)
            assert torch.equal(other[0], other[1]), f"{case} / {tag}: replicas diverged"
            finals[tag] = flat
            if tag == "overlapped graphs":
                assert len(next(iter(step2._graphs.values()))["graphs"]) == 2
        ref_flat = finals["single all-reduce"]
        moved = (ref_flat - TrainStep(build_cuda_model(rec, sd, device=dev)).arena.flat).abs().mean().item()
        for tag in ("overlapped eager", "overlapped graphs"):
            diff = (finals[tag] - ref_flat).abs().mean().item()
            # weight-gradient atomics make two runs of the same step differ in the last bits; three Adam steps later the
            # parameters agree to a small fraction of the distance moved
            assert diff <= 0.05 * moved, (case, tag, diff, moved)
        if rank == 0:
            print(f"[ddp parity] {case}: {checked} gradient tensors equal the dense all-reduce (worst rel err {worst:.2e}); "
                  f"replicas bit-identical after clip + AdamW; overlapped exchange (eager, graphs) == single all-reduce", flush=True)
    dist.barrier()
    if rank == 0:
        print("DDP PARITY OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
