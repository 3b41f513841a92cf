"""TrainStep (flat parameter arena, fused clip + AdamW, optional CUDA-graph replay of forward+loss+backward)
against torch: the same model driven through autograd + torch.optim.AdamW + clip_grad_norm_
(the optimiser calls of the reference trainer, train_multi.py:593-598)."""
import copy

import pytest
import torch

from tests._util import build_cuda_model, golden

pytestmark = pytest.mark.gpu


def _batches(rec):
    out = []
    for dom in rec["domains"]:
        r = rec[dom]
        out.append((r["input_ids"].cuda(), r["labels"].cuda(), r["actions"].cuda(), [dom, dom]))
    return out


def reference_param_groups(model, weight_decay):
    """The optimizer grouping of the reference trainer, verbatim (train_multi.py:906-917)."""
    no_decay = ["bias", "layer_norm.weight"]
    return [
        {"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
        {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]


def test_weight_decay_groups_follow_the_reference():
    """Zero gradients isolate the decay term: after a step with lr*wd = 0.5 every parameter the reference decays is
    halved and every `bias` is untouched (the kernel decays exactly the [decayed | not decayed] prefix of each range)."""
    from hma_b200.train import TrainStep

    rec, cfg, sd = golden()
    model = build_cuda_model(rec, sd)
    with torch.no_grad():
        for p in model.parameters():
            p.fill_(1.0)
    step = TrainStep(model, lr=1.0, weight_decay=0.5, max_grad_norm=None)
    step.grad.zero_()
    dom = rec["domains"][0]
    step._apply(dom)  # m = v = g = 0: the Adam update is 0 / (0 + eps) = 0, only the decay acts
    torch.cuda.synchronize()
    touched = set(step.engine.active_param_names(dict(model.named_parameters()), step.engine.dims(1, cfg.T, cfg.S, True), dom, True))
    n_dec = n_keep = 0
    for k, p in model.named_parameters():
        if k not in touched:
            assert torch.all(p == 1.0), k  # other domains / unused tensors are not in this step's ranges
        elif "bias" in k or "layer_norm.weight" in k:
            assert torch.all(p == 1.0), k
            n_keep += 1
        else:
            assert torch.all(p == 0.5), k
            n_dec += 1
    assert n_dec > 20 and n_keep > 20


def test_optimizer_state_and_checkpoint_roundtrip(tmp_path):
    """save_pretrained after the parameters moved into the flat arena (train_multi.py:310-321), TrainStep.state_dict /
    load_state_dict resume: the resumed run continues bit for bit like the uninterrupted one (eager path)."""
    from hma_b200 import STMaskGIT
    from hma_b200.train import TrainStep

    rec, cfg, sd = golden()
    batches = _batches(rec)
    kw = dict(lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
    a_model = build_cuda_model(rec, sd)
    a = TrainStep(a_model, **kw)
    for it in range(3):
        a(*batches[it % 2])
    a_model.save_pretrained(tmp_path / "ckpt")
    state = a.state_dict()
    b_model = STMaskGIT.from_pretrained(tmp_path / "ckpt").cuda()
    for (k, p), q in zip(a_model.named_parameters(), b_model.parameters()):
        assert torch.equal(p.detach(), q.detach()), k
    b = TrainStep(b_model, **kw)
    b.load_state_dict(state)
    assert b.step_count == 3 and sorted(b.dom_steps.values()) == sorted(a.dom_steps.values())
    for it in range(3, 5):
        la = a(*batches[it % 2])[0].item()
        lb = b(*batches[it % 2])[0].item()
        assert abs(la - lb) <= 1e-3 * abs(la)  # weight-gradient atomics: run-to-run rounding only
    for (k, p), q in zip(a_model.named_parameters(), b_model.parameters()):
        moved = (p.detach() - sd[k].cuda()).abs().mean().item()
        assert (p.detach() - q.detach()).abs().mean().item() <= 0.1 * moved + 1e-8, k


def test_train_step_matches_autograd_adamw_and_graph_replay_matches_eager():
    from hma_b200.train import TrainStep

    rec, cfg, sd = golden()
    ref_model = build_cuda_model(rec, sd)
    eager_model = build_cuda_model(rec, sd)
    graph_model = build_cuda_model(rec, sd)
    kw = dict(lr=1e-4, weight_decay=0.05, max_grad_norm=1.0)
    opt = torch.optim.AdamW(reference_param_groups(ref_model, kw["weight_decay"]), lr=kw["lr"], betas=(0.9, 0.999), eps=1e-8)
    eager = TrainStep(eager_model, **kw)
    graph = TrainStep(graph_model, cuda_graphs=True, **kw)
    batches = _batches(rec)
    losses = {"ref": [], "eager": [], "graph": []}
    for it in range(6):  # each domain three times: graph path = eager warm-up, capture, replay
        ids, labels, acts, dom = batches[it % len(batches)]
        opt.zero_grad(set_to_none=True)
        out = ref_model(ids, labels, action_ids=acts, domain=dom)
        out.loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in ref_model.parameters() if p.grad is not None], kw["max_grad_norm"])
        opt.step()
        losses["ref"].append(out.loss.item())
        losses["eager"].append(eager(ids, labels, acts, dom)[0].item())
        losses["graph"].append(graph(ids, labels, acts, dom)[0].item())
    assert len(graph._graphs) == len(batches)
    # graph replay runs the very same kernels on the very same data; the weight-gradient kernels accumulate with
    # fp32 atomics whose order varies from run to run, so the two trajectories agree to rounding, not bit for bit
    for a, b in zip(losses["graph"], losses["eager"]):
        assert abs(a - b) <= 1e-3 * abs(a), (losses["graph"], losses["eager"])
    for (k, a), b in zip(eager_model.named_parameters(), graph_model.parameters()):
        moved = (a.detach() - sd[k].cuda()).abs().mean().item()  # (Adam turns a sign flip of a ~0 gradient into 2*lr)
        assert (a.detach() - b.detach()).abs().mean().item() <= 0.1 * moved + 1e-8, k
    # and TrainStep follows autograd + AdamW + clip: same loss trajectory, parameters within bf16-gradient noise
    for a, b in zip(losses["ref"], losses["eager"]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses["ref"], losses["eager"])
    worst = 0.0
    for (k, a), b in zip(ref_model.named_parameters(), eager_model.parameters()):
        if a.grad is None:
            continue
        moved = (a.detach() - sd[k].cuda()).abs().mean().item()
        diff = (a.detach() - b.detach()).abs().mean().item()
        worst = max(worst, diff / max(moved, 1e-12))
        assert diff <= 0.25 * moved + 1e-8, (k, diff, moved)
    print("worst parameter deviation relative to the distance moved:", worst)
