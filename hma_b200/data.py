"""On-device MaskGIT training collator: drop-in for `get_maskgit_collator(config)` of the reference (hma/data.py:28-98).

Same call (`collate_fn(features) -> dict` with input_ids, labels, action_ids, domain, h, w), same corruption / masking
distribution, same consumption order of Python's `random`; the tensor draws come from torch's generator ON THE DEVICE
(the reference draws them on the device of `input_ids`, i.e. the CPU in its dataloader workers) and the arithmetic is one
CUDA kernel (csrc/collate.cu). `collate_from_draws` is the deterministic half: given the draws it is bit-exact against
the reference (tests/test_collator.py), which is how parity is checked.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional

import torch

from . import ops


def draw_on_device(cfg, B: int, h: int, w: int, device) -> Dict[str, object]:
    """The random draws of one collate call (data.py:43-83) as device tensors."""
    T, nv, vs = cfg.T, cfg.num_factored_vocabs, cfg.factored_vocab_size
    d: Dict[str, object] = {"first_masked_frame": 1}
    if cfg.dataloader_apply_corruption:
        d["corrupt_r"] = torch.rand(B, T, h, w, nv, device=device)
        d["u01"] = torch.rand((), device=device)
        d["rand_vals"] = torch.randint(low=0, high=vs, size=(B, T, h, w, nv), dtype=torch.long, device=device)
    if random.random() < cfg.non_mlm_ratio:
        fmf = random.randint(cfg.num_prompt_frames, T - 1)
        d["first_masked_frame"] = fmf
        rate = random.uniform(cfg.dataloader_mask_ratio_min, 1.0)
        rates = []
        for _ in range(T - fmf):
            rate *= random.uniform(0.9, 1.0)
            rates.append(rate)
        d["frame_rates"] = torch.tensor(rates, dtype=torch.float64)
        d["frame_r"] = torch.rand(B, T - fmf, h, w, nv, device=device)
    if cfg.dataloader_apply_mask:
        fmf = d["first_masked_frame"]
        while True:  # "we could get unlucky and mask no tokens" (data.py:72)
            prob = torch.cos(torch.rand(B, T - fmf, 1, 1, device=device) * (math.pi / 2))
            r = torch.rand(B, T - fmf, h, w, device=device)
            if bool((r < prob).any()):
                break
        d["mask_prob"], d["mask_r"] = prob, r
    return d


def collate_from_draws(tokens: torch.Tensor, d: Dict[str, object], cfg, h: int, w: int):
    """tokens: i64 [B, T*h*w] on the device; draws as produced by draw_on_device (or the oracle's draw(), moved to the
    device). Returns (input_ids, labels), i64 [B, T*h*w]."""
    if not tokens.is_cuda:
        raise RuntimeError("hma_b200.data collates on a CUDA device only (no CPU path exists)")
    dev = tokens.device
    if not cfg.dataloader_apply_mask:
        # data.py:69-83: the corrupted factors are only folded back into token ids (unfactorize_token_ids) inside
        # `if config.dataloader_apply_mask`; without it the reference returns the ORIGINAL tokens as input_ids — the
        # corruption draws are consumed and then discarded. Mirrored.
        return tokens.clone(), tokens.clone()

    def f32(x):
        return None if x is None else x.to(device=dev, dtype=torch.float32).contiguous()

    u01 = d.get("u01")
    thresh = float(cfg.max_corrupt_rate * (u01.item() if u01 is not None else 0.0))
    if u01 is not None:  # the reference compares float32 < (Python float * 0-d float32 tensor) = a float32 scalar
        thresh = float((cfg.max_corrupt_rate * u01.to(torch.float32).cpu()).item())
    rand_vals = d.get("rand_vals")
    if "frame_r" in d and rand_vals is None:
        raise NameError("the non-MLM branch uses random_values, which only exists when dataloader_apply_corruption is set "
                        "(data.py:46,62: the reference raises NameError here too)")
    frame_rates = d.get("frame_rates")
    return ops.collate_maskgit(tokens.contiguous(), tokens.shape[0], cfg.T, h * w, cfg.num_factored_vocabs, cfg.factored_vocab_size,
                               cfg.image_vocab_size, f32(d.get("corrupt_r")), thresh,
                               None if rand_vals is None else rand_vals.to(dev).contiguous(), int(d["first_masked_frame"]),
                               f32(frame_rates), f32(d.get("frame_r")),
                               f32(d.get("mask_prob")) if cfg.dataloader_apply_mask else None,
                               f32(d.get("mask_r")) if cfg.dataloader_apply_mask else None)


def get_maskgit_collator(config, device="cuda"):
    """Same contract as the reference's get_maskgit_collator(config); batches are assembled and corrupted on `device`."""
    dev = torch.device(device)

    def collate_fn(features: List[dict]) -> Dict[str, object]:
        h, w = features[0]["h"], features[0]["w"]
        tokens = torch.stack([ex["input_ids"] for ex in features]).to(dev, non_blocking=True)
        draws = draw_on_device(config, len(features), h, w, dev)
        input_ids, labels = collate_from_draws(tokens, draws, config, h, w)
        out: Dict[str, object] = {"input_ids": input_ids, "labels": labels}
        if "action_ids" in features[0]:
            out["action_ids"] = torch.stack([ex["action_ids"] for ex in features]).to(dev, non_blocking=True)
        out["domain"] = [ex["domain"] for ex in features]
        out["h"] = [ex["h"] for ex in features]
        out["w"] = [ex["w"] for ex in features]
        return out

    return collate_fn


def cosine_schedule(u: torch.Tensor) -> torch.Tensor:
    """hma/data.py cosine_schedule (st_mask_git.py:116-125 on tensors): cos(pi/2 * u)."""
    return torch.cos(u * math.pi / 2)  # same association as the reference: (u * pi) / 2


def get_maskgit_collator_feature(config, device=None):
    """Drop-in for the reference's continuous-latent collator (hma/data.py:100-157), which STMAR training uses: the latents
    pass through unchanged (input_ids, labels = a copy) and `masked_tokens_indicator` [B, T, h, w] marks, for the frames
    from `first_masked_frame` on, the positions drawn with a per-(sample, frame) cosine-schedule rate — STMAR.forward puts
    its mask token there (st_mar.py:240). Same consumption of Python's `random` and of torch's default generator as the
    reference (so the same seeds give the same masks); with `device` set the batch is moved there first and the tensor
    draws happen on that device."""
    def collate_fn(features: List[dict]) -> Dict[str, object]:
        h, w = features[0]["h"], features[0]["w"]
        B, T = len(features), config.T
        input_ids = torch.stack([ex["input_ids"] for ex in features])
        if device is not None:
            input_ids = input_ids.to(device, non_blocking=True)
        dev = input_ids.device
        x = input_ids.reshape(B, T, h, w, -1)
        first_masked_frame = T
        mask = torch.zeros(1).long()
        indicator = torch.zeros((B, T, h, w)).long()
        if config.dataloader_apply_mask:
            if random.random() < config.non_mlm_ratio:
                first_masked_frame = random.randint(config.num_prompt_frames, T - 1)
            else:
                first_masked_frame = 1
            while mask.max() == 0:  # "we could get unlucky and mask no tokens"
                rand = torch.rand(B, T - first_masked_frame, 1, 1, device=dev if device is not None else None)
                rate = cosine_schedule(rand * (1 - config.dataloader_mask_ratio_min) + config.dataloader_mask_ratio_min)
                r = torch.rand_like(x[:, first_masked_frame:, ..., 0], dtype=torch.float)
                mask = r < rate.to(r.device)
            indicator = torch.cat([torch.zeros((B, first_masked_frame, h, w), dtype=mask.dtype, device=mask.device), mask], dim=1)
        out: Dict[str, object] = {"input_ids": x.reshape(B, T * h * w, -1), "labels": x.clone().reshape(B, T * h * w, -1),
                                  "masked_tokens_indicator": indicator}
        if "action_ids" in features[0]:
            a = torch.stack([ex["action_ids"] for ex in features])
            out["action_ids"] = a.to(device, non_blocking=True) if device is not None else a
        out["domain"] = [ex["domain"] for ex in features]
        out["h"] = [ex["h"] for ex in features]
        out["w"] = [ex["w"] for ex in features]
        return out

    return collate_fn
