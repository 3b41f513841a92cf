"""CPU restatement of the reference's MaskGIT training collator (hma/data.py:28-98, `get_maskgit_collator`).

TEST INFRASTRUCTURE ONLY (used by tests/ to check hma_b200.data): never imported by the product path.

The reference interleaves its random draws with the arithmetic. Here the two are separated:
`draw()` consumes the torch CPU generator and Python's `random` in EXACTLY the reference's order and
returns the draws as tensors; `apply()` is the deterministic integer arithmetic on them. With the same
seeds, `apply(tokens, draw(...))` reproduces the reference collator bit for bit
(tests/test_collator.py pins that against the live reference when /root/reference is present and against
tests/golden/collator_*.pt otherwise).
"""
import math
import random

import torch


def cosine_schedule_t(u: torch.Tensor) -> torch.Tensor:  # st_mask_git.py:116-125
    return torch.cos(u * math.pi / 2)


def draw(cfg, B: int, h: int, w: int):
    """Random draws of one collate call, in the reference's order (data.py:43-83). Uses the global torch CPU
    generator and the global `random` state, like the reference."""
    T, nv, vs = cfg.T, cfg.num_factored_vocabs, cfg.factored_vocab_size
    d = {"first_masked_frame": 1}
    if cfg.dataloader_apply_corruption:
        d["corrupt_r"] = torch.rand(B, T, h, w, nv)                                   # :43
        d["u01"] = torch.rand(())                                                     # :44
        d["rand_vals"] = torch.randint(low=0, high=vs, size=(B, T, h, w, nv), dtype=torch.long)  # :46-47
    if random.random() < cfg.non_mlm_ratio:                                           # :50
        fmf = random.randint(cfg.num_prompt_frames, T - 1)                            # :53
        d["first_masked_frame"] = fmf
        rate = random.uniform(cfg.dataloader_mask_ratio_min, 1.0)                     # :58
        rates, rs = [], []
        for _ in range(T - fmf):
            rate *= random.uniform(0.9, 1.0)                                          # :60
            rates.append(rate)
            rs.append(torch.rand((B, h, w, nv)))                                      # :61
        d["frame_rates"] = torch.tensor(rates, dtype=torch.float64)
        d["frame_r"] = torch.stack(rs, dim=1)                                         # [B, T-fmf, h, w, nv]
    if cfg.dataloader_apply_mask:
        fmf = d["first_masked_frame"]
        tries = []
        while True:                                                                   # :72-78
            prob = cosine_schedule_t(torch.rand(B, T - fmf, 1, 1))
            r = torch.rand(B, T - fmf, h, w)
            tries.append((prob, r))
            if (r < prob).max() != 0:
                break
        d["mask_prob"], d["mask_r"] = tries[-1]
        d["mask_tries"] = len(tries)
    return d


def apply(tokens: torch.Tensor, d: dict, cfg, h: int, w: int):
    """tokens: i64 [B, T*h*w]. Returns (input_ids, labels) exactly as data.py:35-88 computes them from these draws."""
    B = tokens.shape[0]
    T, nv, vs = cfg.T, cfg.num_factored_vocabs, cfg.factored_vocab_size
    x = tokens.reshape(B, T, h, w)
    labels = x.clone()
    powers = vs ** torch.arange(nv)
    f = (x.unsqueeze(-1) // powers) % vs                                              # factorize_token_ids
    if cfg.dataloader_apply_corruption:
        m = d["corrupt_r"] < cfg.max_corrupt_rate * d["u01"]
        f[m] = d["rand_vals"][m]
    fmf = d["first_masked_frame"]
    if "frame_r" in d:
        view = f[:, fmf:]
        for i in range(view.size(1)):
            # the reference compares a float32 tensor with a Python float: the scalar is rounded to float32 first
            m = d["frame_r"][:, i] > torch.tensor(d["frame_rates"][i].item(), dtype=torch.float32)
            view[:, i][m] = d["rand_vals"][:, fmf + i][m]
    out = (f * powers).sum(-1)                                                        # unfactorize_token_ids
    if cfg.dataloader_apply_mask:
        mask = d["mask_r"] < d["mask_prob"]
        out[:, fmf:][mask] = cfg.image_vocab_size
    else:
        # data.py:69-83: x_THWC is folded back into x_THW (unfactorize_token_ids) only inside `if dataloader_apply_mask`;
        # otherwise input_ids are the original tokens and the corruption is discarded (checked against the live reference)
        out = x
    return out.reshape(B, -1), labels.reshape(B, -1)
