"""Oracle for the token -> pixel decode: LFQ code lookup + the MagViT2 convolutional decoder + the
[-1, 1] -> uint8 mapping, as plain fp32 PyTorch.

TEST INFRASTRUCTURE ONLY (same rules as stmaskgit_oracle.py). Restates, functionally, with weights in the reference's
state_dict layout (the `decoder.*` keys of `VQModel`):
  * hma/visualize.py:136-151  decode_latents: get_codebook_entry(tokens).flip(1) -> model.decode -> unnormalize_imgs
  * external/magvit2/modules/vqvae/lookup_free_quantize.py:181-194  LFQ.get_codebook_entry (token_factorization=False)
  * external/magvit2/modules/diffusionmodules/improved_model.py:12-51,124-183,185-234  ResBlock, Decoder, depth_to_space, Upsampler
  * external/magvit2/models/lfqgan.py:131-133  VQModel.decode = decoder(quant)
  * hma/visualize.py:112-121  unnormalize_imgs
Pinned on the reference's own classes run here (oracle/make_decoder_golden.py -> tests/golden/magvit_decoder.pt).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class DecoderConfig:
    """The fields of external/magvit2/config.py:VQConfig the decode path reads (same defaults)."""

    z_channels: int = 18
    out_channels: int = 3
    base_channels: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    num_res_blocks: int = 2
    codebook_size: int = 262144


def codebook_entry(tokens_BHW: Tensor, codebook_dim: int = 18) -> Tensor:
    """lookup_free_quantize.py:181-194 followed by visualize.py:150's `.flip(1)`: big-endian bits of the id as +-1 channels,
    then the channel axis reversed. Returns fp32 [B, codebook_dim, H, W]."""
    B, H, W = tokens_BHW.shape
    mask = 2 ** torch.arange(codebook_dim - 1, -1, -1, device=tokens_BHW.device, dtype=torch.long)
    x = (tokens_BHW.reshape(B, H * W).unsqueeze(-1) & mask) != 0
    x = x * 2.0 - 1.0
    x = x.reshape(B, H, W, codebook_dim).permute(0, 3, 1, 2)
    return x.flip(1).float()


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def res_block(x: Tensor, sd: SD, p: str) -> Tensor:
    """improved_model.py:36-51 (GroupNorm(32, eps=1e-6), bias-free 3x3 convolutions, 1x1 `nin_shortcut` when widths differ)."""
    h = F.group_norm(x, 32, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    h = F.conv2d(swish(h), sd[p + "conv1.weight"], None, padding=1)
    h = F.group_norm(h, 32, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.conv2d(swish(h), sd[p + "conv2.weight"], None, padding=1)
    if p + "nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], None)
    elif p + "conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "conv_shortcut.weight"], None, padding=1)
    return h + x


def depth_to_space(x: Tensor, bs: int = 2) -> Tensor:
    """improved_model.py:185-217 (DCR order: channel = (i * bs + j) * C' + c)."""
    B, C, H, W = x.shape
    x = x.view(B, bs, bs, C // (bs * bs), H, W).permute(0, 3, 4, 1, 5, 2)
    return x.contiguous().view(B, C // (bs * bs), H * bs, W * bs)


def decoder(z: Tensor, sd: SD, cfg: DecoderConfig, prefix: str = "") -> Tensor:
    """improved_model.py:162-183. z: [B, z_channels, h, w] -> [B, out_channels, h * 2^(levels-1), w * 2^(levels-1)]."""
    p = prefix
    x = F.conv2d(z, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    for r in range(cfg.num_res_blocks):
        x = res_block(x, sd, p + f"mid_block.{r}.")
    for lvl in reversed(range(len(cfg.ch_mult))):
        for r in range(cfg.num_res_blocks):
            x = res_block(x, sd, p + f"up.{lvl}.block.{r}.")
        if lvl > 0:
            q = p + f"up.{lvl}.upsample.conv1."
            x = depth_to_space(F.conv2d(x, sd[q + "weight"], sd[q + "bias"], padding=1), 2)
    x = F.group_norm(x, 32, sd[p + "norm_out.weight"], sd[p + "norm_out.bias"], 1e-6)
    return F.conv2d(swish(x), sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)


def unnormalize_imgs(x: Tensor) -> Tensor:
    """visualize.py:112-121: clamp to [-1, 1], (x + 1) * 127.5, clamp to [0, 255], truncate to uint8."""
    x = torch.clamp(x, -1, 1)
    return torch.clamp((x.detach().cpu() + 1) * 127.5, 0, 255).to(dtype=torch.uint8)


def decode_tokens(tokens_BHW: Tensor, sd: SD, cfg: DecoderConfig, prefix: str = ""):
    """visualize.py:147-158 for a VQModel. Returns (uint8 [B, 3, H', W'], fp32 decoder output)."""
    img = decoder(codebook_entry(tokens_BHW, cfg.z_channels), sd, cfg, prefix)
    return unnormalize_imgs(img), img


def make_state_dict(cfg: DecoderConfig, seed: int = 0) -> SD:
    """Deterministic non-degenerate weights in the reference Decoder's key layout (fan-in scaled so that activations stay O(1)
    through the 19 convolutions; norm gains around 1)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def conv(name, cout, cin, k, bias):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5
        if bias:
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def norm(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def resblock(name, cin, cout):
        norm(name + "norm1", cin)
        norm(name + "norm2", cout)
        conv(name + "conv1", cout, cin, 3, False)
        conv(name + "conv2", cout, cout, 3, False)
        if cin != cout:
            conv(name + "nin_shortcut", cout, cin, 1, False)

    nb = len(cfg.ch_mult)
    block_in = cfg.base_channels * cfg.ch_mult[nb - 1]
    conv("conv_in", block_in, cfg.z_channels, 3, True)
    for r in range(cfg.num_res_blocks):
        resblock(f"mid_block.{r}.", block_in, block_in)
    for lvl in reversed(range(nb)):
        block_out = cfg.base_channels * cfg.ch_mult[lvl]
        for r in range(cfg.num_res_blocks):
            resblock(f"up.{lvl}.block.{r}.", block_in, block_out)
            block_in = block_out
        if lvl > 0:
            conv(f"up.{lvl}.upsample.conv1", block_in * 4, block_in, 3, True)
    norm("norm_out", block_in)
    conv("conv_out", cfg.out_channels, block_in, 3, True)
    return sd
