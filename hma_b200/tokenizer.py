"""Token -> pixel decode on the B200 CUDA path: the MagViT2 decoder the reference uses to turn
generated token grids into frames (hma/visualize.py:124-169 `decode_latents_wrapper`, sim/simulator.py's display path).

`MagVitDecoder` mirrors external/magvit2/modules/diffusionmodules/improved_model.py:124-183 — same constructor argument
(a VQConfig-like object), same parameter names and shapes, so the `decoder.*` entries of a reference `VQModel` checkpoint
load with strict=True — and adds `decode_tokens`, the body of `decode_latents` for a VQModel (LFQ code lookup, channel
flip, decode, [-1, 1] -> uint8). The sub-modules are parameter containers; every FLOP runs in libhma_b200.so:
3x3 convolutions are tcgen05 contractions over zero-bordered NHWC images whose nine taps are nine TMA row offsets of the
same matrix (hma_conv3x3_nhwc — no im2col), GroupNorm + swish, depth-to-space and the uint8 mapping are streaming
kernels (csrc/vqdecode.cu). There is no CPU path.
"""
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

BF16, F32 = torch.bfloat16, torch.float32


@dataclass
class VQConfig:
    """external/magvit2/config.py:11-21 — the fields the decode path reads, same names and defaults."""

    in_channels: int = 3
    z_channels: int = 18
    out_channels: int = 3
    base_channels: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    num_res_blocks: int = 2
    num_codebooks: int = 1
    codebook_size: int = 262144
    token_factorization: bool = False


class _ResBlock(nn.Module):  # improved_model.py:12-34
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.in_filters, self.out_filters = cin, cout
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1, padding=0, bias=False)


class _Upsampler(nn.Module):  # improved_model.py:220-229
    def __init__(self, dim: int):
        super().__init__()
        self.conv1 = nn.Conv2d(dim, dim * 4, 3, padding=1)


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class MagVitDecoder(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        config = config if config is not None else VQConfig()
        self.config = config
        if config.base_channels % 128 != 0 or getattr(config, "token_factorization", False) or getattr(config, "num_codebooks", 1) != 1:
            raise NotImplementedError("MagVitDecoder: base_channels must be a multiple of 128, one codebook, no token factorisation")
        nb = len(config.ch_mult)
        self.num_blocks, self.num_res_blocks = nb, config.num_res_blocks
        block_in = config.base_channels * config.ch_mult[nb - 1]
        self.conv_in = nn.Conv2d(config.z_channels, block_in, 3, padding=1, bias=True)
        self.mid_block = nn.ModuleList([_ResBlock(block_in, block_in) for _ in range(config.num_res_blocks)])
        self.up = nn.ModuleList()
        for lvl in reversed(range(nb)):
            block_out = config.base_channels * config.ch_mult[lvl]
            blocks = nn.ModuleList()
            for _ in range(config.num_res_blocks):
                blocks.append(_ResBlock(block_in, block_out))
                block_in = block_out
            up = nn.Module()
            up.block = blocks
            if lvl > 0:
                up.upsample = _Upsampler(block_in)
            self.up.insert(0, up)
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6)
        self.conv_out = nn.Conv2d(block_in, config.out_channels, 3, padding=1)
        self._w: Dict[str, tuple] = {}

    # ------------------------------------------------------------------ operand preparation (once per weight version)
    def _conv_operand(self, name: str, conv: nn.Conv2d):
        """bf16 [Cout_p, k*k*Cin_p] with k index = (ky*3+kx)*Cin_p + ci (zero-padded: Cin to 64, Cout to 128) + fp32 bias."""
        w = conv.weight
        ver = (w._version, w.data_ptr(), None if conv.bias is None else conv.bias._version)
        hit = self._w.get(name)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        cout, cin, kh, kw = w.shape
        cin_p, cout_p = _pad_to(cin, 64), _pad_to(cout, 128)
        wt = torch.zeros(cout_p, kh, kw, cin_p, device=w.device, dtype=F32)
        wt[:cout, :, :, :cin] = w.detach().permute(0, 2, 3, 1)
        wt = wt.reshape(cout_p, kh * kw * cin_p).to(BF16).contiguous()
        bias = None
        if conv.bias is not None:
            bias = torch.zeros(cout_p, device=w.device, dtype=F32)
            bias[:cout] = conv.bias.detach()
        self._w[name] = (ver, wt, bias)
        return wt, bias

    # ------------------------------------------------------------------ stages (all on zero-bordered NHWC matrices)
    @staticmethod
    def _conv3(x16: torch.Tensor, wt: torch.Tensor, bias, resid, W: int) -> torch.Tensor:
        rows, cin = x16.shape
        cout = wt.shape[0]
        out = torch.empty(rows, cout, device=x16.device, dtype=F32)
        ops._call(f"conv3x3[{cin}->{cout}]", 2.0 * rows * cout * 9 * cin, "hma_conv3x3_nhwc", x16.data_ptr(), x16.stride(0), wt.data_ptr(),
                  wt.stride(0), rows, cin, cout, W + 2, out.data_ptr(), out.stride(0), ops._p(bias), ops._p(resid),
                  resid.stride(0) if resid is not None else 0, ops._s())
        return out

    @staticmethod
    def _gn_swish(x32: torch.Tensor, norm: Optional[nn.GroupNorm], images: int, H: int, W: int) -> torch.Tensor:
        C = x32.shape[1]
        out = torch.empty(x32.shape, device=x32.device, dtype=BF16)
        if norm is None:  # plain bf16 copy with a zero border
            ops._call("vq_cast", x32.numel() * 6.0, "hma_gn_swish", x32.data_ptr(), None, None, None, images, H, W, C, 0.0, 1, out.data_ptr(),
                      ops._s())
            return out
        sums = torch.empty(images, 32, 2, device=x32.device, dtype=F32)
        scratch = torch.empty(images, 64, 64, device=x32.device, dtype=F32)
        ops._call("gn_stats", x32.numel() * 4.0, "hma_gn_stats", x32.data_ptr(), images, H, W, C, scratch.data_ptr(), sums.data_ptr(),
                  ops._s())
        ops._call("gn_swish", x32.numel() * 6.0, "hma_gn_swish", x32.data_ptr(), sums.data_ptr(), norm.weight.data_ptr(),
                  norm.bias.data_ptr(), images, H, W, C, float(norm.eps), 0, out.data_ptr(), ops._s())
        return out

    def _res_block(self, x32: torch.Tensor, blk: _ResBlock, name: str, images: int, H: int, W: int) -> torch.Tensor:
        w1, _ = self._conv_operand(name + "conv1", blk.conv1)
        w2, _ = self._conv_operand(name + "conv2", blk.conv2)
        h = self._conv3(self._gn_swish(x32, blk.norm1, images, H, W), w1, None, None, W)
        h16 = self._gn_swish(h, blk.norm2, images, H, W)
        resid = x32
        if hasattr(blk, "nin_shortcut"):  # 1x1 convolution of the block input: a plain GEMM over the same rows
            wn, _ = self._conv_operand(name + "nin_shortcut", blk.nin_shortcut)
            resid = ops.gemm_nt(self._gn_swish(x32, None, images, H, W), wn, ops.EPI_RESID)
        return self._conv3(h16, w2, None, resid, W)  # x + residual rides on the convolution's epilogue

    def _decode_padded(self, z16: torch.Tensor, images: int, H: int, W: int):
        """z16: bf16 [images*(H+2)*(W+2), 64] zero-bordered code image. Returns (fp32 [images*(H'+2)*(W'+2), 128], H', W')."""
        if not z16.is_cuda:
            raise RuntimeError("hma_b200.MagVitDecoder runs on a CUDA device only (no CPU path exists)")
        wt, b = self._conv_operand("conv_in", self.conv_in)
        x = self._conv3(z16, wt, b, None, W)
        for r, blk in enumerate(self.mid_block):
            x = self._res_block(x, blk, f"mid_block.{r}.", images, H, W)
        for lvl in reversed(range(self.num_blocks)):
            up = self.up[lvl]
            for r, blk in enumerate(up.block):
                x = self._res_block(x, blk, f"up.{lvl}.block.{r}.", images, H, W)
            if lvl > 0:
                wt, b = self._conv_operand(f"up.{lvl}.upsample.conv1", up.upsample.conv1)
                y = self._conv3(self._gn_swish(x, None, images, H, W), wt, b, None, W)
                Co = y.shape[1] // 4
                x = torch.empty(images * (2 * H + 2) * (2 * W + 2), Co, device=y.device, dtype=F32)
                ops._call("depth_to_space", y.numel() * 8.0, "hma_depth_to_space", y.data_ptr(), images, H, W, Co, x.data_ptr(), ops._s())
                H, W = 2 * H, 2 * W
        wt, b = self._conv_operand("conv_out", self.conv_out)
        out = self._conv3(self._gn_swish(x, self.norm_out, images, H, W), wt, b, None, W)
        return out, H, W

    # ------------------------------------------------------------------ improved_model.py:162-183
    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """z: [B, z_channels, h, w] -> fp32 [B, out_channels, 16h, 16w] (the reference Decoder.forward contract)."""
        B, C, H, W = z.shape
        zp = torch.zeros(B, H + 2, W + 2, _pad_to(C, 64), device=z.device, dtype=BF16)
        zp[:, 1:-1, 1:-1, :C] = z.permute(0, 2, 3, 1)
        out, Ho, Wo = self._decode_padded(zp.view(-1, zp.shape[-1]), B, H, W)
        oc = self.config.out_channels
        return out.view(B, Ho + 2, Wo + 2, -1)[:, 1:-1, 1:-1, :oc].permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ visualize.py:136-151 (the VQModel branch)
    @torch.no_grad()
    def decode_tokens(self, tokens_BHW: torch.Tensor) -> torch.Tensor:
        """i64 token grids [B, h, w] -> uint8 frames [B, 3, 16h, 16w]: get_codebook_entry(...).flip(1) -> decode ->
        unnormalize_imgs, all on the device."""
        if not tokens_BHW.is_cuda:
            raise RuntimeError("hma_b200.MagVitDecoder runs on a CUDA device only (no CPU path exists)")
        B, H, W = tokens_BHW.shape
        bits = self.config.z_channels
        zc = _pad_to(bits, 64)
        z16 = torch.empty(B * (H + 2) * (W + 2), zc, device=tokens_BHW.device, dtype=BF16)
        ops._call("lfq_entry", z16.numel() * 2.0, "hma_lfq_entry", tokens_BHW.contiguous().data_ptr(), B, H, W, bits, zc, z16.data_ptr(),
                  ops._s())
        out, Ho, Wo = self._decode_padded(z16, B, H, W)
        img = torch.empty(B, self.config.out_channels, Ho, Wo, device=out.device, dtype=torch.uint8)
        ops._call("to_uint8", img.numel() * 5.0, "hma_to_uint8", out.data_ptr(), B, Ho, Wo, out.shape[1], self.config.out_channels,
                  img.data_ptr(), ops._s())
        return img


def decode_latents_wrapper(decoder: MagVitDecoder, batch_size: int = 16, max_images: Optional[int] = None):
    """Same contract as visualize.py:124-169's closure for quantised data: `decode_latents(video_data)` takes an integer array
    (b, h, w) and returns uint8 frames [b, 3, H, W] (a tensor on the host; the reference converts each to a PIL image)."""
    import math

    import numpy as np

    def decode_latents(video_data):
        dev = next(decoder.parameters()).device
        frames = []
        for i in range(math.ceil(len(video_data) / batch_size)):
            shard = video_data[i * batch_size: (i + 1) * batch_size]
            assert shard.ndim == 3, f"{shard.shape=} {shard.dtype=}"
            t = torch.from_numpy(np.asarray(shard).astype(np.int64)) if not torch.is_tensor(shard) else shard.long()
            frames.append(decoder.decode_tokens(t.to(dev)).cpu())
            if max_images and len(frames) * batch_size >= max_images:
                break
        return torch.cat(frames)

    return decode_latents
