"""Pin oracle/magvit_decoder_oracle.py on the REAL reference classes (external/magvit2 Decoder, LFQ.get_codebook_entry,
hma/visualize.unnormalize_imgs arithmetic) and write tests/golden/magvit_decoder.pt.

    python -m oracle.make_decoder_golden

TEST INFRASTRUCTURE ONLY. The checkpoint (data/magvit2.ckpt) is not available offline, so the weights are the deterministic
random ones of magvit_decoder_oracle.make_state_dict, loaded into the reference Decoder with strict=True."""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from oracle import magvit_decoder_oracle as D  # noqa: E402


def reference_decoder_and_lfq():
    from external.magvit2.config import VQConfig
    from external.magvit2.modules.diffusionmodules.improved_model import Decoder
    from external.magvit2.modules.vqvae.lookup_free_quantize import LFQ
    return VQConfig, Decoder, LFQ


def main():
    VQConfig, Decoder, LFQ = reference_decoder_and_lfq()
    cfg = D.DecoderConfig()
    sd = D.make_state_dict(cfg, seed=0)
    ref = Decoder(VQConfig())
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(0, 262144, (2, 4, 4), generator=g)
    # LFQ.get_codebook_entry as an unbound method on a stub carrying the two attributes it reads
    stub = types.SimpleNamespace(token_factorization=False, codebook_dim=18)
    quant = LFQ.get_codebook_entry(stub, tokens.reshape(2, 16), bhwc=(2, 4, 4, 18)).flip(1)  # visualize.py:149-150
    with torch.no_grad():
        img = ref(quant.float())
    u8 = torch.clamp((torch.clamp(img, -1, 1).detach().cpu() + 1) * 127.5, 0, 255).to(dtype=torch.uint8)  # visualize.py:112-121
    out = {"tokens": tokens, "quant": quant.float(), "img": img, "u8": u8, "seed": 0}
    # the restatement must agree before anything is written
    u8_o, img_o = D.decode_tokens(tokens, sd, cfg)
    assert torch.equal(D.codebook_entry(tokens), quant.float())
    assert torch.allclose(img_o, img, rtol=1e-4, atol=1e-5), (img_o - img).abs().max()
    assert torch.equal(u8_o, u8)
    path = ROOT / "tests" / "golden" / "magvit_decoder.pt"
    torch.save(out, path)
    print("wrote", path, tuple(img.shape), float(img.abs().max()), float(img.std()))


if __name__ == "__main__":
    main()
