"""The token -> pixel decode oracle (oracle/magvit_decoder_oracle.py) against the fixture written from the REAL reference
classes (oracle/make_decoder_golden.py: external/magvit2 Decoder + LFQ.get_codebook_entry), and known answers of the code
lookup."""
from pathlib import Path

import torch

from oracle import magvit_decoder_oracle as D

GOLDEN = Path(__file__).parent / "golden" / "magvit_decoder.pt"


def test_decoder_oracle_matches_reference_fixture():
    rec = torch.load(GOLDEN, weights_only=False)
    cfg = D.DecoderConfig()
    sd = D.make_state_dict(cfg, seed=rec["seed"])
    assert torch.equal(D.codebook_entry(rec["tokens"]), rec["quant"])
    u8, img = D.decode_tokens(rec["tokens"], sd, cfg)
    assert torch.allclose(img, rec["img"], rtol=1e-4, atol=1e-5)
    assert torch.equal(u8, rec["u8"])


def test_codebook_entry_known_answers():
    # id 0 -> all -1; id 2^18 - 1 -> all +1; id 1 (lowest bit): big-endian bit 17, flipped to channel 0
    t = torch.tensor([[[0, 262143, 1, 1 << 17]]])
    q = D.codebook_entry(t)
    assert q.shape == (1, 18, 1, 4)
    assert (q[0, :, 0, 0] == -1).all() and (q[0, :, 0, 1] == 1).all()
    assert q[0, 0, 0, 2] == 1 and (q[0, 1:, 0, 2] == -1).all()
    assert q[0, 17, 0, 3] == 1 and (q[0, :17, 0, 3] == -1).all()


def test_depth_to_space_is_dcr():
    x = torch.arange(2 * 8 * 3 * 3, dtype=torch.float32).reshape(2, 8, 3, 3)
    y = D.depth_to_space(x, 2)
    assert y.shape == (2, 2, 6, 6)
    for i in range(2):
        for j in range(2):
            assert torch.equal(y[:, :, i::2, j::2], x[:, (i * 2 + j) * 2:(i * 2 + j) * 2 + 2])
