"""GenieConfig: field-for-field mirror of the reference dataclass (hma/config.py:8-81) so that a
reference `config.json` loads verbatim. Fields the CUDA path does not implement are accepted here
and rejected, loudly, by hma_b200.engine.check_config."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Optional


def nth_root(x: int, n: int) -> int:
    """factorization_utils.py:98-102"""
    root = round(x ** (1 / n))
    assert root ** n == x, (x, n, root)
    return root


@dataclass
class GenieConfig:
    num_layers: int
    num_heads: int
    d_model: int
    T: int = 12
    S: int = 256
    image_vocab_size: Optional[int] = 262144
    use_mup: bool = False
    dataloader_apply_mask: bool = True
    dataloader_apply_corruption: bool = True
    dataloader_mask_ratio_min: float = 0.2
    drop_action_ratio: float = 0.0
    arch: str = "STTransformerDecoder"
    random_dummy_action: bool = True

    num_factored_vocabs: int = 1
    factored_vocab_size: Optional[int] = None

    max_corrupt_rate: float = 0.2
    non_mlm_ratio: float = 0.2
    num_prompt_frames: int = 4

    init_actions: bool = False
    d_action: int = 28
    use_actions: bool = True
    action_domains: Optional[List[str]] = None
    d_actions: Optional[List[int]] = None
    action_stats: Optional[list] = None
    action_network: str = "mlp"
    shared_action_mlps: bool = True
    action_contrastive_loss: bool = False
    jointly_predict_actions: bool = False
    jointly_predict_states: bool = True
    action_token_size: int = 64
    label_drop_prob: float = 0.5
    action_loss_weight: float = 0.5

    qkv_bias: bool = False
    proj_bias: bool = True
    attn_drop: float = 0.0
    qk_norm: bool = True

    mlp_ratio: float = 4.0
    mlp_drop: float = 0.0
    mlp_bias: bool = True

    def save_pretrained(self, json_path):
        with open(json_path, "w") as f:
            json.dump(vars(self), f)

    @classmethod
    def from_pretrained(cls, json_path):
        with open(json_path, "r") as f:
            config = json.load(f)
        return cls(**config)

    def shallow_copy(self):
        return GenieConfig(**vars(self))

    def __post_init__(self):
        if self.image_vocab_size is None:
            self.factored_vocab_size = 64
        else:
            self.factored_vocab_size = nth_root(self.image_vocab_size, self.num_factored_vocabs)
