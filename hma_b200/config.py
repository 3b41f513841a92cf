"""GenieConfig: the reference's model configuration (hma/config.py:8-81) with the same field names and defaults, so that a
reference `config.json` loads verbatim and `GenieConfig(**kwargs)` calls keep working. Fields the CUDA path does not
implement are accepted here and rejected, loudly, by hma_b200.engine.check_config. The fields are grouped by what consumes
them; only the first five keep the reference's positional order (num_layers, num_heads, d_model, T, S)."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Optional


def nth_root(x: int, n: int) -> int:
    """Integer n-th root that must be exact (factorization_utils.py:98-102): 262144 -> 512 for two factored vocabularies."""
    root = round(x ** (1 / n))
    assert root ** n == x, (x, n, root)
    return root


@dataclass
class GenieConfig:
    # ---- transformer trunk (st_transformer.py)
    num_layers: int
    num_heads: int
    d_model: int
    T: int = 12                       # frames per window
    S: int = 256                      # tokens per frame (16 x 16)
    mlp_ratio: float = 4.0
    mlp_bias: bool = True
    mlp_drop: float = 0.0
    qkv_bias: bool = False
    proj_bias: bool = True
    qk_norm: bool = True
    attn_drop: float = 0.0            # constructed but never applied by the reference (attention.py:29)
    use_mup: bool = False
    arch: str = "STTransformerDecoder"

    # ---- vocabulary (factorization_utils.py)
    image_vocab_size: Optional[int] = 262144
    num_factored_vocabs: int = 1
    factored_vocab_size: Optional[int] = None   # derived in __post_init__

    # ---- action conditioning (st_mask_git.py:201-251)
    use_actions: bool = True
    init_actions: bool = False
    action_network: str = "mlp"
    action_token_size: int = 64
    action_domains: Optional[List[str]] = None
    d_actions: Optional[List[int]] = None
    action_stats: Optional[list] = None
    d_action: int = 28
    shared_action_mlps: bool = True
    random_dummy_action: bool = True
    drop_action_ratio: float = 0.0
    jointly_predict_actions: bool = False
    jointly_predict_states: bool = True
    action_contrastive_loss: bool = False
    action_loss_weight: float = 0.5
    label_drop_prob: float = 0.5

    # ---- training collator (data.py:28-98)
    dataloader_apply_mask: bool = True
    dataloader_apply_corruption: bool = True
    dataloader_mask_ratio_min: float = 0.2
    max_corrupt_rate: float = 0.2
    non_mlm_ratio: float = 0.2
    num_prompt_frames: int = 4

    def __post_init__(self):
        # config.py:77-81: one vocabulary of image_vocab_size entries factored into num_factored_vocabs digits
        self.factored_vocab_size = 64 if self.image_vocab_size is None else nth_root(self.image_vocab_size, self.num_factored_vocabs)

    def save_pretrained(self, json_path):
        with open(json_path, "w") as f:
            json.dump(vars(self), f)

    @classmethod
    def from_pretrained(cls, json_path):
        with open(json_path, "r") as f:
            return cls(**json.load(f))

    def shallow_copy(self):
        return GenieConfig(**vars(self))
