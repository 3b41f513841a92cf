"""Import shim for `mup` (pinned by the reference as git+https://github.com/janEbert/mup.git@fsdp-fix,
requirements.txt:11; not vendored, not installable offline). Only what hma/model/st_mask_git.py
touches at import/forward time is provided. TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py
in the authoring container to run the real reference)."""
import torch
import torch.nn as nn


class MuReadout(nn.Linear):
    """Published muP readout: y = Linear(output_mult * x / width_mult). The reference hard-codes its
    base shape to d_model=256 (st_mask_git.py:755-760), so width_mult = in_features / 256."""

    def __init__(self, *args, readout_zero_init=False, output_mult=1.0, **kwargs):
        self.output_mult = output_mult
        self.readout_zero_init = readout_zero_init
        super().__init__(*args, **kwargs)

    def width_mult(self):
        return self.in_features / 256.0

    def forward(self, x):
        return super().forward(self.output_mult * x / self.width_mult())


def set_base_shapes(model, base, rescale_params=True, **kwargs):
    return model


def normal_(tensor, mean=0.0, std=1.0):
    return torch.nn.init.normal_(tensor, mean=mean, std=std)


class MuAdamW(torch.optim.AdamW):
    pass
