"""Stand-in for `ema_pytorch.EMA` at sampling time.  sample.py only uses it as a container whose state dict has
`initted`, `step`, `online_model.*`, `ema_model.*` (ema_pytorch 0.0.8; sample.py:154-167) and then reads
`.ema_model` (:179, :224).  Training-time averaging is out of scope."""
from __future__ import annotations

import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model: nn.Module, **_unused):
        super().__init__()
        self.online_model = model
        self.ema_model = model          # one copy: only the averaged weights are ever sampled from
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("step", torch.tensor([0]))

    def load_state_dict(self, state_dict, strict: bool = True):
        ema = {k[len("ema_model."):]: v for k, v in state_dict.items() if k.startswith("ema_model.")}
        if not ema:
            raise KeyError("checkpoint has no 'ema_model.*' entries")
        res = self.ema_model.load_state_dict(ema, strict=strict)
        for k in ("initted", "step"):
            if k in state_dict:
                getattr(self, k).copy_(state_dict[k].reshape(getattr(self, k).shape))
        return res
