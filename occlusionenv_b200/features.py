"""Row N-1 of SURVEY.md section 8(f): the FROZEN encoder of the reference's policy, evaluated on the env ranks.

``PPO.select_action`` (``/root/reference/PPO.py:152-164``) feeds every observation through
``ActorCritic.extract_features`` = ``FullNetwork(8, dilation=2, separable=True)`` (``PPO.py:46-47,61-63``;
``model.py:156-166``) under ``no_grad`` and buffers the 256 pooled features, the optimiser only sees the two linear
heads (``PPO.py:116-119``).  The encoder is therefore a fixed function of the observation: running it where the
observation is produced turns the 262 144 B/env that the learner gather moves at 128^2 into 1 024 B/env.

This module restates the forward pass of ``Encoder`` + the global average pool (``model.py:8-23,36-52,87-107,
161-162``) functionally, on the state dict of the reference module -- the key names are the reference's, so a
checkpoint saved from ``FullNetwork`` loads as it is:

    initial            Conv(4, ch, 3, dilation 1, separable)                    model.py:92
    features.i  x5     ConvBlock(ch 2^i): 2 x Conv(c, c, 3, dilation d, separable) [+ x, residual], then
                       down = Conv(c, 2c, 3, stride 2, dilation 1, dense)          model.py:36-52
    Conv               depthwise (3,1) -> depthwise (1,3) -> 1x1 (+bias) | dense 3x3 ; then ReLU ; then BatchNorm
                       (eval statistics: the encoder is frozen)                  model.py:8-23
    pooled = mean over H, W of the last ``down`` output                       model.py:161-162

The convolutions are library calls (cuDNN through torch, channels-last): this is policy-side arithmetic next to the
hot path, not one of its hand-written kernels.  No CPU fallback is involved in the product path: it runs on whatever
device the observations live on (the CPU tests run it on the host against the golden fixture).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default


def random_state_dict(ch: int = 8, levels: int = 5, layers: int = 2, seed: int = 0) -> Dict[str, torch.Tensor]:
    """A state dict with the reference's key names and shapes and seeded random values (the reference's checkpoint,
    ``./models/bestSegModel_final2_dice_l1_dilated_res_sep.pt`` at ``PPO.py:48``, is not in its tree)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def rnd(*shape, scale=1.0):
        return (torch.rand(*shape, generator=g) - 0.5) * 2.0 * scale

    def conv(prefix, cin, cout, separable):
        if separable:
            sd[f"{prefix}.conv.0.weight"] = rnd(cin, 1, 3, 1, scale=0.8)
            sd[f"{prefix}.conv.1.weight"] = rnd(cin, 1, 1, 3, scale=0.8)
            sd[f"{prefix}.conv.2.weight"] = rnd(cout, cin, 1, 1, scale=(3.0 / cin) ** 0.5)
            sd[f"{prefix}.conv.2.bias"] = rnd(cout, scale=0.1)
        else:
            sd[f"{prefix}.conv.weight"] = rnd(cout, cin, 3, 3, scale=(1.0 / (3.0 * cin)) ** 0.5)
            sd[f"{prefix}.conv.bias"] = rnd(cout, scale=0.1)
        sd[f"{prefix}.bn.weight"] = 1.0 + rnd(cout, scale=0.2)
        sd[f"{prefix}.bn.bias"] = rnd(cout, scale=0.1)
        sd[f"{prefix}.bn.running_mean"] = 0.3 + rnd(cout, scale=0.2)
        sd[f"{prefix}.bn.running_var"] = 0.5 + torch.rand(cout, generator=g)

    conv("encoder.initial", 4, ch, True)
    for i in range(levels):
        c = ch * 2 ** i
        for k in range(layers):
            conv(f"encoder.features.{i}.net.Layer {k + 1}", c, c, True)
        conv(f"encoder.features.{i}.down", c, 2 * c, False)
    return sd


class FrozenEncoder:
    """``features = FrozenEncoder(state_dict)(obs)``: obs (N,4,S,S) -> (N, ch * 2^levels) pooled features, what
    ``ActorCritic.extract_features`` returns (``PPO.py:61-63``)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None, dilation: int = 2, residual: bool = True,
                 dtype: torch.dtype = torch.float32, chunk: int = 1024):
        self.dilation, self.residual, self.dtype, self.chunk = int(dilation), bool(residual), dtype, int(chunk)
        sd = {k[len("encoder."):]: v for k, v in state_dict.items() if k.startswith("encoder.")}
        if "initial.conv.0.weight" not in sd:
            raise KeyError("state dict has no 'encoder.initial.conv.0.weight': not a FullNetwork(separable=True) checkpoint")
        self.levels = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("features."))
        self.layers = max(int(k.split("Layer ")[1].split(".")[0]) for k in sd if "Layer " in k)
        dev = torch.device(device) if device is not None else sd["initial.conv.0.weight"].device
        self.device = dev
        self.p: Dict[str, torch.Tensor] = {}
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                continue
            self.p[k] = v.detach().to(device=dev, dtype=torch.float32)
        # BatchNorm with frozen statistics = one scale and shift per channel
        for k in [k for k in self.p if k.endswith(".bn.weight")]:
            pre = k[: -len(".weight")]
            scale = self.p[pre + ".weight"] / torch.sqrt(self.p[pre + ".running_var"] + EPS)
            self.p[pre + ".scale"] = scale.view(1, -1, 1, 1).to(dtype)
            self.p[pre + ".shift"] = (self.p[pre + ".bias"] - self.p[pre + ".running_mean"] * scale).view(1, -1, 1, 1).to(dtype)
        for k in [k for k in self.p if ".conv" in k]:
            self.p[k] = self.p[k].to(dtype)
        self.out_features = int(self.p[f"features.{self.levels - 1}.down.conv.weight"].shape[0])

    def _conv(self, x, prefix, separable, stride, dilation):
        p = self.p
        if separable:
            pad = (3 + dilation - 1) // 2
            c = x.shape[1]
            x = F.conv2d(x, p[prefix + ".conv.0.weight"], None, 1, (pad, 0), (dilation, 1), c)
            x = F.conv2d(x, p[prefix + ".conv.1.weight"], None, 1, (0, pad), (1, dilation), c)
            x = F.conv2d(x, p[prefix + ".conv.2.weight"], p[prefix + ".conv.2.bias"])
        else:
            x = F.conv2d(x, p[prefix + ".conv.weight"], p[prefix + ".conv.bias"], stride, (3 + dilation - 1) // 2, dilation)
        return torch.relu(x) * p[prefix + ".bn.scale"] + p[prefix + ".bn.shift"]     # bn(relu(conv(x)))   model.py:22-23

    @torch.no_grad()
    def _forward(self, x):
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = self._conv(x, "initial", True, 1, 1)
        for i in range(self.levels):
            y = x
            for k in range(self.layers):
                y = self._conv(y, f"features.{i}.net.Layer {k + 1}", True, 1, self.dilation)
            if self.residual:
                y = y + x
            x = self._conv(y, f"features.{i}.down", False, 2, 1)
        return x.float().mean(dim=(2, 3))

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = obs.shape[0]
        if out is None:
            out = torch.empty(n, self.out_features, dtype=torch.float32, device=obs.device)
        for lo in range(0, n, self.chunk):           # bounded activation memory: ~3 MB per env at 128^2
            out[lo:lo + self.chunk] = self._forward(obs[lo:lo + self.chunk])
        return out
