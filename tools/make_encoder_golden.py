"""Golden fixture for occlusionenv_b200/features.py: the REFERENCE module (``/root/reference/model.py``, pure torch,
importable in the build container) with seeded random weights -- its own checkpoint is not in the tree -- evaluated on
a seeded observation batch.  Writes tests/golden/encoder_ch2.npz: the encoder part of the state dict, the input and
``pooled_features`` of ``FullNetwork(2, dilation=2, separable=True)`` in eval mode (ch = 2 keeps the file small; the
architecture is that of ``PPO.py:47`` otherwise).  Run once here; the GPU box has no /root/reference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import model  # noqa: E402  (the reference's model.py)

torch.manual_seed(0)
net = model.FullNetwork(2, dilation=2, separable=True)
with torch.no_grad():  # non-trivial BatchNorm statistics and affine parameters
    for k, v in net.state_dict().items():
        if k.endswith("running_mean"):
            v.copy_(0.3 * torch.randn_like(v))
        elif k.endswith("running_var"):
            v.copy_(0.5 + torch.rand_like(v))
        elif k.endswith("bn.weight"):
            v.copy_(1.0 + 0.2 * torch.randn_like(v))
        elif k.endswith("bn.bias"):
            v.copy_(0.1 * torch.randn_like(v))
net.eval()
x = torch.rand(3, 4, 64, 64)
x[:, 3] = torch.where(x[:, 3] > 0.5, 3.0 + x[:, 3], torch.full_like(x[:, 3], -1.0))  # depth plane: -1 on the background
with torch.no_grad():
    pooled, _, _ = net(x)
out = {"input": x.numpy(), "pooled": pooled.numpy()}
for k, v in net.state_dict().items():
    if k.startswith("encoder.") and not k.endswith("num_batches_tracked"):
        out["sd:" + k] = v.numpy()
path = os.path.join(ROOT, "tests", "golden", "encoder_ch2.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes; pooled", pooled.shape)
