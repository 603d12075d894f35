"""Frozen-encoder throughput on one GPU for layouts / dtypes (row N-1): ms per 1024 observations at 128^2."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.features import FrozenEncoder, random_state_dict
import torch.nn.functional as F
n, S = 1024, 128
x = torch.rand(n, 4, S, S, device="cuda")
sd = random_state_dict(8)
def run(tag, enc):
    for _ in range(2): enc(x)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(3): f = enc(x)
    torch.cuda.synchronize(); dt = (time.time() - t) / 3
    print(f"{tag:40s} {dt*1e3:8.2f} ms / {n} obs -> {n/dt:10.0f} obs/s  (feat mean {float(f.mean()):.4f})", flush=True)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    run(f"fp32 channels_last tf32={tf32}", FrozenEncoder(sd, device="cuda", chunk=1024))
run("bf16 channels_last", FrozenEncoder(sd, device="cuda", dtype=torch.bfloat16, chunk=1024))
run("fp16 channels_last", FrozenEncoder(sd, device="cuda", dtype=torch.float16, chunk=1024))
# NCHW variant: monkeypatch the layout conversion
class NCHW(FrozenEncoder):
    @torch.no_grad()
    def _forward(self, x):
        x = x.to(self.dtype).contiguous()
        x = self._conv(x, "initial", True, 1, 1)
        for i in range(self.levels):
            y = x
            for k in range(self.layers):
                y = self._conv(y, f"features.{i}.net.Layer {k + 1}", True, 1, self.dilation)
            if self.residual: y = y + x
            x = self._conv(y, f"features.{i}.down", False, 2, 1)
        return x.float().mean(dim=(2, 3))
torch.backends.cudnn.allow_tf32 = True
run("fp32 NCHW", NCHW(sd, device="cuda", chunk=1024))
run("bf16 NCHW", NCHW(sd, device="cuda", dtype=torch.bfloat16, chunk=1024))
torch.backends.cudnn.benchmark = True
run("fp32 NCHW cudnn.benchmark", NCHW(sd, device="cuda", chunk=1024))
run("fp32 channels_last cudnn.benchmark", FrozenEncoder(sd, device="cuda", chunk=1024))
