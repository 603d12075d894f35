"""Stage-by-stage check of the peer-memory delivery (run with torchrun --nproc-per-node 2, CUDA_LAUNCH_BLOCKING=1)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
from occlusionenv_b200.dist import _share_cuda_tensor
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import default_scene
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
n, S = 4, 32
buf = torch.zeros(2 * n, 4, S, S, device=dev) if rank == 0 else None
peer = _share_cuda_tensor(buf, 0)
say("peer view ptr", hex(peer.data_ptr()), "current device", torch.cuda.current_device())
dist.barrier(); torch.cuda.synchronize()
eng = OcclusionEngine(default_scene("box"), n, RasterConfig(image_size=S), device=dev)
eng.reset(radius=4.0, azimuth=1.5, elevation=0.1)
torch.cuda.synchronize()
say("local reset ok")
sl = peer[rank * n:(rank + 1) * n]
say("slice ptr", hex(sl.data_ptr()))
eng.step(torch.zeros(n, 2, device=dev), obs=sl)
torch.cuda.synchronize()
say("step into peer slice ok")
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    say("means per rank slice:", float(buf[:n].mean()), float(buf[n:].mean()), "local obs mean", float(eng.obs.mean()))
dist.barrier()
dist.destroy_process_group()
