"""Where does the fused peer-memory delivery lose its 12 % against NCCL under incast?  (torchrun, >= 3 GPUs)
  A  stores only: every rank renders into its slice of the learner's buffer, nothing else is exchanged
  B  the bench leg: + reward / done through one NCCL group on the same stream, every step
  C  the reward / done exchange on a side stream (copies of the two small tensors), the next step does not wait for it
usage: python -m torch.distributed.run --nproc-per-node N tools/p2p_legs_probe.py [planes]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
import bench
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.dist import LearnerGather
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import default_scene

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N, S, K = 8192, 128, 20
h = bench.Harness(world, dev)
az, el, actions = bench.make_poses(N, 0, offset=rank * 1000003)
acts = actions.to(dev)
eng = OcclusionEngine(default_scene("box"), N, RasterConfig(image_size=S, obs_planes=planes), device=dev)
eng.reset(radius=4.0, azimuth=az, elevation=el)
lg = LearnerGather(N, (planes, S, S), dev, dst=0, transport="p2p")
buf = lg.obs_send_buffer()
main = torch.cuda.current_stream()
comm = torch.cuda.Stream(device=dev)
small = [(torch.empty_like(eng.reward), torch.empty(N, dtype=torch.uint8, device=dev)) for _ in range(2)]
sent = [torch.cuda.Event() for _ in range(2)]
for e in sent:
    e.record(comm)


def step_a(i):
    eng.step(acts[i % 8], obs=buf)


def step_b(i):
    eng.step(acts[i % 8], obs=buf)
    lg.gather(buf, eng.reward, eng.done)


def step_c(i):
    k = i & 1
    eng.step(acts[i % 8], obs=buf)
    main.wait_event(sent[k])               # the staging pair is free again
    small[k][0].copy_(eng.reward); small[k][1].copy_(eng.done)
    ready = torch.cuda.Event(); ready.record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(ready)
        lg.gather(buf, small[k][0], small[k][1])
        sent[k].record(comm)


nbytes = (world - 1) * N * planes * S * S * 4
for name, fn in (("A stores only", step_a), ("B + reward/done on the same stream", step_b), ("C + reward/done on a side stream", step_c)):
    for i in range(4):
        fn(i)
    main.wait_stream(comm)
    def timed(i, fn=fn):
        fn(i)
        if i == K - 1:
            main.wait_stream(comm)
    ms = h.timed(timed, K)
    if rank == 0:
        print(f"{name:40s} {world * N * K / (ms * 1e-3) / 1e6:6.2f} M env-steps/s  {nbytes / (ms / K * 1e-3) / 1e9:6.1f} GB/s into the learner", flush=True)
dist.destroy_process_group()
