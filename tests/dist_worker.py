"""Worker of tests/test_gpu_multi.py, one process per GPU (launched with torch.distributed.run, NCCL over NVLink).

Checks SURVEY.md section 4 (iv) on hardware:
  1. the sharded run equals the single-GPU run env for env: every rank steps its contiguous shard of the global batch
     (dist.shard_range, env_offset) and compares its rows with a full-batch run on its own GPU -- bit-exact, the
     integer accumulators of the rasteriser make every output deterministic;
  2. LearnerGather delivers rank-ordered rows: the learner rank compares the gathered obs / reward / done (and the
     fused variant that renders straight into the gather buffer) with the full-batch run;
  3. scatter_actions hands every rank its slice.
Exits non-zero on any mismatch.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    from occlusionenv_b200.dist import LearnerGather, scatter_actions, shard_range
    import bench

    n_total, S, steps = 64 * world, 64, 3
    lo, hi = shard_range(n_total, rank, world)
    n_local = hi - lo
    az, el, actions = bench.make_poses(n_total, 0)            # the same global inputs on every rank
    full = BatchedOcclusionVecEnv(n_total, data="box", img_size=S, device=dev, auto_reset=False)
    mine = BatchedOcclusionVecEnv(n_local, data="box", img_size=S, device=dev, auto_reset=False, env_offset=lo)
    full.engine.reset(radius=4.0, azimuth=az, elevation=el)
    mine.engine.reset(radius=4.0, azimuth=az[lo:hi], elevation=el[lo:hi])
    lg = LearnerGather(n_local, (4, S, S), dev, dst=0)
    fused_obs = lg.obs_send_buffer()                          # the rasteriser writes obs straight into the comm buffer
    ok = True
    for t in range(steps):
        a_all = actions[t % 8]
        a_mine = scatter_actions(a_all if rank == 0 else None, n_local, src=0, device=dev)
        ok &= torch.equal(a_mine.cpu(), a_all[lo:hi])
        f_obs, f_rew, f_done, _ = full.step(a_all.to(dev))
        full_obs, full_rew, full_done = f_obs.clone(), f_rew.clone(), full.engine.done.clone()
        mine.engine.step(a_mine, obs=fused_obs)
        m_obs, m_rew, m_done = fused_obs, mine.engine.reward, mine.engine.done
        for name, got, want in (("obs", m_obs, full_obs[lo:hi]), ("reward", m_rew, full_rew[lo:hi]),
                                ("done", m_done, full_done[lo:hi]), ("occl", mine.engine.occl, full.engine.occl[lo:hi]),
                                ("n_visible", mine.engine.n_visible, full.engine.n_visible[lo:hi])):
            if not torch.equal(got, want):
                ok = False
                print(f"rank {rank} step {t}: shard != single-GPU run in {name}", flush=True)
        g_obs, g_rew, g_done = lg.gather(m_obs, m_rew, m_done)
        torch.cuda.synchronize()
        if rank == 0:
            for name, got, want in (("obs", g_obs, full_obs), ("reward", g_rew, full_rew), ("done", g_done, full_done)):
                if not torch.equal(got, want):
                    ok = False
                    print(f"step {t}: gathered {name} is not the rank-ordered full batch", flush=True)
    # 4. the fused transport: the rasteriser stores the observation rows straight into the learner's HBM (peer memory
    #    over NVLink), compact 2-plane layout; only reward + done go through NCCL
    from occlusionenv_b200.config import RasterConfig
    from occlusionenv_b200.dist import expand_compact_obs
    mine2 = BatchedOcclusionVecEnv(n_local, data="box", img_size=S, device=dev, auto_reset=False, env_offset=lo,
                                   cfg=RasterConfig(image_size=S, obs_planes=2))
    full.engine.reset(radius=4.0, azimuth=az, elevation=el)
    mine2.engine.reset(radius=4.0, azimuth=az[lo:hi], elevation=el[lo:hi])
    lg2 = LearnerGather(n_local, (2, S, S), dev, dst=0, transport="p2p")
    peer_obs = lg2.obs_send_buffer()
    for t in range(2):
        a_all = actions[t % 8]
        f_obs, f_rew, _, _ = full.step(a_all.to(dev))
        mine2.engine.step(a_all[lo:hi].to(dev), obs=peer_obs)
        g_obs, g_rew, g_done = lg2.gather(peer_obs, mine2.engine.reward, mine2.engine.done)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            if not (torch.equal(expand_compact_obs(g_obs), f_obs) and torch.equal(g_rew, f_rew)
                    and torch.equal(g_done, full.engine.done)):
                ok = False
                print(f"step {t}: peer-memory delivery differs from the full batch", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print(f"dist_worker OK: world {world}, {n_total} envs, {steps} steps", flush=True)


if __name__ == "__main__":
    main()
