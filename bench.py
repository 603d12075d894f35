#!/usr/bin/env python
"""Benchmark of the OcclusionEnv transition (env-steps/s) -- see BASELINE.json / DESIGN.md section 6.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA, sm_100a)
  python bench.py --impl reference [--gpus N] --steps K --warmup W  # CPU reference arm (oracle on host cores)

A "step" is one transition of every environment of the batch: pose -> projection -> rasterisation of the
target + occluder (soft silhouettes, K=100) and of the scene (K=1, flat-shaded RGBD) -> occlusion loss ->
reward / done.  Workload at N=1: BASELINE config 2 (teapot + box occluder, 4096 envs, 128x128, forward +
reward); with N GPUs every rank owns 4096 envs (weak scaling, no collective inside the step).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
METRIC = "env-steps/sec (render+reward, fwd)"
UNIT = "env-steps/s"


def workload_name(args):
    return (f"teapot + {args.occluder} occluder, {args.envs} envs/GPU, {args.size}x{args.size}, "
            f"{'fwd+bwd to action' if args.grad else 'forward render + reward'}")


def algorithmic_bytes_per_env_step(S: int, grad: bool) -> int:
    # S^2 * (16 B RGBD obs + 4 B occlusion map) + 32 B (action, state, reward, done, loss) [+12 B grad]
    return S * S * 20 + 32 + (12 if grad else 0)


def make_poses(n, seed, offset=0):
    """BASELINE C2 inputs: r=4, az ~ U(pi/2-0.6, pi/2+0.6), el ~ U(-0.3, 0.3) (seed 0); actions ~ N(0,1)^2 (seed 1)."""
    import torch
    g = torch.Generator().manual_seed(seed + offset)
    az = (math.pi / 2 - 0.6) + 1.2 * torch.rand(n, generator=g)
    el = -0.3 + 0.6 * torch.rand(n, generator=g)
    ga = torch.Generator().manual_seed(seed + 1 + offset)
    a = torch.randn(4, n, 2, generator=ga)
    # a, -a, b, -b, ...: unit steps that cancel in pairs, so the pose distribution (and with it the work per
    # step) stays the one of the reset however many steps are timed
    actions = torch.stack([a[0], -a[0], a[1], -a[1], a[2], -a[2], a[3], -a[3]])
    return az, el, actions


# --------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port driven like SimpleVecEnv (sequential per process)
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    occluder, S, n_env, n_steps, seed = args
    import numpy as np
    from occlusionenv_b200.meshes import default_scene
    from oracle import oracle as O
    sc = default_scene(occluder)
    rng = np.random.default_rng(seed)
    envs = [O.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S) for _ in range(n_env)]
    vec = O.OracleSimpleVecEnv(envs)
    for e in envs:  # reset is not timed (auto-reset excluded from the metric)
        e.reset(radius=4.0, azimuth=float(rng.uniform(math.pi / 2 - 0.6, math.pi / 2 + 0.6)),
                elevation=float(rng.uniform(-0.3, 0.3)))
    times = []
    for _ in range(n_steps):
        a = rng.normal(size=(n_env, 2)).astype(np.float32)
        t = time.perf_counter()
        for i, e in enumerate(envs):  # SimpleVecEnv.step_wait loop without the auto-reset
            e.step(a[i])
        times.append(time.perf_counter() - t)
    return times


def cpu_reference_run(occluder, S, procs, envs_per_proc, steps, warmup):
    """`procs` processes, each stepping its own sequential shard; returns (env-steps/s, ms per step)."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    work = [(occluder, S, envs_per_proc, steps + warmup, 100 + p) for p in range(procs)]
    if procs == 1:
        res = [_cpu_worker(work[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, work)
    # a "step" ends when the slowest process has finished its shard
    per_step = [max(r[i] for r in res) for i in range(warmup, warmup + steps)]
    total = sum(per_step)
    return procs * envs_per_proc * steps / total, 1e3 * total / steps


def run_reference(args, rank, world):
    if rank != 0:
        return
    procs = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(procs, 64))  # one process per usable host core (capped: start-up cost)
    n = args.steps + args.warmup
    est = 0.45 if args.size == 128 else 0.45 * (args.size / 128.0) ** 2  # s per env-step per core (measured)
    envs_per_proc = int(max(1, min(8, 150.0 / (n * est))))
    v, ms = cpu_reference_run(args.occluder, args.size, procs, envs_per_proc, args.steps, args.warmup)
    sample = (f"{procs} processes x {envs_per_proc} envs stepped sequentially per step (SimpleVecEnv loop), "
              f"same scene/poses distribution as the GPU workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU oracle port of the reference's pytorch3d-naive path "
                   "(pytorch3d itself is not installable offline); restated semantics, not the pytorch3d binary"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "display_clock_setting": 0x100,
                 "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on ONE host core (scalar port): ~12-25 s
        n_env = 96 if args.size <= 128 else 24
        v, ms = cpu_reference_run(args.occluder, args.size, 1, n_env, 1, 0)
        cpu_base = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{n_env} env-steps of the same workload, sequential SimpleVecEnv loop on 1 core "
                              "(oracle port of the pytorch3d-naive CPU path)"}

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        if "OCCL_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["OCCL_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    from occlusionenv_b200.config import RasterConfig

    N, S, K, W = args.envs, args.size, args.steps, args.warmup
    cfg = RasterConfig(image_size=S, tile_w=args.tile_w, tile_h=args.tile_h)
    venv = BatchedOcclusionVecEnv(N, data=args.occluder, img_size=S, device=dev, auto_reset=False, cfg=cfg,
                                  env_offset=rank * N)
    eng = venv.engine
    az, el, actions_host = make_poses(N, 0, offset=rank * 1000003)
    actions_dev = actions_host.to(dev)
    actions_pinned = actions_host.pin_memory()

    def reset():
        eng.reset(radius=4.0, azimuth=az, elevation=el)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    grad = bool(args.grad)

    # ---- (1) device-resident throughput: the fused C-ABI chain, inputs already in HBM -----------------
    def dev_step(i):
        eng.step(actions_dev[i % 8], with_grad=grad)

    reset()
    for i in range(W):
        dev_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(dev_step, K)
    clocks = sampler.stop()
    status = int(eng.status.max().item())
    value = world * N * K / (ms_total * 1e-3)

    # ---- (2) the rasteriser alone (dominant kernel), CUDA events around its launch -------------------
    reset()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(min(W, 3)):
        eng.step_staged(actions_dev[i % 8], with_grad=grad)
    torch.cuda.synchronize()
    for i in range(K):
        eng.step_staged(actions_dev[i % 8], with_grad=grad, raster_events=evs[i])
    torch.cuda.synchronize()
    raster_ms = sum(a.elapsed_time(b) for a, b in evs) / K

    # ---- (3) end to end through the public API: host actions in, host rewards/dones out ---------------
    reset()
    rew_host = torch.empty(N, dtype=torch.float32).pin_memory()
    done_host = torch.empty(N, dtype=torch.uint8).pin_memory()

    def e2e_step(i):
        a = actions_pinned[i % 8]
        if grad:
            a = a.to(dev, non_blocking=True).requires_grad_(True)
        obs, rews, dones, infos = venv.step(a)
        rew_host.copy_(rews.detach(), non_blocking=True)
        done_host.copy_(eng.done, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the learner needs reward/done before the next action

    for i in range(min(W, 3)):
        e2e_step(i)
    ms_e2e = timed(e2e_step, K)
    e2e_value = world * N * K / (ms_e2e * 1e-3)

    # ---- (4) optional: learner-boundary gather over NVLink (config 5) ---------------------------------
    gather = None
    if world > 1 and args.gather:
        from occlusionenv_b200.dist import LearnerGather
        obs_bytes = (world - 1) * N * (4 * S * S * 4 + 5)
        # (a) serial: step all envs, then gather (obs, reward, done) to rank 0
        lg = LearnerGather(N, (4, S, S), dev, dst=0)
        reset()

        def gather_step(i):
            eng.step(actions_dev[i % 8], with_grad=False)
            lg.gather(eng.obs, eng.reward, eng.done)

        for i in range(min(W, 3)):
            gather_step(i)
        ms_g = timed(gather_step, K)
        gather = {"serial": {"value": world * N * K / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g / K},
                  "bytes_to_learner_per_step": obs_bytes, "collective": "gather(obs,reward,done)->rank0 (NCCL over NVLink)"}
        del lg
        # (b) double-buffered: the envs of a rank form two half batches; while half A's observations travel to the
        # learner (NCCL on a side stream) half B renders.  Same work and same bytes per step, on-policy per half.
        from occlusionenv_b200.engine import OcclusionEngine
        from occlusionenv_b200.meshes import default_scene
        H = N // 2
        sc = default_scene(args.occluder)
        halves = [OcclusionEngine(sc, H, cfg, device=dev) for _ in range(2)]
        for h, e2 in enumerate(halves):
            e2.reset(radius=4.0, azimuth=az[h * H:(h + 1) * H], elevation=el[h * H:(h + 1) * H])
        lgs = [LearnerGather(H, (4, S, S), dev, dst=0) for _ in range(2)]
        comm = torch.cuda.Stream(device=dev)
        rendered = [torch.cuda.Event() for _ in range(2)]
        gathered = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream()

        def gather_step_db(i):
            a = actions_dev[i % 8]
            for h in range(2):
                main.wait_event(gathered[h])            # half h's buffers are free again
                halves[h].step(a[h * H:(h + 1) * H].contiguous(), with_grad=False)
                rendered[h].record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(rendered[h])
                    lgs[h].gather(halves[h].obs, halves[h].reward, halves[h].done)
                    gathered[h].record(comm)

        for h in range(2):
            gathered[h].record(comm)
        for i in range(min(W, 3)):
            gather_step_db(i)
        main.wait_stream(comm)

        def db_and_drain(i):
            gather_step_db(i)
            if i == K - 1:
                main.wait_stream(comm)

        ms_db = timed(db_and_drain, K)
        gather["double_buffered"] = {"value": world * N * K / (ms_db * 1e-3), "unit": UNIT, "ms_per_step": ms_db / K,
                                     "note": "two half batches per rank; gather of one half overlaps the render of the other"}
        gather["value"] = gather["double_buffered"]["value"]
        gather["unit"] = UNIT

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpe = algorithmic_bytes_per_env_step(S, grad)
    achieved = N * bpe / (raster_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "raster_ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("envs") == N and t.get("size") == S and t.get("occluder") == args.occluder and bool(t.get("grad")) == grad:
            traffic = t.get("dram_bytes_per_launch")
    line = {
        "metric": METRIC if not grad else "env-steps/sec (render+reward, fwd+bwd)", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "envs_per_gpu": N, "image_size": S, "occluder": args.occluder,
                   "faces": int(eng.c.n_faces), "faces_per_pixel": 100, "tile": [int(eng.c.tile_w), int(eng.c.tile_h)],
                   "l2": "outputs (obs+occlusion map: %.0f MB/step/GPU) larger than L2; no flush needed" % (N * S * S * 20 / 1e6),
                   "auto_reset": "excluded", "status_or": status},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * 2 * 4, "d2h_bytes_per_step": N * 5,
                "ms_per_step": ms_e2e / K,
                "note": "BatchedOcclusionVecEnv.step(pinned host actions) -> rewards+dones copied to pinned host, stream sync "
                        "every step; observations stay in HBM for the policy, as in the reference (device tensors)"},
        "gpu_launches": 6 * K,  # pose, project, face_setup, raster, raster_clip (cut faces; CTAs leave at once otherwise), finalize
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "raster_kernel", "kernel_ms": raster_ms,
                     "kernel_share_of_step": raster_ms / (ms_total / K),
                     "algorithmic_bytes_per_env_step": bpe, "peak_source": peak_src,
                     "note": "issue-slot bound, not HBM bound: see DESIGN.md section 5"},
        "cpu_baseline": cpu_base,
    }
    if gather is not None:
        line["gather"] = gather
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--occluder", default="box", choices=["box", "teapot"])
    ap.add_argument("--grad", action="store_true", help="differentiable step (config 4)")
    ap.add_argument("--gather", action="store_true", help="also time step + NCCL gather to rank 0 (config 5)")
    ap.add_argument("--tile-w", type=int, default=0)
    ap.add_argument("--tile-h", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints "NCCL version ..." on fd 1) and any stray
    # print are sent to stderr, the JSON line goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
