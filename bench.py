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
def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Harness:
    """Timing helpers shared by the legs: barrier + synchronize on both sides, CUDA events on the launching stream,
    max over ranks."""

    def __init__(self, world, dev):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        self.barrier()
        return ms


def raster_kernel_ms(eng, actions_dev, grad, steps, warm):
    """Average duration of the rasteriser launches (face setup + raster + clip) of one step: CUDA events around
    occl_raster on the launching stream, the transition driven through the split C-ABI."""
    import torch
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(warm):
        eng.step_staged(actions_dev[i % len(actions_dev)], with_grad=grad)
    torch.cuda.synchronize()
    for i in range(steps):
        eng.step_staged(actions_dev[i % len(actions_dev)], with_grad=grad, raster_events=evs[i])
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def c3_scenes(n_distinct):
    from occlusionenv_b200.meshes import procedural_scene
    return [procedural_scene(s, n_obj=3, subdiv=5) for s in range(n_distinct)]


def run_config3(h, dev, n_envs, steps, warm):
    """BASELINE config 3: per-env procedural meshes of three 20 480-face objects (ShapeNet layout), 256x256."""
    import numpy as np
    import torch
    from occlusionenv_b200.config import RasterConfig
    from occlusionenv_b200.engine import OcclusionEngine
    S, n_distinct = 256, 64
    scenes = c3_scenes(n_distinct)
    # engine with per-env mesh slots; the 64 distinct scenes are replicated on the device (every env still reads
    # ITS OWN 1.1 MB of vertices and faces from HBM, which is what the algorithmic bytes count)
    eng = OcclusionEngine(None, n_envs, RasterConfig(image_size=S), device=dev, per_env_scenes=scenes, replicate_scenes=True)
    g = torch.Generator().manual_seed(0)
    az = -0.5 + torch.rand(n_envs, generator=g)
    eng.reset(radius=4.0, azimuth=az, elevation=0.1)
    a = torch.randn(2, n_envs, 2, generator=g)
    acts = torch.stack([a[0], -a[0], a[1], -a[1]]).to(dev)

    def step(i):
        eng.step(acts[i % 4])

    for i in range(warm):
        step(i)
    ms = h.timed(step, steps)
    kms = raster_kernel_ms(eng, acts, False, max(3, steps // 4), 1)
    V, F = int(eng.c.n_verts), int(eng.c.n_faces)
    bpe = S * S * 20 + 32 + 12 * (V + F)
    peak, _ = _peak()
    ach = n_envs * bpe / (kms * 1e-3) / 1e9
    out = {"value": n_envs * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warm,
           "envs": n_envs, "image_size": S, "faces_per_env": F, "verts_per_env": V, "objects": 3,
           "distinct_scenes": n_distinct, "tile": [int(eng.c.tile_w), int(eng.c.tile_h)],
           "workspace_gb": eng.workspace.numel() / 2 ** 30, "status_or": int(eng.check_status(raise_on=0)),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "kernel_ms": kms, "algorithmic_bytes_per_env_step": bpe},
           "workload": "config 3: 3 x 20480-face procedural meshes per env (ShapeNet layout), 256x256, forward render + reward"}
    del eng
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on ONE host core (scalar port): one warm-up pass, one timed pass
        n_env = 32 if args.size <= 128 else 8
        v, ms = cpu_reference_run(args.occluder, args.size, 1, n_env, 1, 1)
        cpu_base = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{n_env} env-steps of the same workload (after {n_env} warm-up env-steps), sequential "
                              "SimpleVecEnv loop on 1 core (oracle port of the pytorch3d-naive CPU path)"}

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        if "OCCL_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["OCCL_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    from occlusionenv_b200.config import RasterConfig

    h = Harness(world, dev)
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

    grad = bool(args.grad)

    # ---- (1) device-resident throughput: the fused C-ABI chain, inputs already in HBM -----------------
    def dev_step(i):
        eng.step(actions_dev[i % 8], with_grad=grad)

    reset()
    for i in range(W):
        dev_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = h.timed(dev_step, K)
    value = world * N * K / (ms_total * 1e-3)
    # the same measurement over a longer region (>= 20 warm-up and >= 100 timed steps), so that the clock
    # sampler sees a loaded GPU for more than a few samples whatever K the caller asked for
    LW, LK = max(W, 20), max(K, 100)
    for i in range(LW):
        dev_step(i)
    ms_long = h.timed(dev_step, LK)
    clocks = sampler.stop()
    status = int(eng.check_status(raise_on=0))
    long_run = {"value": world * N * LK / (ms_long * 1e-3), "unit": UNIT, "steps": LK, "warmup": LW,
                "ms_per_step": ms_long / LK}

    # ---- (2) the rasteriser alone (dominant kernel), CUDA events around its launch -------------------
    reset()
    raster_ms = raster_kernel_ms(eng, actions_dev, grad, K, min(W, 3))

    # ---- (3) end to end through the public API: host actions in, host rewards/dones out ---------------
    reset()
    rew_host = torch.empty(N, dtype=torch.float32).pin_memory()
    done_host = torch.empty(N, dtype=torch.uint8).pin_memory()

    def e2e_step(i, env=venv):
        a = actions_pinned[i % 8]
        if grad:
            a = a.to(dev, non_blocking=True).requires_grad_(True)
        obs, rews, dones, infos = env.step(a)
        rew_host.copy_(rews.detach(), non_blocking=True)
        done_host.copy_(env.engine.done, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the learner needs reward/done before the next action

    for i in range(min(W, 3)):
        e2e_step(i)
    ms_e2e = h.timed(e2e_step, K)
    e2e_value = world * N * K / (ms_e2e * 1e-3)
    e2e_extra = {}
    if world == 1 and not grad:
        # (3b) the same with the auto-reset of SimpleVecEnv.step_wait enabled (masked device-side reset every step)
        venv.auto_reset = True
        for i in range(3):
            e2e_step(i)
        ms_ar = h.timed(e2e_step, K)
        done_frac = float(eng.done.float().mean().item())
        venv.auto_reset = False
        e2e_extra["auto_reset"] = {"value": N * K / (ms_ar * 1e-3), "unit": UNIT, "ms_per_step": ms_ar / K,
                                   "done_fraction_last_step": done_frac,
                                   "note": "as e2e, plus the masked device-side auto-reset of SimpleVecEnv.step_wait: the envs that "
                                           "finish are re-rendered at the reset pose (azimuth 0: unoccluded, so they finish again at "
                                           "once, SURVEY B-9) inside the same call -- extra renders, not idle launches"}
        # (3c) a host-side consumer of the observations (datasetGenerator.py, SimpleVecEnv users on the CPU): the
        # whole (N,4,S,S) observation batch is copied to pinned host memory every step as well
        reset()
        obs_host = torch.empty(N, 4, S, S, dtype=torch.float32).pin_memory()
        kk = max(3, min(K, 10))

        def e2e_obs_step(i):
            obs, rews, dones, infos = venv.step(actions_pinned[i % 8])
            obs_host.copy_(obs, non_blocking=True)
            rew_host.copy_(rews, non_blocking=True)
            done_host.copy_(eng.done, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_obs_step(0)
        ms_o = h.timed(e2e_obs_step, kk)
        e2e_extra["obs_to_host"] = {"value": N * kk / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o / kk, "steps": kk,
                                    "d2h_bytes_per_step": N * (4 * S * S * 4 + 5),
                                    "note": "observations copied to pinned host memory every step too (PCIe bound)"}
        del obs_host

    # ---- (4) learner-boundary exchange over NVLink (config 5): 8192 envs per GPU, by default when world > 1 -----------
    gather = None
    if world > 1 and not args.no_gather and not grad:
        gather = run_gather_legs(h, args, rank, world, dev, K, W)

    # ---- (5) the other BASELINE configurations, N = 1 only --------------------------------------------------------------
    config_results = None
    if world == 1 and not args.no_configs and not grad and args.size == 128 and args.occluder == "box":
        config_results = {}
        # config 4: differentiable step, same scene and batch
        ck = max(K, 20)
        reset()

        def grad_step(i):
            eng.step(actions_dev[i % 8], with_grad=True)

        for i in range(max(W, 5)):
            grad_step(i)
        ms_g = h.timed(grad_step, ck)
        kms_g = raster_kernel_ms(eng, actions_dev, True, ck, 3)
        peak, _ = _peak()
        bpe_g = algorithmic_bytes_per_env_step(S, True)
        ach_g = N * bpe_g / (kms_g * 1e-3) / 1e9
        config_results["c4_grad"] = {
            "value": N * ck / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g / ck, "steps": ck, "warmup": max(W, 5),
            "envs": N, "image_size": S, "status_or": int(eng.check_status(raise_on=0)),
            "roofline": {"bound": "hbm", "achieved": ach_g, "peak": peak, "unit": "GB/s", "frac": ach_g / peak,
                         "kernel_ms": kms_g, "algorithmic_bytes_per_env_step": bpe_g},
            "workload": "config 4: teapot + box occluder, 4096 envs, 128x128, differentiable step (fwd + gradient to the action)"}
        # config 3: dense per-env meshes at the BASELINE shape (8192 envs x 256^2); the C2 engine is released first
        del venv, eng
        torch.cuda.empty_cache()
        try:
            config_results["c3_dense"] = run_config3(h, dev, args.c3_envs, max(3, min(K, args.c3_steps)), 2)
        except Exception as ex:  # reported, never silently dropped
            config_results["c3_dense"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peak()
    bpe = algorithmic_bytes_per_env_step(S, grad)
    achieved = N * bpe / (raster_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "raster_ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("envs") == N and t.get("size") == S and t.get("occluder") == args.occluder and bool(t.get("grad")) == grad:
            traffic = t.get("dram_bytes_per_launch")
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * 2 * 4, "d2h_bytes_per_step": N * 5,
           "ms_per_step": ms_e2e / K,
           "note": "BatchedOcclusionVecEnv.step(pinned host actions) -> rewards+dones copied to pinned host, stream sync "
                   "every step; observations stay in HBM for the policy, as in the reference (device tensors)"}
    e2e.update(e2e_extra)
    line = {
        "metric": METRIC if not grad else "env-steps/sec (render+reward, fwd+bwd)", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "envs_per_gpu": N, "image_size": S, "occluder": args.occluder,
                   "faces": 2476 if args.occluder == "box" else 4928, "faces_per_pixel": 100, "tile": [int(eng.c.tile_w), int(eng.c.tile_h)],
                   "l2": "outputs (obs+occlusion map: %.0f MB/step/GPU) larger than L2; no flush needed" % (N * S * S * 20 / 1e6),
                   "auto_reset": "excluded (see e2e.auto_reset)", "status_or": status},
        "clocks": clocks,
        "long_run": long_run,
        "e2e": e2e,
        # pose, project, face_setup, raster, raster_clip (cut faces; CTAs leave at once otherwise), finalize
        "gpu_launches": 6 * K,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "raster_kernel (timed together with the face_setup_kernel and "
                     "raster_clip_kernel launches of the same occl_raster call: 86.5 % / 11.4 % / 0.1 % of the step in the "
                     "ncu launch list, profiles/r02_launches_summary.txt; traffic is raster_kernel's alone)",
                     "kernel_ms": raster_ms,
                     "kernel_share_of_step": raster_ms / (ms_total / K),
                     "algorithmic_bytes_per_env_step": bpe, "peak_source": peak_src,
                     "note": "issue-slot bound, not HBM bound: see DESIGN.md section 5"},
        "cpu_baseline": cpu_base,
    }
    if gather is not None:
        line["gather"] = gather
    if config_results is not None:
        line["config_results"] = config_results
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gather_legs(h, args, rank, world, dev, K, W):
    """BASELINE config 5: `--gather-envs` (8192) envs per GPU, step + delivery of obs / reward / done to the learner rank.
    Four ways to deliver (all leave rank-ordered (N_total, ...) tensors in the learner's HBM):
      serial           step, then ONE grouped NCCL transfer (no staging copy: the rasteriser renders into the send
                       buffer; on the learner into its slice of the gathered tensor)
      double_buffered  two half batches per rank; the transfer of one half (side stream) overlaps the render of the other
      p2p_fused        the rasteriser's epilogue stores the observation rows straight into the learner's HBM over NVLink
                       (peer-mapped gather buffer); only reward + done (5 B/env) go through NCCL
      p2p_fused_gray   the same with the compact 2-plane observation (grey + depth: the reference's R = G = B)
      double_buffered_gray  the double-buffered NCCL transfer with the compact observation
      p2p_fused[_gray]_incremental  the fused delivery storing only the tiles that changed (lossless: background tiles that
                       were background in the learner's buffer already stay as they are)
      features         row N-1: the frozen encoder on the env ranks, 256 floats per env to the learner"""
    import torch
    from occlusionenv_b200.config import RasterConfig
    from occlusionenv_b200.dist import LearnerGather
    from occlusionenv_b200.engine import OcclusionEngine
    from occlusionenv_b200.meshes import default_scene
    N, S = args.gather_envs, args.size
    sc = default_scene(args.occluder)
    az, el, actions_host = make_poses(N, 0, offset=rank * 1000003)
    actions_dev = actions_host.to(dev)
    warm = min(W, 3)
    out = {"envs_per_gpu": N, "envs_total": N * world, "image_size": S,
           "collective": "obs/reward/done of every rank -> learner rank 0 over NVLink"}

    def leg(name, planes, transport, halves, incremental=False):
        cfg = RasterConfig(image_size=S, obs_planes=planes)
        H = N // halves
        engs = [OcclusionEngine(sc, H, cfg, device=dev) for _ in range(halves)]
        for k, e in enumerate(engs):
            e.reset(radius=4.0, azimuth=az[k * H:(k + 1) * H], elevation=el[k * H:(k + 1) * H])
        lgs = [LearnerGather(H, (planes, S, S), dev, dst=0, transport=transport) for _ in range(halves)]
        bufs = [lg.obs_send_buffer() for lg in lgs]
        # incremental delivery: tiles that were background in the learner's buffer and still are do not travel again
        states = [e.incremental_obs(b) for e, b in zip(engs, bufs)] if incremental else None
        main = torch.cuda.current_stream()
        if halves == 1:
            def step(i):
                engs[0].step(actions_dev[i % 8], obs=bufs[0])
                lgs[0].gather(bufs[0], engs[0].reward, engs[0].done)
        else:
            comm = torch.cuda.Stream(device=dev)
            rendered = [torch.cuda.Event() for _ in range(halves)]
            gathered = [torch.cuda.Event() for _ in range(halves)]
            for k in range(halves):
                gathered[k].record(comm)

            def step(i):
                a = actions_dev[i % 8]
                for k in range(halves):
                    main.wait_event(gathered[k])            # half k's buffers are free again
                    engs[k].step(a[k * H:(k + 1) * H].contiguous(), obs=bufs[k])
                    rendered[k].record(main)
                    with torch.cuda.stream(comm):
                        comm.wait_event(rendered[k])
                        lgs[k].gather(bufs[k], engs[k].reward, engs[k].done)
                        gathered[k].record(comm)
                if i == -1:
                    main.wait_stream(comm)
        for i in range(warm):
            step(i)
        if halves > 1:
            main.wait_stream(comm)

        def timed_step(i):
            step(i)
            if halves > 1 and i == K - 1:
                main.wait_stream(comm)

        ms = h.timed(timed_step, K)
        nbytes = (world - 1) * N * (planes * S * S * 4 + 5)
        out[name] = {"value": world * N * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "obs_planes": planes,
                     "transport": transport, "bytes_to_learner_per_step": nbytes,
                     "learner_ingest_gbs": nbytes / (ms / K * 1e-3) / 1e9}
        if incremental:
            n_tiles = (S // 32) ** 2 if S % 32 == 0 and (S // 32) ** 2 <= 32 else None
            if n_tiles:
                st = states[0][:, 8].to(torch.int64) & 0xffffffff
                kept = 1.0 - float(sum(((st >> b) & 1).sum() for b in range(n_tiles))) / (st.numel() * n_tiles)
                t = torch.tensor([kept], device=dev)
                if world > 1:
                    h.dist.all_reduce(t)
                kept = float(t.item()) / world
                out[name].update({"incremental": True, "tiles_stored_fraction_last_step": kept,
                                  "bytes_to_learner_per_step": int(nbytes * kept),
                                  "learner_ingest_gbs": nbytes * kept / (ms / K * 1e-3) / 1e9,
                                  "note": "lossless: tiles that were background in the learner's buffer and still are "
                                          "are not stored again (OcclOutputs.obs_tile_state)"})
        del engs, lgs, bufs
        torch.cuda.empty_cache()

    def step_only():
        e = OcclusionEngine(sc, N, RasterConfig(image_size=S), device=dev)
        e.reset(radius=4.0, azimuth=az, elevation=el)

        def s(i):
            e.step(actions_dev[i % 8])

        for i in range(warm):
            s(i)
        ms = h.timed(s, K)
        out["step_only"] = {"value": world * N * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K}

    def features_leg():
        # row N-1: the policy's frozen encoder runs on the env ranks; 256 floats per env travel instead of the image
        from occlusionenv_b200.dist import FeatureGather
        from occlusionenv_b200.features import FrozenEncoder, random_state_dict
        enc = FrozenEncoder(random_state_dict(8), device=dev, chunk=1024)
        e = OcclusionEngine(sc, N, RasterConfig(image_size=S), device=dev)
        e.reset(radius=4.0, azimuth=az, elevation=el)
        fg = FeatureGather(N, enc.out_features, dev, dst=0)
        feats = torch.empty(N, enc.out_features, dtype=torch.float32, device=dev)

        def s(i):
            e.step(actions_dev[i % 8])
            enc(e.obs, out=feats)
            fg.gather(feats, e.reward, e.done)

        for i in range(warm):
            s(i)
        ms = h.timed(s, K)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            enc(e.obs, out=feats)
        ev1.record()
        torch.cuda.synchronize()
        out["features"] = {"value": world * N * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
                           "encoder_ms_per_step": ev0.elapsed_time(ev1) / 3, "bytes_to_learner_per_step": (world - 1) * N * (enc.out_features * 4 + 5),
                           "encoder": "FullNetwork(8, dilation=2, separable=True).encoder + global average pool, random weights "
                                      "(the reference's checkpoint is not in its tree), torch/cuDNN convolutions, "
                                      f"fp32 storage, allow_tf32={bool(torch.backends.cudnn.allow_tf32)}"}

    only = set(x for x in args.gather_legs.split(",") if x)
    want = lambda name: not only or name in only
    if want("step_only"):
        step_only()
    for name, planes, halves in (("serial", 4, 1), ("double_buffered", 4, 2), ("double_buffered_gray", 2, 2)):
        if want(name):
            leg(name, planes, "nccl", halves)
    for name, planes, inc in (("p2p_fused", 4, False), ("p2p_fused_gray", 2, False), ("p2p_fused_incremental", 4, True),
                              ("p2p_fused_gray_incremental", 2, True)):
        if not want(name):
            continue
        try:
            leg(name, planes, "p2p", 1, incremental=inc)
        except Exception as ex:  # CUDA IPC unavailable (e.g. a container without peer access): reported, not hidden
            out[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    if want("features"):
        try:
            features_leg()
        except Exception as ex:
            out["features"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    best = max((v["value"] for k, v in out.items() if isinstance(v, dict) and "value" in v and k != "step_only"), default=None)
    out["value"], out["unit"] = best, UNIT
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--occluder", default="box", choices=["box", "teapot"])
    ap.add_argument("--grad", action="store_true", help="differentiable step (config 4)")
    ap.add_argument("--no-gather", action="store_true", help="world > 1: skip the learner-boundary legs (config 5)")
    ap.add_argument("--gather-envs", type=int, default=8192, help="envs per GPU of the config-5 legs")
    ap.add_argument("--gather-legs", default="", help="comma-separated subset of the config-5 legs (default: all)")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip config_results (config 4 and config 3)")
    ap.add_argument("--c3-envs", type=int, default=8192)
    ap.add_argument("--c3-steps", type=int, default=20)
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
