"""Peer-store bandwidth against the store segment size (needs 2 GPUs): device 0 writes into device 1's HBM.
usage: python tools/p2p_store_probe.py"""
import ctypes, os, subprocess, sys
import torch
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, root)
so = os.path.join(root, "build", "p2p_store_probe.so")
if not os.path.exists(so):
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-Xcompiler", "-fPIC", "-shared", "-o", so,
                           os.path.join(root, "tools", "p2p_store_probe.cu")])
lib = ctypes.CDLL(so)
lib.probe_launch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
n = 1 << 28  # 1 GiB of floats
for target in ((0, 1) if torch.cuda.device_count() > 1 else (0,)):
    dst = torch.empty(n, dtype=torch.float32, device=f"cuda:{target}")
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda:0")
    if target:
        from occlusionenv_b200 import _lib as L
        L.check(L.load().occl_enable_peer_access(target), "occl_enable_peer_access")   # kernels of device 0 may store to device 1
    st = torch.cuda.current_stream().cuda_stream
    for vec, pitch, rows, tag in ((1, 128, 32, "128 B segments, 512 B pitch (32-px tile rows)"), (2, 128, 32, "256 B segments, 512 B pitch"),
                                  (4, 128, 32, "512 B segments = full rows"), (1, 32, 1, "128 B segments, contiguous stream"),
                                  (4, 128, 1, "512 B segments, contiguous stream")):
        for _ in range(2):
            lib.probe_launch(dst.data_ptr(), n, vec, pitch, rows, 148 * 8, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rc = lib.probe_launch(dst.data_ptr(), n, vec, pitch, rows, 148 * 8, st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0 and float(dst[12345]) == 1.0
        print(f"device 0 -> device {target}: {tag:48s} {5 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9:7.1f} GB/s", flush=True)

# incast: every other GPU stores into device 0 at the same time (the learner's situation in config 5)
nd = torch.cuda.device_count()
if nd > 2:
    import time
    from occlusionenv_b200 import _lib as L
    lib2 = L.load()
    dst = torch.empty(n, dtype=torch.float32, device="cuda:0")
    part = n // (nd - 1) // 4096 * 4096
    for d in range(1, nd):
        torch.cuda.set_device(d)
        torch.zeros(1, device=f"cuda:{d}")
        L.check(lib2.occl_enable_peer_access(0), "occl_enable_peer_access")
    for vec, pitch, rows, tag in ((1, 128, 32, "128 B segments, 512 B pitch"), (4, 128, 32, "512 B segments")):
        for rep in range(2):   # first pass warms up
            for d in range(1, nd):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for _ in range(5):
                for d in range(1, nd):
                    torch.cuda.set_device(d)
                    lib.probe_launch(dst.data_ptr() + (d - 1) * part * 4, part, vec, pitch, rows, 148 * 8, torch.cuda.current_stream(d).cuda_stream)
            for d in range(1, nd):
                torch.cuda.synchronize(d)
            dt = time.perf_counter() - t0
        print(f"incast {nd - 1} GPUs -> device 0: {tag:40s} {5 * (nd - 1) * part * 4 / dt / 1e9:7.1f} GB/s into device 0", flush=True)
