"""Instruction / stall-sample shares per region of raster_kernel (regions delimited by source markers).
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
import os
lines = open(os.environ.get('OCCL_SRC', '/root/repo/occlusionenv_b200/csrc/occl_b200.cu')).read().split('\n')


def find(pat):
    for i, l in enumerate(lines):
        if pat in l:
            return i + 1
    raise SystemExit("marker not found: " + pat)


marks = [
    ("exact helpers (eval_pair, bary, seg)", "__device__ __forceinline__ float seg_dist", "// raw SFU approximations"),
    ("div_rn_hoisted / sfu", "// raw SFU approximations", "// clipped-barycentric depth"),
    ("soft_accumulate / soft_term", "// soft accumulator of a (pixel, object) slot", "__device__ __forceinline__ unsigned long long pack2f"),
    ("raster_pair (pair evaluation)", "// Round record of a face", "// Tangent terms of one soft hit"),
    ("tile prologue (mask, init)", "__device__ __forceinline__ void raster_tile(", "// ---- main phase: rounds of up to R faces"),
    ("cut faces (z-clip)", "// z-clip: pytorch3d renderer/mesh/clip.py::clip_faces", "// kernel 2: per-env face setup"),
    ("round fill + scan", "// ---- main phase: rounds of up to R faces", "// (b) this warp's share of the round's pairs"),
    ("pair loop (advance, dispatch, depth queue)", "// (b) this warp's share of the round's pairs", "// ---- cut faces (z-clip) ---"),
    ("top-K overflow (pre-scan)", "// ---- pixels with more than K hits: keep the K nearest", "// ---- epilogue: blend, shade"),
    ("top-K overflow: warp-per-slot + rounds", "// K-overflow resolution of one tile.", "// ---- pixels with more hits than a round holds"),
    ("top-K overflow: one-pixel fallback", "// ---- pixels with more hits than a round holds", "// TW, TH: compile-time tile shape"),
    ("hit_tangent", "// Tangent terms of one soft hit", "__device__ __forceinline__ double warp_sum"),
    ("epilogue", "// ---- epilogue: blend, shade", "// kernel 6: per-env finalisation"),
]
marks = [(n, find(a), find(b)) for n, a, b in marks]
hdr = cur = None
out = []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
        ie, si, ti_ = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    elif hdr and r and r[0].isdigit() and len(r) > ie:
        try:
            out.append((cur.split("/")[-1], int(r[0]), int(r[ie] or 0), int(r[si] or 0), int(r[ti_] or 0)))
        except ValueError:
            pass
ti, ts = sum(o[2] for o in out), sum(o[3] for o in out)
acc = {}
for f, l, i, s, t in out:
    key = "other: " + f
    if f == "occl_b200.cu":
        key = "other occl_b200.cu"
        for name, a, b in marks:
            if a <= l < b:
                key = name
    v = acc.setdefault(key, [0, 0, 0])
    v[0] += i
    v[1] += s
    v[2] += t
print(f"total warp instructions {ti}, stall samples {ts}")
for k, (i, s, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} inst {i/ti*100:5.1f}%  samples {s/ts*100:5.1f}%  active lanes/inst {t/max(i,1):5.1f}")
