"""SASS instruction count and opcode histogram of the production raster kernels (cuobjdump -sass on libocclb200.so).
usage: python tools/sass_hist.py [lib.so] > profiles/rNN_raster_sass_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "occlusionenv_b200/libocclb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, kernels = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.+?);", line)
    if m and cur is not None:
        ins = m.group(1).strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        kernels[cur].append(ins.split()[0])
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
want = ("raster_kernel<false, 32, 32, false>", "raster_kernel<true, 32, 32, false>", "raster_kernel<false, 128, 4, false>",
        "face_setup_kernel<true>")
for (mangled, ops), name in zip(kernels.items(), names):
    if not any(w in name.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "") for w in want):
        continue
    h = collections.Counter(o.split(".")[0] for o in ops)
    print(f"== {name}: {len(ops)} SASS instructions ({len(ops) * 16 / 1024:.0f} KB)")
    print("   " + "  ".join(f"{k} {v}" for k, v in h.most_common(28)))
    blackwell = {k: v for k, v in h.items() if k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP"))}
    print(f"   tensor-core / TMA opcodes: {blackwell or 'none (this path has no contraction and gathers 64-byte records by index)'}")
    print(f"   shared-memory atomics: ATOMS {h.get('ATOMS', 0)} (of which CAS loops: "
          f"{sum(1 for o in ops if o.startswith('ATOMS.CAST'))}), LDS {h.get('LDS', 0)}, STS {h.get('STS', 0)}, BAR {h.get('BAR', 0)}")
