"""Per-source-line SASS instruction counts inside a line range of occl_b200.cu, normalised by the
execution count of a marker line (per loop iteration).
usage: python tools/ncu_loop.py report.ncu-rep "<start pattern>" "<end pattern>" "<marker pattern>" """
import collections
import csv
import io
import subprocess
import sys

rep, p_start, p_end, p_mark = sys.argv[1:5]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
lines = open('/root/repo/occlusionenv_b200/csrc/occl_b200.cu').read().split('\n')


def find(pat, lo=0, hi=10**9):
    for i, l in enumerate(lines):
        if pat in l and lo <= i + 1 < hi:
            return i + 1
    raise SystemExit("pattern not found: " + pat)


a, b = find(p_start), find(p_end)
mark = find(p_mark, a, b)
hdr = curfile = curline = None
per_line = collections.OrderedDict()
for r in rows:
    if r and r[0] == "File Path":
        curfile = r[1]
    elif r and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
    elif hdr and r:
        if r[0].isdigit():
            curline = (curfile.split('/')[-1], int(r[0]))
        elif r[0] == "" and len(r) > ie and curline and curline[0] == "occl_b200.cu" and a <= curline[1] < b:
            toks = r[3].strip().split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
            try:
                n = int(r[ie])
            except ValueError:
                continue
            per_line.setdefault(curline[1], collections.Counter())[op.split('.')[0]] += n
it = max(per_line[mark].values())
print("marker executions (warp level):", it)
tot = 0
for l, c in sorted(per_line.items()):
    s = sum(c.values())
    tot += s
    if s / it > 1.0:
        print(f"{l:5d} {s/it:6.1f}/it {dict(c.most_common(5))} | {lines[l-1].strip()[:64]}")
print("total per marker execution", tot / it)
