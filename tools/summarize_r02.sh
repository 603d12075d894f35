#!/bin/bash
# Turns the reports / logs a collection run left in gpurun_out/ into the tracked summaries under profiles/ (run at HEAD).
cd "$(dirname "$0")/.."
P=profiles
hdr() { echo "# $1"; echo "# captured at $(git rev-parse --short HEAD) with: $2"; echo; }
for k in raster setup raster_grad raster_c3; do
  rep=gpurun_out/${k}_r02.ncu-rep
  [ -f $rep ] || continue
  case $k in
    raster) cmd="ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs   (C2: teapot + box, 4096 envs, 128^2, forward)";;
    setup) cmd="ncu --set full ... -k regex:face_setup -s 6 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs";;
    raster_grad) cmd="ncu --set full ... -k regex:raster_kernel -s 6 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grad   (C4)";;
    raster_c3) cmd="ncu --set full ... -k regex:raster_kernel -s 3 -c 1 python tools/c3_probe.py 64 256   (config-3 meshes, 64 envs, 256^2, 128x4 tile)";;
  esac
  { hdr "ncu summary of ${k}" "$cmd"; python tools/ncu_lines.py $rep 40; echo; echo "== instruction / stall-sample shares per region (tools/ncu_regions.py)"; python tools/ncu_regions.py $rep; } > $P/r02_${k}_ncu_summary.txt 2>&1
done
if [ -f gpurun_out/launches_r02.csv ]; then
  cp gpurun_out/launches_r02.csv $P/r02_launches_ncu.csv
  { hdr "launch list" "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs"; python tools/launch_summary.py gpurun_out/launches_r02.csv; } > $P/r02_launches_summary.txt 2>&1
fi
python - <<'PY'
import csv, io, json, subprocess, os
rep = "gpurun_out/raster_r02.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    def val(name):
        i = h.index(name); x = float(v[i]); unit = u[i].lower()
        return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    json.dump({"envs": 4096, "size": 128, "occluder": "box", "grad": False, "dram_bytes_per_launch": int(b),
               "source": "profiles/r02_raster_ncu_summary.txt (ncu --set full, raster_kernel<0,32,32,0>, one launch)"},
              open("profiles/raster_ncu_traffic.json", "w"))
    print("traffic", b)
    # hot loop SASS with executed counts
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; ie = hdr.index("Instructions Executed"); it = hdr.index("Avg. Threads Executed"); isamp = hdr.index("# Samples")
    data = [(r[1].strip(), int(r[ie] or 0), r[it], int(r[isamp] or 0)) for r in rows[2:] if len(r) > ie]
    mx = max(d[1] for d in data)
    import collections
    cnts = collections.Counter(d[1] for d in data if d[1] > mx * 0.3)
    plateau = cnts.most_common(1)[0][0]
    idx = [i for i, d in enumerate(data) if d[1] >= plateau * 0.55]
    lo, hi = min(idx), max(idx)
    ops = collections.Counter()
    for d in data[lo:hi + 1]:
        op = d[0].split()[1] if d[0].startswith("@") else d[0].split()[0]
        ops[op.split(".")[0]] += d[1]
    with open("profiles/r02_raster_sass_hot_loop.txt", "w") as f:
        f.write("# pair loop of raster_kernel<0,32,32,0> (C2): SASS with executed warp-instruction counts (M), stall samples, avg active threads\n")
        f.write(f"# one trip = {plateau} executions per launch; instructions per trip by opcode: " +
                ", ".join(f"{k} {v / plateau:.1f}" for k, v in ops.most_common(24)) + f"; total {sum(ops.values()) / plateau:.0f}\n\n")
        for d in data[lo:hi + 1]:
            f.write(f"{d[1] / 1e6:7.2f}M s{d[3]:5d} t{d[2]:>5s}  {d[0][:100]}\n")
PY
python tools/sass_hist.py > $P/r02_raster_sass_histogram.txt 2>&1
for f in bench_r02 bench_r02_teapot bench_r02_teapot_grad bench_r02_reference; do [ -s gpurun_out/$f.json ] && cp gpurun_out/$f.json $P/r02_$f.json; done
[ -s gpurun_out/single_env.log ] && cp gpurun_out/single_env.log $P/r02_single_env_latency.txt
[ -s gpurun_out/bench_2gpu.json ] && cp gpurun_out/bench_2gpu.json $P/r02_bench_2gpu.json
[ -s gpurun_out/bench_8gpu.json ] && cp gpurun_out/bench_8gpu.json $P/r02_bench_8gpu.json
ls -la $P | grep r02
