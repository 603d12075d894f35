#!/bin/bash
# A/B of library variants under build/variants/ on C2 (box), the reference scene (teapot), C4 (grad) and the C3 probe
for f in build/variants/*.so; do
  export OCCL_B200_LIB=$PWD/$f
  a=$(python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["roofline"]["kernel_ms"],3))')
  b=$(python bench.py --steps 40 --warmup 5 --no-cpu-baseline --occluder teapot 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  g=$(python bench.py --steps 40 --warmup 5 --no-cpu-baseline --grad 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  c=$(python tools/c3_probe.py 512 256 2>&1 | tail -1 | sed 's/.*-> //; s/;.*//')
  echo "$f box=$a teapot=$b grad=$g c3=$c"
done
