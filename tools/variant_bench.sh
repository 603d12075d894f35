#!/bin/bash
# times every library variant under build/variants/ on the C2 workloads (+ the config-3 probe)
for f in build/variants/*.so; do
  export OCCL_B200_LIB=$PWD/$f
  a=$(python bench.py --steps 20 --warmup 5 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  b=$(python bench.py --steps 20 --warmup 5 --occluder teapot 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  c=$(python tools/c3_probe.py 256 256 2>&1 | tail -1 | sed 's/.*-> //; s/;.*//')
  echo "$f box=$a teapot=$b c3=$c"
done
