#!/bin/bash
# config-3 probe (and the config-2 box bench) for every library variant under build/variants/
for f in build/variants/*.so; do
  export OCCL_B200_LIB=$PWD/$f
  c=$(python tools/c3_probe.py 512 256 2>&1 | tail -1 | sed 's/.*-> //; s/;.*//')
  a=$(python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  echo "$f c3=$c box=$a"
done
