#!/bin/bash
# differentiable-step bench (config 4) for every library variant under build/variants/ and the in-tree library
for f in build/variants/*.so occlusionenv_b200/libocclb200.so; do
  export OCCL_B200_LIB=$PWD/$f
  a=$(python bench.py --grad --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  b=$(python bench.py --grad --occluder teapot --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"]))')
  echo "$f grad box=$a teapot=$b"
done
