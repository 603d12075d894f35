#!/bin/bash
# C3-only A/B of library variants under build/variants/
for f in build/variants/*.so; do
  export OCCL_B200_LIB=$PWD/$f
  c=$(python tools/c3_probe.py 512 256 2>&1 | tail -1 | sed 's/.*-> //; s/;.*//')
  echo "$f c3=$c"
done
