#!/bin/bash
# registers / spills / stack of every kernel of libocclb200.so (nvcc -Xptxas -v), one line per kernel
cd "$(dirname "$0")/.."
python occlusionenv_b200/build.py 2>&1 | awk '
/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)}
/bytes stack frame/ && name!="" {spill=$0}
/Used [0-9]+ registers/ && name!="" {print name " | " $0 " |" spill; name=""}' | sed 's/ptxas info    : //' | c++filt | cut -c1-230
