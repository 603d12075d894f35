#!/bin/bash
# usage: tools/sass_kernel.sh <lib.so> <mangled-name-substring>  -> SASS of the first kernel whose name contains it
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on {print}'
