#!/bin/bash
# second GPU call of the round-2 collection (gpurun merges at most 64 MiB per call): the differentiable and config-3 captures
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grad > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -f -o gpurun_out/raster_grad_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grad > gpurun_out/ncu_g.log 2>&1
python tools/c3_probe.py 64 256 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -f -o gpurun_out/raster_c3_r02 python tools/c3_probe.py 64 256 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*_r02*.ncu-rep
