#!/bin/bash
# One GPU call: parity tests, the bench lines, the ncu launch list and one full capture of the raster kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_box.json 2> gpurun_out/bench_box.err
python bench.py --steps 30 --warmup 5 --occluder teapot > gpurun_out/bench_teapot.json 2>/dev/null
python bench.py --steps 30 --warmup 5 --grad > gpurun_out/bench_box_grad.json 2>/dev/null
python bench.py --steps 30 --warmup 5 --grad --occluder teapot > gpurun_out/bench_teapot_grad.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python tools/c3_probe.py 512 256 > gpurun_out/c3.log 2>&1
python tools/misc_probe.py > gpurun_out/misc.log 2>&1
python tools/single_env_probe.py > gpurun_out/single_env.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -f -o gpurun_out/raster_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:face_setup -s 6 -c 1 -f -o gpurun_out/setup_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -f -o gpurun_out/raster_c3 \
    python tools/c3_probe.py 64 256 > gpurun_out/ncu_c3.log 2>&1
cut -c1-200 gpurun_out/bench_box.json gpurun_out/bench_teapot.json gpurun_out/bench_box_grad.json gpurun_out/bench_teapot_grad.json gpurun_out/bench_reference.json
tail -1 gpurun_out/c3.log
