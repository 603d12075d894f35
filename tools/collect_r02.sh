#!/bin/bash
# One GPU call at HEAD: parity tests, the bench lines, the ncu launch list and full captures of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python bench.py --steps 30 --warmup 5 --occluder teapot --no-configs --no-cpu-baseline > gpurun_out/bench_r02_teapot.json 2>/dev/null
python bench.py --steps 30 --warmup 5 --occluder teapot --grad --no-cpu-baseline > gpurun_out/bench_r02_teapot_grad.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null
python tools/single_env_probe.py > gpurun_out/single_env.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_l.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs"
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -f -o gpurun_out/raster_r02 $B2 > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:face_setup -s 6 -c 1 -f -o gpurun_out/setup_r02 $B2 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/*_r02*.ncu-rep
