#!/bin/bash
# ncu evidence for round 1 (run under gpurun on one B200):  bash tools/profile_r1.sh
# Afterwards, here:  python tools/profile_summarise.py   (writes profiles/r01_*)
set -x
mkdir -p gpurun_out
# (1) launch list of the bench command: every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_launches_bench.log 2>&1
# (2) full capture of the dominant kernel (one launch)
ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 1 -c 1 -o gpurun_out/r01_pair_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r01_pair_tile.log 2>&1
# (3) full capture of the back-mapping kernels: forward, backward with and without bond-angle gradients
ncu --set full --clock-control none --import-source on -k regex:backmap_ -c 4 -o gpurun_out/r01_backmap \
    python tools/run_backmap_once.py > gpurun_out/r01_backmap.log 2>&1
ls -la gpurun_out
