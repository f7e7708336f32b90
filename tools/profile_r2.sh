#!/bin/bash
# ncu evidence for round 2 (run under gpurun on one B200):  bash tools/profile_r2.sh
# The captures are summarised on the box (gpurun_out/summary/r02_*); copy those into profiles/ here.
set -x
mkdir -p gpurun_out
# (1) launch list of the bench command: every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
# (2) full capture of the dominant kernel (one launch)
ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 1 -c 1 -o gpurun_out/r02_pair_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_pair_tile.log 2>&1
# (3) full capture of the back-mapping kernels: forward (lane-per-frame fwd6 at 262144 frames), backward with and
#     without bond-angle gradients
ncu --set full --clock-control none --import-source on -k regex:backmap_fwd6 -s 1 -c 1 -o gpurun_out/r02_backmap_fwd6 \
    python tools/profile_backmap_fwd.py 6 262144 1500 > gpurun_out/r02_backmap_fwd6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backmap_ -c 4 -o gpurun_out/r02_backmap \
    python tools/run_backmap_once.py > gpurun_out/r02_backmap.log 2>&1
# (4) the fused Cartesian branch and the PairwiseDistances kernels at the ADC training shape
ncu --set full --clock-control none --import-source on -k regex:"cart_|pairwise_" -c 8 -o gpurun_out/r02_cart \
    python tools/train_profile.py adc_fused > gpurun_out/r02_cart.log 2>&1
# (5) PairwiseDistances forward / backward at 65 536 x 100 atoms, and the narrow-input / small-tile cost kernels at training sizes
ncu --set full --clock-control none --import-source on -k regex:"pairwise_flat3" -s 2 -c 2 -o gpurun_out/r02_pairwise \
    python tools/profile_pairwise.py > gpurun_out/r02_pairwise.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"small_cost_kernel|Geom<64" -c 4 -o gpurun_out/r02_small_cost \
    python tools/bench_small_cost.py > gpurun_out/r02_small_cost.log 2>&1
# (6) back-mapping with side chains at the training batch (256 frames x 448 atoms)
ncu --set full --clock-control none --import-source on -k regex:sidechain -s 4 -c 2 -o gpurun_out/r02_sidechain \
    python tools/experiments/profile_sidechain.py > gpurun_out/r02_sidechain.log 2>&1
# summarise on the box and drop the captures (gpurun_out/ is limited to 64 MiB)
python tools/profile_summarise.py r02 gpurun_out/summary > gpurun_out/r02_summarise.log 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out gpurun_out/summary
