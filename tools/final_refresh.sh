#!/bin/bash
# Everything the round-end evidence under profiles/ is made of, in one gpurun call (one B200):  bash tools/final_refresh.sh
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputest.txt
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.txt 2>&1
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
{ python tools/bench_small_cost.py; python tools/bench_small_ops.py; python tools/bench_backmap.py; python tools/bench_generation.py;
  python tools/bench_sidechain.py; } > gpurun_out/r02_kernel_table.txt 2> gpurun_out/r02_kernel_table.err
{ echo "# compute-sanitizer on tools/sanitize_round2.py (round-2 kernels), B200"; echo "## memcheck";
  compute-sanitizer --tool memcheck python tools/sanitize_round2.py 2>&1 | tail -4; echo "## racecheck";
  compute-sanitizer --tool racecheck python tools/sanitize_round2.py 2>&1 | tail -4; } > gpurun_out/r02_sanitizer.txt
bash tools/profile_r2.sh > gpurun_out/r02_profile.log 2>&1
tail -3 gpurun_out/r02_gputest.txt gpurun_out/r02_smoke.txt gpurun_out/r02_sanitizer.txt
head -c 600 gpurun_out/r02_bench_1gpu.json
