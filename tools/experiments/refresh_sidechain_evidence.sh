set -x
mkdir -p gpurun_out/summary2
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python tools/bench_sidechain.py > gpurun_out/r02_sidechain_table.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:sidechain -s 4 -c 2 -o gpurun_out/r02_sidechain python tools/experiments/profile_sidechain.py > gpurun_out/r02_sidechain.log 2>&1
mkdir -p /tmp/only && cp gpurun_out/r02_sidechain.ncu-rep /tmp/only/
python - <<'PY'
import sys, pathlib
sys.path.insert(0, 'tools')
import profile_summarise as P
P.GP = pathlib.Path('/tmp/only')
sys.argv = ['x', 'r02', 'gpurun_out/summary2']
P.main()
PY
rm -f gpurun_out/*.ncu-rep
ls gpurun_out/summary2
