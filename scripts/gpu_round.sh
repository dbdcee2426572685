#!/bin/bash
# One GPU-box visit: parity tests, scan timing, ncu capture + launch list, bench line.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
for n in 1000 8000; do
  echo "== n=$n" | tee -a gpurun_out/sweep.log
  python scripts/scan_only.py $n 12 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
done
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/scan_$TAG -f python scripts/scan_only.py 1000 6 > gpurun_out/ncu_scan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-big > gpurun_out/ncu_bench.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_$TAG.log
python scripts/run_configs.py 3,5,4r,4 2>&1 | tail -10 > gpurun_out/configs_$TAG.jsonl
python scripts/ncu_summary.py gpurun_out/scan_$TAG.ncu-rep > gpurun_out/scan_$TAG.txt 2>&1
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
cut -c1-300 gpurun_out/configs_$TAG.jsonl
