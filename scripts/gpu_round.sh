#!/bin/bash
# One GPU-box visit: parity tests, scan-kernel tile-height sweep, ncu capture of the scan kernel, bench line.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
for ti in 0 16 20 24 32 48; do
  echo "== MPRG_TILE_ITERS=$ti" | tee -a gpurun_out/sweep.log
  if [ "$ti" = 0 ]; then python scripts/scan_only.py 1000 12 2>&1 | tail -2 | tee -a gpurun_out/sweep.log
  else MPRG_TILE_ITERS=$ti python scripts/scan_only.py 1000 12 2>&1 | tail -2 | tee -a gpurun_out/sweep.log; fi
done
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/scan_r1_v8 -f python scripts/scan_only.py 1000 6 > gpurun_out/ncu_scan.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
cat gpurun_out/bench7.json
