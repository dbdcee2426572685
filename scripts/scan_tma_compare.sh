#!/bin/bash
# GPU box: root-level scan of the bench workload (105 MB) and of the batch repeated 8x (840 MB), plain 128-bit loads
# against the bulk-async-copy (TMA) variant of scan_kernel (MPRG_SCAN_TMA=1); CUDA events around the launch, L2 flushed.
for n in 1000 8000; do
  for t in 0 1; do
    if [ $t = 1 ]; then export MPRG_SCAN_TMA=1; else unset MPRG_SCAN_TMA; fi
    echo "== n=$n TMA=$t  $(python scripts/scan_only.py $n 12 2>&1 | tail -1)"
  done
done
unset MPRG_SCAN_TMA
