"""GPU box, under ncu: one warm-up build, then cudaProfilerStart and ONE build of the chosen workload on a
single worker (every kernel launch of a level covers the whole batch).  Used by scripts/ncu_kernels.sh.

    python scripts/build_once.py bench      # BASELINE config #2, 1,000 loci
    python scripts/build_once.py deep       # BASELINE config #4, one 10,000 x 20,000 locus (-N 10)"""
import os, sys
from pathlib import Path
os.environ.setdefault("MPRG_WORKERS", "1")
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np, torch
import bench
from make_prg_b200 import device, synth
what = sys.argv[1] if len(sys.argv) > 1 else "bench"
ctx = device.Context(0)
if what == "deep":
    mats, N = [synth.config_msa(4, 0)], 10
else:
    data = bench.workload(0, 1000)
    mats, N = (data.reshape(-1), [(bench.ROWS, bench.COLS)] * 1000), 5
batch = ctx.upload(mats)
res = ctx.build(batch, N, 7); res.free()
torch.cuda.synchronize()
torch.cuda.profiler.start()
res = ctx.build(batch, N, 7)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", res.status(0), len(res.prg(0)))
