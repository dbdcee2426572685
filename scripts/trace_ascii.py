"""GPU box: per-range timeline of mprg_build_ascii on the bench workload (MPRG_TRACE=1 MPRG_TRACE_ALL=1)."""
import os, sys, time
from pathlib import Path
os.environ["MPRG_TRACE"] = "1"; os.environ["MPRG_TRACE_ALL"] = "1"
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np, torch
import bench
from make_prg_b200 import device
n = 1000
data = bench.workload(0, n)
host = torch.from_numpy(data.reshape(-1)).pin_memory().numpy()
ctx = device.Context(0)
shapes = [(bench.ROWS, bench.COLS)] * n
for it in range(4):
    print(f"==== call {it}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    b, res = ctx.build_ascii((host, shapes), 5, 7)
    print(f"==== call {it}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr, flush=True)
    res.free(); b.free()
