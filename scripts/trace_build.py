"""GPU box: wall-clock phase trace of mprg_build on the bench workload (MPRG_TRACE=1)."""
import os, sys, time
from pathlib import Path
os.environ["MPRG_TRACE"] = "1"
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np
import bench
from make_prg_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
data = bench.workload(0, n)
ctx = device.Context(0)
shapes = [(bench.ROWS, bench.COLS)] * n
for it in range(3):
    t0 = time.perf_counter()
    batch = ctx.upload((data.reshape(-1), shapes))
    t1 = time.perf_counter()
    res = ctx.build(batch, 5, 7)
    t2 = time.perf_counter()
    prgs = [res.prg(i) for i in range(n)]
    t3 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.1f} ms build {1e3*(t2-t1):.1f} ms fetch {1e3*(t3-t2):.1f} ms", file=sys.stderr)
    res.free(); batch.free()
