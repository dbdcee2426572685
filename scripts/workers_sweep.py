"""GPU box: mprg_build wall time on the bench workload vs. number of worker threads."""
import os, sys, time
from pathlib import Path
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
print("cpu cores", os.cpu_count())
ref = None
for W in (1, 2, 4, 8, 12, 16):
    ctx.set_workers(W)
    ts = []; tu = []
    for it in range(6):
        t0 = time.perf_counter()
        batch = ctx.upload((host, shapes))
        t1 = time.perf_counter()
        res = ctx.build(batch, 5, 7)
        t2 = time.perf_counter()
        if it == 0:
            prgs = [res.prg(i) for i in range(n)]
            if ref is None: ref = prgs
            assert prgs == ref
        res.free(); batch.free()
        tu.append(t1 - t0); ts.append(t2 - t1)
    print(f"workers {W}: build {1e3*min(ts[2:]):.2f} ms (median {1e3*sorted(ts[2:])[2]:.2f}) upload {1e3*min(tu[2:]):.2f} ms -> {n/min(ts[2:]):.0f} loci/s resident", flush=True)
